"""CPU-side checks of the drop-in boundary: the in-tree library loads, exports every symbol
include/dab_b200.h declares, and fails loudly (no CPU fallback) without a CUDA device."""
import ctypes
import os
import re

import pytest

from dataassimbench_b200 import _native, _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _native.load()


def test_header_symbols_all_exported(lib):
    header = open(os.path.join(ROOT, "include", "dab_b200.h")).read()
    declared = set(re.findall(r"DAB_API\s+[\w\s\*]+?\b(dab_\w+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(_native.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_layout_matches_header():
    # 8+4+4+8*4+8+8+4*5 = 84 -> padded to 88 (8-byte alignment)
    assert ctypes.sizeof(_native.CycleConfig) == 88
    assert _native.CycleConfig.n_obs_times.offset == 48
    assert _native.CycleConfig.world.offset == 80


def test_version(lib):
    assert lib.dab_version() >= 100


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(_native.DabError) as ei:
        _native.Handle(0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)
