"""CPU-side checks of the drop-in boundary: the in-tree library loads, exports every symbol
include/dab_b200.h declares, and fails loudly (no CPU fallback) without a CUDA device."""
import ctypes
import os
import re
import shutil
import subprocess

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


C_PROGRAM = r"""
#include <stdio.h>
#include <string.h>
#include "dab_b200.h"

int main(void) {
  struct dab_cycle_config cfg;
  dab_handle* h = (dab_handle*)0;
  int rc;
  memset(&cfg, 0, sizeof cfg);
  printf("version %d sizeof_cfg %d max_ens %d max_ranks %d ipc %d\n", dab_version(), (int)sizeof cfg,
         DAB_MAX_ENSEMBLE, DAB_MAX_RANKS, DAB_IPC_HANDLE_BYTES);
  rc = dab_create(&h, 0);
  if (rc != 0) {                       /* no GPU here: must fail loudly and leave no handle */
    printf("create rc %d handle %s msg %s\n", rc, h ? "set" : "null", dab_last_error((dab_handle*)0));
    return 0;
  }
  printf("create ok\n");
  dab_set_pdl(1);
  rc = dab_cycler_setup(h, &cfg, 0, 0, 0);       /* system_dim 0: argument error, not a crash */
  printf("setup rc %d msg %s\n", rc, dab_last_error(h));
  dab_destroy(h);
  return 0;
}
"""


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_header_is_plain_c_and_links_from_c(lib, tmp_path):
    """The boundary is a C ABI: include/dab_b200.h compiles as strict C99 (no C++, no torch types), and a
    C program linked against the in-tree library can call it; without a GPU dab_create returns an error
    code and a message instead of a handle."""
    inc = os.path.join(ROOT, "include")
    src = tmp_path / "use_dab.c"
    src.write_text(C_PROGRAM)
    exe = tmp_path / "use_dab"
    libdir = os.path.dirname(_build.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe),
                           "-L", libdir, "-ldab_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    first = out.stdout.splitlines()[0].split()
    assert int(first[1]) >= 100 and int(first[3]) == ctypes.sizeof(_native.CycleConfig)
    import torch
    if torch.cuda.is_available():
        assert "create ok" in out.stdout and "setup rc -1" in out.stdout
    else:
        assert "create rc" in out.stdout and "handle null" in out.stdout
        assert "no CPU fallback" in out.stdout or "CUDA" in out.stdout
