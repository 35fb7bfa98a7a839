"""CPU-side checks of the drop-in boundary: the in-tree library loads, exports every symbol
include/dab_b200.h declares, and fails loudly (no CPU fallback) without a CUDA device."""
import ctypes

import numpy as np
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


@pytest.mark.parametrize("n,k,S,ctas,variant", [
    (65536, 64, 20, 148, 0), (4194304, 4, 20, 148, 0), (8192, 64, 20, 148, 1), (65536, 4, 19, 5, 0), (600, 3, 7, 148, 0),
    (36, 10, 20, 148, 0), (65536, 4, 1, 3, 1), (1048576, 2, 32, 16, 0), (40, 20, 20, 7, 1)])
def test_chained_forecast_plan_covers_every_point_once(n, k, S, ctas, variant):
    """The host-side tiling of the chained forecast kernel (csrc/l96_forecast_v6.cu), checked without a GPU: every grid
    point of every member is stored by exactly one tile; a fresh tile keeps 8 S points of halo on its left and 4 S of
    stale points on its right inside its window; a continuing tile starts exactly where its predecessor's window
    advanced to (so that the boundary history it reads is its left neighbour's), in the same chain and member."""
    import ctypes as C
    lib = _native.load()
    dims = (C.c_int * 4)()
    assert lib.dab_debug_l96_chain_plan(n, k, S, ctas, variant, dims, None, 0, None, 0) == 0
    grid, ts, cs, hist_t = list(dims)
    tiles = np.zeros(grid * ts * 4, dtype=np.int32)
    chains = np.zeros(grid * cs, dtype=np.int32)
    assert lib.dab_debug_l96_chain_plan(n, k, S, ctas, variant, dims, tiles.ctypes.data, tiles.size, chains.ctypes.data,
                                        chains.size) == 0
    W = 2048 if variant == 0 else 1024
    A = ((W - 4 * S) // 16) * 16
    assert hist_t == A // 16 - 1 and 1 <= grid <= ctas
    tiles = tiles.reshape(grid, ts, 4)
    chains = chains.reshape(grid, cs)
    cover = np.zeros((k, n), dtype=np.int32)
    for b in range(grid):
        nc = chains[b, 0]
        starts = chains[b, 1:nc + 2]
        assert nc >= 1 and starts[0] == 0 and np.all(np.diff(starts) >= 1) and starts[-1] <= ts
        for c in range(nc):
            prev = None
            for ti in range(starts[c], starts[c + 1]):
                x, g0, p, q = (int(v) for v in tiles[b, ti])
                mem, cont = x >> 1, x & 1
                assert 0 <= mem < k and 0 <= p < q <= n
                if cont:
                    assert prev is not None and ti > starts[c]            # only inside a chain
                    assert prev[0] == mem and g0 == prev[1] + A and p == g0   # window advanced by A, same member
                else:
                    assert ti == starts[c] or prev[0] != mem               # a chain or a member starts fresh
                    assert g0 == p - 8 * S
                assert q <= g0 + W - 4 * S                                 # inside the still-valid part of the window
                assert p % 2 == 0 and q % 2 == 0 and g0 % 2 == 0
                cover[mem, p:q] += 1
                prev = (mem, g0)
    assert cover.min() == 1 and cover.max() == 1
