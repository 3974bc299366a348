"""Bit-exact parity at the sizes BASELINE.json's configs name, against the CPU oracle on the same CSR inputs.
Multi-GB results are compared through an order-sensitive digest computed on both sides (B200_Matrix_digest on the device,
orc_digest on the host: nvals, sum mix(row << 32 | col), sum mix(key + GOLD * CSR position)) plus exact flops; BFS levels and
parents are compared element for element.  The graph is generated on the device and exported once, so both sides read the
same CSR.  Each test skips only when the box lacks the memory it needs (free HBM / host RAM probe), never by default.

These modules sort after the hardware-verified ones on purpose (pytest -x): they are the longest tests, and not yet run on hardware in
the form committed here -- the session that wrote them lost its GPU access when the first version of the config-4 oracle (saxpy form: every
unmasked row materialised, ~1e11 entries on RMAT-24) exhausted a test box's RAM.  The oracle now evaluates masked products in dot
form (oracle/grb_oracle.c: mxm_masked_dot, a few GB), every test states its host-memory need up front (`need`), and
tests/conftest.py runs a resident-set watchdog that kills the test process long before the host is in danger.

This module: the headline chain (512 and 1000 sources, RMAT-24) and config 5 (BFS on RMAT-26) -- the two whose computations
bench.py's own parity legs already verified on a B200.  The row-wise / masked SpGEMM configs (2, 3, 4) are
tests/test_zz3_full_size_spgemm.py, after the new-feature tests (pytest -x stops at the first failure: most certain first)."""
import os
import time

import numpy as np
import pytest

import falkordb_b200 as fb
import oracle as orc
from falkordb_b200.grb import Matrix, Descriptor
from fullsize_util import need, rmat_both, same_digest, rows_of, dev_of, defaults, _GRAPHS

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _defaults():
    yield from defaults()


# ------------------------------------------------------------------------------------------ headline: 3-hop chain, RMAT-24
@pytest.mark.parametrize("nsrc", [512, 1000])
def test_chain_rmat24_all_rows(nsrc):
    """the bench workload, every row: 3-hop F*A*A*A from `nsrc` random non-isolated sources on RMAT-24 (W = 8 / 16 word columns:
    the lane-split pull, long-row segments, the multi-word materialise)"""
    need(40, 60)
    A, Ao = rmat_both(24)
    A.prepare(True)
    deg = np.diff(Ao.p)
    rng = np.random.default_rng(1000003)
    src = rng.choice(np.nonzero(deg > 0)[0], size=nsrc, replace=False).astype(np.uint64)
    _, flops, dg, busy = orc.chain(Ao, src, 3, keep=False)
    F = Matrix(nsrc, Ao.nrows, bool)
    F.build(np.arange(nsrc, dtype=np.uint64), src)
    fl = 0
    for _ in range(3):
        F.lmxm(A)
        fl += fb.get_stat("last_flops")
    assert fb.get_stat("last_path") == 3, "the last hop must take the pull direction at this size"
    F.wait()
    assert fl == flops, "flops (edges traversed) differ from the oracle's"
    same_digest(F, dg, f"3-hop chain, {nsrc} sources, RMAT-24")
    # the bitmap hand-off of the same result: row populations must match the CSR row lengths
    p = np.empty(nsrc + 1, np.uint64)
    fb.check(fb.lib().B200_Matrix_export_CSR(F.h, p.ctypes.data, None, None, 0))
    G = Matrix(nsrc, Ao.nrows, bool)
    G.build(np.arange(nsrc, dtype=np.uint64), src)
    for _ in range(3):
        G.lmxm(A)
    wpr = (Ao.nrows + 63) // 64
    bm = np.empty((nsrc, wpr), np.uint64)
    G.export_bitmap(bm)
    assert np.array_equal(np.bitwise_count(bm).sum(axis=1).astype(np.int64), np.diff(p.astype(np.int64)))


# ------------------------------------------------------------------------------------------ config 5: BFS, RMAT-26
def test_config5_bfs_rmat26_levels_and_parents():
    """BASELINE config 5's graph on one GPU: BFS level (bit-exact) and min-id parent vectors from 3 random sources with out-edges
    on RMAT scale-26 (n = 67.1 M, ~1.05e9 edges) against the oracle; algo.BFS's conventions (algo_procedures.rs:1098-1148)."""
    need(90, 40)
    A, Ao = rmat_both(26)
    A.prepare(True)               # the transpose mirror: B200_bfs then runs the direction-optimising engine (bfs_do.cu)
    deg = np.diff(Ao.p)
    rng = np.random.default_rng(3)
    srcs = rng.choice(np.nonzero(deg > 0)[0], size=3, replace=False)
    for s in srcs:
        lvl, par, edges = fb.bfs(A, int(s))
        wl, wp = orc.bfs(Ao, int(s))
        assert np.array_equal(lvl, wl), f"levels differ from the oracle (source {s})"
        assert np.array_equal(par, wp), f"min-id parents differ from the oracle (source {s})"
        assert edges == int(deg[wl >= 0].sum()), "edges traversed (Graph500 convention) differ"
    _GRAPHS.clear()


