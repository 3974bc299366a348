#!/usr/bin/env python3
"""Regenerates tests/golden/oracle_vectors.npz: seeded inputs and the oracle's outputs for the operations on the path
(3-hop ANY_PAIR chain on RMAT-8, masked mxm in its RSC form, eWiseAdd / eWiseMult, transpose, delta_lmxm, BFS).
The reference itself cannot run here (see reference_known_answers.json), so these are REGRESSION vectors of the oracle --
they freeze its behaviour at the point where it was checked against the reference's known answers, scipy and the algebraic
identities (tests/test_oracle.py, tests/test_golden.py).  Run from the repo root: python tests/golden/make_oracle_vectors.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle as orc  # noqa: E402


def pack(prefix, m, out):
    out[prefix + "_shape"] = np.array([m.nrows, m.ncols], np.int64)
    out[prefix + "_p"] = m.p
    out[prefix + "_j"] = m.j
    if m.x is not None:
        out[prefix + "_x"] = m.x


def main():
    out = {}
    A = orc.rmat_csr(8, 8, 123)
    n = A.nrows
    rng = np.random.default_rng(2024)
    src = rng.choice(n, size=40, replace=False)
    F = orc.build_matrix(40, n, np.arange(40), src)
    pack("A", A, out)
    out["sources"] = src.astype(np.int64)
    W = F
    for h in range(3):
        W = orc.mxm(W, A)
        pack(f"chain{h + 1}", W, out)
    M = orc.build_matrix(40, n, rng.integers(0, 40, 600), rng.integers(0, n, 600))
    pack("mask", M, out)
    pack("chain1_rsc", orc.mxm(F, A, M, mask_mode=2), out)
    B = orc.build_matrix(n, n, rng.integers(0, n, 500), rng.integers(0, n, 500))
    pack("B", B, out)
    pack("A_union_B", orc.ewise_add(A, B), out)
    pack("A_inter_B", orc.ewise_mult(A, B), out)
    pack("A_T", orc.transpose(A), out)
    rows, cols, _ = A.tuples()
    dels = rng.choice(A.nnz, 60, replace=False)
    dm = orc.build_matrix(n, n, rows[dels], cols[dels])
    dp = orc.build_matrix(n, n, rng.integers(0, n, 60), rng.integers(0, n, 60))
    pack("dm", dm, out)
    pack("dp", dp, out)
    pack("delta_lmxm", orc.delta_lmxm(F, A, dp, dm), out)
    lvl, par = orc.bfs(A, int(src[0]))
    out["bfs_src"] = np.array([int(src[0])], np.int64)
    out["bfs_level"] = lvl
    out["bfs_parent"] = par
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith("_p")})


if __name__ == "__main__":
    main()
