"""Golden fixtures (tests/golden/): the reference's own known answers as data, checked against the ORACLE here on the CPU
(the CUDA path is checked against the same statements in test_gpu_parity.py / test_host_*.py / test_abi.py), and the frozen
oracle vectors that guard the oracle itself against regressions, plus independent scipy re-derivations of those vectors."""
import json
import os

import numpy as np
import scipy.sparse as sp

import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "reference_known_answers.json")))
VEC = np.load(os.path.join(HERE, "golden", "oracle_vectors.npz"))


def unpack(prefix):
    nr, nc = VEC[prefix + "_shape"]
    x = VEC[prefix + "_x"] if prefix + "_x" in VEC.files else None
    return orc.CSR(int(nr), int(nc), VEC[prefix + "_p"], VEC[prefix + "_j"], x)


def same(a, b):
    return a.nrows == b.nrows and a.ncols == b.ncols and np.array_equal(a.p, b.p) and np.array_equal(a.j, b.j)


def pat(m):
    s = m.to_scipy().astype(bool).astype(np.int64)
    s.sort_indices()
    return s


def test_reference_build_and_grown_known_answers():
    k = KAT["build_bool_duplicates_collapse"]
    m = orc.build_matrix(k["nrows"], k["ncols"], k["rows"], k["cols"])
    assert m.nnz == k["nvals"] and m.tuple_set() == {tuple(t) for t in k["present"]}
    g = KAT["grown_shapes"]
    r0, c0 = g["r0"], g["c0"]
    coords = sorted({(i, (i * 7) % c0) for i in range(r0)} | {(i, (i * 11 + 3) % c0) for i in range(r0)})
    src = orc.build_matrix(r0, c0, [c[0] for c in coords], [c[1] for c in coords], np.arange(len(coords), dtype=np.uint64))
    want = {(i, j, v) for v, (i, j) in enumerate(coords)}
    for nr, nc in g["shapes"]:                                   # grown = same tuples in a larger frame
        grown = orc.CSR(nr, nc, np.concatenate([src.p, np.full(nr - r0, src.p[-1])]), src.j, src.x)
        assert grown.nnz == len(coords) and grown.tuple_set() == want
    b = KAT["grown_bool_pattern"]
    mb = orc.build_matrix(b["nrows"], b["ncols"], b["rows"], b["cols"])
    assert mb.nnz == b["nvals"] and mb.x is None                 # a bool layer stays a pattern


def test_reference_query_level_known_answers():
    k = KAT["motogp"]
    rides = orc.build_matrix(6, 6, [e[0] for e in k["rides"]], [e[1] for e in k["rides"]])
    team = 3 + k["teams"].index("Yamaha")
    riders = orc.transpose(rides)                                 # team -> rider
    f = orc.mxm(orc.build_matrix(1, 6, [0], [team]), riders)
    assert [k["riders"][j] for j in f.j] == k["yamaha_riders"] and f.nnz == k["yamaha_count"]
    b = KAT["bfs_flow"]
    names = b["nodes"]
    adj = lambda es: orc.build_matrix(5, 5, [e[0] for e in es], [e[1] for e in es])    # noqa: E731
    reach = lambda A, s, d=-1: [names[v] for v in np.nonzero(orc.bfs(A, s, d)[0] >= 1)[0]]    # noqa: E731
    ALL, E1 = adj(b["E1"] + b["E2"]), adj(b["E1"])
    assert reach(ALL, 0) == b["all_from_a"] and reach(E1, 0) == b["E1_from_a"] and reach(ALL, 0, 1) == b["all_depth1_from_a"]
    for s, nm in enumerate(names):
        assert reach(E1, s) == b["E1_all_sources"].get(nm, [])
        assert reach(ALL, s, 1) == b["all_depth1_all_sources"].get(nm, [])
    v = KAT["variable_length_chain"]
    A = adj(v["edges"])
    A = orc.build_matrix(4, 4, [e[0] for e in v["edges"]], [e[1] for e in v["edges"]])
    hop = orc.mxm(orc.build_matrix(4, 4, range(4), range(4)), A)
    assert [[v["nodes"][i], v["nodes"][j]] for i, j in sorted(hop.tuple_set())] == v["one_hop"]
    F, R = orc.build_matrix(4, 4, range(4), range(4)), orc.build_matrix(4, 4, [], [])
    while True:
        F = orc.mxm(F, A, R, mask_mode=2) if R.nnz else orc.mxm(F, A)
        if F.nnz == 0:
            break
        R = orc.ewise_add(R, F)
    assert R.nnz == v["var_len_pairs"]


def test_fold_threshold_arithmetic():
    """versioned_matrix.rs:140-200: fold when tx_added > 0, delta >= MIN_FOLD_DELTA and (2*delta >= base or
    delta^2 >= K * tx_added); the balance points the reference pins (versioned_matrix.rs:1278-1330).  The C++ mirror of
    the same policy (csrc/host/versioned_matrix.hpp) is exercised by tests/test_host_versioned.py."""
    k = KAT["fold_thresholds"]
    MIN, RK, WK, HUGE = k["MIN_FOLD_DELTA"], k["READ_FOLD_K"], k["WRITE_FOLD_K"], 1 << 62

    def fold(delta, tx, base, K):
        return tx > 0 and delta >= MIN and (2 * delta >= base or delta * delta >= K * tx)

    def threshold(K, tx):
        d = MIN
        while not fold(d, tx, HUGE, K):
            d += 1
        return d

    assert [threshold(RK, 1), threshold(WK, 1), threshold(RK, 100)] == k["expected"]
    assert threshold(WK, 1) // threshold(RK, 1) == 15
    for base in (1_000_000, 10_000_000, 100_000_000, HUGE):
        assert not fold(286, 1, base, RK) and fold(287, 1, base, RK)
    assert fold(512, (1 << 64) - 1, 1024, WK) and fold(512, (1 << 64) - 1, 1024, RK)        # delta comparable to the base
    assert not fold(MIN - 1, 1, 0, WK) and not fold((1 << 64) - 1, 0, 1024, RK)             # tiny deltas, read-only tx


def test_oracle_reproduces_its_frozen_vectors():
    A = unpack("A")
    src = VEC["sources"]
    F = orc.build_matrix(40, A.nrows, np.arange(40), src)
    W = F
    for h in range(3):
        W = orc.mxm(W, A)
        assert same(W, unpack(f"chain{h + 1}")), f"hop {h + 1}"
    assert same(orc.mxm(F, A, unpack("mask"), mask_mode=2), unpack("chain1_rsc"))
    B = unpack("B")
    assert same(orc.ewise_add(A, B), unpack("A_union_B")) and same(orc.ewise_mult(A, B), unpack("A_inter_B"))
    assert same(orc.transpose(A), unpack("A_T"))
    assert same(orc.delta_lmxm(F, A, unpack("dp"), unpack("dm")), unpack("delta_lmxm"))
    lvl, par = orc.bfs(A, int(VEC["bfs_src"][0]))
    assert np.array_equal(lvl, VEC["bfs_level"]) and np.array_equal(par, VEC["bfs_parent"])


def test_frozen_vectors_agree_with_scipy():
    """the same vectors re-derived by an independent implementation (scipy.sparse boolean algebra)"""
    A, B = unpack("A"), unpack("B")
    sA, sB = pat(A), pat(B)
    src = VEC["sources"]
    sF = sp.csr_matrix((np.ones(40, np.int64), (np.arange(40), src)), shape=(40, A.nrows))
    W = sF
    for h in range(3):
        W = (W @ sA).astype(bool).astype(np.int64)
        got = pat(unpack(f"chain{h + 1}"))
        assert (W != got).nnz == 0, f"hop {h + 1}"
    m = pat(unpack("mask")).astype(bool)
    one = (sF @ sA).astype(bool)
    assert ((one.astype(np.int64) - one.multiply(m).astype(np.int64)) != pat(unpack("chain1_rsc"))).nnz == 0
    assert (((sA + sB) > 0).astype(np.int64) != pat(unpack("A_union_B"))).nnz == 0
    assert (sA.multiply(sB).astype(bool).astype(np.int64) != pat(unpack("A_inter_B"))).nnz == 0
    assert (sA.T.tocsr() != pat(unpack("A_T"))).nnz == 0
    from scipy.sparse.csgraph import shortest_path
    d = shortest_path(sA, method="D", unweighted=True, indices=int(VEC["bfs_src"][0]))
    assert np.array_equal(np.where(np.isinf(d), -1, d).astype(np.int64), VEC["bfs_level"])


def test_container_layout_fixture_matches_the_python_mirror():
    from falkordb_b200 import serial
    k = KAT["container_struct"]
    assert serial.CONTAINER_STRUCT_SIZE == k["size"]
    for f, off in k["offsets"].items():
        assert getattr(serial.Container, f).offset == off, f
