"""The reference's own unit tests for the delta matrix (graph/src/graph/graphblas/versioned_matrix.rs:1278-1523),
transcribed into the C++ host mirror (falkordb_b200/csrc/host/host_capi.cpp) and run against libb200grb.so.
The fold-policy arithmetic is host-only; everything that folds / masks / merges runs CUDA kernels."""
import ctypes as C
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def host():
    L = C.CDLL(os.path.join(ROOT, "falkordb_b200", "libfdbhost.so"))
    L.fdbh_run_test.argtypes = [C.c_char_p]
    L.fdbh_last_message.restype = C.c_char_p
    return L


def run(name):
    L = host()
    rc = L.fdbh_run_test(name.encode())
    assert rc == 0, f"{name}: {L.fdbh_last_message().decode()}"


@pytest.mark.parametrize("name", [
    "read_path_balance_point_is_flat_in_base_size",          # versioned_matrix.rs:1278-1287 (287)
    "write_path_is_16x_looser_than_read_path",               # :1289-1297 (4528, ratio 15)
    "balance_point_grows_as_sqrt_of_transaction_size",       # :1299-1312 (2864)
    "delta_comparable_to_base_always_folds",                 # :1314-1321
    "tiny_deltas_and_read_only_transactions_never_fold",     # :1323-1330
])
def test_fold_policy_arithmetic(name):
    run(name)


@pytest.mark.gpu
def test_delta_invariants_hold_across_mutation_sequences():
    """versioned_matrix.rs:1399-1472: 4,000-step LCG (seed 0x5eed_1234) over set/remove/set_all/remove_mask/dup/
    wait/fold_oversized against a set model; checks dp n m = 0, dm subset m, nvals arithmetic and the 3-way Iter."""
    run("delta_invariants_hold_across_mutation_sequences")


@pytest.mark.gpu
def test_folded_entry_deleted_and_re_added_stays_out_of_dp():
    """versioned_matrix.rs:1481-1523"""
    run("folded_entry_deleted_and_re_added_stays_out_of_dp")
