"""The reference's unit tests for the per-relationship-type edge store `Tensor` (graph/src/graph/graphblas/tensor.rs:
1340-1615), transcribed into the C++ host mirror (falkordb_b200/csrc/host/tensor.hpp) and run against libb200grb.so.
Element bookkeeping (promotion / demotion, the host-resident `me` store) runs on the host; every fold / mask / extract /
intersection count goes through the CUDA set-algebra kernels, so the whole file needs the GPU."""
import ctypes as C
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(name):
    L = C.CDLL(os.path.join(ROOT, "falkordb_b200", "libfdbhost.so"))
    L.fdbh_run_test.argtypes = [C.c_char_p]
    L.fdbh_last_message.restype = C.c_char_p
    assert L.fdbh_run_test(name.encode()) == 0, f"{name}: {L.fdbh_last_message().decode()}"


@pytest.mark.gpu
@pytest.mark.parametrize("name", [
    "multi_pairs_after_within_batch_duplicates",      # tensor.rs:1340-1380
    "multi_pairs_matches_the_sentinel_count",         # tensor.rs:1382-1425
])
def test_tensor_multi_edge_bookkeeping(name):
    run(name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", [
    "bulk_remove_and_extract_edge_id_zero",           # tensor.rs:1427-1476  (edge id 0 survives bool round trips)
    "deleting_everything_folds_the_tombstones_away",  # tensor.rs:1503-1530
    "batch_demote_leaves_every_survivor_inline",      # tensor.rs:1548-1571
    "batch_can_demote_and_then_empty_the_same_pair",  # tensor.rs:1573-1615
    "traverse_over_tensor_operand",                   # cond_traverse.rs:83 (TraversalMatrix::U64)
])
def test_tensor_on_device(name):
    run(name)
