"""CondVarLenTraverse's trail enumerator in the C++ host mirror (falkordb_b200/csrc/host/cond_var_len_traverse.hpp, mirroring
graph/src/runtime/ops/cond_var_len_traverse.rs:152-386) against a brute-force enumeration of every relationship-unique path:
outgoing / incoming / bidirectional expansion, hop windows incl. 0 and windows longer than any trail, a fixed destination, and
the adjacency-order emission of a frame.  The DFS itself makes no GraphBLAS call, so this runs without a GPU; the same check over
a Tensor (adjacency fetched through the row iterators of the C ABI) is tests/test_zz2_after_last_gpu_session.py."""
from test_host_tensor import run


def test_var_len_trail_enumeration_logic_host_only():
    run("var_len_trails_logic")


def test_var_len_reference_flow_test_goldens_host_only():
    """tests/flow/test_variable_length_traversals.py test02 / 06 / 07 / 11 / 12 / 13: row counts, (from, to) pairs and path lengths"""
    run("var_len_flow_goldens")


def test_output_batch_repack_logic_host_only():
    """batch.rs:81, 274-287: <= 1024 rows per output batch, NodeIds + u16 selection vector, order preserved (cond_traverse.hpp: repack)"""
    run("repack_logic")
