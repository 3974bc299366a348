"""The C-compatible RDB form of a relationship Tensor (graph/src/graph/graphblas/tensor.rs:1049-1204) and of a VersionedMatrix
(versioned_matrix.rs:1082-1113) in the C++ host mirror (csrc/host/serial.hpp, tensor.hpp, versioned_matrix.hpp), over the C ABI's
serialization entry points.  Decoding a stream laid out the way C FalkorDB writes it -- (count | MSB) in the forward matrix for a
multi-edge pair, the pair's ids as the INDICES of a BOOL vector blob -- touches host-resident matrices only, so these run without
a GPU; the round trip after device-side mutations is tests/test_zz2_after_last_gpu_session.py."""
import pytest

from test_host_tensor import run


@pytest.mark.parametrize("name", ["tensor_decodes_the_c_written_form", "versioned_matrix_encode_decode"])
def test_rdb_forms_host_only(name):
    run(name)
