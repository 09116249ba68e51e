"""randblas_b200 -- B200-native sketching hot path behind the RandBLAS API.

The product is the CUDA library `librandblas_b200.so` (C ABI: include/randblas_b200.h) plus the header-only
C++ drop-in layer under include/RandBLAS/. This Python package is a thin host-side mirror of the same API
(same names, argument order and error behaviour as the reference) used by the tests and by bench.py.
"""
from .api import (Axis, COOMatrix, CSCMatrix, CSRMatrix, DenseDist, DenseSkOp, Layout, Op, RNGState, ScalarDist,  # noqa
                  SparseDist, SparseSkOp, fill_dense, fill_dense_unpacked, fill_sparse, fill_sparse_unpacked_nosub,
                  philox_words, boxmuller_words, repeated_fisher_yates, sketch_general, sketch_sparse, sketch_vector,
                  left_spmm, right_spmm, coo_to_csr, coo_to_csc, csr_to_coo, csc_to_coo,
                  sketch_symmetric, sample_indices_iid, sample_indices_iid_uniform, weights_to_cdf,
                  random_coo, random_csr, random_csc, sorted_idxs_to_compressed_ptr, csr_column_block,
                  csc_column_block)
from ._lib import RandBLASError, counter, get_option, release_workspace, set_option  # noqa
