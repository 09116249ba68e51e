"""CPU-only: the C-ABI library loads and exports every symbol include/randblas_b200.h declares; the host-side
integer arithmetic (RNG state, distributions, next_state) matches the fixtures without touching a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "randblas_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "randblas_b200", "librandblas_b200.so"))
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/randblas_b200.h but not exported"


def test_host_side_state_arithmetic_matches_goldens(gold):
    import randblas_b200 as rb
    for c in gold.m["rngstate_u64"]:
        s = rb.RNGState(c["k"])
        assert s.counter == c["ctr"] and s.key == c["key"]
    for c in gold.m["ctr_incr"]:
        assert rb.RNGState(counter=c["ctr"], key=[0, 0]).incr(c["n"]).counter == c["out"]
    for c in gold.m["next_state"]:
        st = rb.RNGState(1997)
        if c["kind"] == "dense":
            r, cc, fam, ax = c["D"]
            D = rb.DenseDist(r, cc, fam, ax)
            assert dict(dim_major=D.dim_major, dim_minor=D.dim_minor, natural_layout=D.natural_layout,
                        isometry_scale=D.isometry_scale) == c["info"]
            assert rb.DenseSkOp(D, st).next_state.counter == c["next_ctr"]
        else:
            r, cc, vn, ax = c["D"]
            D = rb.SparseDist(r, cc, vn, ax)
            assert dict(dim_major=D.dim_major, dim_minor=D.dim_minor, full_nnz=D.full_nnz,
                        isometry_scale=D.isometry_scale) == c["info"]
            assert rb.SparseSkOp(D, st).next_state.counter == c["next_ctr"]


def test_argument_errors_without_gpu():
    import randblas_b200 as rb
    with pytest.raises(rb.RandBLASError):
        rb.DenseDist(0, 5)
    with pytest.raises(rb.RandBLASError):
        rb.SparseDist(4, 8, 5)          # vec_nnz > dim_major
    D = rb.DenseDist(4, 8)
    buf = np.zeros(32, np.float32)
    with pytest.raises(rb.RandBLASError):  # window exceeds the distribution: rejected before any CUDA call
        rb.fill_dense_unpacked("R", D, 4, 8, 1, 0, buf, rb.RNGState(0))
    S = rb.DenseSkOp(D, rb.RNGState(0))
    A = np.zeros(8 * 3, np.float32)
    B = np.zeros(4 * 3, np.float32)
    with pytest.raises(rb.RandBLASError):  # lda too small (skge.hh:186-192)
        rb.sketch_general("C", "N", "N", 4, 3, 8, 1.0, S, 0, 0, A, 7, 0.0, B, 4)
    with pytest.raises(rb.RandBLASError):  # full-operator overload dimension check (skge.hh:1089-1095)
        rb.sketch_general("C", "N", "N", 3, 3, 8, 1.0, S, A, 8, 0.0, B, 4)


def test_sampling_utilities_argument_checks_and_state_arithmetic_without_gpu():
    """The index-sampling entry points validate their arguments and advance the state before any CUDA call:
    k == 0 returns the state unchanged (util.hh:505-511, 538-544: no block is consumed), negative sizes are rejected."""
    import randblas_b200 as rb
    st = rb.RNGState(7).incr(3456)
    s = np.zeros(1, np.int64)
    assert rb.sample_indices_iid_uniform(40, 0, s, st) == st
    assert rb.sample_indices_iid(1, np.ones(1, np.float32), 0, s, st) == st
    with pytest.raises(rb.RandBLASError):
        rb.sample_indices_iid_uniform(40, -1, s, st)
    with pytest.raises(rb.RandBLASError):
        rb.sample_indices_iid(-1, np.ones(1, np.float32), 1, s, st)
    with pytest.raises(rb.RandBLASError):          # int32 samples cannot index 2^40 values
        rb.sample_indices_iid_uniform(1 << 40, 4, np.zeros(4, np.int32), st)
    rb.weights_to_cdf(0, np.zeros(0, np.float32))      # empty: nothing to do (util.hh:461-472 with n = 0)
    assert rb.get_option("tc_cluster") in (0, 1, 2)
    rb.set_option("tc_cluster", rb.get_option("tc_cluster"))
    with pytest.raises(rb.RandBLASError):
        rb.set_option("no_such_option", 1)


def test_kernel_selection_options_round_trip_without_gpu():
    """The switches that choose among kernels computing the same result (include/randblas_b200.h, "tuning") exist in the
    library with the documented defaults, can be set and read back, and an unknown name is rejected by the setter."""
    import randblas_b200 as rb
    defaults = {"dense_path": 0, "saso_path": 0, "saso_fill_path": 0, "saso_bin_path": 0, "saso_rows": 1, "tc_xmn": 1,
                "tc_ymn": 0, "tc_pair": 1, "fill_rep": 1, "fill_unroll": 1, "spdata_path": 0}
    for name, want in defaults.items():
        assert rb.get_option(name) == want, (name, rb.get_option(name))
        rb.set_option(name, 1 - want if want in (0, 1) else 0)
        assert rb.get_option(name) == (1 - want if want in (0, 1) else 0), name
        rb.set_option(name, want)
        assert rb.get_option(name) == want
    with pytest.raises(rb.RandBLASError):
        rb.set_option("no_such_option", 1)
