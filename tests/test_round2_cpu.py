"""CPU-only checks added in round 2: host-side arithmetic of the multi-GPU entry points, the periodic structure of
the reference's random_coo stream that the device generator relies on, and the buffer validation of the Python layer."""
import ctypes
import math

import numpy as np
import pytest

import oracle_lib as ol


def test_mshard_block_matches_python_partition():
    from randblas_b200._lib import lib
    from randblas_b200.sharding import block
    fn = lib().rb_mshard_block
    fn.restype = ctypes.c_int
    for total in (0, 1, 3, 4, 7, 100, 1003, 4000000, 4000001):
        for world in (1, 2, 3, 5, 8):
            covered = 0
            for r in range(world):
                s, c = ctypes.c_int64(-1), ctypes.c_int64(-1)
                assert fn(ctypes.c_int64(total), world, r, ctypes.byref(s), ctypes.byref(c)) == 0
                assert (s.value, c.value) == block(total, r, world, 4)
                assert s.value == covered and (s.value % 4 == 0 or s.value == total)
                covered += c.value
            assert covered == total
    s, c = ctypes.c_int64(), ctypes.c_int64()
    assert fn(ctypes.c_int64(10), 2, 2, ctypes.byref(s), ctypes.byref(c)) != 0       # rank out of range


def test_comm_entry_points_fail_cleanly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import randblas_b200 as rb
    from randblas_b200.sharding import Comm
    with pytest.raises(rb.RandBLASError):
        Comm(1, 0)


def test_random_coo_stream_is_periodic_in_philox_blocks(ref, port):
    """The fact rb_random_coo_* is built on (csrc/random_matrix.cu): in the reference's random_coo
    (sparse_data/random_matrix.hh:290-355) stored entries 2b and 2b+1 consume exactly the words of Philox block b --
    word 0 and word 3 are their geometric skips, words 1 and 2 the Box-Muller pair they share -- and the returned
    state is seed + nnz / 2 + 1. Checked here on the CPU against the compiled reference."""
    for (m, n, dens, k) in ((50, 70, 0.1, 42), (9, 1000, 0.03, 7), (200, 3, 0.6, 1 << 35)):
        ctr, key = ol.state_from_u64(k)
        v, rows, cols, nnz, nxt = ref.random_sparse(2, m, n, dens, ctr, key, np.float64)
        L, total = math.log(1.0 - dens), m * n
        pos, q, out = None, 0, []
        while True:
            w = ref.philox(ol.ctr_add(ctr, q // 2), key)
            word = int(w[3] if q & 1 else w[0])
            sk = math.floor(math.log(1.0 - (word * 2.0 ** -32 + 2.0 ** -33)) / L)
            pos = sk if pos is None else pos + 1 + sk
            if pos >= total:
                break
            g = port.boxmuller(int(w[1]), int(w[2]))
            out.append((pos // n, pos % n, float(g[q & 1])))
            q += 1
        assert len(out) == nnz
        assert [o[0] for o in out] == list(rows) and [o[1] for o in out] == list(cols)
        assert np.array_equal(np.array([o[2] for o in out]), v)
        assert list(ol.ctr_add(ctr, nnz // 2 + 1)) == list(nxt)


def test_reference_random_csr_csc_are_what_the_wrappers_say(ref):
    """Sanity of the new oracle exports: random_csr / random_csc of the reference are valid compressed matrices with
    sorted, in-range indices and about m*n*density entries."""
    ctr, key = ol.state_from_u64(3)
    m, n, dens = 300, 500, 0.05
    for which in (0, 1):
        v, i0, i1, nnz, _ = ref.random_sparse(which, m, n, dens, ctr, key, np.float32)
        ptr, idx, n_major, n_minor = (i0, i1, m, n) if which == 0 else (i1, i0, n, m)
        assert len(ptr) == n_major + 1 and ptr[0] == 0 and ptr[-1] == nnz and np.all(np.diff(ptr) >= 0)
        assert idx.min() >= 0 and idx.max() < n_minor
        for r in range(0, n_major, 37):
            seg = idx[ptr[r]:ptr[r + 1]]
            assert np.all(np.diff(seg) > 0)
        assert abs(nnz - m * n * dens) < 6 * math.sqrt(m * n * dens)


def test_python_layer_rejects_strided_and_mismatched_buffers():
    import randblas_b200 as rb
    S = rb.DenseSkOp(rb.DenseDist(4, 8), rb.RNGState(0))
    A = np.zeros((8, 6), np.float32)
    B = np.zeros(4 * 3, np.float32)
    with pytest.raises(rb.RandBLASError):          # a strided view would be read as if it were dense
        rb.sketch_general("R", "N", "N", 4, 3, 8, 1.0, S, 0, 0, A[:, ::2], 3, 0.0, B, 3)
    with pytest.raises(rb.RandBLASError):          # A and B must share the scalar type
        rb.sketch_general("R", "N", "N", 4, 3, 8, 1.0, S, 0, 0, np.zeros(24, np.float64), 3, 0.0, B, 3)
    with pytest.raises(rb.RandBLASError):          # d * n not divisible by the number of ranks
        from randblas_b200.sharding import sketch_general_mshard
        sketch_general_mshard("C", 3, 3, 16, 1.0, S, A, 8, B, B, 0, 2)
