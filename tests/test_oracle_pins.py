"""CPU-only: pins the oracle (oracle/rb_oracle.c) before anything is allowed to trust it.

1. the reference's own known-answer vectors (r123_kat_vectors.txt:19-21, 50-52),
2. the reference's counter-carry and RNGState(uint64) cases (test_r123.cc:679-797),
3. the uneg11/u01 histograms hard-coded at test_r123.cc:612,618,
4. every fixture under tests/golden/ (outputs of the reference itself, oracle/make_goldens.py),
5. the compiled reference live, where oracle/_ref exists.
"""
import numpy as np
import pytest

import oracle_lib as ol
from golden_util import dense_data


def test_philox_kat(port, gold):
    kats = gold.kat("philox4x32 10")
    assert len(kats) == 3
    for w in kats:
        assert list(port.philox(w[0:4], w[4:6])) == w[6:10]


def test_threefry_kat(port, gold):
    kats = gold.kat("threefry4x32 20")
    assert len(kats) == 3
    for w in kats:
        assert list(port.threefry(w[0:4], w[4:8])) == w[8:12]


def test_counter_carry_like_reference_tests(port):
    # test_r123.cc:710-797: incr across limb boundaries, little-endian limbs
    assert list(port.ctr_incr([0xFFFFFFFF, 0, 0, 0], 1)) == [0, 1, 0, 0]
    assert list(port.ctr_incr([0xFFFFFFFF, 0xFFFFFFFF, 0, 0], 1)) == [0, 0, 1, 0]
    assert list(port.ctr_incr([0xFFFFFFFF] * 4, 1)) == [0, 0, 0, 0]
    assert list(port.ctr_incr([0, 0, 0, 0], (1 << 64) - 1)) == [0xFFFFFFFF, 0xFFFFFFFF, 0, 0]
    assert list(port.ctr_incr([5, 0, 0, 0], (1 << 64) - 1)) == [4, 0, 1, 0]
    # test_r123.cc:679-698: RNGState(uint64) puts the integer into the key, low limb first
    c, k = port.rngstate_from_u64(0x0000000100000002)
    assert list(c) == [0, 0, 0, 0] and list(k) == [2, 1]


def test_uniform_histograms_from_reference_test(port):
    # test_r123.cc:568-671: 1000 Threefry4x32-20 blocks, ctr incremented before use, key 0, 26 bins on [-1,1]
    ref_u01 = " 0 0 0 0 0 0 0 0 0 0 0 0 0 301 330 326 320 295 291 298 287 305 307 310 316 314"
    ref_uneg11 = " 156 139 148 146 159 148 159 168 142 160 156 161 153 143 158 150 180 174 152 163 157 129 166 151 140 142"
    for fn, want in ((port.u01, ref_u01), (port.uneg11, ref_uneg11)):
        hist = [0] * 26
        c = np.zeros(4, np.uint32)
        for _ in range(1000):
            c = port.ctr_incr(c, 1)
            for w in port.threefry(c, [0, 0, 0, 0]):
                u = fn(w)
                assert -1.0 <= u <= 1.0
                hist[int((np.float32(u) + np.float32(1.0)) * np.float32(13))] += 1
        assert "".join(f" {h}" for h in hist) == want


def test_golden_philox_ctr_state(port, gold):
    for c in gold.m["philox"]:
        assert list(port.philox(c["ctr"], c["key"])) == c["out"]
    for c in gold.m["ctr_incr"]:
        assert list(port.ctr_incr(c["ctr"], c["n"])) == c["out"]
    for c in gold.m["rngstate_u64"]:
        cc, kk = port.rngstate_from_u64(c["k"])
        assert list(cc) == c["ctr"] and list(kk) == c["key"]


def test_golden_blocks(port, gold):
    for b in gold.m["blocks"]:
        w = port.philox(b["ctr"], b["key"])
        u = np.array([port.uneg11(x) for x in w], np.float32)
        assert np.array_equal(u.view(np.uint32), gold.arr(b["uneg11"]).view(np.uint32))
        g = np.array(port.boxmuller(w[0], w[1]) + port.boxmuller(w[2], w[3]), np.float32)
        assert np.array_equal(g.view(np.uint32), gold.arr(b["boxmul"]).view(np.uint32))


def test_golden_fill_dense(port, gold):
    assert len(gold.m["fill_dense"]) > 300
    for c in gold.m["fill_dense"]:
        r, cc, fam, ax = c["D"]
        nr, nc, ro, co = c["sub"]
        buf, nxt = port.fill_dense_unpacked(c["layout"], r, cc, fam, ax, nr, nc, ro, co, c["ctr"], c["key"],
                                            np.dtype(c["dtype"]))
        want = gold.arr(c["buff"])
        assert np.array_equal(buf.view(np.uint8), want.view(np.uint8)), c
        assert list(nxt) == c["next_ctr"], c


def test_golden_next_state(port, gold):
    ctr, key = ol.state_from_u64(1997)
    for c in gold.m["next_state"]:
        if c["kind"] == "dense":
            r, cc, fam, ax = c["D"]
            assert port.dense_dist_info(r, cc, fam, ax) == c["info"]
            assert list(port.dense_next_state(r, cc, fam, ax, ctr, key)) == c["next_ctr"]
        else:
            r, cc, vn, ax = c["D"]
            assert port.sparse_dist_info(r, cc, vn, ax) == c["info"]
            assert list(port.sparse_next_state(r, cc, vn, ax, ctr, key)) == c["next_ctr"]


def test_golden_saso(port, gold):
    for c in gold.m["saso"]:
        r, cc, vn, ax = c["D"]
        idt = np.dtype(c["idx"])
        vals, rows, cols, nnz, nxt = port.fill_sparse(r, cc, vn, ax, c["ctr"], c["key"], np.float32, idt)
        assert nnz == c["nnz"] and list(nxt) == c["next_ctr"]
        assert np.array_equal(vals.astype(np.int8), gold.arr(c["vals"]))
        assert np.array_equal(rows.astype(np.int32), gold.arr(c["rows"]))
        assert np.array_equal(cols.astype(np.int32), gold.arr(c["cols"]))
    for c in gold.m["rfy"]:
        s, nxt = port.repeated_fisher_yates(c["k"], c["n"], c["r"], c["ctr"], c["key"])
        assert np.array_equal(s.astype(np.int32), gold.arr(c["samples"])) and list(nxt) == c["next_ctr"]


def _tol(dtype):
    # north_star tolerances on relative Frobenius error
    return 1e-5 if np.dtype(dtype) == np.float32 else 1e-12


def sketch_case_inputs(fill, c):
    """Rebuild (A or sparse A, B0) for a golden sketch case from the shared recipe."""
    dt = np.dtype(c["dtype"])
    lay = c["layout"]
    kind = c["kind"]
    if kind in ("lskge3", "lskges", "lsksp3"):
        d, n, m = c["dims"]
        rB, cB = d, n
    else:
        m, d, n = c["dims"]
        rB, cB = m, d
    rA, cA = (m, n) if c["opA"] == "N" else (n, m)
    Arm = dense_data(fill, rA, cA, 99, dt)
    B0 = dense_data(fill, rB, cB, 42, dt)
    B = np.ascontiguousarray(B0.T if lay == "C" else B0).ravel().copy()
    ldb = rB if lay == "C" else cB
    if kind == "lsksp3":
        import scipy.sparse as sp
        Arm = Arm.copy()
        Arm[np.abs(Arm) <= c["thresh"]] = 0
        M = sp.csr_matrix(Arm)
        if c["fmt"] == 0:
            spA = (m, n, M.nnz, M.data.astype(dt), M.indptr.astype(np.int64), M.indices.astype(np.int64))
        elif c["fmt"] == 1:
            Mc = M.tocsc()
            spA = (m, n, Mc.nnz, Mc.data.astype(dt), Mc.indices.astype(np.int64), Mc.indptr.astype(np.int64))
        else:
            Mo = M.tocoo()
            spA = (m, n, Mo.nnz, Mo.data.astype(dt), Mo.row.astype(np.int64), Mo.col.astype(np.int64))
        return spA, None, B, ldb
    A = np.ascontiguousarray(Arm.T if lay == "C" else Arm).ravel()
    lda = rA if lay == "C" else cA
    return A, lda, B, ldb


def run_sketch_case(impl, c, A, lda, B, ldb):
    ctr, key = ol.state_from_u64(1997)
    dt = np.dtype(c["dtype"]).type
    al, be = dt(c["alpha"]), dt(c["beta"])
    lay, opS, opA = c["layout"], c["opS"], c["opA"]
    D = tuple(c["D"])
    ro, co = c["off"]
    if c["kind"] == "lskge3":
        d, n, m = c["dims"]
        impl.lskge3(lay, opS, opA, d, n, m, al, D, ctr, key, ro, co, A, lda, be, B, ldb)
    elif c["kind"] == "rskge3":
        m, d, n = c["dims"]
        impl.rskge3(lay, opA, opS, m, d, n, al, A, lda, D, ctr, key, ro, co, be, B, ldb)
    elif c["kind"] == "lskges":
        d, n, m = c["dims"]
        impl.lskges(lay, opS, opA, d, n, m, al, D, ctr, key, ro, co, A, lda, be, B, ldb)
    elif c["kind"] == "lsksp3":
        d, n, m = c["dims"]
        impl.lsksp3(c["fmt"], lay, opS, opA, d, n, m, al, D, ctr, key, ro, co, A, be, B, ldb)
    else:
        raise KeyError(c["kind"])


def _laso_vectors(rows, cols, vals, n_rows, n_cols):
    """(short index, long index, value) triples sorted inside each long-axis vector: the order-free view."""
    sh, lg = (rows, cols) if n_rows <= n_cols else (cols, rows)
    return sorted(zip(sh.tolist(), lg.tolist(), vals.tolist()))


def laso_golden_cases():
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_laso.npz"))
    tags = sorted({k[: -len("_rows")] for k in z.files if k.endswith("_rows")})
    for t in tags:
        shape, kk, key, itag = t[len("laso_"):].split("_")
        r, c = map(int, shape.split("x"))
        yield t, r, c, int(kk[1:]), int(key[3:]), (np.int32 if itag == "i32" else np.int64), z


def test_golden_laso(port):
    """LASO operators (RandBLAS/sparse_skops.hh:534-564) against fixtures generated from the compiled reference:
    nnz, next state, per-vector (index, value) sets identical; draw order identical where a vector has no repeats
    (the reference's order for vectors with repeats is std::unordered_map's and is not part of the contract)."""
    n = 0
    for t, r, c, vn, k, idt, z in laso_golden_cases():
        ctr, key = ol.state_from_u64(k)
        for dt, vtag in ((np.float32, "_v32"), (np.float64, "_v64")):
            vals, rows, cols, nnz, nxt = port.fill_sparse(r, c, vn, "L", ctr, key, dt, idt)
            assert nnz == len(z[t + "_rows"]), t
            assert list(nxt) == list(z[t + "_next"]), t
            assert _laso_vectors(rows[:nnz], cols[:nnz], vals[:nnz], r, c) == \
                _laso_vectors(z[t + "_rows"], z[t + "_cols"], z[t + vtag], r, c), t
            if nnz == vn * min(r, c):      # no vector had a repeat: same order, element for element
                assert np.array_equal(rows[:nnz], z[t + "_rows"]) and np.array_equal(cols[:nnz], z[t + "_cols"]), t
                assert np.array_equal(vals[:nnz], z[t + vtag]), t
        n += 1
    assert n >= 80


def check_sampling_goldens(impl):
    """Index-sampling utilities (RandBLAS/util.hh:459-560) of `impl` against tests/golden/ref_sampling.npz, generated
    from the compiled reference by oracle/make_goldens_sampling.py: samples, Rademacher signs and next states
    element for element; CDFs bit for bit (the sum is serial in T on every side); where the reference throws,
    the implementation must fail too and leave the same partial result."""
    import os
    import sampling_cases as sc
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_sampling.npz"))
    n_checked = 0
    for (n, k, key, off) in sc.uniform_cases():
        ctr, kk = ol.state_from_u64(key)
        ctr = ol.ctr_add(ctr, off)
        tag = f"unif_n{n}_k{k}_key{key}_off{off}"
        for idt in ((np.int64,) if n > 2147483647 else (np.int32, np.int64)):
            s, _, nxt = impl.sample_indices_iid_uniform(n, k, ctr, kk, idt, None)
            assert np.array_equal(s.astype(np.int64), z[tag + "_plain"]), tag
            assert list(nxt) == list(z[tag + "_plain_next"]), tag
            for rdt in (np.float32, np.float64):
                s, r, nxt = impl.sample_indices_iid_uniform(n, k, ctr, kk, idt, rdt)
                assert np.array_equal(s.astype(np.int64), z[tag + "_rad"]), tag
                assert np.array_equal(r, z[tag + "_rad_signs"].astype(rdt)), tag
                assert list(nxt) == list(z[tag + "_rad_next"]), tag
        n_checked += 1
    cdfs = {}
    for name, (w, eib) in sc.weight_vectors().items():
        for dt, dtag in ((np.float32, "f32"), (np.float64, "f64")):
            cdf, ok = impl.weights_to_cdf(w.astype(dt), eib)
            assert ok == bool(z[f"w2c_{name}_{dtag}_ok"][0]), (name, dtag)
            assert np.array_equal(cdf, z[f"w2c_{name}_{dtag}"]), (name, dtag)        # bit for bit, partial results too
            cdfs[name, dtag] = cdf
            n_checked += 1
    for name, k in sc.CDF_CASES:
        for dtag in ("f32", "f64"):
            cdf = cdfs[name, dtag]
            for key in sc.KEYS:
                ctr, kk = ol.state_from_u64(key)
                ctr = ol.ctr_add(ctr, sc.CDF_COUNTER_OFFSET)
                tag = f"iid_{name}_k{k}_key{key}_{dtag}"
                for idt in (np.int32, np.int64):
                    s, nxt = impl.sample_indices_iid(len(cdf), cdf, k, ctr, kk, idt)
                    assert np.array_equal(s.astype(np.int32), z[tag]), tag
                    assert list(nxt) == list(z[tag + "_next"]), tag
                n_checked += 1
    assert n_checked >= 100


def test_golden_sampling_utilities(port):
    check_sampling_goldens(port)


def test_sampling_utilities_port_equals_reference_live(port, ref):
    """Random shapes beyond the fixtures: the C restatement against the compiled reference."""
    rng = np.random.default_rng(5)
    for t in range(40):
        n = int(rng.integers(1, 5000))
        k = int(rng.integers(0, 3000))
        ctr, kk = ol.state_from_u64(int(rng.integers(0, 1 << 40)))
        ctr = ol.ctr_add(ctr, int(rng.integers(0, 1 << 34)))
        idt = (np.int32, np.int64)[t % 2]
        rdt = (None, np.float32, np.float64)[t % 3]
        a, b = port.sample_indices_iid_uniform(n, k, ctr, kk, idt, rdt), ref.sample_indices_iid_uniform(n, k, ctr, kk, idt, rdt)
        assert np.array_equal(a[0], b[0]) and list(a[2]) == list(b[2])
        assert (a[1] is None and b[1] is None) or np.array_equal(a[1], b[1])
        dt = (np.float32, np.float64)[(t // 2) % 2]
        w = rng.random(n).astype(dt) ** 3
        ca, oka = port.weights_to_cdf(w)
        cb, okb = ref.weights_to_cdf(w)
        assert oka == okb and np.array_equal(ca, cb)
        if oka:
            a, b = port.sample_indices_iid(n, ca, k, ctr, kk, idt), ref.sample_indices_iid(n, cb, k, ctr, kk, idt)
            assert np.array_equal(a[0], b[0]) and list(a[1]) == list(b[1])


def test_golden_sketch_products(port, gold):
    assert len(gold.m["sketch"]) >= 30
    for c in gold.m["sketch"]:
        A, lda, B, ldb = sketch_case_inputs(port.fill_dense_unpacked, c)
        run_sketch_case(port, c, A, lda, B, ldb)
        want = gold.arr(c["B"])
        err = np.linalg.norm(B.astype(np.float64) - want) / np.linalg.norm(want)
        assert err < _tol(c["dtype"]), (c["kind"], c["dtype"], err)


def test_port_equals_reference_live(port, ref):
    """Where the compiled reference is available, sweep more shapes than the fixtures hold."""
    rng = np.random.default_rng(5)
    for _ in range(60):
        r, c = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        fam, ax, lay = "GU"[rng.integers(2)], "LS"[rng.integers(2)], "RC"[rng.integers(2)]
        nr, nc = int(rng.integers(1, r + 1)), int(rng.integers(1, c + 1))
        ro, co = int(rng.integers(0, r - nr + 1)), int(rng.integers(0, c - nc + 1))
        ctr, key = ol.state_from_u64(int(rng.integers(0, 1 << 62)))
        ctr = ol.ctr_add(ctr, int(rng.integers(0, 1 << 62)))
        for dt in (np.float32, np.float64):
            a, n1 = port.fill_dense_unpacked(lay, r, c, fam, ax, nr, nc, ro, co, ctr, key, dt)
            b, n2 = ref.fill_dense_unpacked(lay, r, c, fam, ax, nr, nc, ro, co, ctr, key, dt)
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8)) and list(n1) == list(n2)
    for _ in range(30):
        r, c = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        vn = int(rng.integers(1, min(r, c) + 1))
        ctr, key = ol.state_from_u64(int(rng.integers(0, 1 << 62)))
        a = port.fill_sparse(r, c, vn, "S", ctr, key, np.float64, np.int64)
        b = ref.fill_sparse(r, c, vn, "S", ctr, key, np.float64, np.int64)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


def test_reference_argument_errors(ref):
    ctr, key = ol.state_from_u64(0)
    with pytest.raises(ol.OracleError):
        ref.fill_dense_unpacked("R", 4, 8, "G", "L", 4, 8, 1, 0, ctr, key, np.float32)  # window exceeds D
    with pytest.raises(ol.OracleError):
        ref.sparse_dist_info(4, 8, 5, "S")  # vec_nnz > dim_major


def test_port_argument_errors(port):
    ctr, key = ol.state_from_u64(0)
    with pytest.raises(ol.OracleError):
        port.fill_dense_unpacked("R", 4, 8, "G", "L", 4, 8, 1, 0, ctr, key, np.float32)
    with pytest.raises(ol.OracleError):
        port.sparse_dist_info(4, 8, 5, "S")
