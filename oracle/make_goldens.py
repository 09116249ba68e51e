#!/usr/bin/env python
"""TEST INFRASTRUCTURE. Regenerates tests/golden/ from the reference itself.

Runs the UNMODIFIED reference (oracle/_ref/librb_ref.so = /root/reference headers behind
oracle/ref_capi.cc + shims for Random123/BLAS++) and stores its outputs as small fixtures, so
that the parity tests do not need /root/reference (absent on the GPU box).

Usage (in the build container, where /root/reference exists):
    python oracle/make_goldens.py

Cases follow SURVEY.md 8c: Philox words (KAT inputs + carry-boundary counters), fill_dense for
dim_major % 4 in {0,1,2,3} with offsets and both layouts, SASO arrays for the shapes of
test/test_datastructures/test_sparseskop.cc:209-261, and sketch products for the shapes of
test_lskge3.cc / test_lskges.cc / test_sketch_sparse.cc with A = fill_dense(DenseDist(m,n),
RNGState(99)) and B0 = fill_dense(..., RNGState(42)) as in test/test_matmul_cores/linop_common.hh:276-277.
"""
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
KAT_SRC = "/root/reference/test/test_basic_rng/r123_kat_vectors.txt"


def kat_fixture():
    """The Philox4x32-10 / Threefry4x32-20 known-answer lines (r123_kat_vectors.txt:19-21,50-52)."""
    keep = []
    for ln, line in enumerate(open(KAT_SRC), 1):
        if re.match(r"^(philox4x32 10|threefry4x32 20) ", line):
            keep.append(f"{line.strip()}  # r123_kat_vectors.txt:{ln}")
    with open(os.path.join(GOLD, "kat_vectors.txt"), "w") as f:
        f.write("# Known-answer vectors copied from the reference's test fixtures\n"
                "# (test/test_basic_rng/r123_kat_vectors.txt; format: name rounds ctr.. key.. expected..)\n")
        f.write("\n".join(keep) + "\n")
    return len(keep)


def main():
    R = ol.ref()
    assert R is not None, "oracle/_ref/librb_ref.so missing: run `make -C oracle ref` where /root/reference exists"
    os.makedirs(GOLD, exist_ok=True)
    arrays, manifest = {}, {"philox": [], "fill_dense": [], "saso": [], "rfy": [], "sketch": [], "next_state": [],
                            "blocks": []}

    def put(name, a):
        arrays[name] = np.ascontiguousarray(a)
        return name

    nkat = kat_fixture()

    # 1. Philox words
    ctrs = [[0, 0, 0, 0], [1, 0, 0, 0], [0xFFFFFFFF, 0, 0, 0], [0, 1, 0, 0], [0xFFFFFFFF, 0xFFFFFFFF, 0, 0],
            [0xFFFFFFFF] * 4, [0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344]]
    keys = [[0, 0], [42, 0], [1997, 0], [0xFFFFFFFF, 0xFFFFFFFF], [0xA4093822, 0x299F31D0]]
    for c in ctrs:
        for k in keys:
            manifest["philox"].append({"ctr": c, "key": k, "out": [int(x) for x in R.philox(c, k)]})
    # counter arithmetic + RNGState(uint64)
    manifest["ctr_incr"] = []
    for c in ctrs:
        for n in [0, 1, 3, 0xFFFFFFFF, 1 << 32, (1 << 64) - 1, 123456789012345]:
            manifest["ctr_incr"].append({"ctr": c, "n": n, "out": [int(x) for x in R.ctr_incr(c, n)]})
    manifest["rngstate_u64"] = []
    for k in [0, 1, 42, 1997, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, 0x0123456789ABCDEF]:
        c_, k_ = R.rngstate_from_u64(k)
        manifest["rngstate_u64"].append({"k": k, "ctr": [int(x) for x in c_], "key": [int(x) for x in k_]})
    # raw transforms of single blocks (uneg11 and Box-Muller lanes)
    for i, (c, k) in enumerate([(ctrs[0], keys[0]), (ctrs[1], keys[2]), (ctrs[5], keys[3]), (ctrs[6], keys[4])]):
        manifest["blocks"].append({"ctr": c, "key": k, "uneg11": put(f"blk{i}_u", R.uneg11_block(c, k)),
                                   "boxmul": put(f"blk{i}_g", R.boxmul_block(c, k))})

    # 2. fill_dense
    fd_cases = []
    for (r, c) in [(3, 5), (4, 8), (2, 4), (13, 7), (7, 22), (6, 17), (97, 47), (5, 5), (1, 9), (9, 1), (10, 203)]:
        for fam in "GU":
            for ax in "LS":
                fd_cases.append((r, c, fam, ax, r, c, 0, 0))
    # submatrices (offsets chosen to hit every (offset % 4, length % 4) combination)
    for (r, c, nr, nc, ro, co) in [(7, 22, 4, 9, 2, 7), (7, 22, 1, 1, 6, 21), (22, 7, 9, 3, 5, 2), (6, 17, 5, 13, 1, 3),
                                   (17, 6, 11, 2, 6, 1), (10, 203, 3, 100, 4, 50), (10, 203, 10, 5, 0, 198),
                                   (16, 16, 7, 6, 5, 9)]:
        for fam in "GU":
            for ax in "LS":
                fd_cases.append((r, c, fam, ax, nr, nc, ro, co))
    seeds = [1997, 0]
    idx = 0
    for case in fd_cases:
        r, c, fam, ax, nr, nc, ro, co = case
        for seed in seeds[: (2 if (ro, co) == (0, 0) and r * c < 100 else 1)]:
            ctr, key = ol.state_from_u64(seed)
            ctr = ol.ctr_add(ctr, 0 if seed else (1 << 32) - 5)   # seed 0: start just below a limb carry
            for dt in (np.float32, np.float64):
                for lay in "RC":
                    buf, nxt = R.fill_dense_unpacked(lay, r, c, fam, ax, nr, nc, ro, co, ctr, key, dt)
                    manifest["fill_dense"].append({
                        "layout": lay, "D": [r, c, fam, ax], "sub": [nr, nc, ro, co], "ctr": [int(x) for x in ctr],
                        "key": [int(x) for x in key], "dtype": np.dtype(dt).name, "buff": put(f"fd{idx}", buf),
                        "next_ctr": [int(x) for x in nxt]})
                    idx += 1
    # next_state of operators (DenseSkOp / SparseSkOp constructors)
    for (r, c) in [(3, 5), (5, 3), (13, 7), (1024, 100000), (8192, 1000000), (4096, 4000000)]:
        for ax in "LS":
            ctr, key = ol.state_from_u64(1997)
            manifest["next_state"].append({"kind": "dense", "D": [r, c, "G", ax],
                                           "info": R.dense_dist_info(r, c, "G", ax),
                                           "next_ctr": [int(x) for x in R.dense_next_state(r, c, "G", ax, ctr, key)]})
    for (r, c, vn) in [(7, 20, 3), (20, 7, 2), (2048, 8000000, 8), (15, 7, 7)]:
        for ax in "SL":
            ctr, key = ol.state_from_u64(1997)
            manifest["next_state"].append({"kind": "sparse", "D": [r, c, vn, ax],
                                           "info": R.sparse_dist_info(r, c, vn, ax),
                                           "next_ctr": [int(x) for x in R.sparse_next_state(r, c, vn, ax, ctr, key)]})

    # 3. SASO arrays
    idx = 0
    for (r, c) in [(7, 20), (20, 7), (15, 7), (7, 15), (5, 5), (2048, 1000)]:
        for vn in ([1, 2, 3, 7] if r * c < 1000 else [8]):
            if vn > min(r, c):
                continue
            for seed in ([42, 0, 1] if r * c < 1000 else [1997]):
                ctr, key = ol.state_from_u64(seed)
                for idt in (np.int32, np.int64):
                    vals, rows, cols, nnz, nxt = R.fill_sparse(r, c, vn, "S", ctr, key, np.float32, idt)
                    vals64 = R.fill_sparse(r, c, vn, "S", ctr, key, np.float64, idt)[0]
                    assert np.array_equal(vals.astype(np.float64), vals64)
                    manifest["saso"].append({
                        "D": [r, c, vn, "S"], "ctr": [int(x) for x in ctr], "key": [int(x) for x in key],
                        "idx": np.dtype(idt).name, "nnz": nnz, "vals": put(f"ss{idx}_v", vals.astype(np.int8)),
                        "rows": put(f"ss{idx}_r", rows.astype(np.int32)), "cols": put(f"ss{idx}_c", cols.astype(np.int32)),
                        "next_ctr": [int(x) for x in nxt]})
                    idx += 1
    for i, (k, n, r) in enumerate([(3, 10, 6), (10, 10, 3), (1, 4, 9), (4, 4000, 5)]):
        ctr, key = ol.state_from_u64(7)
        s, nxt = R.repeated_fisher_yates(k, n, r, ctr, key)
        manifest["rfy"].append({"k": k, "n": n, "r": r, "ctr": [int(x) for x in ctr], "key": [int(x) for x in key],
                                "samples": put(f"rfy{i}", s.astype(np.int32)), "next_ctr": [int(x) for x in nxt]})

    # 4. sketch products (inputs regenerated by the tests from the same recipe; only outputs stored)
    def data(rows, cols, seed, dt):
        ctr, key = ol.state_from_u64(seed)
        info = R.dense_dist_info(rows, cols, "G", "L")
        buf, _ = R.fill_dense_unpacked("R", rows, cols, "G", "L", rows, cols, 0, 0, ctr, key, dt)
        return buf  # row-major rows x cols

    idx = 0
    ctrS, keyS = ol.state_from_u64(1997)
    for dt in (np.float32, np.float64):
        # dense operator, left: (d, n, m) and operator dims/offsets
        for (lay, opS, opA, d, n, m, Dr, Dc, fam, ax, ro, co, alpha, beta) in [
            ("C", "N", "N", 30, 12, 200, 30, 200, "G", "L", 0, 0, 1.0, 0.0),
            ("R", "N", "N", 30, 12, 200, 30, 200, "U", "L", 0, 0, 1.0, 0.0),
            ("C", "N", "N", 19, 12, 201, 19, 201, "U", "S", 0, 0, 1.0, 0.0),
            ("R", "T", "N", 19, 12, 201, 201, 19, "G", "L", 0, 0, 1.0, 0.0),
            ("C", "N", "T", 3, 7, 10, 8, 12, "G", "L", 3, 1, 0.5, -1.5),
            ("R", "T", "T", 3, 7, 10, 12, 8, "U", "S", 1, 3, 2.0, 1.0),
            ("C", "N", "N", 64, 40, 515, 70, 600, "G", "L", 4, 44, 1.0, 0.0),
        ]:
            rA, cA = (m, n) if opA == "N" else (n, m)
            Arm = data(rA, cA, 99, dt).reshape(rA, cA)
            A = np.ascontiguousarray(Arm.T if lay == "C" else Arm).ravel()
            lda = rA if lay == "C" else cA
            B0rm = data(d, n, 42, dt).reshape(d, n)
            B = np.ascontiguousarray(B0rm.T if lay == "C" else B0rm).ravel().copy()
            ldb = d if lay == "C" else n
            R.lskge3(lay, opS, opA, d, n, m, alpha, (Dr, Dc, fam, ax), ctrS, keyS, ro, co, A, lda, beta, B, ldb)
            manifest["sketch"].append({"kind": "lskge3", "dtype": np.dtype(dt).name, "layout": lay, "opS": opS,
                                       "opA": opA, "dims": [d, n, m], "D": [Dr, Dc, fam, ax], "off": [ro, co],
                                       "alpha": alpha, "beta": beta, "B": put(f"sk{idx}", B)})
            idx += 1
        # dense operator, right: B(m x d) = op(A)(m x n) op(S)(n x d)
        for (lay, opA, opS, m, d, n, Dr, Dc, fam, ax, ro, co, alpha, beta) in [
            ("C", "N", "N", 12, 30, 200, 200, 30, "G", "L", 0, 0, 1.0, 0.0),
            ("R", "T", "T", 7, 3, 10, 8, 12, "U", "S", 3, 1, 0.5, -1.5),
        ]:
            rA, cA = (m, n) if opA == "N" else (n, m)
            Arm = data(rA, cA, 99, dt).reshape(rA, cA)
            A = np.ascontiguousarray(Arm.T if lay == "C" else Arm).ravel()
            lda = rA if lay == "C" else cA
            B0rm = data(m, d, 42, dt).reshape(m, d)
            B = np.ascontiguousarray(B0rm.T if lay == "C" else B0rm).ravel().copy()
            ldb = m if lay == "C" else d
            R.rskge3(lay, opA, opS, m, d, n, alpha, A, lda, (Dr, Dc, fam, ax), ctrS, keyS, ro, co, beta, B, ldb)
            manifest["sketch"].append({"kind": "rskge3", "dtype": np.dtype(dt).name, "layout": lay, "opS": opS,
                                       "opA": opA, "dims": [m, d, n], "D": [Dr, Dc, fam, ax], "off": [ro, co],
                                       "alpha": alpha, "beta": beta, "B": put(f"sk{idx}", B)})
            idx += 1
        # SASO operator, left
        for (lay, opS, opA, d, n, m, Dr, Dc, vn, ro, co, alpha, beta) in [
            ("R", "N", "N", 30, 12, 200, 30, 200, 4, 0, 0, 1.0, 0.0),
            ("C", "N", "N", 19, 12, 201, 19, 201, 3, 0, 0, 1.0, 0.0),
            ("C", "T", "N", 19, 12, 201, 201, 19, 2, 0, 0, 1.0, 0.0),
            ("R", "N", "T", 3, 7, 10, 8, 12, 2, 3, 1, 0.5, -1.5),
        ]:
            rA, cA = (m, n) if opA == "N" else (n, m)
            Arm = data(rA, cA, 99, dt).reshape(rA, cA)
            A = np.ascontiguousarray(Arm.T if lay == "C" else Arm).ravel()
            lda = rA if lay == "C" else cA
            B0rm = data(d, n, 42, dt).reshape(d, n)
            B = np.ascontiguousarray(B0rm.T if lay == "C" else B0rm).ravel().copy()
            ldb = d if lay == "C" else n
            R.lskges(lay, opS, opA, d, n, m, alpha, (Dr, Dc, vn, "S"), ctrS, keyS, ro, co, A, lda, beta, B, ldb)
            manifest["sketch"].append({"kind": "lskges", "dtype": np.dtype(dt).name, "layout": lay, "opS": opS,
                                       "opA": opA, "dims": [d, n, m], "D": [Dr, Dc, vn, "S"], "off": [ro, co],
                                       "alpha": alpha, "beta": beta, "B": put(f"sk{idx}", B)})
            idx += 1
        # sketch_sparse, left, A = deterministic sparse pattern: keep entries of data(m, n, 99) with |x| > 1.2
        for (fmt, lay, opS, d, n, m, Dr, Dc, ro, co, alpha, beta) in [
            (0, "C", "N", 8, 40, 60, 8, 60, 0, 0, 1.0, 0.0),
            (1, "R", "N", 8, 40, 60, 10, 70, 1, 5, 0.5, -1.5),
            (2, "C", "T", 8, 40, 60, 60, 8, 0, 0, 1.0, 0.0),
        ]:
            Arm = data(m, n, 99, dt).reshape(m, n).copy()
            Arm[np.abs(Arm) <= 1.2] = 0
            import scipy.sparse as sp
            M = sp.csr_matrix(Arm)
            if fmt == 0:
                spA = (m, n, M.nnz, M.data.astype(dt), M.indptr.astype(np.int64), M.indices.astype(np.int64))
            elif fmt == 1:
                Mc = M.tocsc()
                spA = (m, n, Mc.nnz, Mc.data.astype(dt), Mc.indices.astype(np.int64), Mc.indptr.astype(np.int64))
            else:
                Mo = M.tocoo()
                spA = (m, n, Mo.nnz, Mo.data.astype(dt), Mo.row.astype(np.int64), Mo.col.astype(np.int64))
            B0rm = data(d, n, 42, dt).reshape(d, n)
            B = np.ascontiguousarray(B0rm.T if lay == "C" else B0rm).ravel().copy()
            ldb = d if lay == "C" else n
            R.lsksp3(fmt, lay, opS, "N", d, n, m, alpha, (Dr, Dc, "G", "L"), ctrS, keyS, ro, co, spA, beta, B, ldb)
            manifest["sketch"].append({"kind": "lsksp3", "fmt": fmt, "dtype": np.dtype(dt).name, "layout": lay,
                                       "opS": opS, "opA": "N", "dims": [d, n, m], "D": [Dr, Dc, "G", "L"],
                                       "off": [ro, co], "alpha": alpha, "beta": beta, "thresh": 1.2,
                                       "B": put(f"sk{idx}", B)})
            idx += 1

    manifest["meta"] = {"generator": "oracle/make_goldens.py", "reference": "BallisticLA/RandBLAS @ 9b73a25",
                        "blas": R.blas_config(), "libm": "glibc 2.39 (Ubuntu 2.39-0ubuntu8.5)", "kat_lines": nkat}
    np.savez_compressed(os.path.join(GOLD, "ref_goldens.npz"), **arrays)
    with open(os.path.join(GOLD, "ref_goldens.json"), "w") as f:
        json.dump(manifest, f, indent=0)
    sz = os.path.getsize(os.path.join(GOLD, "ref_goldens.npz"))
    print(f"wrote {len(arrays)} arrays ({sz/1e6:.2f} MB), {sum(len(v) for v in manifest.values() if isinstance(v, list))} cases")


if __name__ == "__main__":
    main()
