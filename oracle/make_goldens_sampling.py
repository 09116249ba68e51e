#!/usr/bin/env python
"""TEST INFRASTRUCTURE. Generates tests/golden/ref_sampling.npz from the compiled reference (oracle/_ref): outputs of
RandBLAS::weights_to_cdf, sample_indices_iid and sample_indices_iid_uniform (RandBLAS/util.hh:459-560) for the cases
of tests/sampling_cases.py (shapes of test/test_basic_rng/test_discrete.cc plus block-boundary and large-n cases).
Run in the build container (needs /root/reference): python oracle/make_goldens_sampling.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402
import sampling_cases as sc  # noqa: E402


def main():
    R = ol.ref()
    assert R is not None, "needs oracle/_ref/librb_ref.so"
    out = {}
    n_cases = 0
    for (n, k, key, off) in sc.uniform_cases():
        ctr, kk = ol.state_from_u64(key)
        ctr = ol.ctr_add(ctr, off)
        tag = f"unif_n{n}_k{k}_key{key}_off{off}"
        res = {}
        for idt in ((np.int64,) if n > 2147483647 else (np.int32, np.int64)):
            s, _, nxt = R.sample_indices_iid_uniform(n, k, ctr, kk, idt, None)
            s2, r32, nxt2 = R.sample_indices_iid_uniform(n, k, ctr, kk, idt, np.float32)
            r64 = R.sample_indices_iid_uniform(n, k, ctr, kk, idt, np.float64)[1]
            assert np.array_equal(r32.astype(np.float64), r64)
            res[idt] = (s.astype(np.int64), s2.astype(np.int64), r32.astype(np.int8), list(nxt), list(nxt2))
        if len(res) == 2:      # the index type does not change the values (n < 2^31): one fixture serves both
            a, b = res[np.int32], res[np.int64]
            assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3])) and a[3:] == b[3:]
        s, s2, sg, nxt, nxt2 = res[np.int64]
        out[tag + "_plain"] = s
        out[tag + "_plain_next"] = np.asarray(nxt, np.uint32)
        out[tag + "_rad"] = s2
        out[tag + "_rad_signs"] = sg
        out[tag + "_rad_next"] = np.asarray(nxt2, np.uint32)
        n_cases += 1
    for name, (w, eib) in sc.weight_vectors().items():
        for dt, dtag in ((np.float32, "f32"), (np.float64, "f64")):
            cdf, ok = R.weights_to_cdf(w.astype(dt), eib)
            out[f"w2c_{name}_{dtag}"] = cdf
            out[f"w2c_{name}_{dtag}_ok"] = np.array([1 if ok else 0], np.int8)
            n_cases += 1
    for name, k in sc.CDF_CASES:
        w, eib = sc.weight_vectors()[name]
        for dt, dtag in ((np.float32, "f32"), (np.float64, "f64")):
            cdf, ok = R.weights_to_cdf(w.astype(dt), eib)
            assert ok
            for key in sc.KEYS:
                ctr, kk = ol.state_from_u64(key)
                ctr = ol.ctr_add(ctr, sc.CDF_COUNTER_OFFSET)
                s32, nxt32 = R.sample_indices_iid(len(cdf), cdf, k, ctr, kk, np.int32)
                s, nxt = R.sample_indices_iid(len(cdf), cdf, k, ctr, kk, np.int64)
                assert np.array_equal(s32.astype(np.int64), s) and list(nxt32) == list(nxt)
                tag = f"iid_{name}_k{k}_key{key}_{dtag}"
                out[tag] = s.astype(np.int32)
                out[tag + "_next"] = np.asarray(nxt, np.uint32)
                n_cases += 1
    path = os.path.join(ROOT, "tests", "golden", "ref_sampling.npz")
    np.savez_compressed(path, **out)
    print("wrote", n_cases, "sampling goldens,", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
