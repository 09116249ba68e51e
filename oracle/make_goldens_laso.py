#!/usr/bin/env python
"""TEST INFRASTRUCTURE. Generates tests/golden/ref_laso.npz from the compiled reference (oracle/_ref): LASO
(Axis::Long) operators sampled by RandBLAS::fill_sparse_unpacked_nosub (RandBLAS/sparse_skops.hh:534-564) for the
shapes of the reference's own LASO tests (test/test_datastructures/test_sparseskop.cc:209-261: 7x20, 15x7,
vec_nnz 1,2,3,7, keys 42,0,1) plus shapes with many repeated indices and a benchmark-like one. Run in the build
container (needs /root/reference): python oracle/make_goldens_laso.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402

CASES = [(7, 20, 1), (7, 20, 2), (7, 20, 3), (7, 20, 7), (15, 7, 1), (15, 7, 2), (15, 7, 3), (15, 7, 7),
         (5, 40, 6), (40, 5, 6), (3, 9, 9), (12, 3000, 8), (64, 5000, 33), (2000, 31, 4)]


def main():
    R = ol.ref()
    assert R is not None, "needs oracle/_ref/librb_ref.so"
    out = {}
    n = 0
    for (r, c, vn) in CASES:
        for k in (42, 0, 1):
            ctr, key = ol.state_from_u64(k)
            for idt, itag in ((np.int32, "i32"), (np.int64, "i64")):
                v32, rows, cols, nnz, nxt = R.fill_sparse(r, c, vn, "L", ctr, key, np.float32, idt)
                v64 = R.fill_sparse(r, c, vn, "L", ctr, key, np.float64, idt)[0]
                tag = f"laso_{r}x{c}_k{vn}_key{k}_{itag}"
                out[tag + "_rows"] = rows[:nnz].copy()
                out[tag + "_cols"] = cols[:nnz].copy()
                out[tag + "_v32"] = v32[:nnz].copy()
                out[tag + "_v64"] = v64[:nnz].copy()
                out[tag + "_next"] = np.asarray(nxt, np.uint32)
                n += 1
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_laso.npz"), **out)
    print("wrote", n, "LASO goldens")


if __name__ == "__main__":
    main()
