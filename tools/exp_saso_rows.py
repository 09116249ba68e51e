#!/usr/bin/env python
"""Timing experiment (not a bench): SASO apply, an 8-lane group per 8 rows of the tile (saso_rows = 0) against a lane per
row (saso_rows = 1), over vec_nnz, d and n; the two results are compared with each other."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402


def timeit(f, reps=5):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    torch.cuda.set_device(0)
    rb.set_option("saso_path", 2)
    cases = [(np.float32, 2048, 8000000, 256, 8), (np.float64, 2048, 4000000, 256, 8)]
    cases += [(np.float32, 2048, 4000000, 256, k) for k in (1, 2, 3, 4, 6, 12, 16, 32)]
    cases += [(np.float32, d, 4000000, 256, 8) for d in (256, 512, 1024, 4096, 8192)]
    cases += [(np.float32, 2048, 4000000, n, 8) for n in (32, 64, 128, 512)]
    for dt, d, m, n, k in cases:
        tdt = torch.float32 if dt == np.float32 else torch.float64
        S = rb.SparseSkOp(rb.SparseDist(d, m, k), rb.RNGState(1997), dtype=dt)
        A = torch.randn(m * n, dtype=tdt, device="cuda")
        out, ts = {}, {}
        for mode in (0, 1):
            rb.set_option("saso_rows", mode)
            B = torch.zeros(d * n, dtype=tdt, device="cuda")
            ts[mode] = timeit(lambda: rb.sketch_general("R", "N", "N", d, n, m, 1.0, S, 0, 0, A, n, 0.0, B, n))
            out[mode] = B.clone()
        err = float(torch.linalg.norm(out[0].double() - out[1].double()) / torch.linalg.norm(out[0].double()))
        print(f"{np.dtype(dt).name} d={d} m={m} n={n} vec_nnz={k}: groups {ts[0]:.3f} ms, lane per row {ts[1]:.3f} ms "
              f"({ts[0] / ts[1]:.2f}x), {m * n * A.element_size() / ts[1] / 1e6:.0f} GB/s of A, rel diff {err:.1e}", flush=True)
        del A


if __name__ == "__main__":
    main()
