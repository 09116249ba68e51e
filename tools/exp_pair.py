#!/usr/bin/env python
"""Timing + bit-identity experiment (not a bench): float tensor-core sketch with and without CTA pairs (cta_group::2)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402


def timeit(f, reps=10):
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
    d, m, n = 1024, 100000, 1024
    A = torch.randn(m * n, dtype=torch.float32, device="cuda")
    flops = 2.0 * d * m * n
    for fam in (rb.ScalarDist.Uniform, rb.ScalarDist.Gaussian):
        for axis in (rb.Axis.Long, rb.Axis.Short):
            S = rb.DenseSkOp(rb.DenseDist(d, m, fam, axis), rb.RNGState(1997), np.float32)
            out = {}
            for pair in (0, 2):
                rb.set_option("tc_pair", pair)
                B = torch.zeros(d * n, dtype=torch.float32, device="cuda")
                f = lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, m, 0.0, B, d)
                t = timeit(f)
                out[pair] = (t, B.clone())
                print(f"{fam} axis {axis} tc_pair={pair}: {t:.3f} ms, {flops / t / 1e9:.1f} TFLOP/s", flush=True)
            same = torch.equal(out[0][1], out[1][1])
            rel = float((out[0][1].double() - out[1][1].double()).norm() / out[0][1].double().norm())
            print(f"   pair vs single-CTA result: bit-identical {same}, rel diff {rel:.2e}", flush=True)
    rb.set_option("tc_pair", 1)


if __name__ == "__main__":
    main()
