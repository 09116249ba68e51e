#!/usr/bin/env python
"""Timing + bit-identity experiment (not a bench): Gaussian float tensor-core sketch, fused generation against
"generate each K panel once, then the materialised-operator kernel" (tc_materialise)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import randblas_b200 as rb  # noqa: E402
from exp_pair import timeit  # noqa: E402


def main():
    torch.cuda.set_device(0)
    d, m = 1024, 100000
    for n in (256, 512, 768, 1024, 2048):
        A = torch.randn(m * n, dtype=torch.float32, device="cuda")
        flops = 2.0 * d * m * n
        for axis in (rb.Axis.Long, rb.Axis.Short):
            S = rb.DenseSkOp(rb.DenseDist(d, m, rb.ScalarDist.Gaussian, axis), rb.RNGState(1997), np.float32)
            out = {}
            for mat in (0, 2):
                rb.set_option("tc_materialise", mat)
                B = torch.zeros(d * n, dtype=torch.float32, device="cuda")
                t = timeit(lambda: rb.sketch_general("C", "N", "N", d, n, m, 1.0, S, 0, 0, A, m, 0.0, B, d), reps=5)
                out[mat] = B.clone()
                print(f"n={n} axis {axis} tc_materialise={mat}: {t:.3f} ms, {flops / t / 1e9:.1f} TFLOP/s", flush=True)
            rel = float((out[0].double() - out[2].double()).norm() / out[0].double().norm())
            print(f"   bit-identical {torch.equal(out[0], out[2])}, rel diff {rel:.2e}", flush=True)
        del A
    rb.set_option("tc_materialise", 1)


if __name__ == "__main__":
    main()
