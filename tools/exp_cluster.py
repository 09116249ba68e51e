#!/usr/bin/env python
"""Timing / equality experiment (not a bench): float dense sketch with and without the 2-CTA cluster that shares the
generated operator tile (rb.set_option("tc_cluster", 0 | 1 | 2)).  python tools/exp_cluster.py"""
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


OPT = ("tc_cluster", 2, "cluster of 2")


def run(d, m, n, fam, lay, ro=0, co=0):
    D = rb.DenseDist(d + ro, m + co, fam, "L")
    S = rb.DenseSkOp(D, rb.RNGState(1997), np.float32)
    A = torch.randn(m * n, dtype=torch.float32, device="cuda")
    lda, ldb = (m, d) if lay == "C" else (n, n)
    outs = {}
    name, on, label = OPT
    default_opt = rb.get_option(name)
    for opt in (0, on):
        rb.set_option(name, opt)
        B = torch.zeros(d * n, dtype=torch.float32, device="cuda")
        f = lambda: rb.sketch_general(lay, "N", "N", d, n, m, 1.0, S, ro, co, A, lda, 0.0, B, ldb)
        ms = timeit(f)
        outs[opt] = (ms, B.clone())
    rb.set_option(name, default_opt)
    same = bool(torch.equal(outs[0][1], outs[on][1]))
    print(f"d={d} m={m} n={n} {fam} data {'K' if lay == 'C' else 'Q'}-contiguous window ({ro},{co}): off {outs[0][0]:.3f} ms "
          f"({2.0 * d * m * n / outs[0][0] / 1e9:.0f} TFLOP/s), {label} {outs[on][0]:.3f} ms "
          f"({2.0 * d * m * n / outs[on][0] / 1e9:.0f} TFLOP/s), bit-identical {same}", flush=True)


if __name__ == "__main__":
    torch.cuda.set_device(0)
    if len(sys.argv) > 1 and sys.argv[1] == "halves":
        OPT = ("tc_halves", 1, "two halves")
        rb.set_option("tc_cluster", 0)
    run(256, 4096, 512, "U", "C")
    run(1024, 100000, 1024, "U", "C")
    run(1024, 100000, 1024, "G", "C")
    run(1024, 100000, 1024, "U", "R")
    run(1024, 100000, 1024, "G", "R")
    run(1000, 50001, 700, "G", "C", 3, 5)
    run(4096, 20000, 2048, "G", "C")
