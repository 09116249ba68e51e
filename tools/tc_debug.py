#!/usr/bin/env python
"""Debug driver for the tensor-core dense kernels: runs one case per subprocess (a device trap poisons the CUDA
context) and prints the relative error against a float64 torch product of the materialised operator."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (dtype, d, n, m, Dr, Dc, ro, co, fam, alpha, beta, lda_pad)
    "u_aligned": ("f32", 128, 256, 4096, 128, 4096, 0, 0, "U", 1.0, 0.0),
    "g_aligned": ("f32", 128, 256, 4096, 128, 4096, 0, 0, "G", 1.0, 0.0),
    "u_kshift": ("f32", 128, 256, 4096, 128, 5000, 0, 6, "U", 1.0, 0.0),
    "u_raggedq": ("f32", 128, 300, 4096, 128, 4096, 0, 0, "U", 1.0, 0.0),
    "u_raggedp": ("f32", 200, 256, 4096, 210, 4096, 3, 0, "U", 1.0, 0.0),
    "u_raggedk": ("f32", 128, 256, 5003, 128, 6000, 0, 0, "U", 1.0, 0.0),
    "g_all": ("f32", 200, 300, 5003, 210, 6000, 3, 6, "G", 0.5, -1.5),
    "chain1": ("f32", 128, 256, 65536, 128, 65536, 0, 0, "U", 1.0, 0.0),
    "chain4": ("f32", 128, 256, 65536, 128, 65536, 0, 0, "U", 1.0, 0.0),
    "chain16": ("f32", 128, 256, 65536, 128, 65536, 0, 0, "U", 1.0, 0.0),
    "chain64": ("f32", 128, 256, 65536, 128, 65536, 0, 0, "U", 1.0, 0.0),
    "chain256": ("f32", 128, 256, 65536, 128, 65536, 0, 0, "U", 1.0, 0.0),
    "d_aligned": ("f64", 128, 128, 4096, 128, 4096, 0, 0, "G", 1.0, 0.0),
    "d_all": ("f64", 200, 300, 5003, 210, 6000, 3, 6, "G", 0.5, -1.5),
    "d_unif": ("f64", 256, 512, 8192, 256, 8192, 0, 0, "U", 1.0, 0.0),
}


def run_case(name):
    import numpy as np
    import torch
    import randblas_b200 as rb
    dts, d, n, m, Dr, Dc, ro, co, fam, alpha, beta = CASES[name]
    dt = torch.float32 if dts == "f32" else torch.float64
    npdt = np.float32 if dts == "f32" else np.float64
    torch.manual_seed(0)
    lda = m + (4 - m % 4) % 4
    A = torch.randn(n * lda, dtype=dt, device="cuda")
    B0 = torch.randn(n * d, dtype=dt, device="cuda")
    B = B0.clone()
    D = rb.DenseDist(Dr, Dc, fam, "L")
    S = rb.DenseSkOp(D, rb.RNGState(1997), npdt)
    if name.startswith("chain"):
        rb.set_option("tc_splits", int(name[5:]))
    t0 = rb.counter("tensor_core_launches")
    rb.sketch_general("C", "N", "N", d, n, m, alpha, S, ro, co, A, lda, beta, B, d)
    torch.cuda.synchronize()
    used_tc = rb.counter("tensor_core_launches") - t0
    # reference: materialise the operator with fill_dense (bit-exact vs the oracle) and multiply in float64
    Sm = torch.empty(Dr * Dc, dtype=dt, device="cuda")
    rb.fill_dense(D, Sm, rb.RNGState(1997))
    Sm = Sm.view(Dr, Dc)[ro:ro + d, co:co + m].double()
    Am = A.view(n, lda)[:, :m].double().t()          # m x n
    want = alpha * (Sm @ Am) + beta * B0.view(n, d).double().t()
    got = B.view(n, d).double().t()
    err = ((got - want).norm() / want.norm()).item()
    print(f"{name}: tc_launches={used_tc} relerr={err:.3e}")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_case(sys.argv[2])
    else:
        names = sys.argv[1:] or list(CASES)
        for nm in names:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", nm], capture_output=True, text=True,
                               timeout=120, env=dict(os.environ, CUDA_LAUNCH_BLOCKING="1"))
            out = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
            if r.returncode != 0:
                err = [ln for ln in r.stderr.strip().splitlines() if "rror" in ln][-2:]
                print(f"{nm}: FAILED rc={r.returncode} {' | '.join(err)}")
            else:
                print(out)
