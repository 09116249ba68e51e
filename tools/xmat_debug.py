#!/usr/bin/env python
"""Debug helper: materialised-operator sketches case by case with a synchronize after each."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import randblas_b200 as rb

dt = np.float32 if (len(sys.argv) < 2 or sys.argv[1] == "f32") else np.float64
tdt = torch.float32 if dt == np.float32 else torch.float64
d, n, m = 200, 300, 5000
for (Dr, Dc, ro, co) in ((d, m, 0, 0), (d + 8, m + 12, 3, 5)):
    D = rb.DenseDist(Dr, Dc, "G", "L")
    S0 = rb.DenseSkOp(D, rb.RNGState(1997), dt)
    S1 = rb.DenseSkOp(D, rb.RNGState(1997), dt)
    rb.fill_dense(S1)
    for lay in ("C", "R"):
        A = torch.randn(m * n, dtype=tdt, device="cuda")
        B0 = torch.zeros(d * n, dtype=tdt, device="cuda")
        B1 = torch.zeros(d * n, dtype=tdt, device="cuda")
        rb.sketch_general(lay, "N", "N", d, n, m, 1.0, S0, ro, co, A, m if lay == "C" else n, 0.0, B0, d if lay == "C" else n)
        torch.cuda.synchronize()
        print("fused ok", Dr, Dc, lay, flush=True)
        rb.sketch_general(lay, "N", "N", d, n, m, 1.0, S1, ro, co, A, m if lay == "C" else n, 0.0, B1, d if lay == "C" else n)
        torch.cuda.synchronize()
        print("xmat ok", Dr, Dc, lay, "equal", bool(torch.equal(B0, B1)), "maxdiff", float((B0 - B1).abs().max()), flush=True)
