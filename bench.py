#!/usr/bin/env python
"""Benchmark of the sketching hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--impl ours|reference]

One JSON line on stdout (rank 0). Default workload = BASELINE.json configs[1] (c2): fill_dense, Gaussian double,
8192 x 1,000,000 operator, one full fill per step. Other workloads are the remaining BASELINE.json configs.
Multi-GPU (torchrun, one rank per GPU): the path shards with no data-path collective except c3 (m-sharded
left sketch -> NCCL reduce-scatter of the d x n partials); scaling is "weak" (per-GPU work fixed).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d.get("bf16_tflops", 1590.0)),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", 1400.0)), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def measured_gemm_tflops(torch, dtype, n=8192, reps=5):
    """Library GEMM throughput measured on this GPU in this run: the denominator for the tensor-bound configs
    (MEASURED_PEAKS.json only has bf16). torch.matmul = cuBLAS; TF32 enabled for float32."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, dtype=dtype, device="cuda")
        b = torch.randn(n, n, dtype=dtype, device="cuda")
        c = torch.empty(n, n, dtype=dtype, device="cuda")
        torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        del a, b, c
        return 2.0 * n ** 3 / 1e12 / (best / 1e3)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, dev_index):
        self.dev = dev_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.dev)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(pw)), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- workloads
class Workload:
    name = ""
    metric = ""
    unit = ""
    dtype = ""

    def setup(self, rb, torch, rank, world):
        raise NotImplementedError

    def step(self):                 # one pass of the hot path, device-resident inputs
        raise NotImplementedError

    def units_per_step(self):       # per rank, in `unit` numerator units (samples or bytes of A)
        raise NotImplementedError


class C2FillDense(Workload):
    """fill_dense<double>(DenseDist(8192, 1e6, Gaussian, Long), buff, RNGState(1997)); rank g fills rows
    [8192 g, 8192 (g+1)) of the (8192 * world) x 1e6 operator (same stream, disjoint counters)."""
    name = "c2: fill_dense Gaussian double 8192x1000000 (RowMajor, ld=1e6), one full fill per step"
    metric = "fill_dense Gsamples/s"
    unit = "Gsamples/s"
    dtype = "f64"
    rows, cols = 8192, 1000000

    def setup(self, rb, torch, rank, world):
        self.rb, self.torch, self.rank = rb, torch, rank
        self.D = rb.DenseDist(self.rows * world, self.cols, rb.ScalarDist.Gaussian, rb.Axis.Long)
        self.buf = torch.empty(self.rows * self.cols, dtype=torch.float64, device="cuda")
        self.seed = rb.RNGState(1997)
        self.ro = self.rows * rank

    def step(self):
        self.rb.fill_dense_unpacked("R", self.D, self.rows, self.cols, self.ro, 0, self.buf, self.seed)

    def units_per_step(self):
        return self.rows * self.cols / 1e9

    def roofline(self, kernel_ms, pk):
        gbs = self.rows * self.cols * 8 / 1e9 / (kernel_ms / 1e3)
        return {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                "traffic": self.rows * self.cols * 8 * (1.989540 + 0.000055) / 2.048,
                "traffic_source": "ncu --set full on a 256 x 1e6 slice of this launch (profiles/r01_ncu_prof_fill_gauss_f64"
                                  ".txt): dram write 1.990 GB + read 0.0001 GB for 2.048 GB algorithmic, scaled by rows",
                "kernel": "fill_dense_tiled_kernel<double, GAUSS, 8, 5>", "peak_source": pk["source"],
                "algorithmic_bytes_per_launch": self.rows * self.cols * 8}

    # end to end: host (pinned) destination through the C ABI, D2H inside the timed region
    def e2e_setup(self):
        torch = self.torch
        rows = 1024
        while rows > 16:
            try:
                self.hbuf = torch.empty(rows * self.cols, dtype=torch.float64, pin_memory=True)
                break
            except RuntimeError:
                rows //= 2
        self.e2e_rows = rows
        self.hnp = self.hbuf.numpy()
        return {"sample": f"{rows} x {self.cols} row window of the same operator per step, pinned host destination"}

    def e2e_step(self):
        self.rb.fill_dense_unpacked("R", self.D, self.e2e_rows, self.cols, self.ro, 0, self.hnp, self.seed)

    def e2e_units(self):
        return self.e2e_rows * self.cols / 1e9, 0, self.e2e_rows * self.cols * 8

    # CPU leg: a 64-row slice of the same operator per step (the full output is 65.5 GB)
    cpu_sample = "64 x 1000000 row slice of the operator via fill_dense_unpacked per step"

    def cpu_setup(self, impl, rng):
        self.cpu_i = 0

    def cpu_step(self, impl):
        rows = 64
        impl.fill_dense_unpacked("R", self.rows, self.cols, "G", "L", rows, self.cols, (64 * self.cpu_i) % 8192, 0,
                                 [0, 0, 0, 0], [1997, 0], np.float64)
        self.cpu_i += 1
        return rows * self.cols / 1e9


class C1DenseSketchF32(Workload):
    """sketch_general<float>(ColMajor, N, N, d=1024, n=1024, m=100000, 1, DenseSkOp(DenseDist(1024,100000,Uniform)),
    A, lda=m, 0, B, ldb=d); rank g owns its own 1024 columns of A and B (n-sharded: no communication)."""
    name = "c1: sketch_general float Uniform d=1024 m=100000 n=1024 ColMajor, S unfilled (fused)"
    metric = "sketch_general GB/s of A"
    unit = "GB/s"
    dtype = "f32"
    d, m, n = 1024, 100000, 1024

    def setup(self, rb, torch, rank, world):
        self.rb, self.torch = rb, torch
        self.S = rb.DenseSkOp(rb.DenseDist(self.d, self.m, rb.ScalarDist.Uniform), rb.RNGState(1997), np.float32)
        self.A = torch.empty(self.m * self.n, dtype=torch.float32, device="cuda")
        rb.fill_dense(rb.DenseDist(self.m, self.n), self.A, rb.RNGState(99 + rank))
        self.B = torch.zeros(self.d * self.n, dtype=torch.float32, device="cuda")

    def step(self):
        self.rb.sketch_general("C", "N", "N", self.d, self.n, self.m, 1.0, self.S, 0, 0, self.A, self.m, 0.0, self.B,
                               self.d)

    def units_per_step(self):
        return self.m * self.n * 4 / 1e9

    def roofline(self, kernel_ms, pk):
        tf = 2.0 * self.d * self.m * self.n / 1e12 / (kernel_ms / 1e3)
        tf32 = measured_gemm_tflops(self.torch, self.torch.float32)
        peak = tf32 / 3.0                              # 3xTF32 issues 3 MMAs per fp32 product
        return {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "traffic": None,
                "kernel": "skge3_tc_kernel (tcgen05 3xTF32) + splitk_reduce_kernel",
                "peak_source": f"measured in this run: cuBLAS TF32 GEMM 8192^3 = {tf32:.1f} TFLOP/s, / 3 "
                               f"(MEASURED_PEAKS bf16 {pk['bf16_tflops']:.0f} / 2 / 3 = {pk['bf16_tflops'] / 6:.1f})",
                "algorithmic_flops_per_launch": 2.0 * self.d * self.m * self.n,
                "hbm_gbs_of_A": self.m * self.n * 4 / 1e9 / (kernel_ms / 1e3)}

    def e2e_setup(self):
        torch = self.torch
        self.hA = torch.empty(self.m * self.n, dtype=torch.float32, pin_memory=True)
        self.hA.copy_(self.A)
        self.hB = torch.zeros(self.d * self.n, dtype=torch.float32, pin_memory=True)
        return {"sample": "full config, A and B in pinned host memory"}

    def e2e_step(self):
        self.rb.sketch_general("C", "N", "N", self.d, self.n, self.m, 1.0, self.S, 0, 0, self.hA.numpy(), self.m, 0.0,
                               self.hB.numpy(), self.d)

    def e2e_units(self):
        return self.m * self.n * 4 / 1e9, self.m * self.n * 4, self.d * self.n * 4

    cpu_sample = "full config (the reference materialises the 1024 x 100000 operator, then SGEMM) per step"

    def cpu_setup(self, impl, rng):
        self.cA = rng.standard_normal(self.m * self.n, dtype=np.float32)
        self.cB = np.zeros(self.d * self.n, np.float32)

    def cpu_step(self, impl):
        impl.lskge3("C", "N", "N", self.d, self.n, self.m, np.float32(1), (self.d, self.m, "U", "L"), [0, 0, 0, 0],
                    [1997, 0], 0, 0, self.cA, self.m, np.float32(0), self.cB, self.d)
        return self.m * self.n * 4 / 1e9


class C4SasoApply(Workload):
    """SparseSkOp SASO vec_nnz=8, d=2048, m=8e6, n=256 float RowMajor, operator unsampled (fused generate+apply)."""
    name = "c4: sketch_general float SASO vec_nnz=8 d=2048 m=8000000 n=256 RowMajor (fused fill_sparse + apply)"
    metric = "sketch_general GB/s of A"
    unit = "GB/s"
    dtype = "f32"
    d, m, n, k = 2048, 8000000, 256, 8

    def setup(self, rb, torch, rank, world):
        self.rb, self.torch = rb, torch
        self.S = rb.SparseSkOp(rb.SparseDist(self.d, self.m, self.k), rb.RNGState(1997), dtype=np.float32)
        self.A = torch.empty(self.m * self.n, dtype=torch.float32, device="cuda")
        rb.fill_dense(rb.DenseDist(self.m, self.n), self.A, rb.RNGState(99 + rank))
        self.B = torch.zeros(self.d * self.n, dtype=torch.float32, device="cuda")

    def step(self):
        self.rb.sketch_general("R", "N", "N", self.d, self.n, self.m, 1.0, self.S, 0, 0, self.A, self.n, 0.0, self.B,
                               self.n)

    def units_per_step(self):
        return self.m * self.n * 4 / 1e9

    def roofline(self, kernel_ms, pk):
        gbs = self.m * self.n * 4 / 1e9 / (kernel_ms / 1e3)
        return {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                "traffic": self.m * self.n * 4 * (1.064040 + 0.006197) / 1.024,
                "traffic_source": "ncu --set full on the first 1e6 rows of A (profiles/r01_ncu_prof_saso_binned.txt): dram "
                                  "read 1.064 GB + write 0.006 GB for 1.024 GB algorithmic, scaled by rows",
                "kernel": "saso_bin_kernel<8> + saso_binned_kernel (saso_binned.cu)",
                "peak_source": pk["source"],
                "algorithmic_bytes_per_launch": self.m * self.n * 4}

    def extra(self, pk):
        """fill_sparse alone, int64 COO arrays written to HBM."""
        torch = self.torch
        Sf = self.rb.SparseSkOp(self.rb.SparseDist(self.d, self.m, self.k), self.rb.RNGState(1997), dtype=np.float32)
        self.rb.fill_sparse(Sf)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            self.rb.fill_sparse(Sf)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        nnz = self.k * self.m
        gbs = nnz * 20 / 1e9 / (ms / 1e3)
        del Sf
        return {"fill_sparse": {"ms": ms, "nnz": nnz, "Mnnz_per_s": nnz / ms / 1e3, "bytes_written": nnz * 20,
                                "achieved_GBs": gbs, "frac_hbm": gbs / pk["hbm_gbs"],
                                "kernel": "saso_fill_group_kernel<int64, float, 8>"}}

    def e2e_setup(self):
        torch = self.torch
        self.e2e_m = 1000000
        self.hA = torch.empty(self.e2e_m * self.n, dtype=torch.float32, pin_memory=True)
        self.hA.copy_(self.A[: self.e2e_m * self.n])
        self.hB = torch.zeros(self.d * self.n, dtype=torch.float32, pin_memory=True)
        return {"sample": "first 1,000,000 rows of A (m/8) per step, pinned host A and B"}

    def e2e_step(self):
        self.rb.sketch_general("R", "N", "N", self.d, self.n, self.e2e_m, 1.0, self.S, 0, 0, self.hA.numpy(), self.n,
                               0.0, self.hB.numpy(), self.n)

    def e2e_units(self):
        return self.e2e_m * self.n * 4 / 1e9, self.e2e_m * self.n * 4, self.d * self.n * 4

    cpu_m = 400000
    cpu_sample = ("m/20 = 400000 rows of A with a 2048 x 400000 SASO operator (fill_sparse + COO->CSC sort + apply, "
                  "as the reference does for an unsampled operator) per step")

    def cpu_setup(self, impl, rng):
        self.cA = rng.standard_normal(self.cpu_m * self.n, dtype=np.float32)
        self.cB = np.zeros(self.d * self.n, np.float32)

    def cpu_step(self, impl):
        impl.lskges("R", "N", "N", self.d, self.n, self.cpu_m, np.float32(1), (self.d, self.cpu_m, self.k, "S"),
                    [0, 0, 0, 0], [1997, 0], 0, 0, self.cA, self.n, np.float32(0), self.cB, self.n)
        return self.cpu_m * self.n * 4 / 1e9


class C3DenseSketchF64(Workload):
    """sketch_general<double> Gaussian d=4096 n=512, m-sharded: every rank owns 500,000 rows of A (the 8-GPU shard
    of m = 4,000,000) and the partial products are summed with an NCCL reduce-scatter."""
    name = "c3: sketch_general double Gaussian d=4096 n=512, m-sharded, 500000 rows of A per GPU + reduce-scatter"
    metric = "sketch_general GB/s of A"
    unit = "GB/s"
    dtype = "f64"
    d, n, m_local = 4096, 512, 500000

    def setup(self, rb, torch, rank, world):
        self.rb, self.torch, self.rank, self.world = rb, torch, rank, world
        self.S = rb.DenseSkOp(rb.DenseDist(self.d, self.m_local * world, rb.ScalarDist.Gaussian), rb.RNGState(1997),
                              np.float64)
        self.A = torch.empty(self.m_local * self.n, dtype=torch.float64, device="cuda")
        rb.fill_dense(rb.DenseDist(self.m_local, self.n), self.A, rb.RNGState(99 + rank))
        self.B = torch.zeros(self.d * self.n, dtype=torch.float64, device="cuda")
        self.Bshard = torch.zeros(self.d * self.n // world, dtype=torch.float64, device="cuda")

    def step(self):
        # ColMajor A (lda = m_local): rank g holds rows [g m_local, (g+1) m_local) of A <=> columns co_s.. of S;
        # partial products are summed by one NCCL reduce-scatter (randblas_b200/sharding.py)
        from randblas_b200.sharding import sketch_general_mshard
        sketch_general_mshard("C", self.d, self.n, self.m_local * self.world, 1.0, self.S, self.A, self.m_local, self.B,
                              self.Bshard, self.rank, self.world)

    def units_per_step(self):
        return self.m_local * self.n * 8 / 1e9

    def roofline(self, kernel_ms, pk):
        tf = 2.0 * self.d * self.m_local * self.n / 1e12 / (kernel_ms / 1e3)
        peak = measured_gemm_tflops(self.torch, self.torch.float64, n=6144, reps=3)
        return {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak, "traffic": None,
                "kernel": "skge3_dmma_ws_kernel (mma.sync m8n8k4 f64, warp-specialised) + splitk_reduce_f64_kernel",
                "peak_source": "measured in this run: cuBLAS DGEMM 6144^3 (nominal B200 FP64: 40 TFLOP/s)",
                "algorithmic_flops_per_launch": 2.0 * self.d * self.m_local * self.n}

    def e2e_setup(self):
        # bounded sample: 50,000 rows of this rank's shard, A and B in pinned host memory, no collective
        torch = self.torch
        self.e2e_m = 50000
        hA = self.A.view(self.n, self.m_local)[:, : self.e2e_m].contiguous()       # ColMajor, lda = e2e_m
        self.hA = torch.empty(self.e2e_m * self.n, dtype=torch.float64, pin_memory=True)
        self.hA.copy_(hA.view(-1))
        self.hB = torch.zeros(self.d * self.n, dtype=torch.float64, pin_memory=True)
        return {"sample": "first 50,000 rows of the rank's shard of A (m_local/10) per step, pinned host A and B, "
                          "no reduce-scatter"}

    def e2e_step(self):
        self.rb.sketch_general("C", "N", "N", self.d, self.n, self.e2e_m, 1.0, self.S, 0, 0, self.hA.numpy(),
                               self.e2e_m, 0.0, self.hB.numpy(), self.d)

    def e2e_units(self):
        return self.e2e_m * self.n * 8 / 1e9, self.e2e_m * self.n * 8, self.d * self.n * 8

    cpu_m = 10000
    cpu_sample = ("one row block of 10000 rows of A (the reference's blocked form: operator columns [0, 10000) "
                  "materialised, then DGEMM) per step")

    def cpu_setup(self, impl, rng):
        self.cA = rng.standard_normal(self.cpu_m * self.n)
        self.cB = np.zeros(self.d * self.n, np.float64)

    def cpu_step(self, impl):
        impl.lskge3("C", "N", "N", self.d, self.n, self.cpu_m, 1.0, (self.d, 4000000, "G", "L"), [0, 0, 0, 0],
                    [1997, 0], 0, 0, self.cA, self.cpu_m, 0.0, self.cB, self.d)
        return self.cpu_m * self.n * 8 / 1e9


class C5SketchSparse(Workload):
    """sketch_sparse(ColMajor, N, N, d=512, n, m=1e7, 1, DenseSkOp(DenseDist(512,1e7)), 0,0, CSR A, 0, B, ldb=512),
    synthetic CSR with ~100 nonzeros per row; rank g owns n/8 = 125,000 columns (column-sharded, no comm)."""
    name = "c5: sketch_sparse float CSR 1e7 x 125000 per GPU (column shard of 1e7 x 1e6 at 1e-4 density), d=512"
    metric = "sketch_sparse GB/s of A"
    unit = "GB/s"
    dtype = "f32"
    d, m, n_local, per_row = 512, 10000000, 125000, 12.5

    def setup(self, rb, torch, rank, world):
        self.rb, self.torch = rb, torch
        g = torch.Generator(device="cuda")
        g.manual_seed(1234 + rank)
        lens = torch.poisson(torch.full((self.m,), self.per_row, device="cuda"), generator=g).to(torch.int64)
        lens.clamp_(max=self.n_local)
        rowptr = torch.zeros(self.m + 1, dtype=torch.int64, device="cuda")
        torch.cumsum(lens, 0, out=rowptr[1:])
        nnz = int(rowptr[-1].item())
        row_of = torch.repeat_interleave(torch.arange(self.m, device="cuda"), lens)
        j = torch.arange(nnz, device="cuda") - rowptr[row_of]
        L = lens[row_of].to(torch.float64)
        u = torch.rand(nnz, device="cuda", generator=g, dtype=torch.float64)
        col = torch.floor((j.to(torch.float64) + u) * (self.n_local / L)).to(torch.int64).clamp_(max=self.n_local - 1)
        del row_of, j, L, u
        vals = torch.randn(nnz, device="cuda", generator=g, dtype=torch.float32)
        self.nnz = nnz
        self.A = rb.CSRMatrix(self.m, self.n_local, nnz, vals, rowptr, col)
        self.S = rb.DenseSkOp(rb.DenseDist(self.d, self.m), rb.RNGState(1997), np.float32)
        self.B = torch.zeros(self.d * self.n_local, dtype=torch.float32, device="cuda")

    def bytes_A(self):
        return self.nnz * (4 + 8) + (self.m + 1) * 8

    def step(self):
        self.rb.sketch_sparse("C", "N", "N", self.d, self.n_local, self.m, 1.0, self.S, 0, 0, self.A, 0.0, self.B,
                              self.d)

    def units_per_step(self):
        return self.bytes_A() / 1e9

    def roofline(self, kernel_ms, pk):
        gbs = (self.bytes_A() + self.d * self.n_local * 4) / 1e9 / (kernel_ms / 1e3)
        return {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                "traffic": None, "kernel": "spdata_kgroup_kernel<float> (+ zero-fill of B)", "peak_source": pk["source"],
                "algorithmic_bytes_per_launch": self.bytes_A() + self.d * self.n_local * 4,
                "note": "bound by L2 reductions into B (512 B of red.v4 per nonzero), see DESIGN.md"}

    def e2e_setup(self):
        # bounded sample: the first 1,000,000 rows of the CSR shard, all arrays and B in pinned host memory
        torch = self.torch
        self.e2e_m = 1000000
        rp = self.A.rowptr[: self.e2e_m + 1]
        nnz = int(rp[-1].item())
        pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
        self.h_sp = (pin(rp), pin(self.A.colidxs[:nnz]), pin(self.A.vals[:nnz]))
        self.e2e_nnz = nnz
        self.hB = torch.zeros(self.d * self.n_local, dtype=torch.float32, pin_memory=True)
        self.hAm = self.rb.CSRMatrix(self.e2e_m, self.n_local, nnz, self.h_sp[2].numpy(), self.h_sp[0].numpy(),
                                     self.h_sp[1].numpy())
        return {"sample": "first 1,000,000 rows of the CSR shard (m/10) per step, pinned host CSR arrays and B"}

    def e2e_step(self):
        self.rb.sketch_sparse("C", "N", "N", self.d, self.n_local, self.e2e_m, 1.0, self.S, 0, 0, self.hAm, 0.0,
                              self.hB.numpy(), self.d)

    def e2e_units(self):
        b = self.e2e_nnz * 12 + (self.e2e_m + 1) * 8
        return b / 1e9, b, self.d * self.n_local * 4

    cpu_m = 100000
    cpu_sample = ("row block of 100000 rows of the CSR shard (100000 x 125000, ~12.5 nnz/row) against the matching "
                  "512 x 100000 block of the operator (materialised by the reference, then right_spmm) per step")

    def cpu_setup(self, impl, rng):
        mm, nn = self.cpu_m, self.n_local
        lens = rng.poisson(self.per_row, mm).astype(np.int64)
        rowptr = np.zeros(mm + 1, np.int64)
        np.cumsum(lens, out=rowptr[1:])
        nnz = int(rowptr[-1])
        cols = rng.integers(0, nn, nnz, dtype=np.int64)
        # sort the column indices inside each row
        order = np.lexsort((cols, np.repeat(np.arange(mm), lens)))
        cols = np.ascontiguousarray(cols[order])
        vals = rng.standard_normal(nnz, dtype=np.float32)
        self.c_sp = (mm, nn, nnz, vals, rowptr, cols)
        self.c_bytes = nnz * 12 + (mm + 1) * 8
        self.cB = np.zeros(self.d * nn, np.float32)

    def cpu_step(self, impl):
        impl.lsksp3(0, "C", "N", "N", self.d, self.n_local, self.cpu_m, np.float32(1), (self.d, self.cpu_m, "G", "L"),
                    [0, 0, 0, 0], [1997, 0], 0, 0, self.c_sp, np.float32(0), self.cB, self.d)
        return self.c_bytes / 1e9


WORKLOADS = {"c1": C1DenseSketchF32, "c2": C2FillDense, "c3": C3DenseSketchF64, "c4": C4SasoApply, "c5": C5SketchSparse}


def cpu_impl():
    import oracle_lib as ol
    r = ol.ref()
    if r is not None:
        return r, "reference"
    return ol.port(), "port"


def time_cpu(wl, impl, warmup, steps):
    """Time the CPU implementation of the path on a bounded sample of the workload; returns (units/s, s/step)."""
    rng = np.random.default_rng(99)
    wl.cpu_setup(impl, rng)
    for _ in range(warmup):
        wl.cpu_step(impl)
    tot_units, t0 = 0.0, time.perf_counter()
    for _ in range(steps):
        tot_units += wl.cpu_step(impl)
    dt = time.perf_counter() - t0
    return tot_units / dt, dt / steps


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (oracle/_ref when
    it travelled with the checkout, else the C port), rank 0 only, on a bounded sample of the same workload."""
    if rank != 0:
        return
    impl, kind = cpu_impl()
    cores = os.cpu_count() or 1
    impl.set_threads(cores)
    wl = WORKLOADS[args.workload]()
    value, sec = time_cpu(wl, impl, max(args.warmup, 1), args.steps)
    line = {"impl": "reference", "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
            "config": {"workload": wl.name, "reference_sample": wl.cpu_sample},
            "cpu_baseline": {"value": value, "unit": wl.unit, "cores": impl.get_threads(), "kind": kind,
                             "sample": wl.cpu_sample},
            "e2e": {"value": value, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = {"c1": 20, "c2": 20, "c3": 1, "c4": 10, "c5": 3}[args.workload]
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import randblas_b200 as rb
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for kv in os.environ.get("RB_OPTIONS", "").split(","):      # debug knobs, e.g. RB_OPTIONS=tc_debug=1
        if "=" in kv:
            rb.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    wl = WORKLOADS[args.workload]()
    wl.setup(rb, torch, rank, world)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        wl.step()
    barrier()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = rb.counter("kernel_launches")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        wl.step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    ms_local_per_step = ms / args.steps
    launches = rb.counter("kernel_launches") - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = wl.units_per_step() * world / (ms_per_step / 1e3)

    # dominant-kernel duration: CUDA events around single launches on the launching (current) stream
    kms = []
    for _ in range(min(args.steps, 5)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); wl.step(); b.record(); torch.cuda.synchronize()
        kms.append(a.elapsed_time(b))
    pk = peaks()
    # per-launch duration of the hot kernel(s) of one step = CUDA-event time of the timed region / steps (this rank)
    roof = wl.roofline(ms_local_per_step, pk)
    roof["launch_ms"] = ms_local_per_step
    roof["launch_ms_isolated"] = float(np.mean(kms))      # same step timed alone between two synchronisations

    # end to end through the C ABI with host buffers (copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        info = wl.e2e_setup()
        if info is not None:
            wl.e2e_step()
            barrier()
            n_e2e = max(1, min(args.steps, 3))
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                wl.e2e_step()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / n_e2e
            te = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            units, h2d, d2h = wl.e2e_units()
            e2e = {"value": units * world / float(te.item()), "unit": wl.unit, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "ms_per_step": float(te.item()) * 1e3, **info}

    cpu = None
    if rank == 0 and not args.no_cpu:
        impl, kind = cpu_impl()
        cores = os.cpu_count() or 1
        impl.set_threads(cores)
        n_cpu = {"c1": 3, "c2": 20, "c3": 2, "c4": 3, "c5": 3}[args.workload]
        v, _ = time_cpu(wl, impl, 1, n_cpu)
        cpu = {"value": v, "unit": wl.unit, "cores": impl.get_threads(), "kind": kind,
               "sample": wl.cpu_sample + f" ({n_cpu} steps after 1 warm-up)"}

    if rank == 0:
        line = {"metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
                "config": {"workload": wl.name, "l2": "inputs/outputs larger than the 126 MB L2 (no flush needed)",
                           "sharding": "independent shards per rank" + (" + NCCL reduce-scatter" if args.workload == "c3" else ", no collective")},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
                "kernel_ms": float(np.mean(kms))}
        if hasattr(wl, "extra"):
            line["extra"] = wl.extra(pk)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
