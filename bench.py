#!/usr/bin/env python
"""Benchmark of the sketching hot path on B200 (contract: see the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workloads c3,c2,c1,c4,c5] [--impl ours|reference]

ONE JSON line on stdout (rank 0). The default run measures ALL FIVE configurations of BASELINE.json; the top-level
keys (value, ms_per_step, roofline, e2e, cpu_baseline, clocks ...) are those of the HEADLINE configuration

    c3: sketch_general<double>, Gaussian, d=4096, n=512, m=4,000,000 -- the WHOLE problem, strong-scaled:
        rank g of N holds rows block(m, g, N) of A (N=1: all 16.4 GB), regenerates its columns of S on the device
        (never from the host; per K panel into scratch), and the d x n partial products are summed by the NCCL reduce-scatter of rb_lskge3_mshard_f64
        INSIDE the timed region; the result is verified after it (against torch.distributed's own all-reduce of
        independently computed partials, and against the compiled reference on a row block).

and `configs` holds one sub-record per configuration (c1..c5), each with its own value / roofline / e2e /
cpu_baseline / clocks / gpu_launches. --steps / --warmup apply to the headline; the other configurations use the
step counts stated in their sub-record. `--workloads cX[,cY]` restricts the run (the first one named is the
headline). The e2e leg and the CPU leg of one configuration run on the SAME bounded sample of it (in `config`).
`--impl reference`: the reference's own CPU implementation (oracle/_ref) of the same samples, rank 0 only.
"""
import argparse
import gc
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


_GEMM_CACHE = {}


def measured_gemm_tflops(torch, dtype, n=8192, reps=5, sustained_s=0.0):
    """Library GEMM throughput measured on this GPU in this run: the denominator for the tensor-bound configs
    (MEASURED_PEAKS.json only has bf16). torch.matmul = cuBLAS; TF32 enabled for float32. sustained_s > 0: the rate
    of back-to-back GEMMs over that many seconds (the denominator for a kernel timed inside a long step, as
    MEASURED_PEAKS.json does for bf16) instead of the best single launch."""
    if (dtype, n, sustained_s) in _GEMM_CACHE:
        return _GEMM_CACHE[(dtype, n, sustained_s)]
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
        if sustained_s > 0:
            k = max(3, int(sustained_s * 1e3 / best))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(k):
                torch.matmul(a, b, out=c)
            e1.record(); torch.cuda.synchronize()
            best = e0.elapsed_time(e1) / k
        del a, b, c
        _GEMM_CACHE[(dtype, n, sustained_s)] = 2.0 * n ** 3 / 1e12 / (best / 1e3)
        return _GEMM_CACHE[(dtype, n, sustained_s)]
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
        self.nvml = None

    # NVML in-process (a polling thread, first sample taken before start() returns): nvidia-smi needs 0.2-1 s to deliver its
    # first line, longer than the timed region of the short configurations ("no samples" on an 8-GPU box)
    def _start_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        h = None
        try:                                   # the CUDA device's own UUID: immune to device-order differences
            import torch
            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.dev).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
        self.nvml = (pynvml, h)
        self.samples = []
        self._stop = threading.Event()
        self._sample()
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [x.strip() for x in vis.split(",") if x.strip()]
            if self.dev < len(ids) and ids[self.dev].isdigit():
                return int(ids[self.dev])
        return self.dev

    def _sample(self):
        nv, h = self.nvml
        try:
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            self.samples.append((sm, mx, pw, rs))
        except Exception:
            pass

    def _poll(self):
        while not self._stop.wait(0.02):
            self._sample()

    def _stop_nvml(self):
        nv, h = self.nvml
        self._stop.set()
        self.t.join(timeout=1)
        self._sample()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap"))
        reasons = set()
        for _, _, _, rs in self.samples:
            for name, a, b in names:
                bit = getattr(nv, a, None) or getattr(nv, b, 0)
                if bit and (rs & bit):
                    reasons.add(name)
        return {"sm_mhz": float(np.median([x[0] for x in self.samples])), "sm_max_mhz": float(max(x[1] for x in self.samples)),
                "reasons": sorted(reasons), "power_w_max": float(max(x[2] for x in self.samples)), "samples": len(self.samples),
                "source": "nvml"}

    def start(self):
        try:
            self._start_nvml()
            return
        except Exception:
            self.nvml = None
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
        if self.nvml is not None:
            return self._stop_nvml()
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


def bind_to_gpu_numa_node(local):
    """N > 1: pin this rank (and with it the first-touch placement of its pinned host buffers) to the CPU cores
    NVML reports as local to its GPU, so the e2e legs of 8 ranks do not all stream through one NUMA node.
    Returns (the cores used, or None) and the previous mask (restored before the CPU leg)."""
    try:
        old = os.sched_getaffinity(0)
    except Exception:
        return None, None
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        cpus &= old
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus), old
    except Exception:
        pass
    return None, old


COOLDOWN_S = 2.0          # idle time between two configurations of one bench run

# ---------------------------------------------------------------------------------------------- workloads
class Workload:
    key = ""
    name = ""
    metric = ""
    unit = ""
    dtype = ""
    scaling = "weak"
    sharding = "independent shards per rank, no collective"
    default_steps = 10
    cpu_steps = 3
    e2e_steps = 3

    def setup(self, rb, torch, rank, world, comm):
        raise NotImplementedError

    def step(self):                 # one pass of the hot path, device-resident inputs
        raise NotImplementedError

    def units_per_step(self):       # whole-job units per step over all ranks, in `unit` numerator units
        raise NotImplementedError

    def verify(self, dist):         # after the timed region; returns a dict for the record
        return None

    def teardown(self):
        for k in list(self.__dict__):
            if k not in ("rb", "torch"):
                delattr(self, k)


class C3DenseSketchF64(Workload):
    """sketch_general<double> Gaussian d=4096 n=512 m=4,000,000, the whole problem: rank g holds rows
    block(m, g, world) of A (ColMajor, lda = its row count) and calls rb_lskge3_mshard_f64 (per 2 GB K panel of the
    operator: fill kernel into scratch + DMMA kernel; then the NCCL reduce-scatter of the 16.8 MB partials inside the
    library)."""
    key = "c3"
    name = ("c3: sketch_general double Gaussian d=4096 m=4000000 n=512 ColMajor, S unfilled (generated on the fly), m-sharded over the "
            "ranks with the NCCL reduce-scatter inside the timed region (N=1: the whole 16.4 GB A on one GPU)")
    metric = "sketch_general GB/s of A"
    unit = "GB/s"
    dtype = "f64"
    scaling = "strong"
    sharding = "rows of A split over the ranks (block starts multiples of 4) + NCCL reduce-scatter of the d x n partials"
    d, n, m = 4096, 512, 4000000
    sample_m = 50000
    sample = ("a 50,000-row block of A (205 MB) per step against the matching operator columns: e2e = pinned host A and "
              "B through the C ABI (H2D of A, kernels, D2H of B; per rank at N>1); CPU = the reference materialises the "
              "4096 x 50000 operator block, then DGEMM")
    default_steps = 5
    cpu_steps = 2
    e2e_steps = 5

    def setup(self, rb, torch, rank, world, comm):
        from randblas_b200.sharding import block
        self.rb, self.torch, self.rank, self.world, self.comm = rb, torch, rank, world, comm
        self.start, self.count = block(self.m, rank, world, 4)
        self.S = rb.DenseSkOp(rb.DenseDist(self.d, self.m, rb.ScalarDist.Gaussian), rb.RNGState(1997), np.float64)
        self.A = torch.empty(self.count * self.n, dtype=torch.float64, device="cuda")
        # this rank's rows of A: standard normal, the library's own generator (its bits are reproducible on the CPU)
        rb.fill_dense(rb.DenseDist(self.count, self.n), self.A, rb.RNGState(99 + rank))
        self.Bshard = torch.zeros(self.d * self.n // world, dtype=torch.float64, device="cuda")

    def step(self):
        from randblas_b200.sharding import lskge3_mshard
        lskge3_mshard(self.comm, "C", "N", "N", self.d, self.n, self.m, 1.0, self.S, 0, 0, self.A, self.count, 0.0,
                      self.Bshard, mode=0)

    def units_per_step(self):
        return self.m * self.n * 8 / 1e9

    def roofline(self, kernel_ms, pk):
        tf = 2.0 * self.d * self.count * self.n / 1e12 / (kernel_ms / 1e3)
        burst = measured_gemm_tflops(self.torch, self.torch.float64, n=6144, reps=3)
        sustained = measured_gemm_tflops(self.torch, self.torch.float64, n=6144, reps=3, sustained_s=1.5)
        # the step is one ~0.5 s (N=1) launch timed back to back: the sustained DGEMM rate is the matching denominator
        peak = sustained if kernel_ms > 100.0 else burst
        return {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                "traffic": self.count * self.n * 8 * (1017.2 + 1308.4 + 121.1 + 134.2 + 9.2) / 134.217728,
                "traffic_source": "ncu --set full on a 32768-row block (profiles/r02_ncu_summary.txt, dense_f64), per 134.2 MB of "
                                  "A: the fill kernel writes the 1.07 GB operator panel (1017 MB reached DRAM), the DMMA kernel "
                                  "reads 1308 MB (panel + A) and writes 121 MB (split-K partials), the reduce reads 134 MB. The "
                                  "panel is scratch (d/n = 8x the bytes of A at this shape): 0.65 TB/s of DRAM traffic against "
                                  "a tensor-bound kernel; the fused kernel (dmma_materialise=0) moves 240 MB for the same block "
                                  "and is 13% slower",
                "kernel": "skge3_dmma_ws_kernel<materialised operator> (mma.sync m8n8k4 f64, warp-specialised) after "
                          "fill_dense_tiled_kernel per K panel, + splitk_reduce_f64_kernel",
                "peak_source": "measured in this run: cuBLAS DGEMM 6144^3, " + ("back to back for 1.5 s (sustained)"
                               if kernel_ms > 100.0 else "best single launch (burst)") + "; nominal B200 FP64: 40 TFLOP/s",
                "peak_burst": burst, "peak_sustained": sustained, "frac_of_burst": tf / burst,
                "launch_duration": "the whole step: per 2 GB operator panel one fill, one DMMA and one split-K reduce launch "
                                   "(ncu launch list profiles/r02_ncu_launches_bench.csv: DMMA kernel 93.5% of the step, panel "
                                   "fill 6.2%, reduce 0.3%), so `achieved` charges the fill and the reduce to the DMMA kernel",
                "algorithmic_flops_per_launch": 2.0 * self.d * self.count * self.n,
                "hbm_gbs_of_A": self.count * self.n * 8 / 1e9 / (kernel_ms / 1e3)}

    def verify(self, dist):
        """(1) the sharded result against partial products computed by the plain single-GPU entry point and summed by
        torch.distributed's all-reduce; (2) rank 0: a 10,000-row block against the compiled reference."""
        torch, rb = self.torch, self.rb
        out = {}
        Bp = torch.zeros(self.d * self.n, dtype=torch.float64, device="cuda")
        rb.sketch_general("C", "N", "N", self.d, self.n, self.count, 1.0, self.S, 0, self.start, self.A, self.count, 0.0,
                          Bp, self.d)
        if self.world > 1:
            dist.all_reduce(Bp)
        cnt = self.d * self.n // self.world
        ref_slice = Bp[self.rank * cnt:(self.rank + 1) * cnt]
        err = float((self.Bshard - ref_slice).norm() / ref_slice.norm())
        t = torch.tensor([err], dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["sharded_vs_allreduced_partials_relerr"] = float(t.item())
        assert out["sharded_vs_allreduced_partials_relerr"] < 1e-12, out
        if self.rank == 0:
            import oracle_lib as ol
            impl = ol.ref() or ol.port()
            mb, off = 10000, 20000
            off = min(off, max(self.count - mb, 0)) // 4 * 4
            Ablk = self.A.view(self.n, self.count)[:, off:off + mb].contiguous()
            Bd = torch.zeros(self.d * self.n, dtype=torch.float64, device="cuda")
            rb.sketch_general("C", "N", "N", self.d, self.n, mb, 1.0, self.S, 0, self.start + off, Ablk.view(-1), mb, 0.0,
                              Bd, self.d)
            want = np.zeros(self.d * self.n)
            impl.set_threads(os.cpu_count() or 1)
            impl.lskge3("C", "N", "N", self.d, self.n, mb, 1.0, (self.d, self.m, "G", "L"), [0, 0, 0, 0], [1997, 0], 0,
                        self.start + off, Ablk.cpu().numpy().ravel(), mb, 0.0, want, self.d)
            got = Bd.cpu().numpy()
            out["block_vs_" + impl.kind + "_relerr"] = float(np.linalg.norm(got - want) / np.linalg.norm(want))
            assert out["block_vs_" + impl.kind + "_relerr"] < 1e-12, out
        return out

    def e2e_setup(self):
        torch = self.torch
        mm = min(self.sample_m, self.count)
        hA = self.A.view(self.n, self.count)[:, :mm].contiguous()                 # ColMajor, lda = mm
        self.e2e_m = mm
        self.hA = torch.empty(mm * self.n, dtype=torch.float64, pin_memory=True)
        self.hA.copy_(hA.view(-1))
        self.hB = torch.zeros(self.d * self.n, dtype=torch.float64, pin_memory=True)

    def e2e_step(self):
        self.rb.sketch_general("C", "N", "N", self.d, self.n, self.e2e_m, 1.0, self.S, 0, self.start, self.hA.numpy(),
                               self.e2e_m, 0.0, self.hB.numpy(), self.d)

    def e2e_units(self):      # per rank: units, h2d bytes, d2h bytes
        return self.e2e_m * self.n * 8 / 1e9, self.e2e_m * self.n * 8, self.d * self.n * 8

    def cpu_setup(self, impl, rng):
        self.cA = rng.standard_normal(self.sample_m * self.n)
        self.cB = np.zeros(self.d * self.n, np.float64)

    def cpu_step(self, impl):
        impl.lskge3("C", "N", "N", self.d, self.n, self.sample_m, 1.0, (self.d, self.m, "G", "L"), [0, 0, 0, 0],
                    [1997, 0], 0, 0, self.cA, self.sample_m, 0.0, self.cB, self.d)
        return self.sample_m * self.n * 8 / 1e9


class C2FillDense(Workload):
    """fill_dense<double>(DenseDist(8192, 1e6, Gaussian, Long), buff, RNGState(1997)); rank g fills rows
    [8192 g, 8192 (g+1)) of the (8192 * world) x 1e6 operator (same stream, disjoint counters)."""
    key = "c2"
    name = "c2: fill_dense Gaussian double 8192x1000000 (RowMajor, ld=1e6), one full fill per step"
    metric = "fill_dense Gsamples/s"
    unit = "Gsamples/s"
    dtype = "f64"
    rows, cols = 8192, 1000000
    sample_rows = 256
    sample = ("a 256 x 1000000 row window of the operator per step: e2e = pinned host destination through the C ABI "
              "(D2H inside); CPU = fill_dense_unpacked of the same window")
    default_steps = 10
    cpu_steps = 5

    def setup(self, rb, torch, rank, world, comm):
        self.rb, self.torch, self.rank, self.world = rb, torch, rank, world
        self.D = rb.DenseDist(self.rows * world, self.cols, rb.ScalarDist.Gaussian, rb.Axis.Long)
        self.buf = torch.empty(self.rows * self.cols, dtype=torch.float64, device="cuda")
        self.seed = rb.RNGState(1997)
        self.ro = self.rows * rank

    def step(self):
        self.rb.fill_dense_unpacked("R", self.D, self.rows, self.cols, self.ro, 0, self.buf, self.seed)

    def units_per_step(self):
        return self.rows * self.cols * self.world / 1e9

    def roofline(self, kernel_ms, pk):
        gbs = self.rows * self.cols * 8 / 1e9 / (kernel_ms / 1e3)
        return {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                "traffic": self.rows * self.cols * 8 * (1.989540 + 0.000055) / 2.048,
                "traffic_source": "ncu --set full on a 256 x 1e6 slice of this launch (profiles/r02_ncu_summary.txt, "
                                  "fill_gauss_f64): dram write 1.989 GB + read 0.0001 GB for 2.048 GB algorithmic, scaled by rows",
                "kernel": "fill_dense_tiled_kernel<double, GAUSS>", "peak_source": pk["source"],
                "algorithmic_bytes_per_launch": self.rows * self.cols * 8}

    def extra(self, pk):
        """The same kernel and buffer with the Uniform family: the HBM write path without the Box-Muller arithmetic."""
        torch, rb = self.torch, self.rb
        Du = rb.DenseDist(self.rows * self.world, self.cols, rb.ScalarDist.Uniform, rb.Axis.Long)
        rb.fill_dense_unpacked("R", Du, self.rows, self.cols, self.ro, 0, self.buf, self.seed)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            rb.fill_dense_unpacked("R", Du, self.rows, self.cols, self.ro, 0, self.buf, self.seed)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        gbs = self.rows * self.cols * 8 / 1e9 / (ms / 1e3)
        return {"uniform_family": {"ms": ms, "Gsamples_per_s": self.rows * self.cols / 1e9 / (ms / 1e3), "achieved_GBs": gbs,
                                   "frac_hbm": gbs / pk["hbm_gbs"], "kernel": "fill_dense_tiled_kernel<double, UNIFORM>"}}

    def verify(self, dist):
        """a 64 x 5000 corner of this rank's rows against the CPU checker (2 float ulp allowed, 0 expected)"""
        if self.rank != 0:
            return None
        import oracle_lib as ol
        impl = ol.ref() or ol.port()
        want, _ = impl.fill_dense_unpacked("R", self.rows * self.world, self.cols, "G", "L", 64, 5000, self.ro, 0,
                                           [0, 0, 0, 0], [1997, 0], np.float64)
        got = self.buf.view(self.rows, self.cols)[:64, :5000].cpu().numpy().ravel()
        ia = got.astype(np.float32).view(np.int32).astype(np.int64)
        ib = want.astype(np.float32).view(np.int32).astype(np.int64)
        ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
        ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
        ulp = int(np.abs(ia - ib).max())
        assert ulp <= 2, ulp
        return {"corner_vs_" + impl.kind + "_max_float_ulp": ulp}

    def e2e_setup(self):
        torch = self.torch
        self.hbuf = torch.empty(self.sample_rows * self.cols, dtype=torch.float64, pin_memory=True)
        self.hnp = self.hbuf.numpy()

    def e2e_step(self):
        self.rb.fill_dense_unpacked("R", self.D, self.sample_rows, self.cols, self.ro, 0, self.hnp, self.seed)

    def e2e_units(self):
        return self.sample_rows * self.cols / 1e9, 0, self.sample_rows * self.cols * 8

    def cpu_setup(self, impl, rng):
        pass

    def cpu_step(self, impl):
        impl.fill_dense_unpacked("R", self.rows, self.cols, "G", "L", self.sample_rows, self.cols, 0, 0, [0, 0, 0, 0],
                                 [1997, 0], np.float64)
        return self.sample_rows * self.cols / 1e9


class C1DenseSketchF32(Workload):
    """sketch_general<float>(ColMajor, N, N, d=1024, n=1024, m=100000, 1, DenseSkOp(DenseDist(1024,100000,Uniform)),
    A, lda=m, 0, B, ldb=d); rank g owns its own 1024 columns of A and B (n-sharded: no communication)."""
    key = "c1"
    name = "c1: sketch_general float Uniform d=1024 m=100000 n=1024 ColMajor, S unfilled (fused)"
    metric = "sketch_general GB/s of A"
    unit = "GB/s"
    dtype = "f32"
    d, m, n = 1024, 100000, 1024
    sample = ("the full configuration per step: e2e = A and B in pinned host memory through the C ABI; CPU = the "
              "reference materialises the 1024 x 100000 operator, then SGEMM")
    default_steps = 20
    cpu_steps = 2

    def setup(self, rb, torch, rank, world, comm):
        self.rb, self.torch, self.rank, self.world = rb, torch, rank, world
        self.S = rb.DenseSkOp(rb.DenseDist(self.d, self.m, rb.ScalarDist.Uniform), rb.RNGState(1997), np.float32)
        self.A = torch.empty(self.m * self.n, dtype=torch.float32, device="cuda")
        rb.fill_dense(rb.DenseDist(self.m, self.n), self.A, rb.RNGState(99 + rank))
        self.B = torch.zeros(self.d * self.n, dtype=torch.float32, device="cuda")

    def step(self):
        self.rb.sketch_general("C", "N", "N", self.d, self.n, self.m, 1.0, self.S, 0, 0, self.A, self.m, 0.0, self.B,
                               self.d)

    def units_per_step(self):
        return self.m * self.n * 4 * self.world / 1e9

    def roofline(self, kernel_ms, pk):
        tf = 2.0 * self.d * self.m * self.n / 1e12 / (kernel_ms / 1e3)
        tf32 = measured_gemm_tflops(self.torch, self.torch.float32)
        peak = tf32 / 3.0                              # 3xTF32 issues 3 MMAs per fp32 product
        return {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                "traffic": (409.712384 + 307.432192 + 348.154368 + 6.877696) * 1e6,
                "traffic_source": "ncu --set full of this launch (profiles/r02_ncu_summary.txt, dense_f32): main kernel dram read "
                                  "409.7 MB (A once) + write 307.4 MB (83 split-K partial tiles), reduce kernel read 348.2 MB + "
                                  "write 6.9 MB",
                "kernel": "skge3_tc_kernel<PAIR> (tcgen05 cta_group::2, 3xTF32) + splitk_reduce_kernel",
                "peak_source": f"measured in this run: cuBLAS TF32 GEMM 8192^3 = {tf32:.1f} TFLOP/s, / 3 "
                               f"(MEASURED_PEAKS bf16 {pk['bf16_tflops']:.0f} / 2 / 3 = {pk['bf16_tflops'] / 6:.1f})",
                "frac_of_measured_bf16_over_6": tf / (pk["bf16_tflops"] / 6.0),
                "algorithmic_flops_per_launch": 2.0 * self.d * self.m * self.n,
                "hbm_gbs_of_A": self.m * self.n * 4 / 1e9 / (kernel_ms / 1e3)}

    def verify(self, dist):
        """against an fp64 product (cuBLAS DGEMM) of the materialised operator and A: 1e-5 relative Frobenius"""
        torch, rb = self.torch, self.rb
        Sb = torch.empty(self.d * self.m, dtype=torch.float32, device="cuda")
        rb.fill_dense(self.S.dist, Sb, self.S.seed_state)                         # RowMajor d x m
        exact = (Sb.view(self.d, self.m).double() @ self.A.view(self.n, self.m).t().double()).t().contiguous().view(-1)
        err = float((self.B.double() - exact).norm() / exact.norm())
        assert err < 1e-5, err
        return {"vs_fp64_product_of_materialised_operator_relerr": err}

    def e2e_setup(self):
        torch = self.torch
        self.hA = torch.empty(self.m * self.n, dtype=torch.float32, pin_memory=True)
        self.hA.copy_(self.A)
        self.hB = torch.zeros(self.d * self.n, dtype=torch.float32, pin_memory=True)

    def e2e_step(self):
        self.rb.sketch_general("C", "N", "N", self.d, self.n, self.m, 1.0, self.S, 0, 0, self.hA.numpy(), self.m, 0.0,
                               self.hB.numpy(), self.d)

    def e2e_units(self):
        return self.m * self.n * 4 / 1e9, self.m * self.n * 4, self.d * self.n * 4

    def cpu_setup(self, impl, rng):
        self.cA = rng.standard_normal(self.m * self.n, dtype=np.float32)
        self.cB = np.zeros(self.d * self.n, np.float32)

    def cpu_step(self, impl):
        impl.lskge3("C", "N", "N", self.d, self.n, self.m, np.float32(1), (self.d, self.m, "U", "L"), [0, 0, 0, 0],
                    [1997, 0], 0, 0, self.cA, self.m, np.float32(0), self.cB, self.d)
        return self.m * self.n * 4 / 1e9


class C4SasoApply(Workload):
    """SparseSkOp SASO vec_nnz=8, d=2048, m=8e6, n=256 float RowMajor, operator unsampled (fused generate+apply)."""
    key = "c4"
    name = "c4: sketch_general float SASO vec_nnz=8 d=2048 m=8000000 n=256 RowMajor (fused fill_sparse + apply)"
    metric = "sketch_general GB/s of A"
    unit = "GB/s"
    dtype = "f32"
    d, m, n, k = 2048, 8000000, 256, 8
    sample_m = 400000
    sample = ("the first 400,000 rows of A (m/20, 410 MB) with a 2048 x 400000 SASO operator per step: e2e = pinned host A "
              "and B through the C ABI; CPU = fill_sparse + COO->CSC sort + apply, as the reference does for an "
              "unsampled operator")
    default_steps = 10
    cpu_steps = 2

    def setup(self, rb, torch, rank, world, comm):
        self.rb, self.torch, self.rank, self.world = rb, torch, rank, world
        self.S = rb.SparseSkOp(rb.SparseDist(self.d, self.m, self.k), rb.RNGState(1997), dtype=np.float32)
        self.A = torch.empty(self.m * self.n, dtype=torch.float32, device="cuda")
        rb.fill_dense(rb.DenseDist(self.m, self.n), self.A, rb.RNGState(99 + rank))
        self.B = torch.zeros(self.d * self.n, dtype=torch.float32, device="cuda")

    def step(self):
        self.rb.sketch_general("R", "N", "N", self.d, self.n, self.m, 1.0, self.S, 0, 0, self.A, self.n, 0.0, self.B,
                               self.n)

    def units_per_step(self):
        return self.m * self.n * 4 * self.world / 1e9

    def roofline(self, kernel_ms, pk):
        gbs = self.m * self.n * 4 / 1e9 / (kernel_ms / 1e3)
        return {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                "traffic": self.m * self.n * 4 * (1.064040 + 0.006197) / 1.024,
                "traffic_source": "ncu --set full on the first 1e6 rows of A (profiles/r02_ncu_summary.txt, saso_apply): dram "
                                  "read 1.064 GB + write 0.004 GB for 1.024 GB algorithmic, scaled by rows",
                "kernel": "saso_bin_kernel<8> + saso_binned_kernel (saso_binned.cu)",
                "peak_source": pk["source"],
                "algorithmic_bytes_per_launch": self.m * self.n * 4}

    def verify(self, dist):
        """the first 100,000 rows against the CPU checker (the operator's first 100,000 columns are the same SASO
        vectors whatever the total column count: vectors are independent, counters i * vec_nnz)"""
        if self.rank != 0:
            return None
        import oracle_lib as ol
        torch, rb = self.torch, self.rb
        impl = ol.ref() or ol.port()
        mm = 100000
        Bd = torch.zeros(self.d * self.n, dtype=torch.float32, device="cuda")
        rb.sketch_general("R", "N", "N", self.d, self.n, mm, 1.0, self.S, 0, 0, self.A[: mm * self.n], self.n, 0.0, Bd,
                          self.n)
        want = np.zeros(self.d * self.n, np.float32)
        impl.set_threads(os.cpu_count() or 1)
        impl.lskges("R", "N", "N", self.d, self.n, mm, np.float32(1), (self.d, self.m, self.k, "S"), [0, 0, 0, 0],
                    [1997, 0], 0, 0, self.A[: mm * self.n].cpu().numpy(), self.n, np.float32(0), want, self.n)
        got = Bd.cpu().numpy()
        err = float(np.linalg.norm(got.astype(np.float64) - want) / np.linalg.norm(want))
        assert err < 1e-5, err
        return {"first_100000_rows_vs_" + impl.kind + "_relerr": err}

    def extra(self, pk):
        """fill_sparse alone, int64 COO arrays written to HBM."""
        torch = self.torch
        Sf = self.rb.SparseSkOp(self.rb.SparseDist(self.d, self.m, self.k), self.rb.RNGState(1997), dtype=np.float32)
        self.rb.fill_sparse(Sf)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            self.rb.fill_sparse(Sf)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        nnz = self.k * self.m
        gbs = nnz * 20 / 1e9 / (ms / 1e3)
        del Sf
        return {"fill_sparse": {"ms": ms, "nnz": nnz, "Mnnz_per_s": nnz / ms / 1e3, "bytes_written": nnz * 20,
                                "achieved_GBs": gbs, "frac_hbm": gbs / pk["hbm_gbs"],
                                "kernel": "saso_fill_vec_kernel<int64, float, 8>"}}

    def e2e_setup(self):
        torch = self.torch
        self.hA = torch.empty(self.sample_m * self.n, dtype=torch.float32, pin_memory=True)
        self.hA.copy_(self.A[: self.sample_m * self.n])
        self.hB = torch.zeros(self.d * self.n, dtype=torch.float32, pin_memory=True)
        self.S_e2e = self.rb.SparseSkOp(self.rb.SparseDist(self.d, self.sample_m, self.k), self.rb.RNGState(1997),
                                        dtype=np.float32)

    def e2e_step(self):
        self.rb.sketch_general("R", "N", "N", self.d, self.n, self.sample_m, 1.0, self.S_e2e, 0, 0, self.hA.numpy(),
                               self.n, 0.0, self.hB.numpy(), self.n)

    def e2e_units(self):
        return self.sample_m * self.n * 4 / 1e9, self.sample_m * self.n * 4, self.d * self.n * 4

    def cpu_setup(self, impl, rng):
        self.cA = rng.standard_normal(self.sample_m * self.n, dtype=np.float32)
        self.cB = np.zeros(self.d * self.n, np.float32)

    def cpu_step(self, impl):
        impl.lskges("R", "N", "N", self.d, self.n, self.sample_m, np.float32(1), (self.d, self.sample_m, self.k, "S"),
                    [0, 0, 0, 0], [1997, 0], 0, 0, self.cA, self.n, np.float32(0), self.cB, self.n)
        return self.sample_m * self.n * 4 / 1e9


class C5SketchSparse(Workload):
    """sketch_sparse(ColMajor, N, N, d=512, n, m=1e7, 1, DenseSkOp(DenseDist(512,1e7)), 0,0, CSR A, 0, B, ldb=512);
    rank g owns n/8 = 125,000 columns of the 1e7 x 1e6 matrix (column-sharded, no communication). The shard is made
    by the library's random_csr (the reference's random_coo stream, bit for bit, in CSR form) at density 1e-4, so
    the CPU leg sees the same matrix: its first rows are what the reference's random_coo yields for fewer rows."""
    key = "c5"
    name = ("c5: sketch_sparse float CSR 1e7 x 125000 per GPU (column shard of 1e7 x 1e6 at 1e-4 density, random_coo "
            "stream), int64 indices, d=512")
    metric = "sketch_sparse GB/s of A"
    unit = "GB/s"
    dtype = "f32"
    d, m, n_local, density = 512, 10000000, 125000, 1e-4
    sample_m = 200000
    sample = ("the first 200,000 rows of the CSR shard (2.5e6 nonzeros) against the matching 512 x 200000 block of the "
              "operator per step: e2e = pinned host CSR arrays and B through the C ABI; CPU = the reference "
              "materialises the operator block, then right_spmm")
    default_steps = 3
    cpu_steps = 2

    def setup(self, rb, torch, rank, world, comm):
        self.rb, self.torch, self.rank, self.world = rb, torch, rank, world
        self.A, _ = rb.random_csr(self.m, self.n_local, self.density, rb.RNGState(4242 + rank), np.float32, np.int64)
        self.nnz = self.A.nnz
        self.S = rb.DenseSkOp(rb.DenseDist(self.d, self.m), rb.RNGState(1997), np.float32)
        self.B = torch.zeros(self.d * self.n_local, dtype=torch.float32, device="cuda")

    def bytes_A(self):
        return self.nnz * (4 + 8) + (self.m + 1) * 8

    def step(self):
        self.rb.sketch_sparse("C", "N", "N", self.d, self.n_local, self.m, 1.0, self.S, 0, 0, self.A, 0.0, self.B,
                              self.d)

    def units_per_step(self):
        return self.bytes_A() * self.world / 1e9

    def roofline(self, kernel_ms, pk):
        gbs = (self.bytes_A() + self.d * self.n_local * 4) / 1e9 / (kernel_ms / 1e3)
        return {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                "traffic": self.nnz * (4.384612e9 + 7.600831e9) / 12.5e6,
                "traffic_source": "ncu --set full on the first 1e6 rows (1.25e7 nonzeros; profiles/r02_ncu_summary.txt, sksp): dram "
                                  "read 4.38 GB + write 7.60 GB, scaled by nonzeros -- B (256 MB) exceeds the L2, so "
                                  "reductions into non-resident lines spill to DRAM",
                "kernel": "spdata_kgroup_kernel<float> (+ zero-fill of B)", "peak_source": pk["source"],
                "algorithmic_bytes_per_launch": self.bytes_A() + self.d * self.n_local * 4,
                "nnz": self.nnz,
                "l2_reduction_rate_G_red_v4_per_s": self.nnz * (self.d / 4) / 1e9 / (kernel_ms / 1e3),
                "note": "bound by L2 reductions into B (d/4 red.v4 per nonzero; measured ceiling ~370 G/s), see DESIGN.md"}

    def verify(self, dist):
        """the first 50,000 rows against the CPU checker with the matching operator window"""
        if self.rank != 0:
            return None
        import oracle_lib as ol
        torch, rb = self.torch, self.rb
        impl = ol.ref() or ol.port()
        mm = 50000
        rp = self.A.rowptr[: mm + 1].contiguous()
        nnz = int(rp[-1].item())
        sub = rb.CSRMatrix(mm, self.n_local, nnz, self.A.vals[:nnz], rp, self.A.colidxs[:nnz])
        Bd = torch.zeros(self.d * self.n_local, dtype=torch.float32, device="cuda")
        rb.sketch_sparse("C", "N", "N", self.d, self.n_local, mm, 1.0, self.S, 0, 0, sub, 0.0, Bd, self.d)
        want = np.zeros(self.d * self.n_local, np.float32)
        impl.set_threads(os.cpu_count() or 1)
        spA = (mm, self.n_local, nnz, self.A.vals[:nnz].cpu().numpy(), rp.cpu().numpy(), self.A.colidxs[:nnz].cpu().numpy())
        impl.lsksp3(0, "C", "N", "N", self.d, self.n_local, mm, np.float32(1), (self.d, self.m, "G", "L"), [0, 0, 0, 0],
                    [1997, 0], 0, 0, spA, np.float32(0), want, self.d)
        got = Bd.cpu().numpy()
        err = float(np.linalg.norm(got.astype(np.float64) - want) / np.linalg.norm(want))
        assert err < 1e-5, err
        return {"first_50000_rows_vs_" + impl.kind + "_relerr": err, "generator_ambiguous_skips": self.A.ambiguous}

    def e2e_setup(self):
        torch = self.torch
        rp = self.A.rowptr[: self.sample_m + 1]
        nnz = int(rp[-1].item())
        pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t)
        self.h_sp = (pin(rp), pin(self.A.colidxs[:nnz]), pin(self.A.vals[:nnz]))
        self.e2e_nnz = nnz
        self.hB = torch.zeros(self.d * self.n_local, dtype=torch.float32, pin_memory=True)
        self.hAm = self.rb.CSRMatrix(self.sample_m, self.n_local, nnz, self.h_sp[2].numpy(), self.h_sp[0].numpy(),
                                     self.h_sp[1].numpy())

    def e2e_step(self):
        self.rb.sketch_sparse("C", "N", "N", self.d, self.n_local, self.sample_m, 1.0, self.S, 0, 0, self.hAm, 0.0,
                              self.hB.numpy(), self.d)

    def e2e_units(self):
        b = self.e2e_nnz * 12 + (self.sample_m + 1) * 8
        return b / 1e9, b, self.d * self.n_local * 4

    def cpu_setup(self, impl, rng):
        # the same matrix as the GPU legs: the reference's own random_coo for the first sample_m rows, as CSR
        mm, nn = self.sample_m, self.n_local
        if hasattr(impl, "random_sparse"):
            v, r, c, nnz, _ = impl.random_sparse(2, mm, nn, self.density, [0, 0, 0, 0], [4242, 0], np.float32)
            vals, cols, rowptr = impl.coo_to_compressed(0, mm, nn, v, r, c)
        else:                                   # C port without the generator: an iid pattern of the same density
            lens = rng.binomial(nn, self.density, mm).astype(np.int64)
            rowptr = np.zeros(mm + 1, np.int64)
            np.cumsum(lens, out=rowptr[1:])
            nnz = int(rowptr[-1])
            cols = rng.integers(0, nn, nnz, dtype=np.int64)
            order = np.lexsort((cols, np.repeat(np.arange(mm), lens)))
            cols = np.ascontiguousarray(cols[order])
            vals = rng.standard_normal(nnz, dtype=np.float32)
        self.c_sp = (mm, nn, nnz, vals, rowptr, cols)
        self.c_bytes = nnz * 12 + (mm + 1) * 8
        self.cB = np.zeros(self.d * nn, np.float32)

    def cpu_step(self, impl):
        impl.lsksp3(0, "C", "N", "N", self.d, self.n_local, self.sample_m, np.float32(1), (self.d, self.m, "G", "L"),
                    [0, 0, 0, 0], [1997, 0], 0, 0, self.c_sp, np.float32(0), self.cB, self.d)
        return self.c_bytes / 1e9


WORKLOADS = {"c1": C1DenseSketchF32, "c2": C2FillDense, "c3": C3DenseSketchF64, "c4": C4SasoApply, "c5": C5SketchSparse}
DEFAULT_ORDER = ["c3", "c2", "c1", "c4", "c5"]


def config_of(wl_cls):
    """The `config` object of a record: identical in the ours arm and in the reference arm."""
    return {"workload": wl_cls.name, "sample": wl_cls.sample,
            "l2": "inputs/outputs larger than the 126 MB L2 (no flush needed)", "sharding": wl_cls.sharding}


def cpu_impl():
    import oracle_lib as ol
    r = ol.ref()
    if r is not None:
        return r, "reference"
    return ol.port(), "port"


def time_cpu(wl, impl, warmup, steps):
    """Time the CPU implementation of the path on the bounded sample of the workload; returns (units/s, s/step)."""
    rng = np.random.default_rng(99)
    wl.cpu_setup(impl, rng)
    for _ in range(warmup):
        wl.cpu_step(impl)
    tot_units, t0 = 0.0, time.perf_counter()
    for _ in range(steps):
        tot_units += wl.cpu_step(impl)
    dt = time.perf_counter() - t0
    return tot_units / dt, dt / steps


def cpu_record(wl, impl, kind, warmup, steps):
    v, sec = time_cpu(wl, impl, warmup, steps)
    return {"value": v, "unit": wl.unit, "cores": impl.get_threads(), "kind": kind,
            "sample": wl.sample + f" ({steps} steps after {warmup} warm-up)", "ms_per_step": sec * 1e3}


def run_reference_arm(args, rank, world, keys):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (oracle/_ref when
    it travelled with the checkout, else the C port), rank 0 only, on the bounded sample of every workload."""
    if rank != 0:
        return
    impl, kind = cpu_impl()
    cores = os.cpu_count() or 1
    impl.set_threads(cores)
    records = {}
    for i, key in enumerate(keys):
        wl = WORKLOADS[key]()
        warm, steps = (max(args.warmup, 1), args.steps) if i == 0 else (1, wl.cpu_steps)
        cpu = cpu_record(wl, impl, kind, warm, steps)
        records[key] = {"impl": "reference", "metric": wl.metric, "value": cpu["value"], "unit": wl.unit, "steps": steps,
                        "warmup": warm, "ms_per_step": cpu["ms_per_step"], "dtype": wl.dtype, "config": config_of(type(wl)),
                        "cpu_baseline": cpu,
                        "e2e": {"value": cpu["value"], "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        wl.teardown()
    head = records[keys[0]]
    line = {"impl": "reference", "metric": head["metric"], "value": head["value"], "unit": head["unit"],
            "n_gpus": args.gpus, "steps": head["steps"], "warmup": head["warmup"], "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": WORKLOADS[keys[0]].scaling, "vs_baseline": None, "dtype": head["dtype"],
            "data": "synthetic", "config": head["config"], "cpu_baseline": head["cpu_baseline"], "e2e": head["e2e"],
            "gpu_launches": 0, "configs": records}
    print(json.dumps(line))


def run_workload(wl, rb, torch, dist, rank, world, local, comm, steps, warmup, args, old_affinity):
    """All legs of one configuration; returns its record (rank 0) or None."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_setup = time.perf_counter()
    wl.setup(rb, torch, rank, world, comm)
    warm = max(warmup, 3)
    for _ in range(warm):
        wl.step()
    barrier()
    setup_s = time.perf_counter() - t_setup

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = rb.counter("kernel_launches")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        wl.step()
    ev1.record()
    barrier()
    ms_local = ev0.elapsed_time(ev1)
    launches = rb.counter("kernel_launches") - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / steps
    value = wl.units_per_step() / (ms_per_step / 1e3)

    check = wl.verify(dist)

    # the same step timed alone between two synchronisations (cross-check of the per-launch duration)
    kms = []
    for _ in range(min(steps, 2)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); wl.step(); b.record(); torch.cuda.synchronize()
        kms.append(a.elapsed_time(b))
    pk = peaks()
    roof = wl.roofline(ms_local / steps, pk)       # this rank's per-launch duration = event time of the region / steps
    roof["launch_ms"] = ms_local / steps
    roof["launch_ms_isolated"] = float(np.mean(kms))
    extra = wl.extra(pk) if hasattr(wl, "extra") and rank == 0 else None

    # end to end through the C ABI with host buffers (copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        wl.e2e_setup()
        wl.e2e_step()
        barrier()
        n_e2e = wl.e2e_steps
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            wl.e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        te = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        units, h2d, d2h = wl.e2e_units()
        ceiling = None
        if world > 1:
            # what the host side allows when all ranks copy at once: a raw pinned-memory copy of the step's dominant
            # transfer (same direction, up to 512 MB), all ranks concurrently -- the ceiling the e2e leg can reach
            nb = int(min(max(h2d, d2h), 512 << 20))
            hb = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
            db = torch.empty(nb, dtype=torch.uint8, device="cuda")
            src, dst = (hb, db) if h2d >= d2h else (db, hb)
            dst.copy_(src, non_blocking=True)
            barrier()
            tc0 = time.perf_counter()
            for _ in range(3):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            tc = torch.tensor([(time.perf_counter() - tc0) / 3], dtype=torch.float64, device="cuda")
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            ceiling = {"direction": "h2d" if h2d >= d2h else "d2h", "bytes": nb,
                       "gbs_per_rank_all_ranks_concurrent": nb / 1e9 / float(tc.item())}
            del hb, db
        e2e = {"value": units * world / float(te.item()), "unit": wl.unit, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": float(te.item()) * 1e3, "steps": n_e2e,
               "pcie_gbs_per_rank": (h2d + d2h) / 1e9 / float(te.item()), "raw_copy_ceiling": ceiling,
               "sample": wl.sample + (" (one such sample per rank)" if world > 1 else "")}

    cpu = None
    if rank == 0 and not args.no_cpu:
        if old_affinity:
            try:
                os.sched_setaffinity(0, old_affinity)      # the CPU leg uses every host core
            except Exception:
                pass
        impl, kind = cpu_impl()
        impl.set_threads(os.cpu_count() or 1)
        cpu = cpu_record(wl, impl, kind, 1, wl.cpu_steps)

    rec = None
    if rank == 0:
        rec = {"metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": world, "steps": steps, "warmup": warm,
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None,
               "dtype": wl.dtype, "data": "synthetic", "config": config_of(type(wl)), "clocks": clocks, "e2e": e2e,
               "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu, "verified": check,
               "setup_s": round(setup_s, 2)}
        if extra:
            rec["extra"] = extra
    wl.teardown()
    gc.collect()
    torch.cuda.empty_cache()
    rb._lib.lib().rb_release_workspace()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workloads", "--workload", dest="workloads", default=",".join(DEFAULT_ORDER),
                    help="comma-separated subset of c1..c5; the first one is the headline of the JSON line")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    keys = [k.strip() for k in args.workloads.split(",") if k.strip()]
    for k in keys:
        if k not in WORKLOADS:
            ap.error(f"unknown workload {k}")
    if args.steps is None:
        args.steps = WORKLOADS[keys[0]].default_steps
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world, keys)
        return

    import torch
    import torch.distributed as dist
    import randblas_b200 as rb
    from randblas_b200.sharding import Comm
    torch.cuda.set_device(local)
    numa_cpus, old_affinity = (None, None)
    if world > 1:
        numa_cpus, old_affinity = bind_to_gpu_numa_node(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Comm.from_torch()

    for kv in os.environ.get("RB_OPTIONS", "").split(","):      # debug knobs, e.g. RB_OPTIONS=tc_debug=1
        if "=" in kv:
            rb.set_option(kv.split("=")[0], int(kv.split("=")[1]))

    records = {}
    t_all = time.perf_counter()
    for i, key in enumerate(keys):
        if i > 0:
            # the configurations are independent measurements: let the power / clock state of the previous one settle (measured:
            # C1 right after the power-capped C2 runs at 1.05 ms with SM clocks still at 1710-1856 MHz, alone at 1.01 ms)
            torch.cuda.synchronize()
            time.sleep(COOLDOWN_S)
        wl = WORKLOADS[key]()
        steps = args.steps if i == 0 else wl.default_steps
        rec = run_workload(wl, rb, torch, dist, rank, world, local, comm, steps, args.warmup, args, old_affinity)
        if world > 1 and numa_cpus:
            try:
                os.sched_setaffinity(0, set(numa_cpus))
            except Exception:
                pass
        if rec is not None:
            records[key] = rec

    if rank == 0:
        line = dict(records[keys[0]])
        line["comm"] = comm.info()
        line["numa_cpus_rank0"] = f"{numa_cpus[0]}-{numa_cpus[-1]} ({len(numa_cpus)})" if numa_cpus else None
        line["total_run_s"] = round(time.perf_counter() - t_all, 1)
        line["configs"] = records
        print(json.dumps(line))
    comm.destroy()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
