"""Worker of test_mshard_two_ranks_cuda_nccl_vs_single_gpu_and_reference: run under torchrun with >= 2 ranks, one GPU
each. Checks rb_lskge3_mshard_* (CUDA kernels + NCCL collective inside librandblas_b200.so) against the single-GPU
sketch of the whole matrix and against the compiled reference. Prints MSHARD_OK from rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import oracle_lib as ol  # noqa: E402
import randblas_b200 as rb  # noqa: E402
from randblas_b200.sharding import Comm, block, lskge3_mshard  # noqa: E402


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Comm.from_torch()
    info = comm.info()
    assert info["nranks"] == world and info["rank"] == rank and info["nccl_version"] > 0, info
    ref = ol.ref() if rank == 0 else None
    ctr, key = ol.state_from_u64(1997)
    worst = {}
    # (dtype, family, d, n, m, layout, tolerance): small ragged shapes (generic + tensor-core paths), then a block of
    # BASELINE.json configs[2] (d=4096, n=512) with 20000 rows per rank
    # the first case leaves the last rank(s) WITHOUT rows (m < 4 * world): their partial is zero, the collective still runs
    cases = [(np.float64, "G", 8, 6 * world, 3, "C", 1e-12),
             (np.float64, "G", 256, 96, 20006, "C", 1e-12), (np.float32, "U", 128, 64, 9001, "R", 1e-5),
             (np.float32, "G", 256, 256, 16384, "C", 1e-5), (np.float64, "G", 4096, 512, 20000 * world, "C", 1e-12)]
    for (dt, fam, d, n, m, layout, tol) in cases:
        tdt = torch.float32 if dt == np.float32 else torch.float64
        # the full A (logical m x n), identical on every rank: the library's own generator with a fixed seed, written in
        # ColMajor order whatever the distribution's natural layout is
        Afull = torch.empty(m * n, dtype=tdt, device="cuda")
        rb.fill_dense_unpacked("C", rb.DenseDist(m, n), m, n, 0, 0, Afull, rb.RNGState(99))     # ColMajor, lda = m
        A2 = Afull.view(n, m).t()                                            # logical m x n
        start, count = block(m, rank, world, 4)
        if layout == "C":
            Aloc = A2[start:start + count, :].t().contiguous().view(-1)      # ColMajor block, lda = count
            lda_loc, lda_full, Ause = max(count, 1), m, Afull
            ldb = d
        else:
            Aloc = A2[start:start + count, :].contiguous().view(-1)          # RowMajor block, lda = n
            lda_loc, lda_full, Ause = n, n, A2.contiguous().view(-1)
            ldb = n
        S = rb.DenseSkOp(rb.DenseDist(d, m, fam), rb.RNGState(1997), dt)
        # single-GPU result of the whole problem (every rank computes it: the comparison baseline)
        Bone = torch.zeros(d * n, dtype=tdt, device="cuda")
        rb.sketch_general(layout, "N", "N", d, n, m, 1.0, S, 0, 0, Ause, lda_full, 0.0, Bone, ldb)
        # mode 0: reduce-scatter
        cnt = d * n // world
        Bshard = torch.full((cnt,), float("nan"), dtype=tdt, device="cuda")
        lskge3_mshard(comm, layout, "N", "N", d, n, m, 1.0, S, 0, 0, Aloc, lda_loc, 0.0, Bshard, mode=0)
        torch.cuda.synchronize()
        e0 = relerr(Bshard.cpu().numpy(), Bone[rank * cnt:(rank + 1) * cnt].cpu().numpy())
        # mode 1: all-reduce with beta != 0
        B0 = torch.arange(d * n, dtype=tdt, device="cuda") / (d * n)
        Ball = B0.clone()
        lskge3_mshard(comm, layout, "N", "N", d, n, m, 2.0, S, 0, 0, Aloc, lda_loc, -0.5, Ball, mode=1)
        torch.cuda.synchronize()
        e1 = relerr(Ball.cpu().numpy(), (2.0 * Bone - 0.5 * B0).cpu().numpy())
        # mode 0 with beta != 0
        Bs2 = B0[rank * cnt:(rank + 1) * cnt].clone()
        lskge3_mshard(comm, layout, "N", "N", d, n, m, 1.0, S, 0, 0, Aloc, lda_loc, 3.0, Bs2, mode=0)
        torch.cuda.synchronize()
        e2 = relerr(Bs2.cpu().numpy(), (Bone + 3.0 * B0)[rank * cnt:(rank + 1) * cnt].cpu().numpy())
        errs = torch.tensor([e0, e1, e2], dtype=torch.float64, device="cuda")
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        e_ref = 0.0
        if rank == 0 and m * d * n <= 3e11:
            gathered = Ball.cpu().numpy()
            want = B0.cpu().numpy().copy()
            ref.set_threads(os.cpu_count() or 1)
            ref.lskge3(layout, "N", "N", d, n, m, dt(2.0), (d, m, fam, "L"), ctr, key, 0, 0, Ause.cpu().numpy(), lda_full,
                       dt(-0.5), want, ldb)
            e_ref = relerr(gathered, want)
        worst[(np.dtype(dt).name, fam, d, n, m, layout)] = (float(errs.max().item()), e_ref)
        assert float(errs.max().item()) < tol, (dt, d, n, m, layout, errs.tolist())
        assert e_ref < tol, (dt, d, n, m, layout, e_ref)
    # SASO operators through rb_lskges_mshard_* ("SASO apply, m-sharded"): a ragged small case and a block of
    # BASELINE.json configs[3] (d=2048, n=256, vec_nnz 8) with 100000 rows per rank
    for (dt, d, n, m, k, layout, tol) in [(np.float64, 64, 12 * world, 4099, 4, "C", 1e-12),
                                            (np.float32, 2048, 256, 100000 * world, 8, "R", 1e-5)]:
        tdt = torch.float32 if dt == np.float32 else torch.float64
        Afull = torch.empty(m * n, dtype=tdt, device="cuda")
        rb.fill_dense_unpacked("C", rb.DenseDist(m, n), m, n, 0, 0, Afull, rb.RNGState(99))     # ColMajor, lda = m
        A2 = Afull.view(n, m).t()
        start, count = block(m, rank, world, 4)
        if layout == "C":
            Aloc = A2[start:start + count, :].t().contiguous().view(-1)
            lda_loc, lda_full, Ause, ldb = max(count, 1), m, Afull, d
        else:
            Aloc = A2[start:start + count, :].contiguous().view(-1)
            lda_loc, lda_full, Ause, ldb = n, n, A2.contiguous().view(-1), n
        S = rb.SparseSkOp(rb.SparseDist(d, m, k), rb.RNGState(1997), dtype=dt)
        Bone = torch.zeros(d * n, dtype=tdt, device="cuda")
        rb.sketch_general(layout, "N", "N", d, n, m, 1.0, S, 0, 0, Ause, lda_full, 0.0, Bone, ldb)
        cnt = d * n // world
        Bshard = torch.full((cnt,), float("nan"), dtype=tdt, device="cuda")
        lskge3_mshard(comm, layout, "N", "N", d, n, m, 1.0, S, 0, 0, Aloc, lda_loc, 0.0, Bshard, mode=0)
        B0 = torch.arange(d * n, dtype=tdt, device="cuda") / (d * n)
        Ball = B0.clone()
        lskge3_mshard(comm, layout, "N", "N", d, n, m, 2.0, S, 0, 0, Aloc, lda_loc, -0.5, Ball, mode=1)
        torch.cuda.synchronize()
        e0 = relerr(Bshard.cpu().numpy(), Bone[rank * cnt:(rank + 1) * cnt].cpu().numpy())
        e1 = relerr(Ball.cpu().numpy(), (2.0 * Bone - 0.5 * B0).cpu().numpy())
        errs = torch.tensor([e0, e1], dtype=torch.float64, device="cuda")
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        e_ref = 0.0
        if rank == 0:
            want = B0.cpu().numpy().copy()
            ref.lskges(layout, "N", "N", d, n, m, dt(2.0), (d, m, k, "S"), ctr, key, 0, 0, Ause.cpu().numpy(), lda_full,
                       dt(-0.5), want, ldb)
            e_ref = relerr(Ball.cpu().numpy(), want)
        worst[("saso", np.dtype(dt).name, d, n, m, k, layout)] = (float(errs.max().item()), e_ref)
        assert float(errs.max().item()) < tol, ("saso", dt, d, n, m, errs.tolist())
        assert e_ref < tol, ("saso", dt, d, n, m, e_ref)
        assert S.nnz < 0
    dist.barrier()
    if rank == 0:
        for k, v in worst.items():
            print(f"mshard {k}: vs single-GPU {v[0]:.2e}, vs reference {v[1]:.2e}")
        print("MSHARD_OK", f"world={world}", f"nccl={info['nccl_version']}")
    comm.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
