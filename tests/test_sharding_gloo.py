"""world_size-2 (and 3) CPU tests of the multi-GPU host logic over the gloo backend: the partitioning helpers and the
m-sharded left sketch with its reduce-scatter. The local sketch is injected (the oracle stands in for the CUDA
kernel on this GPU-less box); what is tested is the sharding arithmetic -- co_s offsets, block alignment, the
layout of the reduced slices -- against one unsharded oracle product."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from randblas_b200.sharding import block, sketch_general_mshard  # noqa: E402


def test_block_partition_properties():
    for total in (0, 1, 7, 100, 4000000, 1003):
        for world in (1, 2, 3, 8):
            for align in (1, 4):
                covered = 0
                for r in range(world):
                    s, c = block(total, r, world, align)
                    assert s == covered and c >= 0
                    assert s % align == 0 or s == total
                    covered += c
                assert covered == total
    # near-even: sizes differ by at most one aligned unit
    sizes = [block(4000000, r, 8, 4)[1] for r in range(8)]
    assert max(sizes) - min(sizes) <= 4


def _worker(rank, world, port_no, d, n, m, tmpdir, kind="dense"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_lib as ol
        port = ol.port()
        port.set_threads(1)
        ctr, key = ol.state_from_u64(1997)
        rng = np.random.default_rng(7)
        A = rng.standard_normal((m, n))                       # full A, row-major logical
        start, count = block(m, rank, world, 4)

        class Op:                                             # what the injected sketch needs to know about S
            dist_t = (d, m, "G", "L") if kind == "dense" else (d, m, 3, "S")      # DenseDist / SASO SparseDist, vec_nnz 3

        sketch = port.lskge3 if kind == "dense" else port.lskges

        def local_sketch(layout, opS, opA, d_, n_, m_, alpha, S, ro_s, co_s, A_loc, lda, beta, B, ldb):
            sketch(layout, opS, opA, d_, n_, m_, float(alpha), S.dist_t, ctr, key, ro_s, co_s, A_loc.numpy(), lda,
                   float(beta), B.numpy(), ldb)

        # ColMajor local block: rows [start, start+count) of A, leading dimension = count
        A_loc = torch.from_numpy(np.ascontiguousarray(A[start:start + count, :].T).ravel().copy())
        Bp = torch.zeros(d * n, dtype=torch.float64)
        Bs = torch.zeros(d * n // world, dtype=torch.float64)
        sketch_general_mshard("C", d, n, m, 1.0, Op(), A_loc, max(count, 1), Bp, Bs, rank, world,
                              local_sketch=local_sketch)
        gathered = [torch.zeros_like(Bs) for _ in range(world)]
        dist.all_gather(gathered, Bs)
        if rank == 0:
            got = torch.cat(gathered).numpy()
            want = np.zeros(d * n)
            sketch("C", "N", "N", d, n, m, 1.0, Op.dist_t, ctr, key, 0, 0, np.ascontiguousarray(A.T).ravel(), m, 0.0, want, d)
            err = np.linalg.norm(got - want) / np.linalg.norm(want)
            with open(os.path.join(tmpdir, "result.txt"), "w") as f:
                f.write(repr(float(err)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kind", [(2, "dense"), (3, "dense"), (2, "saso")])
def test_m_sharded_left_sketch_reduce_scatter(world, kind, tmp_path):
    d, n, m = 12, 6 * world, 203                 # d*n divisible by world; m not a multiple of 4 * world
    port_no = 29500 + (os.getpid() % 2000) + world + (7 if kind == "saso" else 0)
    mp.spawn(_worker, args=(world, port_no, d, n, m, str(tmp_path), kind), nprocs=world, join=True)
    err = float(open(tmp_path / "result.txt").read())
    assert err < 1e-12, err
