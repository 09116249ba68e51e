"""Multi-GPU partitioning of the sketching hot path (one process per GPU, torch.distributed for the plumbing).

The path shards without communication wherever the output is partitioned (column blocks of A and B, row blocks
of a filled operator, column blocks of a sparse data matrix): every rank regenerates the part of S it needs from
the shared (key, counter) with (ro_s, co_s) offsets -- the reference's own submatrix semantics (RandBLAS/skge.hh:
174-181; rtd/source/tutorial/sketch_updates.rst:198-213). The only exchange step is the m-sharded left sketch:
rank g holds rows M_g of A, computes the d x n partial S[:, M_g] A[M_g, :], and the partials are summed with one
reduce-scatter over NVLink.
"""


def block(total, rank, world, align=1):
    """Contiguous block of `total` items owned by `rank`: (start, count). Block starts are multiples of `align`
    (row blocks of A use align=4 so that a Philox block of the operator is never split between ranks)."""
    units = (total + align - 1) // align
    per, rem = divmod(units, world)
    u0 = rank * per + min(rank, rem)
    u1 = u0 + per + (1 if rank < rem else 0)
    start, stop = min(u0 * align, total), min(u1 * align, total)
    return start, stop - start


def sketch_general_mshard(layout, d, n, m_total, alpha, S, A_local, lda, B_partial, B_shard, rank, world,
                          local_sketch=None, reduce_scatter=None):
    """Left sketch B = alpha * S * A of an A whose rows are sharded: `A_local` holds rows block(m_total, rank, world, 4)
    of A in `layout` with leading dimension `lda`. `B_partial` (d*n) receives this rank's partial product, `B_shard`
    (d*n / world) this rank's slice of the reduced result (flat slices of B in memory order).

    local_sketch / reduce_scatter are injection points for the CPU tests (gloo + oracle); the defaults are the CUDA
    path (randblas_b200.sketch_general) and torch.distributed.reduce_scatter_tensor (NCCL)."""
    start, count = block(m_total, rank, world, 4)
    if local_sketch is None:
        from . import api
        local_sketch = api.sketch_general
    ldb = d if layout == "C" else n
    # operator columns [start, start + count) <-> rows of A owned here
    local_sketch(layout, "N", "N", d, n, count, alpha, S, 0, start, A_local, lda, 0.0, B_partial, ldb)
    if world > 1:
        if reduce_scatter is None:
            import torch.distributed as dist
            reduce_scatter = dist.reduce_scatter_tensor
        reduce_scatter(B_shard, B_partial)
    else:
        B_shard.copy_(B_partial) if hasattr(B_shard, "copy_") else B_shard.__setitem__(slice(None), B_partial)
    return start, count
