"""Multi-GPU partitioning of the sketching hot path (one process per GPU, torch.distributed for the plumbing).

The path shards without communication wherever the output is partitioned (column blocks of A and B, row blocks
of a filled operator, column blocks of a sparse data matrix): every rank regenerates the part of S it needs from
the shared (key, counter) with (ro_s, co_s) offsets -- the reference's own submatrix semantics (RandBLAS/skge.hh:
174-181; rtd/source/tutorial/sketch_updates.rst:198-213). The only exchange step is the m-sharded left sketch:
rank g holds rows M_g of A, computes the d x n partial S[:, M_g] A[M_g, :], and the partials are summed with one
NCCL reduce-scatter over NVLink. That step lives INSIDE the library (rb_comm_* / rb_lskge3_mshard_*, csrc/comm.cu):
this module only carries the 128-byte NCCL id from rank 0 to the other ranks and forwards the call.
"""
import ctypes

import numpy as np

from . import _lib, api
from ._lib import RandBLASError, call


def block(total, rank, world, align=4):
    """Contiguous block of `total` items owned by `rank`: (start, count). Block starts are multiples of `align`
    (row blocks of A use align=4 so that a Philox block of the operator is never split between ranks). Same
    arithmetic as rb_mshard_block (csrc/comm.cu)."""
    units = (total + align - 1) // align
    per, rem = divmod(units, world)
    u0 = rank * per + min(rank, rem)
    u1 = u0 + per + (1 if rank < rem else 0)
    start, stop = min(u0 * align, total), min(u1 * align, total)
    return start, stop - start


class Comm:
    """Handle of rb_comm_t for this process' GPU. `Comm.from_torch()` bootstraps it from an initialised
    torch.distributed process group (any backend: only the 128-byte id travels through it)."""

    def __init__(self, nranks, rank, unique_id=None):
        self.nranks, self.rank = int(nranks), int(rank)
        self._h = ctypes.c_void_p()
        idbuf = (ctypes.c_char * 128).from_buffer_copy(bytes(unique_id)) if unique_id is not None else None
        call("rb_comm_init_rank", "iipp", self.nranks, self.rank, idbuf if idbuf is not None else 0, self._h)

    @staticmethod
    def unique_id():
        buf = (ctypes.c_char * 128)()
        call("rb_comm_unique_id", "p", buf)
        return bytes(buf)

    @classmethod
    def from_torch(cls, group=None):
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return cls(1, 0)
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world == 1:
            return cls(1, 0)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls(world, rank, box[0])

    def info(self):
        out = (ctypes.c_int64 * 4)()
        call("rb_comm_info", "pp", self._h.value, out)
        return {"nranks": int(out[0]), "rank": int(out[1]), "device": int(out[2]), "nccl_version": int(out[3])}

    def destroy(self):
        if self._h:
            call("rb_comm_destroy", "p", self._h.value)
            self._h = ctypes.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.destroy()
        except Exception:
            pass


def lskge3_mshard(comm, layout, opS, opA, d, n, m_total, alpha, S, ro_s, co_s, A_local, lda, beta, B_out, mode=0):
    """rb_lskge3_mshard_{f32,f64}: B_out = alpha * op(S[ro_s:, co_s:]) * op(A) + beta * B_out with the m_total rows of
    op(A) sharded over the ranks of `comm` (this rank holds rows block(m_total, rank, nranks) as A_local).
    mode 0: reduce-scatter, B_out has d*n/nranks entries (this rank's slice of the packed result);
    mode 1: all-reduce, B_out has d*n entries. Device buffers only; asynchronous on the current stream."""
    if isinstance(S, api.SparseSkOp):
        return lskges_mshard(comm, layout, opS, opA, d, n, m_total, alpha, S, ro_s, co_s, A_local, lda, beta, B_out, mode)
    if not isinstance(S, api.DenseSkOp) or S.buff is not None:
        raise RandBLASError("lskge3_mshard takes an unfilled DenseSkOp (the operator columns are regenerated per rank)")
    dt = api._same_dtype(A_local, B_out)
    sfx, t = api._sfx(dt)
    seed, D = S.seed_state, S.dist
    call(f"rb_lskge3_mshard_{sfx}", "pcccqqq" + t + "qqccpp" + "qqpq" + t + "pip", comm._h.value, layout, opS, opA, int(d), int(n),
         int(m_total), alpha, D.n_rows, D.n_cols, D.family, D.major_axis, seed._c(), seed._k(), int(ro_s), int(co_s),
         api._ptr(A_local), int(lda), beta, api._ptr(B_out), int(mode), api._stream(A_local, B_out))


def lskges_mshard(comm, layout, opS, opA, d, n, m_total, alpha, S, ro_s, co_s, A_local, lda, beta, B_out, mode=0):
    """rb_lskges_mshard_{f32,f64}: lskge3_mshard for an unsampled SASO SparseSkOp (Axis::Short): every rank regenerates the
    minor-axis vectors of S that meet its rows of op(A); same modes, same buffers."""
    if not isinstance(S, api.SparseSkOp) or S.nnz >= 0 or S.dist.major_axis != "S":
        raise RandBLASError("lskges_mshard takes an unsampled SASO SparseSkOp (the operator is regenerated per rank)")
    dt = api._same_dtype(A_local, B_out)
    sfx, t = api._sfx(dt)
    seed, D = S.seed_state, S.dist
    call(f"rb_lskges_mshard_{sfx}", "pcccqqq" + t + "qqqpp" + "qqpq" + t + "pip", comm._h.value, layout, opS, opA, int(d), int(n),
         int(m_total), alpha, D.n_rows, D.n_cols, D.vec_nnz, seed._c(), seed._k(), int(ro_s), int(co_s),
         api._ptr(A_local), int(lda), beta, api._ptr(B_out), int(mode), api._stream(A_local, B_out))


def sketch_general_mshard(layout, d, n, m_total, alpha, S, A_local, lda, B_partial, B_shard, rank, world,
                          local_sketch=None, reduce_scatter=None, comm=None):
    """Left sketch B = alpha * S * A of an A whose rows are sharded: `A_local` holds rows block(m_total, rank, world)
    of A in `layout` with leading dimension `lda`; `B_shard` (d*n / world entries) receives this rank's slice of the
    reduced result (flat slices of the packed B in memory order).

    With `comm` (a Comm) this is one call of rb_lskge3_mshard_* -- the CUDA kernels and the NCCL reduce-scatter
    both run inside librandblas_b200.so and `B_partial` is unused. Without it, `local_sketch` / `reduce_scatter`
    are the injection points the CPU tests use (gloo + oracle) to check the block arithmetic; their defaults are
    randblas_b200.sketch_general and torch.distributed.reduce_scatter_tensor."""
    if (d * n) % world != 0:
        raise RandBLASError(f"(d * n) % world == 0 was required, but did not hold (d={d}, n={n}, world={world})")
    start, count = block(m_total, rank, world, 4)
    if comm is not None:
        lskge3_mshard(comm, layout, "N", "N", d, n, m_total, alpha, S, 0, 0, A_local, lda, 0.0, B_shard, mode=0)
        return start, count
    if local_sketch is None:
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
