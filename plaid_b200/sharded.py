"""Column-sharded scoring: one process (rank) per GPU, samples split across ranks.

The product, the ranking and the per-column medians are column-independent (the axis
`chunked_crossprod` already splits on, R/plaid.R:115-119), so there is NO data-path
collective.  What does couple the shards are a handful of scalars:
  * min / max of X                      (replaid.scse auto removeLog2, R/plaid.R:160-161)
  * max(rX)                             (ssgsea / ucell / aucell, R/plaid.R:251,278,306)
  * min(scores) and mean(col medians)   (normalize_medians, R/plaid.R:557,572)
They are exchanged with torch.distributed (NCCL on GPUs, gloo in the CPU tests): all-reduce
(min / max) of single doubles and an all-gather of the per-column medians, which every rank
then combines in GLOBAL COLUMN ORDER (plaidgpu_combine_medians), so results are bit-identical
for any number of shards.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib as L


class LocalComm:
    """world_size == 1."""
    rank, world = 0, 1

    def allreduce_min(self, v: float) -> float:
        return v

    def allreduce_max(self, v: float) -> float:
        return v

    def allgather_vec(self, v: np.ndarray) -> np.ndarray:
        return v


class TorchComm:
    """torch.distributed adapter (backend nccl -> tensors on the rank's GPU; gloo -> CPU)."""

    def __init__(self, device=None, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device if device is not None else "cpu"

    def _red(self, v: float, op):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op, group=self.group)
        return float(t.item())

    def allreduce_min(self, v: float) -> float:
        return self._red(v, self.dist.ReduceOp.MIN)

    def allreduce_max(self, v: float) -> float:
        return self._red(v, self.dist.ReduceOp.MAX)

    def allgather_vec(self, v: np.ndarray) -> np.ndarray:
        """concatenate variable-length float64 vectors of all ranks in rank order"""
        torch, dist = self.torch, self.dist
        n = torch.tensor([v.size], dtype=torch.int64, device=self.device)
        sizes = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(sizes, n, group=self.group)
        sizes = [int(s.item()) for s in sizes]
        m = max(sizes) if sizes else 0
        buf = torch.zeros(max(m, 1), dtype=torch.float64, device=self.device)
        if v.size:
            buf[:v.size] = torch.from_numpy(np.ascontiguousarray(v)).to(self.device)
        parts = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(parts, buf, group=self.group)
        return np.concatenate([p[:s].cpu().numpy() for p, s in zip(parts, sizes)]) if sizes else v


def combine_scalars(comm, local: L.Scalars) -> L.Scalars:
    """step 2 of the protocol in include/plaidgpu.h: x_min (min), x_max (max), rank_max (max)."""
    g = L.Scalars()
    C.memmove(C.byref(g), C.byref(local), C.sizeof(L.Scalars))
    g.x_min = comm.allreduce_min(local.x_min)
    g.x_max = comm.allreduce_max(local.x_max)
    g.rank_max = comm.allreduce_max(local.rank_max)
    return g


def combine_medians(lib, comm, ignore_zero_opt: int, scal: L.Scalars, med_all: np.ndarray, med_nz: np.ndarray):
    """step 4: global score_min, all medians in column order -> ignore_zero flag + mean(medx)."""
    smin = comm.allreduce_min(scal.score_min)
    ga = np.ascontiguousarray(comm.allgather_vec(med_all), dtype=np.float64)
    gz = np.ascontiguousarray(comm.allgather_vec(med_nz), dtype=np.float64)
    rc = lib.plaidgpu_combine_medians(int(ignore_zero_opt), smin, ga.ctypes.data, gz.ctypes.data, ga.size, C.byref(scal))
    if rc != L.OK:
        raise L.PlaidGpuError(rc, "plaidgpu_combine_medians failed")
    return scal


def score_shard(ctx, comm, M: L.Matrix, rowmap: np.ndarray, opts: L.Opts, out_ptr: int, n_cols: int):
    """Run the begin / compute / finish protocol for this rank's column shard."""
    lib = ctx.lib
    local = L.Scalars()
    ctx.check(lib.plaidgpu_score_begin(ctx.h, C.byref(M), rowmap.ctypes.data, C.byref(opts), C.byref(local)))
    scal = combine_scalars(comm, local)
    ctx.check(lib.plaidgpu_score_compute(ctx.h, C.byref(scal), out_ptr))
    needs_norm = (opts.scorer in (L.SSGSEA, L.UCELL, L.AUCELL)) or (opts.scorer == L.PLAID and opts.normalize)
    if needs_norm:
        ma = np.empty(n_cols, dtype=np.float64)
        mz = np.empty(n_cols, dtype=np.float64)
        ctx.check(lib.plaidgpu_get_col_medians(ctx.h, ma.ctypes.data, mz.ctypes.data))
        combine_medians(lib, comm, opts.ignore_zero, scal, ma, mz)
    ctx.check(lib.plaidgpu_score_finish(ctx.h, C.byref(scal), out_ptr))
    return scal


def shard_columns(n_total: int, world: int, rank: int):
    """contiguous, equal-count column ranges (equal output bytes, the dominant cost)"""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def score_multi(ctxs, mats, rowmap: np.ndarray, opts_list, out_ptrs, n_cols):
    """The same protocol with several contexts driven by ONE process (one context per GPU, or several
    on one GPU): shards in column order.  This is how a single R process would use more than one GPU."""
    lib = ctxs[0].lib
    k = len(ctxs)
    locs = [L.Scalars() for _ in range(k)]
    for c, M, o, loc in zip(ctxs, mats, opts_list, locs):
        c.check(lib.plaidgpu_score_begin(c.h, C.byref(M), rowmap.ctypes.data, C.byref(o), C.byref(loc)))
    g = L.Scalars()
    C.memmove(C.byref(g), C.byref(locs[0]), C.sizeof(L.Scalars))
    g.x_min = min(l.x_min for l in locs)
    g.x_max = max(l.x_max for l in locs)
    g.rank_max = max(l.rank_max for l in locs)
    scal = []
    for c, outp in zip(ctxs, out_ptrs):
        s = L.Scalars()
        C.memmove(C.byref(s), C.byref(g), C.sizeof(L.Scalars))
        c.check(lib.plaidgpu_score_compute(c.h, C.byref(s), outp))
        scal.append(s)
    o0 = opts_list[0]
    if (o0.scorer in (L.SSGSEA, L.UCELL, L.AUCELL, L.GSVA)) or (o0.scorer == L.PLAID and o0.normalize):
        mas, mzs = [], []
        for c, n in zip(ctxs, n_cols):
            ma = np.empty(n)
            mz = np.empty(n)
            c.check(lib.plaidgpu_get_col_medians(c.h, ma.ctypes.data, mz.ctypes.data))
            mas.append(ma)
            mzs.append(mz)
        ga, gz = np.concatenate(mas), np.concatenate(mzs)
        smin = min(s.score_min for s in scal)
        for s in scal:
            rc = lib.plaidgpu_combine_medians(int(o0.ignore_zero), smin, ga.ctypes.data, gz.ctypes.data, ga.size, C.byref(s))
            if rc != L.OK:
                raise L.PlaidGpuError(rc, "plaidgpu_combine_medians failed")
    for c, s, outp in zip(ctxs, scal, out_ptrs):
        c.check(lib.plaidgpu_score_finish(c.h, C.byref(s), outp))
    return scal
