"""World-size-2 gloo tests (CPU) of the host-side sharded protocol: column partition, scalar
all-reduces and the column-ordered median combine must give bit-identical global scalars for any
number of shards.  The kernels themselves are covered by the -m gpu tests."""
import ctypes as C
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from plaid_b200 import _lib as L, sharded


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = sharded.TorchComm(device="cpu")
    rng = np.random.default_rng(123)  # same stream on every rank = the "global" data
    med_all = rng.normal(size=n_total)
    med_nz = rng.normal(size=n_total)
    colmin = np.abs(rng.normal(size=n_total))
    colmin[n_total // 3] = 0.0  # the global min == 0 lives in exactly one shard
    lo, hi = sharded.shard_columns(n_total, world, rank)
    local = L.Scalars()
    local.x_min, local.x_max, local.rank_max = float(lo), float(hi), 100.0 + rank
    local.score_min = float(colmin[lo:hi].min()) if hi > lo else float("inf")
    g = sharded.combine_scalars(comm, local)
    g.score_min = local.score_min
    sharded.combine_medians(L.load(), comm, -1, g, med_all[lo:hi].copy(), med_nz[lo:hi].copy())
    q.put((rank, lo, hi, g.x_min, g.x_max, g.rank_max, g.score_min, g.ignore_zero, g.med_mean))
    dist.barrier()
    dist.destroy_process_group()


def _single(n_total):
    rng = np.random.default_rng(123)
    med_all = rng.normal(size=n_total)
    med_nz = rng.normal(size=n_total)
    s = L.Scalars()
    L.load().plaidgpu_combine_medians(-1, 0.0, med_all.ctypes.data, med_nz.ctypes.data, n_total, C.byref(s))
    return s.ignore_zero, s.med_mean


def test_two_rank_protocol_matches_single_shard_bit_for_bit():
    n_total, world = 1001, 2  # odd: ragged shards
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1:3] for r in res] == [(0, 501), (501, 1001)]
    iz, mean = _single(n_total)
    for r in res:
        assert r[3] == 0.0 and r[4] == 1001.0 and r[5] == 101.0  # min / max / max over shards
        assert r[6] == 0.0 and r[7] == iz == 1
        assert r[8] == mean  # bit-identical mean(medx): combined in global column order


def test_shard_columns_cover_everything():
    for n, w in [(10, 3), (7, 8), (125000 * 8, 8), (0, 2)]:
        spans = [sharded.shard_columns(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
