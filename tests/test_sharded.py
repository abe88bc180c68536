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


# ---- replaid.gsva on column shards: the column -> row exchange ---------------
def _ecdf_rows(block):
    """ecdf(x)(x) of every row (R/plaid.R:346), numpy"""
    return np.stack([(r[None, :] <= r[:, None]).sum(axis=1) / r.size for r in block])


def _exchange_worker(rank, world, port, P, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    comm = sharded.TorchComm(device="cpu")
    X = np.round(np.random.default_rng(7).normal(size=(P, n_total)), 1)  # rounding: ties across samples
    lo, hi = sharded.shard_columns(n_total, world, rank)
    block, counts = sharded.exchange_to_rows(comm, X[:, lo:hi])
    g0, g1 = sharded.shard_columns(P, world, rank)
    ok_rows = np.array_equal(block, X[g0:g1, :]) and block.flags.c_contiguous
    back = sharded.exchange_to_columns(comm, _ecdf_rows(block), counts)
    ok_back = np.array_equal(back, _ecdf_rows(X)[:, lo:hi]) and back.flags.f_contiguous
    sums = comm.allreduce_sum_vec(X[:, lo:hi].sum(axis=1))
    ref = X[:, :sharded.shard_columns(n_total, world, 0)[1]].sum(axis=1)
    for r in range(1, world):
        a, b = sharded.shard_columns(n_total, world, r)
        ref = ref + X[:, a:b].sum(axis=1)
    q.put((rank, ok_rows, ok_back, counts, bool(np.array_equal(sums, ref))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_row_exchange_round_trip():
    P, n_total, world = 37, 23, 2  # both axes ragged
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_exchange_worker, args=(r, world, port, P, n_total, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert r[1] and r[2] and r[4]
        assert r[3] == [12, 11]


def test_thread_comm_collectives():
    import threading
    world = 3
    comms = sharded.ThreadComm.group(world)
    X = np.random.default_rng(3).normal(size=(10, 8))
    out = [None] * world

    def run(r):
        c = comms[r]
        lo, hi = sharded.shard_columns(8, world, r)
        block, counts = sharded.exchange_to_rows(c, X[:, lo:hi])
        back = sharded.exchange_to_columns(c, block * 2.0, counts)
        out[r] = (c.allreduce_min(float(r)), c.allreduce_max(float(r)), c.allgather_vec(np.array([float(r)])),
                  c.allreduce_sum_vec(np.array([1.0, r])), block, back, lo, hi)

    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=60)
    for r, o in enumerate(out):
        g0, g1 = sharded.shard_columns(10, world, r)
        assert o[0] == 0.0 and o[1] == 2.0 and list(o[2]) == [0.0, 1.0, 2.0] and list(o[3]) == [3.0, 3.0]
        assert np.array_equal(o[4], X[g0:g1]) and np.array_equal(o[5], 2.0 * X[:, o[6]:o[7]])
