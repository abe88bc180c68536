"""pageable (malloc) host buffers on both sides, as an R caller has them: timeline with PLAIDGPU_TRACE=1 (development)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import plaid_b200 as pb
from plaid_b200 import _lib as L, sharded, synth
from plaid_b200.api import _opts
Nc = int(os.environ.get("CELLS", "125000"))
G, xp, xi, xx = bench.make_inputs("cuda:0", 0, Nc)
names = synth.gene_names(bench.P_GENES)
rowmap = pb.make_rowmap(names, names)
ctx = pb.Context(0); ctx.set_genesets(G)
S = bench.S_SETS
hp = xp.cpu().numpy().copy(); hi = xi.cpu().numpy().copy(); hx = xx.cpu().numpy().copy()
Mh = L.Matrix(); Mh.kind, Mh.location, Mh.P, Mh.N = L.CSC, L.HOST, bench.P_GENES, Nc
Mh.p, Mh.i, Mh.x = hp.ctypes.data, hi.ctypes.data, hx.ctypes.data
oh = _opts(ctx.lib, scorer=L.PLAID, stats_mean=1, normalize=1, out_location=L.HOST)
pout = np.empty(S * Nc, dtype=np.float64)
comm = sharded.LocalComm()
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sharded.score_shard(ctx, comm, Mh, rowmap, oh, pout.ctypes.data, Nc)
    print(f"call {it}: {1e3 * (time.perf_counter() - t0):.1f} ms", file=sys.stderr, flush=True)
