"""per-step wall time vs kernel time of the bench step (device-resident), to find host-side overheads"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import plaid_b200 as pb
from plaid_b200 import _lib as L, sharded, synth
from plaid_b200.api import _matrix_struct, _opts
Nc = int(os.environ.get("CELLS", "125000"))
G, xp, xi, xx = bench.make_inputs("cuda:0", 0, Nc)
names = synth.gene_names(bench.P_GENES)
rowmap = pb.make_rowmap(names, names)
ctx = pb.Context(0); ctx.set_genesets(G)
out = torch.empty(bench.S_SETS * Nc, dtype=torch.float64, device="cuda:0")
keep = []
M = _matrix_struct(pb.DeviceCSC(xp, xi, xx, (bench.P_GENES, Nc)), keep)
opts = _opts(ctx.lib, scorer=L.PLAID, stats_mean=1, normalize=1, out_location=L.DEVICE)
comm = sharded.LocalComm()
import ctypes as C
for i in range(14):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    loc = L.Scalars()
    ctx.check(ctx.lib.plaidgpu_score_begin(ctx.h, C.byref(M), rowmap.ctypes.data, C.byref(opts), C.byref(loc)))
    t1 = time.perf_counter()
    ctx.check(ctx.lib.plaidgpu_score_compute(ctx.h, C.byref(loc), out.data_ptr()))
    t2 = time.perf_counter()
    sharded.combine_medians_of(ctx, comm, -1, loc, Nc)
    t3 = time.perf_counter()
    ctx.check(ctx.lib.plaidgpu_score_finish(ctx.h, C.byref(loc), out.data_ptr()))
    torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f"step {i}: total {1e3*(t4-t0):.1f} ms | begin {1e3*(t1-t0):.1f} compute {1e3*(t2-t1):.1f} medians {1e3*(t3-t2):.1f} finish {1e3*(t4-t3):.1f} | kernels {[round(ctx.kernel_ms(k),1) for k in range(3)]}", flush=True)
