"""GPU probe: time the score / colstats / fixup / rank kernels on a C4-shaped shard (device-resident)."""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plaid_b200 as pb
from plaid_b200 import _lib as L, synth

P, S = 20000, 30000
N = int(os.environ.get("PROBE_CELLS", "16384"))
dev = "cuda"
t = time.time()
Gp, Gi = synth.genesets_torch(P, S, seed=synth.SEED0 + 3, device=dev)
G = sp.csc_matrix((np.ones(Gi.size), Gi, Gp), shape=(P, S))
print(f"G nnz {G.nnz} gen {time.time()-t:.1f}s", flush=True)
t = time.time()
p, i, x = synth.sparse_x_torch(P, N, seed=synth.SEED0 + 3, device=dev)
torch.cuda.synchronize()
print(f"X nnz {x.numel()} ({x.numel()/N:.0f}/cell) gen {time.time()-t:.1f}s", flush=True)
deg = torch.tensor(np.asarray(G.sum(1)).ravel(), device=dev)
adds = float(deg[i.long()].sum().item())
print(f"adds {adds:.3e} = {adds/N:.0f}/cell = {adds/N/S:.2f}/output")
names = synth.gene_names(P)
Xd = pb.NamedMatrix(pb.DeviceCSC(p, i, x, (P, N)), names)
Gn = pb.NamedMatrix(G, names)
out = torch.empty(S * N, dtype=torch.float64, device=dev)
for warps in os.environ.get("PROBE_WARPS", "8,4").split(","):
    os.environ["PLAIDGPU_WARPS"] = warps
    ctx = pb.Context(0)
    for norm in (False, True):
        for rep in range(3):
            t = time.time()
            pb.plaid(Xd, Gn, normalize=norm, ctx=ctx, out=out)
            dt = time.time() - t
        ms = [ctx.kernel_ms(k) for k in range(4)]
        info = ctx.plan_info()
        byts = x.numel() * 12 + S * N * 8
        print(f"warps {warps} norm {norm}: wall {dt*1e3:.1f} ms | score {ms[0]:.2f} ms colstats {ms[1]:.2f} fixup {ms[2]:.2f} | "
              f"{S*N/ (ms[0]*1e-3):.3e} cells*sets/s (score only) | {byts/ms[0]/1e6:.0f} GB/s | adds/cycle/SM@1.9GHz "
              f"{adds/(ms[0]*1e-3)/148/1.9e9:.2f} | {info}", flush=True)
    ctx.close()
# rank kernel
ctx = pb.Context(0)
for sc, name in ((pb.replaid_ssgsea, "ssgsea"), (pb.replaid_ucell, "ucell"), (pb.replaid_sing, "sing")):
    for rep in range(2):
        sc(Xd, Gn, ctx=ctx, out=out)
    ms = [ctx.kernel_ms(k) for k in range(4)]
    print(f"{name}: rank {ms[3]:.2f} ms score {ms[0]:.2f} colstats {ms[1]:.2f} fixup {ms[2]:.2f}", flush=True)
