"""Device-resident timings of the other BASELINE.json configs (C2 dense bulk, C3 rank scorers), for DESIGN.md.
Not the bench contract (bench.py is); prints one line per config."""
import os, sys, time
import numpy as np, scipy.sparse as sp, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plaid_b200 as pb
from plaid_b200 import synth

dev = "cuda"
P, S = 20000, 30000
Gp, Gi = synth.genesets_torch(P, S, seed=synth.SEED0 + 3, device=dev)
G = sp.csc_matrix((np.ones(Gi.size), Gi, Gp), shape=(P, S))
names = synth.gene_names(P)
Gn = pb.NamedMatrix(G, names)
ctx = pb.Context(0)

def timeit(f, reps=3):
    f(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps

# C2: dense bulk 20k x 1000, 30k sets
N2 = 1000
gen = torch.Generator(device=dev); gen.manual_seed(synth.SEED0 + 1)
mu = torch.rand(P, generator=gen, device=dev, dtype=torch.float64) * 12 + 2
sg = torch.rand(P, generator=gen, device=dev, dtype=torch.float64) * 1.2 + 0.3
Xd = (mu[None, :] + sg[None, :] * torch.randn((N2, P), generator=gen, device=dev, dtype=torch.float64)).contiguous()  # (N,P) contiguous = P x N col-major
out2 = torch.empty(S * N2, dtype=torch.float64, device=dev)
X2 = pb.NamedMatrix(pb.DeviceDense(Xd, (P, N2)), names)
for norm in (False, True):
    dt = timeit(lambda: pb.plaid(X2, Gn, normalize=norm, ctx=ctx, out=out2))
    print(f"C2 dense plaid(normalize={norm}) 20000x{N2} x {S} sets: {dt*1e3:.2f} ms wall, score kernels {ctx.kernel_ms(0):.2f} ms, "
          f"{S*N2/dt:.3e} cells*sets/s, alg bytes {(P*N2*8+G.nnz*4+S*N2*8)/1e6:.0f} MB -> {(P*N2*8+G.nnz*4+S*N2*8)/ctx.kernel_ms(0)/1e6:.0f} GB/s", flush=True)
del Xd, out2
# C3: rank scorers on 20k x 100k sparse (subsample 32768 cells to keep the probe short)
N3 = int(os.environ.get("C3_CELLS", "32768"))
p, i, x = synth.sparse_x_torch(P, N3, seed=synth.SEED0 + 2, device=dev)
X3 = pb.NamedMatrix(pb.DeviceCSC(p, i, x, (P, N3)), names)
out3 = torch.empty(S * N3, dtype=torch.float64, device=dev)
for name, f in (("ssgsea(alpha=0)", lambda: pb.replaid_ssgsea(X3, Gn, ctx=ctx, out=out3)), ("sing", lambda: pb.replaid_sing(X3, Gn, ctx=ctx, out=out3)),
                ("ucell", lambda: pb.replaid_ucell(X3, Gn, ctx=ctx, out=out3)), ("aucell", lambda: pb.replaid_aucell(X3, Gn, ctx=ctx, out=out3)),
                ("scse", lambda: pb.replaid_scse(X3, Gn, ctx=ctx, out=out3))):
    dt = timeit(f, reps=2)
    ms = [ctx.kernel_ms(k) for k in range(4)]
    print(f"C3 {name} 20000x{N3} sparse x {S} sets: {dt*1e3:.1f} ms wall | rank {ms[3]:.2f} score {ms[0]:.2f} colstats {ms[1]:.2f} fixup {ms[2]:.2f} ms | "
          f"{S*N3/dt:.3e} cells*sets/s | rank-only {x.numel()*16/ms[3]/1e6:.0f} GB/s (nnz*16 B)", flush=True)
# rank only: colranks kernel time
