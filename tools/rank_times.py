"""k_rank timings on the three column shapes of the path: tied single-cell columns (counting
path), all-distinct sparse columns (shared-memory sort path), dense bulk columns of 20,000 distinct values.
Prints rank-kernel ms (ctx.kernel_ms(3)) and the rate over the 16 B per entry the kernel reads and writes."""
import os, sys
import numpy as np, scipy.sparse as sp, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plaid_b200 as pb
from plaid_b200 import synth

dev = "cuda"
P, S = 20000, 2000
Gp, Gi = synth.genesets_torch(P, S, seed=synth.SEED0 + 3, device=dev)
G = sp.csc_matrix((np.ones(Gi.size), Gi, Gp), shape=(P, S))
names = synth.gene_names(P)
Gn = pb.NamedMatrix(G, names)
ctx = pb.Context(0)
N = int(os.environ.get("CELLS", "32768"))
p, i, x = synth.sparse_x_torch(P, N, seed=synth.SEED0 + 2, device=dev)
out = torch.empty(S * N, dtype=torch.float64, device=dev)
gen = torch.Generator(device=dev); gen.manual_seed(7)
cases = [("sparse, tied values (single-cell counts)", pb.DeviceCSC(p, i, x, (P, N)), x.numel()),
         ("sparse, all-distinct values", pb.DeviceCSC(p, i, x + torch.rand(x.numel(), generator=gen, device=dev, dtype=torch.float64), (P, N)), x.numel())]
Nd = 1000
Xd = torch.randn((Nd, P), generator=gen, device=dev, dtype=torch.float64).contiguous()
cases.append(("dense bulk 20000 x 1000, all distinct", pb.DeviceDense(Xd, (P, Nd)), P * Nd))
sel = os.environ.get("CASE")
for ci, (name, X, nent) in enumerate(cases):
    if sel is not None and int(sel) != ci:
        continue
    Xn = pb.NamedMatrix(X, names)
    for _ in range(3):
        pb.replaid_sing(Xn, Gn, ctx=ctx, out=out)
    torch.cuda.synchronize()
    ms = ctx.kernel_ms(3)
    print(f"{name}: rank {ms:.3f} ms for {nent} entries = {nent * 16 / ms / 1e6:.0f} GB/s over 16 B/entry ({ms * 1e6 / nent:.2f} ns/entry)", flush=True)
