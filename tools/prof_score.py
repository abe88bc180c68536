"""Tiny driver for ncu: a few plaid() calls on a device-resident C4-shaped shard."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plaid_b200 as pb
from plaid_b200 import synth

P, S = 20000, 30000
N = int(os.environ.get("PROBE_CELLS", "4096"))
NORM = os.environ.get("PROBE_NORM", "0") == "1"
REPS = int(os.environ.get("PROBE_REPS", "3"))
Gp, Gi = synth.genesets_torch(P, S, seed=synth.SEED0 + 3, device="cuda")
G = sp.csc_matrix((np.ones(Gi.size), Gi, Gp), shape=(P, S))
p, i, x = synth.sparse_x_torch(P, N, seed=synth.SEED0 + 3, device="cuda")
names = synth.gene_names(P)
out = torch.empty(S * N, dtype=torch.float64, device="cuda")
ctx = pb.Context(0)
for _ in range(REPS):
    pb.plaid(pb.NamedMatrix(pb.DeviceCSC(p, i, x, (P, N)), names), pb.NamedMatrix(G, names), normalize=NORM, ctx=ctx, out=out)
print("score ms", ctx.kernel_ms(0), "colstats", ctx.kernel_ms(1), "fixup", ctx.kernel_ms(2))
