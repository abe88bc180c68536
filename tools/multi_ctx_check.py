"""one process, one context per visible GPU: column shards via sharded.score_multi must equal the single-GPU result"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plaid_b200 as pb
from plaid_b200 import _lib as L, sharded, synth
from plaid_b200.api import _matrix_struct, _opts
ng = torch.cuda.device_count()
P, N, S = 3000, 257, 5000
X = synth.sparse_x_numpy(P, N, seed=5)
G = synth.genesets_numpy(P, S, seed=6, size_cap=(5, 300))
names = synth.gene_names(P)
rowmap = pb.make_rowmap(names, names)
ctxs = [pb.Context(d) for d in range(ng)]
for c in ctxs:
    c.set_genesets(G)
keep = []
for scorer, kw in [(L.PLAID, dict(normalize=1)), (L.UCELL, dict(rmax=300.0))]:
    whole = np.empty((S, N), order="F")
    o = _opts(ctxs[0].lib, scorer=scorer, out_location=L.HOST, **kw)
    ctxs[0].check(ctxs[0].lib.plaidgpu_score(ctxs[0].h, _matrix_struct(X, keep), rowmap.ctypes.data, o, whole.ctypes.data))
    spans = [sharded.shard_columns(N, ng, r) for r in range(ng)]
    outs = [np.empty((S, hi - lo), order="F") for lo, hi in spans]
    mats = [_matrix_struct(X[:, lo:hi], keep) for lo, hi in spans]
    opts = [_opts(ctxs[0].lib, scorer=scorer, out_location=L.HOST, **kw) for _ in spans]
    sharded.score_multi(ctxs, mats, rowmap, opts, [a.ctypes.data for a in outs], [hi - lo for lo, hi in spans])
    print("scorer", scorer, "gpus", ng, "bit-identical:", np.array_equal(np.concatenate(outs, axis=1), whole))
