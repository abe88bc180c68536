"""stress of the gsva shard test (development): repeats the body of test_gsva_on_column_shards and reports which step goes wrong"""
import sys, threading, numpy as np
sys.path.insert(0,'/root/repo')
import plaid_b200 as pb
from plaid_b200 import synth, api, _lib as L, sharded
from plaid_b200.api import _opts
from oracle import plaid_oracle as O
P, N, S = 900, 53, 700
Draw = synth.dense_x_numpy(P, N, seed=61)
Dtie = np.round(Draw, 1)
G = synth.genesets_numpy(P, S, seed=62, size_cap=(5, 150))
names = synth.gene_names(P)
rowmap = pb.make_rowmap(names, names)
Go = O.Named(G, names)
want = {}
for rowtf, tau in (("ecdf", 0.0), ("z", 0.0), ("ecdf", 0.5)):
    D = Dtie if rowtf == "ecdf" else Draw
    want[(rowtf,tau)] = O.replaid_gsva(O.Named(D, names), Go, tau=tau, rowtf=rowtf).mat
def report(tag, got, ref):
    err = np.abs(got - ref)
    if err.max() > 1e-6:
        bad = np.argwhere(err > 1e-6)
        print("FAIL %s maxerr=%g nbad=%d cols=%s rows[%d..%d]" % (tag, err.max(), len(bad), sorted(set(bad[:,1]))[:12], bad[:,0].min(), bad[:,0].max()), flush=True)
for mode in (False, True):
    api.EXACT_FP64 = mode
    for it in range(25):
        ctxs = [pb.Context(0) for _ in range(3)]
        for c in ctxs: c.set_genesets(G)
        for rowtf, tau in (("ecdf", 0.0), ("z", 0.0), ("ecdf", 0.5)):
            D = Dtie if rowtf == "ecdf" else Draw
            whole = pb.replaid_gsva(pb.NamedMatrix(D, names), pb.NamedMatrix(G, names), tau=tau, rowtf=rowtf, ctx=ctxs[0]).mat
            report(f"whole exact={mode} it={it} {rowtf} {tau}", whole, want[(rowtf,tau)])
            for world in (2, 3):
                comms = sharded.ThreadComm.group(world)
                spans = [sharded.shard_columns(N, world, r) for r in range(world)]
                outs = [np.empty((S, hi - lo), order="F") for lo, hi in spans]
                def run(r):
                    lo, hi = spans[r]
                    o = _opts(ctxs[r].lib, scorer=L.GSVA, out_location=L.HOST, tau=tau)
                    sharded.gsva_shard(ctxs[r], comms[r], D[:, lo:hi], rowmap, o, outs[r].ctypes.data, rowtf=rowtf)
                ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
                for t in ts: t.start()
                for t in ts: t.join(timeout=120)
                got = np.concatenate(outs, axis=1)
                report(f"shards exact={mode} it={it} {rowtf} {tau} world={world}", got, want[(rowtf,tau)])
        for c in ctxs: c.close()
print("done")
