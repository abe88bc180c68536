"""The same call through every output route must give the same bits: device buffer, pinned host buffer (early shipping +
host-side fix-up of the early part), pageable host buffer (pinned ring + copy threads), column-chunked host path, and
n contexts from one process.  Random shapes with several column chunks / tail tiles."""
import os, sys
import numpy as np, scipy.sparse as sp, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plaid_b200 as pb
from plaid_b200 import synth

rng = np.random.default_rng(int(os.environ.get("SEED", "9")))
ctx = pb.Context(0)
bad = 0
for case in range(int(os.environ.get("CASES", "8"))):
    P = int(rng.choice([1500, 3000]))
    S = int(rng.choice([1100, 2300]))
    N = int(rng.choice([1057, 5000, 17000, 40000]))
    X = synth.sparse_x_numpy(P, N, seed=int(rng.integers(1 << 30)))
    G = synth.genesets_numpy(P, S, seed=int(rng.integers(1 << 30)), size_cap=(3, 400))
    names = synth.gene_names(P)
    Xn, Gn = pb.NamedMatrix(X, names), pb.NamedMatrix(G, names)
    name, fn = [("plaid", lambda **k: pb.plaid(Xn, Gn, **k)), ("ssgsea", lambda **k: pb.replaid_ssgsea(Xn, Gn, **k)),
                ("ucell", lambda **k: pb.replaid_ucell(Xn, Gn, rmax=500, **k)), ("plaid_raw", lambda **k: pb.plaid(Xn, Gn, normalize=False, **k))][int(rng.integers(4))]
    ref = fn(ctx=ctx).mat                                   # pageable numpy result
    dev = torch.empty(S * N, dtype=torch.float64, device="cuda")
    fn(ctx=ctx, out=dev)
    pin = torch.empty(S * N, dtype=torch.float64).pin_memory()
    fn(ctx=ctx, out=pin); fn(ctx=ctx, out=pin)              # second call ships early with measured rates
    os.environ["PLAIDGPU_MAX_OUT_BYTES"] = str(int(S * 8 * max(1056, N // 3)))
    chunked = fn(ctx=ctx).mat
    del os.environ["PLAIDGPU_MAX_OUT_BYTES"]
    r = ref.ravel(order="F")
    oks = [np.array_equal(dev.cpu().numpy(), r), np.array_equal(pin.numpy(), r), np.array_equal(chunked.ravel(order="F"), r)]
    print(f"case {case} {name:9s} P={P} S={S} N={N}: device {oks[0]} pinned {oks[1]} chunked {oks[2]}", flush=True)
    bad += 0 if all(oks) else 1
print("failures", bad)
sys.exit(0 if bad == 0 else 1)
