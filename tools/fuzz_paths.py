"""Randomised consistency sweep: the default fixed-point path (tensor-core block + integer tail) against the all-fp64
path of the same library on random shapes / densities / signs / scorers (both go through the C ABI; the fp64 path is
the one the parity tests pin to the oracle at 1e-11).  Prints the worst relative deviation; exits non-zero above 2e-8."""
import os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plaid_b200 as pb
from plaid_b200 import api, synth

rng = np.random.default_rng(int(os.environ.get("SEED", "7")))
ctx = pb.Context(0)
worst = 0.0
ncase = int(os.environ.get("CASES", "40"))
for case in range(ncase):
    P = int(rng.integers(300, 6000))
    S = int(rng.integers(1030, 7000))
    N = int(rng.choice([1, 2, 47, 48, 49, 200, 1055, 1056, 1057, 2500]))
    dens = float(rng.choice([0.01, 0.05, 0.2, 0.5]))
    X = sp.random(P, N, density=dens, format="csc", random_state=int(rng.integers(1 << 30)), data_rvs=lambda n: rng.lognormal(0.5, 0.8, n))
    mode = int(rng.integers(0, 4))
    if mode == 1:
        X.data *= rng.choice([-1.0, 1.0], size=X.data.size)        # signed values
    elif mode == 2:
        X.data = np.round(X.data, 1)                                 # heavy ties
    elif mode == 3:
        X.data *= 10.0 ** rng.integers(-6, 7, size=X.data.size)     # wide dynamic range inside a column
    G = synth.genesets_numpy(P, S, seed=int(rng.integers(1 << 30)), size_cap=(3, max(10, P // 4)))
    names = synth.gene_names(P)
    Xn, Gn = pb.NamedMatrix(X, names), pb.NamedMatrix(G, names)
    fns = [("plaid", lambda: pb.plaid(Xn, Gn, ctx=ctx)), ("plaid_sum_raw", lambda: pb.plaid(Xn, Gn, stats="sum", normalize=False, ctx=ctx)),
           ("ssgsea", lambda: pb.replaid_ssgsea(Xn, Gn, ctx=ctx)), ("ucell", lambda: pb.replaid_ucell(Xn, Gn, rmax=min(1500, P // 2), ctx=ctx)),
           ("sing", lambda: pb.replaid_sing(Xn, Gn, ctx=ctx)), ("scse", lambda: pb.replaid_scse(Xn, Gn, removeLog2=False, ctx=ctx))]
    name, fn = fns[int(rng.integers(len(fns)))]
    api.EXACT_FP64 = False
    a = fn().mat
    api.EXACT_FP64 = True
    try:
        b = fn().mat
    finally:
        api.EXACT_FP64 = False
    scale = float(np.max(np.abs(b))) or 1.0
    err = float(np.max(np.abs(a - b))) / scale
    nanmis = int(np.sum(np.isnan(a) != np.isnan(b)))
    info = ctx.plan_info()
    print(f"case {case:3d} {name:13s} P={P:5d} S={S:5d} N={N:5d} dens={dens:.2f} mode={mode} tc_rows={info.get('tc_rows')} tail_rows={info.get('tail_rows')} "
          f"err={err:.2e} nan_mismatch={nanmis}", flush=True)
    worst = max(worst, err if nanmis == 0 else 1.0)
print("worst", worst)
sys.exit(0 if worst < 2e-8 else 1)
