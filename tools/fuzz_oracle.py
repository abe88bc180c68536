"""Randomised parity sweep against the CPU oracle on small random shapes: every scorer, sparse and dense input, signed
values, ties, NaN-free; ranks bit-exact, scores within 2e-8 of the oracle (fixed-point path) — shapes chosen around the
switch points of the library (S = 1024 for the tensor-core path, 48-cell tiles, 4096-row columns for the sampled median)."""
import os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plaid_b200 as pb
from oracle import plaid_oracle as O
from plaid_b200 import synth

rng = np.random.default_rng(int(os.environ.get("SEED", "3")))
ctx = pb.Context(0)
worst = 0.0
bad = 0
for case in range(int(os.environ.get("CASES", "60"))):
    P = int(rng.choice([40, 300, 1500, 4200]))
    S = int(rng.choice([1, 7, 900, 1023, 1024, 1025, 4500]))
    N = int(rng.choice([1, 3, 47, 48, 49, 130]))
    dens = float(rng.choice([0.02, 0.2, 0.6]))
    X = sp.random(P, N, density=dens, format="csc", random_state=int(rng.integers(1 << 30)), data_rvs=lambda n: np.round(rng.lognormal(0.5, 0.8, n), int(rng.integers(1, 6))))
    signed = bool(rng.integers(2))
    if signed:
        X.data *= rng.choice([-1.0, 1.0], size=X.data.size)
    dense = bool(rng.integers(3) == 0)
    G = synth.genesets_numpy(P, S, seed=int(rng.integers(1 << 30)), size_cap=(1, max(2, P // 3)))
    names = synth.gene_names(P)
    Xi = X.toarray() if dense else X
    Xg, Gg, Xo, Go = pb.NamedMatrix(Xi, names), pb.NamedMatrix(G, names), O.Named(Xi, names), O.Named(G, names)
    cases = [("plaid", lambda: pb.plaid(Xg, Gg, ctx=ctx).mat, lambda: O.plaid(Xo, Go).mat),
             ("plaid_sum", lambda: pb.plaid(Xg, Gg, stats="sum", normalize=False, ctx=ctx).mat, lambda: O.plaid(Xo, Go, stats="sum", normalize=False).mat),
             ("sing", lambda: pb.replaid_sing(Xg, Gg, ctx=ctx).mat, lambda: O.replaid_sing(Xo, Go).mat),
             ("ssgsea", lambda: pb.replaid_ssgsea(Xg, Gg, alpha=0.25 * int(rng.integers(2)), ctx=ctx).mat, None),
             ("ucell", lambda: pb.replaid_ucell(Xg, Gg, rmax=max(5, P // 3), ctx=ctx).mat, lambda: O.replaid_ucell(Xo, Go, rmax=max(5, P // 3)).mat),
             ("aucell", lambda: pb.replaid_aucell(Xg, Gg, ctx=ctx).mat, lambda: O.replaid_aucell(Xo, Go).mat),
             ("scse", lambda: pb.replaid_scse(Xg, Gg, ctx=ctx).mat, lambda: O.replaid_scse(Xo, Go).mat),
             ("gsva_z", lambda: pb.replaid_gsva(Xg, Gg, ctx=ctx).mat, lambda: O.replaid_gsva(Xo, Go).mat),
             ("colranks", lambda: pb.colranks(Xi, signed=signed, ctx=ctx), lambda: O.colranks(Xi, signed=signed)),
             ("normalize_medians", None, None)]
    name, gf, of = cases[int(rng.integers(len(cases)))]
    try:
        if name == "ssgsea":
            al = 0.25 * int(rng.integers(2))
            got, want = pb.replaid_ssgsea(Xg, Gg, alpha=al, ctx=ctx).mat, O.replaid_ssgsea(Xo, Go, alpha=al).mat
        elif name == "normalize_medians":
            M = np.round(rng.normal(size=(int(rng.choice([5, 500, 4096, 9000])), N)), 2)
            M[rng.random(M.shape) < 0.3] = 0.0
            iz = [None, False, True][int(rng.integers(3))]
            got, want = pb.normalize_medians(M, ignore_zero=iz, ctx=ctx), O.normalize_medians(M, ignore_zero=iz)
        elif name == "gsva_z" and N < 3:
            continue
        else:
            got, want = gf(), of()
    except Exception as ex:
        print(f"case {case:3d} {name:17s} P={P} S={S} N={N} dens={dens} dense={dense} signed={signed}: EXCEPTION {type(ex).__name__}: {str(ex)[:120]}", flush=True)
        bad += 1
        continue
    if got is None or want is None:
        ok = got is None and want is None
        err = 0.0 if ok else 1.0
    else:
        got, want = np.asarray(got.todense() if sp.issparse(got) else got), np.asarray(want.todense() if sp.issparse(want) else want)
        nanm = int(np.sum(np.isnan(got) != np.isnan(want)))
        scale = float(np.nanmax(np.abs(want))) if np.isfinite(np.nanmax(np.abs(want), initial=0.0)) and np.nanmax(np.abs(want), initial=0.0) > 0 else 1.0
        err = float(np.nanmax(np.abs(got - want), initial=0.0)) / scale if nanm == 0 else 1.0
        if name == "colranks":
            err = 0.0 if (nanm == 0 and np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])) else 1.0
    flag = "" if err < 2e-8 else "   <-- FAIL"
    if flag:
        bad += 1
    print(f"case {case:3d} {name:17s} P={P:5d} S={S:5d} N={N:4d} dens={dens:.2f} dense={int(dense)} signed={int(signed)} err={err:.2e}{flag}", flush=True)
    worst = max(worst, err)
print("worst", worst, "failures", bad)
sys.exit(0 if bad == 0 else 1)
