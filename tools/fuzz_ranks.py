"""Randomised bit-exactness sweep of k_rank against the oracle: short and long columns (counting path, bucket path with
256- and 1,024-thread CTAs, sorting network for first / last / dense), ties of every density, signed ranks, NaN, stored
zeros, dense semantics with implicit zeros."""
import os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plaid_b200 as pb
from oracle import plaid_oracle as O

rng = np.random.default_rng(int(os.environ.get("SEED", "5")))
ctx = pb.Context(0)
bad = 0


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


for case in range(int(os.environ.get("CASES", "40"))):
    P = int(rng.choice([60, 255, 256, 257, 700, 4000, 8191, 8192, 8193, 20000, 30000]))
    N = int(rng.choice([1, 2, 5, 9]))
    nd = int(rng.choice([1, 3, 40, 447, 448, 449, 2000, 10 ** 9]))   # distinct values per column (10^9: all distinct)
    dens = float(rng.choice([0.05, 0.3, 1.0]))
    vals = rng.normal(size=(P, N))
    if nd < 10 ** 9:
        pool = rng.normal(size=nd)
        vals = pool[rng.integers(0, nd, size=(P, N))]
    mask = rng.random((P, N)) < dens
    D = np.where(mask, vals, 0.0)
    if rng.integers(2):
        D[rng.random((P, N)) < 0.01] = np.nan
    if rng.integers(2):
        D[rng.random((P, N)) < 0.02] = -0.0
    X = sp.csc_matrix(D)            # NaN stays stored; explicit zeros dropped by scipy
    if rng.integers(2) and X.nnz:
        X.data[rng.integers(0, X.nnz, size=max(1, X.nnz // 50))] = 0.0   # stored zeros
    ties = str(rng.choice(["average", "min", "max", "first", "last", "dense"]))
    signed = bool(rng.integers(2))
    kind = str(rng.choice(["dense", "sparse_stored", "sparse_full"]))
    try:
        if kind == "dense":
            ok = same(pb.colranks(D, signed=signed, ties_method=ties, ctx=ctx), O.colranks(D, signed=signed, ties_method=ties))
        elif kind == "sparse_stored":
            if ties == "dense":
                ties = "last"
            ok = same(pb.sparse_colranks(X, signed=signed, ties_method=ties, ctx=ctx).data, O.sparse_colranks(X, signed=signed, ties_method=ties).data)
        else:
            if ties in ("first", "last", "dense"):
                ties = "average"
            ok = same(pb.colranks(X, signed=signed, ties_method=ties, ctx=ctx), O.colranks(X, signed=signed, ties_method=ties))
    except Exception as ex:
        ok = False
        print("EXCEPTION", type(ex).__name__, str(ex)[:160])
    print(f"case {case:3d} {kind:13s} P={P:5d} N={N} distinct={nd:10d} dens={dens:.2f} ties={ties:7s} signed={int(signed)} {'ok' if ok else 'FAIL'}", flush=True)
    bad += 0 if ok else 1
print("failures", bad)
sys.exit(0 if bad == 0 else 1)
