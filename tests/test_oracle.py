"""Oracle self-consistency (CPU): every oracle function against a second, independent
implementation, the reference vignette's known answers, and hand-computed miniature cases.
The reference's own tests pin nothing on this path (tests/testthat/test-plaid.R:1-3); the p-values its
vignette prints pin plaid() + normalize_medians (tests/test_reference_known_answers.py), and this file is
what guards the rest of the oracle ("unpinned", oracle/__init__.py)."""
import math

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.stats import rankdata

from oracle import plaid_oracle as O
from plaid_b200 import synth

from conftest import rel_err


def test_rank_vector_miniature():
    # SURVEY.md §8c hand-computed case
    x = np.array([1.1, 0, 1.1, 2.5, 0, -0.3, 2.5, 2.5, 0])
    assert np.array_equal(O.rank_vector(x, "average"), [5.5, 3, 5.5, 8, 3, 1, 8, 8, 3])
    assert np.array_equal(O.rank_vector(x, "min"), [5, 2, 5, 7, 2, 1, 7, 7, 2])
    assert np.array_equal(O.rank_vector(x, "max"), [6, 4, 6, 9, 4, 1, 9, 9, 4])


@pytest.mark.parametrize("ties", ["average", "min", "max"])
def test_rank_vector_vs_scipy(ties):
    rng = np.random.default_rng(1)
    for n in [1, 2, 7, 100, 1000]:
        v = rng.integers(-5, 6, size=n).astype(float) * 0.25  # heavy ties, negatives, zeros
        v[rng.random(n) < 0.1] = -0.0
        assert np.array_equal(O.rank_vector(v, ties), rankdata(v, method=ties))


def test_rank_vector_order_of_appearance_methods():
    """ties.method first / last (base::rank) and dense (matrixStats::colRanks), R's documented example first:
    rank(c(3, 1, 4, 1, 5, 9, 2, 6, 5, 3, 5)) from ?rank"""
    x = np.array([3, 1, 4, 1, 5, 9, 2, 6, 5, 3, 5], dtype=float)
    assert np.array_equal(O.rank_vector(x, "first"), [4, 1, 6, 2, 7, 11, 3, 10, 8, 5, 9])   # ?rank: rank(x2, ties = "first")
    assert np.array_equal(O.rank_vector(x, "last"), [5, 2, 6, 1, 9, 11, 3, 10, 8, 4, 7])
    assert np.array_equal(O.rank_vector(x, "dense"), [3, 1, 4, 1, 5, 7, 2, 6, 5, 3, 5])
    rng = np.random.default_rng(2)
    for n in [1, 2, 9, 500]:
        v = rng.integers(-4, 5, size=n).astype(float) * 0.5
        v[rng.random(n) < 0.1] = -0.0
        assert np.array_equal(O.rank_vector(v, "first"), rankdata(v, method="ordinal"))
        assert np.array_equal(O.rank_vector(v, "dense"), rankdata(v, method="dense"))
        # last = first of the reversed vector, reversed back
        assert np.array_equal(O.rank_vector(v, "last"), rankdata(v[::-1], method="ordinal")[::-1])


def test_rank_vector_nan_kept():
    v = np.array([3.0, np.nan, 1.0, 3.0])
    r = O.rank_vector(v, "average")
    assert np.isnan(r[1]) and np.array_equal(r[[0, 2, 3]], [2.5, 1.0, 2.5])


def test_sparse_colranks_ranks_stored_entries_only():
    # explicit stored zero is ranked; implicit zeros are not (R/plaid.R:631-650)
    X = sp.csc_matrix((np.array([2.0, 0.0, -1.0, 5.0]), np.array([0, 2, 3, 1]), np.array([0, 3, 4])), shape=(4, 2))
    r = O.sparse_colranks(X)
    assert np.array_equal(r.data, [3, 2, 1, 1])
    rs = O.sparse_colranks(X, signed=True)
    assert np.array_equal(rs.data, [3, 0, -2, 1])  # rank(|x|) = 3,1,2 ; sign 1,0,-1


def test_colranks_zero_group_identity():
    """Dense-semantics ranks == ranks of stored entries shifted by the implicit-zero group."""
    X = synth.sparse_x_numpy(300, 20, seed=5, density=0.2)
    X.data[::7] *= -1.0
    X.data[::11] = 0.0  # explicit stored zeros join the zero group
    for ties in ["average", "min", "max"]:
        full = O.colranks(X, ties_method=ties)
        for j in range(X.shape[1]):
            col = X[:, j].toarray().ravel()
            assert np.array_equal(full[:, j], rankdata(col, method=ties))


def test_col_medians_and_normalize_medians_dual():
    rng = np.random.default_rng(3)
    x = rng.normal(size=(41, 9))
    x[rng.random(x.shape) < 0.3] = 0.0
    x[5, 2] = np.nan
    med = O.col_medians_narm(x)
    assert np.allclose(med, np.nanmedian(x, axis=0), rtol=0, atol=0)
    got = O.normalize_medians(x)  # min == 0 is false here (negatives) -> zeros kept
    assert np.nanmin(x) < 0
    want = x - np.nanmedian(x, axis=0)[None, :] + np.mean(np.nanmedian(x, axis=0))
    assert rel_err(got, want) < 1e-14
    y = np.abs(x)
    y[np.isnan(y)] = 1.0
    got = O.normalize_medians(y)  # min == 0 -> zeros ignored
    z = y.copy()
    z[z == 0] = np.nan
    m = np.nanmedian(z, axis=0)
    want = y - m[None, :] + m.mean()
    assert rel_err(got, want) < 1e-14
    # all-zero column: median NA -> 0 (R/plaid.R:566)
    y[:, 0] = 0
    got = O.normalize_medians(y)
    assert np.all(got[:, 0] == got[0, 0])


def _dense_plaid(X, xr, G, gr, stats="mean"):
    """independent restatement: plain dense numpy, explicit loops over sets"""
    D = X.toarray() if sp.issparse(X) else np.asarray(X)
    Gd = G.toarray()
    gpos = {n: k for k, n in reversed(list(enumerate(gr)))}
    out = np.zeros((G.shape[1], D.shape[1]))
    seen = set()
    rows = []
    for r, n in enumerate(xr):
        if n in seen or n not in gpos:
            continue
        seen.add(n)
        rows.append((r, gpos[n]))
    for s in range(G.shape[1]):
        mem = [r for r, g in rows if Gd[g, s] != 0]
        w = 1.0 / (1e-8 + len(mem)) if stats == "mean" else 1.0
        for r in mem:
            out[s] += D[r] * w
    return out


def test_plaid_dual_implementation_and_name_alignment():
    P, N, S = 120, 15, 12
    X = synth.sparse_x_numpy(P, N, seed=7, density=0.2)
    G = synth.genesets_numpy(150, S, seed=8, size_cap=(3, 40))
    xr = [f"g{k}" for k in range(P)]
    xr[5] = xr[3]  # duplicated rowname in X: first occurrence wins
    gr = [f"g{k}" for k in np.random.default_rng(0).permutation(200)[:150]]
    for stats in ["mean", "sum"]:
        got = O.plaid(O.Named(X, xr, None), O.Named(G, gr, None), stats=stats, normalize=False).mat
        assert rel_err(got, _dense_plaid(X, xr, G, gr, stats)) < 1e-13
    assert O.plaid(O.Named(X, ["zz%d" % k for k in range(P)], None), O.Named(G, gr, None)) is None


def test_chunked_crossprod_chunks_agree():
    X = synth.sparse_x_numpy(80, 50, seed=9, density=0.2)
    G = synth.genesets_numpy(80, 7, seed=10, size_cap=(3, 30))
    a = O.chunked_crossprod(G, X)
    b = O.chunked_crossprod(G, X, chunk=7)  # forces the loop (R/plaid.R:109-119)
    assert np.array_equal(a, b)
    assert int(round(0.8 * 2147483647 / 30000)) == 57266 and int(round(0.8 * 2147483647 / 50)) == 34359738


def test_fixture_known_answers(fixture_mats, golden):
    X, xr, xc, G, gr, gc = fixture_mats
    assert X.shape == (7728, 50) and X.nnz == 38744
    assert G.shape == (4386, 50)  # doc/plaid-vignette.html: dim(matG)
    assert golden["plaid_mean_norm"].shape == (50, 50)  # dim(gsetX)
    r = O.plaid(O.Named(X, xr, xc), O.Named(G, gr, gc))
    assert np.array_equal(r.mat, golden["plaid_mean_norm"])
    # independent dense restatement on the real fixture
    raw = _dense_plaid(X, xr, G, gr)
    assert rel_err(golden["plaid_mean_raw"], raw) < 1e-12
    assert (golden["plaid_mean_raw"] == 0).sum() > 0 and golden["plaid_mean_raw"].min() == 0  # ignore.zero case


def test_scorers_dual_on_small_input():
    """rank scorers through the closed zero-group form == literal dense restatement."""
    P, N, S = 90, 11, 9
    X = synth.sparse_x_numpy(P, N, seed=11, density=0.25)
    G = synth.genesets_numpy(P, S, seed=12, size_cap=(3, 30))
    names = [f"g{k}" for k in range(P)]
    Xn, Gn = O.Named(X, names, None), O.Named(G, names, None)
    D = X.toarray()
    Gd = G.toarray()
    w = Gd / (1e-8 + Gd.sum(0))[None, :]

    def norm(x):
        z = x.copy()
        if np.min(x) == 0:
            z[z == 0] = np.nan
        m = np.nan_to_num(np.nanmedian(z, axis=0)) if np.min(x) == 0 else np.median(x, axis=0)
        return x - m[None, :] + m.mean()

    r_min = np.column_stack([rankdata(D[:, j], method="min") for j in range(N)])
    assert rel_err(O.replaid_sing(Xn, Gn).mat, w.T @ (r_min / P - 0.5)) < 1e-12
    r_avg = np.column_stack([rankdata(D[:, j], method="average") for j in range(N)])
    u = np.minimum(r_avg.max() - r_avg, 1501)
    S_ = norm(w.T @ u)
    want = 1 - S_ / 1500 + ((Gd != 0).sum(0) + 1)[:, None] / 3000
    assert rel_err(O.replaid_ucell(Xn, Gn).mat, want) < 1e-12
    A = math.ceil(0.05 * P)
    ww = 1.08 * np.maximum((r_avg - (r_avg.max() - A)) / A, 0)
    assert rel_err(O.replaid_aucell(Xn, Gn).mat, norm(w.T @ ww)) < 1e-12
    # ssgsea: ranks among stored entries, zeros stay 0
    rs = np.zeros_like(D)
    for j in range(N):
        nz = D[:, j] != 0
        rs[nz, j] = rankdata(D[nz, j], method="average")
    for alpha in [0.0, 0.25]:
        t = rs ** (1 + alpha) if alpha else rs
        assert rel_err(O.replaid_ssgsea(Xn, Gn, alpha=alpha).mat, norm(w.T @ (t / t.max() - 0.5))) < 1e-12
    # scse
    e = D.copy()
    e[D != 0] = 2.0 ** D[D != 0]
    want = (Gd.T @ e) / (np.abs(e).sum(0) + 1e-8)[None, :] * 100
    assert rel_err(O.replaid_scse(Xn, Gn).mat, want) < 1e-12


def test_r_mean_matches_numpy():
    v = np.random.default_rng(2).normal(size=1001)
    assert abs(O.r_mean(v) - v.mean()) < 1e-15
    assert math.isnan(O.r_mean(np.array([np.nan])))


def test_plaid_test_against_scipy_ttests():
    """plaid.test restatement (R/plaid.R:392-537): its "one" and "lm" legs are textbook t-tests."""
    from scipy import stats
    P, N, S = 300, 24, 20
    X = synth.dense_x_numpy(P, N, seed=5)
    G = synth.genesets_numpy(P, S, seed=6, size_cap=(5, 60))
    names = synth.gene_names(P)
    y = np.arange(N) % 2
    tab, cols, rows = O.plaid_test(O.Named(X, names, None), y, O.Named(G, names, synth.set_names(S)))
    assert cols == ["gsetFC", "p.one", "p.two", "p.lm", "p.meta", "q.meta"] and tab.shape == (S, 6)
    assert np.all(np.diff(tab[:, cols.index("p.meta")]) >= 0)  # sorted by p.meta
    fc = X[:, y == 1].mean(1) - X[:, y == 0].mean(1)
    gx = O.plaid(O.Named(X, names, None), O.Named(G, names, None)).mat
    g = G.toarray() != 0
    for s in (0, 3, 11):
        r = rows.index(f"SET{s:05d}")
        assert abs(tab[r, 1] - stats.ttest_1samp(fc[g[:, s]], 0).pvalue) < 1e-6
        assert abs(tab[r, 3] - stats.ttest_ind(gx[s, y == 0], gx[s, y == 1], equal_var=False).pvalue) < 1e-10
    # BH adjustment
    p = np.array([0.01, 0.04, 0.03, 0.5])
    assert np.allclose(O.p_adjust_fdr(p), [0.04, 0.04 * 4 / 3, 0.04 * 4 / 3, 0.5])
