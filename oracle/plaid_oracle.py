"""CPU oracle: numpy/scipy fp64 restatement of the reference's gene-set scoring hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE (see oracle/__init__.py).  PARITY: plaid() +
normalize_medians are PINNED by the p-values the reference's vignette prints for its fixture
(tests/test_reference_known_answers.py, 7 digits); replaid.sing / replaid.ssgsea(alpha=0) /
replaid.scse — and through them colranks(ties="min") with implicit zeros and
sparse_colranks(ties="average") — are pinned TO PLOT RESOLUTION (about 1 % of a score's range) by
the pairs() figure of the same vignette, the only output of the rank scorers the reference
published (tests/test_reference_figure.py); replaid.ucell / .aucell / .gsva, colranks on dense
input and the remaining ties methods are UNPINNED (the reference is R, cannot run here, and
published no outputs for them).  Every function cites
the reference lines it restates (paths relative to /root/reference) and is cross-checked by a
second independent implementation in tests/test_oracle.py.

Matrices travel as `Named(mat, rownames, colnames)` where `mat` is a
scipy.sparse.csc_matrix (R `dgCMatrix`) or a 2-D numpy array (R base matrix).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import scipy.sparse as sp

INT_MAX = 2147483647  # .Machine$integer.max


@dataclass
class Named:
    mat: object
    rownames: Optional[Sequence[str]] = None
    colnames: Optional[Sequence[str]] = None

    @property
    def shape(self):
        return self.mat.shape


def _is_sparse(m) -> bool:
    return sp.issparse(m)


def _as_csc(m) -> sp.csc_matrix:
    m = sp.csc_matrix(m)
    m.sort_indices()
    return m


# ----------------------------------------------------------------------------------
# base::rank / matrixStats::colRanks / sparseMatrixStats::colRanks semantics
# ----------------------------------------------------------------------------------
def rank_vector(v: np.ndarray, ties: str = "average") -> np.ndarray:
    """base::rank(v, ties.method=ties, na.last="keep") for ties in {average,min,max,first,last} and the "dense"
    method of matrixStats::colRanks.

    Explicit stable-sort + tie-run pass (the second implementation, scipy.stats.rankdata,
    lives in tests/test_oracle.py).  Ties are exact fp64 equality (-0 == +0); NaN -> NaN and
    does not take part in the ranking (reference call sites: R/plaid.R:639,642 via
    base::rank; :605,608 sparseMatrixStats::colRanks; :614,617 matrixStats::colRanks).
    """
    v = np.asarray(v, dtype=np.float64)
    out = np.full(v.shape, np.nan)
    ok = ~np.isnan(v)
    w = v[ok]
    n = w.size
    if n == 0:
        return out
    o = np.argsort(w, kind="stable")
    s = w[o]
    head = np.empty(n, dtype=bool)
    head[0] = True
    head[1:] = s[1:] != s[:-1]
    start = np.flatnonzero(head)  # 0-based first position of each tie run
    end = np.append(start[1:], n)  # exclusive
    run = np.cumsum(head) - 1
    if ties == "average":
        r = (start[run] + 1 + end[run]) / 2.0
    elif ties == "min":
        r = (start[run] + 1).astype(np.float64)
    elif ties == "max":
        r = end[run].astype(np.float64)
    elif ties == "first":  # ties in order of appearance: the stable order itself
        r = np.arange(1, n + 1, dtype=np.float64)
    elif ties == "last":  # ... in reverse order of appearance
        r = (start[run] + end[run] - np.arange(n)).astype(np.float64)
    elif ties == "dense":  # matrixStats::colRanks only: consecutive ranks of the distinct values
        r = (run + 1).astype(np.float64)
    else:  # "random" draws from R's RNG
        raise ValueError(f"unsupported ties.method {ties!r}")
    res = np.empty(n)
    res[o] = r
    out[ok] = res
    return out


def sparse_colranks(X: sp.csc_matrix, signed: bool = False, ties_method: str = "average") -> sp.csc_matrix:
    """`sparse_colranks` (R/plaid.R:631-650): rank only the STORED entries of each CSC
    column (explicit zeros are ranked like any other stored value); pattern is kept."""
    X = _as_csc(X)
    out = X.copy().astype(np.float64)
    p = X.indptr
    for j in range(X.shape[1]):
        v = X.data[p[j]:p[j + 1]]
        if signed:  # :637-640
            out.data[p[j]:p[j + 1]] = np.sign(v) * rank_vector(np.abs(v), ties_method)
        else:  # :642
            out.data[p[j]:p[j + 1]] = rank_vector(v, ties_method)
    return out


def dense_colranks(M: np.ndarray, ties_method: str = "average") -> np.ndarray:
    """t(matrixStats::colRanks(M, ties.method)) == column-wise rank, shape preserved
    (R/plaid.R:614,617; same semantics for sparseMatrixStats::colRanks at :605,608 where
    implicit zeros take part as the value 0)."""
    M = np.asarray(M, dtype=np.float64)
    out = np.empty(M.shape)
    for j in range(M.shape[1]):
        out[:, j] = rank_vector(M[:, j], ties_method)
    return out


def colranks(X, sparse: Optional[bool] = None, signed: bool = False, keep_zero: bool = False,
             ties_method: str = "average"):
    """`colranks` (R/plaid.R:589-623).  Returns csc_matrix for (sparse & keep.zero), else
    a dense ndarray (the reference returns the dense N x P colRanks transposed)."""
    if sparse is None:
        sparse = _is_sparse(X)  # :595-596
    if sparse:
        X = _as_csc(X)  # :599
        if keep_zero:
            return sparse_colranks(X, signed=signed, ties_method=ties_method)  # :601
        D = X.toarray()
        if signed:  # :603-606  (sign(0) == 0 -> zeros stay 0)
            return dense_colranks(np.abs(D), ties_method) * np.sign(D)
        return dense_colranks(D, ties_method)  # :608
    D = X.toarray() if _is_sparse(X) else np.asarray(X, dtype=np.float64)
    if signed:  # :612-615
        return np.sign(D) * dense_colranks(np.abs(D), ties_method)
    return dense_colranks(D, ties_method)  # :617


# ----------------------------------------------------------------------------------
# normalize_medians
# ----------------------------------------------------------------------------------
def col_medians_narm(x: np.ndarray) -> np.ndarray:
    """matrixStats::colMedians(x, na.rm=TRUE): NaN dropped; even count -> mean of the two
    middle values; empty -> NaN (R/plaid.R:565,569)."""
    x = np.asarray(x, dtype=np.float64)
    out = np.full(x.shape[1], np.nan)
    for j in range(x.shape[1]):
        c = x[:, j]
        c = np.sort(c[~np.isnan(c)])
        m = c.size
        if m == 0:
            continue
        out[j] = c[m // 2] if (m & 1) else (c[m // 2 - 1] + c[m // 2]) / 2.0
    return out


def r_mean(v: np.ndarray) -> float:
    """base::mean(v, na.rm=TRUE): long-double sum / n, then one refinement pass
    (R's summary.c real_mean)."""
    v = np.asarray(v, dtype=np.float64)
    v = v[~np.isnan(v)]
    if v.size == 0:
        return float("nan")
    ld = np.longdouble  # sequential (cumsum) like R's LDOUBLE loop, not numpy's pairwise sum
    w = v.astype(ld)
    s = np.cumsum(w)[-1] / ld(v.size)
    t = np.cumsum(w - s)[-1] / ld(v.size)
    return float(s + t)


def normalize_medians(x: np.ndarray, ignore_zero: Optional[bool] = None) -> np.ndarray:
    """`normalize_medians` (R/plaid.R:554-575)."""
    x = np.array(x, dtype=np.float64, copy=True)
    if x.ndim == 1:
        x = x[:, None]
    if ignore_zero is None:  # :556-557  GLOBAL over the whole matrix
        ignore_zero = bool(np.nanmin(x) == 0) if np.any(~np.isnan(x)) else False
    if ignore_zero:  # :561-566
        zx = x.copy()
        zx[x == 0] = np.nan
        medx = col_medians_narm(zx)
        medx[np.isnan(medx)] = 0.0
    else:  # :569
        medx = col_medians_narm(x)
    return (x - medx[None, :]) + r_mean(medx)  # :572


# ----------------------------------------------------------------------------------
# plaid / chunked_crossprod
# ----------------------------------------------------------------------------------
def chunked_crossprod(x, y, chunk: Optional[int] = None) -> np.ndarray:
    """`chunked_crossprod` (R/plaid.R:100-123): t(x) %*% y, column-chunked when
    ncol(y) >= chunk.  Always returns a dense ndarray (the reference returns a
    Matrix-class object on the un-chunked branch :107; values are identical)."""
    ncx = x.shape[1]
    if chunk is None or chunk < 0:  # :101-105
        chunk = int(round(0.8 * INT_MAX / ncx))  # R round(): half-even, as Python's
    xt = (x.T.tocsr() if _is_sparse(x) else np.asarray(x).T)

    def _prod(yy):
        r = xt @ yy
        return r.toarray() if _is_sparse(r) else np.asarray(r)

    if y.shape[1] < chunk:  # :107
        return _prod(y)
    k = int(math.ceil(y.shape[1] / chunk))  # :110
    out = np.full((ncx, y.shape[1]), np.nan)
    ycsc = y.tocsc() if _is_sparse(y) else y
    for i in range(k):  # :115-119
        j0, j1 = i * chunk, min(y.shape[1], (i + 1) * chunk)
        out[:, j0:j1] = _prod(ycsc[:, j0:j1])
    return out


def _intersect_rows(xr: Sequence[str], gr: Sequence[str]):
    """gg <- intersect(rownames(X), rownames(matG)) (R/plaid.R:65) and the row indices that
    `X[gg,]` / `matG[gg,]` select (first occurrence of each name, :71-72)."""
    gpos = {}
    for k, n in enumerate(gr):
        gpos.setdefault(n, k)
    seen = set()
    xi, gi = [], []
    for k, n in enumerate(xr):
        if n in seen:
            continue
        seen.add(n)
        if n in gpos:
            xi.append(k)
            gi.append(gpos[n])
    return np.asarray(xi, dtype=np.int64), np.asarray(gi, dtype=np.int64)


def plaid(X: Named, matG: Named, stats: str = "mean", chunk=None, normalize: bool = True):
    """`plaid` (R/plaid.R:60-87).  Returns Named(dense S x N) or None (no overlap, :66-69)."""
    Xm = X.mat
    if getattr(Xm, "ndim", 2) == 1:  # :63
        Xm = np.asarray(Xm, dtype=np.float64)[:, None]
    xi, gi = _intersect_rows(X.rownames, matG.rownames)
    if xi.size == 0:
        return None
    if _is_sparse(Xm):
        Xs = _as_csc(Xm).tocsr()[xi].tocsc()
    else:
        Xs = np.asarray(Xm, dtype=np.float64)[xi]
    Gm = _as_csc(matG.mat).tocsr()[gi].tocsc()
    G = Gm.copy().astype(np.float64)
    G.data = (G.data != 0).astype(np.float64)  # :73  1*(matG != 0)
    G.eliminate_zeros()
    if stats == "mean":  # :74-77
        sumG = 1e-8 + np.asarray(G.sum(axis=0)).ravel()
        G = G @ sp.diags(1.0 / sumG)  # colScale: scale BEFORE the product
    elif stats != "sum":
        # any other string silently behaves like "sum" in the reference (:74); keep that
        pass
    gsetX = chunked_crossprod(sp.csc_matrix(G), Xs, chunk=None)  # :80 (plaid's chunk arg is dead)
    if normalize:
        gsetX = normalize_medians(gsetX)  # :83
    return Named(gsetX, list(matG.colnames) if matG.colnames is not None else None,
                 list(X.colnames) if X.colnames is not None else None)


# ----------------------------------------------------------------------------------
# replaid.* scorers
# ----------------------------------------------------------------------------------
def _gmin(m):
    if _is_sparse(m):
        m = _as_csc(m)
        v = np.nanmin(m.data) if m.nnz else np.inf
        return min(v, 0.0) if m.nnz < m.shape[0] * m.shape[1] else v
    return np.nanmin(m)


def _gmax(m):
    if _is_sparse(m):
        m = _as_csc(m)
        v = np.nanmax(m.data) if m.nnz else -np.inf
        return max(v, 0.0) if m.nnz < m.shape[0] * m.shape[1] else v
    return np.nanmax(m)


def replaid_scse(X: Named, matG: Named, removeLog2: Optional[bool] = None, scoreMean: bool = False):
    """`replaid.scse` (R/plaid.R:155-190)."""
    Xm = X.mat
    if removeLog2 is None:  # :160-161
        removeLog2 = bool(_gmin(Xm) == 0 and _gmax(Xm) < 20)
    if removeLog2:
        if _is_sparse(Xm):  # :165-166   2**X@x on ALL stored entries (explicit zeros -> 1)
            Xm = _as_csc(Xm).copy().astype(np.float64)
            Xm.data = np.exp2(Xm.data)
        else:  # :168-169  only where X > 0
            Xm = np.array(Xm, dtype=np.float64, copy=True)
            nz = Xm > 0
            Xm[nz] = np.exp2(Xm[nz])
    Xn = Named(Xm, X.rownames, X.colnames)
    absX = abs(Xm)
    if scoreMean:  # :172-176
        s = plaid(Xn, matG, stats="mean", normalize=False)
        if s is None:
            return None
        sumx = np.asarray(absX.mean(axis=0)).ravel() + 1e-8
        out = s.mat * (1.0 / sumx)[None, :]
    else:  # :178-182
        s = plaid(Xn, matG, stats="sum", normalize=False)
        if s is None:
            return None
        sumx = np.asarray(absX.sum(axis=0)).ravel() + 1e-8
        out = s.mat * (1.0 / sumx)[None, :] * 100
    return Named(out, s.rownames, X.colnames)


def replaid_sing(X: Named, matG: Named):
    """`replaid.sing` (R/plaid.R:213-219)."""
    rX = colranks(X.mat, ties_method="min")
    rX = rX / X.mat.shape[0] - 0.5
    return plaid(Named(rX, X.rownames, X.colnames), matG, normalize=False)


def replaid_ssgsea(X: Named, matG: Named, alpha: float = 0.0):
    """`replaid.ssgsea` (R/plaid.R:244-255)."""
    rX = colranks(X.mat, keep_zero=True, ties_method="average")
    if _is_sparse(rX):
        rX = rX.toarray()  # "- 0.5" at :251 densifies
    if alpha != 0:
        rX = rX ** (1 + alpha)  # :246-250
    rX = rX / np.nanmax(rX) - 0.5  # :251 (max(): NA would propagate; inputs here are finite)
    return plaid(Named(rX, X.rownames, X.colnames), matG, stats="mean", normalize=True)


def replaid_ucell(X: Named, matG: Named, rmax: float = 1500):
    """`replaid.ucell` (R/plaid.R:276-282)."""
    rX = colranks(X.mat, ties_method="average")
    rX = np.minimum(np.max(rX) - rX, rmax + 1)
    S = plaid(Named(rX, X.rownames, X.colnames), matG)
    if S is None:
        return None
    Gm = _as_csc(matG.mat)
    gsz = np.asarray((Gm != 0).sum(axis=0)).ravel()  # colSums over ALL rows of matG (:280)
    out = 1 - S.mat / rmax + ((gsz + 1) / (2 * rmax))[:, None]
    return Named(out, S.rownames, S.colnames)


def replaid_aucell(X: Named, matG: Named, aucMaxRank: Optional[float] = None):
    """`replaid.aucell` (R/plaid.R:304-309)."""
    if aucMaxRank is None:
        aucMaxRank = math.ceil(0.05 * X.mat.shape[0])
    rX = colranks(X.mat, ties_method="average")
    ww = 1.08 * np.maximum((rX - (np.max(rX) - aucMaxRank)) / aucMaxRank, 0)
    return plaid(Named(ww, X.rownames, X.colnames), matG, stats="mean")


def mat_rowsds(Xm) -> np.ndarray:
    """`mat.rowsds` (R/plaid.R:365-370): sample SD (n-1) per row, two-pass."""
    D = Xm.toarray() if _is_sparse(Xm) else np.asarray(Xm, dtype=np.float64)
    n = D.shape[1]
    mu = row_means(D)[:, None]
    if n < 2:
        return np.full(D.shape[0], np.nan)
    ss = ((D - mu) ** 2).astype(np.longdouble).sum(axis=1).astype(np.float64)
    return np.sqrt(ss / (n - 1))


def row_means(D: np.ndarray) -> np.ndarray:
    """rowMeans(X): R accumulates each row in long double, then divides by ncol."""
    return (D.astype(np.longdouble).sum(axis=1).astype(np.float64)) / D.shape[1]


def replaid_gsva(X: Named, matG: Named, tau: float = 0.0, rowtf: str = "z"):
    """`replaid.gsva` (R/plaid.R:338-363), rowtf in {"z", "ecdf"}."""
    D = X.mat.toarray() if _is_sparse(X.mat) else np.asarray(X.mat, dtype=np.float64)
    if rowtf == "z":  # :341-343
        zX = (D - row_means(D)[:, None]) / (1e-8 + mat_rowsds(D))[:, None]
    elif rowtf == "ecdf":  # :344-346  ecdf(x)(x) = fraction of row values <= x
        zX = np.empty_like(D)
        for g in range(D.shape[0]):
            s = np.sort(D[g])
            zX[g] = np.searchsorted(s, D[g], side="right") / D.shape[1]
    else:
        raise ValueError("Error: unknown row transform" + str(rowtf))
    rX = colranks(zX, signed=True, ties_method="average")  # dense branch (:612-615)
    rX = rX / np.max(np.abs(rX))  # :352
    if tau > 0:
        rX = np.sign(rX) * np.abs(rX) ** (1 + tau)  # :356
    return plaid(Named(rX, X.rownames, X.colnames), matG)


# ----------------------------------------------------------------------------------
# plaid.test  ("next" row f1 of the scope table; the statistics downstream of the scores)
# ----------------------------------------------------------------------------------
def _pt_upper2(t, df):
    """2 * pt(|t|, df, lower.tail = FALSE)"""
    from scipy import stats
    return 2.0 * stats.t.sf(np.abs(t), df)


def matrix_onesample_ttest(F: np.ndarray, G) -> dict:
    """`matrix_onesample_ttest` (R/plaid.R:476-486); F: genes x k, G: genes x sets."""
    Gb = (_as_csc(G) != 0).astype(np.float64)
    F = np.asarray(F, dtype=np.float64).reshape(Gb.shape[0], -1)
    sumG = np.asarray(Gb.sum(axis=0)).ravel()
    sum_sq = np.asarray(Gb.T @ (F ** 2))
    meanx = np.asarray(Gb.T @ F) / (1e-8 + sumG)[:, None]
    with np.errstate(invalid="ignore", divide="ignore"):
        sdx = np.sqrt((sum_sq - meanx ** 2 * sumG[:, None]) / (sumG - 1)[:, None])
        t = meanx / (1e-8 + sdx) * np.sqrt(sumG)[:, None]
    p = _pt_upper2(t, np.maximum(sumG - 1, 1)[:, None])
    return {"mean": meanx, "t": t, "p": p}


def matrix_twosample_ttest(F: np.ndarray, G) -> dict:
    """`matrix_twosample_ttest` (R/plaid.R:488-520)."""
    Gb = (_as_csc(G) != 0).astype(np.float64)
    F = np.asarray(F, dtype=np.float64).reshape(Gb.shape[0], -1)
    sum1 = np.asarray(Gb.sum(axis=0)).ravel()[:, None]
    sum0 = Gb.shape[0] - sum1
    F2 = F ** 2
    ssq1 = np.asarray(Gb.T @ F2)
    ssq0 = -ssq1 + F2.sum(axis=0)[None, :]
    mean1 = np.asarray(Gb.T @ F)
    mean0 = -mean1 + F.sum(axis=0)[None, :]
    mean1 = mean1 / (1e-8 + sum1)
    mean0 = mean0 / (1e-8 + sum0)
    with np.errstate(invalid="ignore", divide="ignore"):
        var0 = (ssq0 - mean0 ** 2 * sum0) / (sum0 - 1)
        var1 = (ssq1 - mean1 ** 2 * sum1) / (sum1 - 1)
        varsum = var0 / sum0 + var1 / sum1
        dof = varsum ** 2 / (var0 / sum0 * (sum0 - 1) + var1 / sum1 * (sum1 - 1))  # exactly as written at :510
        f = mean1 - mean0
        t = f / np.sqrt(varsum)
    p = _pt_upper2(t, np.maximum(dof, 1))
    return {"diff": f, "t": t, "p": p}


def ttests_welch(M: np.ndarray, y: np.ndarray) -> dict:
    """Rfast::ttests(t(M), ina = y + 1) (R/plaid.R:429): per row of M a Welch two-sample t-test between the
    columns with y == 0 (ina 1) and y == 1 (ina 2)."""
    M = np.asarray(M, dtype=np.float64)
    a, b = M[:, y == 0], M[:, y == 1]
    n1, n2 = a.shape[1], b.shape[1]
    m1, m2 = a.mean(axis=1), b.mean(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        v1 = ((a ** 2).sum(axis=1) - n1 * m1 ** 2) / (n1 - 1)
        v2 = ((b ** 2).sum(axis=1) - n2 * m2 ** 2) / (n2 - 1)
        fac = v1 / n1 + v2 / n2
        stat = (m1 - m2) / np.sqrt(fac)
        dof = fac ** 2 / ((v1 / n1) ** 2 / (n1 - 1) + (v2 / n2) ** 2 / (n2 - 1))
    return {"stat": stat, "pvalue": _pt_upper2(stat, dof), "dof": dof}


def p_adjust_fdr(p: np.ndarray) -> np.ndarray:
    """stats::p.adjust(p, method = "fdr") (Benjamini-Hochberg)."""
    p = np.asarray(p, dtype=np.float64)
    n = p.size
    o = np.argsort(-p, kind="stable")
    ro = np.argsort(o, kind="stable")
    q = np.minimum.accumulate(p[o] * n / np.arange(n, 0, -1))
    return np.minimum(q, 1.0)[ro]


def plaid_test(X: Named, y, G: Named, gsetX: Optional[Named] = None, tests=("one", "two", "lm"),
               metap_method: str = "fisher", sort_by: str = "p.meta"):
    """`plaid.test` (R/plaid.R:392-474).  Returns (table ndarray, column names, row names) sorted by `sort_by`."""
    from scipy import stats
    y = np.asarray(y)
    if not np.all(np.isin(np.unique(y), [0, 1])):
        raise ValueError("elements of y must be 0 or 1")
    gg = [g for g in dict.fromkeys(G.rownames) if g in set(X.rownames)]  # intersect(rownames(G), rownames(X)), :402
    xpos = {}
    for k, n in enumerate(X.rownames):
        xpos.setdefault(n, k)
    gpos = {}
    for k, n in enumerate(G.rownames):
        gpos.setdefault(n, k)
    xi = np.array([xpos[g] for g in gg])
    gi = np.array([gpos[g] for g in gg])
    Xm = X.mat.tocsr()[xi].toarray() if _is_sparse(X.mat) else np.asarray(X.mat, dtype=np.float64)[xi]
    Gm = _as_csc(G.mat).tocsr()[gi].tocsc()
    fc = Xm[:, y == 1].mean(axis=1) - Xm[:, y == 0].mean(axis=1)  # :406-408
    P, Fs = {}, {}
    if "one" in tests:
        r = matrix_onesample_ttest(fc, Gm)
        P["one"], Fs["one"] = r["p"][:, 0], r["mean"][:, 0]
    if "two" in tests:
        r = matrix_twosample_ttest(fc, Gm)
        P["two"], Fs["two"] = r["p"][:, 0], r["diff"][:, 0]
    if "lm" in tests:
        if gsetX is None:
            gsetX = plaid(Named(Xm, gg, X.colnames), Named(Gm, gg, G.colnames))
        r = ttests_welch(gsetX.mat, y)
        P["lm"] = r["pvalue"]
        Fs["lm"] = gsetX.mat[:, y == 1].mean(axis=1) - gsetX.mat[:, y == 0].mean(axis=1)
    for k in P:  # :440-445
        p1 = np.where(np.isnan(P[k]), 1.0, P[k])
        P[k] = np.minimum(np.maximum(p1, 1e-99), 1 - 1e-99)
    keys = [k for k in ("one", "two", "lm") if k in P]
    Fm = np.column_stack([Fs[k] for k in keys])
    gsetFC = Fm.mean(axis=1)
    if len(keys) > 1:  # matrix_combine_p (:522-537)
        if metap_method in ("fisher", "sumlog"):
            chisq = -2.0 * sum(np.log(P[k]) for k in keys)
            pmeta = stats.chi2.sf(chisq, 2 * len(keys))
        elif metap_method in ("stouffer", "sumz"):
            zz = sum(stats.norm.isf(P[k]) for k in keys) / math.sqrt(len(keys))
            pmeta = stats.norm.sf(zz)
        else:
            raise ValueError("Invalid method: " + metap_method)
    else:
        pmeta = P[keys[0]]
    qmeta = p_adjust_fdr(pmeta)
    cols = ["gsetFC"] + ["p." + k for k in keys] + ["p.meta", "q.meta"]
    tab = np.column_stack([gsetFC] + [P[k] for k in keys] + [pmeta, qmeta])
    rows = list(G.colnames) if G.colnames is not None else [str(k) for k in range(tab.shape[0])]
    if sort_by in cols:
        o = np.argsort(tab[:, cols.index(sort_by)], kind="stable")
        tab, rows = tab[o], [rows[k] for k in o]
    return tab, cols, rows
