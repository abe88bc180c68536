"""Oracle restatement of the reference's GMT helpers (TEST INFRASTRUCTURE ONLY).

Follows reference `R/gmt-utils.R:99-125` (read.gmt) and `R/gmt-utils.R:19-66` (gmt2mat).
Only used to turn the bundled `hallmarks.gmt` into the gene x geneset incidence
matrix that the hot path consumes; GMT I/O itself is out of scope.
"""
from __future__ import annotations

from collections import Counter, OrderedDict

import numpy as np
import scipy.sparse as sp


def read_gmt(path: str) -> "OrderedDict[str, list[str]]":
    """`read.gmt` (R/gmt-utils.R:99-125): name <tab> source <tab> genes...; drops "", "NA"."""
    out: "OrderedDict[str, list[str]]" = OrderedDict()
    with open(path, "r", encoding="utf-8", errors="replace") as fh:
        for line in fh:
            line = line.rstrip("\r\n")
            if not line or line.startswith("#"):
                continue
            f = line.split("\t")
            name = f[0]
            genes = " ".join(f[2:]) if len(f) >= 3 else ""
            toks = [g for g in genes.replace("\t", " ").split(" ")]
            seen, gs = set(), []
            for g in toks:  # setdiff(x, c("", "NA", NA)) also de-duplicates, keeping order
                if g in ("", "NA") or g in seen:
                    continue
                seen.add(g)
                gs.append(g)
            # a list with duplicated names keeps both entries in R; mimic with a suffix-free
            # overwrite guard: later duplicates are dropped by gmt2mat anyway (:26)
            if name not in out:
                out[name] = gs
    return out


def gmt2mat(gmt: "dict[str, list[str]]"):
    """`gmt2mat` (R/gmt-utils.R:19-66) with default arguments.

    Returns (csc_matrix genes x sets of 1.0, rownames, colnames).
    Column order: sets by decreasing size, stable (:25).  Row order: genes by decreasing
    membership count (:31, :62), ties in name order (R's `table` sorts names by the
    session collation; plain code-point order is used here — row order never affects
    scores because plaid() matches rows by name, `R/plaid.R:65-72`).
    """
    names = list(gmt.keys())
    order = sorted(range(len(names)), key=lambda k: -len(gmt[names[k]]))  # stable
    names = [names[k] for k in order]
    cnt = Counter(g for n in names for g in gmt[n])
    bg = sorted(cnt.keys())  # table(): sorted level names
    bg.sort(key=lambda g: -cnt[g])  # sort(decreasing=TRUE), stable
    pos = {g: k for k, g in enumerate(bg)}
    rows, cols = [], []
    for j, n in enumerate(names):
        for g in dict.fromkeys(gmt[n]):  # intersect(gg, s): unique
            rows.append(pos[g])
            cols.append(j)
    D = sp.csc_matrix((np.ones(len(rows)), (rows, cols)), shape=(len(bg), len(names)))
    D.sum_duplicates()
    D.data[:] = 1.0
    rs = np.asarray((D != 0).sum(axis=1)).ravel()
    o = np.argsort(-rs, kind="stable")  # :62
    D = D.tocsr()[o].tocsc()
    D.sort_indices()
    return D, [bg[k] for k in o], names
