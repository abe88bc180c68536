"""Oracle restatement of the reference's GMT helpers (TEST INFRASTRUCTURE ONLY).

Follows reference `R/gmt-utils.R:99-125` (read.gmt) and `R/gmt-utils.R:19-66` (gmt2mat).
Only used to turn the bundled `hallmarks.gmt` into the gene x geneset incidence
matrix that the hot path consumes; GMT I/O itself is out of scope.
"""
from __future__ import annotations

from collections import Counter, OrderedDict

import numpy as np
import scipy.sparse as sp


def read_gmt(path: str) -> "list[tuple[str, list[str]]]":
    """`read.gmt` (R/gmt-utils.R:99-125): name <tab> source <tab> genes...; drops "", "NA".
    Returns (name, genes) pairs in file order: like the R list, duplicated set names are KEPT here
    (gmt2mat drops them after sorting by size, R/gmt-utils.R:25-26)."""
    out: "list[tuple[str, list[str]]]" = []
    with open(path, "r", encoding="utf-8", errors="replace") as fh:
        for line in fh:
            line = line.rstrip("\r\n")
            # utils::read.csv(sep = "!", comment.char = "#")[, 1]  (:108): '#' ends the line wherever it stands and only
            # the text before the first '!' (the separator) is kept; empty lines are skipped (quote handling not restated)
            for ch in "#!":
                k = line.find(ch)
                if k >= 0:
                    line = line[:k]
            if not line:
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
            out.append((name, gs))
    return out


def gmt2mat(gmt):
    """`gmt2mat` (R/gmt-utils.R:19-66) with default arguments.

    Returns (csc_matrix genes x sets of 1.0, rownames, colnames).
    Column order: sets by decreasing size, stable (:25).  Row order: genes by decreasing
    membership count (:31, :62), ties in name order (R's `table` sorts names by the
    session collation; plain code-point order is used here — row order never affects
    scores because plaid() matches rows by name, `R/plaid.R:65-72`).
    """
    pairs = list(gmt.items()) if isinstance(gmt, dict) else list(gmt)
    pairs.sort(key=lambda kv: -len(kv[1]))  # stable: order(-sapply(gmt, length))   (:25)
    seen_names, uniq = set(), []
    for n, gs in pairs:  # gmt[!duplicated(names(gmt))]   (:26): the largest set of a duplicated name survives
        if n not in seen_names:
            seen_names.add(n)
            uniq.append((n, gs))
    names = [n for n, _ in uniq]
    gmt = dict(uniq)
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
