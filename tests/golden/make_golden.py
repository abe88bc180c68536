"""Generate the committed golden fixtures (run in the BUILD container, where /root/reference exists).

    python tests/golden/make_golden.py

Inputs: the reference's own bundled data (inst/extdata/pbmc3k-50cells.rda, hallmarks.gmt),
parsed without R by oracle/rdata.py + oracle/gmt.py.  Outputs: oracle/plaid_oracle.py results
on them.  The reference itself (R) cannot run here, so these vectors pin the ORACLE; the oracle's
plaid() + normalize_medians output is in turn pinned by the p-values the reference's vignette prints
(extract_vignette.py, tests/test_reference_known_answers.py), its replaid.sing / ssgsea / scse output to plot
resolution by the vignette's pairs() figure (extract_vignette_figure.py, tests/test_reference_figure.py); the other
functions are unpinned (see oracle/__init__.py).  The vectors let the GPU box — which has no /root/reference —
test against the reference's real input.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import plaid_oracle as O  # noqa: E402
from oracle.gmt import gmt2mat, read_gmt  # noqa: E402
from oracle.rdata import dgc_to_scipy, read_rda  # noqa: E402

REF = os.environ.get("PLAID_REFERENCE", "/root/reference")


def main():
    d = read_rda(os.path.join(REF, "inst/extdata/pbmc3k-50cells.rda"))
    X, xr, xc = dgc_to_scipy(d["X"])
    gmt = read_gmt(os.path.join(REF, "inst/extdata/hallmarks.gmt"))
    G, gr, gc = gmt2mat(gmt)
    assert G.shape == (4386, 50), G.shape  # vignette known answer
    Xn = O.Named(X, xr, xc)
    Gn = O.Named(G, gr, gc)
    out = {
        "X_indptr": X.indptr.astype(np.int32), "X_indices": X.indices.astype(np.int32), "X_data": X.data,
        "X_shape": np.asarray(X.shape), "X_rownames": np.asarray(xr), "X_colnames": np.asarray(xc),
        "G_indptr": G.indptr.astype(np.int32), "G_indices": G.indices.astype(np.int32),
        "G_shape": np.asarray(G.shape), "G_rownames": np.asarray(gr), "G_colnames": np.asarray(gc),
        "celltype": np.asarray(d["celltype"]),
    }
    r = O.plaid(Xn, Gn)
    assert r.mat.shape == (50, 50)  # vignette known answer dim(gsetX)
    out["plaid_mean_norm"] = r.mat
    out["plaid_mean_raw"] = O.plaid(Xn, Gn, normalize=False).mat
    out["plaid_sum_raw"] = O.plaid(Xn, Gn, stats="sum", normalize=False).mat
    out["scse_default"] = O.replaid_scse(Xn, Gn).mat
    out["scse_mean_nolog"] = O.replaid_scse(Xn, Gn, removeLog2=False, scoreMean=True).mat
    out["sing"] = O.replaid_sing(Xn, Gn).mat
    out["ssgsea_a0"] = O.replaid_ssgsea(Xn, Gn, alpha=0).mat
    out["ssgsea_a025"] = O.replaid_ssgsea(Xn, Gn, alpha=0.25).mat
    out["ucell"] = O.replaid_ucell(Xn, Gn).mat
    out["aucell"] = O.replaid_aucell(Xn, Gn).mat
    out["gsva_z"] = O.replaid_gsva(Xn, Gn).mat
    out["sparse_colranks_avg"] = O.sparse_colranks(X, ties_method="average").data
    out["sparse_colranks_min_signed"] = O.sparse_colranks(X, signed=True, ties_method="min").data
    out["colranks_dense_avg"] = O.colranks(X, ties_method="average")
    out["colranks_dense_min"] = O.colranks(X, ties_method="min")
    np.savez_compressed(os.path.join(HERE, "pbmc3k50_hallmarks.npz"), **out)
    print("wrote", os.path.join(HERE, "pbmc3k50_hallmarks.npz"))


if __name__ == "__main__":
    main()
