"""Known answers the REFERENCE ITSELF published for the hot path (tests/golden/vignette_known_answers.json,
transcribed by tests/golden/extract_vignette.py from the reference's built vignette doc/plaid-vignette.html):
for its bundled fixture (pbmc3k-50cells.rda x hallmarks.gmt) it prints head(plaid.test(X, y, matG, gsetX =
plaid(X, matG, normalize = TRUE), tests = c("one", "lm"))).

  * `p.lm` is the Welch t-test of B vs T cells on each row of gsetX: a function of all 50 normalised scores of a
    set.  A relative perturbation of 1e-8 of the scores moves these p-values by ~1e-6, so agreeing with the 7
    printed digits pins read-fixture -> gmt2mat -> row alignment -> plaid() -> normalize_medians to about 1e-8
    relative — this is what pins the oracle (and, in the -m gpu test, the CUDA path) to real reference output.
  * `p.one` is the one-sample t-test on the per-gene fold changes inside each set: pins the membership / alignment
    of f2 (GMT ingestion) and the reductions of f1 (plaid.test).  Printed by a revision without today's 1e-8
    regulariser in t (R/plaid.R:482), which moves p by ~2e-7: tolerance 1e-6.
  * gsetFC / p.meta / q.meta were printed by an older plaid.test (other fold-change summary and combination
    rule) and are not used.
"""
import json
import os

import numpy as np
import pytest
import scipy.stats as st

from oracle import plaid_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
KA = json.load(open(os.path.join(HERE, "golden", "vignette_known_answers.json")))
SETS = list(KA["head_res"])
P_LM = np.array([KA["head_res"][s]["p.lm"] for s in SETS])
P_ONE = np.array([KA["head_res"][s]["p.one"] for s in SETS])
PRINT_TOL = 6e-7  # 7 significant digits printed: half a unit in the last place is <= 5e-7 relative


def welch_p(scores, rownames, y):
    idx = [rownames.index(s) for s in SETS]
    a = np.asarray(scores)[idx]
    return st.ttest_ind(a[:, y == 1], a[:, y == 0], axis=1, equal_var=False).pvalue


def test_printed_dims(fixture_mats):
    X, xr, xc, G, gr, gc = fixture_mats
    assert KA["dims_printed"] == [list(G.shape), [G.shape[1], X.shape[1]]]  # dim(matG), dim(gsetX)


def test_oracle_reproduces_the_reference_printed_p_values(fixture_mats, golden):
    X, xr, xc, G, gr, gc = fixture_mats
    y = (golden["celltype"] == "B").astype(int)  # y <- 1 * (celltype == "B")
    gs = O.plaid(O.Named(X, xr, xc), O.Named(G, gr, gc), normalize=True)
    p = welch_p(gs.mat, gs.rownames, y)
    assert np.max(np.abs(p / P_LM - 1.0)) < PRINT_TOL
    # the committed golden scores are these very scores
    assert np.array_equal(gs.mat, golden["plaid_mean_norm"])
    tab, cols, rows = O.plaid_test(O.Named(X, xr, xc), y, O.Named(G, gr, gc), gsetX=gs, tests=("one", "lm"))
    idx = [rows.index(s) for s in SETS]
    assert np.max(np.abs(tab[idx, cols.index("p.lm")] / P_LM - 1.0)) < PRINT_TOL
    assert np.max(np.abs(tab[idx, cols.index("p.one")] / P_ONE - 1.0)) < 1e-6
    # un-normalised scores do NOT reproduce the table: the check is sensitive to the normalisation step
    raw = O.plaid(O.Named(X, xr, xc), O.Named(G, gr, gc), normalize=False)
    assert np.max(np.abs(welch_p(raw.mat, raw.rownames, y) / P_LM - 1.0)) > 0.1


@pytest.mark.gpu
def test_cuda_path_reproduces_the_reference_printed_p_values(fixture_mats, golden, gpu_ctx):
    import plaid_b200 as pb
    X, xr, xc, G, gr, gc = fixture_mats
    y = (golden["celltype"] == "B").astype(int)
    Xn, Gn = pb.NamedMatrix(X, xr, xc), pb.NamedMatrix(G, gr, gc)
    gs = pb.plaid(Xn, Gn, normalize=True, ctx=gpu_ctx)
    assert np.max(np.abs(welch_p(gs.mat, list(gs.rownames), y) / P_LM - 1.0)) < PRINT_TOL
    tab, cols, rows = pb.plaid_test(Xn, y, Gn, gsetX=gs, tests=("one", "lm"), ctx=gpu_ctx)
    idx = [rows.index(s) for s in SETS]
    assert np.max(np.abs(tab[idx, cols.index("p.lm")] / P_LM - 1.0)) < PRINT_TOL
    assert np.max(np.abs(tab[idx, cols.index("p.one")] / P_ONE - 1.0)) < 1e-6
