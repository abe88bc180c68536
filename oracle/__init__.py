"""CPU oracle for the plaid gene-set scoring hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s CPU-baseline / `--impl reference` legs may import this package.
The product path (`plaid_b200`, `libplaidgpu.so`) never does.

PARITY: PINNED for the default path (7 digits), PINNED TO PLOT RESOLUTION for sing / ssgsea / scse,
UNPINNED for the rest.
The reference (bigomics/plaid, pure R) cannot be executed in this environment (no R, no
rpy2) and its own test-suite pins nothing on this path (`tests/testthat/test-plaid.R:1-3`
asserts 2*2==4).  What the reference DID publish, in its built vignette (doc/plaid-vignette.html), are
the p-values printed for its bundled fixture (head(plaid.test(X, y, matG, gsetX = plaid(X, matG,
normalize = TRUE))), committed as tests/golden/vignette_known_answers.json) and a pairs() figure of
plaid / replaid.sing / replaid.ssgsea(alpha=0) / replaid.scse for cell 1 (tests/golden/vignette_pairs.png):
  * PINNED — read fixture -> gmt2mat -> row alignment -> plaid(mean) -> normalize_medians:
    the oracle reproduces all 6 printed `p.lm` values (Welch t-test on the rows of gsetX) to
    the 7 printed digits, which bounds the scores to ~1e-8 relative; and the 6 printed `p.one`
    values (gene fold changes inside each set) to 1e-6 (tests/test_reference_known_answers.py).
  * PINNED TO PLOT RESOLUTION (about 1 % of a score's range) — replaid.sing, replaid.ssgsea(alpha=0),
    replaid.scse(removeLog2=TRUE, scoreMean=FALSE), and through them colranks(ties="min") with the
    implicit-zero group and sparse_colranks(ties="average"): every one of the 600 points of the figure is
    reproduced from the oracle's values through the figure's own axis ticks, and misreadings of the
    reference (ties="average" in sing, per-column max rank or alpha=0.25 in ssgsea, scoreMean=TRUE)
    are rejected (tests/test_reference_figure.py).
  * UNPINNED — replaid.ucell / .aucell / .gsva, colranks on dense input, the other ties methods: no
    reference output exists for them.  They are restatements of `R/plaid.R` +
    the documented semantics of the CRAN/Bioconductor routines it calls (Matrix::crossprod,
    matrixStats::colRanks/colMedians/rowSds, sparseMatrixStats::colRanks/rowSds, base::rank),
    guarded by dual independent implementations per function (tests/test_oracle.py), the
    vignette's printed dims ((4386, 50) and (50, 50)) and hand-computed miniature cases.
"""
