"""CPU oracle for the plaid gene-set scoring hot path.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s CPU-baseline / `--impl reference` legs may import this package.
The product path (`plaid_b200`, `libplaidgpu.so`) never does.

PARITY UNPINNED: the reference (bigomics/plaid, pure R) cannot be executed in this
environment (no R, no rpy2) and its own test-suite pins nothing on this path
(`tests/testthat/test-plaid.R:1-3` asserts 2*2==4).  The oracle is therefore a
restatement of `R/plaid.R` + the documented semantics of the CRAN/Bioconductor
routines it calls (Matrix::crossprod, matrixStats::colRanks/colMedians/rowSds,
sparseMatrixStats::colRanks/rowSds, base::rank), guarded by
  * dual independent implementations per function (tests/test_oracle.py),
  * the two usable known answers of the reference vignette
    (dim(gmt2mat(read.gmt(hallmarks.gmt))) == (4386, 50); dim(plaid(X, matG)) == (50, 50)),
  * hand-computed miniature cases.
"""
