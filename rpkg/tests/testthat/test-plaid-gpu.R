# Runs wherever R + the built package + a B200 exist (not in the build container: no R there).
# Same fixture as tests/test_gpu_parity.py; expected values are recomputed with the reference's own
# CPU code paths (Matrix / matrixStats), which is the literal reference.
test_that("plaid() matches Matrix::crossprod on the bundled fixture", {
  skip_if_not_installed("Matrix")
  load(system.file("extdata", "pbmc3k-50cells.rda", package = "plaid"))
  gmt <- read.gmt(system.file("extdata", "hallmarks.gmt", package = "plaid"))
  matG <- gmt2mat(gmt)
  got <- plaid(X, matG, normalize = FALSE)
  gg <- intersect(rownames(X), rownames(matG))
  G <- 1 * (matG[gg, ] != 0)
  G <- Matrix::colScale(G, 1 / (1e-8 + Matrix::colSums(G)))
  want <- as.matrix(Matrix::crossprod(G, X[gg, ]))
  expect_equal(dim(got), c(50L, 50L))
  expect_equal(unname(got), unname(want), tolerance = 1e-11)
})

test_that("sparse_colranks() is bit-exact against base::rank", {
  load(system.file("extdata", "pbmc3k-50cells.rda", package = "plaid"))
  got <- colranks(X, keep.zero = TRUE)
  want <- unlist(lapply(split(X@x, rep.int(seq_len(ncol(X)), diff(X@p))), rank))
  expect_identical(got@x, as.numeric(want))
})
