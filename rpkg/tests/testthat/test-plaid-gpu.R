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

test_that("ties.method first / last / dense follow base::rank and matrixStats::colRanks", {
  load(system.file("extdata", "pbmc3k-50cells.rda", package = "plaid"))
  for (tm in c("first", "last")) {
    got <- colranks(X, keep.zero = TRUE, ties.method = tm)
    want <- unlist(lapply(split(X@x, rep.int(seq_len(ncol(X)), diff(X@p))), rank, ties.method = tm))
    expect_identical(got@x, as.numeric(want))
  }
  skip_if_not_installed("matrixStats")
  D <- round(as.matrix(X[1:500, 1:10]), 1)
  for (tm in c("first", "last", "dense", "average", "min", "max"))
    expect_equal(unname(colranks(D, ties.method = tm)),
                 unname(t(matrixStats::colRanks(D, ties.method = tm))), tolerance = 0)
  expect_error(colranks(X, ties.method = "random"))
})

test_that("plaid.test without gsetX (scores reduced on the device) equals plaid.test on the materialised scores", {
  load(system.file("extdata", "pbmc3k-50cells.rda", package = "plaid"))
  matG <- gmt2mat(read.gmt(system.file("extdata", "hallmarks.gmt", package = "plaid")))
  y <- as.integer(seq_len(ncol(X)) %% 2)
  a <- plaid.test(X, y, matG, tests = "lm")
  b <- plaid.test(X, y, matG, gsetX = plaid(X, matG), tests = "lm")
  expect_equal(a[rownames(b), ], b, tolerance = 1e-12)
})

test_that("scores keep the reference's zero pattern on columns with a huge dynamic range", {
  skip_if_not_installed("Matrix")
  set.seed(1)
  X <- Matrix::rsparsematrix(2000, 40, 0.02, rand.x = function(n) rlnorm(n) * 10^sample(-6:6, n, TRUE))
  rownames(X) <- paste0("g", seq_len(nrow(X)))
  G <- Matrix::rsparsematrix(2000, 1200, 0.01, rand.x = function(n) rep(1, n))
  rownames(G) <- rownames(X); colnames(G) <- paste0("s", seq_len(ncol(G)))
  Gs <- Matrix::colScale(1 * (G != 0), 1 / (1e-8 + Matrix::colSums(G != 0)))
  want <- as.matrix(Matrix::crossprod(Gs, X))
  got <- plaid(X, G, normalize = FALSE)
  expect_identical(unname(got == 0), unname(want == 0))
  expect_equal(unname(got), unname(want), tolerance = 1e-11)
})
