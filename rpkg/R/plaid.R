## Drop-in R wrappers: same signatures, defaults, messages and return types as the reference
## (R/plaid.R of bigomics/plaid; line numbers cited per function).  Only name matching,
## dimnames and argument defaults stay in R; all arithmetic runs in libplaidgpu through .Call.

.plaid_env <- new.env(parent = emptyenv())

.ctx <- function() {
  if (is.null(.plaid_env$ctx)) .plaid_env$ctx <- .Call(C_plaidgpu_ctx, 0L)
  .plaid_env$ctx
}

## X -> list(kind, p, i, x, dim): dgCMatrix slots are passed as they are (no copy)
.as_x <- function(X) {
  if (is.null(dim(X))) X <- cbind(X)                                  # R/plaid.R:63
  if (inherits(X, "CsparseMatrix")) {
    X <- methods::as(X, "CsparseMatrix")
    if (!inherits(X, "dgCMatrix")) X <- methods::as(X, "generalMatrix")
    list(kind = 0L, p = X@p, i = X@i, x = X@x, dim = dim(X), dimnames = dimnames(X))
  } else {
    X <- as.matrix(X)
    storage.mode(X) <- "double"
    list(kind = 1L, p = NULL, i = NULL, x = X, dim = dim(X), dimnames = dimnames(X))
  }
}

.score <- function(X, matG, opts) {
  x <- .as_x(X)
  rn <- x$dimnames[[1]]
  gg <- intersect(rn, rownames(matG))                                 # R/plaid.R:65
  if (length(gg) == 0) {
    message("[plaid] ERROR. No overlapping features.")               # R/plaid.R:66-69
    return(NULL)
  }
  ## rowmap[r] = 0-based row of matG aligned with X row r, or -1; first occurrences only,
  ## exactly what X[gg,] / matG[gg,] select (R/plaid.R:71-72)
  rowmap <- match(rn, rownames(matG)) - 1L
  rowmap[is.na(rowmap) | duplicated(rn)] <- -1L
  G <- methods::as(matG, "CsparseMatrix")
  if (!inherits(G, "dgCMatrix")) G <- methods::as(methods::as(G, "dMatrix"), "generalMatrix")
  out <- .Call(C_plaidgpu_score, .ctx(), x$kind, x$p, x$i, x$x, as.integer(x$dim),
               G@p, G@i, G@x, as.integer(dim(G)), as.integer(rowmap), opts)
  dimnames(out) <- list(colnames(matG), x$dimnames[[2]])
  out
}

plaid <- function(X, matG, stats = c("mean", "sum"), chunk = NULL, normalize = TRUE) {
  stats <- stats[1]                                                   # R/plaid.R:62
  ## `chunk` is accepted and ignored exactly like the reference (R/plaid.R:80 passes NULL)
  .score(X, matG, list(scorer = 0L, stats_mean = as.integer(stats == "mean"),
                       normalize = as.integer(isTRUE(normalize))))
}

normalize_medians <- function(x, ignore.zero = NULL) {                # R/plaid.R:554-575
  x <- as.matrix(x)
  storage.mode(x) <- "double"
  iz <- if (is.null(ignore.zero)) -1L else as.integer(isTRUE(ignore.zero))
  out <- .Call(C_plaidgpu_normalize_medians, .ctx(), x, iz)
  dimnames(out) <- dimnames(x)
  out
}

sparse_colranks <- function(X, signed = FALSE, ties.method = "average") {   # R/plaid.R:631-650
  X <- methods::as(X, "CsparseMatrix")
  rX <- X
  rX@x <- .Call(C_plaidgpu_colranks, .ctx(), 0L, X@p, X@i, X@x, as.integer(dim(X)),
                .ties(ties.method), as.integer(signed), 1L)
  rX
}

.ties <- function(m) {
  k <- match(m, c("average", "min", "max"))
  if (is.na(k)) stop("ties.method '", m, "' is not available on the GPU path (average, min, max)")
  k - 1L
}

colranks <- function(X, sparse = NULL, signed = FALSE, keep.zero = FALSE,   # R/plaid.R:589-623
                     ties.method = "average") {
  if (is.null(sparse)) sparse <- inherits(X, "CsparseMatrix")
  if (sparse) {
    X <- methods::as(X, "CsparseMatrix")
    if (keep.zero) return(sparse_colranks(X, signed = signed, ties.method = ties.method))
    r <- .Call(C_plaidgpu_colranks, .ctx(), 0L, X@p, X@i, X@x, as.integer(dim(X)),
               .ties(ties.method), as.integer(signed), 0L)
  } else {
    M <- as.matrix(X)
    storage.mode(M) <- "double"
    r <- .Call(C_plaidgpu_colranks, .ctx(), 1L, NULL, NULL, M, as.integer(dim(M)),
               .ties(ties.method), as.integer(signed), 0L)
  }
  r <- matrix(r, nrow = nrow(X), ncol = ncol(X), dimnames = dimnames(X))
  r
}

replaid.scse <- function(X, matG, removeLog2 = NULL, scoreMean = FALSE) {   # R/plaid.R:155-190
  rl <- if (is.null(removeLog2)) -1L else as.integer(isTRUE(removeLog2))
  if (isTRUE(removeLog2)) message("[replaid.scse] Converting data to linear scale (removing log2)...")
  .score(X, matG, list(scorer = 1L, remove_log2 = rl, score_mean = as.integer(isTRUE(scoreMean))))
}

replaid.sing <- function(X, matG) {                                         # R/plaid.R:213-219
  .score(X, matG, list(scorer = 2L, nrow_x = NROW(X)))
}

replaid.ssgsea <- function(X, matG, alpha = 0) {                            # R/plaid.R:244-255
  .score(X, matG, list(scorer = 3L, alpha = as.numeric(alpha)))
}

replaid.ucell <- function(X, matG, rmax = 1500) {                           # R/plaid.R:276-282
  .score(X, matG, list(scorer = 4L, rmax = as.numeric(rmax),
                       matg_full_colsums = as.numeric(Matrix::colSums(matG != 0))))
}

replaid.aucell <- function(X, matG, aucMaxRank = ceiling(0.05 * nrow(X))) { # R/plaid.R:304-309
  .score(X, matG, list(scorer = 5L, auc_max_rank = as.numeric(aucMaxRank)))
}

replaid.gsva <- function(X, matG, tau = 0, rowtf = c("z", "ecdf")[1]) {      # R/plaid.R:338-363
  rowtf <- rowtf[1]
  if (rowtf == "ecdf") stop("replaid.gsva(rowtf = 'ecdf') is not available on the GPU path")
  if (rowtf != "z") stop("Error: unknown row transform", rowtf)              # R/plaid.R:348
  .score(X, matG, list(scorer = 6L, tau = as.numeric(tau)))
}
