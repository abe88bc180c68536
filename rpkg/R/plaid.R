## Drop-in R wrappers: same signatures, defaults, messages and return types as the reference
## (R/plaid.R of bigomics/plaid; line numbers cited per function).  Only name matching,
## dimnames and argument defaults stay in R; all arithmetic runs in libplaidgpu through .Call.

.plaid_env <- new.env(parent = emptyenv())

.ctx <- function() {
  if (is.null(.plaid_env$ctx)) .plaid_env$ctx <- .Call(C_plaidgpu_ctx, 0L)
  .plaid_env$ctx
}

## options(plaid.gpus = n): the scorers split the columns of X over n devices (one context each; the axis of
## chunked_crossprod's column loop, R/plaid.R:110-119).  Results are identical for any n.
.ctxs <- function() {
  n <- as.integer(getOption("plaid.gpus", 1L))
  if (is.na(n) || n <= 1L) return(.ctx())
  have <- length(.plaid_env$ctxs)
  if (have < n) {
    if (have == 0L) .plaid_env$ctxs <- list(.ctx())
    for (d in seq.int(length(.plaid_env$ctxs), n - 1L))
      .plaid_env$ctxs[[d + 1L]] <- .Call(C_plaidgpu_ctx, as.integer(d))
  }
  .plaid_env$ctxs[seq_len(n)]
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

.score <- function(X, matG, opts, y = NULL) {
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
  if (!is.null(y)) {  # plaid.test("lm"): scores reduced per set and sample group on the device, S x 4 back
    out <- .Call(C_plaidgpu_score_group_moments, .ctxs(), x$kind, x$p, x$i, x$x, as.integer(x$dim),
                 G@p, G@i, G@x, as.integer(dim(G)), as.integer(rowmap), opts, as.integer(y))
    rownames(out) <- colnames(matG)
    return(out)
  }
  out <- .Call(C_plaidgpu_score, .ctxs(), x$kind, x$p, x$i, x$x, as.integer(x$dim),
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

## ties.method -> PLAIDGPU_TIES_*: "first" / "last" (base::rank, matrixStats) and "dense" (matrixStats) are ranked in
## order of appearance on the device; "random" draws from R's RNG and is not available
.ties <- function(m) {
  k <- match(m, c("average", "min", "max", "first", "last", "dense"))
  if (is.na(k)) stop("ties.method '", m, "' is not available on the GPU path (average, min, max, first, last, dense)")
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
  if (!rowtf %in% c("z", "ecdf")) stop("Error: unknown row transform", rowtf) # R/plaid.R:348
  .score(X, matG, list(scorer = 6L, tau = as.numeric(tau), gsva_ecdf = as.integer(rowtf == "ecdf")))
}

## plaid.test (R/plaid.R:392-474): same arguments and result table; the score matrix, the set-wise sums of
## logFC / logFC^2 and the per-set group moments come from the GPU, pt / pchisq / p.adjust stay in R.
plaid.test <- function(X, y, G, gsetX, tests = c("one", "two", "lm"),
                       metap.method = "fisher", sort.by = "p.meta") {
  if (!all(unique(y) %in% c(0, 1))) stop("elements of y must be 0 or 1")
  if (is.list(G)) G <- gmt2mat(G)                                      # R/plaid.R: a GMT list is converted first
  gg <- intersect(rownames(G), rownames(X))
  X <- X[gg, , drop = FALSE]
  G <- G[gg, , drop = FALSE]
  n1 <- sum(y == 1); n0 <- sum(y == 0)
  fc <- Matrix::rowMeans(X[, y == 1, drop = FALSE]) - Matrix::rowMeans(X[, y == 0, drop = FALSE])
  B <- 1 * (G != 0)
  sumG <- Matrix::colSums(B)
  pv <- list(); ff <- list()
  if (any(c("one", "two") %in% tests)) {
    sq <- chunked_crossprod(B, cbind(fc, fc^2))           # GPU: t(G != 0) %*% [F, F^2]
    s1 <- sq[, 1]; q1 <- sq[, 2]
  }
  if ("one" %in% tests) {
    mx <- s1 / (1e-8 + sumG)
    sdx <- sqrt((q1 - mx^2 * sumG) / (sumG - 1))
    tt <- mx / (1e-8 + sdx) * sqrt(sumG)
    pv$one <- 2 * stats::pt(abs(tt), df = pmax(sumG - 1, 1), lower.tail = FALSE); ff$one <- mx
  }
  if ("two" %in% tests) {
    sum0 <- nrow(B) - sumG
    mean1 <- s1 / (1e-8 + sumG); mean0 <- (sum(fc) - s1) / (1e-8 + sum0)
    var0 <- ((sum(fc^2) - q1) - mean0^2 * sum0) / (sum0 - 1)
    var1 <- (q1 - mean1^2 * sumG) / (sumG - 1)
    vs <- var0 / sum0 + var1 / sumG
    dof <- vs^2 / (var0 / sum0 * (sum0 - 1) + var1 / sumG * (sumG - 1))
    pv$two <- 2 * stats::pt(abs((mean1 - mean0) / sqrt(vs)), df = pmax(dof, 1), lower.tail = FALSE)
    ff$two <- mean1 - mean0
  }
  if ("lm" %in% tests) {
    if (missing(gsetX) || is.null(gsetX)) {
      ## the S x N score matrix is never materialised for the host: plaid() and the group sums run as one device call
      message("[plaid.test] computing plaid scores...")
      gm <- .score(X, G, list(scorer = 0L, stats_mean = 1L, normalize = 1L), y = y)
    } else {
      gm <- .Call(C_plaidgpu_group_moments, .ctx(), as.matrix(gsetX), as.integer(y))
    }
    m0 <- gm[, 1] / n0; m1 <- gm[, 3] / n1
    v0 <- (gm[, 2] - n0 * m0^2) / (n0 - 1); v1 <- (gm[, 4] - n1 * m1^2) / (n1 - 1)
    fac <- v0 / n0 + v1 / n1
    dof <- fac^2 / ((v0 / n0)^2 / (n0 - 1) + (v1 / n1)^2 / (n1 - 1))
    pv$lm <- 2 * stats::pt(abs((m0 - m1) / sqrt(fac)), df = dof, lower.tail = FALSE); ff$lm <- m1 - m0
  }
  pv <- lapply(pv, function(p) { p[is.na(p)] <- 1; pmin(pmax(p, 1e-99), 1 - 1e-99) })
  gsetFC <- rowMeans(do.call(cbind, ff))
  if (length(pv) > 1) {
    if (metap.method %in% c("fisher", "sumlog")) {
      pmeta <- stats::pchisq(-2 * Reduce(`+`, lapply(pv, log)), 2 * length(pv), lower.tail = FALSE)
    } else if (metap.method %in% c("stouffer", "sumz")) {
      pmeta <- stats::pnorm(Reduce(`+`, lapply(pv, stats::qnorm, lower.tail = FALSE)) / sqrt(length(pv)), lower.tail = FALSE)
    } else stop("Invalid method: ", metap.method)
  } else pmeta <- pv[[1]]
  P <- do.call(cbind, pv); colnames(P) <- paste0("p.", names(pv))
  res <- cbind(gsetFC = gsetFC, P, p.meta = pmeta, q.meta = stats::p.adjust(pmeta, method = "fdr"))
  rownames(res) <- colnames(G)
  if (sort.by %in% colnames(res)) res <- res[order(res[, sort.by]), ]
  res
}

chunked_crossprod <- function(x, y, chunk = NULL) {                         # R/plaid.R:100-123
  x <- methods::as(x, "CsparseMatrix")
  sc <- rep(1, ncol(x))                      # x must be column-scaled binary (what plaid() builds)
  nz <- diff(x@p) > 0
  sc[nz] <- x@x[x@p[-length(x@p)][nz] + 1L]
  if (is.null(chunk) || chunk < 0) chunk <- round(0.8 * .Machine$integer.max / ncol(x))
  if (NCOL(y) >= chunk) message("[chunked_crossprod] chunked compute: chunk = ", chunk)
  yy <- .as_x(y)
  out <- .Call(C_plaidgpu_crossprod, .ctx(), yy$kind, yy$p, yy$i, yy$x, as.integer(yy$dim),
               x@p, x@i, x@x, as.integer(dim(x)), as.numeric(sc))
  dimnames(out) <- list(colnames(x), yy$dimnames[[2]])
  out
}
