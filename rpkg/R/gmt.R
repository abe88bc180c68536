## GMT helpers with the reference's signatures and conventions (R/gmt-utils.R of bigomics/plaid; line numbers
## cited per function), so that the reference's own flow read.gmt() -> gmt2mat() -> plaid() runs after the swap.
## Text handling stays on the host in R, like in the reference; the bulk path (tens of thousands of sets) is
## plaidgpu_gmt_read / plaidgpu_gmt_to_matrix in libplaidgpu (include/plaidgpu.h), which follow the same rules.

## R/gmt-utils.R:99-125: one set per line "name <TAB> source <TAB> gene <TAB> gene ..."; lines are cut at '#';
## empty strings and "NA" are dropped from the members; duplicated members are kept once (setdiff)
read.gmt <- function(gmt.file, dir = NULL, add.source = FALSE, nrows = -1) {
  f0 <- gmt.file
  if (substr(gmt.file, 1, 1) == "/") dir <- NULL
  if (!is.null(dir)) f0 <- file.path(sub("/$", "", dir), gmt.file)
  lines <- readLines(f0, n = nrows, warn = FALSE)
  lines <- sub("#.*$", "", lines)
  lines <- lines[nzchar(trimws(lines))]
  fields <- strsplit(lines, "\t", fixed = FALSE)
  nm <- vapply(fields, function(f) f[1], "")
  src <- vapply(fields, function(f) if (length(f) >= 2) f[2] else NA_character_, "")
  gset <- lapply(fields, function(f) {
    if (length(f) < 3) return(character(0))
    g <- unlist(strsplit(f[-(1:2)], "[ \t]"))
    unique(g[!is.na(g) & g != "" & g != "NA"])
  })
  names(gset) <- if (add.source) paste0(nm, " (", src, ")") else nm
  gset
}

## R/gmt-utils.R:139-144
write.gmt <- function(gmt, file, source = NA) {
  if (length(source) == 1 && is.na(source)) source <- names(gmt)
  genes <- vapply(gmt, paste, "", collapse = "\t")
  writeLines(paste(names(gmt), source, genes, sep = "\t"), con = file)
  invisible(NULL)
}

## R/gmt-utils.R:19-66: sets by decreasing size (stable), first of a duplicated name kept; genes = the background
## (default: all members by decreasing number of sets); rows of the result by decreasing membership count
gmt2mat <- function(gmt, max.genes = -1, ntop = -1, sparse = TRUE, bg = NULL, use.multicore = TRUE) {
  gmt <- gmt[order(-lengths(gmt))]
  gmt <- gmt[!duplicated(names(gmt))]
  if (ntop > 0) gmt <- lapply(gmt, utils::head, n = ntop)
  if (is.null(names(gmt))) names(gmt) <- paste0("gmt.", seq_along(gmt))
  if (is.null(bg)) bg <- names(sort(table(unlist(gmt)), decreasing = TRUE))
  if (max.genes < 0) max.genes <- length(bg)
  gg <- utils::head(bg, n = max.genes)
  row <- lapply(gmt, function(s) { r <- match(unique(s), gg); r[!is.na(r)] })
  D <- Matrix::sparseMatrix(i = unlist(row), j = rep.int(seq_along(row), lengths(row)), x = 1,
                            dims = c(length(gg), length(gmt)), dimnames = list(gg, names(gmt)))
  if (!sparse) D <- as.matrix(D)
  D[order(-Matrix::rowSums(D != 0, na.rm = TRUE)), , drop = FALSE]
}

## R/gmt-utils.R:80-85
mat2gmt <- function(mat) {
  mat <- methods::as(methods::as(mat, "CsparseMatrix"), "generalMatrix")
  nz <- methods::as(mat != 0, "TsparseMatrix")
  members <- split(rownames(mat)[nz@i + 1L], factor(nz@j + 1L, levels = seq_len(ncol(mat))))
  members <- members[lengths(members) > 0]
  names(members) <- colnames(mat)[as.integer(names(members))]
  members
}
