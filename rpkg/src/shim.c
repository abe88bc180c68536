/*
 * .Call shim between R and libplaidgpu (include/plaidgpu.h).  Thin by design: it only unpacks
 * SEXPs into plain pointers, allocates the result, and maps status codes to R conditions.
 * No arithmetic happens here.  R's API is single-threaded: SEXPs are touched on the calling
 * thread only; the library joins its streams before it returns.
 *
 * Replaces nothing in the reference (it has no src/); it is the binding a maintainer adds to
 * route R/plaid.R:60-87,100-123,155-309,554-650 to the GPU.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "plaidgpu.h"

static void ctx_finalizer(SEXP ptr) {
  plaidgpu_ctx* c = (plaidgpu_ctx*)R_ExternalPtrAddr(ptr);
  if (c) {
    plaidgpu_destroy(c);
    R_ClearExternalPtr(ptr);
  }
}

static plaidgpu_ctx* get_ctx(SEXP s) {
  plaidgpu_ctx* c = (plaidgpu_ctx*)R_ExternalPtrAddr(s);
  if (!c) Rf_error("plaidgpu: context has been released");
  return c;
}

SEXP C_plaidgpu_ctx(SEXP device) {
  plaidgpu_ctx* c = NULL;
  int rc = plaidgpu_init(Rf_asInteger(device), &c);
  if (rc != PLAIDGPU_OK) Rf_error("plaidgpu_init failed (%d): no usable B200; this package has no CPU fallback", rc);
  SEXP p = PROTECT(R_MakeExternalPtr(c, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(p, ctx_finalizer, TRUE);
  UNPROTECT(1);
  return p;
}

static SEXP list_get(SEXP lst, const char* name) {
  SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
  for (R_xlen_t k = 0; k < XLENGTH(lst); ++k)
    if (strcmp(CHAR(STRING_ELT(names, k)), name) == 0) return VECTOR_ELT(lst, k);
  return R_NilValue;
}
static int opt_int(SEXP lst, const char* n, int def) { SEXP v = list_get(lst, n); return v == R_NilValue ? def : Rf_asInteger(v); }
static double opt_dbl(SEXP lst, const char* n, double def) { SEXP v = list_get(lst, n); return v == R_NilValue ? def : Rf_asReal(v); }

static void fill_matrix(plaidgpu_matrix* M, int kind, SEXP p, SEXP i, SEXP x, SEXP dim) {
  memset(M, 0, sizeof(*M));
  M->kind = kind;
  M->location = PLAIDGPU_HOST;
  M->P = INTEGER(dim)[0];
  M->N = INTEGER(dim)[1];
  M->p = kind == PLAIDGPU_CSC ? INTEGER(p) : NULL;
  M->i = kind == PLAIDGPU_CSC ? INTEGER(i) : NULL;
  M->x = REAL(x);
}

/* named list of options from R -> plaidgpu_opts (defaults for whatever is absent) */
static void fill_opts(plaidgpu_opts* op, SEXP opts) {
  plaidgpu_opts o;
  plaidgpu_default_opts(&o);
  o.scorer = opt_int(opts, "scorer", o.scorer);
  o.stats_mean = opt_int(opts, "stats_mean", o.stats_mean);
  o.normalize = opt_int(opts, "normalize", o.normalize);
  o.remove_log2 = opt_int(opts, "remove_log2", o.remove_log2);
  o.score_mean = opt_int(opts, "score_mean", o.score_mean);
  o.alpha = opt_dbl(opts, "alpha", o.alpha);
  o.rmax = opt_dbl(opts, "rmax", o.rmax);
  o.auc_max_rank = opt_dbl(opts, "auc_max_rank", o.auc_max_rank);
  o.tau = opt_dbl(opts, "tau", o.tau);
  o.gsva_ecdf = opt_int(opts, "gsva_ecdf", 0);
  o.nrow_x = (int64_t)opt_dbl(opts, "nrow_x", 0.0);
  SEXP csums = list_get(opts, "matg_full_colsums");
  o.matg_full_colsums = csums == R_NilValue ? NULL : REAL(csums);
  *op = o;
}

/* plaid() and the replaid.* scorers.  `ctx` is one context (external pointer) or a list of contexts on different
 * devices (options(plaid.gpus = n) in R/plaid.R): the columns are then split over them by plaidgpu_score_multi, one
 * host thread per device inside the library, the shards landing directly in the result matrix. */
SEXP C_plaidgpu_score(SEXP ctx, SEXP kind, SEXP Xp, SEXP Xi, SEXP Xx, SEXP Xdim, SEXP Gp, SEXP Gi, SEXP Gx,
                      SEXP Gdim, SEXP rowmap, SEXP opts) {
  plaidgpu_ctx* cs[16];
  int nctx = 1;
  if (TYPEOF(ctx) == VECSXP) {
    nctx = (int)XLENGTH(ctx);
    if (nctx < 1 || nctx > 16) Rf_error("plaidgpu: between 1 and 16 contexts expected");
    for (int k = 0; k < nctx; ++k) cs[k] = get_ctx(VECTOR_ELT(ctx, k));
  } else {
    cs[0] = get_ctx(ctx);
  }
  plaidgpu_ctx* c = cs[0];
  const int PG = INTEGER(Gdim)[0], S = INTEGER(Gdim)[1];
  /* re-registering the same pattern is a memcmp inside the library: the plan is kept across calls */
  for (int k = 0; k < nctx; ++k) {
    int rc0 = plaidgpu_set_genesets(cs[k], PG, S, INTEGER(Gp), INTEGER(Gi), REAL(Gx));
    if (rc0 != PLAIDGPU_OK) Rf_error("plaidgpu_set_genesets: %s", plaidgpu_last_error(cs[k]));
  }
  int rc;
  plaidgpu_matrix M;
  fill_matrix(&M, Rf_asInteger(kind), Xp, Xi, Xx, Xdim);
  plaidgpu_opts o;
  fill_opts(&o, opts);
  o.out_location = PLAIDGPU_HOST;
  R_CheckUserInterrupt(); /* before the (blocking) device call; the library itself never touches the R API */
  /* S x N may exceed 2^31 elements: Rf_allocMatrix takes ints for the dims, the product is R_xlen_t */
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, S, (int)M.N));
  if (nctx > 1 && !(o.scorer == PLAIDGPU_GSVA && o.gsva_ecdf == PLAIDGPU_ROWTF_ECDF))
    rc = plaidgpu_score_multi(cs, nctx, &M, INTEGER(rowmap), &o, REAL(out));
  else
    rc = plaidgpu_score(c, &M, INTEGER(rowmap), &o, REAL(out));
  if (rc != PLAIDGPU_OK) {
    UNPROTECT(1);
    Rf_error("plaidgpu_score: %s", plaidgpu_last_error(c)); /* longjmp: nothing left to release */
  }
  UNPROTECT(1);
  return out;
}

/* chunked_crossprod(x, y): t(x) %*% y for a column-scaled binary sparse x */
SEXP C_plaidgpu_crossprod(SEXP ctx, SEXP kind, SEXP Yp, SEXP Yi, SEXP Yx, SEXP Ydim, SEXP Gp, SEXP Gi, SEXP Gx,
                          SEXP Gdim, SEXP colscale) {
  plaidgpu_ctx* c = get_ctx(ctx);
  const int PG = INTEGER(Gdim)[0], S = INTEGER(Gdim)[1];
  int rc = plaidgpu_set_genesets(c, PG, S, INTEGER(Gp), INTEGER(Gi), REAL(Gx));
  if (rc != PLAIDGPU_OK) Rf_error("plaidgpu_set_genesets: %s", plaidgpu_last_error(c));
  plaidgpu_matrix M;
  fill_matrix(&M, Rf_asInteger(kind), Yp, Yi, Yx, Ydim);
  if (M.P != PG) Rf_error("non-conformable arguments");
  int* rowmap = (int*)R_alloc((size_t)PG, sizeof(int)); /* identity: same rows on both sides */
  for (int r = 0; r < PG; ++r) rowmap[r] = r;
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, S, (int)M.N));
  rc = plaidgpu_crossprod(c, &M, rowmap, REAL(colscale), PLAIDGPU_HOST, REAL(out));
  if (rc != PLAIDGPU_OK) {
    UNPROTECT(1);
    Rf_error("plaidgpu_crossprod: %s", plaidgpu_last_error(c));
  }
  UNPROTECT(1);
  return out;
}

/* normalize_medians(x, ignore.zero) */
SEXP C_plaidgpu_normalize_medians(SEXP ctx, SEXP x, SEXP ignore_zero) {
  plaidgpu_ctx* c = get_ctx(ctx);
  SEXP dim = Rf_getAttrib(x, R_DimSymbol);
  const int S = INTEGER(dim)[0], N = INTEGER(dim)[1];
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, S, N));
  int rc = plaidgpu_normalize_medians(c, REAL(x), S, N, Rf_asInteger(ignore_zero), PLAIDGPU_HOST, REAL(out));
  if (rc != PLAIDGPU_OK) {
    UNPROTECT(1);
    Rf_error("plaidgpu_normalize_medians: %s", plaidgpu_last_error(c));
  }
  UNPROTECT(1);
  return out;
}

/* colranks() / sparse_colranks(): returns nnz ranks (keep_zero) or P*N ranks */
SEXP C_plaidgpu_colranks(SEXP ctx, SEXP kind, SEXP Xp, SEXP Xi, SEXP Xx, SEXP Xdim, SEXP ties, SEXP is_signed,
                         SEXP keep_zero) {
  plaidgpu_ctx* c = get_ctx(ctx);
  plaidgpu_matrix M;
  fill_matrix(&M, Rf_asInteger(kind), Xp, Xi, Xx, Xdim);
  const int kz = Rf_asInteger(keep_zero);
  const R_xlen_t n = (M.kind == PLAIDGPU_CSC && kz) ? XLENGTH(Xx) : (R_xlen_t)M.P * (R_xlen_t)M.N;
  SEXP out = PROTECT(Rf_allocVector(REALSXP, n));
  int rc = plaidgpu_colranks(c, &M, Rf_asInteger(ties), Rf_asInteger(is_signed), kz, PLAIDGPU_HOST, REAL(out));
  if (rc != PLAIDGPU_OK) {
    UNPROTECT(1);
    Rf_error("plaidgpu_colranks: %s", plaidgpu_last_error(c));
  }
  UNPROTECT(1);
  return out;
}

/* per-set sums / sums of squares of gsetX by group y (plaid.test "lm"): 4 x S matrix */
SEXP C_plaidgpu_group_moments(SEXP ctx, SEXP x, SEXP y) {
  plaidgpu_ctx* c = get_ctx(ctx);
  SEXP dim = Rf_getAttrib(x, R_DimSymbol);
  const int S = INTEGER(dim)[0], N = INTEGER(dim)[1];
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, S, 4)); /* columns: sum0, sumsq0, sum1, sumsq1 */
  int rc = plaidgpu_group_moments(c, REAL(x), S, N, INTEGER(y), PLAIDGPU_HOST, REAL(out));
  if (rc != PLAIDGPU_OK) {
    UNPROTECT(1);
    Rf_error("plaidgpu_group_moments: %s", plaidgpu_last_error(c));
  }
  UNPROTECT(1);
  return out;
}

/* plaid.test(tests = "lm") without gsetX (R/plaid.R:423-431): plaid() fused with the per-set group reductions —
 * the S x N scores stay on the device, an S x 4 matrix (sum0, sumsq0, sum1, sumsq1) comes back */
SEXP C_plaidgpu_score_group_moments(SEXP ctx, SEXP kind, SEXP Xp, SEXP Xi, SEXP Xx, SEXP Xdim, SEXP Gp, SEXP Gi, SEXP Gx,
                                    SEXP Gdim, SEXP rowmap, SEXP opts, SEXP y) {
  plaidgpu_ctx* c = get_ctx(TYPEOF(ctx) == VECSXP ? VECTOR_ELT(ctx, 0) : ctx);
  const int PG = INTEGER(Gdim)[0], S = INTEGER(Gdim)[1];
  if (plaidgpu_set_genesets(c, PG, S, INTEGER(Gp), INTEGER(Gi), REAL(Gx)) != PLAIDGPU_OK)
    Rf_error("plaidgpu_set_genesets: %s", plaidgpu_last_error(c));
  plaidgpu_matrix M;
  fill_matrix(&M, Rf_asInteger(kind), Xp, Xi, Xx, Xdim);
  if ((int64_t)XLENGTH(y) != M.N) Rf_error("plaidgpu: length(y) must equal ncol(X)");
  plaidgpu_opts o;
  fill_opts(&o, opts);
  R_CheckUserInterrupt();
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, S, 4));
  int rc = plaidgpu_score_group_moments(c, &M, INTEGER(rowmap), &o, INTEGER(y), REAL(out));
  if (rc != PLAIDGPU_OK) {
    UNPROTECT(1);
    Rf_error("plaidgpu_score_group_moments: %s", plaidgpu_last_error(c));
  }
  UNPROTECT(1);
  return out;
}

static const R_CallMethodDef call_methods[] = {
    {"C_plaidgpu_ctx", (DL_FUNC)&C_plaidgpu_ctx, 1},
    {"C_plaidgpu_score", (DL_FUNC)&C_plaidgpu_score, 12},
    {"C_plaidgpu_normalize_medians", (DL_FUNC)&C_plaidgpu_normalize_medians, 3},
    {"C_plaidgpu_colranks", (DL_FUNC)&C_plaidgpu_colranks, 9},
    {"C_plaidgpu_group_moments", (DL_FUNC)&C_plaidgpu_group_moments, 3},
    {"C_plaidgpu_score_group_moments", (DL_FUNC)&C_plaidgpu_score_group_moments, 13},
    {"C_plaidgpu_crossprod", (DL_FUNC)&C_plaidgpu_crossprod, 11},
    {NULL, NULL, 0}};

void R_init_plaid(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
