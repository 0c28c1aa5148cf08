/* .Call shim between R and libhaystack_b200.so (C ABI: include/haystack_b200.h).
 *
 * Neither the build container nor the GPU boxes have R (no Rinternals.h), so this file is compiled and RUN against a
 * miniature stand-in for R's C API (tests/rmock; tests/test_rshim_cpu.py, tests/test_gpu_rshim.py) -- it has not met
 * a real R yet.  It is the binding a maintainer adds to put the B200 path behind the unchanged haystack() API
 * (package skeleton in rshim/).  Rules it follows:
 *   - REAL()/INTEGER() pointers go straight through: R matrices are column-major doubles, dgRMatrix
 *     slots are @p int, @j int (0-based), @x double -- exactly what the C ABI takes.
 *   - outputs are allocated with allocVector under PROTECT;
 *   - the context lives in an external pointer with a finalizer;
 *   - a non-zero return code becomes Rf_error() AFTER the call returned (never from inside the library).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>

#include "haystack_b200.h"

static void ctx_finalizer(SEXP ptr) {
  hs_ctx* ctx = (hs_ctx*)R_ExternalPtrAddr(ptr);
  if (ctx) hs_destroy(ctx);
  R_ClearExternalPtr(ptr);
}

static hs_ctx* get_ctx(SEXP ptr) {
  hs_ctx* ctx = (hs_ctx*)R_ExternalPtrAddr(ptr);
  if (!ctx) Rf_error("haystack_b200: context has been destroyed");
  return ctx;
}

static void check(hs_ctx* ctx, int rc, const char* what) {
  if (rc != HS_OK) Rf_error("%s failed (%d): %s", what, rc, hs_last_error(ctx));
}

SEXP hsb_create(SEXP device) {
  hs_ctx* ctx = NULL;
  int rc = hs_create(asInteger(device), &ctx);
  if (rc != HS_OK) Rf_error("hs_create failed (%d): %s", rc, hs_last_error(NULL));
  SEXP ptr = PROTECT(R_MakeExternalPtr(ctx, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, ctx_finalizer, TRUE);
  UNPROTECT(1);
  return ptr;
}

/* x: scaled N x d matrix, grid: G x d matrix, w: NULL or numeric(N), want_dens: logical.
 * returns list(bandwidth, Q, densities | NULL)  -- R/haystack_continuous.R:168-186 */
SEXP hsb_density(SEXP sctx, SEXP x, SEXP grid, SEXP w, SEXP want_dens) {
  hs_ctx* ctx = get_ctx(sctx);
  const int64_t n = nrows(x);
  const int d = ncols(x), g = nrows(grid);
  SEXP q = PROTECT(allocVector(REALSXP, g));
  SEXP dens = R_NilValue;
  if (asLogical(want_dens)) dens = allocMatrix(REALSXP, (int)n, g);
  PROTECT(dens);
  double bw = 0.0;
  int rc = hs_density(ctx, REAL(x), n, d, REAL(grid), g, isNull(w) ? NULL : REAL(w), &bw, REAL(q),
                      isNull(dens) ? NULL : REAL(dens), NULL);
  check(ctx, rc, "hs_density");
  SEXP out = PROTECT(allocVector(VECSXP, 3));
  SET_VECTOR_ELT(out, 0, ScalarReal(bw));
  SET_VECTOR_ELT(out, 1, q);
  SET_VECTOR_ELT(out, 2, dens);
  UNPROTECT(3);
  return out;
}

/* expression: numeric matrix genes x cells -- loop R/haystack_continuous.R:202-206 */
SEXP hsb_dkl_dense(SEXP sctx, SEXP expression) {
  hs_ctx* ctx = get_ctx(sctx);
  const int64_t m = nrows(expression);
  SEXP out = PROTECT(allocVector(REALSXP, m));
  check(ctx, hs_dkl_dense(ctx, REAL(expression), m, m, REAL(out)), "hs_dkl_dense");
  UNPROTECT(1);
  return out;
}

/* dgRMatrix slots -- loop R/haystack_continuous.R:207-212 */
SEXP hsb_dkl_csr(SEXP sctx, SEXP p, SEXP j, SEXP xval) {
  hs_ctx* ctx = get_ctx(sctx);
  const int64_t m = XLENGTH(p) - 1;
  SEXP out = PROTECT(allocVector(REALSXP, m));
  check(ctx, hs_dkl_csr(ctx, INTEGER(p), 0, INTEGER(j), REAL(xval), m, REAL(out)), "hs_dkl_csr");
  UNPROTECT(1);
  return out;
}

/* v: numeric(N); perm: integer N x R matrix, column r = sample.int(N) -- R/haystack_continuous.R:250-265 */
SEXP hsb_dkl_perm_dense(SEXP sctx, SEXP v, SEXP perm) {
  hs_ctx* ctx = get_ctx(sctx);
  const int r = ncols(perm);
  SEXP out = PROTECT(allocVector(REALSXP, r));
  check(ctx, hs_dkl_perm_dense(ctx, REAL(v), INTEGER(perm), r, 1, REAL(out)), "hs_dkl_perm_dense");
  UNPROTECT(1);
  return out;
}

/* gene_ptr: numeric(S+1) prefix offsets (doubles: may exceed int), val: numeric, ind: integer vector
 * laid out gene by gene as nnz_s x R column-major blocks -- R/haystack_continuous.R:266-283 */
SEXP hsb_dkl_perm_csr(SEXP sctx, SEXP gene_ptr, SEXP val, SEXP ind, SEXP n_perm) {
  hs_ctx* ctx = get_ctx(sctx);
  const int64_t s = XLENGTH(gene_ptr) - 1;
  const int r = asInteger(n_perm);
  int64_t* gp = (int64_t*)R_alloc((size_t)s + 1, sizeof(int64_t));
  for (int64_t i = 0; i <= s; ++i) gp[i] = (int64_t)REAL(gene_ptr)[i];
  SEXP out = PROTECT(allocMatrix(REALSXP, (int)s, r));
  check(ctx, hs_dkl_perm_csr(ctx, gp, s, REAL(val), INTEGER(ind), r, 1, REAL(out)), "hs_dkl_perm_csr");
  UNPROTECT(1);
  return out;
}

/* Asynchronous pair of hsb_dkl_perm_csr: _begin starts the scoring of one batch of genes on the library's worker
 * thread and returns, so R can draw the next batch with sample() meanwhile; _wait returns the batch's
 * n.genes x randomization.count matrix.  val and ind are read until the wait: they are preserved here and released
 * there (the caller need not keep them alive). */
static SEXP pending_val = NULL, pending_ind = NULL;
static int64_t pending_rows = 0;
static int pending_perm = 0;

SEXP hsb_dkl_perm_csr_begin(SEXP sctx, SEXP gene_ptr, SEXP val, SEXP ind, SEXP n_perm) {
  hs_ctx* ctx = get_ctx(sctx);
  if (pending_val) Rf_error("hsb_dkl_perm_csr_begin: a batch is already in flight");
  const int64_t s = XLENGTH(gene_ptr) - 1;
  const int r = asInteger(n_perm);
  int64_t* gp = (int64_t*)R_alloc((size_t)s + 1, sizeof(int64_t));
  for (int64_t i = 0; i <= s; ++i) gp[i] = (int64_t)REAL(gene_ptr)[i];
  R_PreserveObject(val);
  R_PreserveObject(ind);
  const int rc = hs_dkl_perm_csr_begin(ctx, gp, s, REAL(val), INTEGER(ind), r, 1);
  if (rc != HS_OK) {
    R_ReleaseObject(val);
    R_ReleaseObject(ind);
    check(ctx, rc, "hs_dkl_perm_csr_begin");
  }
  pending_val = val;
  pending_ind = ind;
  pending_rows = s;
  pending_perm = r;
  return R_NilValue;
}

SEXP hsb_dkl_perm_csr_wait(SEXP sctx) {
  hs_ctx* ctx = get_ctx(sctx);
  if (!pending_val) Rf_error("hsb_dkl_perm_csr_wait: no batch in flight");
  SEXP out = PROTECT(allocMatrix(REALSXP, (int)pending_rows, pending_perm));
  const int rc = hs_dkl_perm_csr_wait(ctx, REAL(out));
  R_ReleaseObject(pending_val);
  R_ReleaseObject(pending_ind);
  pending_val = pending_ind = NULL;
  check(ctx, rc, "hs_dkl_perm_csr_wait");
  UNPROTECT(1);
  return out;
}

/* Seeded form: the draws of R/haystack_continuous.R:275 are made on the device.  seed: numeric(1) holding
 * a non-negative integer < 2^53 (e.g. sample.int(.Machine$integer.max, 1)). */
SEXP hsb_dkl_perm_csr_seeded(SEXP sctx, SEXP gene_ptr, SEXP val, SEXP n_perm, SEXP seed) {
  hs_ctx* ctx = get_ctx(sctx);
  const int64_t s = XLENGTH(gene_ptr) - 1;
  const int r = asInteger(n_perm);
  int64_t* gp = (int64_t*)R_alloc((size_t)s + 1, sizeof(int64_t));
  for (int64_t i = 0; i <= s; ++i) gp[i] = (int64_t)REAL(gene_ptr)[i];
  SEXP out = PROTECT(allocMatrix(REALSXP, (int)s, r));
  check(ctx, hs_dkl_perm_csr_seeded(ctx, gp, s, REAL(val), r, (uint64_t)asReal(seed), 0, REAL(out)),
        "hs_dkl_perm_csr_seeded");
  UNPROTECT(1);
  return out;
}

/* Dense seeded form: R/haystack_continuous.R:259 drawn on the device; stream_base = (i - 1) * R for gene i */
SEXP hsb_dkl_perm_dense_seeded(SEXP sctx, SEXP v, SEXP n_perm, SEXP seed, SEXP stream_base) {
  hs_ctx* ctx = get_ctx(sctx);
  const int r = asInteger(n_perm);
  SEXP out = PROTECT(allocVector(REALSXP, r));
  check(ctx, hs_dkl_perm_dense_seeded(ctx, REAL(v), r, (uint64_t)asReal(seed), (uint64_t)asReal(stream_base), REAL(out)),
        "hs_dkl_perm_dense_seeded");
  UNPROTECT(1);
  return out;
}

/* The whole dense loop R/haystack_continuous.R:250-265 in one call: rows sel (1-based, order of genes.to.randomize) of
 * the genes x cells matrix; returns the n.genes.to.randomize x randomization.count matrix all.D_KL.randomized. */
SEXP hsb_dkl_perm_dense_seeded_batch(SEXP sctx, SEXP expression, SEXP sel, SEXP n_perm, SEXP seed) {
  hs_ctx* ctx = get_ctx(sctx);
  const int64_t ld = (int64_t)nrows(expression), s = XLENGTH(sel);
  const int r = asInteger(n_perm);
  int64_t* rows = (int64_t*)R_alloc((size_t)s + 1, sizeof(int64_t));
  for (int64_t i = 0; i < s; ++i) rows[i] = (int64_t)INTEGER(sel)[i] - 1;
  SEXP out = PROTECT(allocMatrix(REALSXP, (int)s, r));
  check(ctx, hs_dkl_perm_dense_seeded_batch(ctx, REAL(expression), ld, rows, s, r, (uint64_t)asReal(seed), 0, REAL(out)),
        "hs_dkl_perm_dense_seeded_batch");
  UNPROTECT(1);
  return out;
}

/* Scores of all genes and the seeded randomization of rows sel (1-based, order of genes.to.randomize) in one
 * call; returns list(D_KL, all.D_KL.randomized). */
SEXP hsb_dkl_csr_randomized(SEXP sctx, SEXP p, SEXP j, SEXP x, SEXP sel, SEXP n_perm, SEXP seed) {
  hs_ctx* ctx = get_ctx(sctx);
  const int64_t m = XLENGTH(p) - 1, s = XLENGTH(sel);
  const int r = asInteger(n_perm);
  int64_t* rows = (int64_t*)R_alloc((size_t)s + 1, sizeof(int64_t));
  for (int64_t i = 0; i < s; ++i) rows[i] = (int64_t)INTEGER(sel)[i] - 1;
  SEXP dkl = PROTECT(allocVector(REALSXP, (R_xlen_t)m));
  SEXP rnd = PROTECT(allocMatrix(REALSXP, (int)s, r));
  check(ctx, hs_dkl_csr_randomized(ctx, INTEGER(p), 0, INTEGER(j), REAL(x), m, rows, s, r, (uint64_t)asReal(seed), 0,
                                   REAL(dkl), REAL(rnd)), "hs_dkl_csr_randomized");
  SEXP out = PROTECT(allocVector(VECSXP, 2));
  SET_VECTOR_ELT(out, 0, dkl);
  SET_VECTOR_ELT(out, 1, rnd);
  UNPROTECT(3);
  return out;
}

/* ids of one draw (1-based), for reproducing a seeded randomization in R */
SEXP hsb_perm_sample(SEXP seed, SEXP stream, SEXP n, SEXP size) {
  const int64_t k = (int64_t)asReal(size);
  SEXP out = PROTECT(allocVector(INTSXP, (R_xlen_t)k));
  if (hs_perm_sample((uint64_t)asReal(seed), (uint64_t)asReal(stream), (int64_t)asReal(n), k, 1, INTEGER(out)) != HS_OK) {
    UNPROTECT(1);
    Rf_error("hs_perm_sample: %s", hs_last_error(NULL));
  }
  UNPROTECT(1);
  return out;
}

static const R_CallMethodDef call_methods[] = {
    {"hsb_create", (DL_FUNC)&hsb_create, 1},
    {"hsb_density", (DL_FUNC)&hsb_density, 5},
    {"hsb_dkl_dense", (DL_FUNC)&hsb_dkl_dense, 2},
    {"hsb_dkl_csr", (DL_FUNC)&hsb_dkl_csr, 4},
    {"hsb_dkl_perm_dense", (DL_FUNC)&hsb_dkl_perm_dense, 3},
    {"hsb_dkl_perm_csr", (DL_FUNC)&hsb_dkl_perm_csr, 5},
    {"hsb_dkl_perm_csr_begin", (DL_FUNC)&hsb_dkl_perm_csr_begin, 5},
    {"hsb_dkl_perm_csr_wait", (DL_FUNC)&hsb_dkl_perm_csr_wait, 1},
    {"hsb_dkl_perm_csr_seeded", (DL_FUNC)&hsb_dkl_perm_csr_seeded, 5},
    {"hsb_dkl_perm_dense_seeded", (DL_FUNC)&hsb_dkl_perm_dense_seeded, 5},
    {"hsb_dkl_perm_dense_seeded_batch", (DL_FUNC)&hsb_dkl_perm_dense_seeded_batch, 5},
    {"hsb_dkl_csr_randomized", (DL_FUNC)&hsb_dkl_csr_randomized, 7},
    {"hsb_perm_sample", (DL_FUNC)&hsb_perm_sample, 4},
    {NULL, NULL, 0}};

void R_init_haystackB200(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
