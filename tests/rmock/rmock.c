/* TEST INFRASTRUCTURE: the few R API functions the shim calls, plus helpers the Python tests use to build arguments
 * and to call a registered routine with R's error semantics (Rf_error -> longjmp back, message kept). */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Rinternals.h"
#include "R_ext/Rdynload.h"

static struct rmock_sexp nil_rec = {NILSXP, 0, 0, 0, NULL, NULL};
SEXP R_NilValue = &nil_rec;
static jmp_buf* err_jmp = NULL;
static char err_msg[1024];

SEXP Rf_allocVector(SEXPTYPE t, R_xlen_t n) {
  SEXP s = (SEXP)calloc(1, sizeof(*s));
  const size_t el = t == REALSXP ? sizeof(double) : t == VECSXP ? sizeof(SEXP) : sizeof(int);
  s->type = t;
  s->length = n;
  s->data = calloc((size_t)(n > 0 ? n : 1), el);
  if (t == VECSXP) for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
  return s;
}
SEXP Rf_allocMatrix(SEXPTYPE t, int nr, int nc) {
  SEXP s = Rf_allocVector(t, (R_xlen_t)nr * nc);
  s->nrow = nr;
  s->ncol = nc;
  return s;
}
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }
int Rf_asInteger(SEXP s) { return s->type == REALSXP ? (int)REAL(s)[0] : INTEGER(s)[0]; }
int Rf_asLogical(SEXP s) { return Rf_asInteger(s) != 0; }
double Rf_asReal(SEXP s) { return s->type == REALSXP ? REAL(s)[0] : (double)INTEGER(s)[0]; }
int Rf_nrows(SEXP s) { return s->nrow ? s->nrow : (int)s->length; }
int Rf_ncols(SEXP s) { return s->ncol ? s->ncol : 1; }
char* R_alloc(size_t n, int size) { return (char*)calloc(n ? n : 1, (size_t)size); }
void Rf_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_msg, sizeof(err_msg), fmt, ap);
  va_end(ap);
  if (err_jmp) longjmp(*err_jmp, 1);
  fprintf(stderr, "Rf_error outside rmock_call: %s\n", err_msg);
  abort();
}
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) {
  (void)tag; (void)prot;
  SEXP s = (SEXP)calloc(1, sizeof(*s));
  s->type = EXTPTRSXP;
  s->data = p;
  return s;
}
void* R_ExternalPtrAddr(SEXP s) { return s->data; }
void R_ClearExternalPtr(SEXP s) { s->data = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t f, Rboolean onexit) { (void)onexit; s->finalizer = f; }
int rmock_preserved = 0;   /* objects held by R_PreserveObject and not yet released */
void R_PreserveObject(SEXP s) { (void)s; ++rmock_preserved; }
void R_ReleaseObject(SEXP s) { (void)s; --rmock_preserved; }

static DllInfo the_dll;
int R_registerRoutines(DllInfo* d, const void* c, const R_CallMethodDef* call, const void* f, const void* e) {
  (void)c; (void)f; (void)e;
  d->call_methods = call;
  return 1;
}
int R_useDynamicSymbols(DllInfo* d, int v) { d->dynamic_symbols = v; return 1; }

/* ---- helpers for the Python harness ---- */
void R_init_haystackB200(DllInfo*);
const R_CallMethodDef* rmock_load(void) { R_init_haystackB200(&the_dll); return the_dll.call_methods; }
const char* rmock_last_error(void) { return err_msg; }
void rmock_finalize(SEXP s) { if (s && s->finalizer) s->finalizer(s); }
SEXP rmock_real(const double* v, R_xlen_t n, int nr, int nc) {
  SEXP s = Rf_allocVector(REALSXP, n);
  memcpy(s->data, v, sizeof(double) * (size_t)n);
  s->nrow = nr; s->ncol = nc;
  return s;
}
SEXP rmock_int(const int* v, R_xlen_t n, int nr, int nc) {
  SEXP s = Rf_allocVector(INTSXP, n);
  memcpy(s->data, v, sizeof(int) * (size_t)n);
  s->nrow = nr; s->ncol = nc;
  return s;
}
SEXP rmock_lgl(int v) { SEXP s = Rf_allocVector(LGLSXP, 1); INTEGER(s)[0] = v; return s; }
/* .Call(name, args...): NULL when the routine raised an R error (message in rmock_last_error) */
SEXP rmock_call(const char* name, int nargs, SEXP* a) {
  typedef SEXP (*F1)(SEXP); typedef SEXP (*F2)(SEXP, SEXP); typedef SEXP (*F3)(SEXP, SEXP, SEXP);
  typedef SEXP (*F4)(SEXP, SEXP, SEXP, SEXP); typedef SEXP (*F5)(SEXP, SEXP, SEXP, SEXP, SEXP);
  typedef SEXP (*F7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
  const R_CallMethodDef* m = the_dll.call_methods;
  for (; m && m->name; ++m) if (!strcmp(m->name, name)) break;
  if (!m || !m->name) { snprintf(err_msg, sizeof(err_msg), "no registered routine %s", name); return NULL; }
  if (m->numArgs != nargs) { snprintf(err_msg, sizeof(err_msg), "%s takes %d arguments, got %d", name, m->numArgs, nargs); return NULL; }
  jmp_buf jb;
  err_jmp = &jb;
  SEXP out = NULL;
  if (!setjmp(jb)) {
    switch (nargs) {
      case 1: out = ((F1)m->fun)(a[0]); break;
      case 2: out = ((F2)m->fun)(a[0], a[1]); break;
      case 3: out = ((F3)m->fun)(a[0], a[1], a[2]); break;
      case 4: out = ((F4)m->fun)(a[0], a[1], a[2], a[3]); break;
      case 5: out = ((F5)m->fun)(a[0], a[1], a[2], a[3], a[4]); break;
      case 7: out = ((F7)m->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
      default: snprintf(err_msg, sizeof(err_msg), "rmock_call: %d arguments not supported", nargs);
    }
  }
  err_jmp = NULL;
  return out;
}
