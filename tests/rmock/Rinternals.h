/* TEST INFRASTRUCTURE: a miniature stand-in for R's C API, just enough of it to compile, link and RUN
 * rshim/src/haystack_b200_shim.c where no R exists (this image, the GPU boxes).  SEXPs are heap records that are
 * never freed (tests are short-lived); PROTECT is a no-op; Rf_error records the message and longjmps to the
 * test harness (rmock_call in rmock.c) -- like R, it does not return.  Not a general R emulation. */
#ifndef RMOCK_RINTERNALS_H
#define RMOCK_RINTERNALS_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
typedef enum { NILSXP = 0, LGLSXP = 10, INTSXP = 13, REALSXP = 14, VECSXP = 19, EXTPTRSXP = 22 } SEXPTYPE;
typedef ptrdiff_t R_xlen_t;
typedef struct rmock_sexp {
  SEXPTYPE type;
  R_xlen_t length;
  int nrow, ncol;      /* 0, 0 for plain vectors */
  void* data;          /* int[] / double[] / SEXP[] / external address */
  void (*finalizer)(struct rmock_sexp*);
}* SEXP;
typedef enum { FALSE = 0, TRUE } Rboolean;
typedef void (*R_CFinalizer_t)(SEXP);

extern SEXP R_NilValue;
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_ScalarReal(double);
int Rf_asInteger(SEXP);
int Rf_asLogical(SEXP);
double Rf_asReal(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
void Rf_error(const char*, ...) __attribute__((noreturn));
char* R_alloc(size_t, int);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
void R_PreserveObject(SEXP);
void R_ReleaseObject(SEXP);

#define PROTECT(x) (x)
#define UNPROTECT(n) ((void)(n))
#define REAL(x) ((double*)(x)->data)
#define INTEGER(x) ((int*)(x)->data)
#define XLENGTH(x) ((x)->length)
#define SET_VECTOR_ELT(v, i, x) (((SEXP*)(v)->data)[i] = (x))
#define VECTOR_ELT(v, i) (((SEXP*)(v)->data)[i])
#define isNull(x) ((x) == R_NilValue || (x)->type == NILSXP)
#define allocVector Rf_allocVector
#define allocMatrix Rf_allocMatrix
#define ScalarReal Rf_ScalarReal
#define asInteger Rf_asInteger
#define asLogical Rf_asLogical
#define asReal Rf_asReal
#define nrows Rf_nrows
#define ncols Rf_ncols
#ifdef __cplusplus
}
#endif
#endif
