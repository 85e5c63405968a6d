/* Minimal stand-in for R's C API, ONLY so that r-package/src/r_shim.c can be syntax- and
 * type-checked where R is not installed (tests/test_r_shim_syntax.py).  Declarations follow
 * R >= 4.4's Rinternals.h.  mock_runtime.c implements them for tests/test_r_shim_runtime.py. */
#ifndef MOCK_RINTERNALS_H
#define MOCK_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef unsigned int SEXPTYPE;
typedef enum { FALSE = 0, TRUE } Rboolean;
#define INTSXP 13
#define REALSXP 14
extern SEXP R_NilValue;
Rboolean Rf_isReal(SEXP);
Rboolean Rf_isInteger(SEXP);
Rboolean Rf_isNull(SEXP);
Rboolean Rf_isMatrix(SEXP);
int Rf_nrows(SEXP);
int Rf_ncols(SEXP);
int LENGTH(SEXP);
R_xlen_t XLENGTH(SEXP);
double* REAL(SEXP);
int* INTEGER(SEXP);
double Rf_asReal(SEXP);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocArray(SEXPTYPE, SEXP);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
void Rf_error(const char*, ...) __attribute__((noreturn));
#define isReal Rf_isReal
#define isInteger Rf_isInteger
#define isNull Rf_isNull
#define isMatrix Rf_isMatrix
#define nrows Rf_nrows
#define ncols Rf_ncols
#define asReal Rf_asReal
#define allocMatrix Rf_allocMatrix
#define allocVector Rf_allocVector
#define allocArray Rf_allocArray
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
#define error Rf_error
#endif
