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
#define CLOSXP 3
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define LANGSXP 6
extern SEXP R_NilValue;
extern SEXP R_GlobalEnv;
extern SEXP R_DimSymbol;
struct R_allocator;
Rboolean Rf_isFunction(SEXP);
int Rf_asLogical(SEXP);
SEXP Rf_allocVector3(SEXPTYPE, R_xlen_t, struct R_allocator*);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_ScalarReal(double);
SEXP Rf_lang3(SEXP, SEXP, SEXP);
SEXP R_tryEval(SEXP, SEXP, int*);
Rboolean R_ToplevelExec(void (*fun)(void*), void* data);
void R_CheckUserInterrupt(void);
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
#define isFunction Rf_isFunction
#define asLogical Rf_asLogical
#define allocVector3 Rf_allocVector3
#define setAttrib Rf_setAttrib
#define ScalarReal Rf_ScalarReal
#define lang3 Rf_lang3
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
#define error Rf_error
#endif
