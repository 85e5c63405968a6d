/* Minimal EXECUTABLE stand-in for the part of R's C API that r-package/src/r_shim.c uses, plus a
 * driver that plays R's role around the two .Call routines (tests/test_r_shim_runtime.py):
 *
 *   shim_driver <in.bin> <out.bin>
 *
 * in.bin  : int64 header {mode (0 IDW, 1 Cressman), n_cells, n_st, n_dates, n_radii}, then doubles:
 *           cell_x, cell_y, st_x, st_y, obs (column-major n_dates x n_st), p | radii_m
 * out.bin : the doubles of the returned array.
 * exit 0 on success; exit 3 with the message on stderr when the shim raised an R error.
 * Test infrastructure only — nothing here ships. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Rinternals.h"
#include "R_ext/Rdynload.h"

struct SEXPREC {
    SEXPTYPE type;
    R_xlen_t length;
    int nrow, ncol;   /* ncol > 0: has a dim attribute of length 2 */
    void* data;
};

static struct SEXPREC nil_rec = {0, 0, 0, 0, NULL};
SEXP R_NilValue = &nil_rec;
static jmp_buf error_jmp;
static char error_msg[1024];
static int protect_depth = 0;

static SEXP new_vec(SEXPTYPE type, R_xlen_t n) {
    SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
    s->type = type;
    s->length = n;
    s->data = malloc((size_t)(n > 0 ? n : 1) * (type == REALSXP ? sizeof(double) : sizeof(int)));
    return s;
}
Rboolean Rf_isReal(SEXP s) { return s->type == REALSXP; }
Rboolean Rf_isInteger(SEXP s) { return s->type == INTSXP; }
Rboolean Rf_isNull(SEXP s) { return s == R_NilValue; }
Rboolean Rf_isMatrix(SEXP s) { return s->ncol > 0; }
int Rf_nrows(SEXP s) { return s->nrow; }
int Rf_ncols(SEXP s) { return s->ncol; }
int LENGTH(SEXP s) { return (int)s->length; }
R_xlen_t XLENGTH(SEXP s) { return s->length; }
double* REAL(SEXP s) { return (double*)s->data; }
int* INTEGER(SEXP s) { return (int*)s->data; }
double Rf_asReal(SEXP s) { return ((double*)s->data)[0]; }
SEXP Rf_allocVector(SEXPTYPE t, R_xlen_t n) { return new_vec(t, n); }
SEXP Rf_allocMatrix(SEXPTYPE t, int nr, int nc) {
    SEXP s = new_vec(t, (R_xlen_t)nr * nc);
    s->nrow = nr;
    s->ncol = nc;
    return s;
}
SEXP Rf_allocArray(SEXPTYPE t, SEXP dims) {
    R_xlen_t n = 1;
    for (int i = 0; i < LENGTH(dims); i++) n *= INTEGER(dims)[i];
    return new_vec(t, n);
}
SEXP Rf_protect(SEXP s) {
    protect_depth++;
    return s;
}
void Rf_unprotect(int n) { protect_depth -= n; }
void Rf_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(error_msg, sizeof error_msg, fmt, ap);
    va_end(ap);
    longjmp(error_jmp, 1);   /* what R does */
}
static const R_CallMethodDef* registered = NULL;
int R_registerRoutines(DllInfo* dll, const void* c, const R_CallMethodDef* call, const void* f, const void* e) {
    (void)dll; (void)c; (void)f; (void)e;
    registered = call;
    return 1;
}
Rboolean R_useDynamicSymbols(DllInfo* dll, Rboolean v) {
    (void)dll;
    return v;
}

void R_init_InterpolateR(DllInfo* dll);
typedef SEXP (*call7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

static SEXP read_real(FILE* f, R_xlen_t n) {
    SEXP s = new_vec(REALSXP, n);
    if (fread(s->data, sizeof(double), (size_t)n, f) != (size_t)n) {
        fprintf(stderr, "short input\n");
        exit(2);
    }
    return s;
}

int main(int argc, char** argv) {
    if (argc != 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int64_t h[5];
    if (fread(h, sizeof(int64_t), 5, f) != 5) return 2;
    const int mode = (int)h[0];
    const R_xlen_t n_cells = (R_xlen_t)h[1];
    const int n_st = (int)h[2], n_dates = (int)h[3], n_radii = (int)h[4];
    SEXP cell_x = read_real(f, n_cells), cell_y = read_real(f, n_cells);
    SEXP st_x = read_real(f, n_st), st_y = read_real(f, n_st);
    SEXP obs = read_real(f, (R_xlen_t)n_dates * n_st);
    obs->nrow = n_dates;
    obs->ncol = n_st;
    SEXP last = read_real(f, mode == 0 ? 1 : n_radii);
    fclose(f);

    R_init_InterpolateR(NULL);   /* library(InterpolateR): routine registration */
    const char* want = mode == 0 ? "C_ipr_idw" : "C_ipr_cressman";
    call7 fn = NULL;
    for (const R_CallMethodDef* m = registered; m && m->name; m++)
        if (strcmp(m->name, want) == 0 && m->numArgs == 7) fn = (call7)m->fun;
    if (!fn) {
        fprintf(stderr, "%s is not registered with 7 arguments\n", want);
        return 2;
    }
    if (setjmp(error_jmp)) {
        fprintf(stderr, "R error: %s\n", error_msg);
        return 3;
    }
    SEXP out = fn(cell_x, cell_y, st_x, st_y, obs, last, R_NilValue);   /* .Call(...) */
    if (protect_depth != 0) {
        fprintf(stderr, "protect stack imbalance: %d\n", protect_depth);
        return 4;
    }
    const R_xlen_t expect = n_cells * n_dates * (mode == 0 ? 1 : n_radii);
    if (!Rf_isReal(out) || XLENGTH(out) != expect) {
        fprintf(stderr, "unexpected result shape\n");
        return 4;
    }
    FILE* g = fopen(argv[2], "wb");
    if (!g) return 2;
    fwrite(REAL(out), sizeof(double), (size_t)expect, g);
    fclose(g);
    return 0;
}
