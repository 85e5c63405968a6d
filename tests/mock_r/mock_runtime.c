/* Minimal EXECUTABLE stand-in for the part of R's C API that r-package/src/r_shim.c uses, plus a
 * driver that plays R's role around the .Call routines (tests/test_r_shim_runtime.py):
 *
 *   shim_driver <in.bin> <out.bin>
 *
 * in.bin  : int64 header {mode (0 IDW, 1 Cressman, 2 Kriging IDW fallback, 3 ordinary kriging), n_cells, n_st,
 *           n_dates, n_radii}, then doubles: cell_x, cell_y, st_x, st_y, obs (column-major n_dates x n_st),
 *           p | radii_m | (nothing) | model codes, nugget, sill, range (n_dates each)
 * out.bin : the doubles of the returned array.
 * environment: MOCK_PINNED=1            pinned = TRUE (result through the R_allocator_t of the shim)
 *              MOCK_PROGRESS=1          pass a progress closure; its calls are logged on stderr
 *              MOCK_PROGRESS_FAIL_AT=n  the closure raises an R error at its n-th call
 *              MOCK_INTERRUPT_AT=n      the n-th R_CheckUserInterrupt() finds a pending interrupt
 * exit 0 on success; exit 3 with the message on stderr when the shim raised an R error.
 * Test infrastructure only — nothing here ships. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "Rinternals.h"
#include "R_ext/Rallocators.h"
#include "R_ext/Rdynload.h"

struct SEXPREC {
    SEXPTYPE type;
    R_xlen_t length;
    int nrow, ncol;   /* ncol > 0: has a dim attribute of length 2 */
    int ndim;
    void* data;
    struct R_allocator alloc;   /* allocVector3: how the block is returned */
    void* block;
    SEXP car[3];                /* LANGSXP */
};

static struct SEXPREC nil_rec, env_rec, dim_rec;
SEXP R_NilValue = &nil_rec;
SEXP R_GlobalEnv = &env_rec;
SEXP R_DimSymbol = &dim_rec;
static jmp_buf error_jmp;
static jmp_buf* toplevel_jmp = NULL;   /* innermost R_ToplevelExec / R_tryEval context */
static char error_msg[1024];
static int protect_depth = 0;
static int interrupt_checks = 0, progress_calls = 0;
static int pinned_allocs = 0, pinned_frees = 0;

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

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
Rboolean Rf_isMatrix(SEXP s) { return s->ncol > 0 && s->ndim == 2; }
Rboolean Rf_isFunction(SEXP s) { return s->type == CLOSXP; }
int Rf_nrows(SEXP s) { return s->nrow; }
int Rf_ncols(SEXP s) { return s->ncol; }
int LENGTH(SEXP s) { return (int)s->length; }
R_xlen_t XLENGTH(SEXP s) { return s->length; }
double* REAL(SEXP s) { return (double*)s->data; }
int* INTEGER(SEXP s) { return (int*)s->data; }
double Rf_asReal(SEXP s) { return ((double*)s->data)[0]; }
int Rf_asLogical(SEXP s) { return s->type == LGLSXP ? ((int*)s->data)[0] : 0; }
SEXP Rf_allocVector(SEXPTYPE t, R_xlen_t n) { return new_vec(t, n); }
/* like R: ONE block from the allocator holds the header and the data; the data start inside it */
SEXP Rf_allocVector3(SEXPTYPE t, R_xlen_t n, struct R_allocator* al) {
    if (!al) return new_vec(t, n);
    const size_t header = 48;
    SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
    s->type = t;
    s->length = n;
    s->alloc = *al;
    s->block = al->mem_alloc(&s->alloc, header + (size_t)n * sizeof(double));
    if (!s->block) Rf_error("cannot allocate vector of size %.1f Gb", (double)n * 8 / 1073741824.0);
    pinned_allocs++;
    s->data = (char*)s->block + header;
    return s;
}
SEXP Rf_allocMatrix(SEXPTYPE t, int nr, int nc) {
    SEXP s = new_vec(t, (R_xlen_t)nr * nc);
    s->nrow = nr;
    s->ncol = nc;
    s->ndim = 2;
    return s;
}
SEXP Rf_allocArray(SEXPTYPE t, SEXP dims) {
    R_xlen_t n = 1;
    for (int i = 0; i < LENGTH(dims); i++) n *= INTEGER(dims)[i];
    return new_vec(t, n);
}
SEXP Rf_setAttrib(SEXP x, SEXP name, SEXP val) {
    if (name == R_DimSymbol) {
        R_xlen_t n = 1;
        for (int i = 0; i < LENGTH(val); i++) n *= INTEGER(val)[i];
        if (n != x->length) Rf_error("dims do not match the length of object");
        x->ndim = LENGTH(val);
        x->nrow = INTEGER(val)[0];
        x->ncol = INTEGER(val)[1];
    }
    return val;
}
SEXP Rf_ScalarReal(double v) {
    SEXP s = new_vec(REALSXP, 1);
    REAL(s)[0] = v;
    return s;
}
SEXP Rf_lang3(SEXP f, SEXP a, SEXP b) {
    SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
    s->type = LANGSXP;
    s->car[0] = f;
    s->car[1] = a;
    s->car[2] = b;
    return s;
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
    longjmp(toplevel_jmp ? *toplevel_jmp : error_jmp, 1);   /* what R does: unwind to the innermost context */
}
/* the only closure of this runtime: progress(done, total) */
static SEXP eval_progress(SEXP call) {
    progress_calls++;
    fprintf(stderr, "progress %d: %.0f / %.0f\n", progress_calls, REAL(call->car[1])[0], REAL(call->car[2])[0]);
    if (progress_calls == env_int("MOCK_PROGRESS_FAIL_AT", -1)) Rf_error("progress bar exploded");
    return R_NilValue;
}
SEXP R_tryEval(SEXP call, SEXP env, int* failed) {
    (void)env;
    jmp_buf here, *outer = toplevel_jmp;
    const int depth = protect_depth;
    *failed = 0;
    if (setjmp(here)) {
        toplevel_jmp = outer;
        protect_depth = depth;   /* R restores the protect stack of the context */
        *failed = 1;
        return NULL;
    }
    toplevel_jmp = &here;
    SEXP r = eval_progress(call);
    toplevel_jmp = outer;
    return r;
}
void R_CheckUserInterrupt(void) {
    interrupt_checks++;
    if (interrupt_checks == env_int("MOCK_INTERRUPT_AT", -1)) {
        snprintf(error_msg, sizeof error_msg, "<interrupt>");
        longjmp(toplevel_jmp ? *toplevel_jmp : error_jmp, 1);
    }
}
Rboolean R_ToplevelExec(void (*fun)(void*), void* data) {
    jmp_buf here, *outer = toplevel_jmp;
    const int depth = protect_depth;
    if (setjmp(here)) {
        toplevel_jmp = outer;
        protect_depth = depth;
        return FALSE;
    }
    toplevel_jmp = &here;
    fun(data);
    toplevel_jmp = outer;
    return TRUE;
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
void R_unload_InterpolateR(DllInfo* dll);
typedef SEXP (*call9)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*call8)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*call10)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*call0)(void);

static SEXP read_real(FILE* f, R_xlen_t n) {
    SEXP s = new_vec(REALSXP, n);
    if (fread(s->data, sizeof(double), (size_t)n, f) != (size_t)n) {
        fprintf(stderr, "short input\n");
        exit(2);
    }
    return s;
}

static DL_FUNC lookup(const char* want, int nargs) {
    for (const R_CallMethodDef* m = registered; m && m->name; m++)
        if (strcmp(m->name, want) == 0 && m->numArgs == nargs) return m->fun;
    fprintf(stderr, "%s is not registered with %d arguments\n", want, nargs);
    exit(2);
}

/* R's garbage collector releasing a vector that came from allocVector3 */
static void gc_release(SEXP s) {
    if (s && s->block) {
        s->alloc.mem_free(&s->alloc, s->block);
        pinned_frees++;
        s->block = NULL;
    }
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
    obs->ndim = 2;
    SEXP last = (mode == 2 || mode == 3) ? R_NilValue : read_real(f, mode == 0 ? 1 : n_radii);
    SEXP k_model = R_NilValue, k_nug = R_NilValue, k_sill = R_NilValue, k_range = R_NilValue;
    if (mode == 3) {
        SEXP codes = read_real(f, n_dates);
        k_model = new_vec(INTSXP, n_dates);
        for (int i = 0; i < n_dates; i++) INTEGER(k_model)[i] = (int)REAL(codes)[i];
        k_nug = read_real(f, n_dates);
        k_sill = read_real(f, n_dates);
        k_range = read_real(f, n_dates);
    }
    fclose(f);
    SEXP pinned = new_vec(LGLSXP, 1);
    INTEGER(pinned)[0] = env_int("MOCK_PINNED", 0);
    SEXP progress = R_NilValue;
    if (env_int("MOCK_PROGRESS", 0)) {
        progress = (SEXP)calloc(1, sizeof(struct SEXPREC));
        progress->type = CLOSXP;
    }

    R_init_InterpolateR(NULL);   /* library(InterpolateR): routine registration */
    SEXP volatile out = NULL;   /* assigned between setjmp and a possible longjmp */
    int code = 0;
    if (setjmp(error_jmp)) {
        fprintf(stderr, "R error: %s\n", error_msg);
        code = 3;
    } else {
        /* .Call(...) */
        if (mode == 3)
            out = ((call10)lookup("C_ipr_kriging", 10))(cell_x, cell_y, st_x, st_y, obs, k_model, k_nug, k_sill, k_range,
                                                        R_NilValue);
        else if (mode == 2)
            out = ((call8)lookup("C_ipr_idw_kriging", 8))(cell_x, cell_y, st_x, st_y, obs, R_NilValue, progress, pinned);
        else
            out = ((call9)lookup(mode == 0 ? "C_ipr_idw" : "C_ipr_cressman", 9))(cell_x, cell_y, st_x, st_y, obs, last,
                                                                               R_NilValue, progress, pinned);
        if (protect_depth != 0) {
            fprintf(stderr, "protect stack imbalance: %d\n", protect_depth);
            return 4;
        }
        const R_xlen_t expect = n_cells * n_dates * (mode == 1 ? n_radii : 1);
        const int want_dims = mode == 1 ? 3 : 2;
        SEXP o = out;
        if (!Rf_isReal(o) || XLENGTH(o) != expect || o->ndim != want_dims || o->nrow != (int)n_cells ||
            o->ncol != n_dates) {
            fprintf(stderr, "unexpected result shape\n");
            return 4;
        }
        if (env_int("MOCK_PINNED", 0) && mode != 3 && pinned_allocs != 1) {
            fprintf(stderr, "pinned = TRUE did not go through the allocator\n");
            return 4;
        }
        FILE* g = fopen(argv[2], "wb");
        if (!g) return 2;
        fwrite(REAL(o), sizeof(double), (size_t)expect, g);
        fclose(g);
    }
    gc_release(out);
    ((call0)lookup("C_ipr_release", 0))();
    R_unload_InterpolateR(NULL);
    fprintf(stderr, "interrupt checks %d, progress calls %d, pinned allocs %d frees %d\n", interrupt_checks,
            progress_calls, pinned_allocs, pinned_frees);
    return code;
}
