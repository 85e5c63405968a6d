/*
 * .Call shim between R and the interpolater_b200 C ABI (include/interpolater_b200.h).
 *
 * This file is the WHOLE R-facing native surface: argument validation, the job loop, error translation.
 * It cannot be compiled in the build container (no R headers there); it is type-checked and EXECUTED against a
 * mock of R's C API (tests/mock_r, tests/test_r_shim_*.py).  Registered routines:
 *
 *   C_ipr_idw(cell_x, cell_y, st_x, st_y, obs, p, devices, progress, pinned)            n_cells x n_dates
 *   C_ipr_cressman(cell_x, cell_y, st_x, st_y, obs, radii_m, devices, progress, pinned) n_cells x n_dates x n_radii
 *   C_ipr_idw_kriging(cell_x, cell_y, st_x, st_y, obs, devices, progress, pinned)       n_cells x n_dates
 *   C_ipr_kriging(cell_x, cell_y, st_x, st_y, obs, model, nugget, sill, range, devices)  n_cells x n_dates
 *   C_ipr_release()                                                                      NULL
 *
 * obs is the n_dates x n_stations double matrix (column-major, NA = missing) whose columns are aligned with
 * st_x / st_y.  NA cells of the result carry NA_real_.
 *
 * Threading: the core runs the job on its own worker threads (ipr_job_start_*) and never calls
 * into R.  The R main thread stays in this file: every 100 ms it polls the job, lets R see a pending user
 * interrupt (R_CheckUserInterrupt inside R_ToplevelExec, so that the job can be cancelled and joined BEFORE the
 * interrupt unwinds the stack) and calls `progress(done, total)` — the reference wraps the same loop in a
 * pbapply timer bar (R/IDW.R:350-352, R/Cressman.R:338-340).  An interrupt or an error in `progress` cancels
 * the job at its next 128-date chunk; the R error is raised only after ipr_job_finish() has joined the workers.
 *
 * Result memory: by default an ordinary R allocation (pageable; the core fills it through its pinned staging
 * ring).  pinned = TRUE allocates the vector through an R_allocator_t backed by cudaHostAlloc, so that the D2H
 * copies land in the REALSXP directly and R's GC returns the pages through the allocator; measured on the B200
 * box, page-locking 29 GB costs 12 s (and 2.7 s to release) against 0.4-0.8 s saved per call, so it only pays
 * for results that are small or for hosts where pinning is cheap — it is an option, not the default.
 */
#include <limits.h>
#include <stdint.h>

#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rallocators.h>
#include <R_ext/Rdynload.h>

#include "interpolater_b200.h"

static void check_real(SEXP x, const char* what) {
    if (!isReal(x)) error("%s must be a double vector", what);
}

static int device_list(SEXP devices, int** out) {
    if (isNull(devices) || LENGTH(devices) == 0) {
        *out = NULL;
        return 0;
    }
    if (!isInteger(devices)) error("devices must be an integer vector of CUDA ordinals");
    *out = INTEGER(devices);
    return LENGTH(devices);
}

/* ---- result allocation -------------------------------------------------------------------------------- */
static void* pinned_alloc(R_allocator_t* a, size_t bytes) {
    (void)a;
    return ipr_host_alloc(bytes);   /* NULL makes R raise its own "cannot allocate" error */
}
static void pinned_free(R_allocator_t* a, void* p) {
    (void)a;
    ipr_host_free(p);
}

/* n_cells x n_dates [x n_stacks] doubles with the dim attribute set; protected by the caller */
static SEXP alloc_result(R_xlen_t n_cells, int n_dates, int n_stacks, int as_array, int pinned) {
    const R_xlen_t n = n_cells * (R_xlen_t)n_dates * (R_xlen_t)n_stacks;
    SEXP out;
    if (pinned) {
        R_allocator_t al = {pinned_alloc, pinned_free, NULL, NULL};   /* copied by allocVector3 */
        out = PROTECT(allocVector3(REALSXP, n, &al));
    } else {
        out = PROTECT(allocVector(REALSXP, n));
    }
    SEXP dims = PROTECT(allocVector(INTSXP, as_array ? 3 : 2));
    INTEGER(dims)[0] = (int)n_cells;
    INTEGER(dims)[1] = n_dates;
    if (as_array) INTEGER(dims)[2] = n_stacks;
    setAttrib(out, R_DimSymbol, dims);
    UNPROTECT(2);
    return out;
}

/* ---- the job loop on R's main thread ------------------------------------------------------------------ */
static void check_interrupt(void* unused) {
    (void)unused;
    R_CheckUserInterrupt();
}

/* returns the job's code; *interrupted / *callback_failed say why it was cancelled, if it was */
static int drive_job(ipr_job* job, SEXP progress, int* interrupted, int* callback_failed, char* err, size_t errlen) {
    int64_t done = 0, total = 0, shown = -1;
    *interrupted = 0;
    *callback_failed = 0;
    for (;;) {
        const int finished = ipr_job_wait(job, 100, &done, &total);
        if (!*interrupted && !*callback_failed) {
            if (!R_ToplevelExec(check_interrupt, NULL)) {
                *interrupted = 1;
                ipr_job_cancel(job);
            } else if (!isNull(progress) && done != shown) {
                int failed = 0;
                SEXP d = PROTECT(ScalarReal((double)done));
                SEXP t = PROTECT(ScalarReal((double)total));
                SEXP call = PROTECT(lang3(progress, d, t));
                R_tryEval(call, R_GlobalEnv, &failed);
                UNPROTECT(3);
                shown = done;
                if (failed) {
                    *callback_failed = 1;
                    ipr_job_cancel(job);
                }
            }
        }
        if (finished) break;
    }
    return ipr_job_finish(job, err, errlen);   /* joins the workers; nothing of the job is alive after this */
}

static void raise_job_error(int rc, int interrupted, int callback_failed, const char* err) {
    if (interrupted) error("interpolater_b200: interrupted by the user (the GPU job was cancelled)");
    if (callback_failed) error("interpolater_b200: the progress callback failed (the GPU job was cancelled)");
    if (rc != IPR_OK) error("interpolater_b200 (code %d): %s", rc, err);
}

typedef struct {
    R_xlen_t n_cells;
    int n_st, n_dates;
    int* dev;
    int n_dev;
} common_args;

static common_args check_common(SEXP cell_x, SEXP cell_y, SEXP st_x, SEXP st_y, SEXP obs, SEXP devices, SEXP progress) {
    common_args a;
    check_real(cell_x, "cell_x"); check_real(cell_y, "cell_y");
    check_real(st_x, "st_x");     check_real(st_y, "st_y");
    check_real(obs, "obs");
    a.n_cells = XLENGTH(cell_x);
    a.n_st = LENGTH(st_x);
    if (XLENGTH(cell_y) != a.n_cells || LENGTH(st_y) != a.n_st) error("coordinate vectors differ in length");
    if (a.n_cells > INT_MAX) error("more than 2^31 - 1 cells: a matrix dimension cannot hold that");
    if (!isMatrix(obs) || ncols(obs) != a.n_st) error("obs must be an n_dates x n_stations matrix");
    a.n_dates = nrows(obs);
    if (!isNull(progress) && !isFunction(progress)) error("progress must be NULL or a function(done, total)");
    a.dev = NULL;
    a.n_dev = device_list(devices, &a.dev);
    return a;
}

SEXP C_ipr_idw(SEXP cell_x, SEXP cell_y, SEXP st_x, SEXP st_y, SEXP obs, SEXP p, SEXP devices, SEXP progress,
               SEXP pinned) {
    const common_args a = check_common(cell_x, cell_y, st_x, st_y, obs, devices, progress);
    check_real(p, "p");
    SEXP out = PROTECT(alloc_result(a.n_cells, a.n_dates, 1, 0, asLogical(pinned) == TRUE));
    char err[512] = "";
    ipr_job* job = NULL;
    int interrupted = 0, cb_failed = 0;
    int rc = ipr_job_start_idw(&job, REAL(cell_x), REAL(cell_y), (int64_t)a.n_cells, REAL(st_x), REAL(st_y), a.n_st,
                               REAL(obs), a.n_dates, asReal(p), 0, 0, NULL, a.dev, a.n_dev, REAL(out), err, sizeof err);
    if (rc == IPR_OK) rc = drive_job(job, progress, &interrupted, &cb_failed, err, sizeof err);
    UNPROTECT(1);
    raise_job_error(rc, interrupted, cb_failed, err);
    return out;
}

SEXP C_ipr_idw_kriging(SEXP cell_x, SEXP cell_y, SEXP st_x, SEXP st_y, SEXP obs, SEXP devices, SEXP progress,
                       SEXP pinned) {
    /* the fallback dates of a kriging run are few: the blocking call is enough (no progress, no cancel) */
    const common_args a = check_common(cell_x, cell_y, st_x, st_y, obs, devices, progress);
    SEXP out = PROTECT(alloc_result(a.n_cells, a.n_dates, 1, 0, asLogical(pinned) == TRUE));
    char err[512] = "";
    const int rc = ipr_idw_kriging_f64(REAL(cell_x), REAL(cell_y), (int64_t)a.n_cells, REAL(st_x), REAL(st_y), a.n_st,
                                       REAL(obs), a.n_dates, a.dev, a.n_dev, REAL(out), err, sizeof err);
    UNPROTECT(1);
    raise_job_error(rc, 0, 0, err);
    return out;
}

/* ordinary kriging: model (integer, 0 = IDW fallback, 1..4 = exponential / spherical / gaussian / linear) and the
 * fitted nugget / sill / range, one entry per date (row of obs) */
SEXP C_ipr_kriging(SEXP cell_x, SEXP cell_y, SEXP st_x, SEXP st_y, SEXP obs, SEXP model, SEXP nugget, SEXP sill,
                   SEXP range, SEXP devices) {
    const common_args a = check_common(cell_x, cell_y, st_x, st_y, obs, devices, R_NilValue);
    check_real(nugget, "nugget"); check_real(sill, "sill"); check_real(range, "range");
    if (!isInteger(model)) error("model must be an integer vector (0 = IDW fallback, 1..4 = variogram model)");
    if (LENGTH(model) != a.n_dates || LENGTH(nugget) != a.n_dates || LENGTH(sill) != a.n_dates ||
        LENGTH(range) != a.n_dates)
        error("model, nugget, sill and range need one entry per date (row of obs)");
    SEXP out = PROTECT(alloc_result(a.n_cells, a.n_dates, 1, 0, 0));
    char err[512] = "";
    const int rc = ipr_kriging_f64(REAL(cell_x), REAL(cell_y), (int64_t)a.n_cells, REAL(st_x), REAL(st_y), a.n_st,
                                   REAL(obs), a.n_dates, INTEGER(model), REAL(nugget), REAL(sill), REAL(range), a.dev,
                                   a.n_dev, REAL(out), err, sizeof err);
    UNPROTECT(1);
    raise_job_error(rc, 0, 0, err);
    return out;
}

SEXP C_ipr_cressman(SEXP cell_x, SEXP cell_y, SEXP st_x, SEXP st_y, SEXP obs, SEXP radii_m, SEXP devices,
                    SEXP progress, SEXP pinned) {
    const common_args a = check_common(cell_x, cell_y, st_x, st_y, obs, devices, progress);
    check_real(radii_m, "radii_m");
    const int n_radii = LENGTH(radii_m);
    SEXP out = PROTECT(alloc_result(a.n_cells, a.n_dates, n_radii, 1, asLogical(pinned) == TRUE));
    char err[512] = "";
    ipr_job* job = NULL;
    int interrupted = 0, cb_failed = 0;
    int rc = ipr_job_start_cressman(&job, REAL(cell_x), REAL(cell_y), (int64_t)a.n_cells, REAL(st_x), REAL(st_y),
                                    a.n_st, REAL(obs), a.n_dates, REAL(radii_m), n_radii, 0, 0, NULL, a.dev, a.n_dev,
                                    REAL(out), err, sizeof err);
    if (rc == IPR_OK) rc = drive_job(job, progress, &interrupted, &cb_failed, err, sizeof err);
    UNPROTECT(1);
    raise_job_error(rc, interrupted, cb_failed, err);
    return out;
}

/* returns the device buffers the library keeps between calls (tens of GB after a large job) to the driver */
SEXP C_ipr_release(void) {
    ipr_release_cached_memory();
    return R_NilValue;
}

static const R_CallMethodDef call_methods[] = {
    {"C_ipr_idw", (DL_FUNC)&C_ipr_idw, 9},
    {"C_ipr_cressman", (DL_FUNC)&C_ipr_cressman, 9},
    {"C_ipr_idw_kriging", (DL_FUNC)&C_ipr_idw_kriging, 8},
    {"C_ipr_kriging", (DL_FUNC)&C_ipr_kriging, 10},
    {"C_ipr_release", (DL_FUNC)&C_ipr_release, 0},
    {NULL, NULL, 0}};

void R_init_InterpolateR(DllInfo* dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}

void R_unload_InterpolateR(DllInfo* dll) {
    (void)dll;
    ipr_release_cached_memory();
}
