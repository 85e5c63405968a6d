/*
 * .Call shim between R and the interpolater_b200 C ABI (include/interpolater_b200.h).
 *
 * This file is the WHOLE R-facing native surface: argument validation, two calls into the R-free
 * core, error translation.  It cannot be compiled in the build container (no R headers there);
 * it is kept trivially reviewable on purpose.  Registered routines:
 *
 *   C_ipr_idw(cell_x, cell_y, st_x, st_y, obs, p, devices)            -> n_cells x n_dates matrix
 *   C_ipr_cressman(cell_x, cell_y, st_x, st_y, obs, radii_m, devices) -> n_cells x n_dates x n_radii array
 *
 * obs is the n_dates x n_stations double matrix (column-major, NA = missing) whose columns are
 * aligned with st_x/st_y.  The result is an R-allocated (pageable) REALSXP that the core fills in
 * place through its pinned staging ring; NA cells carry NA_real_.  The call blocks until the result
 * is complete (no R_CheckUserInterrupt inside).
 * The core never calls into R; on failure the shim raises an R error AFTER the core has cleaned up.
 */
#include <limits.h>

#include <R.h>
#include <Rinternals.h>
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

SEXP C_ipr_idw(SEXP cell_x, SEXP cell_y, SEXP st_x, SEXP st_y, SEXP obs, SEXP p, SEXP devices) {
    check_real(cell_x, "cell_x"); check_real(cell_y, "cell_y");
    check_real(st_x, "st_x");     check_real(st_y, "st_y");
    check_real(obs, "obs");       check_real(p, "p");
    const R_xlen_t n_cells = XLENGTH(cell_x);
    const int n_st = LENGTH(st_x);
    if (XLENGTH(cell_y) != n_cells || LENGTH(st_y) != n_st) error("coordinate vectors differ in length");
    if (n_cells > INT_MAX) error("more than 2^31 - 1 cells: a matrix dimension cannot hold that");
    if (!isMatrix(obs) || ncols(obs) != n_st) error("obs must be an n_dates x n_stations matrix");
    const int n_dates = nrows(obs);
    int* dev = NULL;
    const int n_dev = device_list(devices, &dev);
    SEXP out = PROTECT(allocMatrix(REALSXP, (int)n_cells, n_dates));
    char err[512] = "";
    const int rc = ipr_idw_f64(REAL(cell_x), REAL(cell_y), (int64_t)n_cells, REAL(st_x), REAL(st_y), n_st,
                               REAL(obs), n_dates, asReal(p), dev, n_dev, REAL(out), err, sizeof err);
    UNPROTECT(1);
    if (rc != IPR_OK) error("interpolater_b200 (code %d): %s", rc, err);
    return out;
}

SEXP C_ipr_cressman(SEXP cell_x, SEXP cell_y, SEXP st_x, SEXP st_y, SEXP obs, SEXP radii_m, SEXP devices) {
    check_real(cell_x, "cell_x"); check_real(cell_y, "cell_y");
    check_real(st_x, "st_x");     check_real(st_y, "st_y");
    check_real(obs, "obs");       check_real(radii_m, "radii_m");
    const R_xlen_t n_cells = XLENGTH(cell_x);
    const int n_st = LENGTH(st_x);
    if (XLENGTH(cell_y) != n_cells || LENGTH(st_y) != n_st) error("coordinate vectors differ in length");
    if (n_cells > INT_MAX) error("more than 2^31 - 1 cells: a matrix dimension cannot hold that");
    if (!isMatrix(obs) || ncols(obs) != n_st) error("obs must be an n_dates x n_stations matrix");
    const int n_dates = nrows(obs);
    const int n_radii = LENGTH(radii_m);
    int* dev = NULL;
    const int n_dev = device_list(devices, &dev);
    SEXP dims = PROTECT(allocVector(INTSXP, 3));
    INTEGER(dims)[0] = (int)n_cells; INTEGER(dims)[1] = n_dates; INTEGER(dims)[2] = n_radii;
    SEXP out = PROTECT(allocArray(REALSXP, dims));
    char err[512] = "";
    const int rc = ipr_cressman_f64(REAL(cell_x), REAL(cell_y), (int64_t)n_cells, REAL(st_x), REAL(st_y), n_st,
                                    REAL(obs), n_dates, REAL(radii_m), n_radii, dev, n_dev, REAL(out), err,
                                    sizeof err);
    UNPROTECT(2);
    if (rc != IPR_OK) error("interpolater_b200 (code %d): %s", rc, err);
    return out;
}

static const R_CallMethodDef call_methods[] = {
    {"C_ipr_idw", (DL_FUNC)&C_ipr_idw, 7},
    {"C_ipr_cressman", (DL_FUNC)&C_ipr_cressman, 7},
    {NULL, NULL, 0}};

void R_init_InterpolateR(DllInfo* dll) {
    R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}
