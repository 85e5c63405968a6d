/*
 * interpolater_b200 — C ABI of the B200-native IDW / Cressman engine.
 *
 * This is the drop-in boundary for InterpolateR's data-parallel hot path.  The reference is pure
 * R and has NO FFI today (NAMESPACE has no useDynLib, there is no src/); the seam is created by
 * replacing two R code regions with one .Call each:
 *
 *   ipr_idw_f64       replaces  R/IDW.R:249-355       (grid_xy -> dist_mat -> w_base ->
 *                                                      pblapply(call_idw) -> list of rasters)
 *   ipr_cressman_f64  replaces  R/Cressman.R:268-345  (dist_mat -> weights_list ->
 *                                                      pblapply(call_crsm_opt) -> per-radius stacks)
 *
 * Plain pointers and sizes only; no torch / R / C++ types cross the boundary.  All matrices are
 * column-major doubles exactly as R stores them.  Missing values on input are any NaN (R's
 * is.na() is TRUE for NA_real_ and NaN); missing values on output are written with R's
 * NA_real_ bit pattern 0x7FF00000000007A2.
 *
 * Every entry point returns 0 on success or a negative IPR_E* code, and writes a message to
 * `err` (if non-NULL).  There is NO CPU fallback: without a usable sm_100 device the calls fail.
 * No C++ exception leaves the library (host allocation failures come back as IPR_ENOMEM), and the
 * library never calls back into the host language.
 *
 * Threading: the one-shot calls may be issued concurrently from several threads; a plan must be
 * used by one thread at a time.  Device work buffers, streams and pinned staging slots are cached
 * process-wide between calls (ipr_release_cached_memory() returns the device memory).
 */
#ifndef INTERPOLATER_B200_H
#define INTERPOLATER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPR_OK 0
#define IPR_EINVAL (-1)   /* bad argument */
#define IPR_ECUDA (-2)    /* CUDA runtime error / no device */
#define IPR_ENOMEM (-3)   /* device or pinned-host allocation failed */
#define IPR_EUNSUPPORTED (-4)
#define IPR_ECANCELLED (-5)  /* ipr_job_cancel() stopped the job */

/* library version: major*10000 + minor*100 + patch */
int ipr_version(void);
/* number of visible CUDA devices of compute capability 10.x; negative on error */
int ipr_device_count(void);
/* R's NA_real_ as written by the engine */
double ipr_na_real(void);

/* ---------------------------------------------------------------------------------------
 * One-shot host-buffer entry points (what the R .Call shim binds).
 *
 *   cell_x, cell_y   n_cells      grid-cell centres in terra cell order (R/IDW.R:249-257)
 *   st_x, st_y       n_st         training-station coordinates, aligned with obs columns
 *   obs              n_dates x n_st, column-major (each station's series contiguous, as the
 *                    columns of BD_Obs are); NaN/NA = missing
 *   devices          n_devices CUDA ordinals (NULL / 0 => device 0); the cells are split into
 *                    contiguous ranges (64-cell tiles) over them, every device computing all dates
 *                    of its cells and writing its own columns of every layer: nothing is computed
 *                    twice and there is no inter-GPU exchange
 *   out              caller-owned, column-major n_cells x n_dates (one layer per date); page-locked
 *                    memory receives the D2H copies directly, pageable memory (an R vector) is
 *                    filled through a pinned staging ring with a multi-threaded first touch
 * ------------------------------------------------------------------------------------- */
int ipr_idw_f64(const double* cell_x, const double* cell_y, int64_t n_cells,
                const double* st_x, const double* st_y, int32_t n_st,
                const double* obs, int32_t n_dates, double p,
                const int* devices, int n_devices,
                double* out, char* err, size_t errlen);

/* radii_m: n_radii search radii in metres, ascending and unique (R/Cressman.R:216-218 does the
 * km->m, sort(unique()) before the seam).  out: n_radii consecutive column-major
 * n_cells x n_dates stacks (radius-major). */
int ipr_cressman_f64(const double* cell_x, const double* cell_y, int64_t n_cells,
                     const double* st_x, const double* st_y, int32_t n_st,
                     const double* obs, int32_t n_dates,
                     const double* radii_m, int32_t n_radii,
                     const int* devices, int n_devices,
                     double* out, char* err, size_t errlen);

/* Kriging_Ordinary's IDW fallback (first slice): what perform_kriging computes for a date whose
 * variogram is flat or whose kriging matrix is singular / ill-conditioned (R/Kriging_Ordinary.R:465-476), with
 * its per-date prelude (:406-423).  Per date, over the stations available on it:
 *   fewer than two  -> the layer is their mean (0 when there is none)
 *   all equal       -> the layer is that value
 *   otherwise       -> sum(w * v) / sum(w), w = 1 / d^2 with d == 0 replaced by 1e-10
 * Never NA.  The caller passes the dates for which it has chosen the fallback (the choice depends on the fitted
 * variogram and stays in R); arguments as for ipr_idw_f64. */
int ipr_idw_kriging_f64(const double* cell_x, const double* cell_y, int64_t n_cells,
                        const double* st_x, const double* st_y, int32_t n_st,
                        const double* obs, int32_t n_dates,
                        const int* devices, int n_devices,
                        double* out, char* err, size_t errlen);

/* Ordinary kriging, the prediction step of perform_kriging for every date (R/Kriging_Ordinary.R:405-508).
 * The empirical variogram, its fit (optim) and the decision to fall back stay in R; per date t the caller passes
 *   model[t]  0 = fall back to IDW (flat variogram, det == 0 or kappa > 1e12, :428-462)
 *             1 exponential | 2 spherical | 3 gaussian | 4 linear        (variogram_models, :262-275)
 *   nugget[t], sill[t], range[t]   the fitted parameters (ignored when model[t] == 0)
 * Per date: the prelude of ipr_idw_kriging_f64 (constant layers), then for model > 0
 *   Gamma = variogram matrix of the n stations available on the date (nugget on the diagonal, Lagrange row / column),
 *   alpha = Gamma^-1 [v; 0]                       ONE (n+1)^2 LU per date — the reference solves per cell
 *   out[c, t] = max(0, alpha_n + sum_i alpha_i gamma_t(d(c, s_i)))       = max(0, sum(w * v)) with Gamma w = [gamma(c); 1]
 * and mean(v) where Gamma is exactly singular (the reference's tryCatch, :494).  The two orders of solving agree to
 * kappa(Gamma) * 2^-53. */
int ipr_kriging_f64(const double* cell_x, const double* cell_y, int64_t n_cells,
                    const double* st_x, const double* st_y, int32_t n_st,
                    const double* obs, int32_t n_dates,
                    const int32_t* model, const double* nugget, const double* sill, const double* range,
                    const int* devices, int n_devices,
                    double* out, char* err, size_t errlen);

/* ---------------------------------------------------------------------------------------
 * The same two calls with the reference's next steps fused on the device, so that a large stack
 * needs no further host pass:
 *   round_mode 0 none | 1 R's round(v, n_round) as in terra::app(x, function(v) round(v, n)),
 *              R/IDW.R:356-358 | 2 terra's round(SpatRaster, n), R/Cressman.R:353-355
 *   cell_keep  NULL, or n_cells bytes: cells with 0 are written as NA (terra::mask by the study
 *              area, R/IDW.R:376-381 — the caller rasterises the polygon)
 * The reference validates on the rounded, UN-masked stack (R/IDW.R:364-372): a caller that masks
 * here evaluates the few test-station cells with a separate (tiny) call.
 * ------------------------------------------------------------------------------------- */
int ipr_idw_post_f64(const double* cell_x, const double* cell_y, int64_t n_cells,
                     const double* st_x, const double* st_y, int32_t n_st,
                     const double* obs, int32_t n_dates, double p,
                     int32_t round_mode, int32_t n_round, const uint8_t* cell_keep,
                     const int* devices, int n_devices,
                     double* out, char* err, size_t errlen);
int ipr_cressman_post_f64(const double* cell_x, const double* cell_y, int64_t n_cells,
                          const double* st_x, const double* st_y, int32_t n_st,
                          const double* obs, int32_t n_dates,
                          const double* radii_m, int32_t n_radii,
                          int32_t round_mode, int32_t n_round, const uint8_t* cell_keep,
                          const int* devices, int n_devices,
                          double* out, char* err, size_t errlen);

/* ---------------------------------------------------------------------------------------
 * Asynchronous form of the one-shot calls, for a host language whose main thread must stay
 * responsive: the reference wraps its per-date loop in a pbapply progress bar (R/IDW.R:350-352,
 * R/Cressman.R:338-340) and R code is interruptible between dates.  ipr_job_start_* takes the
 * arguments of ipr_idw_post_f64 / ipr_cressman_post_f64 (round_mode 0 and cell_keep NULL give the
 * plain calls), starts the job on worker threads and returns at once; every array except radii_m must
 * stay valid until ipr_job_finish().  The main thread then alternates
 *     ipr_job_wait(job, 100, &done, &total)   -> 1 finished | 0 still running; progress in work items
 *                                                (one 128-date chunk of one stack on one device, counted
 *                                                when its layers have reached the host)
 * with its own event handling (R: R_CheckUserInterrupt inside R_ToplevelExec, progress bar update), may
 * call ipr_job_cancel() (the workers stop at the next chunk; what was written so far stays), and
 * always ends with ipr_job_finish(), which joins, frees the job and returns its code
 * (IPR_ECANCELLED after a cancel).  timeout_ms < 0 waits until the job has finished.
 * ------------------------------------------------------------------------------------- */
typedef struct ipr_job ipr_job;
int ipr_job_start_idw(ipr_job** job,
                      const double* cell_x, const double* cell_y, int64_t n_cells,
                      const double* st_x, const double* st_y, int32_t n_st,
                      const double* obs, int32_t n_dates, double p,
                      int32_t round_mode, int32_t n_round, const uint8_t* cell_keep,
                      const int* devices, int n_devices,
                      double* out, char* err, size_t errlen);
int ipr_job_start_cressman(ipr_job** job,
                           const double* cell_x, const double* cell_y, int64_t n_cells,
                           const double* st_x, const double* st_y, int32_t n_st,
                           const double* obs, int32_t n_dates,
                           const double* radii_m, int32_t n_radii,
                           int32_t round_mode, int32_t n_round, const uint8_t* cell_keep,
                           const int* devices, int n_devices,
                           double* out, char* err, size_t errlen);
int ipr_job_wait(ipr_job* job, int timeout_ms, int64_t* done, int64_t* total);
void ipr_job_cancel(ipr_job* job);
int ipr_job_finish(ipr_job* job, char* err, size_t errlen);

/* The cell ranges the one-shot calls give to n_devices devices: bounds[0..n_devices], device g owns
 * cells [bounds[g], bounds[g+1]) — contiguous, on 64-cell tile boundaries, equal to within one tile
 * (no output element depends on another cell, so there is nothing to exchange). */
void ipr_cell_split_bounds(int64_t n_cells, int n_devices, int64_t* bounds);

/* ---------------------------------------------------------------------------------------
 * Device-resident plan API (used by the one-shot calls, the benchmark and the Python host
 * mirror): inputs stay in HBM, outputs are written to caller-supplied DEVICE pointers.
 * A plan is bound to one device and owns one compute stream.
 * ------------------------------------------------------------------------------------- */
typedef struct ipr_plan ipr_plan;

int ipr_plan_create(ipr_plan** plan, int device,
                    const double* cell_x, const double* cell_y, int64_t n_cells,
                    const double* st_x, const double* st_y, int32_t n_st,
                    int32_t n_dates, char* err, size_t errlen);
void ipr_plan_destroy(ipr_plan* plan);

/* H2D copy of obs (host, column-major n_dates x n_st) followed by the device-side preparation
 * (NA -> 0 / mask split, per-date early-out flags).  Asynchronous w.r.t. the host except for a
 * n_dates-byte flag read-back. */
int ipr_plan_set_obs(ipr_plan* plan, const double* obs, char* err, size_t errlen);
/* same, obs already on the device (n_dates x n_st column-major) */
int ipr_plan_set_obs_device(ipr_plan* plan, const double* d_obs, char* err, size_t errlen);

/* Interpolate dates [t0, t1) into d_out (DEVICE pointer): layer t is written at
 * d_out + (t - t0) * ld_out, ld_out >= n_cells.  Launches are asynchronous on the plan stream. */
int ipr_plan_idw(ipr_plan* plan, double p, int32_t t0, int32_t t1,
                 double* d_out, int64_t ld_out, char* err, size_t errlen);
/* the Kriging fallback rule of ipr_idw_kriging_f64 */
int ipr_plan_idw_kriging(ipr_plan* plan, int32_t t0, int32_t t1,
                         double* d_out, int64_t ld_out, char* err, size_t errlen);
/* ordinary kriging (ipr_kriging_f64); model / nugget / sill / range are HOST arrays of n_dates entries */
int ipr_plan_kriging(ipr_plan* plan, const int32_t* model, const double* nugget, const double* sill, const double* range,
                     int32_t t0, int32_t t1, double* d_out, int64_t ld_out, char* err, size_t errlen);
/* radius ri's stack starts at d_out + ri * stack_stride (elements) */
int ipr_plan_cressman(ipr_plan* plan, const double* radii_m, int32_t n_radii,
                      int32_t t0, int32_t t1, double* d_out, int64_t ld_out,
                      int64_t stack_stride, char* err, size_t errlen);

/* Drops the cached A operand (the cell x station weights, kept in HBM between calls because the
 * geometry of a plan never changes), so that the next call regenerates it.  The benchmark calls
 * this every step so that a timed step contains the whole hot path. */
void ipr_plan_reset_cache(ipr_plan* plan);

/* post-pass applied by later ipr_plan_idw / ipr_plan_cressman calls (see ipr_idw_post_f64);
 * cell_keep is a HOST pointer (copied), NULL clears the mask */
int ipr_plan_set_post(ipr_plan* plan, int32_t round_mode, int32_t n_round, const uint8_t* cell_keep,
                      char* err, size_t errlen);

int ipr_plan_sync(ipr_plan* plan);
/* CUDA events on the plan's stream, for timing exactly what the plan launched */
int ipr_plan_timer_start(ipr_plan* plan);
int ipr_plan_timer_stop(ipr_plan* plan, float* elapsed_ms); /* synchronises */
/* counters since plan creation: kernels launched by this library on the plan's stream */
int64_t ipr_plan_launch_count(const ipr_plan* plan);
/* Cumulative counts, since plan creation, of (64-cell tile, station stage) blocks of the A
 * operand — 16-station stages for IDW, 8-station half stages for Cressman: `dense` = what the dense
 * product visits, `executed` = blocks holding any non-zero weight (all of them for IDW; for Cressman
 * the blocks wholly beyond the radius are skipped).  Only the ratio of differences taken around
 * calls of ONE kind is meaningful; it lets a caller report executed vs dense-equivalent flops.
 * Synchronises the plan stream. */
int ipr_plan_stage_blocks(ipr_plan* plan, uint64_t* executed, uint64_t* dense);
/* Per-phase CUDA-event timing of what the plan launches, on the plan's own stream (off by default: an event
 * pair costs a few microseconds per launch group).  ipr_plan_profile_read() synchronises the stream, adds the
 * elapsed milliseconds of every finished span to ms[phase] (n >= IPR_N_PHASES doubles, zeroed by the call
 * first) and forgets them.  This is how bench.py measures the dominant kernels live for the roofline. */
#define IPR_PHASE_PREP 0       /* obs preparation: NA split, flags, NA lists, byte mask */
#define IPR_PHASE_WEIGHTS 1    /* A operand generation (weights_kernel) */
#define IPR_PHASE_CONTRACT 2   /* FP64 DMMA contraction (IDW: unmasked, numerator or masked GEMM) */
#define IPR_PHASE_SLICE 3      /* digit planes of the weights (int8 mask path) */
#define IPR_PHASE_MASK_GEMM 4  /* tcgen05 kind::i8 mask contraction */
#define IPR_PHASE_FIXUP 5      /* zero-distance / sparse-NA / direct fixes and the conditional FP64 fallback */
#define IPR_PHASE_POST 6       /* optional rounding / polygon mask pass */
#define IPR_PHASE_RADIUS0 8    /* Cressman: the contraction of radius r is phase IPR_PHASE_RADIUS0 + r, r < 24 */
#define IPR_N_PHASES 32
int ipr_plan_profile(ipr_plan* plan, int enable);
int ipr_plan_profile_read(ipr_plan* plan, double* ms, int n);
/* number of (cell, station) pairs at exactly zero distance found for this geometry */
int64_t ipr_plan_zero_pairs(const ipr_plan* plan);
/* raw cudaStream_t of the plan (as void*) so a host framework can order its own work */
void* ipr_plan_stream(const ipr_plan* plan);

/* The library keeps freed multi-MB device work buffers for reuse by later calls (cudaMalloc/cudaFree
 * of the 16 GB weight cache and the output ring cost up to hundreds of ms per call); this returns
 * them to the driver.  IPR_POOL_MAX_GB (default 64) bounds what is kept. */
void ipr_release_cached_memory(void);

/* pinned host memory helpers for callers that want zero-copy-staging D2H (the benchmark) */
void* ipr_host_alloc(size_t bytes);
void ipr_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* INTERPOLATER_B200_H */
