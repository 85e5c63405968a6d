"""Array-level Python front of the CUDA engine (thin: argument marshalling only).

``idw_grid`` / ``cressman_grid`` are the exact Python twins of the two one-shot C entry points the
R ``.Call`` shim binds (host buffers in, host buffers out); ``Plan`` wraps the device-resident
plan API used by the benchmark.  All compute happens in ``lib/libinterpolater_b200.so``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native


def _f64(a, name):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim != 1:
        raise ValueError(f"{name} must be one-dimensional")
    return a


def _obs_colmajor(obs, n_st):
    """obs as (n_dates, n_st) -> Fortran-ordered float64 (each station's series contiguous)."""
    obs = np.asarray(obs, dtype=np.float64)
    if obs.ndim != 2 or obs.shape[1] != n_st:
        raise ValueError(f"obs must have shape (n_dates, {n_st}); got {obs.shape}")
    return np.asfortranarray(obs)


def _devices(devices):
    if devices is None:
        return None, 0
    arr = (C.c_int * len(devices))(*[int(d) for d in devices])
    return arr, len(devices)


ROUND_NONE, ROUND_R, ROUND_TERRA = 0, 1, 2


def _post_args(n_round, round_mode, cell_keep, n_cells):
    if n_round is None:
        round_mode, n_round = ROUND_NONE, 0
    keep = None
    if cell_keep is not None:
        keep = np.ascontiguousarray(cell_keep, dtype=np.uint8)
        if keep.shape != (n_cells,):
            raise ValueError("cell_keep must have one entry per cell")
    return int(round_mode), int(n_round), keep


def idw_grid(cell_x, cell_y, st_x, st_y, obs, p=2.0, devices=None, out=None, n_round=None, round_mode=ROUND_R,
             cell_keep=None):
    """IDW over every cell x date.  Returns (n_cells, n_dates) float64, Fortran order (one layer
    per date, like the values of a terra stack); NaN (R's NA_real_ payload) where the reference
    yields NA.  Replaces R/IDW.R:249-355."""
    lib = _native.load()
    cx, cy = _f64(cell_x, "cell_x"), _f64(cell_y, "cell_y")
    sx, sy = _f64(st_x, "st_x"), _f64(st_y, "st_y")
    if cx.shape != cy.shape or sx.shape != sy.shape:
        raise ValueError("coordinate arrays must have matching lengths")
    ob = _obs_colmajor(obs, sx.shape[0])
    n_dates = ob.shape[0]
    if out is None:
        out = np.empty((cx.shape[0], n_dates), dtype=np.float64, order="F")
    elif out.shape != (cx.shape[0], n_dates) or not out.flags.f_contiguous or out.dtype != np.float64:
        raise ValueError("out must be a Fortran-ordered float64 (n_cells, n_dates) array")
    dev, ndev = _devices(devices)
    err = _native.errbuf()
    rmode, nr, keep = _post_args(n_round, round_mode, cell_keep, cx.shape[0])
    if rmode == ROUND_NONE and keep is None:
        rc = lib.ipr_idw_f64(cx.ctypes.data, cy.ctypes.data, cx.shape[0], sx.ctypes.data, sy.ctypes.data, sx.shape[0],
                             ob.ctypes.data, n_dates, float(p), dev, ndev, out.ctypes.data, err, _native.ERRLEN)
    else:
        rc = lib.ipr_idw_post_f64(cx.ctypes.data, cy.ctypes.data, cx.shape[0], sx.ctypes.data, sy.ctypes.data,
                                  sx.shape[0], ob.ctypes.data, n_dates, float(p), rmode, nr,
                                  keep.ctypes.data if keep is not None else None, dev, ndev, out.ctypes.data, err,
                                  _native.ERRLEN)
    _native.check(rc, err)
    return out


def idw_kriging_fallback_grid(cell_x, cell_y, st_x, st_y, obs, devices=None, out=None):
    """The IDW fallback of Kriging_Ordinary's perform_kriging (R/Kriging_Ordinary.R:406-423, :465-476) over every
    cell x date: constant layers for dates with fewer than two (or only equal) available values, otherwise the
    inverse-square-distance mean with d == 0 -> 1e-10.  Returns (n_cells, n_dates) float64, Fortran order."""
    lib = _native.load()
    cx, cy = _f64(cell_x, "cell_x"), _f64(cell_y, "cell_y")
    sx, sy = _f64(st_x, "st_x"), _f64(st_y, "st_y")
    if cx.shape != cy.shape or sx.shape != sy.shape:
        raise ValueError("coordinate arrays must have matching lengths")
    ob = _obs_colmajor(obs, sx.shape[0])
    n_dates = ob.shape[0]
    if out is None:
        out = np.empty((cx.shape[0], n_dates), dtype=np.float64, order="F")
    elif out.shape != (cx.shape[0], n_dates) or not out.flags.f_contiguous or out.dtype != np.float64:
        raise ValueError("out must be a Fortran-ordered float64 (n_cells, n_dates) array")
    dev, ndev = _devices(devices)
    err = _native.errbuf()
    rc = lib.ipr_idw_kriging_f64(cx.ctypes.data, cy.ctypes.data, cx.shape[0], sx.ctypes.data, sy.ctypes.data,
                                 sx.shape[0], ob.ctypes.data, n_dates, dev, ndev, out.ctypes.data, err, _native.ERRLEN)
    _native.check(rc, err)
    return out


#: variogram model codes of ipr_kriging_f64 (0 = the date falls back to IDW)
VARIOGRAM_MODELS = {"idw": 0, "exponential": 1, "spherical": 2, "gaussian": 3, "linear": 4}


def kriging_grid(cell_x, cell_y, st_x, st_y, obs, model, nugget, sill, range_, devices=None, out=None):
    """Ordinary kriging prediction for every cell x date (the prediction step of perform_kriging,
    R/Kriging_Ordinary.R:405-508).  model / nugget / sill / range_: one entry per date — the variogram the caller
    fitted for it (model by name or code; "idw" / 0 = the date falls back to IDW).  Returns (n_cells, n_dates)
    float64, Fortran order; never NA."""
    lib = _native.load()
    cx, cy = _f64(cell_x, "cell_x"), _f64(cell_y, "cell_y")
    sx, sy = _f64(st_x, "st_x"), _f64(st_y, "st_y")
    if cx.shape != cy.shape or sx.shape != sy.shape:
        raise ValueError("coordinate arrays must have matching lengths")
    ob = _obs_colmajor(obs, sx.shape[0])
    n_dates = ob.shape[0]
    md = np.ascontiguousarray([VARIOGRAM_MODELS[m] if isinstance(m, str) else int(m) for m in np.atleast_1d(model)],
                              dtype=np.int32)
    par = [np.ascontiguousarray(np.atleast_1d(v), dtype=np.float64) for v in (nugget, sill, range_)]
    if md.shape != (n_dates,) or any(v.shape != (n_dates,) for v in par):
        raise ValueError("model, nugget, sill and range_ need one entry per date")
    if out is None:
        out = np.empty((cx.shape[0], n_dates), dtype=np.float64, order="F")
    elif out.shape != (cx.shape[0], n_dates) or not out.flags.f_contiguous or out.dtype != np.float64:
        raise ValueError("out must be a Fortran-ordered float64 (n_cells, n_dates) array")
    dev, ndev = _devices(devices)
    err = _native.errbuf()
    rc = lib.ipr_kriging_f64(cx.ctypes.data, cy.ctypes.data, cx.shape[0], sx.ctypes.data, sy.ctypes.data, sx.shape[0],
                             ob.ctypes.data, n_dates, md.ctypes.data, par[0].ctypes.data, par[1].ctypes.data,
                             par[2].ctypes.data, dev, ndev, out.ctypes.data, err, _native.ERRLEN)
    _native.check(rc, err)
    return out


def cressman_grid(cell_x, cell_y, st_x, st_y, obs, radii_m, devices=None, out=None, n_round=None,
                  round_mode=ROUND_TERRA, cell_keep=None):
    """Cressman analysis for each radius (metres, ascending unique).  Returns
    (n_radii, n_cells, n_dates) with each [ri] Fortran-ordered.  Replaces R/Cressman.R:268-345."""
    lib = _native.load()
    cx, cy = _f64(cell_x, "cell_x"), _f64(cell_y, "cell_y")
    sx, sy = _f64(st_x, "st_x"), _f64(st_y, "st_y")
    if cx.shape != cy.shape or sx.shape != sy.shape:
        raise ValueError("coordinate arrays must have matching lengths")
    ob = _obs_colmajor(obs, sx.shape[0])
    n_dates = ob.shape[0]
    rad = _f64(radii_m, "radii_m")
    n_cells = cx.shape[0]
    if out is None:
        flat = np.empty(rad.shape[0] * n_cells * n_dates, dtype=np.float64)
    else:
        flat = out
        if flat.ndim != 1 or flat.shape[0] != rad.shape[0] * n_cells * n_dates or flat.dtype != np.float64:
            raise ValueError("out must be a flat float64 array of n_radii*n_cells*n_dates elements")
    dev, ndev = _devices(devices)
    err = _native.errbuf()
    rmode, nr, keep = _post_args(n_round, round_mode, cell_keep, n_cells)
    if rmode == ROUND_NONE and keep is None:
        rc = lib.ipr_cressman_f64(cx.ctypes.data, cy.ctypes.data, n_cells, sx.ctypes.data, sy.ctypes.data, sx.shape[0],
                                  ob.ctypes.data, n_dates, rad.ctypes.data, rad.shape[0], dev, ndev,
                                  flat.ctypes.data, err, _native.ERRLEN)
    else:
        rc = lib.ipr_cressman_post_f64(cx.ctypes.data, cy.ctypes.data, n_cells, sx.ctypes.data, sy.ctypes.data,
                                       sx.shape[0], ob.ctypes.data, n_dates, rad.ctypes.data, rad.shape[0], rmode, nr,
                                       keep.ctypes.data if keep is not None else None, dev, ndev, flat.ctypes.data,
                                       err, _native.ERRLEN)
    _native.check(rc, err)
    # radius-major, each stack column-major (n_cells fastest): view as (n_radii, n_dates, n_cells)^T
    return flat.reshape(rad.shape[0], n_dates, n_cells).transpose(0, 2, 1)


def cell_split_bounds(n_cells, n_devices):
    """The cell ranges the one-shot calls give to `n_devices` GPUs (ipr_cell_split_bounds): device g owns
    cells [b[g], b[g+1])."""
    b = (C.c_int64 * (n_devices + 1))()
    _native.load().ipr_cell_split_bounds(int(n_cells), int(n_devices), b)
    return [int(v) for v in b]


class Job:
    """Asynchronous one-shot call (ipr_job_*): the job runs on the library's worker threads while the caller
    polls — the shape the R shim uses to keep R's main thread free for the progress bar of R/IDW.R:350-352
    and for user interrupts.  ``Job.idw(...)`` / ``Job.cressman(...)`` start, ``wait(ms)`` polls and
    returns (finished, done, total), ``cancel()`` asks to stop, ``finish()`` joins and returns the array
    (raises NativeError, code IPR_ECANCELLED after a cancel)."""

    def __init__(self):
        self._lib = _native.load()
        self._h = C.c_void_p()
        self._keep = None
        self._result = None

    @classmethod
    def idw(cls, cell_x, cell_y, st_x, st_y, obs, p=2.0, devices=None, n_round=None, round_mode=ROUND_R,
            cell_keep=None):
        self = cls()
        cx, cy = _f64(cell_x, "cell_x"), _f64(cell_y, "cell_y")
        sx, sy = _f64(st_x, "st_x"), _f64(st_y, "st_y")
        ob = _obs_colmajor(obs, sx.shape[0])
        out = np.empty((cx.shape[0], ob.shape[0]), dtype=np.float64, order="F")
        dev, ndev = _devices(devices)
        rmode, nr, keep = _post_args(n_round, round_mode, cell_keep, cx.shape[0])
        err = _native.errbuf()
        self._keep = (cx, cy, sx, sy, ob, keep, dev, out)   # must outlive the job
        rc = self._lib.ipr_job_start_idw(C.byref(self._h), cx.ctypes.data, cy.ctypes.data, cx.shape[0], sx.ctypes.data,
                                         sy.ctypes.data, sx.shape[0], ob.ctypes.data, ob.shape[0], float(p), rmode, nr,
                                         keep.ctypes.data if keep is not None else None, dev, ndev, out.ctypes.data,
                                         err, _native.ERRLEN)
        _native.check(rc, err)
        self._result = out
        return self

    @classmethod
    def cressman(cls, cell_x, cell_y, st_x, st_y, obs, radii_m, devices=None, n_round=None, round_mode=ROUND_TERRA,
                 cell_keep=None):
        self = cls()
        cx, cy = _f64(cell_x, "cell_x"), _f64(cell_y, "cell_y")
        sx, sy = _f64(st_x, "st_x"), _f64(st_y, "st_y")
        ob = _obs_colmajor(obs, sx.shape[0])
        rad = _f64(radii_m, "radii_m")
        flat = np.empty(rad.shape[0] * cx.shape[0] * ob.shape[0], dtype=np.float64)
        dev, ndev = _devices(devices)
        rmode, nr, keep = _post_args(n_round, round_mode, cell_keep, cx.shape[0])
        err = _native.errbuf()
        self._keep = (cx, cy, sx, sy, ob, keep, dev, flat)
        rc = self._lib.ipr_job_start_cressman(C.byref(self._h), cx.ctypes.data, cy.ctypes.data, cx.shape[0],
                                              sx.ctypes.data, sy.ctypes.data, sx.shape[0], ob.ctypes.data, ob.shape[0],
                                              rad.ctypes.data, rad.shape[0], rmode, nr,
                                              keep.ctypes.data if keep is not None else None, dev, ndev,
                                              flat.ctypes.data, err, _native.ERRLEN)
        _native.check(rc, err)
        self._result = flat.reshape(rad.shape[0], ob.shape[0], cx.shape[0]).transpose(0, 2, 1)
        return self

    def wait(self, timeout_ms=100):
        done, total = C.c_int64(), C.c_int64()
        fin = self._lib.ipr_job_wait(self._h, int(timeout_ms), C.byref(done), C.byref(total))
        if fin < 0:
            raise _native.NativeError(fin, "job handle is not valid")
        return bool(fin), int(done.value), int(total.value)

    def cancel(self):
        self._lib.ipr_job_cancel(self._h)

    def finish(self):
        err = _native.errbuf()
        h, self._h = self._h, C.c_void_p()
        rc = self._lib.ipr_job_finish(h, err, _native.ERRLEN)
        res, self._result, self._keep = self._result, None, None
        _native.check(rc, err)
        return res

    def __del__(self):
        try:
            if self._h:
                self.cancel()
                self._lib.ipr_job_finish(self._h, None, 0)
        except Exception:
            pass


class Plan:
    """Device-resident plan (include/interpolater_b200.h, ipr_plan_*): inputs live in HBM, outputs
    are written to caller-supplied device pointers (e.g. ``torch.Tensor.data_ptr()``)."""

    def __init__(self, device, cell_x, cell_y, st_x, st_y, n_dates):
        self._lib = _native.load()
        cx, cy = _f64(cell_x, "cell_x"), _f64(cell_y, "cell_y")
        sx, sy = _f64(st_x, "st_x"), _f64(st_y, "st_y")
        self.n_cells, self.n_st, self.n_dates = cx.shape[0], sx.shape[0], int(n_dates)
        self._h = C.c_void_p()
        err = _native.errbuf()
        rc = self._lib.ipr_plan_create(C.byref(self._h), int(device), cx.ctypes.data, cy.ctypes.data, self.n_cells,
                                       sx.ctypes.data, sy.ctypes.data, self.n_st, self.n_dates, err, _native.ERRLEN)
        _native.check(rc, err)

    def close(self):
        if self._h:
            self._lib.ipr_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_obs(self, obs):
        ob = _obs_colmajor(obs, self.n_st)
        if ob.shape[0] != self.n_dates:
            raise ValueError("obs has the wrong number of dates")
        err = _native.errbuf()
        _native.check(self._lib.ipr_plan_set_obs(self._h, ob.ctypes.data, err, _native.ERRLEN), err)

    def set_obs_device(self, d_ptr):
        err = _native.errbuf()
        _native.check(self._lib.ipr_plan_set_obs_device(self._h, C.c_void_p(d_ptr), err, _native.ERRLEN), err)

    def idw(self, d_out, p=2.0, t0=0, t1=None, ld_out=None):
        t1 = self.n_dates if t1 is None else t1
        ld_out = self.n_cells if ld_out is None else ld_out
        err = _native.errbuf()
        _native.check(self._lib.ipr_plan_idw(self._h, float(p), t0, t1, C.c_void_p(d_out), ld_out, err,
                                             _native.ERRLEN), err)

    def cressman(self, d_out, radii_m, t0=0, t1=None, ld_out=None, stack_stride=None):
        t1 = self.n_dates if t1 is None else t1
        ld_out = self.n_cells if ld_out is None else ld_out
        stack_stride = ld_out * (t1 - t0) if stack_stride is None else stack_stride
        rad = _f64(radii_m, "radii_m")
        err = _native.errbuf()
        _native.check(self._lib.ipr_plan_cressman(self._h, rad.ctypes.data, rad.shape[0], t0, t1, C.c_void_p(d_out),
                                                  ld_out, stack_stride, err, _native.ERRLEN), err)

    def reset_cache(self):
        """Forget the cached A operand (weights) so that the next call regenerates it."""
        self._lib.ipr_plan_reset_cache(self._h)

    def sync(self):
        if self._lib.ipr_plan_sync(self._h) != 0:
            raise _native.NativeError(-2, "stream synchronisation failed")

    def timer_start(self):
        self._lib.ipr_plan_timer_start(self._h)

    def timer_stop(self) -> float:
        ms = C.c_float()
        if self._lib.ipr_plan_timer_stop(self._h, C.byref(ms)) != 0:
            raise _native.NativeError(-2, "timer stop failed")
        return float(ms.value)

    @property
    def launch_count(self) -> int:
        return int(self._lib.ipr_plan_launch_count(self._h))

    def stage_blocks(self):
        """(executed, dense) cumulative counts of A-operand blocks (see ipr_plan_stage_blocks)."""
        ex, de = C.c_uint64(), C.c_uint64()
        if self._lib.ipr_plan_stage_blocks(self._h, C.byref(ex), C.byref(de)) != 0:
            raise _native.NativeError(-2, "stage block query failed")
        return int(ex.value), int(de.value)

    def profile(self, enable=True):
        """Switch the per-phase CUDA-event timing of the plan's launches on or off."""
        self._lib.ipr_plan_profile(self._h, 1 if enable else 0)

    def profile_read(self):
        """Milliseconds per phase since the last read (synchronises): {"prep", "weights", "contract", "slice",
        "mask_gemm", "fixup", "post", "radius": [ms per Cressman radius]}."""
        ms = (C.c_double * _native.N_PHASES)()
        if self._lib.ipr_plan_profile_read(self._h, ms, _native.N_PHASES) != 0:
            raise _native.NativeError(-2, "profile read failed")
        res = {k: float(ms[i]) for k, i in _native.PHASES.items()}
        res["radius"] = [float(ms[i]) for i in range(_native.PHASE_RADIUS0, _native.N_PHASES)]
        return res

    @property
    def zero_pairs(self) -> int:
        return int(self._lib.ipr_plan_zero_pairs(self._h))

    @property
    def stream(self) -> int:
        return int(self._lib.ipr_plan_stream(self._h) or 0)
