#!/usr/bin/env python
"""Benchmark of the IDW / Cressman hot path: interpolated cell-days per second.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|idw|cressman|idw_masked]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # the reference algorithm on the host cores

One "step" is one pass of the hot path over the whole synthetic workload (SURVEY §8d cfg 3/4/5:
1,000,000 cells x 2,000 stations x 3,650 days).  The top-level `value` / `roofline` / `e2e` are cfg 3
(IDW p=2, the configuration the metric is quoted on); with the default `--workload all` the same JSON
line carries full sub-records for cfg 4 (Cressman, 4 radii) and cfg 5 (IDW, 20 % NA) under `workloads`.

`value` is the device-resident throughput (raw inputs already in HBM, outputs left in HBM, CUDA events
on the plan's own stream, max over ranks); `e2e` is the same job through the host-buffer C ABI call
(`ipr_idw_f64` / `ipr_cressman_f64`), H2D of the inputs and D2H of every output layer inside the timed
region.  Multi-GPU (default `--scaling strong`): ONE 1M x 2,000 x 3,650 job, the cells split over the
GPUs on 64-cell tiles, no collective on the data path (NCCL carries only the barrier and the
max-over-ranks of the timings); `--scaling weak` gives every GPU its own 3,650-day block instead.
"""
from __future__ import annotations

import argparse
import ctypes as Ct
import json
import os
import shutil
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "interpolated cell-days/sec (IDW p=2, Cressman) at 1/2/4/8 B200; % of roofline"
UNIT = "cell-days/s"
RADII_M = [25e3, 50e3, 100e3, 200e3]  # cfg 4
WORKLOADS = ("idw", "cressman", "idw_masked")
PARITY_CELLS = 1000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", *WORKLOADS],
                    help="all: cfg 3 IDW at the top level + cfg 4 / cfg 5 sub-records under `workloads`")
    ap.add_argument("--ncol", type=int, default=1000)
    ap.add_argument("--nrow", type=int, default=1000)
    ap.add_argument("--stations", type=int, default=2000)
    ap.add_argument("--dates", type=int, default=3650)
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default): ONE job, the cells are split over the GPUs on 64-cell tiles; "
                         "weak: every GPU owns its own --dates-day block of the record")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--pageable-steps", type=int, default=3,
                    help="steps of the pageable-output e2e variant (what R's own allocation would see); 0 = skip")
    ap.add_argument("--inproc-steps", type=int, default=3,
                    help="N > 1: steps of the one-process, device-list variant of e2e run by rank 0; 0 = skip")
    ap.add_argument("--cpu-sample-cells", type=int, default=10000)
    ap.add_argument("--cpu-sample-dates", type=int, default=48)
    return ap.parse_args()


def workload_name(a, wl):
    c = a.ncol * a.nrow
    base = f"{c:,} cells x {a.stations:,} stations x {a.dates:,} days"
    if wl == "idw":
        return f"synthetic IDW (cfg 3): {base}, p=2, no NA"
    if wl == "idw_masked":
        return (f"synthetic IDW (cfg 5): {base}, p=2, 20% per-date NA (masked contraction); all {a.stations:,} stations "
                "train here — the training=0.8 hold-out + validation of configs[4] runs at full size in "
                "tests/test_gpu_parity.py::test_full_size_cfg5_through_the_drop_in_signature")
    return f"synthetic Cressman (cfg 4): {base}, 4 radii (25/50/100/200 km)"


def make_config(a, wl, world):
    """`config` of the JSON line: identical for the engine's arm and the reference arm."""
    from interpolater_b200 import sharding

    C = a.ncol * a.nrow
    weak = a.scaling == "weak"
    b = sharding.cell_block_bounds(C, 1 if weak else world)
    return {
        "workload": workload_name(a, wl),
        "n_cells": C, "n_cells_rank0": b[1] - b[0], "n_stations": a.stations,
        "n_dates_total": a.dates * (world if weak else 1), "n_dates_rank0": a.dates,
        "n_stacks": len(RADII_M) if wl == "cressman" else 1,
        "sharding": ("one independent %d-day date block per GPU" % a.dates if weak else
                     "%d cells split over %d GPU(s) on 64-cell tiles, every GPU all %d days" % (C, world, a.dates))
                    + ", no data-path collective",
        "l2": "outputs (%.1f GB/step) >> 126 MB L2: every step streams through and flushes L2"
              % ((len(RADII_M) if wl == "cressman" else 1) * C * a.dates * 8e-9),
    }


def make_inputs(a, wl, rank):
    from interpolater_b200 import synthetic

    na = 0.2 if wl == "idw_masked" else 0.0
    cx, cy, sx, sy, obs = synthetic.synthetic_case(a.ncol, a.nrow, a.stations, a.dates, na_frac=na,
                                                   seed=20260101 + 1000 * rank)
    if rank > 0:  # same geometry on every rank; only the date block (obs) differs
        _, _, sx, sy, _ = synthetic.synthetic_case(a.ncol, a.nrow, a.stations, 1, seed=20260101)
    return cx, cy, sx, sy, obs


# ------------------------------------------------------------------------------------------
# clocks sampler (NVML) — runs during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._th = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
            0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
            0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit and nm != "gpu_idle":
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv is not None:
            self._th = threading.Thread(target=self._loop, daemon=True)
            self._th.start()

    def stop(self):
        self._stop.set()
        if self._th:
            self._th.join()
        s = sorted(self.samples)
        return {
            "sm_mhz": (s[len(s) // 2] if s else None),
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(s),
        }


# ------------------------------------------------------------------------------------------
# CPU baseline: the oracle's literal restatement of the reference algorithm on a bounded sample
# ------------------------------------------------------------------------------------------
def cpu_reference_sample(a, wl, threads, steps=1):
    """Times oracle.reference_port (dense dist/weight matrices + per-date loop, exactly the
    reference's algorithm) on `cpu_sample_cells` cells x `cpu_sample_dates` dates of the workload,
    all stations.  threads > 1 runs independent date sub-blocks in a thread pool (NumPy releases the
    GIL inside the dense passes), which is how the serial R loop would be spread over cores."""
    from concurrent.futures import ThreadPoolExecutor

    from interpolater_b200 import synthetic
    from oracle import reference_port as ref  # the timed CPU baseline (allowed use of oracle/)

    side = int(round(a.cpu_sample_cells ** 0.5))
    nd = a.cpu_sample_dates
    na = 0.2 if wl == "idw_masked" else 0.0
    cx, cy, sx, sy, obs = synthetic.synthetic_case(side, side, a.stations, nd, na_frac=na)
    ncell = cx.shape[0]

    radii = np.asarray(RADII_M)

    ref.LONG_DOUBLE_ROWSUMS = False  # timing leg: float64 row sums (R's C loop is at least that fast)
    n_full = a.dates

    def one_pass():
        # date-independent setup once (as the reference does), then the per-date loop; with
        # threads > 1 the dates are spread over a pool sharing the dense matrices
        t0 = time.perf_counter()
        if wl == "cressman":
            weights = ref.cressman_setup(cx, cy, sx, sy, radii)
            fn = lambda t: ref.cressman_one_date(weights, obs[t])  # noqa: E731
        else:
            dist_mat, w_base = ref.idw_setup(cx, cy, sx, sy, 2.0)
            fn = lambda t: ref.idw_one_date(dist_mat, w_base, obs[t])  # noqa: E731
        t1 = time.perf_counter()
        if threads <= 1:
            for t in range(nd):
                fn(t)
        else:
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(fn, range(nd)))
        t2 = time.perf_counter()
        return t1 - t0, (t2 - t1) / nd

    setup_s, date_s = [], []
    # BLAS itself single-threaded (R's reference BLAS is); with threads > 1 the
    # parallelism is the pool over dates, so `cores` is exactly the number of busy threads
    try:
        from threadpoolctl import threadpool_limits
        blas_limit = threadpool_limits(limits=1)
    except Exception:
        blas_limit = None
    for _ in range(steps):
        s_, d_ = one_pass()
        setup_s.append(s_)
        date_s.append(d_)
    if blas_limit is not None:
        blas_limit.restore_original_limits()
    ref.LONG_DOUBLE_ROWSUMS = True
    setup = float(np.mean(setup_s))
    per_date = float(np.mean(date_s))
    # the reference amortises its dense setup over all dates of the job: extrapolate the sample's
    # per-date cost linearly to the workload's full date count
    sec_full = setup + per_date * n_full
    return {
        "value": ncell * n_full / sec_full,
        "unit": UNIT,
        "cores": threads,
        "kind": "port",
        "sample": f"{ncell} cells x {a.stations} stations; dense setup timed once ({setup:.2f} s) + per-date loop "
                  f"timed on {nd} dates ({per_date * 1e3:.1f} ms/date), extrapolated linearly to {n_full} dates "
                  f"(oracle/reference_port.py = the reference's algorithm in NumPy/OpenBLAS; R is not installed)",
        "seconds": setup + per_date * nd,
    }


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(max(a.gpus, 1))))
    threads = os.cpu_count() or 1
    wls = list(WORKLOADS) if a.workload == "all" else [a.workload]
    recs = {}
    for i, wl in enumerate(wls):
        if i == 0:
            for _ in range(min(a.warmup, 1)):
                cpu_reference_sample(a, wl, threads, 1)
        # the other workloads are context for the sub-records: one bounded pass each
        recs[wl] = cpu_reference_sample(a, wl, threads, max(1, a.steps) if i == 0 else 1)
    res = recs[wls[0]]
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": res["value"],
        "unit": UNIT,
        "n_gpus": a.gpus,
        "steps": a.steps,
        "warmup": a.warmup,
        "ms_per_step": res["seconds"] * 1e3,
        "higher_is_better": True,
        "scaling": a.scaling,
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": make_config(a, wls[0], world),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "rscript": shutil.which("Rscript"),
    }
    if len(wls) > 1:
        line["workloads"] = {
            wl: {"value": recs[wl]["value"], "unit": UNIT, "config": make_config(a, wl, world),
                 "cpu_baseline": {k: recs[wl][k] for k in ("value", "unit", "cores", "kind", "sample")}}
            for wl in wls[1:]
        }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# the engine's arm
# ------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide state shared by the workloads of one bench run."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        from interpolater_b200 import _native

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.cpu_group = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            # a CPU-side group: ranks that must WAIT while rank 0 drives every GPU from one process block on a
            # socket, not inside an NCCL kernel spinning on their GPU (which would time-slice with rank 0's work)
            self.cpu_group = dist.new_group(backend="gloo")
        self.lib = _native.load()
        with open(os.path.join(ROOT, "profiles", "fp64_peaks.json")) as f:
            self.peaks = json.load(f)
        self.hbm_peak_gbs, self.hbm_peak_src = 6546.2, "fallback: round-1 MEASURED_PEAKS.json value"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                self.hbm_peak_gbs = float(json.load(f)["hbm_gbs"])
                self.hbm_peak_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
        self.traffic = {}
        try:
            with open(os.path.join(ROOT, "profiles", "dram_traffic.json")) as f:
                self.traffic = json.load(f)
        except Exception:
            pass
        self._pin_ptr, self._pin_bytes = None, 0

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def cpu_barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group)

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def pinned(self, nbytes):
        """One caller-owned pinned output buffer, reused by all workloads (cudaHostAlloc of tens of GB takes seconds)."""
        if self._pin_bytes < nbytes:
            if self._pin_ptr:
                self.lib.ipr_host_free(self._pin_ptr)
            self._pin_ptr = self.lib.ipr_host_alloc(nbytes)
            if not self._pin_ptr:
                raise SystemExit("pinned host allocation failed")
            self._pin_bytes = nbytes
        return self._pin_ptr

    def close(self):
        if self._pin_ptr:
            self.lib.ipr_host_free(self._pin_ptr)
            self._pin_ptr, self._pin_bytes = None, 0
        if self.world > 1:
            self.dist.destroy_process_group()


def parity_sample(wl, out, cx, cy, sx, sy, obs, radii):
    """Seeded sample of cells x ALL dates of the timed output (still in HBM) against the oracle's matrix
    form (oracle/reference_port.py; held to the literal per-date restatement by tests/test_oracle.py).
    The oracle is the checker here, never the thing measured."""
    import torch

    from oracle import reference_port as ref

    C = cx.shape[0]
    rng = np.random.default_rng(12345)
    idx = np.sort(rng.choice(C, size=min(PARITY_CELLS, C), replace=False))
    got = out[:, :, torch.from_numpy(idx).cuda()].cpu().numpy()       # (n_stacks, D, n_sample)
    if wl == "cressman":
        want = ref.cressman_reference_matrix(cx[idx], cy[idx], sx, sy, obs, radii)
    else:
        want = [ref.idw_reference_matrix(cx[idx], cy[idx], sx, sy, obs, 2.0)]
    worst, na_equal, n_val, n_na = 0.0, True, 0, 0
    for s, w in enumerate(want):
        g = got[s].T                                                    # (n_sample, D)
        na_g, na_w = np.isnan(g), np.isnan(w)
        na_equal &= bool(np.array_equal(na_g, na_w))
        m = ~na_w & ~na_g
        if m.any():
            rel = np.abs(g[m] - w[m]) / np.maximum(np.abs(w[m]), 1e-12)
            worst = max(worst, float(rel.max()))
        n_val += int(m.sum())
        n_na += int(na_w.sum())
    return {"parity_sample_max_rel": worst, "parity_sample_na_pattern_equal": na_equal,
            "parity_sample": "%d seeded cells x all %d dates x %d stack(s) of the timed output vs "
                             "oracle/reference_port.py (matrix form): %d values, %d NA; tolerance 1e-9"
                             % (idx.size, obs.shape[0], len(want), n_val, n_na),
            "parity_sample_ok": bool(na_equal and worst <= 1e-9)}


def run_workload(a, wl, ctx):
    torch = ctx.torch
    from interpolater_b200 import _native, engine, sharding

    lib, world, rank, local_rank = ctx.lib, ctx.world, ctx.rank, ctx.local_rank
    weak = a.scaling == "weak"
    cx, cy, sx, sy, obs = make_inputs(a, wl, rank if weak else 0)
    total_cells = cx.shape[0]
    total_dates = obs.shape[0] * (world if weak else 1)
    cx_all, cy_all = cx, cy
    if not weak:
        # one job over all GPUs: contiguous cell ranges (nothing computed twice, equal shares)
        b = sharding.cell_block_bounds(total_cells, world)
        cx, cy = np.ascontiguousarray(cx[b[rank]:b[rank + 1]]), np.ascontiguousarray(cy[b[rank]:b[rank + 1]])
        if cx.shape[0] == 0:
            raise SystemExit("strong scaling: fewer 64-cell tiles than GPUs")
    C, S, D = cx.shape[0], sx.shape[0], obs.shape[0]
    n_stacks = len(RADII_M) if wl == "cressman" else 1
    radii = np.asarray(RADII_M)
    int8_mask = os.environ.get("IPR_MASK_INT8", "1") != "0"

    # ---------------- device-resident arm (`value`) ----------------
    plan = engine.Plan(local_rank, cx, cy, sx, sy, D)
    obs_f = np.asfortranarray(obs)
    d_obs = torch.from_numpy(obs_f.T.copy()).cuda()  # (n_st, n_dates) row-major == column-major n_dates x n_st
    out = torch.empty((n_stacks, D, C), dtype=torch.float64, device="cuda")

    def step():
        # the whole hot path from HBM-resident raw inputs: NA split + early-out flags, weights, contraction
        # (+ fix-ups), outputs left in HBM
        plan.reset_cache()  # weights (A operand) and digit planes are regenerated inside every step
        plan.set_obs_device(d_obs.data_ptr())
        if wl == "cressman":
            plan.cressman(out.data_ptr(), radii)
        else:
            plan.idw(out.data_ptr(), 2.0)

    for _ in range(a.warmup):
        step()
    plan.sync()
    sampler = ClockSampler(local_rank)
    ctx.barrier()
    l0 = plan.launch_count
    ex0, de0 = plan.stage_blocks()
    plan.profile(True)   # per-phase CUDA events on the plan stream, inside the timed region (a few us per launch group)
    plan.profile_read()
    sampler.start()
    plan.timer_start()
    for _ in range(a.steps):
        step()
    ms_total = plan.timer_stop()
    ctx.barrier()
    clocks = sampler.stop()
    phases = plan.profile_read()
    plan.profile(False)
    launches = plan.launch_count - l0
    ex1, de1 = plan.stage_blocks()
    ms_step = ctx.max_over_ranks(ms_total / a.steps)
    value = total_cells * total_dates / (ms_step * 1e-3)
    ph = {k: (v / a.steps if k != "radius" else [x / a.steps for x in v[:n_stacks]]) for k, v in phases.items()}

    # ---- roofline of the dominant kernel(s), from the phase timings of the timed region itself ----
    dmma_peak = ctx.peaks["dmma_tflops"]
    peak_src = ("profiles/fp64_peaks.json: DMMA m8n8k4 probe measured on this pool's B200 (MEASURED_PEAKS.json "
                "has no FP64 entry); cuBLAS DGEMM 8192^3 = %.1f TFLOP/s" % ctx.peaks["cublas_dgemm_tflops_burst"])
    cd = float(C) * D
    extra = {}
    if wl == "cressman":
        # (cell tile, half stage) blocks beyond the radius are exactly zero and are skipped: the FP64 pipe
        # executes dense * executed_block_frac; the small radii are bound by the store of their stack
        executed_frac = (ex1 - ex0) / max(de1 - de0, 1)
        kernel_ms = sum(ph["radius"])
        dense_flops = 2.0 * S * cd * n_stacks
        achieved = dense_flops * executed_frac / (kernel_ms * 1e-3) * 1e-12
        roofline = {
            "bound": "tensor", "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s", "frac": achieved / dmma_peak,
            "traffic": ctx.traffic.get(wl),
            "kernel": "ipr::cressman_kernel<8,2> (persistent, 2 CTAs/SM; per cell tile: station list + weights, then "
                      "gathered-row FP64 DMMA.8x8x4 contraction over all date tiles; one launch per radius)",
            "flops_counted": "executed: 2 x 8 stations x 64 cells x 128 dates per pipeline stage run (a cell tile's "
                             "stations within reach, padded to whole stages), as a fraction of the dense 2*S per "
                             "cell-day and radius; the dense-equivalent rate is in `dense_equivalent_tflops`",
            "executed_block_frac": executed_frac,
            "dense_equivalent_tflops": dense_flops / (kernel_ms * 1e-3) * 1e-12,
            "kernel_ms": kernel_ms,
            "per_radius": [
                {"radius_km": RADII_M[i] / 1e3, "ms": ph["radius"][i],
                 "store_gbs": cd * 8 / (ph["radius"][i] * 1e-3) * 1e-9,
                 "store_frac_of_hbm_peak": cd * 8 / (ph["radius"][i] * 1e-3) * 1e-9 / ctx.hbm_peak_gbs}
                for i in range(n_stacks)],
            "hbm_peak_gbs": ctx.hbm_peak_gbs, "hbm_peak_source": ctx.hbm_peak_src,
            "weights_ms": ph["weights"], "weights_share_of_step": ph["weights"] / (ms_total / a.steps),
        }
    else:
        kernel_ms = ph["contract"]
        achieved = 2.0 * S * cd / (kernel_ms * 1e-3) * 1e-12      # flops issued on the FP64 pipe by the numerator GEMM
        masked_fp64 = wl == "idw_masked" and not int8_mask
        if masked_fp64:
            achieved *= 2.0                                       # numerator and mask accumulators both on DMMA
        roofline = {
            "bound": "tensor", "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s", "frac": achieved / dmma_peak,
            "traffic": ctx.traffic.get(wl),
            "kernel": ("ipr::contract_kernel<8,false,false,4> (FP64 DMMA.8x8x4, 2 CTAs/SM)" if wl == "idw" else
                       "ipr::contract_kernel<8,false,true,4> (numerator GEMM, FP64 DMMA; divides by the denominators the "
                       "mask GEMM left in the output buffer)" if int8_mask else
                       "ipr::contract_kernel<4,true,false,4> (numerator + mask accumulators, FP64 DMMA.8x8x4)"),
            "flops_counted": ("2*S per cell-day (W.Z0)" if not masked_fp64 else "4*S per cell-day (W.Z0 and W.M on DMMA)"),
            "kernel_ms": kernel_ms,
            "output_stream_gbs": cd * 8 / (kernel_ms * 1e-3) * 1e-9,
        }
        if wl == "idw_masked" and int8_mask:
            i8_peak = ctx.peaks.get("tcgen05_i8_n224_tops")
            ops = 2.0 * 7 * S * cd                                    # seven 8-bit digit planes of W against the byte mask
            extra["mask_gemm"] = {
                "kernel": "ipr::denom_umma_kernel (tcgen05.mma kind::i8, M128 N224 K32, s32 accumulators in TMEM, persistent)",
                "ms": ph["mask_gemm"], "tops": ops / (ph["mask_gemm"] * 1e-3) * 1e-12, "peak_tops": i8_peak,
                "frac": (ops / (ph["mask_gemm"] * 1e-3) * 1e-12 / i8_peak) if i8_peak else None,
                "peak_source": "tools/umma_i8_probe.cu on this pool's B200: 112 cycles per M128 N224 K32 MMA",
                "slice_ms": ph["slice"], "pipe": "int8 (exact)",
            }
    roofline["peak_source"] = peak_src
    roofline["frac_of_cublas_dgemm"] = roofline["achieved"] / ctx.peaks["cublas_dgemm_tflops_burst"]
    roofline["phases_ms_per_step"] = {k: v for k, v in ph.items() if k != "radius"}

    parity = {}
    if rank == 0 and not a.no_parity:
        parity = parity_sample(wl, out, cx, cy, sx, sy, obs, radii)
    del out
    torch.cuda.empty_cache()

    # ---------------- end-to-end arm (`e2e`): host buffers through the C ABI ----------------
    e2e = None
    pageable = None
    inproc = None
    if not a.no_e2e:
        bytes_out = n_stacks * C * D * 8
        with open("/proc/meminfo") as f:
            avail = [int(l.split()[1]) for l in f if l.startswith("MemAvailable")][0] * 1024
        budget = int((avail + ctx._pin_bytes) * 0.45 / world)
        n_chunks = 1
        while bytes_out / n_chunks > budget or bytes_out / n_chunks > (64 << 30):
            n_chunks *= 2
        d_chunk = (D + n_chunks - 1) // n_chunks
        hbuf_bytes = n_stacks * C * d_chunk * 8
        hptr = ctx.pinned(hbuf_bytes)  # pinned, caller-owned output
        host_out = np.ctypeslib.as_array((Ct.c_double * (hbuf_bytes // 8)).from_address(hptr))
        obs_pin_ptr = lib.ipr_host_alloc(obs_f.nbytes)
        obs_pin = np.ctypeslib.as_array((Ct.c_double * obs_f.size).from_address(obs_pin_ptr)).reshape(S, D).T
        obs_pin[...] = obs_f  # F-ordered view over pinned memory
        err = _native.errbuf()

        def call(devs, xs, ys, n, optr, t0=0, t1=D):
            ob = obs_pin[t0:t1, :]
            if t1 - t0 != D:
                ob = np.asfortranarray(ob)
            dev = (Ct.c_int * len(devs))(*devs)
            if wl == "cressman":
                rc = lib.ipr_cressman_f64(xs.ctypes.data, ys.ctypes.data, n, sx.ctypes.data, sy.ctypes.data, S,
                                          ob.ctypes.data, t1 - t0, radii.ctypes.data, len(RADII_M), dev, len(devs),
                                          optr, err, _native.ERRLEN)
            else:
                rc = lib.ipr_idw_f64(xs.ctypes.data, ys.ctypes.data, n, sx.ctypes.data, sy.ctypes.data, S,
                                     ob.ctypes.data, t1 - t0, 2.0, dev, len(devs), optr, err, _native.ERRLEN)
            _native.check(rc, err)

        def e2e_step():
            for ci in range(n_chunks):
                t0, t1 = ci * d_chunk, min(D, (ci + 1) * d_chunk)
                if t0 < t1:
                    call([local_rank], cx, cy, C, hptr, t0, t1)

        # PCIe D2H probe on this box right now (1 GiB into pinned memory): the floor of e2e
        probe = torch.empty(1 << 27, dtype=torch.float64, device="cuda")
        probe_h = torch.empty(1 << 27, dtype=torch.float64, pin_memory=True)
        d2h_gbs = 0.0
        for _ in range(3):
            torch.cuda.synchronize()
            tp = time.perf_counter()
            probe_h.copy_(probe, non_blocking=True)
            torch.cuda.synchronize()
            d2h_gbs = max(d2h_gbs, probe.numel() * 8 / (time.perf_counter() - tp) * 1e-9)
        del probe, probe_h
        for _ in range(min(a.warmup, 2)):
            e2e_step()
        ctx.barrier()
        step_s = []
        t0 = time.perf_counter()
        for _ in range(a.steps):
            ts = time.perf_counter()
            e2e_step()
            step_s.append(time.perf_counter() - ts)
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / a.steps
        ctx.barrier()
        sec = ctx.max_over_ranks(sec)
        checksum = float(np.nansum(host_out[: min(host_out.size, 1 << 20)]))
        ss = sorted(step_s)
        e2e = {
            "value": total_cells * total_dates / sec,
            "unit": UNIT,
            "h2d_bytes_per_step": int((2 * C + 2 * S + S * D) * 8),
            "d2h_bytes_per_step": int(bytes_out),
            "ms_per_step": sec * 1e3,
            "api": "ipr_cressman_f64" if wl == "cressman" else "ipr_idw_f64",
            "host_buffers": "pinned (ipr_host_alloc), caller-owned; one process per GPU, each its own cell range",
            "calls_per_step": n_chunks,
            "ms_per_step_min_rank0": ss[0] * 1e3,
            "ms_per_step_p90_rank0": ss[min(len(ss) - 1, int(0.9 * len(ss)))] * 1e3,
            "ms_steps_rank0": [round(x * 1e3, 1) for x in step_s],
            "pcie_d2h_gbs_probe": d2h_gbs,
            "pcie_floor_ms": bytes_out / (d2h_gbs * 1e9) * 1e3 if d2h_gbs > 0 else None,
            "checksum_first_1M": checksum,
        }
        # ---- the same call with a PAGEABLE output (what R's own allocVector gives the shim) ----
        if a.pageable_steps > 0 and wl == "idw" and n_chunks == 1:
            pg = np.empty(hbuf_bytes // 8, dtype=np.float64)     # never touched: the first call pays the page faults
            ps = []
            ctx.barrier()
            for _ in range(a.pageable_steps + 1):
                ts = time.perf_counter()
                call([local_rank], cx, cy, C, pg.ctypes.data)
                ps.append(time.perf_counter() - ts)
            ctx.barrier()
            warm = ctx.max_over_ranks(float(np.mean(ps[1:])))
            pageable = {
                "value": total_cells * total_dates / warm, "unit": UNIT, "ms_per_step": warm * 1e3,
                "ms_first_touch_rank0": ps[0] * 1e3, "steps": a.pageable_steps,
                "h2d_bytes_per_step": e2e["h2d_bytes_per_step"], "d2h_bytes_per_step": e2e["d2h_bytes_per_step"],
                "host_buffers": "pageable numpy array (malloc): D2H through the pinned staging ring + parallel memcpy",
            }
            del pg
        # ---- N > 1: ONE process driving every GPU through the device list of the same call ----
        if a.inproc_steps > 0 and world > 1 and not weak and wl == "idw":
            full_bytes = n_stacks * total_cells * D * 8
            ctx.barrier()      # every GPU idle; the other ranks then wait on the CPU-side group
            if rank == 0:
                hp = ctx.pinned(full_bytes)
                ips = []
                for _ in range(a.inproc_steps + 1):
                    ts = time.perf_counter()
                    call(list(range(world)), cx_all, cy_all, total_cells, hp)
                    ips.append(time.perf_counter() - ts)
                best = float(np.mean(ips[1:]))
                inproc = {
                    "value": total_cells * D / best, "unit": UNIT, "ms_per_step": best * 1e3, "steps": a.inproc_steps,
                    "ms_steps": [round(x * 1e3, 1) for x in ips],
                    "d2h_bytes_per_step": int(full_bytes),
                    "api": "ipr_idw_f64(devices = 0..%d) from ONE process: a host thread + plan per GPU, every GPU "
                           "writes its column range of the caller's pinned matrix" % (world - 1),
                }
            ctx.cpu_barrier()
            ctx.barrier()
        lib.ipr_host_free(obs_pin_ptr)

    plan.close()
    del d_obs
    torch.cuda.empty_cache()
    lib.ipr_release_cached_memory()

    rec = {
        "value": value,
        "unit": UNIT,
        "ms_per_step": ms_step,
        "steps": a.steps,
        "config": make_config(a, wl, world),
        "roofline": roofline,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    rec.update(extra)
    rec.update(parity)
    if pageable:
        rec["e2e_pageable"] = pageable
    if inproc:
        rec["e2e_inproc"] = inproc
    return rec


def run_ours(a):
    ctx = Ctx()
    wls = list(WORKLOADS) if a.workload == "all" else [a.workload]
    recs = {}
    for wl in wls:
        recs[wl] = run_workload(a, wl, ctx)
    first = recs[wls[0]]
    cpu = None
    if ctx.rank == 0 and ctx.world == 1 and not a.no_cpu_baseline:
        cpu = cpu_reference_sample(a, wls[0], 1, 1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if ctx.rank == 0:
        line = {
            "metric": METRIC,
            "value": first["value"],
            "unit": UNIT,
            "n_gpus": ctx.world,
            "steps": a.steps,
            "warmup": a.warmup,
            "ms_per_step": first["ms_per_step"],
            "higher_is_better": True,
            "scaling": a.scaling,
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": first["config"],
            "roofline": first["roofline"],
            "cpu_baseline": cpu,
            "e2e": first["e2e"],
            "gpu_launches": first["gpu_launches"],
            "clocks": first["clocks"],
            "rscript": shutil.which("Rscript"),
        }
        for k, v in first.items():
            if k not in line and k not in ("steps",):
                line[k] = v
        if len(wls) > 1:
            line["workloads"] = {wl: recs[wl] for wl in wls[1:]}
            line["gpu_launches_all_workloads"] = int(sum(r["gpu_launches"] for r in recs.values()))
        print(json.dumps(line), flush=True)
    ctx.close()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
