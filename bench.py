#!/usr/bin/env python
"""Benchmark of the IDW / Cressman hot path: interpolated cell-days per second.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload idw|idw_masked|cressman]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      # the reference algorithm on the host cores

One "step" is one pass of the hot path over the whole synthetic workload (SURVEY §8d cfg 3/4/5:
1,000,000 cells x 2,000 stations x 3,650 days).  `value` is the device-resident throughput
(inputs already in HBM, outputs left in HBM, CUDA events on the plan stream, max over ranks);
`e2e` is the same job through the host-buffer C ABI call (`ipr_idw_f64`), H2D of the inputs and
D2H of every output layer inside the timed region.  Multi-GPU: date blocks are independent, so
each rank owns its own 3,650-day block of the same grid (weak scaling, no collective on the data
path; NCCL is used only for the barrier and the max-over-ranks of the timings).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "interpolated cell-days/sec (IDW p=2, Cressman) at 1/2/4/8 B200; % of roofline"
UNIT = "cell-days/s"
RADII_M = [25e3, 50e3, 100e3, 200e3]  # cfg 4


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="idw", choices=["idw", "idw_masked", "cressman"])
    ap.add_argument("--ncol", type=int, default=1000)
    ap.add_argument("--nrow", type=int, default=1000)
    ap.add_argument("--stations", type=int, default=2000)
    ap.add_argument("--dates", type=int, default=3650)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU owns its own --dates-day block of the record (default); "
                         "strong: ONE job, the cells are split over the GPUs on 64-cell tiles")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-cells", type=int, default=10000)
    ap.add_argument("--cpu-sample-dates", type=int, default=48)
    return ap.parse_args()


def workload_name(a):
    c = a.ncol * a.nrow
    base = f"{c:,} cells x {a.stations:,} stations x {a.dates:,} days"
    if a.workload == "idw":
        return f"synthetic IDW (cfg 3): {base}, p=2, no NA"
    if a.workload == "idw_masked":
        return f"synthetic IDW (cfg 5): {base}, p=2, 20% per-date NA (masked contraction)"
    return f"synthetic Cressman (cfg 4): {base}, 4 radii (25/50/100/200 km)"


def make_inputs(a, rank):
    from interpolater_b200 import synthetic

    na = 0.2 if a.workload == "idw_masked" else 0.0
    cx, cy, sx, sy, obs = synthetic.synthetic_case(a.ncol, a.nrow, a.stations, a.dates, na_frac=na,
                                                   seed=20260101 + 1000 * rank)
    if rank > 0:  # same geometry on every rank; only the date block (obs) differs
        cx0, cy0, sx, sy, _ = synthetic.synthetic_case(a.ncol, a.nrow, a.stations, 1, seed=20260101)
    return cx, cy, sx, sy, obs


def flops_per_cell_day(a):
    """Algorithmic flops: weight generation / normalisation are not counted."""
    if a.workload == "idw":
        return 2.0 * a.stations
    if a.workload == "idw_masked":
        return 4.0 * a.stations
    return 2.0 * a.stations * len(RADII_M)


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
            time.sleep(0.05)

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
def cpu_reference_sample(a, threads, steps=1):
    """Times oracle.reference_port (dense dist/weight matrices + per-date loop, exactly the
    reference's algorithm) on `cpu_sample_cells` cells x `cpu_sample_dates` dates of the workload,
    all stations.  threads > 1 runs independent date sub-blocks in a thread pool (NumPy releases the
    GIL inside the dense passes), which is how the serial R loop would be spread over cores."""
    from concurrent.futures import ThreadPoolExecutor

    from interpolater_b200 import synthetic
    from oracle import reference_port as ref  # the timed CPU baseline (allowed use of oracle/)

    side = int(round(a.cpu_sample_cells ** 0.5))
    nd = a.cpu_sample_dates
    na = 0.2 if a.workload == "idw_masked" else 0.0
    cx, cy, sx, sy, obs = synthetic.synthetic_case(side, side, a.stations, nd, na_frac=na)
    ncell = cx.shape[0]

    radii = np.asarray(RADII_M)

    ref.LONG_DOUBLE_ROWSUMS = False  # timing leg: float64 row sums (R's C loop is at least that fast)
    n_full = a.dates

    def one_pass():
        # date-independent setup once (as the reference does), then the per-date loop; with
        # threads > 1 the dates are spread over a pool sharing the dense matrices
        t0 = time.perf_counter()
        if a.workload == "cressman":
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
    threads = os.cpu_count() or 1
    for _ in range(min(a.warmup, 1)):
        cpu_reference_sample(a, threads, 1)
    res = cpu_reference_sample(a, threads, max(1, a.steps))
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
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": res["sample"]},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    from interpolater_b200 import _native, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _native.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cx, cy, sx, sy, obs = make_inputs(a, rank if a.scaling == "weak" else 0)
    total_cells = cx.shape[0]
    total_dates = obs.shape[0] * (world if a.scaling == "weak" else 1)
    if a.scaling == "strong":
        # one job over all GPUs: contiguous cell ranges (nothing computed twice, equal shares)
        from interpolater_b200 import sharding

        b = sharding.cell_block_bounds(total_cells, world)
        cx, cy = np.ascontiguousarray(cx[b[rank]:b[rank + 1]]), np.ascontiguousarray(cy[b[rank]:b[rank + 1]])
        if cx.shape[0] == 0:
            raise SystemExit("strong scaling: fewer 64-cell tiles than GPUs")
    C, S, D = cx.shape[0], sx.shape[0], obs.shape[0]
    n_stacks = len(RADII_M) if a.workload == "cressman" else 1
    radii = np.asarray(RADII_M)

    # ---------------- device-resident arm (`value`) ----------------
    plan = engine.Plan(local_rank, cx, cy, sx, sy, D)
    obs_f = np.asfortranarray(obs)
    d_obs = torch.from_numpy(obs_f.T.copy()).cuda()  # (n_st, n_dates) row-major == column-major n_dates x n_st
    out = torch.empty((n_stacks, D, C), dtype=torch.float64, device="cuda")

    def step():
        # the whole hot path from HBM-resident raw inputs: NA split + early-out flags, contraction
        # (+ zero-distance fix-up), outputs left in HBM
        plan.reset_cache()  # weights (A operand) are regenerated inside every step
        plan.set_obs_device(d_obs.data_ptr())
        if a.workload == "cressman":
            plan.cressman(out.data_ptr(), radii)
        else:
            plan.idw(out.data_ptr(), 2.0)

    for _ in range(a.warmup):
        step()
    plan.sync()
    sampler = ClockSampler(local_rank)
    barrier()
    l0 = plan.launch_count
    sampler.start()
    plan.timer_start()
    for _ in range(a.steps):
        step()
    ms_total = plan.timer_stop()
    barrier()
    clocks = sampler.stop()
    launches = plan.launch_count - l0
    ms_step = max_over_ranks(ms_total / a.steps)
    value = total_cells * total_dates / (ms_step * 1e-3)

    # dominant kernel: the DMMA contraction.  Its launches are timed alone here (same stream, CUDA
    # events; obs preparation excluded, weights cached — for the multi-radius Cressman the cache
    # holds one radius at a time so its weight passes stay inside) to give the roofline numerator.
    plan.sync()
    ex0, de0 = plan.stage_blocks()
    plan.timer_start()
    if a.workload == "cressman":
        plan.cressman(out.data_ptr(), radii)
    else:
        plan.idw(out.data_ptr(), 2.0)
    ms_kernel = plan.timer_stop()
    ex1, de1 = plan.stage_blocks()
    # fraction of the dense (cell tile x station stage) blocks that hold any non-zero weight: 1 for
    # IDW; for Cressman the blocks wholly beyond the radius are exactly zero and are skipped
    executed_frac = (ex1 - ex0) / max(de1 - de0, 1) if de1 > de0 else 1.0
    with open(os.path.join(ROOT, "profiles", "fp64_peaks.json")) as f:
        peaks = json.load(f)
    flops = flops_per_cell_day(a) * C * D
    achieved = flops / (ms_kernel * 1e-3) * 1e-12
    roofline = {
        "bound": "tensor",
        "achieved": achieved,
        "peak": peaks["dmma_tflops"],
        "unit": "TFLOP/s",
        "frac": achieved / peaks["dmma_tflops"],
        "traffic": None,
        "kernel": {
            "idw": "ipr::contract_kernel<8,false,false,4> (FP64 DMMA.8x8x4, 2 CTAs/SM)",
            "cressman": "ipr::contract_kernel<8,false,false,2> (FP64 DMMA.8x8x4, 2 CTAs/SM, 8-station half stages)",
            "idw_masked": ("ipr::contract_kernel<8,false,true,4> (numerator, FP64 DMMA) + ipr::denom_imma_kernel (mask GEMM, IMMA)"
                           if os.environ.get("IPR_MASK_INT8", "1") != "0" else
                           "ipr::contract_kernel<4,true,false,4> (numerator + mask accumulators, FP64 DMMA.8x8x4)"),
        }[a.workload],
        "peak_source": "profiles/fp64_peaks.json: DMMA m8n8k4 probe measured on this pool's B200 "
                       "(MEASURED_PEAKS.json has no FP64 entry); cuBLAS DGEMM 8192^3 = %.1f TFLOP/s"
                       % peaks["cublas_dgemm_tflops_burst"],
        "frac_of_cublas_dgemm": achieved / peaks["cublas_dgemm_tflops_burst"],
        "algorithmic_flops_per_cell_day": flops_per_cell_day(a),
        "kernel_ms": ms_kernel,
        "output_stream_gbs": n_stacks * C * D * 8 / (ms_kernel * 1e-3) * 1e-9,
        # `achieved` counts the DENSE product as SURVEY 8(d) prescribes; what the tensor pipe really
        # executed is reported beside it (equal for IDW)
        # masked IDW: SURVEY 8(d) counts 4*S flop per cell-day (numerator + mask GEMM); the mask GEMM
        # runs error-free on the INTEGER tensor pipe (8-bit digit planes of the weights), so only
        # 2*S of them load the FP64 pipe and `frac` can exceed 1
        "mask_gemm_pipe": ("int8 (mma.sync u8, exact)" if a.workload == "idw_masked"
                           and os.environ.get("IPR_MASK_INT8", "1") != "0" else
                           ("fp64" if a.workload == "idw_masked" else None)),
        "executed_block_frac": executed_frac,
        "executed_tflops": achieved * executed_frac,
        "executed_frac_of_peak": achieved * executed_frac / peaks["dmma_tflops"],
    }
    traffic_file = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                roofline["traffic"] = json.load(f).get(a.workload)
        except Exception:
            pass
    # spot-check of the timed output against nothing-but-itself invariants (finite, non-negative)
    chk = out[0, :2, :1000]
    assert bool(torch.isfinite(chk).all()) or a.workload != "idw", "timed run produced non-finite output"
    del out
    torch.cuda.empty_cache()

    # ---------------- end-to-end arm (`e2e`): host buffers through the C ABI ----------------
    e2e = None
    if not a.no_e2e:
        bytes_out = n_stacks * C * D * 8
        with open("/proc/meminfo") as f:
            avail = [int(l.split()[1]) for l in f if l.startswith("MemAvailable")][0] * 1024
        budget = int(avail * 0.45 / world)
        n_chunks = 1
        while bytes_out / n_chunks > budget or bytes_out / n_chunks > (64 << 30):
            n_chunks *= 2
        d_chunk = (D + n_chunks - 1) // n_chunks
        hbuf_bytes = n_stacks * C * d_chunk * 8
        hptr = lib.ipr_host_alloc(hbuf_bytes)  # pinned, caller-owned output
        if not hptr:
            raise SystemExit("pinned host allocation failed")
        import ctypes as Ct

        host_out = np.ctypeslib.as_array((Ct.c_double * (hbuf_bytes // 8)).from_address(hptr))
        obs_pin_ptr = lib.ipr_host_alloc(obs_f.nbytes)
        obs_pin = np.ctypeslib.as_array((Ct.c_double * obs_f.size).from_address(obs_pin_ptr)).reshape(S, D).T
        obs_pin[...] = obs_f  # F-ordered view over pinned memory
        dev = (Ct.c_int * 1)(local_rank)
        err = _native.errbuf()

        def e2e_step():
            for ci in range(n_chunks):
                t0, t1 = ci * d_chunk, min(D, (ci + 1) * d_chunk)
                if t0 >= t1:
                    break
                ob = obs_pin[t0:t1, :]
                if n_chunks > 1:
                    ob = np.asfortranarray(ob)
                if a.workload == "cressman":
                    rc = lib.ipr_cressman_f64(cx.ctypes.data, cy.ctypes.data, C, sx.ctypes.data, sy.ctypes.data, S,
                                              ob.ctypes.data, t1 - t0, radii.ctypes.data, len(RADII_M), dev, 1,
                                              hptr, err, _native.ERRLEN)
                else:
                    rc = lib.ipr_idw_f64(cx.ctypes.data, cy.ctypes.data, C, sx.ctypes.data, sy.ctypes.data, S,
                                         ob.ctypes.data, t1 - t0, 2.0, dev, 1, hptr, err, _native.ERRLEN)
                _native.check(rc, err)

        # PCIe D2H probe on this box right now (1 GiB from the same pinned buffer): the floor of e2e
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
        barrier()
        step_s = []
        t0 = time.perf_counter()
        for _ in range(a.steps):
            ts = time.perf_counter()
            e2e_step()
            step_s.append(time.perf_counter() - ts)
        torch.cuda.synchronize()
        sec = (time.perf_counter() - t0) / a.steps
        barrier()
        sec = max_over_ranks(sec)
        checksum = float(np.nansum(host_out[: min(host_out.size, 1 << 20)]))
        e2e = {
            "value": total_cells * total_dates / sec,
            "unit": UNIT,
            "h2d_bytes_per_step": int((2 * C + 2 * S + S * D) * 8),
            "d2h_bytes_per_step": int(bytes_out),
            "ms_per_step": sec * 1e3,
            "api": "ipr_cressman_f64" if a.workload == "cressman" else "ipr_idw_f64",
            "host_buffers": "pinned (ipr_host_alloc), caller-owned",
            "calls_per_step": n_chunks,
            "ms_per_step_min_rank0": min(step_s) * 1e3,
            "ms_steps_rank0": [round(x * 1e3, 1) for x in step_s],
            "pcie_d2h_gbs_probe": d2h_gbs,
            "pcie_floor_ms": bytes_out / (d2h_gbs * 1e9) * 1e3 if d2h_gbs > 0 else None,
            "checksum_first_1M": checksum,
        }
        lib.ipr_host_free(hptr)
        lib.ipr_host_free(obs_pin_ptr)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu = cpu_reference_sample(a, 1, 1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": world,
            "steps": a.steps,
            "warmup": a.warmup,
            "ms_per_step": ms_step,
            "higher_is_better": True,
            "scaling": a.scaling,
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": workload_name(a),
                "n_cells": int(total_cells), "n_cells_rank0": C, "n_stations": S, "n_dates_total": int(total_dates),
                "n_dates_rank0": D, "n_stacks": n_stacks,
                "sharding": ("one independent %d-day date block per GPU" % D if a.scaling == "weak" else
                             "%d cells split over %d GPUs on 64-cell tiles, every GPU all %d days"
                             % (total_cells, world, D)) + ", no data-path collective",
                "l2": "outputs (%.1f GB/step) >> 126 MB L2: every step streams through and flushes L2"
                      % (n_stacks * C * D * 8e-9),
            },
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
