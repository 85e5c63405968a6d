"""Development check run on the GPU box: parity of the CUDA path vs the oracle on small cases, then
a mid-size throughput probe.  (The real tests live in tests/; this is a fast iteration aid.)"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from interpolater_b200 import engine, synthetic
from oracle import reference_port as ref


def cmp(name, got, want, z_abs=None, tol=1e-9):
    got = np.asarray(got); want = np.asarray(want)
    na_g, na_w = np.isnan(got), np.isnan(want)
    ok_na = np.array_equal(na_g, na_w)
    m = ~na_w & ~na_g
    err = np.abs(got[m] - want[m])
    scale = np.abs(want[m])
    rel = np.max(err / np.maximum(scale, 1e-300)) if m.any() else 0.0
    bad = np.sum(err > tol * np.maximum(scale, 1e-12)) if m.any() else 0
    status = "OK " if (ok_na and bad == 0) else "FAIL"
    print(f"{status} {name}: NA-match={ok_na} n={m.sum()} max_rel={rel:.3e} bad={bad} na_ref={na_w.sum()} na_got={na_g.sum()}")
    return ok_na and bad == 0


ok = True
# 1. small mixed case with planted zero-distance stations, all-NA and all-zero dates, NAs
cx, cy, sx, sy, obs = synthetic.synthetic_case(64, 64, 64, 32, na_frac=0.1, plant_zero_dist=4, all_na_dates=(3,), all_zero_dates=(7,))
got = engine.idw_grid(cx, cy, sx, sy, obs, p=2.0)
want = ref.idw_reference(cx, cy, sx, sy, obs, 2.0)
ok &= cmp("idw mixed 64x64x64x32", got, want)
assert np.all(got[:, 7] == 0.0), "zero layer must be exact"
# 2. no NA, multiple tiles, ragged sizes
cx, cy, sx, sy, obs = synthetic.synthetic_case(30, 11, 37, 700, plant_zero_dist=2)
got = engine.idw_grid(cx, cy, sx, sy, obs, p=2.0)
want = ref.idw_reference(cx, cy, sx, sy, obs, 2.0)
ok &= cmp("idw noNA 330x37x700", got, want)
# 3. NA only in some dates -> mixed tile variants
obs2 = obs.copy(); obs2[130:140, 3] = np.nan; obs2[600, :] = np.nan
got = engine.idw_grid(cx, cy, sx, sy, obs2, p=2.0)
want = ref.idw_reference(cx, cy, sx, sy, obs2, 2.0)
ok &= cmp("idw partial-NA 330x37x700", got, want)
# 4. other powers
for p in (4.0, 1.0, 1.5, 3.0):
    got = engine.idw_grid(cx, cy, sx, sy, obs2, p=p)
    want = ref.idw_reference(cx, cy, sx, sy, obs2, p)
    ok &= cmp(f"idw p={p}", got, want)
# 5. Cressman
radii = ref.cressman_radii_m([3, 10, 5, 40])
got = engine.cressman_grid(cx, cy, sx, sy, obs2, radii)
want = ref.cressman_reference(cx, cy, sx, sy, obs2, radii)
for i, r in enumerate(radii):
    ok &= cmp(f"cressman R={r/1000:g} km", got[i], want[i])
print("PARITY", "PASS" if ok else "FAIL")

# 6. throughput probe through the plan API (device-resident)
import torch
ncol, nrow, S, D = 500, 400, 2000, 3650
cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, S, D)
C = cx.shape[0]
plan = engine.Plan(0, cx, cy, sx, sy, D)
plan.set_obs(obs)
out = torch.empty((D, C), dtype=torch.float64, device="cuda")
for it in range(3):
    plan.timer_start()
    plan.idw(out.data_ptr(), 2.0)
    ms = plan.timer_stop()
    flops = 2.0 * C * S * D
    print(f"IDW C={C} S={S} D={D}: {ms:.2f} ms  {C*D/ms*1e-6:.3f} G cell-days/s  {flops/ms*1e-9:.2f} TFLOP/s ({flops/ms*1e-9/37.12*100:.1f}% of DMMA peak)")
chk = out[:5].cpu().numpy().T
want = ref.idw_reference(cx[:2000], cy[:2000], sx, sy, obs[:5], 2.0)
ok2 = cmp("idw large sample", chk[:2000], want)
radii = np.array([25e3, 50e3, 100e3, 200e3])
outc = torch.empty((4, D, C), dtype=torch.float64, device="cuda")
for it in range(2):
    plan.timer_start()
    plan.cressman(outc.data_ptr(), radii)
    ms = plan.timer_stop()
    flops = 2.0 * C * S * D * 4
    print(f"Cressman x4 radii: {ms:.2f} ms  {C*D/ms*1e-6:.3f} G cell-days/s  {flops/ms*1e-9:.2f} TFLOP/s ({flops/ms*1e-9/37.12*100:.1f}%)")
wantc = ref.cressman_reference(cx[:2000], cy[:2000], sx, sy, obs[:3], radii)
for i in range(4):
    cmp(f"cressman large sample R{i}", outc[i, :3, :2000].cpu().numpy().T, wantc[i])
# masked IDW
obs3 = obs.copy(); obs3[np.random.default_rng(5).random(obs3.shape) < 0.2] = np.nan
plan.set_obs(obs3)
for it in range(2):
    plan.timer_start()
    plan.idw(out.data_ptr(), 2.0)
    ms = plan.timer_stop()
    flops = 4.0 * C * S * D
    print(f"IDW masked: {ms:.2f} ms  {C*D/ms*1e-6:.3f} G cell-days/s  {flops/ms*1e-9:.2f} TFLOP/s ({flops/ms*1e-9/37.12*100:.1f}%)")
chk = out[:4].cpu().numpy().T
want = ref.idw_reference(cx[:2000], cy[:2000], sx, sy, obs3[:4], 2.0)
cmp("idw masked large sample", chk[:2000], want)
print("launches", plan.launch_count, "zero pairs", plan.zero_pairs)
