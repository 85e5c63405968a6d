"""Host-side mirror of the reference's shared validation module (KEEP-in-R code):
station split (.select_data, R/Intr_Connections_module.R:5-52), point-to-pixel validation
(.validate, :215-287), goodness-of-fit and categorical metrics (metrics, :55-212;
R/Intr_gof.R:7-143).  In an R deployment these stay in R untouched; this NumPy/pandas twin exists
because R is absent here and the drop-in `IDW()` / `Cressman()` must still return the
validation table.  Not on the GPU hot path.
"""
from __future__ import annotations

import math
import sys

import numpy as np
import pandas as pd

from .spatial import r_round


def message(*parts):
    """R's message(): diagnostics on stderr."""
    print(*parts, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------------------
# R's default RNG (Mersenne-Twister, set.seed scrambling, rejection sampling) — needed only to
# reproduce `set.seed(123); sample(columns, floor(training * n))` (ICM:17-22)
# --------------------------------------------------------------------------------------
class RMersenneTwister:
    N, M = 624, 397

    def __init__(self, seed: int):
        s = seed & 0xFFFFFFFF
        for _ in range(50):  # initial scrambling (RNG.c Randomize)
            s = (69069 * s + 1) & 0xFFFFFFFF
        st = []
        for _ in range(self.N + 1):
            s = (69069 * s + 1) & 0xFFFFFFFF
            st.append(s)
        self.mt = st[1:]
        self.mti = self.N  # FixupSeeds: dummy[0] = 624 -> regenerate on first draw

    def _genrand_int32(self):
        N, M = self.N, self.M
        mt = self.mt
        if self.mti >= N:
            for kk in range(N):
                y = (mt[kk] & 0x80000000) | (mt[(kk + 1) % N] & 0x7FFFFFFF)
                v = mt[(kk + M) % N] ^ (y >> 1)
                if y & 1:
                    v ^= 0x9908B0DF
                mt[kk] = v
            self.mti = 0
        y = mt[self.mti]
        self.mti += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def unif_rand(self) -> float:
        v = self._genrand_int32() * 2.3283064365386963e-10
        if v <= 0.0:
            return 0.5 * 2.328306437080797e-10
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * 2.328306437080797e-10
        return v

    def _rbits(self, bits: int) -> int:
        v = 0
        n = 0
        while n <= bits:
            v1 = int(math.floor(self.unif_rand() * 65536))
            v = 65536 * v + v1
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        return v

    def unif_index(self, dn: int) -> int:
        """R_unif_index with sample.kind = "Rejection" (R >= 3.6.0)."""
        if dn <= 0:
            return 0
        bits = int(math.ceil(math.log2(dn)))
        while True:
            dv = self._rbits(bits)
            if dv < dn:
                return dv

    def sample_int(self, n: int, k: int):
        """sample.int(n, k) without replacement -> 1-based indices."""
        x = list(range(n))
        out = []
        for _ in range(k):
            j = self.unif_index(n)
            out.append(x[j] + 1)
            n -= 1
            x[j] = x[n]
        return out


# --------------------------------------------------------------------------------------
# .select_data  (ICM:5-52)
# --------------------------------------------------------------------------------------
def select_data(BD_Obs: pd.DataFrame, BD_Coord: pd.DataFrame, training, stat_validation=None):
    cols = [c for c in BD_Obs.columns if c != "Date"]
    if stat_validation is None:
        if not (0 <= training <= 1):
            raise ValueError("The training parameter must be between 0 and 1.")
        message("Random validation mode.", _r_num(training * 100), "% of the stations will be used for training and",
                _r_num(100 - training * 100), "% for validation.")
        rng = RMersenneTwister(123)                                     # ICM:17
        k = int(math.floor(training * len(cols)))                       # ICM:21
        idx = rng.sample_int(len(cols), k)                              # ICM:19-22
        train_columns = [cols[i - 1] for i in idx]
    else:
        sv = [stat_validation] if isinstance(stat_validation, str) else list(stat_validation)
        message("Manual validation mode. Stations:", ", ".join(sv), "will be used for validation.")
        train_columns = [c for c in cols if c not in set(sv)]           # ICM:29
    train_data = BD_Obs[["Date"] + train_columns]                       # ICM:33
    test_cols = [c for c in BD_Obs.columns if c not in set(train_columns)]  # ICM:36-39 (keeps Date)
    test_data = BD_Obs[test_cols]
    train_cords = BD_Coord[BD_Coord["Cod"].isin(train_columns)]         # ICM:42
    test_cords = BD_Coord[BD_Coord["Cod"].isin([c for c in BD_Obs.columns if c not in set(train_columns)])]  # ICM:45
    return dict(train_data=train_data, test_data=test_data, train_cords=train_cords, test_cords=test_cords)


def _r_num(v):
    """paste() formatting of a number (up to 15 significant digits, no trailing zeros)."""
    return "%.15g" % v


# --------------------------------------------------------------------------------------
# goodness of fit  (R/Intr_gof.R)
# --------------------------------------------------------------------------------------
def _complete(sim, obs):
    sim = np.asarray(sim, dtype=np.float64)
    obs = np.asarray(obs, dtype=np.float64)
    ok = ~(np.isnan(sim) | np.isnan(obs))
    return sim[ok], obs[ok]


def _cor(a, b):
    da, db = a - a.mean(), b - b.mean()
    den = math.sqrt(float(da @ da) * float(db @ db))
    return float(da @ db) / den if den != 0 else float("nan")


def _rank(v):
    from scipy.stats import rankdata

    return rankdata(v, method="average")


def gof_metrics(sim, obs):
    """Unrounded MAE, MSE, RMSE, PBIAS %, NSE, R2, rSpearman, rPearson, KGE (Intr_gof.R:7-143)."""
    s, o = _complete(sim, obs)
    n = o.size
    nan = float("nan")
    r = dict.fromkeys(["MAE", "MSE", "RMSE", "PBIAS %", "NSE", "R2", "rSpearman", "rPearson", "KGE"], nan)
    if n == 0:
        return r
    d = s - o
    r["MAE"] = float(np.abs(d).mean())
    r["MSE"] = float((d * d).mean())
    r["RMSE"] = math.sqrt(r["MSE"])
    so = float(o.sum())
    if so != 0:
        r["PBIAS %"] = 100.0 * float(d.sum()) / so
    ss_tot = float(((o - o.mean()) ** 2).sum())
    if ss_tot != 0:
        r["NSE"] = 1.0 - float((d * d).sum()) / ss_tot
    if n >= 2:
        if ss_tot != 0:
            r["R2"] = 1.0 - float((d * d).sum()) / ss_tot
        r["rSpearman"] = _cor(_rank(o), _rank(s))
        r["rPearson"] = _cor(o, s)
        mo = float(o.mean())
        sdo = float(o.std(ddof=1))
        if mo != 0 and sdo != 0 and not math.isnan(r["rPearson"]):
            alpha = float(s.std(ddof=1)) / sdo
            beta = float(s.mean()) / mo
            r["KGE"] = 1.0 - math.sqrt((r["rPearson"] - 1) ** 2 + (alpha - 1) ** 2 + (beta - 1) ** 2)
    return r


GOF_COLUMNS = ["MAE", "MSE", "RMSE", "PBIAS %", "NSE", "R2", "rSpearman", "rPearson", "KGE"]


def _div(a, b):
    if b == 0:
        if a == 0 or (isinstance(a, float) and math.isnan(a)):
            return float("nan")
        return math.copysign(float("inf"), a)
    return a / b


def _threshold_categories(rain_threshold):
    names = list(rain_threshold.keys())
    mins = []
    for nm in names:
        v = np.asarray(rain_threshold[nm], dtype=np.float64).reshape(-1)
        mins.append(float(v[0]))
    order = sorted(range(len(names)), key=lambda i: mins[i])
    return [mins[i] for i in order] + [float("inf")], [names[i] for i in order]


def categorical_metrics(sim, obs, rain_threshold) -> pd.DataFrame:
    """ICM:73-208, including the HSS / ETS operator-precedence of the reference (ICM:130-152)."""
    thr, cats = _threshold_categories(rain_threshold)
    edges = np.asarray(thr)

    def cut(v):
        v = np.asarray(v, dtype=np.float64)
        idx = np.searchsorted(edges, v, side="right") - 1   # [thr[i], thr[i+1])
        idx = np.where(np.isnan(v) | (idx < 0) | (idx >= len(cats)), -1, idx)
        return idx

    oc, ec = cut(obs), cut(sim)
    valid = (oc >= 0) & (ec >= 0)
    rows = []
    nan = float("nan")
    for ci, cat in enumerate(cats):
        h = int(np.sum(valid & (oc == ci) & (ec == ci)))
        m = int(np.sum(valid & (oc == ci) & (ec != ci)))
        fa = int(np.sum(valid & (ec == ci) & (oc != ci)))
        cn = int(np.sum(valid & (oc != ci) & (ec != ci)))
        POD = h / (h + m) if h + m > 0 else nan
        SR = 1 - fa / (h + fa) if h + fa > 0 else nan
        CSI = h / (h + m + fa) if h + m + fa > 0 else nan
        HB = (h + fa) / (h + m) if h + m > 0 else nan
        FAR = fa / (h + fa) if h + fa > 0 else nan
        HK = POD - _div(fa, fa + cn)
        cond = (h + m) * (m + cn) + (h + fa) * (fa + cn)
        HSS = (_div(2 * (h * cn - m * fa), h + m) * (m + cn) + (h + fa) * (fa + cn)) if cond != 0 else nan
        a_random = _div((h + fa) * (h + m), h + m + fa + cn)
        if math.isnan(a_random) or (h + m + fa - a_random) == 0:
            ETS = nan
        else:
            ETS = h - _div(a_random, h) + m + fa - a_random
        rows.append(dict(Category=cat, POD=POD, SR=SR, CSI=CSI, HB=HB, FAR=FAR, HK=HK, HSS=HSS, ETS=ETS))
    return pd.DataFrame(rows)


def metrics(sim, obs, Rain_threshold=None):
    """metrics() (ICM:55-212): gof row rounded to 2 digits, plus categorical table if requested."""
    g = gof_metrics(sim, obs)
    gof = pd.DataFrame([{k: float(r_round(g[k], 2)) for k in GOF_COLUMNS}])
    if Rain_threshold is not None:
        return dict(gof=gof, categorical_metrics=categorical_metrics(sim, obs, Rain_threshold))
    return gof


# --------------------------------------------------------------------------------------
# .validate  (ICM:215-287)
# --------------------------------------------------------------------------------------
def validate(test_cords: pd.DataFrame, test_data: pd.DataFrame, Ensamble, Rain_threshold=None, sim_pts=None,
             layer_names=None):
    """Samples the stack at the test stations (cell containing the point; NA outside the raster),
    pairs Obs/Sim per station by date and evaluates metrics().  `Ensamble` is a spatial.SpatRaster
    whose layers are the sorted unique dates of the run; alternatively the already-extracted
    (n_test, n_layers) matrix `sim_pts` with its `layer_names` (the engine evaluated just the cells
    containing the test stations, so the full stack never has to be sampled on the host)."""
    message("Validation process in progress. Please wait.")
    test_cols = [c for c in test_data.columns if c != "Date"]
    # ICM:231: ID := as.numeric(Cod) — factor levels follow test_data's column order, and the
    # i-th test coordinate row is matched to the i-th level (ICM:217, :236)
    if sim_pts is None:
        sim_pts = Ensamble.extract(test_cords["X"].to_numpy(dtype=np.float64),
                                   test_cords["Y"].to_numpy(dtype=np.float64))
        layer_names = Ensamble.names
    layer_dates = list(layer_names)
    date_key = [str(d)[:10] for d in pd.to_datetime(test_data["Date"])] if not _is_str_dates(test_data["Date"]) \
        else [str(d) for d in test_data["Date"]]
    pos = {d: i for i, d in enumerate(layer_dates)}
    res = {}
    for j, col in enumerate(test_cols):
        if j >= sim_pts.shape[0]:
            break
        obs = test_data[col].to_numpy(dtype=np.float64)
        sim = np.array([sim_pts[j, pos[d]] if d in pos else np.nan for d in date_key])
        res[col] = metrics(sim, obs, Rain_threshold)
    if Rain_threshold is not None:
        gofs, cats = [], []
        for k, v in res.items():
            g = v["gof"].copy()
            g["ID"] = k
            c = v["categorical_metrics"].copy()
            c["ID"] = k
            gofs.append(g)
            cats.append(c)
        return dict(gof=pd.concat(gofs, ignore_index=True), categorical_metrics=pd.concat(cats, ignore_index=True))
    rows = []
    for k, v in res.items():
        g = v.copy()
        g.insert(0, "ID", k)
        rows.append(g)
    return pd.concat(rows, ignore_index=True)


def _is_str_dates(col):
    return len(col) > 0 and isinstance(col.iloc[0], str)
