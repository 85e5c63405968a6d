"""Observation ingestion: the caller-side producer of the `BD_Obs` table the hot path consumes.

    create_data(file_path, Start_date, End_Date, ncores=None, max_na=None)     R/create_data.R:43-94

Host-side mirror.  One CSV per station (`<station id>.csv`, two columns:
date, value) becomes one wide table — dates as rows, stations as columns — which `IDW()` /
`Cressman()` take as `BD_Obs`.  `obs_matrix()` hands the same table to the engine in the layout the
C ABI wants (column-major n_dates x n_stations, NaN = missing) without going through the
reference's melt/long-table detour (R/IDW.R:213-229).
"""
from __future__ import annotations

import os
import re
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pandas as pd

__all__ = ["create_data", "obs_matrix"]


def _station_files(folder):
    """list.files(path, pattern = "*.csv", full.names = TRUE) (R/create_data.R:50-54): the pattern
    is a regular expression, so it keeps every name that contains "csv" after at least one
    character; names come back sorted.  The station id is the base name minus a trailing ".csv"
    (R/create_data.R:55)."""
    names = sorted(n for n in os.listdir(folder) if re.search(r".csv", n) and os.path.isfile(os.path.join(folder, n)))
    return [(re.sub(r"\.csv$", "", n), os.path.join(folder, n)) for n in names]


def _read_station(item):
    """fread + setnames(c("Date", id)) (R/create_data.R:57-62): exactly two columns or an error."""
    name, path = item
    est = pd.read_csv(path)
    if est.shape[1] != 2:
        raise ValueError(f"Can't assign 2 names to a {est.shape[1]}-column data.table ({os.path.basename(path)})")
    est.columns = ["Date", name]
    est["Date"] = pd.to_datetime(est["Date"]).dt.normalize()
    est[name] = pd.to_numeric(est[name], errors="coerce").astype(np.float64)
    return est


def create_data(file_path, Start_date, End_Date, ncores=None, max_na=None):
    """Merge the station files under `file_path` into one `BD_Obs` table.

    Rows are the union of the dates found in [Start_date, End_Date] (both inclusive,
    R/create_data.R:72-74), ascending (`merge(by = "Date", all = TRUE)` sorts on the key, :76-79);
    a station without a record on a date gets NA.  With `max_na` (percent, 0-100) the result is
    `dict(data=..., Na_stations=...)`; stations are kept when their share of missing values is
    strictly below `max_na` (`na_percentages < max.na`, :89 — the reference's documentation says
    "less than or equal", its code does not).
    """
    files = _station_files(str(file_path))
    if not files:
        raise ValueError("No CSV files found in " + str(file_path))
    if ncores is None:
        tables = [_read_station(f) for f in files]
    else:
        with ThreadPoolExecutor(max_workers=max(1, int(ncores))) as pool:   # future_lapply, :66-70
            tables = list(pool.map(_read_station, files))
    lo, hi = pd.Timestamp(Start_date), pd.Timestamp(End_Date)
    tables = [t[(t["Date"] >= lo) & (t["Date"] <= hi)] for t in tables]
    data_base = tables[0]
    for t in tables[1:]:
        data_base = data_base.merge(t, on="Date", how="outer")
    data_base = data_base.sort_values("Date", kind="stable").reset_index(drop=True)
    if max_na is None:
        return data_base
    stations = [c for c in data_base.columns if c != "Date"]
    n = len(data_base)
    pct = {c: (100.0 * float(data_base[c].isna().sum()) / n if n else float("nan")) for c in stations}
    valid = [c for c in stations if pct[c] < max_na]
    return dict(data=data_base[["Date"] + valid], Na_stations=pd.DataFrame([pct], columns=stations))


def obs_matrix(BD_Obs, stations=None):
    """(dates, station ids, Z) with Z Fortran-ordered n_dates x n_stations float64 — each station's
    series contiguous, the layout `ipr_idw_f64` / `ipr_plan_set_obs` read (include/interpolater_b200.h)."""
    cols = [c for c in BD_Obs.columns[1:]] if stations is None else list(stations)
    z = np.asfortranarray(BD_Obs[cols].to_numpy(dtype=np.float64))
    return BD_Obs.iloc[:, 0].to_numpy(), cols, z
