"""Drop-in mirrors of the reference's exported functions for the accelerated path:

    IDW(BD_Obs, BD_Coord, shapefile, grid_resolution, p=2, Mask=True, n_round=None, training=1,
        stat_validation=None, Rain_threshold=None, save_model=False, name_save=None)      R/IDW.R:120-133
    Cressman(BD_Obs, BD_Coord, shapefile, grid_resolution, search_radius, training=1,
             n_round=None, stat_validation=None, Rain_threshold=None, save_model=False)   R/Cressman.R:137-148

Same argument names, meaning, error strings and messages, same return shapes (a raster stack, or
`dict(Ensamble=..., Validation=...)` when validating).  Everything outside the compute core is the
reference's own orchestration restated on pandas/NumPy objects; the compute core
(R/IDW.R:249-355, R/Cressman.R:268-345) is ONE call into the CUDA engine through the C ABI
(`engine.idw_grid` / `engine.cressman_grid`).  There is no CPU path.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import engine
from .spatial import GridGeometry, SpatRaster, SpatVector
from .validation import message, select_data, validate


def _as_frames(BD_Obs, BD_Coord):
    """names(BD_Coord)[1:3] = c("Cod","X","Y"); names(BD_Obs)[1] = "Date" (R/IDW.R:172-173)."""
    obs = BD_Obs.copy()
    crd = BD_Coord.copy()
    ccols = list(crd.columns)
    ccols[:3] = ["Cod", "X", "Y"]
    crd.columns = ccols
    ocols = list(obs.columns)
    ocols[0] = "Date"
    obs.columns = ocols
    return obs, crd


def _date_layers(train_data):
    """Dates_extracted <- sort(unique(Date)); the row used for a date is its first occurrence
    (`match()` in idw_predict, R/IDW.R:289).  Returns (layer names, row index per layer)."""
    dates = train_data["Date"]
    if len(dates) and not isinstance(dates.iloc[0], str):
        keys = pd.to_datetime(dates)
        names = [str(d)[:10] for d in keys]
    else:
        keys = dates
        names = [str(d) for d in dates]
    order = np.argsort(np.asarray(keys), kind="stable")
    seen, rows, labels = set(), [], []
    for i in order:
        if names[i] in seen:
            continue
        seen.add(names[i])
        rows.append(int(i))
        labels.append(names[i])
    return labels, np.asarray(rows, dtype=np.int64)


def _station_table(train_data, train_cords, sort_by_cod):
    """Points_Train: the training stations with coordinates, in the order the reference feeds them
    to the distance matrix — Cod-sorted for IDW (keyed join, R/IDW.R:222-245), BD_Obs column order
    for Cressman (setorder(ID), R/Cressman.R:258)."""
    cols = [c for c in train_data.columns if c != "Date"]
    if sort_by_cod:
        cols = sorted(cols)  # data.table keys sort in C locale
    crd = train_cords.drop_duplicates("Cod").set_index("Cod")
    missing = [c for c in cols if c not in crd.index]
    if missing:
        raise ValueError("Stations without coordinates in BD_Coord: " + ", ".join(missing))
    sx = crd.loc[cols, "X"].to_numpy(dtype=np.float64)
    sy = crd.loc[cols, "Y"].to_numpy(dtype=np.float64)
    return cols, sx, sy


def IDW(BD_Obs, BD_Coord, shapefile, grid_resolution, p=2, Mask=True, n_round=None, training=1,
        stat_validation=None, Rain_threshold=None, save_model=False, name_save=None, devices=None):
    # ---- Check input data (R/IDW.R:137-173) ----
    if not isinstance(shapefile, SpatVector):
        raise TypeError("shapefile must be a 'SpatVector' object.")
    if shapefile.crs is None:
        raise ValueError("shapefile CRS is undefined.")
    if shapefile.is_lonlat():
        raise ValueError("shapefile must use a projected CRS (meters). Reproject before calling IDW.")
    if not isinstance(BD_Obs, pd.DataFrame):
        raise TypeError("BD_Obs must be a 'data.table' or a 'data.frame'.")
    if not isinstance(BD_Coord, pd.DataFrame):
        raise TypeError("BD_Coord must be a 'data.table' or a 'data.frame'.")
    cod_col = BD_Coord.columns[0] if "Cod" not in BD_Coord.columns else "Cod"
    obs_names = [c for c in BD_Obs.columns if c != "Date"]
    if not all(c in obs_names for c in BD_Coord[cod_col]):
        raise ValueError("The names of the coordinates do not appear in the observed data.")
    if not isinstance(grid_resolution, (int, float, np.integer, np.floating)) or isinstance(grid_resolution, bool):
        raise TypeError("'grid_resolution' must be a single numeric value in km.")
    BD_Obs, BD_Coord = _as_frames(BD_Obs, BD_Coord)

    # ---- Verify if validation is to be done (R/IDW.R:177-194) ----
    validating = training != 1 or stat_validation is not None
    if validating:
        data_val = select_data(BD_Obs, BD_Coord, training=training, stat_validation=stat_validation)
        train_data, train_cords = data_val["train_data"], data_val["train_cords"]
    else:
        message("Training not specified; using all data for training.")
        train_data, train_cords = BD_Obs, BD_Coord

    # ---- Interpolation zone (R/IDW.R:198-209) ----
    geom = GridGeometry.from_extent(shapefile.ext(), float(grid_resolution) * 1000.0, shapefile.crs)
    cell_x, cell_y = geom.cell_centres()

    # ---- Data training (R/IDW.R:213-245) ----
    cols, sx, sy = _station_table(train_data, train_cords, sort_by_cod=True)
    labels, rows = _date_layers(train_data)
    obs = train_data[cols].to_numpy(dtype=np.float64)[rows, :]

    # ---- IDW algorithm: the accelerated core (replaces R/IDW.R:249-355) ----
    # Every cell is independent, so the steps that follow the hot loop in the reference are folded
    # into the engine calls instead of being host passes over the whole stack:
    #   * validation (R/IDW.R:364-372, on the rounded UN-masked stack) needs only the cells that
    #     contain the test stations -> one tiny engine call on those cells;
    #   * rounding (R/IDW.R:356-358) and crop + mask (R/IDW.R:376-381) -> the main call computes just
    #     the cropped window, rounds and masks on the device.
    message("Analysis in progress. Please wait...")
    nr = None if n_round is None else int(n_round)
    if validating:
        tc = data_val["test_cords"]
        cells = geom.cell_from_xy(tc["X"].to_numpy(dtype=np.float64), tc["Y"].to_numpy(dtype=np.float64))
        sim_pts = np.full((len(cells), len(labels)), np.nan)
        inside = cells >= 0
        if inside.any():
            sim_pts[inside] = engine.idw_grid(cell_x[cells[inside]], cell_y[cells[inside]], sx, sy, obs, p=float(p),
                                              devices=devices, n_round=nr, round_mode=engine.ROUND_R)
        final_results = validate(tc, data_val["test_data"], None, Rain_threshold, sim_pts=sim_pts, layer_names=labels)
    if isinstance(Mask, (bool, np.bool_)) and bool(Mask):   # isTRUE(Mask), R/IDW.R:376
        sub, sub_cells = geom.crop_window(shapefile)
        keep = sub.polygon_cover(shapefile)
        values = engine.idw_grid(cell_x[sub_cells], cell_y[sub_cells], sx, sy, obs, p=float(p), devices=devices,
                                 n_round=nr, round_mode=engine.ROUND_R, cell_keep=keep)
        Ensamble = SpatRaster(sub, values, labels)
    else:
        values = engine.idw_grid(cell_x, cell_y, sx, sy, obs, p=float(p), devices=devices, n_round=nr,
                                 round_mode=engine.ROUND_R)
        Ensamble = SpatRaster(geom, values, labels)                   # R/IDW.R:355, :359
    # ---- save (R/IDW.R:385-395) ----
    if save_model is True:
        if name_save is None:
            name_save = "Model_IDW"
        Ensamble.write_cdf(f"{name_save}.nc")
        message("Model saved successfully")
    if validating:
        return dict(Ensamble=Ensamble, Validation=final_results)
    return Ensamble


def Cressman(BD_Obs, BD_Coord, shapefile, grid_resolution, search_radius, training=1, n_round=None,
             stat_validation=None, Rain_threshold=None, save_model=False, devices=None):
    # ---- Check input data (R/Cressman.R:153-193) ----
    if not isinstance(shapefile, SpatVector):
        raise TypeError("shapefile must be a 'SpatVector' with a defined CRS.")
    if shapefile.crs is None:
        raise ValueError("shapefile CRS is undefined.")
    if not isinstance(BD_Obs, pd.DataFrame):
        raise TypeError("BD_Obs must be a 'data.table' or a 'data.frame'.")
    if not isinstance(BD_Coord, pd.DataFrame):
        raise TypeError("BD_Coord must be a 'data.table' or a 'data.frame'.")
    if not isinstance(grid_resolution, (int, float, np.integer, np.floating)) or isinstance(grid_resolution, bool):
        raise TypeError("'grid_resolution' must be a single numeric value (km).")
    try:
        radii = np.asarray(search_radius, dtype=np.float64).reshape(-1)
    except (TypeError, ValueError):
        radii = np.empty(0)
    if radii.size < 1 or isinstance(search_radius, (str, bytes)):
        raise TypeError("'search_radius' must be a numeric vector (km).")
    if shapefile.is_lonlat():
        # documented deviation: the reference lets terra compute geodesic distances here; the
        # engine is planar-only
        raise ValueError("shapefile must use a projected CRS (meters). Reproject before calling Cressman.")
    BD_Obs, BD_Coord = _as_frames(BD_Obs, BD_Coord)
    names_col = [c for c in BD_Obs.columns if c != "Date"]
    if not all(c in names_col for c in BD_Coord["Cod"]):
        faltantes = [c for c in pd.unique(BD_Coord["Cod"]) if c not in names_col]
        raise ValueError("The following codes from BD_Coord do not appear in BD_Obs: " + ", ".join(map(str, faltantes)))

    validating = training != 1 or stat_validation is not None
    if validating:
        data_val = select_data(BD_Obs, BD_Coord, training=training, stat_validation=stat_validation)
        train_data, train_cords = data_val["train_data"], data_val["train_cords"]
    else:
        message("Training not specified; using all data for training.")
        train_data, train_cords = BD_Obs, BD_Coord

    # ---- Zone of interpolation (R/Cressman.R:216-227) ----
    radii_m = np.unique(radii * 1000.0)                                # sort(unique()), :217-218
    geom = GridGeometry.from_extent(shapefile.ext(), float(grid_resolution) * 1000.0, shapefile.crs)
    cell_x, cell_y = geom.cell_centres()

    # ---- Data training (R/Cressman.R:231-263) ----
    cols, sx, sy = _station_table(train_data, train_cords, sort_by_cod=False)
    labels, rows = _date_layers(train_data)
    obs = train_data[cols].to_numpy(dtype=np.float64)[rows, :]

    # ---- Cressman core (replaces R/Cressman.R:268-345) ----
    message("Analysis in progress. Please wait...")
    nr = None if n_round is None else int(n_round)
    names = ["%g km" % (r / 1000.0) for r in radii_m]                  # :357
    if validating:                                                     # :361-372, on the test-station cells only
        tc = data_val["test_cords"]
        cells = geom.cell_from_xy(tc["X"].to_numpy(dtype=np.float64), tc["Y"].to_numpy(dtype=np.float64))
        inside = cells >= 0
        sims = np.full((len(radii_m), len(cells), len(labels)), np.nan)
        if inside.any():
            sims[:, inside, :] = engine.cressman_grid(cell_x[cells[inside]], cell_y[cells[inside]], sx, sy, obs,
                                                      radii_m, devices=devices, n_round=nr,
                                                      round_mode=engine.ROUND_TERRA)
        final_results = {nm: validate(tc, data_val["test_data"], None, Rain_threshold, sim_pts=sims[i],
                                      layer_names=labels) for i, nm in enumerate(names)}
    stacks = engine.cressman_grid(cell_x, cell_y, sx, sy, obs, radii_m, devices=devices, n_round=nr,
                                  round_mode=engine.ROUND_TERRA)      # rounding :353-355 on the device
    Ensamble = {nm: SpatRaster(geom, stacks[i], labels) for i, nm in enumerate(names)}

    if save_model is True:                                             # :377-387
        message("Model saved successfully")
        for nm in names:
            Ensamble[nm].write_cdf(f"Radius_{nm}.nc")
    if validating:
        return dict(Ensamble=Ensamble, Validation=final_results)
    return Ensamble
