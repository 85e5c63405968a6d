"""Generates the committed golden fixtures (run in the build container, where /root/reference is
mounted; the GPU box never reads /root/reference):

  packaged_inputs.npz   the reference's packaged fixture decoded from data/BD_Obs.rda,
                        data/BD_Coord.rda (RDX3, oracle/rdata.py) and inst/extdata/study_area.shp
  packaged_golden.npz   oracle outputs for BASELINE.json configs[0] / configs[1]:
                        IDW(grid_resolution=5, p=2, stat_validation="M001") and
                        Cressman(search_radius=c(50,20,10), stat_validation="M001")
  synthetic_golden.npz  oracle outputs for small seeded synthetic cases (zero-distance stations,
                        all-NA / all-zero dates, NA-heavy dates, radii with empty neighbourhoods)

PARITY UNPINNED: the reference ships no numeric test vectors and R cannot run here, so these
goldens come from the oracle restatement (oracle/reference_port.py), not from the R package.

    python tests/golden/make_golden.py
"""
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from interpolater_b200 import synthetic  # noqa: E402
from oracle import rdata  # noqa: E402
from oracle import reference_port as ref  # noqa: E402

REF = "/root/reference"


def read_shp(path):
    buf = open(path, "rb").read()
    xmin, ymin, xmax, ymax = struct.unpack("<4d", buf[36:68])
    off, rings = 100, []
    while off + 8 <= len(buf):
        clen = struct.unpack(">i", buf[off + 4:off + 8])[0] * 2
        rec = buf[off + 8:off + 8 + clen]
        off += 8 + clen
        nparts, npoints = struct.unpack("<2i", rec[36:44])
        parts = list(struct.unpack(f"<{nparts}i", rec[44:44 + 4 * nparts])) + [npoints]
        pts = np.frombuffer(rec, dtype="<f8", count=2 * npoints, offset=44 + 4 * nparts).reshape(npoints, 2)
        for i in range(nparts):
            rings.append(pts[parts[i]:parts[i + 1]].copy())
    return (xmin, xmax, ymin, ymax), rings


def main():
    obs_df = rdata.data_frame(rdata.load_rda(f"{REF}/data/BD_Obs.rda")["BD_Obs"])
    crd_df = rdata.data_frame(rdata.load_rda(f"{REF}/data/BD_Coord.rda")["BD_Coord"])
    stations = [k for k in obs_df if k != "Date" and "::" not in k]
    dates = obs_df["Date"]                       # R Date = days since 1970-01-01
    obs = np.column_stack([obs_df[s] for s in stations])
    cod = crd_df["Cod"]
    bbox, rings = read_shp(f"{REF}/inst/extdata/study_area.shp")
    prj = open(f"{REF}/inst/extdata/study_area.prj").read().strip()
    assert len(rings) == 1
    np.savez_compressed(os.path.join(HERE, "packaged_inputs.npz"), stations=np.array(stations), dates=dates, obs=obs,
                        cod=np.array(cod), X=crd_df["X"], Y=crd_df["Y"], Z=crd_df["Z"], bbox=np.array(bbox),
                        ring=rings[0], prj=np.array(prj))

    # ---- cfg 1 / cfg 2 through the oracle ----
    grid = ref.grid_from_extent(bbox[0], bbox[1], bbox[2], bbox[3], 5000.0)
    train, test = ref.select_manual(stations, "M001")
    xy = {c: (x, y) for c, x, y in zip(cod, crd_df["X"], crd_df["Y"])}
    col = {s: i for i, s in enumerate(stations)}
    # IDW feeds stations Cod-sorted (R/IDW.R:222-245); Cressman in BD_Obs column order (R/Cressman.R:258)
    tr_idw = sorted(train)
    sx = np.array([xy[c][0] for c in tr_idw]); sy = np.array([xy[c][1] for c in tr_idw])
    ob = obs[:, [col[c] for c in tr_idw]]
    idw = ref.idw_reference(grid["X"], grid["Y"], sx, sy, ob, 2.0)
    tx = np.array([xy[c][0] for c in test]); ty = np.array([xy[c][1] for c in test])
    tobs = obs[:, [col[c] for c in test]]
    val_idw = ref.validate_reference(grid, idw, tx, ty, tobs)
    radii = ref.cressman_radii_m([50, 20, 10])
    sx2 = np.array([xy[c][0] for c in train]); sy2 = np.array([xy[c][1] for c in train])
    ob2 = obs[:, [col[c] for c in train]]
    crs = ref.cressman_reference(grid["X"], grid["Y"], sx2, sy2, ob2, radii)
    rain = dict(no_rain=(0, 1), light_rain=(1, 5), moderate_rain=(5, 20), heavy_rain=(20, 40), extremely_rain=(40, np.inf))
    val_crs = [ref.validate_reference(grid, c, tx, ty, tobs, rain) for c in crs]
    gofcols = ref.GOF_COLUMNS
    catcols = ["POD", "SR", "CSI", "HB", "FAR", "HK", "HSS", "ETS"]
    np.savez_compressed(
        os.path.join(HERE, "packaged_golden.npz"),
        grid_X=grid["X"], grid_Y=grid["Y"], grid_dims=np.array([grid["ncol"], grid["nrow"]]),
        grid_ext=np.array([grid["xmin"], grid["xmax"], grid["ymin"], grid["ymax"]]),
        idw=idw, idw_gof=np.array([[v["gof"][k] for k in gofcols] for v in val_idw]),
        radii_m=radii, cressman=np.stack(crs),
        cressman_gof=np.array([[[v["gof"][k] for k in gofcols] for v in vr] for vr in val_crs]),
        cressman_cat=np.array([[[[row[k] for k in catcols] for row in v["categorical"]] for v in vr] for vr in val_crs]),
        cat_names=np.array([row["Category"] for row in val_crs[0][0]["categorical"]]),
    )

    # ---- small synthetic goldens ----
    out = {}
    cx, cy, sx, sy, ob = synthetic.synthetic_case(24, 20, 40, 24, na_frac=0.1, plant_zero_dist=4, all_na_dates=(3,),
                                                  all_zero_dates=(7,))
    out["a_idw"] = ref.idw_reference(cx, cy, sx, sy, ob, 2.0)
    out["a_idw_p3"] = ref.idw_reference(cx, cy, sx, sy, ob, 3.0)
    out["a_crs"] = np.stack(ref.cressman_reference(cx, cy, sx, sy, ob, ref.cressman_radii_m([2, 4, 8, 30])))
    cx, cy, sx, sy, ob = synthetic.synthetic_case(20, 9, 37, 280, plant_zero_dist=2)
    out["b_idw"] = ref.idw_reference(cx, cy, sx, sy, ob, 2.0)
    np.savez_compressed(os.path.join(HERE, "synthetic_golden.npz"), **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
