"""Drop-in signature at scale: IDW()/Cressman() on a 1000 x 1000 grid, 2000 stations, 365 dates."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
import interpolater_b200 as ipr
from interpolater_b200 import synthetic
ncol = nrow = 1000; S = 2000; D = 365
cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, S, D, na_frac=0.01)
names = [f"S{i:04d}" for i in range(S)]
BD_Obs = pd.DataFrame(obs, columns=names); BD_Obs.insert(0, "Date", pd.date_range("2000-01-01", periods=D))
BD_Coord = pd.DataFrame({"Cod": names, "X": sx, "Y": sy, "Z": 0.0})
th = np.linspace(0, 2 * np.pi, 400)
ring = np.c_[200000 + 500e3 + 450e3 * np.cos(th), 9000000 + 500e3 + 400e3 * np.sin(th)]; ring[-1] = ring[0]
shp = ipr.SpatVector([ring], crs='PROJCS["synthetic"]', bbox=(200000.0, 1200000.0, 9000000.0, 10000000.0))
held = names[::5]
t = time.perf_counter()
out = ipr.IDW(BD_Obs, BD_Coord, shp, grid_resolution=1, p=2, n_round=1, stat_validation=held)
print("IDW() wall", round(time.perf_counter() - t, 2), "s", out["Ensamble"].values.shape, "NA share", float(np.isnan(out["Ensamble"].values[:, 0]).mean()))
print(out["Validation"].head(3))
t = time.perf_counter()
outc = ipr.Cressman(BD_Obs, BD_Coord, shp, grid_resolution=1, search_radius=[50, 100], n_round=1, stat_validation=held[:50])
print("Cressman() wall", round(time.perf_counter() - t, 2), "s", list(outc["Ensamble"]))
