"""Host-side stand-ins for the two terra classes on the path's boundary.

The reference's hot path sits between terra objects: a ``SpatVector`` study area comes in
(R/IDW.R:137-150, :198-204) and ``SpatRaster`` stacks go out (R/IDW.R:355-403,
R/Cressman.R:343-395).  terra (C++/GDAL) stays in R in a real deployment; here, where R is absent,
these two small NumPy classes carry exactly the pieces of state the path touches: the extent and
CRS of the polygon layer, the grid geometry, the layer names and the cell values.

Nothing here is on the GPU hot path.
"""
from __future__ import annotations

import math
import os
import struct

import numpy as np


# --------------------------------------------------------------------------------------
# SpatVector (polygons)
# --------------------------------------------------------------------------------------
class SpatVector:
    """Polygon layer: list of rings (each an (n, 2) float64 array), extent and CRS WKT."""

    def __init__(self, rings, crs="", bbox=None):
        self.rings = [np.asarray(r, dtype=np.float64) for r in rings]
        self.crs = crs if crs is not None else ""
        if bbox is None:
            allp = np.concatenate(self.rings, axis=0)
            bbox = (allp[:, 0].min(), allp[:, 0].max(), allp[:, 1].min(), allp[:, 1].max())
        self.xmin, self.xmax, self.ymin, self.ymax = (float(v) for v in bbox)

    def ext(self):
        """terra::ext(): (xmin, xmax, ymin, ymax)."""
        return self.xmin, self.xmax, self.ymin, self.ymax

    def is_lonlat(self) -> bool:
        """terra::is.lonlat(): geographic (unprojected) CRS."""
        w = self.crs.strip().upper()
        if not w:
            return False
        return (w.startswith("GEOGCS") or w.startswith("GEOGCRS") or w.startswith("GEODCRS")) and "PROJCS" not in w

    @staticmethod
    def from_bbox(xmin, xmax, ymin, ymax, crs="LOCAL_CS[\"synthetic metres\"]"):
        ring = np.array([[xmin, ymin], [xmin, ymax], [xmax, ymax], [xmax, ymin], [xmin, ymin]], dtype=np.float64)
        return SpatVector([ring], crs=crs, bbox=(xmin, xmax, ymin, ymax))


def vect(path: str) -> SpatVector:
    """terra::vect(<file>.shp) for polygon shapefiles (shape types 5/15/25): reads the header
    bounding box, every ring of every record and the sidecar .prj (CRS WKT)."""
    with open(path, "rb") as f:
        buf = f.read()
    if struct.unpack(">i", buf[0:4])[0] != 9994:
        raise ValueError(f"{path} is not an ESRI shapefile")
    shape_type = struct.unpack("<i", buf[32:36])[0]
    xmin, ymin, xmax, ymax = struct.unpack("<4d", buf[36:68])
    if shape_type not in (5, 15, 25):
        raise ValueError(f"only polygon shapefiles are supported (shape type {shape_type})")
    rings = []
    off = 100
    while off + 8 <= len(buf):
        clen = struct.unpack(">i", buf[off + 4:off + 8])[0] * 2
        rec = buf[off + 8:off + 8 + clen]
        off += 8 + clen
        st = struct.unpack("<i", rec[0:4])[0]
        if st == 0:
            continue
        nparts, npoints = struct.unpack("<2i", rec[36:44])
        parts = list(struct.unpack(f"<{nparts}i", rec[44:44 + 4 * nparts])) + [npoints]
        pts = np.frombuffer(rec, dtype="<f8", count=2 * npoints, offset=44 + 4 * nparts).reshape(npoints, 2)
        for i in range(nparts):
            rings.append(pts[parts[i]:parts[i + 1]].copy())
    crs = ""
    prj = os.path.splitext(path)[0] + ".prj"
    if os.path.exists(prj):
        with open(prj, "r", encoding="utf-8", errors="replace") as f:
            crs = f.read().strip()
    return SpatVector(rings, crs=crs, bbox=(xmin, xmax, ymin, ymax))


# --------------------------------------------------------------------------------------
# grid geometry (terra::rast(ext, resolution=))
# --------------------------------------------------------------------------------------
class GridGeometry:
    """Regular raster geometry as terra builds it from an extent and a resolution
    (R/IDW.R:198-204, R/Cressman.R:216-225): (xmin, ymin) are kept, ncol/nrow = round(range/res)
    (at least 1) and (xmax, ymax) are moved to xmin + ncol*res / ymin + nrow*res."""

    def __init__(self, xmin, xmax, ymin, ymax, ncol, nrow, crs=""):
        self.xmin, self.xmax, self.ymin, self.ymax = float(xmin), float(xmax), float(ymin), float(ymax)
        self.ncol, self.nrow = int(ncol), int(nrow)
        self.crs = crs

    @classmethod
    def from_extent(cls, ext, res_m, crs=""):
        xmin, xmax, ymin, ymax = ext
        ncol = int(max(1.0, math.floor((xmax - xmin) / res_m + 0.5)))
        nrow = int(max(1.0, math.floor((ymax - ymin) / res_m + 0.5)))
        return cls(xmin, xmin + ncol * res_m, ymin, ymin + nrow * res_m, ncol, nrow, crs)

    @property
    def n_cells(self):
        return self.ncol * self.nrow

    @property
    def xres(self):
        return (self.xmax - self.xmin) / self.ncol

    @property
    def yres(self):
        return (self.ymax - self.ymin) / self.nrow

    def cell_centres(self):
        """as.data.frame(r, xy=TRUE, cells=TRUE) ordered by cell (R/IDW.R:249-257): row-major from
        the top-left corner.  Returns (X, Y) of length n_cells."""
        xc = self.xmin + (np.arange(self.ncol, dtype=np.float64) + 0.5) * self.xres
        yc = self.ymax - (np.arange(self.nrow, dtype=np.float64) + 0.5) * self.yres
        return np.tile(xc, self.nrow), np.repeat(yc, self.ncol)

    def cell_from_xy(self, x, y):
        """terra cellFromXY: index of the cell containing each point, -1 outside the extent; points
        on the xmax / ymin edge belong to the last column / row."""
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        inside = (x >= self.xmin) & (x <= self.xmax) & (y >= self.ymin) & (y <= self.ymax)
        with np.errstate(invalid="ignore"):
            col = np.floor((x - self.xmin) / self.xres).astype(np.int64)
            row = np.floor((self.ymax - y) / self.yres).astype(np.int64)
        col = np.where(x == self.xmax, self.ncol - 1, col)
        row = np.where(y == self.ymin, self.nrow - 1, row)
        cell = row * self.ncol + col
        return np.where(inside, cell, -1)

    def crop_window(self, v: SpatVector):
        """terra::crop(x, v, snap="out"): the sub-grid covering v's extent snapped outward to cell
        edges (clipped to this grid).  Returns (sub GridGeometry, indices of its cells in this grid)."""
        c0 = max(0, int(math.floor((v.xmin - self.xmin) / self.xres + 1e-9)))
        c1 = min(self.ncol, int(math.ceil((v.xmax - self.xmin) / self.xres - 1e-9)))
        r0 = max(0, int(math.floor((self.ymax - v.ymax) / self.yres + 1e-9)))
        r1 = min(self.nrow, int(math.ceil((self.ymax - v.ymin) / self.yres - 1e-9)))
        if c1 <= c0 or r1 <= r0:
            raise ValueError("extents do not overlap")
        sub = GridGeometry(self.xmin + c0 * self.xres, self.xmin + c1 * self.xres, self.ymax - r1 * self.yres,
                           self.ymax - r0 * self.yres, c1 - c0, r1 - r0, self.crs)
        cells = (np.arange(r0, r1)[:, None] * self.ncol + np.arange(c0, c1)[None, :]).reshape(-1)
        return sub, cells

    def polygon_cover(self, v: SpatVector):
        """Boolean n_cells mask: cell centre inside the polygon layer (even-odd rule over all rings,
        so holes are honoured) — terra::mask / rasterize with touches=FALSE."""
        X, Y = self.cell_centres()
        xc = X[: self.ncol]
        inside = np.zeros((self.nrow, self.ncol), dtype=bool)
        yrows = Y[:: self.ncol]
        for ring in v.rings:
            x0, y0 = ring[:-1, 0], ring[:-1, 1]
            x1, y1 = ring[1:, 0], ring[1:, 1]
            for r, yv in enumerate(yrows):
                cross = (y0 <= yv) != (y1 <= yv)
                if not cross.any():
                    continue
                xs = x0[cross] + (yv - y0[cross]) * (x1[cross] - x0[cross]) / (y1[cross] - y0[cross])
                # a centre is inside if an odd number of crossings lie to its right
                cnt = (xs[None, :] > xc[:, None]).sum(axis=1)
                inside[r] ^= (cnt % 2 == 1)
        return inside.reshape(-1)


# --------------------------------------------------------------------------------------
# SpatRaster stack
# --------------------------------------------------------------------------------------
class SpatRaster:
    """Multi-layer raster: `values` is (n_cells, n_layers), Fortran order (one contiguous layer per
    date, cells in terra order) — exactly what `terra::values(stack) <- matrix` takes."""

    def __init__(self, geom: GridGeometry, values, names):
        values = np.asarray(values)
        if values.ndim != 2 or values.shape[0] != geom.n_cells or values.shape[1] != len(names):
            raise ValueError("values must be (n_cells, n_layers) with one name per layer")
        self.geom = geom
        self.values = values
        self.names = [str(n) for n in names]

    def nlyr(self) -> int:
        return self.values.shape[1]

    def __len__(self):
        return self.nlyr()

    @property
    def crs(self):
        return self.geom.crs

    def layer(self, i):
        """One layer as a (nrow, ncol) array (north-up)."""
        if isinstance(i, str):
            i = self.names.index(i)
        return np.asarray(self.values[:, i]).reshape(self.geom.nrow, self.geom.ncol)

    def to_array(self):
        """(n_layers, nrow, ncol)."""
        return np.ascontiguousarray(self.values.T).reshape(self.nlyr(), self.geom.nrow, self.geom.ncol)

    def extract(self, x, y):
        """terra::extract(stack, points): (n_points, n_layers) values of the cell containing each
        point; NaN rows for points outside the extent."""
        cells = self.geom.cell_from_xy(x, y)
        out = np.full((len(cells), self.nlyr()), np.nan)
        ok = cells >= 0
        out[ok] = self.values[cells[ok], :]
        return out

    def crop_mask(self, v: SpatVector):
        """terra::mask(terra::crop(x, v, snap="out"), v) (R/IDW.R:376-381): crop to the polygon's
        extent snapped outward to cell edges, then set cells whose centre is outside to NA."""
        sub, cells = self.geom.crop_window(v)
        vals = np.array(self.values[cells, :], order="F", copy=True)
        cover = sub.polygon_cover(v)
        vals[~cover, :] = np.nan
        return SpatRaster(sub, vals, self.names)

    def write_cdf(self, filename, varname="precipitation"):
        """terra::writeCDF(x, filename, overwrite=TRUE) (R/IDW.R:389-393): NetCDF-3 classic file with
        dimensions (time, y, x); NA stored as the _FillValue."""
        from scipy.io import netcdf_file

        g = self.geom
        fill = -1.175494e38
        with netcdf_file(filename, "w", version=2) as nc:
            nc.createDimension("x", g.ncol)
            nc.createDimension("y", g.nrow)
            nc.createDimension("time", self.nlyr())
            vx = nc.createVariable("x", "d", ("x",))
            vy = nc.createVariable("y", "d", ("y",))
            vt = nc.createVariable("time", "i", ("time",))
            vx[:] = g.xmin + (np.arange(g.ncol) + 0.5) * g.xres
            vy[:] = g.ymax - (np.arange(g.nrow) + 0.5) * g.yres
            vt[:] = np.arange(self.nlyr(), dtype=np.int32)
            vt.long_name = "layer index; names in the global attribute layer_names"
            var = nc.createVariable(varname, "d", ("time", "y", "x"))
            arr = self.to_array()
            var[:] = np.where(np.isnan(arr), fill, arr)
            var._FillValue = fill
            var.missing_value = fill
            nc.layer_names = ",".join(self.names)
            nc.crs_wkt = g.crs or ""


def r_round(x, digits):
    """R's round(x, digits) (R >= 4.0.0: the candidate nearest to x among floor/ceil at 10^-digits,
    ties to the even candidate), vectorised; used for terra::app(x, function(v) round(v, n)) at
    R/IDW.R:356-358."""
    x = np.asarray(x, dtype=np.float64)
    p10 = 10.0 ** int(digits)
    ax = np.abs(x)
    with np.errstate(invalid="ignore", over="ignore"):
        x10 = p10 * ax
        i10 = np.floor(x10)
        xd = i10 / p10
        xu = np.ceil(x10) / p10
        du = xu - ax
        dd = ax - xd
        pick_d = (dd < du) | ((dd == du) & (np.mod(i10, 2.0) == 0))
        out = np.where(pick_d, xd, xu) * np.sign(x)
    out = np.where(np.isfinite(x), out, x)
    return np.where(x == 0, x, out)


def terra_round(x, digits):
    """terra's round(SpatRaster, digits) (R/Cressman.R:353-355): C++ std::round(v * 10^d) / 10^d,
    i.e. halves away from zero (un-vendored terra; restated from its documented behaviour)."""
    x = np.asarray(x, dtype=np.float64)
    p10 = 10.0 ** int(digits)
    with np.errstate(invalid="ignore", over="ignore"):
        y = x * p10
        out = np.sign(y) * np.floor(np.abs(y) + 0.5) / p10
    return np.where(np.isfinite(x), out, x)
