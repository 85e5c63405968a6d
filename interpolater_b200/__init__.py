"""interpolater_b200 — B200-native engine for InterpolateR's data-parallel hot path:
`IDW()` and `Cressman()` over every grid cell x every date.

Public surface (mirrors the reference's exported functions for this path, NAMESPACE:3-8):
    IDW, Cressman                      — drop-in signatures (api.py)
    vect, SpatVector, SpatRaster       — stand-ins for the terra objects on the boundary
    engine.idw_grid / cressman_grid    — array-level twins of the C ABI (include/interpolater_b200.h)
    create_data                        — station CSVs -> BD_Obs table (R/create_data.R:43-94)

The compute core is hand-written sm_100a CUDA (csrc/), reached through a C-ABI shared library;
there is no CPU fallback.
"""
from .api import IDW, Cressman  # noqa: F401
from .ingest import create_data  # noqa: F401
from .spatial import GridGeometry, SpatRaster, SpatVector, vect  # noqa: F401

__all__ = ["IDW", "Cressman", "create_data", "vect", "SpatVector", "SpatRaster", "GridGeometry"]
__version__ = "0.1.0"
