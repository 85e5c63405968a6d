"""ORACLE / TEST INFRASTRUCTURE — not product code.

Second, independent restatement used ONLY to cross-check oracle/reference_port.py on tiny cases:
scalar Python loops written from the *mathematical* definition of what the reference computes
(one cell, one date at a time; exact `math.fsum` accumulation), so that a slip in the vectorised
restatement (wrong axis, wrong mask, wrong order of the early-outs) cannot hide.
Sources: R/IDW.R:288-342 and R/Cressman.R:290-321.
"""
from __future__ import annotations

import math


def idw_scalar(cx, cy, sx, sy, obs, p=2.0):
    n_cells, n_dates, n_st = len(cx), len(obs), len(sx)
    out = [[float("nan")] * n_dates for _ in range(n_cells)]
    for t in range(n_dates):
        avail = [s for s in range(n_st) if not math.isnan(obs[t][s])]
        if not avail:
            continue                                     # all-NA date -> NA layer
        if all(abs(obs[t][s]) < 1e-12 for s in avail):
            for c in range(n_cells):
                out[c][t] = 0.0                          # exact-zero layer
            continue
        for c in range(n_cells):
            d = [math.sqrt((cx[c] - sx[s]) ** 2 + (cy[c] - sy[s]) ** 2) for s in range(n_st)]
            colo = [s for s in avail if d[s] == 0.0]
            if colo:
                pred, hit = None, 0
                for s in colo:                            # running mean in station order
                    pred = obs[t][s] if pred is None else (pred * hit + obs[t][s]) / (hit + 1)
                    hit += 1
                out[c][t] = pred
                continue
            w = [d[s] ** (-p) for s in avail]
            num = math.fsum(wi * obs[t][s] for wi, s in zip(w, avail))
            den = math.fsum(w)
            v = num / den if den != 0 else float("nan")
            out[c][t] = v if math.isfinite(v) else float("nan")
    return out


def cressman_scalar(cx, cy, sx, sy, obs, radii_m):
    n_cells, n_dates, n_st = len(cx), len(obs), len(sx)
    outs = []
    for R in radii_m:
        out = [[float("nan")] * n_dates for _ in range(n_cells)]
        for c in range(n_cells):
            w = []
            for s in range(n_st):
                d2 = (cx[c] - sx[s]) ** 2 + (cy[c] - sy[s]) ** 2
                d = math.sqrt(d2)
                w.append(0.0 if d > R else (R * R - d * d) / (R * R + d * d))
            den = math.fsum(w)                           # ALL stations, NA or not
            for t in range(n_dates):
                num = math.fsum(w[s] * obs[t][s] for s in range(n_st) if not math.isnan(obs[t][s]))
                if den == 0:
                    continue
                v = num / den
                out[c][t] = v if math.isfinite(v) else float("nan")
        outs.append(out)
    return outs
