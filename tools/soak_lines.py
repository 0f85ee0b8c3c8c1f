#!/usr/bin/env python
"""Soak of the antialiased rasteriser: random line sets, canvases and widths; the warp-balanced kernel (rows, then pixels spread
over the lanes) against the one-thread-per-segment kernel.  any / max must be bit-equal, the atomic f64 sums equal to 1e-6, the float32 count canvas to 1e-4.
    python tools/soak_lines.py [configs=40] [seed=0]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import datashader_b200 as ds
from datashader_b200 import _lib

nconf = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
L = _lib.lib()
ds.config.device_results = True
bad = 0
for it in range(nconf):
    nl, nv = int(rng.integers(1, 3000)), int(rng.integers(2, 60))
    W, H = int(rng.integers(2, 2500)), int(rng.integers(2, 1500))
    lw = float(rng.choice([0.5, 1.0, 2.0, 3.7, 9.0]))
    step = float(rng.choice([0.002, 0.02, 0.2]))
    dt = np.float32 if it % 2 else np.float64
    xs = np.cumsum(rng.normal(0, step, (nl, nv)), axis=1).astype(dt) + dt(rng.random())
    ys = np.cumsum(rng.normal(0, step, (nl, nv)), axis=1).astype(dt) + dt(rng.random())
    ys[rng.random((nl, nv)) < 0.01] = np.nan
    val = rng.standard_normal(nl)
    val[rng.random(nl) < 0.05] = np.nan
    cols = {f"x{j}": xs[:, j] for j in range(nv)}
    cols.update({f"y{j}": ys[:, j] for j in range(nv)})
    cols["val"] = val
    frame = ds.DeviceFrame({k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in cols.items()})
    xc, yc = [f"x{j}" for j in range(nv)], [f"y{j}" for j in range(nv)]
    cvs = ds.Canvas(W, H, x_range=(-0.2, 1.3), y_range=(0.1, 0.9))
    for name, agg in (("any", ds.any()), ("max", ds.max("val")), ("count", ds.count()), ("sum", ds.sum("val")), ("mean", ds.mean("val"))):
        out = []
        for balanced in (1, 0):
            _lib.check(L.dsb_lines_configure(balanced))
            out.append(cvs.line(frame, x=xc, y=yc, axis=1, agg=agg, line_width=lw).data.double())
        a, b = out
        nan_same = torch.equal(torch.isnan(a), torch.isnan(b))
        if name in ("any", "max"):
            same = nan_same and torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))
        else:       # count is a float32 canvas of atomic adds of (coverage - previous coverage): the order of the adds differs
            tol = dict(rtol=1e-4, atol=2e-5) if name == "count" else dict(rtol=1e-6, atol=1e-9)
            same = nan_same and torch.allclose(torch.nan_to_num(a, nan=0.0), torch.nan_to_num(b, nan=0.0), **tol)
        if not same:
            bad += 1
            d = (torch.nan_to_num(a, nan=0.0) - torch.nan_to_num(b, nan=0.0)).abs().max().item() if nan_same else float("nan")
            print("MISMATCH", it, nl, nv, W, H, lw, dt.__name__, name, "nan masks equal:", nan_same, "max abs diff:", d)
    print(it, nl, nv, W, H, lw, dt.__name__, flush=True)
_lib.check(L.dsb_lines_configure(1))
print("soak done, mismatches:", bad)
