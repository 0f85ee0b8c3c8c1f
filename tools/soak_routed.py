#!/usr/bin/env python
"""Soak of the routed path: random canvas shapes (1 .. ~1500 buckets), row counts and ranges; every routed reduction must be
bit-equal to the unbanded generic kernel on the same frame - including first / last through a random head size of the
head-then-filter split (0 = off, rows shuffled or sorted in space) and where(max | min) through the two-pass form.
    python tools/soak_routed.py [configs=40] [seed=0]"""
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
ds.config.priv_count = ds.config.count16 = False
g = torch.Generator(device="cuda")
bad = 0
for it in range(nconf):
    W, H = int(rng.integers(3, 9000)), int(rng.integers(3, 9000))
    if rng.random() < 0.5:                    # small canvases: many rows per pixel, so first / last take the head-then-filter split
        W, H = int(rng.integers(3, 1200)), int(rng.integers(3, 1200))
    if W * H > 70_000_000:
        H = 70_000_000 // W
    n = int(rng.integers(1, 6_000_000))
    g.manual_seed(it)
    lo, span = float(rng.choice([0.0, -2.0, 100.0])), float(rng.choice([1.0, 0.25, 40.0]))
    x = lo + span * (torch.rand(n, generator=g, device="cuda") * 1.1 - 0.05)
    y = lo + span * (torch.rand(n, generator=g, device="cuda") * 1.1 - 0.05)
    v = torch.randn(n, generator=g, device="cuda")
    if rng.random() < 0.5:
        v = torch.round(v * 4) / 4            # ties everywhere: where(max | min) must pick the earliest row
    v[::53] = float("nan")
    if rng.random() < 0.3:                    # rows sorted in space: the head of first / last covers a strip of the canvas only
        order = torch.argsort(y if rng.random() < 0.5 else -y)
        x, y, v = x[order].contiguous(), y[order].contiguous(), v[order].contiguous()
    head = int(rng.choice([0, 1, 2, 5, 10]))
    _lib.check(L.dsb_configure(b"routed_head_per_cell", head))
    frame = ds.DeviceFrame({"x": x, "y": y, "value": v})
    cvs = ds.Canvas(W, H, x_range=(lo, lo + span), y_range=(lo, lo + span))
    for agg in (ds.max("value"), ds.min("value"), ds.first("value"), ds.last("value"), ds.count(), ds.where(ds.max("value")),
                ds.where(ds.min("value")), ds.where(ds.first("value")), ds.where(ds.last("value"))):
        res = {}
        for mode in ("routed", "generic"):
            ds.config.routed = mode == "routed"
            ds.config.routed_min_rows, ds.config.l2_budget_bytes = (0, 1) if mode == "routed" else (1 << 24, 96 << 20)
            if it % 3 == 2 and mode == "routed":      # every third configuration keeps the real L2 budget: max / min then take the
                ds.config.l2_budget_bytes = 96 << 20  # head + threshold-filtered rest (dsb_points_minmax_rest), first / last the split
            ds.config.minmax_split_rows_per_cell, ds.config.minmax_head_rows_per_cell = (2, 1) if mode == "routed" else (0, 128)
            _lib.check(L.dsb_routed_configure(0 if mode == "routed" else 1 << 24))
            _lib.check(L.dsb_configure(b"l2_band_bytes", 0 if mode == "generic" else 96 << 20))
            _lib.check(L.dsb_configure(b"mono", 0 if mode == "generic" else 1))
            res[mode] = cvs.points(frame, "x", "y", agg).data.clone()
            res[mode + "_kernel"] = L.dsb_last_kernel()
            if mode == "routed" and isinstance(agg, ds.max):
                kern_max = L.dsb_last_kernel().decode()
        a, b = res["routed"], res["generic"]
        same = torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0)) if a.dtype.is_floating_point else torch.equal(a, b)
        if not same:
            bad += 1
            print("MISMATCH", it, W, H, n, lo, span, type(agg).__name__, res["routed_kernel"])
    print(it, W, H, n, "head", head, res["routed_kernel"].decode()[:60], "| max:", kern_max[:40], flush=True)
_lib.check(L.dsb_configure(b"routed_head_per_cell", 10))
ds.config.minmax_split_rows_per_cell, ds.config.minmax_head_rows_per_cell = 512, 128
print("soak done, mismatches:", bad)
