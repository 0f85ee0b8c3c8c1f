#!/usr/bin/env python
"""count() at 900x525 for several row counts with the privatised kernel (K2) on and off: finds the crossover
below which the fixed cost of K2 (zeroing + flushing 148 private canvases) loses to plain global REDs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import datashader_b200 as ds  # noqa: E402

torch.cuda.set_device(0)
ds.config.device_results = True
g = torch.Generator(device="cuda"); g.manual_seed(1)
nmax = 400_000_000
x = torch.rand(nmax, generator=g, device="cuda"); y = torch.rand(nmax, generator=g, device="cuda")
v = torch.randn(nmax, generator=g, device="cuda")
cvs = ds.Canvas(900, 525, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
for n in (2_000_000, 5_000_000, 10_000_000, 20_000_000, 50_000_000, 100_000_000, 200_000_000, 400_000_000):
    frame = ds.DeviceFrame({"x": x[:n], "y": y[:n], "value": v[:n]})
    row = {"n": n}
    for agg_name, agg in (("count", ds.count()), ("mean", ds.mean("value"))):
        for mode, thr in (("priv", 0), ("generic", 1 << 62)):
            ds.config.priv_min_rows = thr
            for _ in range(3):
                cvs.points(frame, "x", "y", agg)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                cvs.points(frame, "x", "y", agg)
            e1.record(); torch.cuda.synchronize()
            row[f"{agg_name}_{mode}_ms"] = round(e0.elapsed_time(e1) / 10, 4)
    print(json.dumps(row), flush=True)
