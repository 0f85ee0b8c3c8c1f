#!/usr/bin/env python
"""One routed aggregation at BASELINE config 5's geometry (8192 x 8192), for `ncu -k regex:k_route_...` captures:
    python tools/prof_one_routed.py max|first|count|where_max [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import datashader_b200 as ds

what = sys.argv[1]
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000_000
g = torch.Generator(device="cuda")
g.manual_seed(1)
x = torch.rand(n, generator=g, device="cuda")
y = torch.rand(n, generator=g, device="cuda")
v = torch.randn(n, generator=g, device="cuda")
ds.config.device_results = True
frame = ds.DeviceFrame({"x": x, "y": y, "value": v})
cvs = ds.Canvas(8192, 8192, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
agg = {"max": ds.max("value"), "first": ds.first("value"), "count": ds.count(), "where_max": ds.where(ds.max("value"))}[what]
for _ in range(3):
    cvs.points(frame, "x", "y", agg)
torch.cuda.synchronize()
