#!/usr/bin/env python
"""One aggregation at the headline geometry, for `ncu -k regex:<kernel>` captures:
    python tools/prof_one.py max|first|last|where_max|by_count_c3  [n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import datashader_b200 as ds
from datashader_b200 import config

what = sys.argv[1]
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000_000
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.rand(n, generator=g, device="cuda"); y = torch.rand(n, generator=g, device="cuda")
v = torch.randn(n, generator=g, device="cuda")
config.device_results = True
if what == "by_count_c3":
    cat = torch.randint(0, 16, (n,), generator=g, device="cuda", dtype=torch.int8)
    frame = ds.DeviceFrame({"x": x, "y": y, "cat": cat}, categories={"cat": [f"c{i}" for i in range(16)]})
    cvs = ds.Canvas(1920, 1080, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    agg = ds.by("cat", ds.count())
else:
    frame = ds.DeviceFrame({"x": x, "y": y, "value": v})
    cvs = ds.Canvas(900, 525, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    agg = {"max": ds.max("value"), "first": ds.first("value"), "last": ds.last("value"), "where_max": ds.where(ds.max("value"))}[what]
for _ in range(3):
    cvs.points(frame, "x", "y", agg)
torch.cuda.synchronize()
