#!/usr/bin/env python
"""One LinesAxis1 aggregation for ncu: python tools/prof_lines.py [line_width] [nlines] [max|min]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import datashader_b200 as ds
from datashader_b200 import config
lw = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
nl, nv = (int(sys.argv[2]) if len(sys.argv) > 2 else 20000), 1000
g = torch.Generator(device="cuda"); g.manual_seed(4)
xs = torch.arange(nv, device="cuda", dtype=torch.float32).repeat(nl, 1)
ys = torch.randn(nl, nv, generator=g, device="cuda").cumsum(dim=1)
val = torch.rand(nl, generator=g, device="cuda")
cols = {f"x{j}": xs[:, j].contiguous() for j in range(nv)}
cols.update({f"y{j}": ys[:, j].contiguous() for j in range(nv)})
cols["value"] = val
frame = ds.DeviceFrame(cols)
cvs = ds.Canvas(3840, 2160, x_range=(0.0, float(nv - 1)), y_range=(float(ys.min()), float(ys.max())))
config.device_results = True
AGG = ds.min("value") if (len(sys.argv) > 3 and sys.argv[3] == "min") else ds.max("value")
for _ in range(3):
    cvs.line(frame, x=[f"x{j}" for j in range(nv)], y=[f"y{j}" for j in range(nv)], axis=1, agg=AGG, line_width=lw)
torch.cuda.synchronize()
import cProfile, pstats, time
t0 = time.perf_counter()
for _ in range(3):
    cvs.line(frame, x=[f"x{j}" for j in range(nv)], y=[f"y{j}" for j in range(nv)], axis=1, agg=AGG, line_width=lw)
torch.cuda.synchronize()
print("ms per call (wall):", (time.perf_counter() - t0) / 3 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    cvs.line(frame, x=[f"x{j}" for j in range(nv)], y=[f"y{j}" for j in range(nv)], axis=1, agg=AGG, line_width=lw)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
