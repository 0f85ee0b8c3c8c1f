#!/usr/bin/env python
"""Host-side overhead of one Canvas.points call on a small resident frame (interactive zoom / pan regime)."""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import datashader_b200 as ds
from datashader_b200 import config

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
g = torch.Generator(device="cuda"); g.manual_seed(1)
f = ds.DeviceFrame({"x": torch.rand(n, generator=g, device="cuda"), "y": torch.rand(n, generator=g, device="cuda"),
                    "value": torch.randn(n, generator=g, device="cuda")})
config.device_results = True
cvs = ds.Canvas(900, 525, x_range=(0.2, 0.7), y_range=(0.1, 0.6))
for agg in (ds.count(), ds.mean("value"), ds.max("value")):
    for _ in range(20):
        cvs.points(f, "x", "y", agg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(300):
        cvs.points(f, "x", "y", agg)
    torch.cuda.synchronize()
    print(type(agg).__name__, "us per call:", round((time.perf_counter() - t0) / 300 * 1e6, 1))
pr = cProfile.Profile(); pr.enable()
for _ in range(300):
    cvs.points(f, "x", "y", ds.mean("value"))
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
