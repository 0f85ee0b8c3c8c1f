#!/usr/bin/env python
"""A zoom level of 16 x 16 = 256 tiles (256 x 256 pixels each) over 1e8 resident points: one Canvas.points_batch call vs 256
Canvas.points calls.  Prints ms and checks that every tile is bit-equal."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import datashader_b200 as ds  # noqa: E402
from datashader_b200 import config  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
g = torch.Generator(device="cuda"); g.manual_seed(7)
x = torch.rand(n, generator=g, device="cuda") * 16
y = torch.rand(n, generator=g, device="cuda") * 16
v = torch.randn(n, generator=g, device="cuda")
frame = ds.DeviceFrame({"x": x, "y": y, "v": v})
views = [((float(ix), float(ix + 1)), (float(iy), float(iy + 1))) for iy in range(16) for ix in range(16)]
cvs = ds.Canvas(256, 256)
config.device_results = True
out = {"points": n, "tiles": len(views)}
for name, agg in (("count", ds.count()), ("mean", ds.mean("v")), ("max", ds.max("v"))):
    for _ in range(2):
        batch = cvs.points_batch(frame, "x", "y", agg, views, grid=(16, 16))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        batch = cvs.points_batch(frame, "x", "y", agg, views, grid=(16, 16))
    torch.cuda.synchronize()
    t_batch = (time.perf_counter() - t0) / 3
    t0 = time.perf_counter()
    singles = [ds.Canvas(256, 256, x_range=xr, y_range=yr).points(frame, "x", "y", agg) for xr, yr in views]
    torch.cuda.synchronize()
    t_single = time.perf_counter() - t0
    same = all(torch.equal(torch.nan_to_num(a.data.double(), nan=-7.0), torch.nan_to_num(b.data.double(), nan=-7.0)) if name != "mean"
               else torch.allclose(torch.nan_to_num(a.data, nan=-7.0), torch.nan_to_num(b.data, nan=-7.0), rtol=1e-12, atol=0)
               for a, b in zip(batch, singles))
    out[name] = {"batch_ms": t_batch * 1e3, "per_tile_us": t_batch * 1e6 / len(views), "single_calls_ms": t_single * 1e3,
                 "single_per_tile_us": t_single * 1e6 / len(views), "speedup": t_single / t_batch, "tiles_equal": bool(same)}
print(json.dumps(out))
