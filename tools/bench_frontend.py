#!/usr/bin/env python
"""Front-end cost of the point kernels: count() at 900x525 with a range that drops every point (no scatter at all),
K1 (global-RED kernel) vs K2 (privatised kernel).  Isolates streaming + mapping from the atomics."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import datashader_b200 as ds
torch.cuda.set_device(0)
ds.config.device_results = True
n = 1_000_000_000
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.rand(n, generator=g, device="cuda"); y = torch.rand(n, generator=g, device="cuda")
frame = ds.DeviceFrame({"x": x, "y": y})
for name, xr in (("all points dropped (x_range=(2,3))", (2.0, 3.0)), ("all points kept", (0.0, 1.0))):
    cvs = ds.Canvas(900, 525, x_range=xr, y_range=(0.0, 1.0))
    for mode, thr in (("K2", 0), ("K1", 1 << 62)):
        ds.config.priv_min_rows = thr
        for _ in range(2):
            cvs.points(frame, "x", "y", ds.count())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            cvs.points(frame, "x", "y", ds.count())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(json.dumps({"case": name, "kernel": mode, "ms": round(ms, 3), "gpts": round(n / ms / 1e6, 1), "GBps": round(8 * n / ms / 1e6, 1)}), flush=True)
