#!/usr/bin/env python
"""Timing of the K2 kernels at the headline geometry (900x525, n uniform float32 points, device-resident):
    python tools/bench_k2.py [n] [key=value ...]     # key=value pairs go to dsb_configure before timing
Prints ms / Gpts/s of count() and mean('value') and the kernel the library chose."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import datashader_b200 as ds
from datashader_b200 import _lib, config

n = int(float(sys.argv[1])) if len(sys.argv) > 1 and "=" not in sys.argv[1] else 1_000_000_000
L = _lib.lib()
for kv in sys.argv[1:]:
    if "=" in kv:
        k, v = kv.split("=")
        _lib.check(L.dsb_configure(k.encode(), int(v)), kv)
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.rand(n, generator=g, device="cuda"); y = torch.rand(n, generator=g, device="cuda")
v = torch.randn(n, generator=g, device="cuda")
v[::1000] = float("nan")
frame = ds.DeviceFrame({"x": x, "y": y, "value": v})
cvs = ds.Canvas(900, 525, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
config.device_results = True
config.priv_min_rows = 0
for name, agg in (("count", ds.count()), ("mean", ds.mean("value"))):
    for _ in range(3):
        out = cvs.points(frame, "x", "y", agg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = cvs.points(frame, "x", "y", agg)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:5s} {ms:7.3f} ms  {n / ms / 1e6:7.1f} Gpts/s   {L.dsb_last_kernel().decode()}", flush=True)
