#!/usr/bin/env python
"""first('value') at BASELINE config 5's geometry over the head size of the split (rows per canvas cell that are routed before the
rest is only filtered):  python tools/bench_first.py [n] [head_per_cell ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import datashader_b200 as ds
from datashader_b200 import _lib, config

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000_000
heads = [int(h) for h in sys.argv[2:]] or [8]
g = torch.Generator(device="cuda")
g.manual_seed(5)
x = torch.rand(n, generator=g, device="cuda")
y = torch.rand(n, generator=g, device="cuda")
v = torch.randn(n, generator=g, device="cuda")
frame = ds.DeviceFrame({"x": x, "y": y, "value": v})
cvs = ds.Canvas(8192, 8192, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
config.device_results = True
config.time_kernels = True
L = _lib.lib()
ref = None
for h in heads:
    _lib.check(L.dsb_configure(b"routed_head_per_cell", h))
    for name, agg in (("first", ds.first("value")), ("where_first", ds.where(ds.first("value"))), ("last", ds.last("value"))):
        best = None
        for it in range(4):
            config.kernel_events.clear()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = cvs.points(frame, "x", "y", agg).data
            e1.record()
            torch.cuda.synchronize()
            ms = (e0.elapsed_time(e1), sum(a.elapsed_time(b) for a, b in config.kernel_events))
            if it and (best is None or ms[0] < best[0]):
                best = ms
        same = None
        if name == "first":
            if ref is None:
                ref = r.clone()
            same = bool(torch.equal(torch.nan_to_num(r, nan=-7.0), torch.nan_to_num(ref, nan=-7.0)))
        print(json.dumps({"n": n, "head_per_cell": h, "agg": name, "call_ms": best[0], "kernel_ms": best[1], "same_as_first_run": same,
                          "kernel": L.dsb_last_kernel().decode()}), flush=True)
        del r
