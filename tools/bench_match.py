#!/usr/bin/env python
"""where(max) at BASELINE config 5's geometry, stage by stage (routed max, then dsb_points_match32), over the knobs of the
second pass:  python tools/bench_match.py [n] [shift,variant ...]"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import datashader_b200 as ds
from datashader_b200 import _lib, config

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000_000
combos = [tuple(map(int, c.split(","))) for c in sys.argv[2:]] or [(4, 0)]
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
if hasattr(L, "dsb_match_configure"):
    L.dsb_match_configure.argtypes = [C.c_int, C.c_int]
ref = None
for shift, variant in combos:
    if hasattr(L, "dsb_match_configure"):      # experiment builds only (profiles/r02_where_two_pass.md)
        L.dsb_match_configure(shift, variant)
    else:                                      # shipped library: variant 0 = the queued shared-memory form, 1 = the first form
        _lib.check(L.dsb_configure(b"match_queue", int(variant == 0)))
    best = None
    for it in range(4):
        config.kernel_events.clear()
        r = cvs.points(frame, "x", "y", ds.where(ds.max("value"))).data
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in config.kernel_events]
        if it and (best is None or ms[-1] < best[-1]):
            best = ms
    if ref is None:
        ref = r.clone()
    print(json.dumps({"n": n, "shift": shift, "variant": variant, "stage_ms": best, "same_rows": bool(torch.equal(r, ref)),
                      "kernel": L.dsb_last_kernel().decode()}), flush=True)
    del r
