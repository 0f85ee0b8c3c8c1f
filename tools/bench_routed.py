"""Routed path (csrc/routed.cu) at BASELINE config 5's geometry: pass 2 through the TMA ring vs plain loads, per reduction.
usage: python tools/bench_routed.py [n_points=1e9]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import datashader_b200 as ds
from datashader_b200 import _lib

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
L = _lib.lib()
g = torch.Generator(device="cuda")
g.manual_seed(1)
x = torch.rand(n, generator=g, device="cuda")
y = torch.rand(n, generator=g, device="cuda")
v = torch.randn(n, generator=g, device="cuda")
frame = ds.DeviceFrame({"x": x, "y": y, "value": v})
cvs = ds.Canvas(8192, 8192, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
ds.config.device_results = True
out = {"n": n}
ref = {}
for tma in (1, 0, 1, 0):
    _lib.check(L.dsb_configure(b"routed_tma", tma))
    for name, agg in (("max", ds.max("value")), ("first", ds.first("value")), ("count", ds.count())):
        for _ in range(2):
            r = cvs.points(frame, "x", "y", agg)
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = cvs.points(frame, "x", "y", agg)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        assert b"k_route" in L.dsb_last_kernel()
        d = torch.nan_to_num(r.data, nan=-7.0)
        if name in ref:
            assert torch.equal(ref[name], d), name
        ref[name] = d
        out.setdefault(f"{name}_{'tma' if tma else 'ldg'}_ms", []).append(round(sorted(ts)[2], 3))
print(json.dumps(out))
