#!/usr/bin/env python
"""A canvas with more than 2^31 cells (50 000 x 46 000 = 2.3e9 pixels, 9.2 GB of u32) on one B200: cell indices no longer fit
32 bits.  Checks that need no oracle: count() conserves the rows and nests exactly into the 500 x 460 canvas (x, y are float32
in [0, 1): x * 50 000 and x * 500 are exact in float64, so fine pixel // 100 == coarse pixel); max('value') nests the same way.
    python tools/check_huge_canvas.py [n=2e8]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import datashader_b200 as ds
from datashader_b200 import _lib

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
W, H, F = 50_000, 46_000, 100
g = torch.Generator(device="cuda")
g.manual_seed(3)
x = torch.rand(n, generator=g, device="cuda")
y = torch.rand(n, generator=g, device="cuda")
v = torch.randn(n, generator=g, device="cuda")
frame = ds.DeviceFrame({"x": x, "y": y, "value": v})
ds.config.device_results = True
fine = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
coarse = ds.Canvas(W // F, H // F, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
ok = True


def check(name, cond):
    global ok
    print(f"{name}: {'ok' if cond else 'FAILED'}  [{_lib.lib().dsb_last_kernel().decode()[:80]}]", flush=True)
    ok = ok and bool(cond)


def dev(a):
    d = a.data
    return d if isinstance(d, torch.Tensor) else torch.as_tensor(d, device="cuda")


c = dev(fine.points(frame, "x", "y", ds.count()))
check("count conserves the rows", int(c.sum(dtype=torch.int64)) == n)
nested = c.view(H // F, F, W // F, F).sum(dim=(1, 3), dtype=torch.int64)
del c
want = dev(coarse.points(frame, "x", "y", ds.count())).to(torch.int64)
check("count nests into the 500 x 460 canvas", torch.equal(nested, want))
m = dev(fine.points(frame, "x", "y", ds.max("value")))
nested = torch.nan_to_num(m, nan=-float("inf")).view(H // F, F, W // F, F).amax(dim=(1, 3))
del m
want = dev(coarse.points(frame, "x", "y", ds.max("value")))
check("max nests into the 500 x 460 canvas", torch.equal(nested, want.to(nested.dtype)))
print("huge canvas:", "all ok" if ok else "FAILED")
sys.exit(0 if ok else 1)
