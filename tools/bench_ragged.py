#!/usr/bin/env python
"""LinesAxis1Ragged at BASELINE config 4's geometry (100k lines x 1000 vertices, 3840 x 2160) against the dense LinesAxis1 layout of the
same vertices, and with uneven row lengths of the same total:  python tools/bench_ragged.py [nlines]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import datashader_b200 as ds
from datashader_b200 import config

nl, nv = (int(sys.argv[1]) if len(sys.argv) > 1 else 100_000), 1000
g = torch.Generator(device="cuda")
g.manual_seed(4)
xs = torch.arange(nv, device="cuda", dtype=torch.float32).repeat(nl, 1)
ys = torch.randn(nl, nv, generator=g, device="cuda").cumsum(dim=1)
val = torch.rand(nl, generator=g, device="cuda")
cvs = ds.Canvas(3840, 2160, x_range=(0.0, float(nv - 1)), y_range=(float(ys.min()), float(ys.max())))
config.device_results = True


def timed(fn, warmup=2, steps=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, r


cols = {f"x{j}": xs[:, j].contiguous() for j in range(nv)}
cols.update({f"y{j}": ys[:, j].contiguous() for j in range(nv)})
cols["value"] = val
dense = ds.DeviceFrame(cols)
xc, yc = [f"x{j}" for j in range(nv)], [f"y{j}" for j in range(nv)]
starts = torch.arange(nl, device="cuda", dtype=torch.int64) * nv
even = ds.DeviceFrame({"x": ds.RaggedColumn(xs.reshape(-1), starts), "y": ds.RaggedColumn(ys.reshape(-1), starts), "value": val})
# uneven rows of the same total: lengths 2 .. 2 nv - 2
lens = torch.randint(2, 2 * nv - 1, (nl,), generator=g, device="cuda")
lens = (lens.double() * (nl * nv / lens.sum().double())).long().clamp(min=2)
ustarts = torch.cumsum(lens, 0) - lens
total = int(lens.sum())
flat_x = (torch.arange(total, device="cuda") - torch.repeat_interleave(ustarts, lens)).float() * (float(nv - 1) / (2 * nv))
flat_y = torch.randn(total, generator=g, device="cuda").cumsum(0)
flat_y = flat_y - torch.repeat_interleave(flat_y[ustarts], lens)
uneven = ds.DeviceFrame({"x": ds.RaggedColumn(flat_x, ustarts), "y": ds.RaggedColumn(flat_y, ustarts), "value": val})
out = {"lines": nl, "segments_dense": nl * (nv - 1), "vertices_uneven": total}
for lw, tag in ((0, "bresenham"), (1, "aa")):
    for name, agg in (("max", ds.max("value")),) + ((("first", ds.first("value")),) if lw else ()):
        d, rd_ = timed(lambda: cvs.line(dense, x=xc, y=yc, axis=1, agg=agg, line_width=lw))
        e, re_ = timed(lambda: cvs.line(even, "x", "y", axis=1, agg=agg, line_width=lw))
        u, _ = timed(lambda: cvs.line(uneven, "x", "y", axis=1, agg=agg, line_width=lw))
        a, b = torch.as_tensor(rd_.data), torch.as_tensor(re_.data)
        same = bool(torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0)))
        out[f"{tag}_{name}"] = {"dense_ms": d, "ragged_even_ms": e, "ragged_uneven_ms": u, "ragged_equals_dense": same}
print(json.dumps(out))
