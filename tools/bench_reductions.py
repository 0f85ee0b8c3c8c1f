#!/usr/bin/env python
"""Every points reduction at the headline geometry (900x525, 1e9 f32 points resident in HBM): ms per pass and Gpts/s.
    python tools/bench_reductions.py [--n 1000000000]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import datashader_b200 as ds  # noqa: E402
from datashader_b200 import config  # noqa: E402


def timed(fn, warmup=2, steps=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000_000)
    ap.add_argument("--width", type=int, default=900)
    ap.add_argument("--height", type=int, default=525)
    ap.add_argument("--only", default="", help="comma-separated subset of the reductions")
    ap.add_argument("--configure", action="append", default=[], help="key=value passed to dsb_configure")
    a = ap.parse_args()
    n = a.n
    from datashader_b200 import _lib
    for kv in a.configure:
        k, v = kv.split("=")
        _lib.check(_lib.lib().dsb_configure(k.encode(), int(v)), "dsb_configure")
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    x = torch.rand(n, generator=g, device="cuda"); y = torch.rand(n, generator=g, device="cuda")
    v = torch.randn(n, generator=g, device="cuda")
    cat = torch.randint(0, 16, (n,), generator=g, device="cuda", dtype=torch.int8)
    frame = ds.DeviceFrame({"x": x, "y": y, "value": v, "cat": cat}, categories={"cat": [f"c{i}" for i in range(16)]})
    cvs = ds.Canvas(a.width, a.height, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    config.device_results = True
    aggs = {"count": ds.count(), "any": ds.any(), "sum": ds.sum("value"), "mean": ds.mean("value"), "max": ds.max("value"),
            "min": ds.min("value"), "first": ds.first("value"), "last": ds.last("value"),
            "where_max": ds.where(ds.max("value"), "x"), "where_max_row": ds.where(ds.max("value")),
            "by_count": ds.by("cat", ds.count()), "by_mean": ds.by("cat", ds.mean("value")),
            "summary3": ds.summary(c=ds.count(), m=ds.mean("value"), mx=ds.max("value"))}
    out = {}
    if a.only:
        aggs = {k: v for k, v in aggs.items() if k in a.only.split(",")}
    for name, agg in aggs.items():
        ms = timed(lambda: cvs.points(frame, "x", "y", agg))
        out[name] = {"ms": round(ms, 3), "gpts": round(n / ms / 1e6, 1)}
        print(name, out[name], flush=True)
    print(json.dumps({"n": n, "canvas": [a.width, a.height], "results": out}))


if __name__ == "__main__":
    main()
