#!/usr/bin/env python
"""Secondary measurements: BASELINE.json configs 3, 4 and 5 (scaled to one GPU where noted) on the CUDA path,
each checked against a size-independent property.  Prints one JSON object per config.  Not the driver's
bench (that is bench.py, config 2); results are copied into DESIGN.md / profiles/.

    python tools/bench_configs.py [--scale 1.0] [--configs 3,4,5]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import datashader_b200 as ds  # noqa: E402
from datashader_b200 import config  # noqa: E402


def timed(fn, warmup=2, steps=5):
    for _ in range(warmup):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def config3(scale):
    """Canvas(1920x1080).points 1e9 points, by('cat', count()) 16 categories, then tf.shade(how='eq_hist')."""
    n = int(1e9 * scale)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    x = torch.rand(n, generator=g, device="cuda"); y = torch.rand(n, generator=g, device="cuda")
    cat = torch.randint(0, 16, (n,), generator=g, device="cuda", dtype=torch.int8)
    frame = ds.DeviceFrame({"x": x, "y": y, "cat": cat}, categories={"cat": [f"c{i}" for i in range(16)]})
    cvs = ds.Canvas(1920, 1080, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    config.device_results = True
    ms_agg, agg = timed(lambda: cvs.points(frame, "x", "y", ds.by("cat", ds.count())))
    total = int(agg.data.view(torch.int32).to(torch.int64).sum().item())
    ms_shade, img = timed(lambda: ds.tf.shade(agg, how="eq_hist"))
    config.device_results = False
    return {"config": 3, "points": n, "agg_ms": ms_agg, "agg_gpts": n / ms_agg / 1e6, "agg_hbm_gbs_algorithmic": n * 9 / ms_agg / 1e6,
            "shade_ms": ms_shade, "shade_gpix": 1920 * 1080 / ms_shade / 1e6, "check_total_count": total == n,
            "nonzero_alpha_pixels": int((img.data >> 24 > 0).sum())}


def config4(scale):
    """Canvas(3840x2160).line LinesAxis1, 100k lines x 1000 samples, antialiased, agg=max('value')."""
    nl, nv = int(100_000 * scale), 1000
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    xs = torch.arange(nv, device="cuda", dtype=torch.float32).repeat(nl, 1)
    ys = torch.randn(nl, nv, generator=g, device="cuda").cumsum(dim=1)
    val = torch.rand(nl, generator=g, device="cuda")
    cols = {f"x{j}": xs[:, j].contiguous() for j in range(nv)}
    cols.update({f"y{j}": ys[:, j].contiguous() for j in range(nv)})
    cols["value"] = val
    frame = ds.DeviceFrame(cols)
    xr, yr = (0.0, float(nv - 1)), (float(ys.min()), float(ys.max()))
    cvs = ds.Canvas(3840, 2160, x_range=xr, y_range=yr)
    xc, yc = [f"x{j}" for j in range(nv)], [f"y{j}" for j in range(nv)]
    out = {"config": 4, "lines": nl, "segments": nl * (nv - 1)}
    for lw, tag in ((1, "aa"), (0, "bresenham")):
        config.device_results = True        # the aggregate stays on the device (what config 2's `value` measures)
        ms, agg = timed(lambda: cvs.line(frame, x=xc, y=yc, axis=1, agg=ds.max("value"), line_width=lw), warmup=2, steps=3)
        out[f"{tag}_ms"] = ms
        out[f"{tag}_msegments_per_s"] = nl * (nv - 1) / ms / 1e3
        out[f"{tag}_covered_pixels"] = int((~torch.isnan(torch.as_tensor(agg.data))).sum())
        config.device_results = False       # ... and with the 66 MB f64 aggregate copied to a numpy array
        ms, agg = timed(lambda: cvs.line(frame, x=xc, y=yc, axis=1, agg=ds.max("value"), line_width=lw), warmup=2, steps=3)
        out[f"{tag}_to_host_ms"] = ms
    config.device_results = True
    # the 2-stage antialiased reductions (one CTA per line, per-line stage-1 canvases in scratch memory)
    for name, a2 in (("aa2_min", ds.min("value")), ("aa2_first", ds.first("value")), ("aa2_sum_nsi", ds.sum("value", self_intersect=False))):
        ms, agg = timed(lambda: cvs.line(frame, x=xc, y=yc, axis=1, agg=a2, line_width=1), warmup=1, steps=2)
        out[f"{name}_ms"] = ms
        out[f"{name}_msegments_per_s"] = nl * (nv - 1) / ms / 1e3
    config.device_results = False
    return out


def config5(scale):
    """Canvas(8192x8192).points 4e9 points (scaled), agg=max('value') and where(first('value')) - canvas beyond L2."""
    n = int(4e9 * scale)
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    x = torch.rand(n, generator=g, device="cuda"); y = torch.rand(n, generator=g, device="cuda")
    v = torch.randn(n, generator=g, device="cuda")
    frame = ds.DeviceFrame({"x": x, "y": y, "value": v})
    cvs = ds.Canvas(8192, 8192, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    config.device_results = True
    out = {"config": 5, "points": n}
    for name, agg in (("max", ds.max("value")), ("first", ds.first("value")), ("where_max_row", ds.where(ds.max("value"))),
                      ("count", ds.count())):
        ms, res = timed(lambda: cvs.points(frame, "x", "y", agg), warmup=3, steps=5)   # the first passes run at ramping clocks
        out[f"{name}_ms"] = ms
        out[f"{name}_gpts"] = n / ms / 1e6
    config.device_results = False
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--configs", default="3,4,5")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    t0 = time.time()
    for c in a.configs.split(","):
        fn = {"3": config3, "4": config4, "5": config5}[c.strip()]
        # config 5 at full size needs 48 GB of columns: run it at a quarter (1e9 points) unless told otherwise
        s = a.scale * (0.25 if c.strip() == "5" else 1.0)
        r = fn(s)
        r["scale"] = s
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
    print(json.dumps({"wall_s": time.time() - t0}))


if __name__ == "__main__":
    main()
