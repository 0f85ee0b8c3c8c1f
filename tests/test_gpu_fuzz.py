"""Differential fuzz of Canvas.points against the C oracle: random canvas sizes, ranges (including ranges far from the
origin, where the float32 fast mapping must hand over to the exact one), coordinate dtypes, reductions and NaN patterns,
with every specialised kernel forced on at small n (K2 privatised count, K1 mono, 16-bit packed count; every fourth seed the
routed path - float32 coordinates only, it declines the rest and the call falls back)."""
import os

import numpy as np
import pytest

from helpers import assert_agg_equal, make_agg

pytestmark = pytest.mark.gpu

SPECS = [("count",), ("count", "v32"), ("any",), ("sum", "v32"), ("mean", "v32"), ("max", "v32"), ("min", "v32"),
         ("first", "v32"), ("last", "v32"), ("where", ("max", "v32"), "other"), ("where", ("min", "v32"), None),
         ("max", "vi"), ("mean", "v64"), ("by", "cat", ("count",)), ("by", "cat", ("mean", "v32")),
         ("max", "v64"), ("min", "v64"), ("first", "v64"), ("last", "v64")]


@pytest.fixture()
def forced(monkeypatch):
    import datashader_b200 as ds
    lib = ds._lib.lib()
    monkeypatch.setattr(ds.config, "priv_min_rows", 0)
    monkeypatch.setattr(ds.config, "count16_min_rows", 0)
    ds._lib.check(lib.dsb_configure(b"mono_min_rows", 0), "cfg")
    ds._lib.check(lib.dsb_configure(b"band_min_rows", 0), "cfg")
    old = (ds.config.routed_min_rows, ds.config.l2_budget_bytes, ds.config.priv_count, ds.config.count16)
    yield ds
    ds.config.routed_min_rows, ds.config.l2_budget_bytes, ds.config.priv_count, ds.config.count16 = old
    ds._lib.check(lib.dsb_routed_configure(1 << 24), "cfg")
    ds._lib.check(lib.dsb_configure(b"mono_min_rows", 1 << 20), "cfg")
    ds._lib.check(lib.dsb_configure(b"band_min_rows", 1 << 22), "cfg")
    ds._lib.check(lib.dsb_configure(b"l2_band_bytes", 96 << 20), "cfg")


@pytest.mark.parametrize("seed", range(int(os.environ.get("DSB_FUZZ_SEEDS", "24"))))     # soak: DSB_FUZZ_SEEDS=400
def test_points_fuzz_vs_oracle(forced, monkeypatch, seed):
    import torch
    from oracle import oracle as ora
    ds = forced
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(20_000, 60_000))
    W, H = int(rng.integers(1, 400)), int(rng.integers(1, 300))
    centre = float(rng.choice([0.0, 0.5, -3.0, 1e3, 2.5e5, -7e6]))      # big offsets: the fast mapping's error bound grows
    span = float(rng.choice([1.0, 0.01, 37.5, 1e4]))
    xdt = np.float32 if seed % 3 else np.float64
    x = (centre + span * (rng.random(n) * 1.3 - 0.15)).astype(xdt)
    y = (centre + span * (rng.random(n) * 1.3 - 0.15)).astype(xdt)
    xr = (float(np.float64(centre)), float(np.float64(centre + span)))
    yr = (float(np.float64(centre + 0.1 * span)), float(np.float64(centre + 0.9 * span)))
    # points exactly on pixel edges and on the range bounds
    k = min(n, 500)
    x[:k] = (xr[0] + (xr[1] - xr[0]) * rng.integers(0, W + 1, k) / W).astype(xdt)
    y[k:2 * k] = (yr[0] + (yr[1] - yr[0]) * rng.integers(0, H + 1, k) / H).astype(xdt)
    x[2 * k:2 * k + 5] = np.nan
    ncat = int(rng.integers(2, 6))
    cols = {"x": x, "y": y, "v32": np.round(rng.standard_normal(n), 1).astype(np.float32), "other": rng.random(n).astype(np.float32),
            "vi": rng.integers(-9, 9, n).astype(np.int32), "v64": np.round(rng.standard_normal(n), 1),
            "cat": rng.integers(0, ncat, n).astype(np.int8), "cat__ncat": ncat}
    cols["v32"][rng.integers(0, n, n // 50)] = np.nan
    view = ora.make_view(W, H, xr, yr)
    cvs = ds.Canvas(W, H, x_range=xr, y_range=yr)
    row_offset = 0
    if seed % 5 == 4:        # every fifth seed: a shard whose global rows cross 2^32, walked in slices of a few thousand rows
        row_offset = (1 << 32) - int(rng.integers(1, n))
        monkeypatch.setattr(ds.DeviceFrame, "CHUNK_ROWS", int(rng.integers(5_000, 20_000)))
    frame = ds.DeviceFrame({k_: torch.from_numpy(v).cuda() for k_, v in cols.items() if k_ != "cat__ncat"},
                           categories={"cat": [f"c{i}" for i in range(ncat)]}, row_offset=row_offset)
    picks = [SPECS[i] for i in rng.choice(len(SPECS), 8, replace=False)]
    if seed % 4 == 2:        # every fourth seed: single-accumulator plans take the routed path (bin, then accumulate in shared memory)
        ds.config.routed_min_rows, ds.config.l2_budget_bytes = 0, 1
        ds.config.priv_count, ds.config.count16 = False, False
        ds._lib.check(ds._lib.lib().dsb_routed_configure(0), "cfg")
    if seed % 4 == 3:        # every fourth seed: tiny L2 bands, so the mono / generic kernels run their BANDED forms
        ds._lib.check(ds._lib.lib().dsb_configure(b"l2_band_bytes", int(rng.choice([4096, 32768]))), "cfg")
    for spec in picks:
        want = ora.points(cols, "x", "y", spec, view, npartitions=2 if ("first" in str(spec) or "last" in str(spec)) else 1)
        got = cvs.points(frame, "x", "y", make_agg(spec)).data
        if spec[0] == "where" and spec[2] is None:      # row ids are global: back to rows of this frame
            got = np.where(got >= 0, got - row_offset, got)
        # values are multiples of 0.1 that may cancel exactly in the reference's order: 1e-13 absolute on sums / means
        assert_agg_equal(got, want, f"seed {seed} {W}x{H} centre {centre} span {span} {xdt.__name__} {spec}", atol=1e-13)


@pytest.mark.parametrize("seed", range(8))
def test_lines_and_areas_fuzz_vs_oracle(seed):
    """Random LinesAxis1 frames (float32 / float64, NaN breaks, clipping canvases, ranges away from the origin): Bresenham
    bit-exact, antialiased single-stage and 2-stage reductions to 1e-6; areas (to zero / to line) bit-exact."""
    import pandas as pd
    import datashader_b200 as ds
    from oracle import oracle as ora
    rng = np.random.default_rng(500 + seed)
    nl, nv = int(rng.integers(30, 400)), int(rng.integers(2, 20))
    dt = np.float32 if seed % 2 else np.float64
    centre = float(rng.choice([0.0, 10.0, -250.0]))
    xs = (centre + np.cumsum(rng.normal(0, 0.06, (nl, nv)), axis=1) + rng.random((nl, 1))).astype(dt)
    ys = (centre + np.cumsum(rng.normal(0, 0.06, (nl, nv)), axis=1) + rng.random((nl, 1))).astype(dt)
    ys2 = (ys - np.abs(rng.normal(0.1, 0.05, (nl, nv)))).astype(dt)
    xs[rng.random((nl, nv)) < 0.03] = np.nan
    val = (rng.random(nl) * 6 - 2).astype(np.float32)
    val[rng.integers(0, nl, max(1, nl // 20))] = np.nan
    d = {f"x{j}": xs[:, j] for j in range(nv)}
    d.update({f"y{j}": ys[:, j] for j in range(nv)})
    d.update({f"s{j}": ys2[:, j] for j in range(nv)})
    d["val"] = val
    df = pd.DataFrame(d)
    xc, yc, sc = ([f"{p}{j}" for j in range(nv)] for p in "xys")
    W, H = int(rng.integers(20, 300)), int(rng.integers(20, 200))
    xr, yr = (centre + 0.1, centre + 0.9), (centre + 0.2, centre + 1.1)
    cvs = ds.Canvas(W, H, x_range=xr, y_range=yr)
    view = ora.make_view(W, H, xr, yr)
    key = f"seed {seed} {nl}x{nv} {dt.__name__} {W}x{H} centre {centre}"
    for name, agg in (("any", ds.any()), ("count", ds.count()), ("max", ds.max("val")), ("min", ds.min("val"))):
        got = cvs.line(df, x=xc, y=yc, axis=1, agg=agg).data
        want = ora.lines(xs, ys, view, name, None if name in ("any", "count") else val, 0)
        assert got.dtype == want.dtype and np.array_equal(got, want, equal_nan=got.dtype.kind == "f"), f"{key} bresenham {name}"
    lw = float(rng.choice([0.5, 1.0, 2.0, 3.5]))
    for name, agg in (("any", ds.any()), ("max", ds.max("val")), ("sum", ds.sum("val")), ("mean", ds.mean("val"))):
        got = cvs.line(df, x=xc, y=yc, axis=1, agg=agg, line_width=lw).data
        want = ora.lines(xs, ys, view, name, None if name == "any" else val, lw)
        assert got.dtype == want.dtype and np.array_equal(np.isnan(got), np.isnan(want)), f"{key} aa {name} lw {lw}"
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6, equal_nan=True, err_msg=f"{key} aa {name} lw {lw}")
    for name, agg in (("min", ds.min("val")), ("first", ds.first("val")), ("last", ds.last("val")),
                      ("sum", ds.sum("val", self_intersect=False))):
        got = cvs.line(df, x=xc, y=yc, axis=1, agg=agg, line_width=lw).data
        want = ora.lines_aa2(xs, ys, view, name, val, lw)
        assert np.array_equal(np.isnan(got), np.isnan(want)), f"{key} aa2 {name} lw {lw}"
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6, equal_nan=True, err_msg=f"{key} aa2 {name} lw {lw}")
    for name, agg in (("any", ds.any()), ("count", ds.count()), ("max", ds.max("val"))):
        vals = None if name in ("any", "count") else val
        got = cvs.area(df, x=xc, y=yc, agg=agg, axis=1).data
        assert np.array_equal(got, ora.areas(xs, ys, view, None, name, vals), equal_nan=got.dtype.kind == "f"), f"{key} area zero {name}"
        got = cvs.area(df, x=xc, y=yc, y_stack=sc, agg=agg, axis=1).data
        assert np.array_equal(got, ora.areas(xs, ys, view, ys2, name, vals), equal_nan=got.dtype.kind == "f"), f"{key} area line {name}"


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 9, 15, 17, 1023, 1025])
def test_tiny_inputs_through_the_specialised_kernels(forced, n):
    """Row counts around the vector widths (float4 / double2 loads, 8-point batches, tail rows) with every specialised
    kernel forced on."""
    import torch
    from oracle import oracle as ora
    ds = forced
    rng = np.random.default_rng(n)
    W, H = 37, 23
    view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    for dt in (np.float32, np.float64):
        cols = {"x": rng.random(n).astype(dt), "y": rng.random(n).astype(dt), "v32": rng.standard_normal(n).astype(np.float32),
                "v64": rng.standard_normal(n)}
        if n > 2:
            cols["v32"][1] = np.nan
            cols["v64"][1] = np.nan
        frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items()})
        for spec in (("count",), ("any",), ("mean", "v32"), ("mean", "v64"), ("sum", "v64"), ("max", "v32"), ("first", "v32"),
                     ("last", "v32"), ("where", ("max", "v32"), None), ("max", "v64"), ("min", "v64"), ("first", "v64"),
                     ("last", "v64")):
            want = ora.points(cols, "x", "y", spec, view, npartitions=2 if spec[0] in ("first", "last") else 1)
            assert_agg_equal(cvs.points(frame, "x", "y", make_agg(spec)).data, want, f"n={n} {dt.__name__} {spec}", atol=1e-13)
