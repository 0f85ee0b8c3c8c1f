"""Generate golden vectors from the REAL reference (holoviz/datashader at /root/reference).

Run in the build container only (the reference is not present on the GPU box):

    python tests/golden/make_golden.py

The reference is pure Python + numba; it is imported unmodified.  `xarray`, `toolz` and
`multipledispatch` are missing from this image, so tests/golden/_shims/ provides container-only
stand-ins for them (containers and functional helpers only - no arithmetic).  Outputs are small
.npz fixtures (inputs + reference outputs) committed under tests/golden/; tests/ compare both the
C oracle and the CUDA path against them.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shims"))
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402

import datashader as ds  # noqa: E402
from datashader.compiler import compile_components  # noqa: E402
from datashader.glyphs import Point  # noqa: E402
from datashader.utils import dshape_from_pandas  # noqa: E402

NCAT = 5


def make_frame(seed, n, xy_dtype):
    rng = np.random.default_rng(seed)
    x = rng.random(n).astype(xy_dtype) * 1.3 - 0.15      # some rows fall outside (0,1)
    y = rng.random(n).astype(xy_dtype) * 1.3 - 0.15
    # plant exact edge / pixel-boundary coordinates
    edge = np.array([0.0, 1.0, 0.5, 0.25, 0.75, np.nextafter(np.float32(1), np.float32(0)),
                     np.nextafter(np.float32(0.5), np.float32(0)), 1.0 / 3, 2.0 / 3], dtype=xy_dtype)
    k = len(edge)
    x[:k] = edge
    y[:k] = edge[::-1]
    x[k:2 * k] = edge
    y[k:2 * k] = 0.5
    x[rng.integers(0, n, n // 50)] = np.nan
    y[rng.integers(0, n, n // 50)] = np.nan
    v32 = rng.standard_normal(n).astype(np.float32)
    v32[rng.integers(0, n, n // 20)] = np.nan
    v32[rng.integers(0, n, n // 40)] = np.round(v32[rng.integers(0, n, n // 40)], 1)   # ties
    v64 = rng.standard_normal(n) * 1e3
    v64[rng.integers(0, n, n // 20)] = np.nan
    vi = rng.integers(-50, 50, n).astype(np.int32)
    other = rng.random(n).astype(np.float32) * 100
    other[rng.integers(0, n, n // 30)] = np.nan
    codes = rng.integers(0, NCAT, n).astype(np.int8)
    return dict(x=x, y=y, v32=v32, v64=v64, vi=vi, other=other, cat=codes)


def to_df(cols):
    d = {k: v for k, v in cols.items() if k != "cat"}
    d["cat"] = pd.Categorical.from_codes(cols["cat"], categories=[f"c{i}" for i in range(NCAT)])
    return pd.DataFrame(d)


def reductions():
    return {
        "count": ds.count(), "count_v32": ds.count("v32"), "any": ds.any(), "any_v32": ds.any("v32"),
        "sum_v32": ds.sum("v32"), "sum_v64": ds.sum("v64"), "sum_vi": ds.sum("vi"),
        "mean_v32": ds.mean("v32"), "mean_v64": ds.mean("v64"),
        "min_v32": ds.min("v32"), "max_v32": ds.max("v32"), "min_v64": ds.min("v64"), "max_v64": ds.max("v64"),
        "max_vi": ds.max("vi"),
        "first_v32": ds.first("v32"), "last_v32": ds.last("v32"),
        "where_max_v32_other": ds.where(ds.max("v32"), "other"), "where_min_v32_other": ds.where(ds.min("v32"), "other"),
        "where_max_v32_row": ds.where(ds.max("v32")), "where_min_v32_row": ds.where(ds.min("v32")),
        "where_first_v32_other": ds.where(ds.first("v32"), "other"), "where_last_v32_other": ds.where(ds.last("v32"), "other"),
        "where_first_v32_row": ds.where(ds.first("v32")), "where_last_v32_row": ds.where(ds.last("v32")),
        "where_max_vi_other": ds.where(ds.max("vi"), "other"),
        "by_count": ds.by("cat", ds.count()), "by_count_v32": ds.by("cat", ds.count("v32")),
        "by_sum_v32": ds.by("cat", ds.sum("v32")), "by_mean_v32": ds.by("cat", ds.mean("v32")),
        "by_max_v32": ds.by("cat", ds.max("v32")), "by_min_v32": ds.by("cat", ds.min("v32")),
        "by_any": ds.by("cat", ds.any()),
    }


def points_cases():
    out = {}
    canvases = {
        "c2x2": dict(plot_width=2, plot_height=2, x_range=(0, 1), y_range=(0, 1)),
        "c37x23": dict(plot_width=37, plot_height=23, x_range=(-0.1, 1.05), y_range=(0.1, 0.9)),
        "c90x52": dict(plot_width=90, plot_height=52, x_range=(0, 1), y_range=(0, 1)),
        "cauto": dict(plot_width=31, plot_height=17),
    }
    for xy_dtype in (np.float32, np.float64):
        tag = "f32" if xy_dtype == np.float32 else "f64"
        cols = make_frame(1234 if tag == "f32" else 4321, 6000, xy_dtype)
        df = to_df(cols)
        for k, v in cols.items():
            out[f"in_{tag}_{k}"] = v
        for cname, ckw in canvases.items():
            cvs = ds.Canvas(**ckw)
            for rname, red in reductions().items():
                if tag == "f64" and not (rname in ("count", "mean_v32", "max_v32", "by_count", "where_max_v32_other")):
                    continue
                agg = cvs.points(df, "x", "y", red)
                out[f"pts_{tag}_{cname}_{rname}"] = np.asarray(agg.data)
                if rname == "count":
                    out[f"pts_{tag}_{cname}_xcoords"] = np.asarray(agg.coords["x"])
                    out[f"pts_{tag}_{cname}_ycoords"] = np.asarray(agg.coords["y"])
                    out[f"pts_{tag}_{cname}_xrange"] = np.asarray(agg.attrs["x_range"], dtype=np.float64)
                    out[f"pts_{tag}_{cname}_yrange"] = np.asarray(agg.attrs["y_range"], dtype=np.float64)
    # log axes
    cols = make_frame(99, 4000, np.float32)
    cols["x"] = (10 ** (cols["x"].astype(np.float64) * 3)).astype(np.float32)
    cols["y"] = (10 ** (cols["y"].astype(np.float64) * 2)).astype(np.float32)
    df = to_df(cols)
    for k in ("x", "y", "v32"):
        out[f"in_log_{k}"] = cols[k]
    cvs = ds.Canvas(plot_width=40, plot_height=30, x_range=(1, 1000), y_range=(1, 100), x_axis_type="log", y_axis_type="log")
    out["pts_log_count"] = np.asarray(cvs.points(df, "x", "y", ds.count()).data)
    out["pts_log_max_v32"] = np.asarray(cvs.points(df, "x", "y", ds.max("v32")).data)
    a = cvs.points(df, "x", "y", ds.count())
    out["pts_log_xcoords"] = np.asarray(a.coords["x"])
    out["pts_log_ycoords"] = np.asarray(a.coords["y"])
    return out


def partitioned_cases():
    """The reference's dask path without dask: compile_components(partitioned=True), create+extend per
    row slice with _datashader_row_offset set as data_libraries/dask.py:110-113 does, combine, finalize."""
    out = {}
    cols = make_frame(777, 5000, np.float32)
    df = to_df(cols)
    for k, v in cols.items():
        out[f"in_{k}"] = v
    cvs = ds.Canvas(plot_width=37, plot_height=23, x_range=(-0.1, 1.05), y_range=(0.1, 0.9))
    glyph = Point("x", "y")
    schema = dshape_from_pandas(df)
    x_range, y_range = cvs.x_range, cvs.y_range
    x_st = cvs.x_axis.compute_scale_and_translate(x_range, cvs.plot_width)
    y_st = cvs.y_axis.compute_scale_and_translate(y_range, cvs.plot_height)
    names = ["count", "mean_v32", "sum_v32", "max_v32", "min_v32", "first_v32", "last_v32",
             "where_max_v32_other", "where_min_v32_row", "where_first_v32_other", "where_last_v32_row",
             "by_count", "by_max_v32", "any_v32"]
    reds = reductions()
    nparts = 3
    for rname in names:
        red = reds[rname]
        create, info, append, combine, finalize, aa2, aa2f, _ = compile_components(
            red, schema, glyph, antialias=False, cuda=False, partitioned=True)
        extend = glyph._build_extend(cvs.x_axis.mapper, cvs.y_axis.mapper, info, append, aa2, aa2f)
        parts = []
        n = len(df)
        for p in range(nparts):
            lo, hi = n * p // nparts, n * (p + 1) // nparts
            part = df.iloc[lo:hi]
            object.__setattr__(part, "_datashader_row_offset", lo)
            aggs = create((cvs.plot_height, cvs.plot_width))
            extend(aggs, part, x_st + y_st, x_range + y_range)
            parts.append(aggs)
        res = finalize(combine(parts), cuda=False, coords={"x": None, "y": None}, dims=["y", "x"], attrs={})
        out[f"part3_{rname}"] = np.asarray(res.data)
    return out


def line_frame(seed, nlines, nverts, dtype):
    rng = np.random.default_rng(seed)
    xs = np.tile(np.linspace(-0.2, 1.2, nverts), (nlines, 1)) + rng.normal(0, 0.02, (nlines, nverts))
    ys = np.cumsum(rng.normal(0, 0.08, (nlines, nverts)), axis=1) + 0.5
    xs[rng.random((nlines, nverts)) < 0.03] = np.nan
    ys[rng.random((nlines, nverts)) < 0.03] = np.nan
    # a few degenerate / vertical / horizontal / repeated-vertex segments
    xs[0, :] = 0.5
    ys[1, :] = 0.5
    xs[2, 3] = xs[2, 2]
    ys[2, 3] = ys[2, 2]
    xs[3, :] = np.linspace(0.0, 1.0, nverts)
    ys[3, :] = np.linspace(0.0, 1.0, nverts)     # exact diagonal through pixel corners
    val = rng.random(nlines) * 10 - 3
    val[5] = np.nan
    return xs.astype(dtype), ys.astype(dtype), val.astype(np.float32)


def lines_cases():
    out = {}
    for tag, dtype in (("f32", np.float32), ("f64", np.float64)):
        xs, ys, val = line_frame(2024, 40, 24, dtype)
        nverts = xs.shape[1]
        out[f"in_{tag}_xs"], out[f"in_{tag}_ys"], out[f"in_{tag}_val"] = xs, ys, val
        d = {f"x{j}": xs[:, j] for j in range(nverts)}
        d.update({f"y{j}": ys[:, j] for j in range(nverts)})
        d["val"] = val
        df = pd.DataFrame(d)
        xcols, ycols = [f"x{j}" for j in range(nverts)], [f"y{j}" for j in range(nverts)]
        canvases = {
            "c64x48": dict(plot_width=64, plot_height=48, x_range=(0, 1), y_range=(0, 1)),
            "c33x57": dict(plot_width=33, plot_height=57, x_range=(-0.3, 1.3), y_range=(-0.5, 1.5)),
        }
        for cname, ckw in canvases.items():
            cvs = ds.Canvas(**ckw)
            for lw in (0, 1, 2.5, 0.5):
                aggs = {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val")}
                if lw == 0:
                    aggs["min"] = ds.min("val")
                if tag == "f64" and lw not in (0, 1):
                    continue
                for aname, agg in aggs.items():
                    r = cvs.line(df, x=xcols, y=ycols, axis=1, agg=agg, line_width=lw)
                    out[f"ln_{tag}_{cname}_lw{lw}_{aname}"] = np.asarray(r.data)
    return out


def shade_cases():
    """tf.shade on categorical [H,W,C] count canvases and on 2-D canvases (transfer_functions/__init__.py)."""
    import xarray as xr
    import datashader.transfer_functions as tf
    out = {}
    rng = np.random.default_rng(31)
    H, Wd = 48, 64

    def cat_agg(data):
        C = data.shape[2]
        return xr.DataArray(data, coords={"y": np.arange(H), "x": np.arange(Wd), "cat": [f"c{i}" for i in range(C)]},
                            dims=["y", "x", "cat"])

    def agg2(data):
        return xr.DataArray(data, coords={"y": np.arange(H), "x": np.arange(Wd)}, dims=["y", "x"])

    cats = {}
    a = rng.poisson(2.0, (H, Wd, 5)).astype(np.uint32)
    a[rng.random((H, Wd)) < 0.3] = 0                       # empty pixels
    cats["poisson5"] = a
    b = (rng.pareto(1.2, (H, Wd, 16)) * 40).astype(np.uint32)   # heavy tail: totals far apart, > 65536 levels
    b[rng.random((H, Wd)) < 0.2] = 0
    b[0, 0, :] = 3_000_000
    cats["pareto16"] = b
    c = rng.integers(1, 50, (H, Wd, 3)).astype(np.uint32)  # no empty pixel, baseline > 0
    cats["dense3"] = c
    d = np.zeros((H, Wd, 4), np.uint32)
    d[5, 7, 2] = 9                                           # single non-empty pixel
    cats["single4"] = d
    for name, data in cats.items():
        out[f"cat_{name}_in"] = data
        for how in ("eq_hist", "log", "cbrt", "linear"):
            out[f"cat_{name}_{how}"] = np.asarray(tf.shade(cat_agg(data), how=how).data)
        out[f"cat_{name}_eq_hist_a200_m10"] = np.asarray(tf.shade(cat_agg(data), how="eq_hist", alpha=200, min_alpha=10).data)
        out[f"cat_{name}_eq_hist_rescale"] = np.asarray(
            tf.shade(cat_agg(data), how="eq_hist", rescale_discrete_levels=True).data)

    twod = {
        "u32": cats["poisson5"].sum(axis=2).astype(np.uint32),
        "u32big": cats["pareto16"].sum(axis=2).astype(np.uint32),
        "f64": np.where(rng.random((H, Wd)) < 0.25, np.nan, rng.standard_normal((H, Wd)) * 10),
        "f32": np.where(rng.random((H, Wd)) < 0.25, np.nan, rng.random((H, Wd))).astype(np.float32),
    }
    for name, data in twod.items():
        out[f"d2_{name}_in"] = data
        for how in ("eq_hist", "log", "cbrt", "linear"):
            out[f"d2_{name}_{how}_default"] = np.asarray(tf.shade(agg2(data), how=how).data)
            out[f"d2_{name}_{how}_hot"] = np.asarray(
                tf.shade(agg2(data), cmap=["black", "darkred", "red", "orange", "yellow", "white"], how=how).data)
            out[f"d2_{name}_{how}_single"] = np.asarray(tf.shade(agg2(data), cmap="#3070c0", how=how, min_alpha=20).data)
    return out


SPAN_CASES_2D = {"u32": [(2, 9), (0.5, 7.5)], "f64": [(-5.0, 12.5)], "f32": [(0.2, 0.7)]}
SPAN_CASES_CAT = {"poisson5": [(3, 20), (2.5, 15.5)], "dense3": [(40, 90), (30.5, 100.25)]}


def shade_catfloat_cases():
    """tf.shade on categorical aggregates that are not counts: by(cat, mean | max) -> float64 [H, W, C] with NaN for empty
    cells, a float32 variant, and signed integers (_colorize, transfer_functions/__init__.py:382-452)."""
    import xarray as xr
    import datashader.transfer_functions as tf
    out = {}
    rng = np.random.default_rng(58)
    H, Wd = 40, 56

    def cat_agg(data):
        C = data.shape[2]
        return xr.DataArray(data, coords={"y": np.arange(H), "x": np.arange(Wd), "cat": [f"c{i}" for i in range(C)]},
                            dims=["y", "x", "cat"])

    a = rng.standard_normal((H, Wd, 4)) * 3 + 1
    a[rng.random((H, Wd, 4)) < 0.4] = np.nan
    a[rng.random((H, Wd)) < 0.2] = np.nan                  # pixels with no category at all
    b = (rng.random((H, Wd, 3)) * 100).astype(np.float32)
    b[rng.random((H, Wd, 3)) < 0.5] = np.nan
    c = rng.integers(-5, 40, (H, Wd, 5)).astype(np.int32)
    for name, data in (("f64", a), ("f32", b), ("i32", c)):
        out[f"catf_{name}_in"] = data
        for how in ("eq_hist", "log", "cbrt", "linear"):
            out[f"catf_{name}_{how}"] = np.asarray(tf.shade(cat_agg(data), how=how).data)
        out[f"catf_{name}_linear_base"] = np.asarray(tf.shade(cat_agg(data), how="linear", color_baseline=0.5 if name != "i32" else 2).data)
        out[f"catf_{name}_linear_span"] = np.asarray(tf.shade(cat_agg(data), how="linear", span=(0, 20)).data)
    return out


def sqrt_how(d, m):
    return np.where(m, np.nan, np.sqrt(d))


def fake_cmap(x, bytes=True):     # noqa: A002  (a stand-in for a matplotlib colormap: callable(scaled, bytes=True) -> uint8 RGBA)
    v = np.nan_to_num(np.clip(x, 0, 1))
    out = np.empty(x.shape + (4,), dtype=np.uint8)
    out[..., 0] = (v * 255).astype(np.uint8)
    out[..., 1] = 255 - (v * 200).astype(np.uint8)
    out[..., 2] = 64
    out[..., 3] = 255
    return out


def shade_extra_cases():
    """tf.shade arguments that are Python objects: a callable `how`, a callable cmap (matplotlib-style) and a discrete
    colour key on a 2-D aggregate (transfer_functions/__init__.py:218-231, 340-349, 535-612)."""
    import xarray as xr
    import datashader.transfer_functions as tf
    out = {}
    rng = np.random.default_rng(73)
    H, Wd = 36, 44

    def agg2(data):
        return xr.DataArray(data, coords={"y": np.arange(H), "x": np.arange(Wd)}, dims=["y", "x"])

    f = np.where(rng.random((H, Wd)) < 0.3, np.nan, rng.random((H, Wd)) * 50)
    u = rng.poisson(3.0, (H, Wd)).astype(np.uint32)
    out["x_f64_in"], out["x_u32_in"] = f, u
    for name, data in (("f64", f), ("u32", u)):
        out[f"x_{name}_callhow_list"] = np.asarray(tf.shade(agg2(data), cmap=["black", "red", "white"], how=sqrt_how).data)
        out[f"x_{name}_callhow_single"] = np.asarray(tf.shade(agg2(data), cmap="#3070c0", how=sqrt_how, min_alpha=20).data)
        for how in ("linear", "log", "cbrt"):
            out[f"x_{name}_callcmap_{how}"] = np.asarray(tf.shade(agg2(data), cmap=fake_cmap, how=how, alpha=200).data)
    cats = rng.integers(0, 5, (H, Wd)).astype(np.int32)
    out["x_cats_in"] = cats
    key = {1: "red", 2: "#00ff00", 4: (0, 0, 255)}
    out["x_cats_key"] = np.asarray(tf.shade(agg2(cats), color_key=key).data)
    out["x_cats_key_a100"] = np.asarray(tf.shade(agg2(cats), color_key=key, alpha=100).data)
    out["x_cats_key_base"] = np.asarray(tf.shade(agg2(cats), color_key=key, color_baseline=0.25).data)
    return out


def shade_span_cases():
    """tf.shade with an explicit span (clip + fixed normalisation range), on the shade.npz inputs."""
    import xarray as xr
    import datashader.transfer_functions as tf
    g = np.load(os.path.join(HERE, "shade.npz"))
    out = {}
    for name, spans in SPAN_CASES_2D.items():
        data = g[f"d2_{name}_in"]
        H, Wd = data.shape
        for k, span in enumerate(spans):
            for how in ("log", "cbrt", "linear"):
                agg = xr.DataArray(data.copy(), coords={"y": np.arange(H), "x": np.arange(Wd)}, dims=["y", "x"])
                out[f"d2_{name}_s{k}_{how}_default"] = np.asarray(tf.shade(agg, how=how, span=span).data)
                out[f"d2_{name}_s{k}_{how}_single"] = np.asarray(tf.shade(agg, cmap="#3070c0", how=how, span=span, min_alpha=20).data)
    for name, spans in SPAN_CASES_CAT.items():
        data = g[f"cat_{name}_in"]
        H, Wd, C = data.shape
        for k, span in enumerate(spans):
            for how in ("log", "cbrt", "linear"):
                agg = xr.DataArray(data.copy(), coords={"y": np.arange(H), "x": np.arange(Wd), "cat": [f"c{i}" for i in range(C)]},
                                   dims=["y", "x", "cat"])
                out[f"cat_{name}_s{k}_{how}"] = np.asarray(tf.shade(agg, how=how, span=span).data)
    return out


def lines_extra_cases():
    """Bresenham lines with the reductions beyond any/count/sum/max/min (row-index based ones included)."""
    out = {}
    xs, ys, val = line_frame(2024, 40, 24, np.float32)
    rng = np.random.default_rng(77)
    nl, nverts = xs.shape
    other = rng.random(nl).astype(np.float32) * 5
    codes = rng.integers(0, 4, nl).astype(np.int8)
    out["in_other"], out["in_cat"] = other, codes
    d = {f"x{j}": xs[:, j] for j in range(nverts)}
    d.update({f"y{j}": ys[:, j] for j in range(nverts)})
    d["val"], d["other"] = val, other
    d["cat"] = pd.Categorical.from_codes(codes, categories=["a", "b", "c", "d"])
    df = pd.DataFrame(d)
    xcols, ycols = [f"x{j}" for j in range(nverts)], [f"y{j}" for j in range(nverts)]
    cvs = ds.Canvas(plot_width=64, plot_height=48, x_range=(0, 1), y_range=(0, 1))
    aggs = {
        "count_val": ds.count("val"), "mean_val": ds.mean("val"), "first_val": ds.first("val"), "last_val": ds.last("val"),
        "where_max_val_row": ds.where(ds.max("val")), "where_min_val_other": ds.where(ds.min("val"), "other"),
        "where_first_val_other": ds.where(ds.first("val"), "other"), "by_count": ds.by("cat", ds.count()),
        "by_max_val": ds.by("cat", ds.max("val")), "by_any": ds.by("cat", ds.any()),
    }
    for name, agg in aggs.items():
        out[f"lnx_{name}"] = np.asarray(cvs.line(df, x=xcols, y=ycols, axis=1, agg=agg).data)
    return out


def line_layout_cases():
    """The other line layouts: LineAxis0, LineAxis0Multi, LinesAxis1XConstant, LinesAxis1YConstant
    (core.py:408-452), Bresenham and antialiased."""
    out = {}
    rng = np.random.default_rng(404)
    n = 60
    x = np.linspace(-0.1, 1.1, n) + rng.normal(0, 0.01, n)
    y = np.cumsum(rng.normal(0, 0.07, n)) + 0.5
    x2 = x[::-1].copy() + 0.03
    y2 = np.cumsum(rng.normal(0, 0.05, n)) + 0.4
    y[17] = np.nan
    x2[40] = np.nan
    val = rng.random(n).astype(np.float32) * 4 - 1
    val[5] = np.nan
    df0 = pd.DataFrame({"x": x.astype(np.float32), "y": y.astype(np.float32), "x2": x2.astype(np.float32),
                        "y2": y2.astype(np.float32), "val": val})
    for k in df0.columns:
        out[f"ax0_{k}"] = df0[k].to_numpy()
    cvs = ds.Canvas(plot_width=50, plot_height=40, x_range=(0, 1), y_range=(0, 1))
    aggs0 = {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val"), "mean": ds.mean("val"),
             "first": ds.first("val"), "where_max_row": ds.where(ds.max("val"))}
    for name, agg in aggs0.items():
        out[f"ax0_lw0_{name}"] = np.asarray(cvs.line(df0, "x", "y", agg=agg).data)
        out[f"ax0multi_lw0_{name}"] = np.asarray(cvs.line(df0, x=["x", "x2"], y=["y", "y2"], axis=0, agg=agg).data)
    for name in ("any", "max"):
        out[f"ax0_lw1_{name}"] = np.asarray(cvs.line(df0, "x", "y", agg=aggs0[name], line_width=1).data)
        out[f"ax0multi_lw1_{name}"] = np.asarray(
            cvs.line(df0, x=["x", "x2"], y=["y", "y2"], axis=0, agg=aggs0[name], line_width=1).data)

    # constant-x / constant-y: 12 lines x 9 vertices
    nl, nv = 12, 9
    xc = np.linspace(0.0, 1.0, nv)
    ys = rng.random((nl, nv)).astype(np.float32)
    ys[3, 4] = np.nan
    lval = rng.random(nl).astype(np.float32)
    d = {f"y{j}": ys[:, j] for j in range(nv)}
    d["val"] = lval
    df1 = pd.DataFrame(d)
    out["xc_x"], out["xc_ys"], out["xc_val"] = xc, ys, lval
    ycols = [f"y{j}" for j in range(nv)]
    aggs1 = {"any": ds.any(), "count": ds.count(), "max": ds.max("val"), "mean": ds.mean("val")}
    for name, agg in aggs1.items():
        out[f"xconst_lw0_{name}"] = np.asarray(cvs.line(df1, x=xc, y=ycols, axis=1, agg=agg).data)
    out["xconst_lw1_max"] = np.asarray(cvs.line(df1, x=xc, y=ycols, axis=1, agg=ds.max("val"), line_width=1).data)
    d2 = {f"x{j}": ys[:, j] for j in range(nv)}
    d2["val"] = lval
    df2 = pd.DataFrame(d2)
    xcols = [f"x{j}" for j in range(nv)]
    for name, agg in aggs1.items():
        out[f"yconst_lw0_{name}"] = np.asarray(cvs.line(df2, x=xcols, y=xc, axis=1, agg=agg).data)
    out["yconst_lw1_max"] = np.asarray(cvs.line(df2, x=xcols, y=xc, axis=1, agg=ds.max("val"), line_width=1).data)
    return out


def lines_aa2_cases():
    """Antialiased lines whose reduction needs the 2-stage combine (antialias.py:30-58; compiler.py:198-268;
    line.py:1291-1319): min, first, last and count/sum with self_intersect=False - LinesAxis1 (40 lines, the
    lines.npz inputs), LineAxis0Multi (2 lines) and LineAxis0 (1 line: stage 1 only)."""
    out = {}
    aggs = {"min": ds.min("val"), "first": ds.first("val"), "last": ds.last("val"),
            "sum_nsi": ds.sum("val", self_intersect=False), "count_nsi": ds.count(self_intersect=False),
            "count_val_nsi": ds.count("val", self_intersect=False), "mean": ds.mean("val")}   # mean: single stage
    for tag, dtype in (("f32", np.float32), ("f64", np.float64)):
        xs, ys, val = line_frame(2024, 40, 24, dtype)       # identical to lines.npz in_{tag}_*
        nverts = xs.shape[1]
        d = {f"x{j}": xs[:, j] for j in range(nverts)}
        d.update({f"y{j}": ys[:, j] for j in range(nverts)})
        d["val"] = val
        df = pd.DataFrame(d)
        xcols, ycols = [f"x{j}" for j in range(nverts)], [f"y{j}" for j in range(nverts)]
        cvs = ds.Canvas(plot_width=64, plot_height=48, x_range=(0, 1), y_range=(0, 1))
        for lw in ((1, 2.5) if tag == "f32" else (1,)):
            for aname, agg in aggs.items():
                out[f"aa2_{tag}_lw{lw}_{aname}"] = np.asarray(cvs.line(df, x=xcols, y=ycols, axis=1, agg=agg, line_width=lw).data)
    # antialiased by(cat, r) on the 40-line f32 frame (categories from lines_extra.npz: in_cat)
    xs, ys, val = line_frame(2024, 40, 24, np.float32)
    nverts = xs.shape[1]
    codes = np.load(os.path.join(HERE, "lines_extra.npz"))["in_cat"]
    d = {f"x{j}": xs[:, j] for j in range(nverts)}
    d.update({f"y{j}": ys[:, j] for j in range(nverts)})
    d["val"] = val
    d["cat"] = pd.Categorical.from_codes(codes, categories=["a", "b", "c", "d"])
    df = pd.DataFrame(d)
    xcols, ycols = [f"x{j}" for j in range(nverts)], [f"y{j}" for j in range(nverts)]
    cvs = ds.Canvas(plot_width=64, plot_height=48, x_range=(0, 1), y_range=(0, 1))
    for aname, inner in {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val"), "mean": ds.mean("val")}.items():
        out[f"aaby_lw2_{aname}"] = np.asarray(cvs.line(df, x=xcols, y=ycols, axis=1, agg=ds.by("cat", inner), line_width=2).data)
    # axis=0 layouts: the line_layouts.npz inputs (ax0_x, ax0_y, ax0_x2, ax0_y2, ax0_val)
    g = np.load(os.path.join(HERE, "line_layouts.npz"))
    df0 = pd.DataFrame({k: g[f"ax0_{k}"] for k in ("x", "y", "x2", "y2", "val")})
    cvs = ds.Canvas(plot_width=50, plot_height=40, x_range=(0, 1), y_range=(0, 1))
    for aname, agg in aggs.items():
        out[f"aa2_ax0_lw2_{aname}"] = np.asarray(cvs.line(df0, "x", "y", agg=agg, line_width=2).data)
        out[f"aa2_ax0multi_lw2_{aname}"] = np.asarray(
            cvs.line(df0, x=["x", "x2"], y=["y", "y2"], axis=0, agg=agg, line_width=2).data)
    return out


def lines_aa3_cases():
    """Antialiased lines, composite aggregations (compiler.py:539-554: one 2-stage reduction forces count / sum of the
    whole summary to self_intersect=False; reductions.py:782-787: by() over a 2-stage reduction; :1906-1914: where())
    on the 40-line f32 frame of lines.npz with the categories of lines_extra.npz."""
    out = {}
    xs, ys, val = line_frame(2024, 40, 24, np.float32)
    nverts = xs.shape[1]
    codes = np.load(os.path.join(HERE, "lines_extra.npz"))["in_cat"]
    other = np.random.default_rng(77).random(len(val)).astype(np.float32) * 50 - 10
    other[7] = np.nan
    out["in_other"] = other
    d = {f"x{j}": xs[:, j] for j in range(nverts)}
    d.update({f"y{j}": ys[:, j] for j in range(nverts)})
    d["val"], d["other"] = val, other
    d["cat"] = pd.Categorical.from_codes(codes, categories=["a", "b", "c", "d"])
    df = pd.DataFrame(d)
    xcols, ycols = [f"x{j}" for j in range(nverts)], [f"y{j}" for j in range(nverts)]
    cvs = ds.Canvas(plot_width=64, plot_height=48, x_range=(0, 1), y_range=(0, 1))
    kw = dict(x=xcols, y=ycols, axis=1, line_width=2)
    summaries = {
        "s1": ds.summary(count=ds.count("val"), min=ds.min("val")),                       # min forces count to self_intersect=False
        "s2": ds.summary(count=ds.count("val", self_intersect=True), sum=ds.sum("val", self_intersect=True)),
        "s3": ds.summary(cnt=ds.count(), mx=ds.max("val"), first=ds.first("val"), anyv=ds.any()),
        "s4": ds.summary(count=ds.count(self_intersect=True), sum=ds.sum("val", self_intersect=False)),
        "s5": ds.summary(mean=ds.mean("val"), min=ds.min("val")),                         # mean next to a 2-stage member: its sum and
                                                                                          #   count are combined per line (SUM_2AGG)
        "s6": ds.summary(anyv=ds.any(), mn=ds.min("val"), mx=ds.max("val")),
        "s7": ds.summary(mx=ds.max("val"), last=ds.last("val"), sum=ds.sum("val"), mean=ds.mean("val")),
        "s8": ds.summary(mean=ds.mean("val"), count=ds.count(), sum=ds.sum("val")),       # nothing 2-stage: all single-stage
    }
    for sname, agg in summaries.items():
        res = cvs.line(df, agg=agg, **kw)
        for k in agg.keys:
            out[f"aa3_{sname}_{k}"] = np.asarray(res[k].data)
    for aname, inner in {"min": ds.min("val"), "first": ds.first("val"), "last": ds.last("val"),
                         "sum_nsi": ds.sum("val", self_intersect=False), "count_nsi": ds.count(self_intersect=False)}.items():
        out[f"aa3_by_{aname}"] = np.asarray(cvs.line(df, agg=ds.by("cat", inner), **kw).data)
    for aname, agg in {"by_where_first_row": ds.by("cat", ds.where(ds.first("val"))),
                       "by_where_last_other": ds.by("cat", ds.where(ds.last("val"), "other")),
                       "by_where_max_row": ds.by("cat", ds.where(ds.max("val"))),
                       "by_where_min_other": ds.by("cat", ds.where(ds.min("val"), "other"))}.items():
        out[f"aa3_{aname}"] = np.asarray(cvs.line(df, agg=agg, **kw).data)
    res = cvs.line(df, agg=ds.summary(m=ds.by("cat", ds.mean("val")), mn=ds.min("val")), **kw)
    out["aa3_s9_m"], out["aa3_s9_mn"] = np.asarray(res["m"].data), np.asarray(res["mn"].data)
    for aname, agg in {"where_first_row": ds.where(ds.first("val")), "where_first_other": ds.where(ds.first("val"), "other"),
                       "where_last_row": ds.where(ds.last("val")), "where_last_other": ds.where(ds.last("val"), "other"),
                       "where_max_row": ds.where(ds.max("val")), "where_max_other": ds.where(ds.max("val"), "other"),
                       "where_min_row": ds.where(ds.min("val")), "where_min_other": ds.where(ds.min("val"), "other")}.items():
        out[f"aa3_{aname}"] = np.asarray(cvs.line(df, agg=agg, **kw).data)
    return out


def tiles_cases():
    """Tile arithmetic of the reference's pyramid driver (tiles.py:99-117, 138-300): super tiles and 256-pixel tiles of a
    few extents and zoom levels, flattened to (tx, ty, level, xmin, ymin, xmax, ymax) rows."""
    from datashader.tiles import MercatorTileDefinition, gen_super_tiles
    out = {}
    extents = {"world": (-20037508.34, -20037508.34, 20037508.34, 20037508.34),
               "nyc": (-8242000.0, 4965000.0, -8210000.0, 4990000.0),
               "odd": (-1234567.8, 2345678.9, 3456789.1, 4567890.2)}
    for name, ext in extents.items():
        out[f"extent_{name}"] = np.array(ext)
        for level in (0, 1, 3, 5, 9):
            if name == "world" and level > 5:
                continue
            td = MercatorTileDefinition(x_range=(ext[0], ext[2]), y_range=(ext[1], ext[3]), tile_size=256)
            tiles = td.get_tiles_by_extent(ext, level)
            out[f"tiles_{name}_{level}"] = np.array([[t[0], t[1], t[2], *t[3]] for t in tiles], dtype=np.float64).reshape(-1, 7)
            sup = list(gen_super_tiles(ext, level))
            out[f"super_{name}_{level}"] = np.array([[s["level"], s["tile_size"], *s["x_range"], *s["y_range"]] for s in sup],
                                                    dtype=np.float64).reshape(-1, 6)
    return out


def spread_cases():
    """Post-shade image ops straight from the reference's kernels (composite.py; transfer_functions/__init__.py:748-1051)."""
    from datashader import composite as comp
    from datashader import transfer_functions as tf
    out = {}
    rng = np.random.default_rng(31)
    H, W = 23, 31
    img = rng.integers(0, 2 ** 32, (H, W), dtype=np.uint64).astype(np.uint32)
    img[rng.random((H, W)) < 0.6] = 0
    img[2, 3] = 0x00ffffff                       # colour but alpha 0
    img2 = rng.integers(0, 2 ** 32, (H, W), dtype=np.uint64).astype(np.uint32)
    img2[rng.random((H, W)) < 0.4] = 0
    out["img"], out["img2"] = img, img2
    for how in ("over", "add", "saturate", "source"):
        with np.errstate(divide="ignore", invalid="ignore"):
            out[f"comp_{how}"] = comp.composite_op_lookup[how](img, img2)
            out[f"comp_bg_{how}"] = comp.composite_op_lookup[how](img, np.uint32(0xff204060))
    f64 = rng.normal(size=(H, W))
    f64[rng.random((H, W)) < 0.7] = np.nan
    f32 = f64.astype(np.float32)
    i32 = rng.integers(-5, 9, (H, W)).astype(np.int32)
    i32[rng.random((H, W)) < 0.7] = 0
    u32 = rng.integers(0, 9, (H, W)).astype(np.uint32)
    u32[rng.random((H, W)) < 0.7] = 0
    cat = rng.integers(0, 5, (H, W, 3)).astype(np.uint32)
    cat[rng.random((H, W, 3)) < 0.8] = 0
    out.update(f64=f64, f32=f32, i32=i32, u32=u32, cat=cat)
    masks = {"c1": tf._circle_mask(1), "c2": tf._circle_mask(2), "c3": tf._circle_mask(3), "s1": tf._square_mask(1), "s2": tf._square_mask(2)}
    for mname, mask in masks.items():
        out[f"mask_{mname}"] = mask
        w = mask.shape[0]
        extra = w // 2

        def run(kernel, layer, fill):
            buf = np.full((layer.shape[0] + 2 * extra, layer.shape[1] + 2 * extra), fill, dtype=layer.dtype)
            kernel(layer, mask, buf)
            return buf[extra:-extra, extra:-extra].copy()

        for how in ("over", "add", "saturate", "source"):
            if mname in ("c1", "c2", "s1"):
                out[f"spread_img_{mname}_{how}"] = run(tf._build_spread_kernel(how, True), img, 0)
        for how in ("add", "max", "min", "source"):
            if mname in ("c1", "c3", "s2"):
                out[f"spread_f64_{mname}_{how}"] = run(tf._build_float_kernel(how, w), f64, np.nan)
                out[f"spread_f32_{mname}_{how}"] = run(tf._build_float_kernel(how, w), f32, np.nan)
                out[f"spread_i32_{mname}_{how}"] = run(tf._build_int_kernel(how, w, False), i32, 0)
                out[f"spread_u32_{mname}_{how}"] = run(tf._build_int_kernel(how, w, True), u32, 0)
        if mname == "c2":
            out["spread_cat_c2_add"] = np.dstack([run(tf._build_int_kernel("add", w, True), np.ascontiguousarray(cat[:, :, c]), 0)
                                                  for c in range(3)])
    for px in (1, 2, 4, 6):
        out[f"density_img_{px}"] = np.float64(tf._rgb_density(img, px))
        out[f"density_f64_{px}"] = np.float64(tf._array_density(f64, True, px))
        out[f"density_u32_{px}"] = np.float64(tf._array_density(u32, False, px))
    sparse = np.zeros((40, 50), np.uint32)
    sparse[rng.integers(0, 40, 25), rng.integers(0, 50, 25)] = 0xff0000ff
    out["sparse"] = sparse
    for px in (2, 4, 6):
        out[f"density_sparse_{px}"] = np.float64(tf._rgb_density(sparse, px))
    return out


def area_cases():
    """Canvas.area for the ten non-ragged layouts (core.py:480-709, glyphs/area.py)."""
    out = {}
    rng = np.random.default_rng(808)
    n = 50
    x = np.linspace(-0.1, 1.1, n) + rng.normal(0, 0.01, n)
    y = np.cumsum(rng.normal(0, 0.12, n)) + 0.2
    ys = y - rng.random(n) * 0.4
    x2 = np.linspace(1.05, -0.05, n)
    y2 = np.cumsum(rng.normal(0, 0.1, n))
    y2s = y2 + rng.random(n) * 0.3
    y[11] = np.nan
    ys[30] = np.nan
    val = (rng.random(n) * 3).astype(np.float32)
    val[7] = np.nan
    f32 = np.float32
    df0 = pd.DataFrame({"x": x.astype(f32), "y": y.astype(f32), "ys": ys.astype(f32), "x2": x2.astype(f32),
                        "y2": y2.astype(f32), "y2s": y2s.astype(f32), "val": val})
    for k in df0.columns:
        out[f"a0_{k}"] = df0[k].to_numpy()
    canvases = {"fixed": ds.Canvas(plot_width=45, plot_height=35, x_range=(0, 1), y_range=(-0.5, 1.0)),
                "auto": ds.Canvas(plot_width=33, plot_height=27)}
    aggs = {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val"), "first": ds.first("val")}
    for cn, cvs in canvases.items():
        for an, agg in aggs.items():
            out[f"a0_zero_{cn}_{an}"] = np.asarray(cvs.area(df0, "x", "y", agg=agg).data)
            out[f"a0_line_{cn}_{an}"] = np.asarray(cvs.area(df0, "x", "y", agg=agg, y_stack="ys").data)
            out[f"a0m_zero_{cn}_{an}"] = np.asarray(cvs.area(df0, x=["x", "x2"], y=["y", "y2"], agg=agg, axis=0).data)
            out[f"a0m_line_{cn}_{an}"] = np.asarray(
                cvs.area(df0, x=["x", "x2"], y=["y", "y2"], y_stack=["ys", "y2s"], agg=agg, axis=0).data)
        r = cvs.area(df0, "x", "y")
        out[f"a0_zero_{cn}_yrange"] = np.asarray(r.attrs["y_range"], dtype=np.float64)

    nl, nv = 10, 8
    xm = (np.tile(np.linspace(-0.1, 1.1, nv), (nl, 1)) + rng.normal(0, 0.02, (nl, nv))).astype(f32)
    ym = (rng.random((nl, nv)) * 1.4 - 0.3).astype(f32)
    ysm = (ym - rng.random((nl, nv)).astype(f32) * 0.5).astype(f32)
    ym[2, 3] = np.nan
    lval = (rng.random(nl) * 2).astype(f32)
    out["a1_x"], out["a1_y"], out["a1_ys"], out["a1_val"] = xm, ym, ysm, lval
    d = {f"x{j}": xm[:, j] for j in range(nv)}
    d.update({f"y{j}": ym[:, j] for j in range(nv)})
    d.update({f"s{j}": ysm[:, j] for j in range(nv)})
    d["val"] = lval
    df1 = pd.DataFrame(d)
    xc, yc, sc = [f"x{j}" for j in range(nv)], [f"y{j}" for j in range(nv)], [f"s{j}" for j in range(nv)]
    xconst = np.linspace(0.0, 1.0, nv)
    yconst = np.linspace(0.9, -0.2, nv)
    sconst = yconst - 0.25
    out["a1_xconst"], out["a1_yconst"], out["a1_sconst"] = xconst, yconst, sconst
    cvs = canvases["fixed"]
    aggs1 = {"any": ds.any(), "count": ds.count(), "max": ds.max("val"), "mean": ds.mean("val")}
    for an, agg in aggs1.items():
        out[f"a1_zero_{an}"] = np.asarray(cvs.area(df1, x=xc, y=yc, agg=agg, axis=1).data)
        out[f"a1_line_{an}"] = np.asarray(cvs.area(df1, x=xc, y=yc, y_stack=sc, agg=agg, axis=1).data)
        out[f"a1xc_zero_{an}"] = np.asarray(cvs.area(df1, x=xconst, y=yc, agg=agg, axis=1).data)
        out[f"a1xc_line_{an}"] = np.asarray(cvs.area(df1, x=xconst, y=yc, y_stack=sc, agg=agg, axis=1).data)
        out[f"a1yc_zero_{an}"] = np.asarray(cvs.area(df1, x=xc, y=yconst, agg=agg, axis=1).data)
        out[f"a1yc_line_{an}"] = np.asarray(cvs.area(df1, x=xc, y=yconst, y_stack=sconst, agg=agg, axis=1).data)
    return out


def ragged_cases():
    """The ragged layouts: LinesAxis1Ragged (core.py:443-444, line.py:457-523 + 1538-1600) Bresenham, antialiased single- and
    2-stage; AreaToZeroAxis1Ragged / AreaToLineAxis1Ragged (core.py:676-677, 698-699; area.py:916-1073, 1939-2083).  Rows of
    different lengths, an empty row, a one-vertex row, NaN vertices, x / y rows of different lengths (the shorter one counts),
    float32 and float64 flat arrays."""
    from datashader.datatypes import RaggedArray
    out = {}
    rng = np.random.default_rng(1212)
    lens_x = [7, 1, 0, 12, 5, 9, 2, 30, 3, 16]
    lens_y = [7, 1, 0, 12, 8, 6, 2, 30, 3, 16]          # rows 4 and 5: x and y lengths differ
    lens_s = [7, 1, 0, 12, 8, 6, 2, 25, 3, 16]          # row 7: the stack curve is shorter still
    nl = len(lens_x)

    def walk(n, lo=-0.1, hi=1.1):
        return np.sort(rng.random(n) * (hi - lo) + lo) if n else np.empty(0)

    xr = [walk(n) for n in lens_x]
    yr = [np.clip(np.cumsum(rng.normal(0, 0.15, n)) + rng.random(), -0.4, 1.3) for n in lens_y]
    sr = [yr[i][:n] - rng.random(n) * 0.4 if n <= len(yr[i]) else np.r_[yr[i], yr[i][-1:].repeat(n - len(yr[i]))] - 0.2
          for i, n in enumerate(lens_s)]
    xr[3][2] = np.nan
    yr[7][20] = np.nan
    sr[9][5] = np.nan
    yr[3][:] = yr[3][::-1]
    xr[8] = xr[8][::-1].copy()                           # a right-to-left row
    val = (rng.random(nl) * 5 - 1).astype(np.float32)
    val[6] = np.nan
    cat = rng.integers(0, 3, nl).astype(np.int8)
    for dt in ("float32", "float64"):
        d = {"x": RaggedArray(xr, dtype=dt), "y": RaggedArray(yr, dtype=dt), "s": RaggedArray(sr, dtype=dt), "val": val,
             "cat": pd.Categorical.from_codes(cat, categories=["a", "b", "c"])}
        df = pd.DataFrame(d)
        t = "f32" if dt == "float32" else "f64"
        for k in ("x", "y", "s"):
            out[f"{t}_{k}_flat"] = np.asarray(df[k].array.flat_array)
            out[f"{t}_{k}_starts"] = np.asarray(df[k].array.start_indices).astype(np.int64)
        out["val"], out["cat"] = val, cat
        cvs = ds.Canvas(plot_width=48, plot_height=36, x_range=(0, 1), y_range=(-0.2, 1.1))
        auto = ds.Canvas(plot_width=31, plot_height=23)
        aggs = {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val"), "min": ds.min("val"),
                "mean": ds.mean("val"), "first": ds.first("val"), "last": ds.last("val"),
                "where_max_row": ds.where(ds.max("val")), "by_count": ds.by("cat", ds.count())}
        for name, agg in aggs.items():
            out[f"{t}_line_lw0_{name}"] = np.asarray(cvs.line(df, "x", "y", agg=agg, axis=1).data)
        r = auto.line(df, "x", "y", agg=ds.count(), axis=1)
        out[f"{t}_line_auto_count"] = np.asarray(r.data)
        out[f"{t}_line_auto_ranges"] = np.asarray(list(r.attrs["x_range"]) + list(r.attrs["y_range"]), dtype=np.float64)
        aa = {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val"), "mean": ds.mean("val"),
              "min": ds.min("val"), "first": ds.first("val"), "last": ds.last("val"),
              "count_nsi": ds.count(self_intersect=False), "sum_nsi": ds.sum("val", self_intersect=False),
              "where_max_row": ds.where(ds.max("val")), "by_max": ds.by("cat", ds.max("val"))}
        for name, agg in aa.items():
            for lw in (1, 2.5):
                out[f"{t}_line_lw{lw}_{name}"] = np.asarray(cvs.line(df, "x", "y", agg=agg, axis=1, line_width=lw).data)
        area_aggs = {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val"),
                     "first": ds.first("val"), "by_count": ds.by("cat", ds.count())}
        for name, agg in area_aggs.items():
            out[f"{t}_area_zero_{name}"] = np.asarray(cvs.area(df, "x", "y", agg=agg, axis=1).data)
            out[f"{t}_area_line_{name}"] = np.asarray(cvs.area(df, "x", "y", agg=agg, axis=1, y_stack="s").data)
        r = auto.area(df, "x", "y", agg=ds.count(), axis=1)
        out[f"{t}_area_zero_auto_count"] = np.asarray(r.data)
        out[f"{t}_area_zero_auto_ranges"] = np.asarray(list(r.attrs["x_range"]) + list(r.attrs["y_range"]), dtype=np.float64)
        r = auto.area(df, "x", "y", agg=ds.count(), axis=1, y_stack="s")
        out[f"{t}_area_line_auto_count"] = np.asarray(r.data)
        out[f"{t}_area_line_auto_ranges"] = np.asarray(list(r.attrs["x_range"]) + list(r.attrs["y_range"]), dtype=np.float64)
    return out


def negzero_cases():
    """max / min keep whichever zero ARRIVED FIRST when the extreme of a pixel is a zero (strict compare,
    reductions.py:1178-1183, 1222-1227): columns mixing -0.0 and +0.0 with values on one side of zero only."""
    out = {}
    rng = np.random.default_rng(606)
    n = 4000
    x = rng.random(n).astype(np.float32)
    y = rng.random(n).astype(np.float32)
    zeros = np.where(rng.random(n) < 0.5, np.float32(-0.0), np.float32(0.0))
    neg = -rng.random(n).astype(np.float32) - 0.1
    vmax32 = np.where(rng.random(n) < 0.4, zeros, neg).astype(np.float32)          # the max of most pixels is a zero
    vmin32 = np.where(rng.random(n) < 0.4, zeros, -neg).astype(np.float32)         # the min of most pixels is a zero
    vmax32[rng.integers(0, n, n // 30)] = np.nan
    cols = dict(x=x, y=y, vmax32=vmax32, vmin32=vmin32, vmax64=vmax32.astype(np.float64), vmin64=vmin32.astype(np.float64),
                cat=rng.integers(0, NCAT, n).astype(np.int8))
    for k, v in cols.items():
        out[f"in_{k}"] = v
    df = to_df(cols)
    cvs = ds.Canvas(plot_width=9, plot_height=7, x_range=(0, 1), y_range=(0, 1))
    for name, red in (("max_vmax32", ds.max("vmax32")), ("min_vmin32", ds.min("vmin32")), ("max_vmax64", ds.max("vmax64")),
                      ("min_vmin64", ds.min("vmin64")), ("by_max_vmax32", ds.by("cat", ds.max("vmax32"))),
                      ("max_vmin32", ds.max("vmin32")), ("min_vmax32", ds.min("vmax32"))):
        out[f"nz_{name}"] = np.asarray(cvs.points(df, "x", "y", red).data)
    return out


def main():
    if "--ragged-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "ragged.npz"), **ragged_cases())
        print("ragged.npz", os.path.getsize(os.path.join(HERE, "ragged.npz")) // 1024, "KiB")
        return
    if "--lines-aa3-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "lines_aa3.npz"), **lines_aa3_cases())
        print("lines_aa3.npz", os.path.getsize(os.path.join(HERE, "lines_aa3.npz")) // 1024, "KiB")
        return
    if "--shade-catfloat-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "shade_catfloat.npz"), **shade_catfloat_cases())
        print("shade_catfloat.npz", os.path.getsize(os.path.join(HERE, "shade_catfloat.npz")) // 1024, "KiB")
        return
    if "--shade-extra-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "shade_extra.npz"), **shade_extra_cases())
        print("shade_extra.npz", os.path.getsize(os.path.join(HERE, "shade_extra.npz")) // 1024, "KiB")
        return
    if "--tiles-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "tiles.npz"), **tiles_cases())
        print("tiles.npz", os.path.getsize(os.path.join(HERE, "tiles.npz")) // 1024, "KiB")
        return
    if "--negzero-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "points_negzero.npz"), **negzero_cases())
        print("points_negzero.npz", os.path.getsize(os.path.join(HERE, "points_negzero.npz")) // 1024, "KiB")
        return
    if "--lines-aa2-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "lines_aa2.npz"), **lines_aa2_cases())
        print("lines_aa2.npz", os.path.getsize(os.path.join(HERE, "lines_aa2.npz")) // 1024, "KiB")
        return
    if "--spread-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "spread.npz"), **spread_cases())
        print("spread.npz", os.path.getsize(os.path.join(HERE, "spread.npz")) // 1024, "KiB")
        return
    if "--shade-span-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "shade_span.npz"), **shade_span_cases())
        print("shade_span.npz", os.path.getsize(os.path.join(HERE, "shade_span.npz")) // 1024, "KiB")
        return
    if "--areas-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "areas.npz"), **area_cases())
        print("areas.npz", os.path.getsize(os.path.join(HERE, "areas.npz")) // 1024, "KiB")
        return
    if "--line-layouts-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "line_layouts.npz"), **line_layout_cases())
        print("line_layouts.npz", os.path.getsize(os.path.join(HERE, "line_layouts.npz")) // 1024, "KiB")
        return
    if "--lines-extra-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "lines_extra.npz"), **lines_extra_cases())
        print("lines_extra.npz", os.path.getsize(os.path.join(HERE, "lines_extra.npz")) // 1024, "KiB")
        return
    if "--shade-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "shade.npz"), **shade_cases())
        print("shade.npz", os.path.getsize(os.path.join(HERE, "shade.npz")) // 1024, "KiB")
        return
    np.savez_compressed(os.path.join(HERE, "shade.npz"), **shade_cases())
    np.savez_compressed(os.path.join(HERE, "lines_extra.npz"), **lines_extra_cases())
    np.savez_compressed(os.path.join(HERE, "line_layouts.npz"), **line_layout_cases())
    np.savez_compressed(os.path.join(HERE, "areas.npz"), **area_cases())
    np.savez_compressed(os.path.join(HERE, "ragged.npz"), **ragged_cases())
    np.savez_compressed(os.path.join(HERE, "lines_aa2.npz"), **lines_aa2_cases())
    np.savez_compressed(os.path.join(HERE, "lines_aa3.npz"), **lines_aa3_cases())
    np.savez_compressed(os.path.join(HERE, "tiles.npz"), **tiles_cases())
    np.savez_compressed(os.path.join(HERE, "spread.npz"), **spread_cases())
    np.savez_compressed(os.path.join(HERE, "shade_span.npz"), **shade_span_cases())
    np.savez_compressed(os.path.join(HERE, "shade_catfloat.npz"), **shade_catfloat_cases())
    np.savez_compressed(os.path.join(HERE, "shade_extra.npz"), **shade_extra_cases())
    np.savez_compressed(os.path.join(HERE, "points.npz"), **points_cases())
    np.savez_compressed(os.path.join(HERE, "points_negzero.npz"), **negzero_cases())
    np.savez_compressed(os.path.join(HERE, "partitioned.npz"), **partitioned_cases())
    np.savez_compressed(os.path.join(HERE, "lines.npz"), **lines_cases())
    for f in ("points.npz", "partitioned.npz", "lines.npz", "shade.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
