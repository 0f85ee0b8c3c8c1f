"""LinesAxis1 on the GPU against the reference's golden vectors: Bresenham bit-exact; antialiased
any/max bit-exact, antialiased count/sum within 1e-6 (atomic float adds change the summation order)."""
import numpy as np
import pytest

from helpers import LINE_CANVASES, load

pytestmark = pytest.mark.gpu


def _frame(xs, ys, val):
    import pandas as pd
    nverts = xs.shape[1]
    d = {f"x{j}": xs[:, j] for j in range(nverts)}
    d.update({f"y{j}": ys[:, j] for j in range(nverts)})
    d["val"] = val
    return pd.DataFrame(d), [f"x{j}" for j in range(nverts)], [f"y{j}" for j in range(nverts)]


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("cname", list(LINE_CANVASES))
def test_lines_golden(tag, cname):
    import datashader_b200 as ds
    g = load("lines.npz")
    df, xc, yc = _frame(g[f"in_{tag}_xs"], g[f"in_{tag}_ys"], g[f"in_{tag}_val"])
    cvs = ds.Canvas(**LINE_CANVASES[cname])
    n = 0
    for key in g.files:
        pre = f"ln_{tag}_{cname}_lw"
        if not key.startswith(pre):
            continue
        lw, aname = key[len(pre):].split("_")
        lw = float(lw)
        agg = {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val"), "min": ds.min("val")}[aname]
        got = cvs.line(df, x=xc, y=yc, axis=1, agg=agg, line_width=lw).data
        want = g[key]
        assert got.dtype == want.dtype and got.shape == want.shape, key
        if aname in ("count", "sum") and (lw > 0 or aname == "sum"):
            assert np.array_equal(np.isnan(got), np.isnan(want)), key
            np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6, equal_nan=True, err_msg=key)
        else:
            assert np.array_equal(got, want, equal_nan=(got.dtype.kind == "f")), key
        n += 1
    assert n >= 8


def test_lines_vs_oracle_larger():
    import datashader_b200 as ds
    from oracle import oracle as ora
    rng = np.random.default_rng(8)
    nl, nv = 300, 64
    xs = np.tile(np.arange(nv, dtype=np.float32), (nl, 1))
    ys = np.cumsum(rng.standard_normal((nl, nv)), axis=1).astype(np.float32)
    val = rng.random(nl).astype(np.float32)
    df, xc, yc = _frame(xs, ys, val)
    xr, yr = (0.0, float(nv - 1)), (float(ys.min()), float(ys.max()))
    view = ora.make_view(384, 216, xr, yr)
    cvs = ds.Canvas(384, 216, x_range=xr, y_range=yr)
    for lw in (0, 1):
        got = cvs.line(df, x=xc, y=yc, axis=1, agg=ds.max("val"), line_width=lw).data
        want = ora.lines_axis1(xs, ys, view, agg="max", values=val, line_width=lw)
        assert np.array_equal(got, want, equal_nan=True), lw
    got = cvs.line(df, x=xc, y=yc, axis=1, agg=ds.count()).data
    assert np.array_equal(got, ora.lines_axis1(xs, ys, view, agg="count"))


def test_lines_all_reductions_golden():
    """line_width=0 runs the same accumulator plans as points: mean / first / last / where / by on lines."""
    import pandas as pd
    import datashader_b200 as ds
    g, gx = load("lines.npz"), load("lines_extra.npz")
    df, xc, yc = _frame(g["in_f32_xs"], g["in_f32_ys"], g["in_f32_val"])
    df["other"] = gx["in_other"]
    df["cat"] = pd.Categorical.from_codes(gx["in_cat"], categories=["a", "b", "c", "d"])
    cvs = ds.Canvas(plot_width=64, plot_height=48, x_range=(0, 1), y_range=(0, 1))
    aggs = {
        "count_val": ds.count("val"), "mean_val": ds.mean("val"), "first_val": ds.first("val"), "last_val": ds.last("val"),
        "where_max_val_row": ds.where(ds.max("val")), "where_min_val_other": ds.where(ds.min("val"), "other"),
        "where_first_val_other": ds.where(ds.first("val"), "other"), "by_count": ds.by("cat", ds.count()),
        "by_max_val": ds.by("cat", ds.max("val")), "by_any": ds.by("cat", ds.any()),
    }
    for name, agg in aggs.items():
        got = cvs.line(df, x=xc, y=yc, axis=1, agg=agg).data
        want = gx[f"lnx_{name}"]
        assert got.dtype == want.dtype and got.shape == want.shape, name
        if name == "mean_val":
            np.testing.assert_allclose(got, want, rtol=1e-12, equal_nan=True, err_msg=name)
        else:
            assert np.array_equal(got, want, equal_nan=(got.dtype.kind == "f")), name


def _cmp(got, want, key, float_sum=False):
    assert got.dtype == want.dtype and got.shape == want.shape, key
    if float_sum:
        assert np.array_equal(np.isnan(got), np.isnan(want)), key
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-9, equal_nan=True, err_msg=key)
    else:
        assert np.array_equal(got, want, equal_nan=(got.dtype.kind == "f")), key


def test_line_layouts_golden():
    """LineAxis0, LineAxis0Multi, LinesAxis1XConstant, LinesAxis1YConstant against the real reference."""
    import pandas as pd
    import datashader_b200 as ds
    g = load("line_layouts.npz")
    df0 = pd.DataFrame({k: g[f"ax0_{k}"] for k in ("x", "y", "x2", "y2", "val")})
    cvs = ds.Canvas(plot_width=50, plot_height=40, x_range=(0, 1), y_range=(0, 1))
    aggs0 = {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val"), "mean": ds.mean("val"),
             "first": ds.first("val"), "where_max_row": ds.where(ds.max("val"))}
    for name, agg in aggs0.items():
        r = cvs.line(df0, "x", "y", agg=agg)
        assert tuple(r.dims) == ("y", "x")
        _cmp(r.data, g[f"ax0_lw0_{name}"], f"ax0 {name}", name in ("sum", "mean"))
        _cmp(cvs.line(df0, x=["x", "x2"], y=["y", "y2"], axis=0, agg=agg).data, g[f"ax0multi_lw0_{name}"],
             f"ax0multi {name}", name in ("sum", "mean"))
    for name in ("any", "max"):
        _cmp(cvs.line(df0, "x", "y", agg=aggs0[name], line_width=1).data, g[f"ax0_lw1_{name}"], f"ax0 aa {name}")
        _cmp(cvs.line(df0, x=["x", "x2"], y=["y", "y2"], axis=0, agg=aggs0[name], line_width=1).data,
             g[f"ax0multi_lw1_{name}"], f"ax0multi aa {name}")
    xc, ys, lval = g["xc_x"], g["xc_ys"], g["xc_val"]
    nv = ys.shape[1]
    df1 = pd.DataFrame({**{f"y{j}": ys[:, j] for j in range(nv)}, "val": lval})
    df2 = pd.DataFrame({**{f"x{j}": ys[:, j] for j in range(nv)}, "val": lval})
    ycols, xcols = [f"y{j}" for j in range(nv)], [f"x{j}" for j in range(nv)]
    aggs1 = {"any": ds.any(), "count": ds.count(), "max": ds.max("val"), "mean": ds.mean("val")}
    for name, agg in aggs1.items():
        _cmp(cvs.line(df1, x=xc, y=ycols, axis=1, agg=agg).data, g[f"xconst_lw0_{name}"], f"xconst {name}", name == "mean")
        _cmp(cvs.line(df2, x=xcols, y=xc, axis=1, agg=agg).data, g[f"yconst_lw0_{name}"], f"yconst {name}", name == "mean")
    _cmp(cvs.line(df1, x=xc, y=ycols, axis=1, agg=ds.max("val"), line_width=1).data, g["xconst_lw1_max"], "xconst aa")
    _cmp(cvs.line(df2, x=xcols, y=xc, axis=1, agg=ds.max("val"), line_width=1).data, g["yconst_lw1_max"], "yconst aa")


@pytest.mark.parametrize("dtype", ["f4", "f8"])
def test_line_axis0_vs_oracle_random(dtype):
    """LineAxis0 / LineAxis0Multi at a few thousand vertices with NaN breaks on a clipping canvas vs the C oracle
    (ora_lines with value_per_vertex): Bresenham bit-exact, antialiased to 1e-6 relative (fp contraction only)."""
    import pandas as pd
    import datashader_b200 as ds
    from oracle import oracle as ora
    rng = np.random.default_rng(5)
    n = 5000
    x = np.cumsum(rng.normal(0, 0.01, n)).astype(dtype) + np.asarray(0.5, dtype)
    y = np.cumsum(rng.normal(0, 0.01, n)).astype(dtype) + np.asarray(0.5, dtype)
    x2 = rng.uniform(-0.2, 1.2, n).astype(dtype)
    y2 = rng.uniform(-0.2, 1.2, n).astype(dtype)
    for a in (x, y2):
        a[rng.integers(0, n, 20)] = np.nan
    val = rng.normal(size=n)
    df = pd.DataFrame({"x": x, "y": y, "x2": x2, "y2": y2, "val": val})
    W, H = 240, 180
    cvs = ds.Canvas(plot_width=W, plot_height=H, x_range=(0.2, 0.8), y_range=(0.3, 0.9))
    view = ora.make_view(W, H, (0.2, 0.8), (0.3, 0.9))
    for name, agg in {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val"), "min": ds.min("val")}.items():
        vals = None if name in ("any", "count") else val
        _cmp(cvs.line(df, "x", "y", agg=agg).data, ora.lines(x[None], y[None], view, name, vals, 0, per_vertex=True), f"ax0 {name}", name == "sum")
        _cmp(cvs.line(df, x=["x", "x2"], y=["y", "y2"], agg=agg, axis=0).data,
             ora.lines(np.stack([x, x2]), np.stack([y, y2]), view, name, vals, 0, per_vertex=True), f"ax0multi {name}", name == "sum")
    for name, agg in {"any": ds.any(), "max": ds.max("val"), "count": ds.count(), "sum": ds.sum("val")}.items():
        vals = None if name in ("any", "count") else val
        got = cvs.line(df, "x", "y", agg=agg, line_width=2.5).data
        want = ora.lines(x[None], y[None], view, name, vals, 2.5, per_vertex=True)
        assert got.dtype == want.dtype
        assert np.array_equal(np.isnan(got), np.isnan(want)), name
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6, equal_nan=True, err_msg=name)


AA2 = {"min": ("min", lambda ds: ds.min("val")), "first": ("first", lambda ds: ds.first("val")),
       "last": ("last", lambda ds: ds.last("val")), "sum_nsi": ("sum", lambda ds: ds.sum("val", self_intersect=False)),
       "count_nsi": ("count", lambda ds: ds.count(self_intersect=False)),
       "count_val_nsi": ("count", lambda ds: ds.count("val", self_intersect=False)),
       "mean": ("mean", lambda ds: ds.mean("val"))}      # single-stage, rides along with the same goldens


def _cmp_aa(got, want, key, rtol=1e-6):
    assert got.dtype == want.dtype and got.shape == want.shape, key
    assert np.array_equal(np.isnan(got), np.isnan(want)), key
    np.testing.assert_allclose(got, want, rtol=rtol, atol=1e-7, equal_nan=True, err_msg=key)


def test_lines_aa2_golden():
    """2-stage antialiased reductions (L7: compiler.py:198-268) vs the real reference: LinesAxis1 (40 lines), LineAxis0Multi
    and LineAxis0.  Tolerance 1e-6: the GPU contracts a few f64 expressions of the coverage computation differently."""
    import pandas as pd
    import datashader_b200 as ds
    g, gl, gx = load("lines_aa2.npz"), load("lines.npz"), load("line_layouts.npz")
    for tag in ("f32", "f64"):
        frame, xcols, ycols = _frame(gl[f"in_{tag}_xs"], gl[f"in_{tag}_ys"], gl[f"in_{tag}_val"])
        cvs = ds.Canvas(plot_width=64, plot_height=48, x_range=(0, 1), y_range=(0, 1))
        for lw in ((1, 2.5) if tag == "f32" else (1,)):
            for gname, (_o, mk) in AA2.items():
                got = cvs.line(frame, x=xcols, y=ycols, axis=1, agg=mk(ds), line_width=lw).data
                _cmp_aa(got, g[f"aa2_{tag}_lw{lw}_{gname}"], f"{tag} lw{lw} {gname}")
    df0 = pd.DataFrame({k: gx[f"ax0_{k}"] for k in ("x", "y", "x2", "y2", "val")})
    cvs = ds.Canvas(plot_width=50, plot_height=40, x_range=(0, 1), y_range=(0, 1))
    for gname, (_o, mk) in AA2.items():
        _cmp_aa(cvs.line(df0, "x", "y", agg=mk(ds), line_width=2).data, g[f"aa2_ax0_lw2_{gname}"], f"ax0 {gname}")
        _cmp_aa(cvs.line(df0, x=["x", "x2"], y=["y", "y2"], axis=0, agg=mk(ds), line_width=2).data, g[f"aa2_ax0multi_lw2_{gname}"],
                f"ax0multi {gname}")


def test_lines_aa2_vs_oracle_larger():
    """3000 lines x 16 vertices on a clipping 300x200 canvas: many lines per CTA, heavy overlap."""
    import datashader_b200 as ds
    from oracle import oracle as ora
    rng = np.random.default_rng(99)
    nl, nv = 3000, 16
    xs = (np.cumsum(rng.normal(0, 0.04, (nl, nv)), axis=1) + rng.random((nl, 1))).astype(np.float32)
    ys = (np.cumsum(rng.normal(0, 0.04, (nl, nv)), axis=1) + rng.random((nl, 1))).astype(np.float32)
    xs[rng.random((nl, nv)) < 0.02] = np.nan
    val = (rng.random(nl) * 6 - 2).astype(np.float32)
    val[rng.integers(0, nl, 30)] = np.nan
    frame, xcols, ycols = _frame(xs, ys, val)
    W, H = 300, 200
    cvs = ds.Canvas(plot_width=W, plot_height=H, x_range=(0.1, 0.9), y_range=(0.2, 0.8))
    view = ora.make_view(W, H, (0.1, 0.9), (0.2, 0.8))
    for gname, (oname, mk) in AA2.items():
        got = cvs.line(frame, x=xcols, y=ycols, axis=1, agg=mk(ds), line_width=1.5).data
        if oname == "mean":
            want = ora.lines(xs, ys, view, "mean", val, 1.5)
        else:
            want = ora.lines_aa2(xs, ys, view, oname, None if gname == "count_nsi" else val, 1.5)
        _cmp_aa(got, want, gname, rtol=2e-6 if oname in ("sum", "count") else 1e-6)


def test_lines_aa2_long_lines_overflow_the_stage1_table():
    """Lines that touch far more than 6144 pixels (wide line_width on a 1200x900 canvas) take the global stage-1 canvas;
    the result must match the oracle either way."""
    import datashader_b200 as ds
    from oracle import oracle as ora
    rng = np.random.default_rng(3)
    nl, nv = 24, 40
    xs = np.sort(rng.random((nl, nv)), axis=1).astype(np.float32)
    ys = rng.random((nl, nv)).astype(np.float32)
    xs[:4] = rng.random((4, nv)).astype(np.float32) * 0.02 + 0.5      # 4 short lines stay in the table
    ys[:4] = rng.random((4, nv)).astype(np.float32) * 0.02 + 0.5
    val = (rng.random(nl) * 5 + 1).astype(np.float32)
    frame, xcols, ycols = _frame(xs, ys, val)
    W, H = 1200, 900
    cvs = ds.Canvas(plot_width=W, plot_height=H, x_range=(0, 1), y_range=(0, 1))
    view = ora.make_view(W, H, (0, 1), (0, 1))
    for gname in ("min", "first", "last", "sum_nsi"):
        oname, mk = AA2[gname]
        got = cvs.line(frame, x=xcols, y=ycols, axis=1, agg=mk(ds), line_width=4).data
        _cmp_aa(got, ora.lines_aa2(xs, ys, view, oname, val, 4), gname, rtol=2e-6)


def test_lines_aa_by_category_golden():
    """Antialiased by('cat', any | count | sum | max | mean): every line updates the [H, W] plane of its category."""
    import pandas as pd
    import datashader_b200 as ds
    g, gl, ge = load("lines_aa2.npz"), load("lines.npz"), load("lines_extra.npz")
    frame, xcols, ycols = _frame(gl["in_f32_xs"], gl["in_f32_ys"], gl["in_f32_val"])
    frame["cat"] = pd.Categorical.from_codes(ge["in_cat"], categories=["a", "b", "c", "d"])
    cvs = ds.Canvas(plot_width=64, plot_height=48, x_range=(0, 1), y_range=(0, 1))
    inner = {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val"), "mean": ds.mean("val")}
    for aname, red in inner.items():
        r = cvs.line(frame, x=xcols, y=ycols, axis=1, agg=ds.by("cat", red), line_width=2)
        want = g[f"aaby_lw2_{aname}"]
        assert tuple(r.dims) == ("y", "x", "cat") and list(r.coords["cat"]) == ["a", "b", "c", "d"], aname
        _cmp_aa(r.data, want, f"aa by {aname}")


def test_lines_antialiased_summary_by_where_golden():
    """Composite antialiased aggregations vs the real reference (tests/golden/lines_aa3.npz): summary() where one 2-stage
    reduction forces every count / sum to self_intersect=False (compiler.py:539-554, test_pandas.py:2803
    TestLineAntialiasSummary), by() over 2-stage reductions (reductions.py:782-787) and where(first | last)
    (test_pandas.py:3079 TestLineAntialiasWhere)."""
    import pandas as pd
    import datashader_b200 as ds
    g, gl, ge = load("lines_aa3.npz"), load("lines.npz"), load("lines_extra.npz")
    frame, xcols, ycols = _frame(gl["in_f32_xs"], gl["in_f32_ys"], gl["in_f32_val"])
    frame["other"] = g["in_other"]
    frame["cat"] = pd.Categorical.from_codes(ge["in_cat"], categories=["a", "b", "c", "d"])
    cvs = ds.Canvas(plot_width=64, plot_height=48, x_range=(0, 1), y_range=(0, 1))
    kw = dict(x=xcols, y=ycols, axis=1, line_width=2)
    summaries = {
        "s1": ds.summary(count=ds.count("val"), min=ds.min("val")),
        "s2": ds.summary(count=ds.count("val", self_intersect=True), sum=ds.sum("val", self_intersect=True)),
        "s3": ds.summary(cnt=ds.count(), mx=ds.max("val"), first=ds.first("val"), anyv=ds.any()),
        "s4": ds.summary(count=ds.count(self_intersect=True), sum=ds.sum("val", self_intersect=False)),
        "s5": ds.summary(mean=ds.mean("val"), min=ds.min("val")),      # mean's bases drawn in overwrite mode, combined per line
        "s6": ds.summary(anyv=ds.any(), mn=ds.min("val"), mx=ds.max("val")),
        "s7": ds.summary(mx=ds.max("val"), last=ds.last("val"), sum=ds.sum("val"), mean=ds.mean("val")),
        "s8": ds.summary(mean=ds.mean("val"), count=ds.count(), sum=ds.sum("val")),
    }
    for sname, agg in summaries.items():
        res = cvs.line(frame, agg=agg, **kw)
        for k in agg.keys:
            _cmp_aa(res[k].data, g[f"aa3_{sname}_{k}"], f"summary {sname}.{k}")
    for aname, inner in {"min": ds.min("val"), "first": ds.first("val"), "last": ds.last("val"),
                         "sum_nsi": ds.sum("val", self_intersect=False), "count_nsi": ds.count(self_intersect=False)}.items():
        r = cvs.line(frame, agg=ds.by("cat", inner), **kw)
        assert tuple(r.dims) == ("y", "x", "cat") and list(r.coords["cat"]) == ["a", "b", "c", "d"], aname
        _cmp_aa(r.data, g[f"aa3_by_{aname}"], f"aa by {aname}")
    # by(cat, where(...)): per category plane, row ids of the whole frame; by(cat, mean) next to a 2-stage member
    for aname, agg in {"by_where_first_row": ds.by("cat", ds.where(ds.first("val"))),
                       "by_where_last_other": ds.by("cat", ds.where(ds.last("val"), "other")),
                       "by_where_max_row": ds.by("cat", ds.where(ds.max("val"))),
                       "by_where_min_other": ds.by("cat", ds.where(ds.min("val"), "other"))}.items():
        got, want = cvs.line(frame, agg=agg, **kw).data, g[f"aa3_{aname}"]
        assert got.shape == want.shape and got.dtype == want.dtype and np.array_equal(got, want, equal_nan=got.dtype.kind == "f"), aname
    res = cvs.line(frame, agg=ds.summary(m=ds.by("cat", ds.mean("val")), mn=ds.min("val")), **kw)
    _cmp_aa(res["m"].data, g["aa3_s9_m"], "summary s9.m (by(cat, mean) next to min)")
    _cmp_aa(res["mn"].data, g["aa3_s9_mn"], "summary s9.mn")
    for aname, agg in {"where_first_row": ds.where(ds.first("val")), "where_first_other": ds.where(ds.first("val"), "other"),
                       "where_last_row": ds.where(ds.last("val")), "where_last_other": ds.where(ds.last("val"), "other"),
                       "where_max_row": ds.where(ds.max("val")), "where_max_other": ds.where(ds.max("val"), "other"),
                       "where_min_row": ds.where(ds.min("val")), "where_min_other": ds.where(ds.min("val"), "other")}.items():
        got, want = cvs.line(frame, agg=agg, **kw).data, g[f"aa3_{aname}"]
        assert got.dtype == want.dtype and np.array_equal(got, want, equal_nan=got.dtype.kind == "f"), aname


def test_antialiased_lines_at_production_geometry_vs_oracle():
    """BASELINE config 4's geometry - 3840 x 2160, LinesAxis1 random walks of 1000 samples, line_width 1 - on a 600-line
    sample: antialiased max (single stage) and min (2-stage, lines long enough to overflow the shared-memory stage-1 table)
    against the C oracle."""
    import torch
    import datashader_b200 as ds
    from oracle import oracle as ora
    rng = np.random.default_rng(404)
    nl, nv = 600, 1000
    xs = np.tile(np.arange(nv, dtype=np.float32), (nl, 1))
    ys = np.cumsum(rng.standard_normal((nl, nv)), axis=1).astype(np.float32)
    val = rng.random(nl).astype(np.float32)
    cols = {f"x{j}": torch.from_numpy(np.ascontiguousarray(xs[:, j])).cuda() for j in range(nv)}
    cols.update({f"y{j}": torch.from_numpy(np.ascontiguousarray(ys[:, j])).cuda() for j in range(nv)})
    cols["value"] = torch.from_numpy(val).cuda()
    frame = ds.DeviceFrame(cols)
    xr, yr = (0.0, float(nv - 1)), (float(ys.min()), float(ys.max()))
    W, H = 3840, 2160
    cvs = ds.Canvas(W, H, x_range=xr, y_range=yr)
    view = ora.make_view(W, H, xr, yr)
    xc, yc = [f"x{j}" for j in range(nv)], [f"y{j}" for j in range(nv)]
    got = cvs.line(frame, x=xc, y=yc, axis=1, agg=ds.max("value"), line_width=1).data
    _cmp_aa(got, ora.lines_axis1(xs, ys, view, agg="max", values=val, line_width=1.0), "config-4 geometry, aa max")
    got = cvs.line(frame, x=xc, y=yc, axis=1, agg=ds.min("value"), line_width=1).data
    _cmp_aa(got, ora.lines_aa2(xs, ys, view, "min", val, 1.0), "config-4 geometry, aa min (2-stage)")


def test_balanced_antialiased_kernel_equals_per_thread_kernel():
    """k_lines_aa_balanced (rows of the scan conversion handed out evenly over the warp) against k_lines_axis1 (one thread
    per segment): identical pixels and values for max / any (bit for bit), count / sum / mean within float-add ordering."""
    import torch
    import datashader_b200 as ds
    from datashader_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(77)
    nl, nv = 700, 60
    xs = (np.cumsum(rng.normal(0, 0.02, (nl, nv)), axis=1) + rng.random((nl, 1))).astype(np.float32)
    ys = (np.cumsum(rng.normal(0, 0.03, (nl, nv)), axis=1) + rng.random((nl, 1))).astype(np.float32)
    xs[rng.random((nl, nv)) < 0.02] = np.nan
    val = (rng.random(nl) * 5 - 1).astype(np.float32)
    cols = {f"x{j}": torch.from_numpy(np.ascontiguousarray(xs[:, j])).cuda() for j in range(nv)}
    cols.update({f"y{j}": torch.from_numpy(np.ascontiguousarray(ys[:, j])).cuda() for j in range(nv)})
    cols["val"] = torch.from_numpy(val).cuda()
    frame = ds.DeviceFrame(cols)
    xc, yc = [f"x{j}" for j in range(nv)], [f"y{j}" for j in range(nv)]
    cvs = ds.Canvas(500, 300, x_range=(0.1, 0.9), y_range=(0.0, 1.2))
    for lw in (1.0, 3.5, 0.5):
        for name, agg, exact in (("max", ds.max("val"), True), ("any", ds.any(), True), ("count", ds.count(), False),
                                 ("sum", ds.sum("val"), False), ("mean", ds.mean("val"), False)):
            res = {}
            for balanced in (1, 0):
                _lib.check(L.dsb_lines_configure(balanced))
                try:
                    res[balanced] = cvs.line(frame, x=xc, y=yc, axis=1, agg=agg, line_width=lw).data
                    assert (b"balanced" in L.dsb_last_kernel()) == bool(balanced)
                finally:
                    _lib.check(L.dsb_lines_configure(1))
            if exact:
                assert np.array_equal(res[1], res[0], equal_nan=True), (name, lw)
            else:
                _cmp_aa(res[1], res[0], f"balanced {name} lw{lw}", rtol=1e-5)
