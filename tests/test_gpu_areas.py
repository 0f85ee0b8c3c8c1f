"""Canvas.area on the GPU against the real reference (tests/golden/areas.npz): the ten non-ragged layouts,
to-zero and to-line, integer scan fill -> bit-exact (float sums to 1e-12)."""
import numpy as np
import pytest

from helpers import load

pytestmark = pytest.mark.gpu


def _cmp(got, want, key, atol=0.0):
    """atol: float sums of values that cancel are only reproducible to the accumulation order (f64 atomics)."""
    assert got.dtype == want.dtype and got.shape == want.shape, key
    if "sum" in key or "mean" in key:
        assert np.array_equal(np.isnan(got), np.isnan(want)), key
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=atol, equal_nan=True, err_msg=key)
    else:
        assert np.array_equal(got, want, equal_nan=(got.dtype.kind == "f")), key


def test_area_axis0_layouts():
    import pandas as pd
    import datashader_b200 as ds
    g = load("areas.npz")
    df0 = pd.DataFrame({k: g[f"a0_{k}"] for k in ("x", "y", "ys", "x2", "y2", "y2s", "val")})
    canvases = {"fixed": ds.Canvas(plot_width=45, plot_height=35, x_range=(0, 1), y_range=(-0.5, 1.0)),
                "auto": ds.Canvas(plot_width=33, plot_height=27)}
    aggs = {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val"), "first": ds.first("val")}
    for cn, cvs in canvases.items():
        for an, agg in aggs.items():
            _cmp(cvs.area(df0, "x", "y", agg=agg).data, g[f"a0_zero_{cn}_{an}"], f"a0 zero {cn} {an}")
            _cmp(cvs.area(df0, "x", "y", agg=agg, y_stack="ys").data, g[f"a0_line_{cn}_{an}"], f"a0 line {cn} {an}")
            if an == "first":
                # Multi layouts: the serial CPU loop visits column pairs one after the other, so its `first` is
                # "first in (column, row) order"; the GPU follows the reference's row-index formulation (min row id,
                # reductions.py:1346-1360), which is what its own CUDA / dask paths compute.  Not comparable.
                continue
            _cmp(cvs.area(df0, x=["x", "x2"], y=["y", "y2"], agg=agg, axis=0).data, g[f"a0m_zero_{cn}_{an}"], f"a0m zero {cn} {an}")
            _cmp(cvs.area(df0, x=["x", "x2"], y=["y", "y2"], y_stack=["ys", "y2s"], agg=agg, axis=0).data,
                 g[f"a0m_line_{cn}_{an}"], f"a0m line {cn} {an}")
        r = cvs.area(df0, "x", "y")
        np.testing.assert_array_equal(np.asarray(r.attrs["y_range"], dtype="f8"), g[f"a0_zero_{cn}_yrange"])
        assert r.data.dtype == np.bool_ and tuple(r.dims) == ("y", "x")


def test_area_axis1_layouts():
    import pandas as pd
    import datashader_b200 as ds
    g = load("areas.npz")
    xm, ym, ysm, lval = g["a1_x"], g["a1_y"], g["a1_ys"], g["a1_val"]
    nv = xm.shape[1]
    d = {f"x{j}": xm[:, j] for j in range(nv)}
    d.update({f"y{j}": ym[:, j] for j in range(nv)})
    d.update({f"s{j}": ysm[:, j] for j in range(nv)})
    d["val"] = lval
    df1 = pd.DataFrame(d)
    xc, yc, sc = [f"x{j}" for j in range(nv)], [f"y{j}" for j in range(nv)], [f"s{j}" for j in range(nv)]
    xconst, yconst, sconst = g["a1_xconst"], g["a1_yconst"], g["a1_sconst"]
    cvs = ds.Canvas(plot_width=45, plot_height=35, x_range=(0, 1), y_range=(-0.5, 1.0))
    aggs1 = {"any": ds.any(), "count": ds.count(), "max": ds.max("val"), "mean": ds.mean("val")}
    for an, agg in aggs1.items():
        _cmp(cvs.area(df1, x=xc, y=yc, agg=agg, axis=1).data, g[f"a1_zero_{an}"], f"a1 zero {an}")
        _cmp(cvs.area(df1, x=xc, y=yc, y_stack=sc, agg=agg, axis=1).data, g[f"a1_line_{an}"], f"a1 line {an}")
        _cmp(cvs.area(df1, x=xconst, y=yc, agg=agg, axis=1).data, g[f"a1xc_zero_{an}"], f"a1xc zero {an}")
        _cmp(cvs.area(df1, x=xconst, y=yc, y_stack=sc, agg=agg, axis=1).data, g[f"a1xc_line_{an}"], f"a1xc line {an}")
        _cmp(cvs.area(df1, x=xc, y=yconst, agg=agg, axis=1).data, g[f"a1yc_zero_{an}"], f"a1yc zero {an}")
        _cmp(cvs.area(df1, x=xc, y=yconst, y_stack=sconst, agg=agg, axis=1).data, g[f"a1yc_line_{an}"], f"a1yc line {an}")


@pytest.mark.parametrize("dtype", ["f4", "f8"])
def test_areas_vs_oracle_random(dtype):
    """Seeded random series (NaN gaps, clipping canvas, float32 and float64 columns) against the C oracle's
    restatement of draw_trapezoid_y; integer scan fill -> bit-exact."""
    import pandas as pd
    import datashader_b200 as ds
    from oracle import oracle as ora
    rng = np.random.default_rng(2024)
    n = 4000
    x = np.sort(rng.uniform(-0.2, 1.2, n)).astype(dtype)
    y = (np.sin(np.linspace(0, 40, n)) * 0.6 + rng.normal(0, 0.05, n)).astype(dtype)
    ys = (y - np.abs(rng.normal(0.2, 0.1, n))).astype(dtype)
    x2 = np.sort(rng.uniform(0, 1, n)).astype(dtype)
    y2 = rng.uniform(-1, 1, n).astype(dtype)
    y2s = (y2 * 0.5).astype(dtype)
    for a in (y, y2, ys):
        a[rng.integers(0, n, 25)] = np.nan
    val = rng.normal(size=n)
    df = pd.DataFrame({"x": x, "y": y, "ys": ys, "x2": x2, "y2": y2, "y2s": y2s, "val": val})
    W, H = 300, 200
    cvs = ds.Canvas(plot_width=W, plot_height=H, x_range=(0.1, 0.9), y_range=(-0.4, 0.5))
    view = ora.make_view(W, H, (0.1, 0.9), (-0.4, 0.5))
    for name, agg in {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val"), "min": ds.min("val")}.items():
        vals = None if name in ("any", "count") else val
        _cmp(cvs.area(df, "x", "y", agg=agg).data, ora.areas(x[None], y[None], view, None, name, vals, per_vertex=True), f"zero {name}", 1e-11)
        _cmp(cvs.area(df, "x", "y", y_stack="ys", agg=agg).data, ora.areas(x[None], y[None], view, ys[None], name, vals, per_vertex=True),
             f"line {name}", 1e-11)
        _cmp(cvs.area(df, x=["x", "x2"], y=["y", "y2"], agg=agg, axis=0).data,
             ora.areas(np.stack([x, x2]), np.stack([y, y2]), view, None, name, vals, per_vertex=True), f"multi zero {name}", 1e-11)
        _cmp(cvs.area(df, x=["x", "x2"], y=["y", "y2"], y_stack=["ys", "y2s"], agg=agg, axis=0).data,
             ora.areas(np.stack([x, x2]), np.stack([y, y2]), view, np.stack([ys, y2s]), name, vals, per_vertex=True), f"multi line {name}", 1e-11)


def test_areas_axis1_vs_oracle_random():
    import pandas as pd
    import datashader_b200 as ds
    from oracle import oracle as ora
    rng = np.random.default_rng(77)
    nl, nv = 500, 12
    xm = np.sort(rng.uniform(-0.1, 1.1, (nl, nv)), axis=1).astype("f4")
    ym = rng.uniform(-0.8, 1.2, (nl, nv)).astype("f4")
    sm = (ym - rng.uniform(0, 0.3, (nl, nv))).astype("f4")
    ym[rng.integers(0, nl, 30), rng.integers(0, nv, 30)] = np.nan
    lval = rng.normal(size=nl)
    d = {f"x{j}": xm[:, j] for j in range(nv)}
    d.update({f"y{j}": ym[:, j] for j in range(nv)})
    d.update({f"s{j}": sm[:, j] for j in range(nv)})
    d["val"] = lval
    df = pd.DataFrame(d)
    xc, yc, sc = ([f"{p}{j}" for j in range(nv)] for p in "xys")
    xconst = np.linspace(0, 1, nv)
    W, H = 160, 120
    cvs = ds.Canvas(plot_width=W, plot_height=H, x_range=(0, 1), y_range=(-0.5, 1.0))
    view = ora.make_view(W, H, (0, 1), (-0.5, 1.0))
    for name, agg in {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val")}.items():
        vals = None if name in ("any", "count") else lval
        _cmp(cvs.area(df, x=xc, y=yc, agg=agg, axis=1).data, ora.areas(xm, ym, view, None, name, vals), f"a1 zero {name}", 1e-11)
        _cmp(cvs.area(df, x=xc, y=yc, y_stack=sc, agg=agg, axis=1).data, ora.areas(xm, ym, view, sm, name, vals), f"a1 line {name}", 1e-11)
        # x constant: numpy float64 x against float32 y columns -> the reference promotes nothing: xs stay f64, ys f32;
        # the oracle takes one dtype for all vertex arrays, so compare on float64 copies of the same values.
        df64 = df.astype("f8")
        _cmp(cvs.area(df64, x=xconst, y=yc, agg=agg, axis=1).data, ora.areas(xconst, ym.astype("f8"), view, None, name, vals), f"a1xc zero {name}", 1e-11)
