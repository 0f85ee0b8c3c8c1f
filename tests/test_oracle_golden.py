"""The C oracle against the real reference: committed golden vectors (tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference) and the literal known-answer tables of the
reference's own tests (datashader/tests/test_pandas.py)."""
import numpy as np
import pytest

from oracle import oracle as ora
from helpers import (CANVASES, LINE_CANVASES, SPECS, assert_agg_equal, columns_from_golden, load)


def _view(cols, ckw, x="x", y="y"):
    xr = ckw.get("x_range") or ora.compute_bounds(cols[x])
    yr = ckw.get("y_range") or ora.compute_bounds(cols[y])
    return ora.make_view(ckw["plot_width"], ckw["plot_height"], xr, yr), xr, yr


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("cname", list(CANVASES))
def test_points_golden(tag, cname):
    g = load("points.npz")
    cols = columns_from_golden(g, f"in_{tag}_")
    view, xr, yr = _view(cols, CANVASES[cname])
    n = 0
    for rname, spec in SPECS.items():
        key = f"pts_{tag}_{cname}_{rname}"
        if key not in g.files:
            continue
        assert_agg_equal(ora.points(cols, "x", "y", spec, view), g[key], key)
        n += 1
    assert n >= 5
    # coords and ranges (core.py:83-99, pandas.py:61-66)
    np.testing.assert_array_equal(ora.axis_index((view.sx, view.tx), view.width), g[f"pts_{tag}_{cname}_xcoords"])
    np.testing.assert_array_equal(ora.axis_index((view.sy, view.ty), view.height), g[f"pts_{tag}_{cname}_ycoords"])
    np.testing.assert_array_equal(np.asarray(xr, dtype=np.float64), g[f"pts_{tag}_{cname}_xrange"])
    np.testing.assert_array_equal(np.asarray(yr, dtype=np.float64), g[f"pts_{tag}_{cname}_yrange"])


def test_points_log_axes_golden():
    g = load("points.npz")
    cols = {k: g[f"in_log_{k}"] for k in ("x", "y", "v32")}
    view = ora.make_view(40, 30, (1, 1000), (1, 100), "log", "log")
    assert_agg_equal(ora.points(cols, "x", "y", ("count",), view), g["pts_log_count"], "log count")
    assert_agg_equal(ora.points(cols, "x", "y", ("max", "v32"), view), g["pts_log_max_v32"], "log max")
    np.testing.assert_allclose(ora.axis_index((view.sx, view.tx), 40, "log"), g["pts_log_xcoords"], rtol=1e-15)
    np.testing.assert_allclose(ora.axis_index((view.sy, view.ty), 30, "log"), g["pts_log_ycoords"], rtol=1e-15)


NEGZERO = {"max_vmax32": ("max", "vmax32"), "min_vmin32": ("min", "vmin32"), "max_vmax64": ("max", "vmax64"),
           "min_vmin64": ("min", "vmin64"), "by_max_vmax32": ("by", "cat", ("max", "vmax32")),
           "max_vmin32": ("max", "vmin32"), "min_vmax32": ("min", "vmax32")}


def test_points_negzero_golden():
    """max / min keep the zero that arrived first (-0.0 or +0.0): compared as bit patterns (helpers.assert_agg_equal)."""
    g = load("points_negzero.npz")
    cols = columns_from_golden(g, "in_")
    view = ora.make_view(9, 7, (0, 1), (0, 1))
    for name, spec in NEGZERO.items():
        want = g[f"nz_{name}"]
        assert np.signbit(want[want == 0]).any() or "vmin32" in name and "max" in name or "vmax32" in name and "min" in name
        assert_agg_equal(ora.points(cols, "x", "y", spec, view), want, f"negzero {name}")


@pytest.mark.parametrize("nparts", [1, 3, 4])
def test_partitioned_golden(nparts):
    """dask-style partition + combine (data_libraries/dask.py:168-217) equals the single pass."""
    g = load("partitioned.npz")
    cols = columns_from_golden(g, "in_")
    view = ora.make_view(37, 23, (-0.1, 1.05), (0.1, 0.9))
    for key in g.files:
        if not key.startswith("part3_"):
            continue
        rname = key[len("part3_"):]
        got = ora.points(cols, "x", "y", SPECS[rname], view, npartitions=nparts)
        assert_agg_equal(got, g[key], f"{key} nparts={nparts}")


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("cname", list(LINE_CANVASES))
def test_lines_golden(tag, cname):
    g = load("lines.npz")
    xs, ys, val = g[f"in_{tag}_xs"], g[f"in_{tag}_ys"], g[f"in_{tag}_val"]
    ckw = LINE_CANVASES[cname]
    view = ora.make_view(ckw["plot_width"], ckw["plot_height"], ckw["x_range"], ckw["y_range"])
    n = 0
    for key in g.files:
        pre = f"ln_{tag}_{cname}_lw"
        if not key.startswith(pre):
            continue
        lw, aname = key[len(pre):].split("_")
        lw = float(lw)
        values = None if aname in ("any", "count") else val
        got = ora.lines_axis1(xs, ys, view, agg=aname, values=values, line_width=lw)
        want = g[key]
        assert got.dtype == want.dtype and got.shape == want.shape, key
        if lw > 0 and aname in ("count", "sum"):
            np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6, equal_nan=True, err_msg=key)
        else:
            assert np.array_equal(got, want, equal_nan=(got.dtype.kind == "f")), key
        n += 1
    assert n >= 8


# ---- the reference's own literal tables (datashader/tests/test_pandas.py:27-59, 100, 194-243, 698-738)
def _ref_fixture():
    nan = np.nan
    cols = {
        "x": np.array(([0.] * 10 + [1] * 10)),
        "y": np.array(([0.] * 5 + [1] * 5 + [0] * 5 + [1] * 5)),
        "log_x": np.array(([1.] * 10 + [10] * 10)),
        "log_y": np.array(([1.] * 5 + [10] * 5 + [1] * 5 + [10] * 5)),
        "i32": np.arange(20, dtype="i4"), "i64": np.arange(20, dtype="i8"),
        "f32": np.arange(20, dtype="f4"), "f64": np.arange(20, dtype="f8"),
        "reverse": np.arange(20, 0, -1, dtype="f8"),
        "plusminus": np.arange(20, dtype="f8") * ([1, -1] * 10),
        "cat": np.array([0] * 5 + [1] * 5 + [2] * 5 + [3] * 5, dtype=np.int8), "cat__ncat": 4,
    }
    for c in ("f32", "f64", "reverse", "plusminus"):
        cols[c] = cols[c].copy()
        cols[c][2] = nan
    cols["f64"][2] = nan
    return cols


def test_reference_known_answers():
    cols = _ref_fixture()
    view = ora.make_view(2, 2, (0, 1), (0, 1))
    nan = np.nan
    pts = lambda spec: ora.points(cols, "x", "y", spec, view)   # noqa: E731
    np.testing.assert_array_equal(pts(("count",)), np.array([[5, 5], [5, 5]], dtype="u4"))          # :194-203
    np.testing.assert_array_equal(pts(("count", "f32")), np.array([[4, 5], [5, 5]], dtype="u4"))
    np.testing.assert_array_equal(pts(("any", "f64")), np.array([[True, True], [True, True]]))     # :206-214
    s = pts(("sum", "i32"))                                                                            # :217-225
    np.testing.assert_array_equal(s, cols["i32"].reshape(2, 2, 5).sum(axis=2, dtype="f8").T)
    np.testing.assert_array_equal(pts(("sum", "f64")), np.nansum(cols["f64"].reshape(2, 2, 5), axis=2).T)
    np.testing.assert_array_equal(pts(("min", "f32")), np.nanmin(cols["f64"].reshape(2, 2, 5), axis=2).T)   # :228-234
    np.testing.assert_array_equal(pts(("max", "f32")), np.nanmax(cols["f64"].reshape(2, 2, 5), axis=2).T)   # :237-243
    np.testing.assert_allclose(pts(("mean", "f32")), np.nanmean(cols["f64"].reshape(2, 2, 5), axis=2).T)  # :698-706
    sol = np.array([[[5, 0, 0, 0], [0, 0, 5, 0]], [[0, 5, 0, 0], [0, 0, 0, 5]]], dtype="u4")          # :731-738
    np.testing.assert_array_equal(pts(("by", "cat", ("count",))), sol)
    np.testing.assert_array_equal(pts(("first", "f32")), np.array([[0, 10], [5, 15]], dtype="f8"))     # :991-996
    np.testing.assert_array_equal(pts(("last", "f32")), np.array([[4, 14], [9, 19]], dtype="f8"))      # :999-1004
    np.testing.assert_array_equal(pts(("where", ("first", "f32"), None)), np.array([[0, 10], [5, 15]]))  # :448-462
    np.testing.assert_array_equal(pts(("where", ("last", "f32"), None)), np.array([[4, 14], [9, 19]]))   # :465-479
    np.testing.assert_array_equal(pts(("where", ("max", "f32"), None)), np.array([[4, 14], [9, 19]]))    # :482-494
    np.testing.assert_array_equal(pts(("where", ("min", "f32"), None)), np.array([[0, 10], [5, 15]]))    # :497-509
    np.testing.assert_array_equal(pts(("where", ("max", "f32"), "reverse")), np.array([[16, 6], [11, 1]], dtype="f8"))
    np.testing.assert_array_equal(pts(("where", ("first", "f32"), "reverse")), np.array([[20, 10], [15, 5]], dtype="f8"))
    _ = nan


def test_reference_uniform_points_upper_edge_fold():
    """test_pandas.py:1132-1144: 101 points over [0,100] into 10 bins -> 10,...,10,11."""
    n = 101
    cols = {"x": np.arange(n, dtype="f8"), "y": np.zeros(n)}
    view = ora.make_view(10, 1, (0, 100), (-1, 1))
    got = ora.points(cols, "x", "y", ("count",), view)
    np.testing.assert_array_equal(got, np.array([[10] * 9 + [11]], dtype="u4"))


def test_no_fma_contraction_in_mapping():
    """x*sx+tx must be evaluated unfused like numba does. Find inputs where the fused
    and unfused results truncate to different pixels and check the oracle takes the unfused one."""
    rng = np.random.default_rng(5)
    W = 900
    xr = (-0.3337, 1.2771)
    s, t = ora.scale_and_translate(xr, W)
    x = (rng.random(2_000_000) * (xr[1] - xr[0]) + xr[0]).astype(np.float32).astype(np.float64)
    unfused = (x * s + t).astype(np.int64)
    import math
    # exact fused value via integer-exact long double is not available in numpy: use fractions on candidates
    near = np.nonzero(np.abs((x * s + t) - np.rint(x * s + t)) < 1e-12)[0]
    from fractions import Fraction
    diff = []
    for i in near:
        exact = Fraction(float(x[i])) * Fraction(s) + Fraction(t)
        fused = int(math.floor(float(exact))) if exact >= 0 else int(float(exact))
        if fused != unfused[i]:
            diff.append(i)
    xs = x[diff] if diff else x[near[:8]]
    cols = {"x": xs.astype(np.float32), "y": np.zeros(len(xs), dtype=np.float32)}
    view = ora.make_view(W, 1, xr, (-1, 1))
    got = ora.points(cols, "x", "y", ("count",), view)
    want = np.bincount(np.minimum((xs * s + t).astype(np.int64), W - 1), minlength=W).astype("u4")[None, :]
    np.testing.assert_array_equal(got, want)


def _eq(got, want, key, tol=False):
    assert got.dtype == want.dtype and got.shape == want.shape, key
    if tol:
        np.testing.assert_allclose(got, want, rtol=1e-12, equal_nan=True, err_msg=key)
    else:
        assert np.array_equal(got, want, equal_nan=(got.dtype.kind == "f")), key


def test_line_layouts_golden():
    """LineAxis0 / LineAxis0Multi / LinesAxis1XConstant / YConstant restated in the oracle vs the real reference."""
    g = load("line_layouts.npz")
    view = ora.make_view(50, 40, (0, 1), (0, 1))
    x, y, x2, y2, val = (g[f"ax0_{k}"] for k in ("x", "y", "x2", "y2", "val"))
    for name in ("any", "count", "sum", "max"):
        vals = None if name in ("any", "count") else val
        _eq(ora.lines(x[None, :], y[None, :], view, name, vals, 0, per_vertex=True), g[f"ax0_lw0_{name}"], f"ax0 {name}", name == "sum")
        _eq(ora.lines(np.stack([x, x2]), np.stack([y, y2]), view, name, vals, 0, per_vertex=True), g[f"ax0multi_lw0_{name}"],
            f"ax0multi {name}", name == "sum")
    for name in ("any", "max"):
        vals = None if name == "any" else val
        _eq(ora.lines(x[None, :], y[None, :], view, name, vals, 1.0, per_vertex=True), g[f"ax0_lw1_{name}"], f"ax0 aa {name}")
        _eq(ora.lines(np.stack([x, x2]), np.stack([y, y2]), view, name, vals, 1.0, per_vertex=True), g[f"ax0multi_lw1_{name}"],
            f"ax0multi aa {name}")
    xc, ys, lval = g["xc_x"], g["xc_ys"], g["xc_val"]
    for name in ("any", "count", "max"):
        vals = None if name in ("any", "count") else lval
        _eq(ora.lines(xc, ys, view, name, vals, 0), g[f"xconst_lw0_{name}"], f"xconst {name}")
        _eq(ora.lines(ys, xc, view, name, vals, 0), g[f"yconst_lw0_{name}"], f"yconst {name}")
    _eq(ora.lines(xc, ys, view, "max", lval, 1.0), g["xconst_lw1_max"], "xconst aa")
    _eq(ora.lines(ys, xc, view, "max", lval, 1.0), g["yconst_lw1_max"], "yconst aa")


def test_areas_golden():
    """The area layouts restated in the oracle (ds_oracle_lines.c: draw_trapezoid_y) vs the real reference."""
    g = load("areas.npz")
    x, y, ys, x2, y2, y2s, val = (g[f"a0_{k}"] for k in ("x", "y", "ys", "x2", "y2", "y2s", "val"))
    fixed = ora.make_view(45, 35, (0, 1), (-0.5, 1.0))
    xr = ora.compute_bounds(x)
    b = ora.compute_bounds(y)
    auto = ora.make_view(33, 27, xr, (min(b[0], 0), max(b[1], 0)))          # area.py:71-79: bounds include zero
    np.testing.assert_array_equal(np.array([min(b[0], 0), max(b[1], 0)]), g["a0_zero_auto_yrange"])
    for name in ("any", "count", "sum", "max"):
        vals = None if name in ("any", "count") else val
        tol = name == "sum"
        _eq(ora.areas(x[None, :], y[None, :], fixed, None, name, vals, per_vertex=True), g[f"a0_zero_fixed_{name}"], f"a0 zero {name}", tol)
        _eq(ora.areas(x[None, :], y[None, :], auto, None, name, vals, per_vertex=True), g[f"a0_zero_auto_{name}"], f"a0 zero auto {name}", tol)
        _eq(ora.areas(x[None, :], y[None, :], fixed, ys[None, :], name, vals, per_vertex=True), g[f"a0_line_fixed_{name}"], f"a0 line {name}", tol)
        _eq(ora.areas(np.stack([x, x2]), np.stack([y, y2]), fixed, None, name, vals, per_vertex=True), g[f"a0m_zero_fixed_{name}"],
            f"a0m zero {name}", tol)
        _eq(ora.areas(np.stack([x, x2]), np.stack([y, y2]), fixed, np.stack([ys, y2s]), name, vals, per_vertex=True),
            g[f"a0m_line_fixed_{name}"], f"a0m line {name}", tol)
    xm, ym, ysm, lval = g["a1_x"], g["a1_y"], g["a1_ys"], g["a1_val"]
    xconst, yconst, sconst = g["a1_xconst"], g["a1_yconst"], g["a1_sconst"]
    for name in ("any", "count", "max"):
        vals = None if name in ("any", "count") else lval
        _eq(ora.areas(xm, ym, fixed, None, name, vals), g[f"a1_zero_{name}"], f"a1 zero {name}")
        _eq(ora.areas(xm, ym, fixed, ysm, name, vals), g[f"a1_line_{name}"], f"a1 line {name}")
        _eq(ora.areas(xconst, ym, fixed, None, name, vals), g[f"a1xc_zero_{name}"], f"a1xc zero {name}")
        _eq(ora.areas(xconst, ym, fixed, ysm, name, vals), g[f"a1xc_line_{name}"], f"a1xc line {name}")
        _eq(ora.areas(xm, yconst, fixed, None, name, vals), g[f"a1yc_zero_{name}"], f"a1yc zero {name}")
        _eq(ora.areas(xm, yconst, fixed, sconst, name, vals), g[f"a1yc_line_{name}"], f"a1yc line {name}")


def test_lines_aa2_golden():
    """2-stage antialiased reductions (min / first / last / count, sum with self_intersect=False) vs the reference."""
    g, gl, gx = load("lines_aa2.npz"), load("lines.npz"), load("line_layouts.npz")
    view = ora.make_view(64, 48, (0, 1), (0, 1))
    names = {"min": "min", "first": "first", "last": "last", "sum_nsi": "sum", "count_nsi": "count", "count_val_nsi": "count",
             "mean": "mean"}
    for tag in ("f32", "f64"):
        xs, ys, val = gl[f"in_{tag}_xs"], gl[f"in_{tag}_ys"], gl[f"in_{tag}_val"]
        for lw in ((1, 2.5) if tag == "f32" else (1,)):
            for gname, oname in names.items():
                vals = None if gname == "count_nsi" else val
                got = ora.lines(xs, ys, view, "mean", val, lw) if oname == "mean" else ora.lines_aa2(xs, ys, view, oname, vals, lw)
                want = g[f"aa2_{tag}_lw{lw}_{gname}"]
                assert got.dtype == want.dtype, gname
                assert np.array_equal(np.isnan(got), np.isnan(want)), (tag, lw, gname)
                np.testing.assert_allclose(got, want, rtol=1e-6 if oname == "count" else 1e-12, equal_nan=True, err_msg=f"{tag} {lw} {gname}")
    view0 = ora.make_view(50, 40, (0, 1), (0, 1))
    x, y, x2, y2, val = (gx[f"ax0_{k}"] for k in ("x", "y", "x2", "y2", "val"))
    for gname, oname in names.items():
        vals = None if gname == "count_nsi" else val
        for key, xa, ya in (("ax0", x[None], y[None]), ("ax0multi", np.stack([x, x2]), np.stack([y, y2]))):
            got = (ora.lines(xa, ya, view0, "mean", val, 2, per_vertex=True) if oname == "mean"
                   else ora.lines_aa2(xa, ya, view0, oname, vals, 2, per_vertex=True))
            want = g[f"aa2_{key}_lw2_{gname}"]
            assert got.dtype == want.dtype and np.array_equal(np.isnan(got), np.isnan(want)), (key, gname)
            np.testing.assert_allclose(got, want, rtol=1e-6 if oname == "count" else 1e-12, equal_nan=True, err_msg=f"{key} {gname}")


def test_lines_aa_by_category_golden():
    """Antialiased by('cat', r): the oracle's per-category restatement vs the reference."""
    g, gl, ge = load("lines_aa2.npz"), load("lines.npz"), load("lines_extra.npz")
    view = ora.make_view(64, 48, (0, 1), (0, 1))
    xs, ys, val, codes = gl["in_f32_xs"], gl["in_f32_ys"], gl["in_f32_val"], ge["in_cat"]
    for aname in ("any", "count", "sum", "max", "mean"):
        got = ora.lines_by(xs, ys, view, codes, 4, aname, None if aname in ("any", "count") else val, 2)
        want = g[f"aaby_lw2_{aname}"]
        assert got.dtype == want.dtype and got.shape == want.shape, aname
        assert np.array_equal(np.isnan(got), np.isnan(want)), aname
        np.testing.assert_allclose(got, want, rtol=1e-6 if aname == "count" else 1e-12, equal_nan=True, err_msg=aname)


def test_device_log10f_restatement_equals_the_c_library():
    """datashader_b200/csrc/log10f_glibc.h (what the device evaluates for LogAxis on float32 coordinates) against this box's
    libm, every 61st positive finite float (the full sweep, `oracle/_build/log10f_check 1`, takes 10 s on 8 cores: 0 of
    2 139 095 039 differ on glibc 2.39)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_build", "log10f_check")
    if not os.path.exists(exe):
        ora.build()
    r = subprocess.run([exe, "61"], capture_output=True, text=True)
    assert r.returncode == 0 and " 0 differ" in r.stdout, r.stdout + r.stderr


def _ragged_inputs(g, tag):
    return {k: (g[f"{tag}_{k}_flat"], g[f"{tag}_{k}_starts"]) for k in ("x", "y", "s")}


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_ragged_golden(tag):
    """LinesAxis1Ragged / AreaToZeroAxis1Ragged / AreaToLineAxis1Ragged restated in the oracle (row loops of line.py:1559-1600,
    area.py:1959-2004, 2033-2081) vs the real reference (tests/golden/ragged.npz: uneven rows, an empty and a one-vertex row,
    x / y / stack rows of different lengths, NaN vertices)."""
    g = load("ragged.npz")
    r = _ragged_inputs(g, tag)
    (xf, xs), (yf, ys), (sf, ss) = r["x"], r["y"], r["s"]
    val = g["val"]
    view = ora.make_view(48, 36, (0, 1), (-0.2, 1.1))
    for name in ("any", "count", "sum", "max", "min"):
        vals = None if name in ("any", "count") else val
        _eq(ora.lines_ragged(xf, xs, yf, ys, view, name, vals, 0), g[f"{tag}_line_lw0_{name}"], f"ragged lw0 {name}", name == "sum")
    for lw in (1, 2.5):
        for name in ("any", "count", "sum", "max", "mean"):
            vals = None if name in ("any", "count") else val
            got, want = ora.lines_ragged(xf, xs, yf, ys, view, name, vals, lw), g[f"{tag}_line_lw{lw}_{name}"]
            assert got.dtype == want.dtype and np.array_equal(np.isnan(got), np.isnan(want)), (lw, name)
            np.testing.assert_allclose(got, want, rtol=1e-6 if name in ("any", "count") else 1e-12, equal_nan=True, err_msg=f"{lw} {name}")
        for gname, oname in (("min", "min"), ("first", "first"), ("last", "last"), ("sum_nsi", "sum"), ("count_nsi", "count")):
            got = ora.lines_ragged_aa2(xf, xs, yf, ys, view, oname, None if gname == "count_nsi" else val, lw)
            want = g[f"{tag}_line_lw{lw}_{gname}"]
            assert got.dtype == want.dtype and np.array_equal(np.isnan(got), np.isnan(want)), (lw, gname)
            np.testing.assert_allclose(got, want, rtol=1e-6 if oname == "count" else 1e-12, equal_nan=True, err_msg=f"{lw} {gname}")
    # auto ranges: bounds of the flat arrays (line.py:472-478); area-to-zero includes y = 0 (area.py:932-939)
    xr, yr = ora.compute_bounds(xf), ora.compute_bounds(yf)
    np.testing.assert_array_equal(np.array(list(xr) + list(yr)), g[f"{tag}_line_auto_ranges"])
    _eq(ora.lines_ragged(xf, xs, yf, ys, ora.make_view(31, 23, xr, yr), "count", None, 0), g[f"{tag}_line_auto_count"], "auto count")
    for name in ("any", "count", "sum", "max"):
        vals = None if name in ("any", "count") else val
        _eq(ora.areas_ragged(xf, xs, yf, ys, view, None, None, name, vals), g[f"{tag}_area_zero_{name}"], f"area zero {name}", name == "sum")
        _eq(ora.areas_ragged(xf, xs, yf, ys, view, sf, ss, name, vals), g[f"{tag}_area_line_{name}"], f"area line {name}", name == "sum")
    yz = (min(yr[0], 0), max(yr[1], 0))
    np.testing.assert_array_equal(np.array(list(xr) + list(yz)), g[f"{tag}_area_zero_auto_ranges"])
    _eq(ora.areas_ragged(xf, xs, yf, ys, ora.make_view(31, 23, xr, yz), None, None, "count"), g[f"{tag}_area_zero_auto_count"], "area auto")
