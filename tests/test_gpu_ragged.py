"""The ragged layouts on the GPU: LinesAxis1Ragged (line.py:457-523, 1538-1600), AreaToZeroAxis1Ragged and
AreaToLineAxis1Ragged (area.py:916-1073, 1939-2083) against goldens of the real reference (tests/golden/ragged.npz) and,
on random ragged frames, against the oracle's row loops.  Bresenham / scan fill: bit-exact (float sums 1e-12); antialiased:
1e-6 with identical NaN masks; row ids exact."""
import numpy as np
import pytest

from helpers import load

pytestmark = pytest.mark.gpu


def _cmp(got, want, key, atol=0.0):
    assert got.dtype == want.dtype and got.shape == want.shape, (key, got.dtype, want.dtype, got.shape, want.shape)
    if "sum" in key or "mean" in key:
        assert np.array_equal(np.isnan(got), np.isnan(want)), key
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=atol, equal_nan=True, err_msg=key)
    else:
        assert np.array_equal(got, want, equal_nan=(got.dtype.kind == "f")), key


def _cmp_aa(got, want, key, rtol=1e-6):
    assert got.dtype == want.dtype and got.shape == want.shape, (key, got.dtype, want.dtype)
    assert np.array_equal(np.isnan(got), np.isnan(want)), key
    np.testing.assert_allclose(got, want, rtol=rtol, atol=1e-7, equal_nan=True, err_msg=key)


def _frame(g, tag, reference_array=False):
    import pandas as pd
    from datashader_b200.datatypes import RaggedArray
    d = {k: RaggedArray({"start_indices": g[f"{tag}_{k}_starts"], "flat_array": g[f"{tag}_{k}_flat"]}) for k in ("x", "y", "s")}
    d["val"] = g["val"]
    d["cat"] = pd.Categorical.from_codes(g["cat"], categories=["a", "b", "c"])
    return pd.DataFrame(d)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_ragged_lines_golden(tag):
    import datashader_b200 as ds
    g = load("ragged.npz")
    df = _frame(g, tag)
    cvs = ds.Canvas(plot_width=48, plot_height=36, x_range=(0, 1), y_range=(-0.2, 1.1))
    aggs = {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val"), "min": ds.min("val"),
            "mean": ds.mean("val"), "first": ds.first("val"), "last": ds.last("val"),
            "where_max_row": ds.where(ds.max("val")), "by_count": ds.by("cat", ds.count())}
    for name, agg in aggs.items():
        r = cvs.line(df, "x", "y", agg=agg, axis=1)
        _cmp(r.data, g[f"{tag}_line_lw0_{name}"], f"{tag} lw0 {name}")
    assert tuple(r.dims) == ("y", "x", "cat")            # _PointLike labels: the column names (points.py:134-140)
    r = ds.Canvas(plot_width=31, plot_height=23).line(df, "x", "y", agg=ds.count(), axis=1)
    _cmp(r.data, g[f"{tag}_line_auto_count"], "auto count")
    np.testing.assert_array_equal(np.array(list(r.attrs["x_range"]) + list(r.attrs["y_range"])), g[f"{tag}_line_auto_ranges"])
    aa = {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val"), "mean": ds.mean("val"),
          "min": ds.min("val"), "first": ds.first("val"), "last": ds.last("val"),
          "count_nsi": ds.count(self_intersect=False), "sum_nsi": ds.sum("val", self_intersect=False),
          "where_max_row": ds.where(ds.max("val")), "by_max": ds.by("cat", ds.max("val"))}
    for name, agg in aa.items():
        for lw in (1, 2.5):
            got, want = cvs.line(df, "x", "y", agg=agg, axis=1, line_width=lw).data, g[f"{tag}_line_lw{lw}_{name}"]
            if name == "where_max_row":
                _cmp(got, want, f"{tag} lw{lw} {name}")
            else:
                _cmp_aa(got, want, f"{tag} lw{lw} {name}", rtol=2e-6 if name in ("sum", "count", "sum_nsi", "count_nsi") else 1e-6)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_ragged_areas_golden(tag):
    import datashader_b200 as ds
    g = load("ragged.npz")
    df = _frame(g, tag)
    cvs = ds.Canvas(plot_width=48, plot_height=36, x_range=(0, 1), y_range=(-0.2, 1.1))
    aggs = {"any": ds.any(), "count": ds.count(), "sum": ds.sum("val"), "max": ds.max("val"), "first": ds.first("val"),
            "by_count": ds.by("cat", ds.count())}
    for name, agg in aggs.items():
        _cmp(cvs.area(df, "x", "y", agg=agg, axis=1).data, g[f"{tag}_area_zero_{name}"], f"{tag} zero {name}")
        _cmp(cvs.area(df, "x", "y", agg=agg, axis=1, y_stack="s").data, g[f"{tag}_area_line_{name}"], f"{tag} line {name}")
    auto = ds.Canvas(plot_width=31, plot_height=23)
    for kind, kw in (("zero", {}), ("line", {"y_stack": "s"})):
        r = auto.area(df, "x", "y", agg=ds.count(), axis=1, **kw)
        np.testing.assert_array_equal(np.array(list(r.attrs["x_range"]) + list(r.attrs["y_range"])), g[f"{tag}_area_{kind}_auto_ranges"])
        _cmp(r.data, g[f"{tag}_area_{kind}_auto_count"], f"{tag} area {kind} auto")


def _random_ragged(seed, nrows, maxlen, dtype):
    rng = np.random.default_rng(seed)

    def col(lens, smooth):
        rows = []
        for n in lens:
            a = np.cumsum(rng.normal(0, 0.08, n)) + rng.random() if smooth else np.sort(rng.random(n) * 1.3 - 0.15)
            if n and rng.random() < 0.2:
                a[rng.integers(0, n)] = np.nan
            rows.append(a.astype(dtype))
        lens = np.asarray(lens, dtype=np.int64)
        flat = np.concatenate(rows) if len(rows) and lens.sum() else np.empty(0, dtype)
        return flat, np.cumsum(lens) - lens
    lx = rng.integers(0, maxlen, nrows)
    ly = np.where(rng.random(nrows) < 0.15, rng.integers(0, maxlen, nrows), lx)      # some rows: x and y lengths differ
    ls = np.where(rng.random(nrows) < 0.15, rng.integers(0, maxlen, nrows), ly)
    return col(lx, False), col(ly, True), col(ls, True), (rng.random(nrows) * 4 - 1).astype(np.float32)


@pytest.mark.parametrize("seed,nrows,maxlen,dtype", [(1, 300, 40, np.float32), (2, 2000, 12, np.float64), (3, 50, 700, np.float32),
                                                      (4, 1, 5, np.float32), (5, 5000, 3, np.float32)])
def test_ragged_random_vs_oracle(seed, nrows, maxlen, dtype):
    """Random ragged frames (many short rows, few long rows, a single row) against the oracle's row loops; device-resident
    RaggedColumns through a dict source as well."""
    import torch
    import datashader_b200 as ds
    from datashader_b200.datatypes import RaggedArray
    from oracle import oracle as ora
    import pandas as pd
    (xf, xs), (yf, ys), (sf, ss), val = _random_ragged(seed, nrows, maxlen, dtype)
    df = pd.DataFrame({"x": RaggedArray({"start_indices": xs, "flat_array": xf}), "y": RaggedArray({"start_indices": ys, "flat_array": yf}),
                       "s": RaggedArray({"start_indices": ss, "flat_array": sf}), "val": val})
    W, H = 97, 61
    cvs = ds.Canvas(plot_width=W, plot_height=H, x_range=(0, 1), y_range=(-0.3, 1.4))
    view = ora.make_view(W, H, (0, 1), (-0.3, 1.4))
    for name, agg in (("count", ds.count()), ("sum", ds.sum("val")), ("max", ds.max("val")), ("min", ds.min("val"))):
        vals = None if name == "count" else val
        _cmp(cvs.line(df, "x", "y", agg=agg, axis=1).data, ora.lines_ragged(xf, xs, yf, ys, view, name, vals, 0), f"lw0 {name}", atol=1e-12)
        _cmp(cvs.area(df, "x", "y", agg=agg, axis=1).data, ora.areas_ragged(xf, xs, yf, ys, view, None, None, name, vals), f"zero {name}",
             atol=1e-10)
        _cmp(cvs.area(df, "x", "y", agg=agg, axis=1, y_stack="s").data, ora.areas_ragged(xf, xs, yf, ys, view, sf, ss, name, vals),
             f"line {name}", atol=1e-10)
    for name, agg in (("any", ds.any()), ("count", ds.count()), ("max", ds.max("val")), ("mean", ds.mean("val"))):
        vals = None if name in ("any", "count") else val
        _cmp_aa(cvs.line(df, "x", "y", agg=agg, axis=1, line_width=2).data, ora.lines_ragged(xf, xs, yf, ys, view, name, vals, 2), f"aa {name}",
                rtol=1e-5 if name == "count" else 1e-6)
    for name, agg in (("min", ds.min("val")), ("first", ds.first("val")), ("last", ds.last("val")),
                      ("sum", ds.sum("val", self_intersect=False))):
        _cmp_aa(cvs.line(df, "x", "y", agg=agg, axis=1, line_width=2).data, ora.lines_ragged_aa2(xf, xs, yf, ys, view, name, val, 2),
                f"aa2 {name}", rtol=2e-6)
    # the same columns already on the device (dict source -> DeviceFrame of RaggedColumns)
    dev = {"x": ds.RaggedColumn(torch.from_numpy(xf).cuda(), torch.from_numpy(xs).cuda()),
           "y": ds.RaggedColumn(torch.from_numpy(yf).cuda(), torch.from_numpy(ys).cuda()), "val": torch.from_numpy(val).cuda()}
    _cmp(cvs.line(ds.DeviceFrame(dev), "x", "y", agg=ds.max("val"), axis=1).data, cvs.line(df, "x", "y", agg=ds.max("val"), axis=1).data, "device")


def test_ragged_equal_rows_match_dense_layout():
    """Rows of one length: the ragged kernels must reproduce the dense LinesAxis1 / area results bit for bit."""
    import pandas as pd
    import datashader_b200 as ds
    from datashader_b200.datatypes import RaggedArray
    rng = np.random.default_rng(77)
    nl, nv = 400, 17
    xs = np.sort(rng.random((nl, nv)).astype(np.float32) * 1.2 - 0.1, axis=1)
    ys = (np.cumsum(rng.normal(0, 0.1, (nl, nv)), axis=1) + 0.5).astype(np.float32)
    ys[5, 3] = np.nan
    val = rng.random(nl).astype(np.float32)
    d = {f"x{j}": xs[:, j] for j in range(nv)}
    d.update({f"y{j}": ys[:, j] for j in range(nv)})
    d["val"] = val
    dense = pd.DataFrame(d)
    rag = pd.DataFrame({"x": RaggedArray(list(xs)), "y": RaggedArray(list(ys)), "val": val})
    xc, yc = [f"x{j}" for j in range(nv)], [f"y{j}" for j in range(nv)]
    cvs = ds.Canvas(plot_width=120, plot_height=80, x_range=(0, 1), y_range=(-0.5, 1.5))
    for lw in (0, 1.5):
        for agg in (ds.count(), ds.max("val"), ds.first("val"), ds.where(ds.min("val"))):
            a = cvs.line(rag, "x", "y", agg=agg, axis=1, line_width=lw).data
            b = cvs.line(dense, xc, yc, agg=agg, axis=1, line_width=lw).data
            if lw > 0 and isinstance(agg, ds.count):     # float32 atomic adds: the order of the updates differs between launches
                _cmp_aa(a, b, "aa count", rtol=1e-5)
                continue
            assert a.dtype == b.dtype and np.array_equal(a, b, equal_nan=True), (lw, type(agg).__name__)
    for agg in (ds.count(), ds.max("val")):
        a, b = cvs.area(rag, "x", "y", agg=agg, axis=1).data, cvs.area(dense, xc, yc, agg=agg, axis=1).data
        assert np.array_equal(a, b, equal_nan=True)


def test_ragged_validation():
    import pandas as pd
    import datashader_b200 as ds
    from datashader_b200.datatypes import RaggedArray
    df = pd.DataFrame({"x": RaggedArray([[0.0, 1.0], [0.5]]), "y": [1.0, 2.0]})
    cvs = ds.Canvas(plot_width=8, plot_height=8, x_range=(0, 1), y_range=(0, 1))
    with pytest.raises(ValueError, match="y must be a RaggedArray"):
        cvs.line(df, "x", "y", axis=1)
    with pytest.raises(ValueError, match="x must be a RaggedArray"):
        cvs.area(df, "y", "x", axis=1)
