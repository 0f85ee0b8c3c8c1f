"""Parity of the CUDA path (through Canvas.points -> ctypes -> libdsb200.so) with the reference's golden
vectors and with the C oracle on seeded inputs.  Bit-exact for count / any / min / max / first / last /
where / by; float sums within RTOL_SUM (f64 accumulation on both sides, only the order differs)."""
import numpy as np
import pytest

from helpers import (CANVASES, NCAT, SPECS, assert_agg_equal, columns_from_golden, load, make_agg, pandas_frame)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ds():
    import torch
    assert torch.cuda.is_available()
    import datashader_b200 as ds
    return ds


@pytest.mark.parametrize("tag", ["f32", "f64"])
@pytest.mark.parametrize("cname", list(CANVASES))
def test_points_golden(ds, tag, cname):
    g = load("points.npz")
    cols = columns_from_golden(g, f"in_{tag}_")
    df = pandas_frame(cols)
    cvs = ds.Canvas(**CANVASES[cname])
    n = 0
    for rname, spec in SPECS.items():
        key = f"pts_{tag}_{cname}_{rname}"
        if key not in g.files:
            continue
        agg = cvs.points(df, "x", "y", make_agg(spec))
        assert_agg_equal(agg.data, g[key], key)
        n += 1
    assert n >= 5
    agg = cvs.points(df, "x", "y")
    np.testing.assert_array_equal(agg.coords["x"], g[f"pts_{tag}_{cname}_xcoords"])
    np.testing.assert_array_equal(agg.coords["y"], g[f"pts_{tag}_{cname}_ycoords"])
    np.testing.assert_array_equal(np.asarray(agg.attrs["x_range"], dtype="f8"), g[f"pts_{tag}_{cname}_xrange"])
    np.testing.assert_array_equal(np.asarray(agg.attrs["y_range"], dtype="f8"), g[f"pts_{tag}_{cname}_yrange"])
    assert tuple(agg.dims) == ("y", "x")


@pytest.mark.parametrize("source", ["pandas", "device_mono", "host_chunks"])
def test_points_negzero_golden(ds, source):
    """A pixel whose max / min is a zero keeps the sign of the zero that arrived first (reductions.py:1178-1183,
    1222-1227): the keys fold -0.0 onto +0.0, the pass raises DSB_NOTE_NEGZERO and the host redoes the reduction through
    the row-exact accumulators.  Bit patterns compared (helpers.assert_agg_equal)."""
    import torch
    from datashader_b200 import _lib
    from test_oracle_golden import NEGZERO
    g = load("points_negzero.npz")
    cols = columns_from_golden(g, "in_")
    cvs = ds.Canvas(plot_width=9, plot_height=7, x_range=(0, 1), y_range=(0, 1))
    L = _lib.lib()
    try:
        if source == "pandas":
            src = pandas_frame(cols)
        elif source == "device_mono":         # the single-accumulator kernels (k_points_mono, k_points_mono_f64)
            _lib.check(L.dsb_configure(b"mono_min_rows", 0))
            src = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items() if k not in ("cat__ncat",)},
                                 categories={"cat": [f"c{i}" for i in range(NCAT)]})
        else:                                  # host columns streamed in chunks: the gather runs against the host column
            src = ds.HostFrame({k: torch.from_numpy(v) for k, v in cols.items() if k not in ("cat__ncat", "cat")})
            src.CHUNK_ROWS = 1024
        for name, spec in NEGZERO.items():
            if source == "host_chunks" and spec[0] == "by":
                continue
            assert_agg_equal(cvs.points(src, "x", "y", make_agg(spec)).data, g[f"nz_{name}"], f"negzero {source} {name}")
    finally:
        _lib.check(L.dsb_configure(b"mono_min_rows", 1 << 20))


def test_by_dims_and_labels(ds):
    g = load("points.npz")
    cols = columns_from_golden(g, "in_f32_")
    df = pandas_frame(cols)
    agg = ds.Canvas(**CANVASES["c37x23"]).points(df, "x", "y", ds.count_cat("cat"))
    assert tuple(agg.dims) == ("y", "x", "cat")
    assert list(agg.coords["cat"]) == [f"c{i}" for i in range(NCAT)]
    assert agg.data.dtype == np.uint32 and agg.data.shape == (23, 37, NCAT)


def test_partitioned_golden_single_gpu(ds):
    """first/last/where are row-index based on the GPU, i.e. the reference's partitioned formulation."""
    g = load("partitioned.npz")
    cols = columns_from_golden(g, "in_")
    df = pandas_frame(cols)
    cvs = ds.Canvas(plot_width=37, plot_height=23, x_range=(-0.1, 1.05), y_range=(0.1, 0.9))
    for key in g.files:
        if key.startswith("part3_"):
            rname = key[len("part3_"):]
            assert_agg_equal(cvs.points(df, "x", "y", make_agg(SPECS[rname])).data, g[key], key)


def test_log_axes(ds):
    """Log axes, float32 coordinates: bit-exact with the reference (test_pandas.py:1171-1187).  The device evaluates glibc's
    log10f instruction for instruction (csrc/log10f_glibc.h, proven against the C library for every positive float by
    oracle/log10f_check.c), so boundary points land in the reference's pixel."""
    g = load("points.npz")
    import pandas as pd
    df = pd.DataFrame({k: g[f"in_log_{k}"] for k in ("x", "y", "v32")})
    cvs = ds.Canvas(plot_width=40, plot_height=30, x_range=(1, 1000), y_range=(1, 100), x_axis_type="log", y_axis_type="log")
    assert_agg_equal(cvs.points(df, "x", "y", ds.count()).data, g["pts_log_count"], "log count")
    assert_agg_equal(cvs.points(df, "x", "y", ds.max("v32")).data, g["pts_log_max_v32"], "log max")
    np.testing.assert_allclose(cvs.points(df, "x", "y").coords["x"], g["pts_log_xcoords"], rtol=1e-15)


def test_log_axes_vs_oracle_with_points_on_pixel_edges(ds):
    """2e5 float32 points on log axes, a third of them the float32 values nearest to the pixel edges (and their float32
    neighbours): count, by-count and K2 (forced) are bit-equal to the oracle, which calls the C library's log10f."""
    import torch
    from oracle import oracle as ora
    rng = np.random.default_rng(41)
    n, W, H = 200_000, 300, 200
    xr, yr = (0.5, 2.0e4), (1.0e-3, 7.0)
    x = (10 ** (np.log10(xr[0]) + rng.random(n) * (np.log10(xr[1]) - np.log10(xr[0])))).astype(np.float32)
    y = (10 ** (np.log10(yr[0]) + rng.random(n) * (np.log10(yr[1]) - np.log10(yr[0])))).astype(np.float32)
    ex = (10 ** (np.log10(xr[0]) + np.arange(W + 1) / W * (np.log10(xr[1]) - np.log10(xr[0])))).astype(np.float32)
    ey = (10 ** (np.log10(yr[0]) + np.arange(H + 1) / H * (np.log10(yr[1]) - np.log10(yr[0])))).astype(np.float32)
    k = n // 6
    for arr, e in ((x, ex), (y, ey)):
        pick = e[rng.integers(0, len(e), k)]
        step = rng.integers(-2, 3, k)
        for s_ in (-2, -1, 1, 2):
            m = step == s_
            for _ in range(abs(s_)):
                pick[m] = np.nextafter(pick[m], np.float32(np.inf if s_ > 0 else -np.inf))
        arr[:k] = pick
        rng.shuffle(arr)
    cols = {"x": x, "y": y, "cat": rng.integers(0, NCAT, n).astype(np.int8), "cat__ncat": NCAT}
    view = ora.make_view(W, H, xr, yr, "log", "log")
    cvs = ds.Canvas(W, H, x_range=xr, y_range=yr, x_axis_type="log", y_axis_type="log")
    frame = ds.DeviceFrame({k_: torch.from_numpy(v).cuda() for k_, v in cols.items() if k_ != "cat__ncat"},
                           categories={"cat": [f"c{i}" for i in range(NCAT)]})
    assert_agg_equal(cvs.points(frame, "x", "y", ds.count()).data, ora.points(cols, "x", "y", ("count",), view), "log count vs oracle")
    assert_agg_equal(cvs.points(frame, "x", "y", ds.count_cat("cat")).data, ora.points(cols, "x", "y", ("by", "cat", ("count",)), view),
                     "log by-count vs oracle")
    old = ds.config.priv_min_rows
    ds.config.priv_min_rows = 0
    try:
        assert_agg_equal(cvs.points(frame, "x", "y", ds.count()).data, ora.points(cols, "x", "y", ("count",), view), "log K2 count")
    finally:
        ds.config.priv_min_rows = old


@pytest.mark.parametrize("seed,n,W,H", [(1, 300_000, 900, 525), (2, 200_000, 1920, 1080), (3, 100_000, 17, 3)])
def test_points_vs_oracle_seeded(ds, seed, n, W, H):
    from oracle import oracle as ora
    rng = np.random.default_rng(seed)
    cols = {
        "x": rng.random(n, dtype=np.float32), "y": rng.random(n, dtype=np.float32),
        "v32": rng.standard_normal(n).astype(np.float32), "other": rng.random(n).astype(np.float32),
        "vi": rng.integers(-9, 9, n).astype(np.int32), "v64": rng.standard_normal(n),
        "cat": rng.integers(0, NCAT, n).astype(np.int8), "cat__ncat": NCAT,
    }
    cols["v32"][rng.integers(0, n, n // 1000 + 1)] = np.nan
    # adversarial coordinates: every pixel edge and its float32 neighbours
    edges = (np.arange(W + 1) / W).astype(np.float32)
    k = min(len(edges), n // 3)
    cols["x"][:k] = edges[:k]
    cols["x"][k:2 * k] = np.nextafter(edges[:k], np.float32(2))
    cols["x"][2 * k:3 * k] = np.nextafter(edges[:k], np.float32(-1))
    df = pandas_frame(cols)
    view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    for rname in ("count", "count_v32", "any", "mean_v32", "sum_v32", "max_v32", "min_v32", "max_vi", "max_v64",
                  "first_v32", "last_v32", "where_max_v32_other", "where_min_v32_row", "where_max_vi_other",
                  "where_first_v32_other", "where_last_v32_row", "by_count", "by_mean_v32", "by_max_v32"):
        spec = SPECS[rname]
        got = cvs.points(df, "x", "y", make_agg(spec)).data
        want = ora.points(cols, "x", "y", spec, view, npartitions=2 if "first" in rname or "last" in rname else 1)
        assert_agg_equal(got, want, f"{rname} seed={seed}")


def test_where_max_f64_two_pass(ds):
    from oracle import oracle as ora
    rng = np.random.default_rng(11)
    n = 50_000
    cols = {"x": rng.random(n), "y": rng.random(n), "v64": np.round(rng.standard_normal(n), 1),
            "other": rng.random(n).astype(np.float32)}
    df = pandas_frame(cols)
    view = ora.make_view(31, 29, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(31, 29, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    for spec in (("where", ("max", "v64"), "other"), ("where", ("min", "v64"), None)):
        assert_agg_equal(cvs.points(df, "x", "y", make_agg(spec)).data, ora.points(cols, "x", "y", spec, view), str(spec))


def test_summary_and_device_frame(ds):
    import torch
    from oracle import oracle as ora
    rng = np.random.default_rng(5)
    n = 20_000
    cols = {"x": rng.random(n, dtype=np.float32), "y": rng.random(n, dtype=np.float32),
            "v32": rng.standard_normal(n).astype(np.float32)}
    frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items()})
    cvs = ds.Canvas(50, 40, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    out = cvs.points(frame, "x", "y", ds.summary(n=ds.count(), m=ds.mean("v32"), hi=ds.max("v32")))
    view = ora.make_view(50, 40, (0.0, 1.0), (0.0, 1.0))
    assert_agg_equal(out["n"].data, ora.points(cols, "x", "y", ("count",), view), "summary count")
    assert_agg_equal(out["m"].data, ora.points(cols, "x", "y", ("mean", "v32"), view), "summary mean")
    assert_agg_equal(out["hi"].data, ora.points(cols, "x", "y", ("max", "v32"), view), "summary max")


def test_empty_and_all_nan(ds):
    import pandas as pd
    cvs = ds.Canvas(4, 3, x_range=(0, 1), y_range=(0, 1))
    df = pd.DataFrame({"x": np.array([], dtype="f4"), "y": np.array([], dtype="f4"), "v": np.array([], dtype="f4")})
    assert cvs.points(df, "x", "y").data.sum() == 0
    assert np.isnan(cvs.points(df, "x", "y", ds.max("v")).data).all()
    df = pd.DataFrame({"x": np.full(5, np.nan, "f4"), "y": np.full(5, np.nan, "f4")})
    agg = ds.Canvas(4, 3).points(df, "x", "y")          # auto-range on all-NaN -> (-1, 1) (glyph.py:42-49)
    assert agg.attrs["x_range"] == (-1.0, 1.0) and agg.data.sum() == 0


def test_errors_match_reference(ds):
    import pandas as pd
    df = pd.DataFrame({"x": [0.1, 0.2], "y": [0.1, 0.2], "s": ["a", "b"], "v": [1.0, 2.0]})
    cvs = ds.Canvas(2, 2, x_range=(0, 1), y_range=(0, 1))
    with pytest.raises(ValueError, match="specified column not found"):
        cvs.points(df, "x", "y", ds.mean("nope"))
    with pytest.raises(ValueError, match="input must be categorical"):
        cvs.points(df, "x", "y", ds.count_cat("v"))
    with pytest.raises(ValueError, match="where and its contained reduction cannot use the same column"):
        cvs.points(df, "x", "y", ds.where(ds.max("v"), "v"))
    with pytest.raises(TypeError, match="selector can only be"):
        ds.where(ds.mean("v"), "x")
    with pytest.raises(ValueError, match="Range values must be >0"):
        ds.Canvas(2, 2, x_range=(0, 1), y_range=(0, 1), x_axis_type="log").points(df, "x", "y")
    with pytest.raises(ValueError, match="Invalid size"):
        ds.Canvas(0, 2, x_range=(0, 1), y_range=(0, 1)).points(df, "x", "y")
    with pytest.raises(ValueError, match="coordinates may be specified"):
        cvs.points(df, "x")


def test_host_frame_chunk_streaming(ds, monkeypatch):
    """Host columns streamed in several double-buffered chunks give the same aggregates as one pass
    (global row ids keep first/last/where exact across chunk boundaries)."""
    from oracle import oracle as ora
    monkeypatch.setattr(ds.HostFrame, "CHUNK_ROWS", 7_000)
    rng = np.random.default_rng(21)
    n = 50_000
    cols = {"x": rng.random(n, dtype=np.float32), "y": rng.random(n, dtype=np.float32),
            "v32": np.round(rng.standard_normal(n), 1).astype(np.float32), "other": rng.random(n).astype(np.float32),
            "v64": np.round(rng.standard_normal(n), 1), "cat": rng.integers(0, NCAT, n).astype(np.int8), "cat__ncat": NCAT}
    cols["v32"][rng.integers(0, n, 50)] = np.nan
    df = pandas_frame(cols)
    view = ora.make_view(64, 48, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(64, 48, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    for rname in ("count", "mean_v32", "max_v32", "first_v32", "last_v32", "where_max_v32_other", "where_min_v32_row",
                  "where_first_v32_other", "by_count", "by_max_v32"):
        assert_agg_equal(cvs.points(df, "x", "y", make_agg(SPECS[rname])).data,
                         ora.points(cols, "x", "y", SPECS[rname], view, npartitions=2 if "first" in rname or "last" in rname else 1),
                         f"chunked {rname}")
    spec = ("where", ("max", "v64"), "other")
    assert_agg_equal(cvs.points(df, "x", "y", make_agg(spec)).data, ora.points(cols, "x", "y", spec, view), "chunked 2-pass")
    # auto-ranging over chunks
    got = ds.Canvas(31, 17).points(df, "x", "y")
    v2 = ora.make_view(31, 17, ora.compute_bounds(cols["x"]), ora.compute_bounds(cols["y"]))
    assert_agg_equal(got.data, ora.points(cols, "x", "y", ("count",), v2), "chunked auto-range")


@pytest.fixture
def force_priv(ds):
    old = ds.config.priv_min_rows
    ds.config.priv_min_rows = 0
    yield
    ds.config.priv_min_rows = old


@pytest.mark.parametrize("W,H", [(2, 2), (500, 400), (600, 500), (900, 525), (1024, 768)])
def test_priv_count_kernel_matches_oracle(ds, force_priv, W, H):
    """K2 (shared-memory privatised count): every slot width (8/5/4/3/2 bits), count / count(col) / mean / by."""
    import torch
    from oracle import oracle as ora
    rng = np.random.default_rng(W * 1000 + H)
    n = 400_003                                   # not a multiple of 4: exercises the vector tail
    cols = {"x": rng.random(n, dtype=np.float32), "y": rng.random(n, dtype=np.float32),
            "v32": rng.standard_normal(n).astype(np.float32)}
    cols["v32"][rng.integers(0, n, 400)] = np.nan
    cols["x"][:5] = [0.0, 1.0, 0.5, np.nan, 2.0]
    view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items()})
    for spec in (("count",), ("count", "v32"), ("mean", "v32"), ("sum", "v32"), ("any",), ("any", "v32")):
        assert_agg_equal(cvs.points(frame, "x", "y", make_agg(spec)).data, ora.points(cols, "x", "y", spec, view), f"priv {spec} {W}x{H}")
    # unaligned column start -> scalar-load variant
    frame2 = ds.DeviceFrame({k: torch.from_numpy(v).cuda()[1:] for k, v in cols.items()})
    cols2 = {k: v[1:] for k, v in cols.items()}
    assert_agg_equal(cvs.points(frame2, "x", "y", ds.count()).data, ora.points(cols2, "x", "y", ("count",), view), "priv unaligned")


def test_priv_count_hot_pixel_falls_back_exactly(ds, force_priv):
    """All points in a handful of pixels: guard bits overflow, the carry is detected and the count is redone
    with global REDs - the result must still be exact."""
    import torch
    n = 3_000_000
    x = torch.full((n,), 0.5, dtype=torch.float32, device="cuda")
    y = torch.full((n,), 0.25, dtype=torch.float32, device="cuda")
    x[::7] = 0.75
    frame = ds.DeviceFrame({"x": x, "y": y})
    got = ds.Canvas(900, 525, x_range=(0.0, 1.0), y_range=(0.0, 1.0)).points(frame, "x", "y").data
    assert int(got.sum()) == n
    assert got[131, 450] == n - len(range(0, n, 7)) and got[131, 675] == len(range(0, n, 7))
    hit = ds.Canvas(900, 525, x_range=(0.0, 1.0), y_range=(0.0, 1.0)).points(frame, "x", "y", ds.any()).data
    assert hit.dtype == np.bool_ and int(hit.sum()) == 2 and hit[131, 450] and hit[131, 675]


def test_priv_by_count_small_canvas(ds, force_priv):
    import torch
    from oracle import oracle as ora
    rng = np.random.default_rng(77)
    n = 200_000
    cols = {"x": rng.random(n, dtype=np.float32), "y": rng.random(n, dtype=np.float32),
            "cat": rng.integers(0, NCAT, n).astype(np.int8), "cat__ncat": NCAT}
    frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items() if k != "cat__ncat"},
                           categories={"cat": [f"c{i}" for i in range(NCAT)]})
    view = ora.make_view(300, 200, (0.0, 1.0), (0.0, 1.0))
    got = ds.Canvas(300, 200, x_range=(0.0, 1.0), y_range=(0.0, 1.0)).points(frame, "x", "y", ds.count_cat("cat")).data
    assert_agg_equal(got, ora.points(cols, "x", "y", ("by", "cat", ("count",)), view), "priv by count")


@pytest.mark.parametrize("xr,yr", [((0.0, 1.0), (0.0, 1.0)), ((-0.1, 1.05), (0.1, 0.9)), ((-3.7, 12.9), (5.0, 5.5)),
                                   ((1.0e6, 1.0e6 + 1.0), (0.0, 1.0))])
def test_priv_fast_mapping_adversarial_edges(ds, force_priv, xr, yr):
    """The float32 fast path of K2 must agree with the exact f64 mapping on every pixel edge and its float32
    neighbours (and must switch itself off when the error bound is too large, last case)."""
    import torch
    from oracle import oracle as ora
    W, H = 900, 525
    rng = np.random.default_rng(3)
    ex = xr[0] + (xr[1] - xr[0]) * np.arange(W + 1) / W
    ey = yr[0] + (yr[1] - yr[0]) * np.arange(H + 1) / H
    xs, ys = [], []
    for e, other, out in ((ex, yr, xs), (ey, xr, ys)):
        f = e.astype(np.float32)
        cand = np.concatenate([f, np.nextafter(f, np.float32(np.inf)), np.nextafter(f, np.float32(-np.inf)),
                               np.nextafter(np.nextafter(f, np.float32(np.inf)), np.float32(np.inf))])
        out.append(cand)
    nx, ny = len(xs[0]), len(ys[0])
    x = np.concatenate([xs[0], (xr[0] + (xr[1] - xr[0]) * rng.random(ny)).astype(np.float32),
                        (xr[0] + (xr[1] - xr[0]) * rng.random(200_000)).astype(np.float32)])
    y = np.concatenate([(yr[0] + (yr[1] - yr[0]) * rng.random(nx)).astype(np.float32), ys[0],
                        (yr[0] + (yr[1] - yr[0]) * rng.random(200_000)).astype(np.float32)])
    cols = {"x": x, "y": y}
    view = ora.make_view(W, H, xr, yr)
    frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items()})
    got = ds.Canvas(W, H, x_range=xr, y_range=yr).points(frame, "x", "y").data
    assert_agg_equal(got, ora.points(cols, "x", "y", ("count",), view), f"fast map {xr} {yr}")


def test_l2_banded_passes_match_oracle(ds):
    """Canvases beyond the L2 budget are filled in several passes over bands of canvas rows (with a float32 y
    pre-filter); force tiny bands and compare every kind of accumulator with the oracle."""
    import torch
    from datashader_b200 import _lib
    from oracle import oracle as ora
    L = _lib.lib()
    rng = np.random.default_rng(12)
    n = 150_000
    cols = {"x": rng.random(n, dtype=np.float32) * 1.1 - 0.05, "y": rng.random(n, dtype=np.float32) * 1.1 - 0.05,
            "v32": np.round(rng.standard_normal(n), 1).astype(np.float32), "other": rng.random(n).astype(np.float32),
            "v64": np.round(rng.standard_normal(n), 1), "cat": rng.integers(0, NCAT, n).astype(np.int8), "cat__ncat": NCAT}
    cols["y"][:4] = [0.0, 1.0, np.nan, 0.5]
    df = pandas_frame(cols)
    W, H = 301, 257
    view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    old_priv = ds.config.priv_count
    try:
        _lib.check(L.dsb_configure(b"l2_band_bytes", 64 * 1024))
        _lib.check(L.dsb_configure(b"band_min_rows", 0))
        ds.config.priv_count = False
        # mono_min_rows = 0 sends the single-accumulator plans through the BANDED k_points_mono (the production path of
        # 8192^2 canvases); the default leaves them to the banded generic kernel - both must agree with the oracle
        for mono_rows in (0, 1 << 20):
            _lib.check(L.dsb_configure(b"mono_min_rows", mono_rows))
            for rname in ("count", "count_v32", "mean_v32", "max_v32", "min_v32", "max_v64", "first_v32", "last_v32",
                          "where_max_v32_other", "where_min_v32_other", "where_min_v32_row", "where_first_v32_other",
                          "by_count", "by_max_v32"):
                spec = SPECS[rname]
                got = cvs.points(df, "x", "y", make_agg(spec)).data
                want = ora.points(cols, "x", "y", spec, view, npartitions=2 if "first" in rname or "last" in rname else 1)
                assert_agg_equal(got, want, f"banded {rname} mono_min_rows={mono_rows}")
        spec = ("where", ("max", "v64"), "other")
        assert_agg_equal(cvs.points(df, "x", "y", make_agg(spec)).data, ora.points(cols, "x", "y", spec, view), "banded 2-pass")
    finally:
        ds.config.priv_count = old_priv
        _lib.check(L.dsb_configure(b"l2_band_bytes", 96 << 20))
        _lib.check(L.dsb_configure(b"band_min_rows", 1 << 22))
        _lib.check(L.dsb_configure(b"mono_min_rows", 1 << 20))


def test_arrow_and_dict_sources_match_pandas(ds):
    """The same rows through a pandas.DataFrame, a pyarrow.Table, a dict of host arrays and a dict of device arrays."""
    import pandas as pd
    import torch
    pa = pytest.importorskip("pyarrow")
    rng = np.random.default_rng(8)
    n = 50_000
    df = pd.DataFrame({"x": rng.random(n, dtype=np.float32), "y": rng.random(n, dtype=np.float32), "v": rng.normal(size=n),
                       "cat": pd.Categorical.from_codes(rng.integers(0, 4, n), categories=list("abcd"))})
    df.loc[rng.integers(0, n, 50), "v"] = np.nan
    cvs = ds.Canvas(120, 80, x_range=(0, 1), y_range=(0, 1))
    table = pa.Table.from_pandas(df)
    host = {"x": df.x.to_numpy(), "y": df.y.to_numpy(), "v": df.v.to_numpy()}
    dev = {k: torch.from_numpy(v.copy()).cuda() for k, v in host.items()}
    for agg in (ds.count(), ds.mean("v"), ds.max("v"), ds.where(ds.max("v"))):
        want = cvs.points(df, "x", "y", agg).data
        for src in (table, host, dev):
            got = cvs.points(src, "x", "y", agg).data
            assert got.dtype == want.dtype, (type(src), agg)
            if isinstance(agg, ds.mean):      # f64 atomic adds: the summation order differs from run to run
                np.testing.assert_allclose(got, want, rtol=1e-12, equal_nan=True)
            else:
                assert np.array_equal(got, want, equal_nan=got.dtype.kind == "f"), (type(src), agg)
    want = cvs.points(df, "x", "y", ds.by("cat", ds.count()))
    got = cvs.points(table, "x", "y", ds.by("cat", ds.count()))
    assert np.array_equal(got.data, want.data) and list(got.coords["cat"]) == list(want.coords["cat"])


def test_mono_kernel_matches_oracle(ds):
    """k_points_mono (single monotone accumulator, >= 2^20 float32 rows, L2-resident canvas): max / min / first / last /
    where(max | min) with NaN values, NaN coordinates, points on pixel edges and outside the ranges - bit-exact."""
    import torch
    from oracle import oracle as ora
    rng = np.random.default_rng(4242)
    n = (1 << 20) + 4099
    x = (rng.random(n, dtype=np.float32) * 1.2 - 0.1).astype(np.float32)      # ~17 % out of range
    y = (rng.random(n, dtype=np.float32) * 1.2 - 0.1).astype(np.float32)
    x[:2000] = (rng.integers(0, 301, 2000) / 300).astype(np.float32)            # exactly on pixel edges (W = 300)
    y[2000:4000] = (rng.integers(0, 201, 2000) / 200).astype(np.float32)
    x[5000:5010] = np.nan
    cols = {"x": x, "y": y, "v32": np.round(rng.standard_normal(n), 2).astype(np.float32),     # rounded: many ties
            "other": rng.random(n).astype(np.float32)}
    cols["v32"][rng.integers(0, n, 3000)] = np.nan
    W, H = 300, 200
    view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items()})
    specs = [("max", "v32"), ("min", "v32"), ("first", "v32"), ("last", "v32"), ("where", ("max", "v32"), "other"),
             ("where", ("min", "v32"), None), ("where", ("max", "v32"), None), ("where", ("first", "v32"), "other")]
    lib = ds._lib.lib()
    for spec in specs:
        want = ora.points(cols, "x", "y", spec, view, npartitions=2 if ("first" in str(spec) or "last" in str(spec)) else 1)
        for mono in (1, 0):        # the specialised kernel and the generic one must agree with the oracle
            ds._lib.check(lib.dsb_configure(b"mono", mono), "dsb_configure")
            try:
                assert_agg_equal(cvs.points(frame, "x", "y", make_agg(spec)).data, want, f"mono={mono} {spec}")
            finally:
                ds._lib.check(lib.dsb_configure(b"mono", 1), "dsb_configure")


def test_count16_packed_path_matches_oracle(ds):
    """dsb_points_count16 (16-bit packed counters + checksum) on a by-count canvas: the normal case, and a hot cell with
    more than 65 535 hits that must trip the checksum and fall back to the exact u32 pass."""
    import torch
    from oracle import oracle as ora
    old = (ds.config.count16_min_rows, ds.config.l2_budget_bytes)
    ds.config.count16_min_rows = 0
    W, H, NC = 64, 48, 4
    ds.config.l2_budget_bytes = 3 * W * H * NC          # makes 4*cells > budget >= 2*cells: the packed path is chosen
    try:
        rng = np.random.default_rng(16)
        n = 400_000
        cols = {"x": rng.random(n, dtype=np.float32), "y": rng.random(n, dtype=np.float32),
                "v32": rng.standard_normal(n).astype(np.float32), "cat": rng.integers(0, NC, n).astype(np.int8), "cat__ncat": NC}
        cols["v32"][rng.integers(0, n, 500)] = np.nan
        view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
        cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
        frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items() if k != "cat__ncat"},
                               categories={"cat": [f"c{i}" for i in range(NC)]})
        before = ds._lib.lib().dsb_launch_count()
        got = cvs.points(frame, "x", "y", ds.by("cat", ds.count())).data
        assert_agg_equal(got, ora.points(cols, "x", "y", ("by", "cat", ("count",)), view), "count16 by count")
        got = cvs.points(frame, "x", "y", ds.by("cat", ds.count("v32"))).data
        assert_agg_equal(got, ora.points(cols, "x", "y", ("by", "cat", ("count", "v32")), view), "count16 by count(v32)")
        assert ds._lib.lib().dsb_launch_count() > before
        # warm cell: 1 000 hits on one (pixel, category) -> an 8-bit field wraps -> the gated 16-bit stage redoes the pass
        warm = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in cols.items()}
        warm["x"][:1000], warm["y"][:1000], warm["cat"][:1000] = 0.31, 0.32, 2
        wframe = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in warm.items() if k != "cat__ncat"},
                                categories={"cat": [f"c{i}" for i in range(NC)]})
        got = cvs.points(wframe, "x", "y", ds.by("cat", ds.count())).data
        assert_agg_equal(got, ora.points(warm, "x", "y", ("by", "cat", ("count",)), view), "count16 warm cell (8-bit wrap)")
        assert 1000 <= got.max() < 65536
        # the same without the 8-bit stage
        ds._lib.check(ds._lib.lib().dsb_configure(b"count8", 0), "cfg")
        try:
            got = cvs.points(wframe, "x", "y", ds.by("cat", ds.count())).data
            assert_agg_equal(got, ora.points(warm, "x", "y", ("by", "cat", ("count",)), view), "count16 without the 8-bit stage")
        finally:
            ds._lib.check(ds._lib.lib().dsb_configure(b"count8", 1), "cfg")
        # hot cell: 200 000 hits on one (pixel, category) -> a 16-bit half wraps -> checksum mismatch -> exact redo
        cols["x"][:200_000] = 0.51
        cols["y"][:200_000] = 0.52
        cols["cat"][:200_000] = 1
        frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items() if k != "cat__ncat"},
                               categories={"cat": [f"c{i}" for i in range(NC)]})
        got = cvs.points(frame, "x", "y", ds.by("cat", ds.count())).data
        assert_agg_equal(got, ora.points(cols, "x", "y", ("by", "cat", ("count",)), view), "count16 hot cell")
        assert got.max() >= 200_000
    finally:
        ds.config.count16_min_rows, ds.config.l2_budget_bytes = old


def test_summary_split_into_specialised_groups_matches_oracle(ds, force_priv):
    """summary() with >= 3 accumulators on a small canvas is routed group by group to the specialised kernels
    (pipeline._specialised_groups); every member must still equal the oracle, with and without the split."""
    import torch
    from oracle import oracle as ora
    rng = np.random.default_rng(77)
    n = (1 << 20) + 123
    cols = {"x": rng.random(n, dtype=np.float32), "y": rng.random(n, dtype=np.float32),
            "v32": np.round(rng.standard_normal(n), 2).astype(np.float32)}
    cols["v32"][rng.integers(0, n, 2000)] = np.nan
    W, H = 200, 150
    view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items()})
    specs = {"c": ("count",), "m": ("mean", "v32"), "mx": ("max", "v32"), "mn": ("min", "v32"), "f": ("first", "v32"),
             "s": ("sum", "v32")}
    agg = ds.summary(**{k: make_agg(v) for k, v in specs.items()})
    old = ds.config.split_summary
    try:
        for split in (True, False):
            ds.config.split_summary = split
            got = cvs.points(frame, "x", "y", agg)
            for k, spec in specs.items():
                want = ora.points(cols, "x", "y", spec, view, npartitions=2 if spec[0] == "first" else 1)
                assert_agg_equal(np.asarray(got[k].data), want, f"summary split={split} {k}")
    finally:
        ds.config.split_summary = old


@pytest.mark.parametrize("W,H", [(300, 200), (900, 525), (1024, 768)])
def test_priv_kernel_float64_coordinates(ds, force_priv, W, H):
    """K2 with float64 coordinates (pandas' default dtype): count / count(col) / mean / sum of a float64 column, linear
    and log axes, odd row count; anything else must fall back to the generic kernel and still match."""
    import torch
    from oracle import oracle as ora
    rng = np.random.default_rng(W + H)
    n = 300_001
    cols = {"x": rng.random(n) * 1.2 - 0.1, "y": rng.random(n) * 1.2 - 0.1, "v64": np.round(rng.standard_normal(n), 2),
            "v32": rng.standard_normal(n).astype(np.float32)}
    cols["v64"][rng.integers(0, n, 300)] = np.nan
    cols["x"][:4] = [0.0, 1.0, np.nan, 0.5]
    view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items()})
    for spec in (("count",), ("count", "v64"), ("mean", "v64"), ("sum", "v64"), ("any",), ("mean", "v32"), ("max", "v64")):
        assert_agg_equal(cvs.points(frame, "x", "y", make_agg(spec)).data, ora.points(cols, "x", "y", spec, view),
                         f"f64 priv {spec} {W}x{H}", atol=1e-13)
    lcols = dict(cols, x=np.abs(cols["x"]) + 0.01, y=np.abs(cols["y"]) + 0.01)
    lview = ora.make_view(W, H, (0.01, 1.0), (0.01, 1.0), "log", "log")
    lcvs = ds.Canvas(W, H, x_range=(0.01, 1.0), y_range=(0.01, 1.0), x_axis_type="log", y_axis_type="log")
    lframe = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in lcols.items()})
    got = lcvs.points(lframe, "x", "y", ds.count()).data
    want = ora.points(lcols, "x", "y", ("count",), lview)
    assert int(got.sum()) == int(want.sum())
    assert np.abs(got.astype(np.int64) - want.astype(np.int64)).sum() <= 2 * max(1, n // 20000)   # log10 rounding on pixel edges


def test_large_result_leaves_through_the_staging_ring(ds):
    """Aggregates above 128 MiB (here 4200x4100 f64 = 138 MB) are copied to ordinary host memory through the pinned ring;
    the values must equal the device-resident result."""
    import torch
    rng = np.random.default_rng(1)
    n = 200_000
    frame = ds.DeviceFrame({"x": torch.from_numpy(rng.random(n, dtype=np.float32)).cuda(),
                            "y": torch.from_numpy(rng.random(n, dtype=np.float32)).cuda(),
                            "v": torch.from_numpy(rng.standard_normal(n).astype(np.float32)).cuda()})
    cvs = ds.Canvas(4200, 4100, x_range=(0, 1), y_range=(0, 1))
    host = cvs.points(frame, "x", "y", ds.max("v")).data
    ds.config.device_results = True
    try:
        dev = cvs.points(frame, "x", "y", ds.max("v")).data
    finally:
        ds.config.device_results = False
    assert isinstance(host, np.ndarray) and host.dtype == np.float64 and host.shape == (4100, 4200)
    assert np.array_equal(host, dev.cpu().numpy(), equal_nan=True)
    assert int((~np.isnan(host)).sum()) > 0


def test_summary_mixes_categorical_and_plain_reductions(ds):
    """summary() composes any bases (compiler.py:103-107, reductions.py:2169-2246): by() and plain reductions in one call."""
    from oracle import oracle as ora
    g = load("points.npz")
    cols = columns_from_golden(g, "in_f32_")
    df = pandas_frame(cols)
    ckw = CANVASES["c37x23"]
    cvs = ds.Canvas(**ckw)
    agg = cvs.points(df, "x", "y", ds.summary(cats=ds.by("cat", ds.count()), n=ds.count(), m=ds.mean("v32"),
                                              catmax=ds.by("cat", ds.max("v32"))))
    assert_agg_equal(agg["cats"].data, g["pts_f32_c37x23_by_count"], "summary by_count")
    assert_agg_equal(agg["n"].data, g["pts_f32_c37x23_count"], "summary count")
    assert_agg_equal(agg["m"].data, g["pts_f32_c37x23_mean_v32"], "summary mean_v32")
    assert_agg_equal(agg["catmax"].data, g["pts_f32_c37x23_by_max_v32"], "summary by_max_v32")
    assert tuple(agg["cats"].dims) == ("y", "x", "cat") and tuple(agg["n"].dims) == ("y", "x")


@pytest.mark.gpu
@pytest.mark.parametrize("mono_rows", [0, 1 << 20])
def test_where_rows_of_a_frame_that_crosses_2_to_the_32(ds, mono_rows):
    """where(max | min) keeps the FIRST row of the extreme (reductions.py:2009-2016).  The packed {key32, row} accumulator breaks
    ties on the low 32 bits of the global row id, so a frame whose rows cross a multiple of 2^32 (a shard of a bigger frame) must not
    use it: with values full of ties, the rows found at row_offset = 2^32 - 7000 are the rows found at offset 0, shifted - and both
    are the oracle's."""
    import torch
    from datashader_b200 import _lib
    from oracle import oracle as ora
    rng = np.random.default_rng(77)
    n = 20_000
    cols = {"x": rng.random(n, dtype=np.float32), "y": rng.random(n, dtype=np.float32),
            "v32": rng.integers(-3, 4, n).astype(np.float32), "other": rng.random(n).astype(np.float32)}
    dev = {k: torch.from_numpy(v).cuda() for k, v in cols.items()}
    view = ora.make_view(16, 12, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(16, 12, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    base = (1 << 32) - 7000
    L = _lib.lib()
    _lib.check(L.dsb_configure(b"mono_min_rows", mono_rows))
    try:
        for which in ("max", "min"):
            want = ora.points(cols, "x", "y", ("where", (which, "v32"), None), view)
            r0 = cvs.points(ds.DeviceFrame(dev), "x", "y", ds.where(getattr(ds, which)("v32"))).data
            r1 = cvs.points(ds.DeviceFrame(dev, row_offset=base), "x", "y", ds.where(getattr(ds, which)("v32"))).data
            assert np.array_equal(r0, want), which
            assert np.array_equal(np.where(r1 >= 0, r1 - base, r1), want), which
            o1 = cvs.points(ds.DeviceFrame(dev, row_offset=base), "x", "y", ds.where(getattr(ds, which)("v32"), "other")).data
            assert_agg_equal(o1, ora.points(cols, "x", "y", ("where", (which, "v32"), "other"), view), f"{which} other")
        # per category plane as well: by(cat, where(...)) at the offset == at offset 0
        cat = torch.from_numpy(rng.integers(0, 3, n).astype(np.int8)).cuda()
        cats = {"cat": ["a", "b", "c"]}
        agg = ds.by("cat", ds.where(ds.max("v32")))
        b0 = cvs.points(ds.DeviceFrame({**dev, "cat": cat}, categories=cats), "x", "y", agg).data
        b1 = cvs.points(ds.DeviceFrame({**dev, "cat": cat}, categories=cats, row_offset=base), "x", "y", agg).data
        assert b0.shape == (12, 16, 3) and (b0 >= 0).any() and np.array_equal(np.where(b1 >= 0, b1 - base, b1), b0)
    finally:
        _lib.check(L.dsb_configure(b"mono_min_rows", 1 << 20))


@pytest.mark.parametrize("chunk", [7_001, 8_192])
def test_device_frame_beyond_the_rows_of_one_call(ds, monkeypatch, force_priv, chunk):
    """A resident frame with more rows than one kernel call takes (2^32 in production, a few thousand here) is walked in
    slices of the same device columns: every reduction equals the single pass (global row ids keep first / last / where exact
    across the slices), auto-ranging and a batched tile level included."""
    import torch
    from oracle import oracle as ora
    rng = np.random.default_rng(23)
    n = 50_000
    cols = {"x": rng.random(n, dtype=np.float32), "y": rng.random(n, dtype=np.float32),
            "v32": np.round(rng.standard_normal(n), 1).astype(np.float32), "other": rng.random(n).astype(np.float32),
            "v64": np.round(rng.standard_normal(n), 1), "cat": rng.integers(0, NCAT, n).astype(np.int8), "cat__ncat": NCAT}
    cols["v32"][rng.integers(0, n, 50)] = np.nan
    frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items() if k != "cat__ncat"},
                           categories={"cat": [f"c{i}" for i in range(NCAT)]})
    view = ora.make_view(64, 48, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(64, 48, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    views = [((0.0, 0.5), (0.0, 0.5)), ((0.5, 1.0), (0.0, 0.5)), ((0.0, 0.5), (0.5, 1.0)), ((0.5, 1.0), (0.5, 1.0))]
    tiles_whole = [t.data for t in cvs.points_batch(frame, "x", "y", ds.count(), views, grid=(2, 2))]
    monkeypatch.setattr(ds.DeviceFrame, "CHUNK_ROWS", chunk)
    assert frame.n_chunks() == -(-n // chunk) and ds.DeviceFrame({"x": frame["x"][:2 * chunk - 2]}).n_chunks() == 1
    for rname in ("count", "mean_v32", "max_v32", "first_v32", "last_v32", "where_max_v32_other", "where_min_v32_row",
                  "where_first_v32_other", "by_count", "by_max_v32"):
        assert_agg_equal(cvs.points(frame, "x", "y", make_agg(SPECS[rname])).data,
                         ora.points(cols, "x", "y", SPECS[rname], view, npartitions=2 if "first" in rname or "last" in rname else 1),
                         f"sliced {rname}")
    spec = ("where", ("max", "v64"), "other")
    assert_agg_equal(cvs.points(frame, "x", "y", make_agg(spec)).data, ora.points(cols, "x", "y", spec, view), "sliced 2-pass")
    got = ds.Canvas(31, 17).points(frame, "x", "y")
    v2 = ora.make_view(31, 17, ora.compute_bounds(cols["x"]), ora.compute_bounds(cols["y"]))
    assert_agg_equal(got.data, ora.points(cols, "x", "y", ("count",), v2), "sliced auto-range")
    for a, b in zip(tiles_whole, cvs.points_batch(frame, "x", "y", ds.count(), views, grid=(2, 2))):
        assert np.array_equal(a, b.data)
