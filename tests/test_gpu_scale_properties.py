"""Size-independent properties at sizes the CPU oracle cannot reach in seconds (2e8 points, 900x525):
conservation (sum of counts == rows inside the ranges), K2 (shared-memory privatised) == K1 (global REDs) bit
for bit, partition linearity (count(A) + count(B) == count(A u B)), mean within 1e-12 between the two kernels,
max idempotent under duplication of the input."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    import torch
    import datashader_b200 as ds
    n = 200_000_000
    g = torch.Generator(device="cuda")
    g.manual_seed(99)
    x = torch.rand(n, generator=g, device="cuda") * 1.2 - 0.1
    y = torch.rand(n, generator=g, device="cuda") * 1.2 - 0.1
    v = torch.randn(n, generator=g, device="cuda")
    v[::1009] = float("nan")
    return ds, torch, n, x, y, v


def _run(ds, frame, agg, priv):
    old = ds.config.priv_min_rows
    ds.config.priv_min_rows = 0 if priv else 1 << 62
    try:
        return ds.Canvas(900, 525, x_range=(0.0, 1.0), y_range=(0.0, 1.0)).points(frame, "x", "y", agg).data
    finally:
        ds.config.priv_min_rows = old


def test_count_conservation_and_kernel_equivalence(big):
    ds, torch, n, x, y, v = big
    frame = ds.DeviceFrame({"x": x, "y": y, "value": v})
    inside = int(((x >= 0) & (x <= 1) & (y >= 0) & (y <= 1)).sum().item())
    k2 = _run(ds, frame, ds.count(), True)
    k1 = _run(ds, frame, ds.count(), False)
    assert int(k2.sum(dtype=np.int64)) == inside
    assert np.array_equal(k1, k2)
    # count of a column skips its NaNs
    k2v = _run(ds, frame, ds.count("value"), True)
    inside_v = int(((x >= 0) & (x <= 1) & (y >= 0) & (y <= 1) & ~torch.isnan(v)).sum().item())
    assert int(k2v.sum(dtype=np.int64)) == inside_v
    assert np.array_equal(k2v, _run(ds, frame, ds.count("value"), False))


def test_partition_linearity(big):
    ds, torch, n, x, y, v = big
    h = n // 3
    a = ds.DeviceFrame({"x": x[:h], "y": y[:h]})
    b = ds.DeviceFrame({"x": x[h:], "y": y[h:]}, row_offset=h)
    whole = ds.DeviceFrame({"x": x, "y": y})
    ca, cb, cw = (_run(ds, f, ds.count(), True) for f in (a, b, whole))
    assert np.array_equal(ca + cb, cw)


def test_mean_k2_vs_k1_and_max_idempotence(big):
    ds, torch, n, x, y, v = big
    frame = ds.DeviceFrame({"x": x, "y": y, "value": v})
    m2 = _run(ds, frame, ds.mean("value"), True)
    m1 = _run(ds, frame, ds.mean("value"), False)
    assert np.array_equal(np.isnan(m1), np.isnan(m2))
    np.testing.assert_allclose(m2, m1, rtol=1e-12, atol=1e-15, equal_nan=True)   # f64 sums, order differs only
    half = n // 2
    mx = _run(ds, frame, ds.max("value"), False)
    twice = ds.DeviceFrame({"x": torch.cat([x[:half], x]), "y": torch.cat([y[:half], y]), "value": torch.cat([v[:half], v])})
    assert np.array_equal(_run(ds, twice, ds.max("value"), False), mx, equal_nan=True)


def _configure(ds, key, value):
    ds._lib.check(ds._lib.lib().dsb_configure(key.encode(), int(value)), "dsb_configure")


def test_mono_kernel_equals_generic_at_scale(big):
    """k_points_mono (filtered, vectorised) == k_points_generic bit for bit on 2e8 points for every monotone reduction;
    first <= last row-wise; where(max) returns a row whose value is the pixel's max."""
    ds, torch, n, x, y, v = big
    frame = ds.DeviceFrame({"x": x, "y": y, "value": v})
    aggs = {"max": ds.max("value"), "min": ds.min("value"), "first": ds.first("value"), "last": ds.last("value"),
            "where_max_row": ds.where(ds.max("value")), "where_min_row": ds.where(ds.min("value"))}
    out = {}
    for name, agg in aggs.items():
        _configure(ds, "mono", 1)
        a = _run(ds, frame, agg, False)
        _configure(ds, "mono", 0)
        try:
            b = _run(ds, frame, agg, False)
        finally:
            _configure(ds, "mono", 1)
        assert a.dtype == b.dtype and np.array_equal(a, b, equal_nan=a.dtype.kind == "f"), name
        out[name] = a
    assert np.all((out["min"] <= out["max"]) | np.isnan(out["max"]))
    rows = torch.from_numpy(out["where_max_row"]).cuda()
    hit = rows >= 0
    picked = v[rows[hit]].cpu().numpy()
    assert np.array_equal(picked, out["max"][hit.cpu().numpy()])


def test_count16_equals_u32_count_at_scale(big):
    """by('cat', count()) through the 16-bit packed path == the u32 path, and conserves the number of rows in range."""
    ds, torch, n, x, y, v = big
    m = 60_000_000
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    cat = torch.randint(0, 16, (m,), generator=g, device="cuda", dtype=torch.int8)
    frame = ds.DeviceFrame({"x": x[:m], "y": y[:m], "cat": cat}, categories={"cat": [f"c{i}" for i in range(16)]})
    cvs = ds.Canvas(1920, 1080, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    a = cvs.points(frame, "x", "y", ds.by("cat", ds.count())).data
    old = ds.config.count16
    ds.config.count16 = False
    try:
        b = cvs.points(frame, "x", "y", ds.by("cat", ds.count())).data
    finally:
        ds.config.count16 = old
    assert np.array_equal(a, b)
    inside = int(((x[:m] >= 0) & (x[:m] <= 1) & (y[:m] >= 0) & (y[:m] <= 1)).sum().item())
    assert int(a.sum(dtype=np.int64)) == inside


def test_head_then_filtered_rest_equals_single_pass_at_scale(big):
    """The head-then-filter forms at 2e8 points (423 rows per canvas cell, NaN values, points outside the canvas): max / min /
    where(max | min) through dsb_points over a head + dsb_points_minmax_rest / dsb_points_argminmax_rest over the rest, and
    first / last through the routed head + k_rows_rest, each bit-equal to the same reduction with the split switched off
    (k_points_mono over every row) - and the result is a fixed point: max over the row-reversed frame is the same canvas."""
    ds, torch, n, x, y, v = big
    from datashader_b200 import _lib
    L = _lib.lib()
    frame = ds.DeviceFrame({"x": x, "y": y, "value": v})
    aggs = {"max": ds.max("value"), "min": ds.min("value"), "where_max_row": ds.where(ds.max("value")),
            "where_min_row": ds.where(ds.min("value")), "first": ds.first("value"), "last": ds.last("value")}
    old = (ds.config.minmax_split_rows_per_cell, ds.config.minmax_head_rows_per_cell, ds.config.routed)
    got = {}
    try:
        for name, agg in aggs.items():
            ds.config.minmax_split_rows_per_cell, ds.config.minmax_head_rows_per_cell, ds.config.routed = 64, 32, True
            a = _run(ds, frame, agg, False)
            kern = L.dsb_last_kernel()
            if name in ("max", "min"):
                assert b"k_points_minmax_rest<" in kern, (name, kern)
            elif name.startswith("where"):
                assert b"k_points_argminmax_rest<" in kern, (name, kern)
            else:
                assert b"k_rows_rest<" in kern, (name, kern)
            ds.config.minmax_split_rows_per_cell, ds.config.routed = 0, False
            b = _run(ds, frame, agg, False)
            assert b"rest<" not in L.dsb_last_kernel(), (name, L.dsb_last_kernel())
            assert a.dtype == b.dtype and np.array_equal(a, b, equal_nan=a.dtype.kind == "f"), name
            got[name] = a
        ds.config.minmax_split_rows_per_cell, ds.config.minmax_head_rows_per_cell, ds.config.routed = 64, 32, True
        flipped = ds.DeviceFrame({"x": x.flip(0), "y": y.flip(0), "value": v.flip(0)})
        assert np.array_equal(_run(ds, flipped, ds.max("value"), False), got["max"], equal_nan=True)
        assert np.array_equal(_run(ds, flipped, ds.first("value"), False), got["last"], equal_nan=True)
    finally:
        ds.config.minmax_split_rows_per_cell, ds.config.minmax_head_rows_per_cell, ds.config.routed = old
