"""dsb_points_routed ("bin, then accumulate in shared memory", csrc/routed.cu) against the C oracle and against dsb_points:
forced on at small n, several buckets, clustered data, records overflowing their bucket's region (direct-atomic fallback),
-0.0 values, and - at BASELINE config 5's production geometry (8192 x 8192, 1e8 points) - against the L2-banded kernels."""
import ctypes as C

import numpy as np
import pytest

from helpers import SPECS, assert_agg_equal, make_agg

pytestmark = pytest.mark.gpu

NAMES = ["max_v32", "min_v32", "first_v32", "last_v32", "count", "count_v32", "where_first_v32_other", "where_last_v32_row"]


@pytest.fixture(params=[1, 0], ids=["tma", "ldg"])
def routed(request):
    import datashader_b200 as ds
    from datashader_b200 import _lib
    L = _lib.lib()
    old = (ds.config.routed_min_rows, ds.config.l2_budget_bytes, ds.config.priv_count, ds.config.count16)
    ds.config.routed_min_rows, ds.config.l2_budget_bytes, ds.config.priv_count, ds.config.count16 = 0, 1, False, False
    _lib.check(L.dsb_routed_configure(0))
    _lib.check(L.dsb_configure(b"routed_tma", request.param))      # pass 2 through the TMA ring / through plain loads
    yield ds
    _lib.check(L.dsb_configure(b"routed_tma", 1))
    ds.config.routed_min_rows, ds.config.l2_budget_bytes, ds.config.priv_count, ds.config.count16 = old
    _lib.check(L.dsb_routed_configure(1 << 24))


def _cols(rng, n, clustered=False):
    if clustered:
        c = rng.integers(0, 3, n)
        cx, cy = np.array([0.2, 0.7, 0.71])[c], np.array([0.3, 0.3, 0.8])[c]
        x = (cx + rng.normal(0, 0.01, n)).astype(np.float32)
        y = (cy + rng.normal(0, 0.01, n)).astype(np.float32)
    else:
        x = (rng.random(n) * 1.2 - 0.1).astype(np.float32)
        y = (rng.random(n) * 1.2 - 0.1).astype(np.float32)
    cols = {"x": x, "y": y, "v32": (np.round(rng.standard_normal(n), 1) + 0.0).astype(np.float32), "other": rng.random(n).astype(np.float32)}
    assert not np.signbit(cols["v32"][cols["v32"] == 0]).any()     # no -0.0: the routed pass is the last aggregation kernel
    cols["v32"][rng.integers(0, n, n // 40)] = np.nan
    cols["x"][:3] = [0.0, 1.0, np.nan]
    cols["y"][:3] = [1.0, 0.0, 0.5]
    return cols


@pytest.mark.parametrize("shape,n,clustered", [((301, 257), 200_003, False), ((640, 480), 300_001, False), ((640, 480), 250_000, True)])
def test_routed_matches_oracle(routed, shape, n, clustered):
    import torch
    from datashader_b200 import _lib
    from oracle import oracle as ora
    ds = routed
    W, H = shape
    cols = _cols(np.random.default_rng(n), n, clustered)
    frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items()})
    view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    for name in NAMES:
        spec = SPECS[name]
        got = cvs.points(frame, "x", "y", make_agg(spec)).data
        assert b"k_route" in _lib.lib().dsb_last_kernel(), (name, _lib.lib().dsb_last_kernel())
        want = ora.points(cols, "x", "y", spec, view, npartitions=2 if ("first" in name or "last" in name) else 1)
        assert_agg_equal(got, want, f"routed {name} {shape} clustered={clustered}")


def test_routed_overflow_falls_back_to_canvas_atomics(routed):
    """A scratch buffer far smaller than the records: almost every record overflows its bucket's region and is applied to
    the canvas directly; the result must equal dsb_points' bit for bit."""
    import torch
    from datashader_b200 import _lib
    ds = routed
    L = _lib.lib()
    rng = np.random.default_rng(5)
    n, W, H = 400_000, 640, 480
    cols = _cols(rng, n)
    x, y, v = (torch.from_numpy(cols[k]).cuda() for k in ("x", "y", "v32"))
    from datashader_b200 import pipeline
    view, _, _ = pipeline.make_view(ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0)), (0.0, 1.0), (0.0, 1.0))
    stream = torch.cuda.current_stream().cuda_stream
    nb = -(-W * H // 45056)
    small = (((nb * 20 + 64) + 255) & ~255) + (((n // 32 + 65536) * 12 + 255) & ~255) + (nb * 4096 + 1024) * 8   # the minimum accepted
    for op, dtype, use_chk in ((_lib.OP_MAX32, torch.int32, False), (_lib.OP_MINROW, torch.int64, True), (_lib.OP_MAXROW, torch.int64, True),
                               (_lib.OP_COUNT, torch.int32, False)):
        outs = []
        for routed_call in (True, False):
            canvas = torch.empty(H * W, dtype=dtype, device="cuda")
            _lib.check(L.dsb_init_canvas(op, canvas.data_ptr(), canvas.numel(), stream))
            plan = _lib.Plan()
            plan.nops = 1
            plan.ops[0].op, plan.ops[0].agg = op, canvas.data_ptr()
            if use_chk:
                plan.ops[0].chk_dtype, plan.ops[0].chk = _lib.F32, v.data_ptr()
            else:
                plan.ops[0].val_dtype, plan.ops[0].val = _lib.F32, v.data_ptr()
            if routed_call:
                scratch = torch.empty(small, dtype=torch.uint8, device="cuda")
                _lib.check(L.dsb_points_routed(C.byref(view), x.data_ptr(), y.data_ptr(), _lib.F32, n, 7, C.byref(plan), scratch.data_ptr(),
                                               small, stream), "dsb_points_routed")
            else:
                _lib.check(L.dsb_points(C.byref(view), x.data_ptr(), y.data_ptr(), _lib.F32, n, 7, C.byref(plan), stream), "dsb_points")
            outs.append(canvas.cpu().numpy())
        assert np.array_equal(outs[0], outs[1]), f"op {op}"


def test_routed_everything_filtered(routed):
    """Every row lands in the dummy bucket (out of range, or a NaN value): canvases stay at their initial state."""
    import torch
    from oracle import oracle as ora
    ds = routed
    rng = np.random.default_rng(3)
    n, W, H = 70_001, 301, 257
    cols = {"x": (rng.random(n, dtype=np.float32) + 2.0), "y": rng.random(n, dtype=np.float32),
            "v32": rng.standard_normal(n).astype(np.float32)}
    cols2 = {"x": rng.random(n, dtype=np.float32), "y": rng.random(n, dtype=np.float32), "v32": np.full(n, np.nan, np.float32)}
    view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    for c in (cols, cols2):
        frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in c.items()})
        for name in ("max_v32", "first_v32", "count_v32"):
            got = cvs.points(frame, "x", "y", make_agg(SPECS[name])).data
            assert_agg_equal(got, ora.points(c, "x", "y", SPECS[name], view), f"routed all-filtered {name}")


def test_routed_declines_canvases_with_too_many_buckets(routed):
    """24 000 x 24 000 cells = 12 784 buckets: their tables do not fit pass 1's shared memory; the call must fall back to the
    banded kernels and still be right (compared with the unbanded generic kernel)."""
    import torch
    from datashader_b200 import _lib
    ds = routed
    rng = np.random.default_rng(9)
    n, W, H = 200_000, 24_000, 24_000
    cols = {"x": rng.random(n, dtype=np.float32), "y": rng.random(n, dtype=np.float32), "v32": rng.standard_normal(n).astype(np.float32)}
    frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items()})
    cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    ds.config.device_results = True
    try:
        got = cvs.points(frame, "x", "y", ds.max("v32")).data
        assert b"k_route" not in _lib.lib().dsb_last_kernel()
        xi = np.minimum((cols["x"].astype(np.float64) * W).astype(np.int64), W - 1)
        yi = np.minimum((cols["y"].astype(np.float64) * H).astype(np.int64), H - 1)
        cells = torch.from_numpy(yi * W + xi).cuda()
        want = torch.full((H * W,), float("-inf"), dtype=torch.float32, device="cuda")
        want.scatter_reduce_(0, cells, torch.from_numpy(cols["v32"]).cuda(), "amax")
        hit = torch.isfinite(want)
        assert int(hit.sum()) > 0.99 * n
        assert torch.equal(got.reshape(-1)[hit].float(), want[hit]) and bool(torch.isnan(got.reshape(-1)[~hit]).all())
    finally:
        ds.config.device_results = False


def test_routed_negzero(routed):
    import torch
    from oracle import oracle as ora
    ds = routed
    rng = np.random.default_rng(11)
    n, W, H = 150_000, 301, 257
    zeros = np.where(rng.random(n) < 0.5, np.float32(-0.0), np.float32(0.0))
    cols = {"x": rng.random(n, dtype=np.float32), "y": rng.random(n, dtype=np.float32),
            "v32": np.where(rng.random(n) < 0.5, zeros, -rng.random(n).astype(np.float32) - 0.1).astype(np.float32)}
    frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items()})
    got = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0)).points(frame, "x", "y", ds.max("v32")).data
    want = ora.points(cols, "x", "y", ("max", "v32"), ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0)))
    assert np.signbit(want[want == 0]).any()
    assert_agg_equal(got, want, "routed negzero max")


@pytest.mark.parametrize("shape,n,clustered", [((301, 257), 200_003, False), ((640, 480), 250_000, True)])
def test_where_two_pass_matches_oracle(routed, shape, n, clustered):
    """where(max | min) as two passes (the routed extreme, then dsb_points_match32 with its coarse filter) against the oracle:
    values rounded to 0.1, so nearly every pixel's extreme is held by several rows and the earliest must win."""
    import torch
    from datashader_b200 import _lib
    from oracle import oracle as ora
    ds = routed
    W, H = shape
    cols = _cols(np.random.default_rng(n + 1), n, clustered)
    frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items()})
    view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    for queue in (1, 0):          # the queued shared-memory form of the second pass (shipped) and its first form
        _lib.check(_lib.lib().dsb_configure(b"match_queue", queue))
        for name in ("where_max_v32_other", "where_min_v32_row", "where_max_v32_row", "where_min_v32_other"):
            spec = SPECS.get(name) or ("where", (name.split("_")[1], "v32"), None if name.endswith("row") else "other")
            got = cvs.points(frame, "x", "y", make_agg(spec)).data
            assert b"k_points_match32<" in _lib.lib().dsb_last_kernel(), (name, _lib.lib().dsb_last_kernel())
            assert_agg_equal(got, ora.points(cols, "x", "y", spec, view), f"two-pass {name} {shape} clustered={clustered} queue={queue}")
    _lib.check(_lib.lib().dsb_configure(b"match_queue", 1))
    # summary sharing the extreme's canvas with the where()
    both = cvs.points(frame, "x", "y", ds.summary(m=ds.max("v32"), w=ds.where(ds.max("v32"), "other")))
    assert_agg_equal(both["m"].data, ora.points(cols, "x", "y", ("max", "v32"), view), "summary max")
    assert_agg_equal(both["w"].data, ora.points(cols, "x", "y", ("where", ("max", "v32"), "other"), view), "summary where")
    # a frame 1e7 from the origin: the float32 fast mapping is off, every row takes the exact mapping (k_points_match32_exact)
    far = dict(cols)
    far["x"] = (cols["x"] * 64 + np.float32(1e7)).astype(np.float32)
    fframe = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in far.items()})
    fview = ora.make_view(64, 48, (1e7, 1e7 + 64.0), (0.0, 1.0))
    got = ds.Canvas(64, 48, x_range=(1e7, 1e7 + 64.0), y_range=(0.0, 1.0)).points(fframe, "x", "y", ds.where(ds.max("v32"))).data
    assert b"k_points_match32_exact" in _lib.lib().dsb_last_kernel(), _lib.lib().dsb_last_kernel()
    assert_agg_equal(got, ora.points(far, "x", "y", ("where", ("max", "v32"), None), fview), "two-pass exact mapping")
    # -0.0 and +0.0 tie (reductions.py:1224: strict compare): the earliest of them is the row
    z = {"x": np.full(6, 0.5, np.float32), "y": np.full(6, 0.5, np.float32),
         "v32": np.array([-1.0, -0.0, 0.0, -0.0, np.nan, -2.0], np.float32), "other": np.arange(6, dtype=np.float32)}
    zf = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in z.items()})
    got = ds.Canvas(2, 2, x_range=(0.0, 1.0), y_range=(0.0, 1.0)).points(zf, "x", "y", ds.where(ds.max("v32"))).data
    assert got[1, 1] == 1 and (got.ravel()[[0, 1, 2]] == -1).all()


@pytest.mark.parametrize("order", ["shuffled", "sorted_by_y", "reverse_sorted"])
def test_first_last_head_then_filtered_rest(routed, order):
    """first / last / where(first | last) with the rows split into a routed head (tail) and a filtered rest (k_rows_rest), forced
    on at small n: shuffled rows (the head settles nearly every pixel: the filter does the rest) and rows sorted in space (the head
    covers a strip of the canvas, the device-side sample hands the rest to the routed kernels) against the oracle."""
    import torch
    from datashader_b200 import _lib
    from oracle import oracle as ora
    ds = routed
    L = _lib.lib()
    W, H, n = 301, 257, 400_003
    cols = _cols(np.random.default_rng(99), n)
    if order != "shuffled":
        idx = np.argsort(cols["y"], kind="stable")
        if order == "reverse_sorted":
            idx = idx[::-1]
        cols = {k: np.ascontiguousarray(v[idx]) for k, v in cols.items()}
    frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items()})
    view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    _lib.check(L.dsb_configure(b"routed_head_per_cell", 1))          # head = one row per canvas cell (77 357 rows)
    try:
        for name in ("first_v32", "last_v32", "where_first_v32_other", "where_last_v32_row"):
            got = cvs.points(frame, "x", "y", make_agg(SPECS[name])).data
            assert b"k_rows_rest<" in L.dsb_last_kernel(), (name, L.dsb_last_kernel())
            assert_agg_equal(got, ora.points(cols, "x", "y", SPECS[name], view, npartitions=2), f"head + rest {name} {order}")
    finally:
        _lib.check(L.dsb_configure(b"routed_head_per_cell", 10))


def test_first_last_split_on_an_l2_resident_canvas():
    """first / last with many rows per pixel take the head + filtered-rest form on canvases that fit L2 as well (pipeline gate
    `routed_rows_per_cell_for_first`): the real L2 budget, 301 x 257, against the oracle - and against the mono kernel."""
    import torch
    import datashader_b200 as ds
    from datashader_b200 import _lib
    from oracle import oracle as ora
    L = _lib.lib()
    W, H, n = 301, 257, 500_001
    cols = _cols(np.random.default_rng(7), n)
    frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items()})
    view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    old = (ds.config.routed_min_rows, ds.config.routed_rows_per_cell_for_first)
    ds.config.routed_min_rows, ds.config.routed_rows_per_cell_for_first = 0, 2
    _lib.check(L.dsb_routed_configure(0))
    _lib.check(L.dsb_configure(b"routed_head_per_cell", 1))
    try:
        for name in ("first_v32", "last_v32", "where_first_v32_other", "where_last_v32_row"):
            got = cvs.points(frame, "x", "y", make_agg(SPECS[name])).data
            assert b"k_rows_rest<" in L.dsb_last_kernel(), (name, L.dsb_last_kernel())
            assert_agg_equal(got, ora.points(cols, "x", "y", SPECS[name], view, npartitions=2), f"L2-resident head + rest {name}")
        # rows sorted in space: the head covers a strip of the canvas, the device-side sample hands the rest to the routed kernels -
        # with a record buffer the pipeline sized for the head only (L2-resident canvas): the overflow goes to direct atomics
        idx = np.argsort(cols["y"], kind="stable")
        scols = {k: np.ascontiguousarray(v[idx]) for k, v in cols.items()}
        sframe = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in scols.items()})
        for name in ("first_v32", "last_v32"):
            got = cvs.points(sframe, "x", "y", make_agg(SPECS[name])).data
            assert b"k_rows_rest<" in L.dsb_last_kernel(), (name, L.dsb_last_kernel())
            assert_agg_equal(got, ora.points(scols, "x", "y", SPECS[name], view, npartitions=2), f"L2-resident, sorted rows {name}")
        # the same frame streamed from the host in two row chunks: every chunk call splits on its own, the canvas carries the
        # earlier chunks' rows (first: they settle their pixels; last: the later chunk's rows must still replace them)
        hf = ds.HostFrame(cols)
        chunk_rows = ds.HostFrame.CHUNK_ROWS
        ds.HostFrame.CHUNK_ROWS = 250_001
        try:
            for name in ("first_v32", "last_v32", "where_last_v32_row"):
                got = cvs.points(hf, "x", "y", make_agg(SPECS[name])).data
                assert b"k_rows_rest<" in L.dsb_last_kernel(), (name, L.dsb_last_kernel())
                assert_agg_equal(got, ora.points(cols, "x", "y", SPECS[name], view, npartitions=2), f"chunked head + rest {name}")
        finally:
            ds.HostFrame.CHUNK_ROWS = chunk_rows
        got = cvs.points(frame, "x", "y", ds.max("v32")).data        # other reductions keep their kernels
        assert b"k_route" not in L.dsb_last_kernel() and b"k_rows_rest" not in L.dsb_last_kernel(), L.dsb_last_kernel()
        assert_agg_equal(got, ora.points(cols, "x", "y", ("max", "v32"), view), "max unaffected")
    finally:
        ds.config.routed_min_rows, ds.config.routed_rows_per_cell_for_first = old
        _lib.check(L.dsb_routed_configure(1 << 24))
        _lib.check(L.dsb_configure(b"routed_head_per_cell", 10))


@pytest.mark.parametrize("values", ["normal", "zeros"])
def test_max_min_head_then_threshold_filtered_rest(values):
    """max / min of a float32 column on an L2-resident canvas with many rows per pixel: dsb_points over the head of the rows,
    dsb_points_minmax_rest (block thresholds in shared memory, queued survivors) over the rest - forced on at small n, against the
    oracle bit for bit, including the sign of a zero extreme (-0.0 / +0.0 rows that tie)."""
    import torch
    import datashader_b200 as ds
    from datashader_b200 import _lib
    from oracle import oracle as ora
    L = _lib.lib()
    W, H, n = 301, 257, 600_003
    rng = np.random.default_rng(31)
    cols = _cols(rng, n)
    if values == "zeros":          # extremes that are zeros of either sign, in either order of arrival
        v = np.where(rng.random(n) < 0.5, -np.abs(cols["v32"]), 0.0).astype(np.float32)
        v[rng.random(n) < 0.3] *= np.float32(-1.0)        # -0.0 and positive values appear as well
        v[rng.random(n) < 0.02] = np.nan
        cols["v32"] = v
        assert np.signbit(v[v == 0]).any() and (~np.signbit(v[v == 0])).any()
    frame = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items()})
    view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
    cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    old = (ds.config.routed_min_rows, ds.config.minmax_split_rows_per_cell, ds.config.minmax_head_rows_per_cell)
    ds.config.routed_min_rows, ds.config.minmax_split_rows_per_cell, ds.config.minmax_head_rows_per_cell = 0, 2, 1
    try:
        for name in ("max_v32", "min_v32"):
            got = cvs.points(frame, "x", "y", make_agg(SPECS[name])).data
            if values == "normal":     # (with zeros the -0.0 redo runs last and leaves its own kernel name)
                assert b"k_points_minmax_rest<" in L.dsb_last_kernel(), (name, L.dsb_last_kernel())
            assert_agg_equal(got, ora.points(cols, "x", "y", SPECS[name], view), f"head + threshold rest {name} {values}")
        for name in ("where_max_v32_other", "where_min_v32_row"):      # the packed {key, row} accumulator through the same split
            got = cvs.points(frame, "x", "y", make_agg(SPECS[name])).data
            assert b"k_points_argminmax_rest<" in L.dsb_last_kernel(), (name, L.dsb_last_kernel())
            assert_agg_equal(got, ora.points(cols, "x", "y", SPECS[name], view), f"head + threshold rest {name} {values}")
        both = cvs.points(frame, "x", "y", ds.summary(a=ds.max("v32"), b=ds.min("v32"), c=ds.count()))
        assert_agg_equal(both["a"].data, ora.points(cols, "x", "y", ("max", "v32"), view), f"summary max {values}")
        assert_agg_equal(both["b"].data, ora.points(cols, "x", "y", ("min", "v32"), view), f"summary min {values}")
    finally:
        ds.config.routed_min_rows, ds.config.minmax_split_rows_per_cell, ds.config.minmax_head_rows_per_cell = old


def test_first_last_split_equals_banded_at_production_scale():
    """4096 x 4096 (134 MB of row ids: beyond L2), 1e8 points, head = 2 rows per cell: the split form against the L2-banded kernels."""
    import torch
    import datashader_b200 as ds
    from datashader_b200 import _lib
    L = _lib.lib()
    n = 100_000_000
    g = torch.Generator(device="cuda")
    g.manual_seed(6)
    x = torch.rand(n, generator=g, device="cuda") * 1.02 - 0.01
    y = torch.rand(n, generator=g, device="cuda") * 1.02 - 0.01
    v = torch.randn(n, generator=g, device="cuda")
    v[::13] = float("nan")
    frame = ds.DeviceFrame({"x": x, "y": y, "value": v})
    cvs = ds.Canvas(4096, 4096, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    ds.config.device_results = True
    _lib.check(L.dsb_configure(b"routed_head_per_cell", 2))
    try:
        for agg in (ds.first("value"), ds.last("value"), ds.where(ds.first("value"))):
            res = {}
            for mode in ("split", "banded"):
                ds.config.routed = mode == "split"
                res[mode] = cvs.points(frame, "x", "y", agg).data.clone()
                assert (b"k_rows_rest<" in L.dsb_last_kernel()) == (mode == "split"), (mode, L.dsb_last_kernel())
            a, b = res["split"], res["banded"]
            same = torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0)) if a.dtype.is_floating_point else torch.equal(a, b)
            assert same, f"{agg} split vs banded"
    finally:
        ds.config.device_results = False
        ds.config.routed = True
        _lib.check(L.dsb_configure(b"routed_head_per_cell", 10))


def test_routed_equals_banded_at_production_geometry():
    """8192 x 8192, 1e8 points, real budgets: the routed path, the L2-banded mono kernels and the unbanded generic kernel
    agree bit for bit on max / first / count (BASELINE config 5's geometry)."""
    import torch
    import datashader_b200 as ds
    from datashader_b200 import _lib
    L = _lib.lib()
    n = 100_000_000
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    x = torch.rand(n, generator=g, device="cuda") * 1.02 - 0.01
    y = torch.rand(n, generator=g, device="cuda") * 1.02 - 0.01
    v = torch.randn(n, generator=g, device="cuda")
    v[::997] = float("nan")
    frame = ds.DeviceFrame({"x": x, "y": y, "value": v})
    cvs = ds.Canvas(8192, 8192, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    ds.config.device_results = True
    try:
        for agg in (ds.max("value"), ds.first("value"), ds.count()):
            res = {}
            for mode in ("routed", "banded", "generic"):
                ds.config.routed = mode == "routed"
                _lib.check(L.dsb_configure(b"l2_band_bytes", 0 if mode == "generic" else 96 << 20))
                _lib.check(L.dsb_configure(b"mono", 0 if mode == "generic" else 1))
                res[mode] = cvs.points(frame, "x", "y", agg).data.clone()
                kern = L.dsb_last_kernel()
                assert (b"k_route" in kern) == (mode == "routed"), (mode, kern)
            for mode in ("banded", "generic"):
                a, b = res["routed"], res[mode]
                same = torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0)) if a.dtype.is_floating_point else torch.equal(a, b)
                assert same, f"{agg} routed vs {mode}"
            del res
        # where(max): the two-pass form (routed max + dsb_points_match32) against the packed {key, row} accumulator, L2-banded
        rows = {}
        ds.config.routed = True
        _lib.check(L.dsb_configure(b"l2_band_bytes", 96 << 20))
        _lib.check(L.dsb_configure(b"mono", 1))
        for two in (True, False):
            ds.config.where_two_pass = two
            rows[two] = cvs.points(frame, "x", "y", ds.where(ds.max("value"))).data.clone()
            assert (b"k_points_match32<" in L.dsb_last_kernel()) == two, L.dsb_last_kernel()
        assert torch.equal(rows[True], rows[False]), "where(max): two-pass vs packed accumulator"
    finally:
        ds.config.device_results = False
        ds.config.routed = True
        ds.config.where_two_pass = True
        _lib.check(L.dsb_configure(b"l2_band_bytes", 96 << 20))
        _lib.check(L.dsb_configure(b"mono", 1))
