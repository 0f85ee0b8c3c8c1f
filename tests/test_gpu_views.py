"""Canvas.points_batch (dsb_points_views): every view of a batch equals the single Canvas.points call for that view, bit
for bit (means: rtol 1e-12) - a 4 x 4 tile grid with points exactly on shared tile edges, and a short list of arbitrary,
overlapping views; plus the tile-pyramid driver (datashader_b200.tiles.render_tiles) end to end."""
import numpy as np
import pytest

from helpers import assert_agg_equal

pytestmark = pytest.mark.gpu


def _frame(ds, n, seed):
    import torch
    rng = np.random.default_rng(seed)
    x = (rng.random(n) * 4.4 - 0.2).astype(np.float32)
    y = (rng.random(n) * 4.4 - 0.2).astype(np.float32)
    k = 4000
    x[:k] = rng.integers(0, 5, k).astype(np.float32)            # exactly on the vertical tile edges 0, 1, 2, 3, 4
    y[k:2 * k] = rng.integers(0, 5, k).astype(np.float32)
    x[2 * k:3 * k] = rng.integers(0, 5, k).astype(np.float32)   # tile corners
    y[2 * k:3 * k] = rng.integers(0, 5, k).astype(np.float32)
    v = rng.standard_normal(n).astype(np.float32)
    v[rng.integers(0, n, n // 50)] = np.nan
    cols = {"x": x, "y": y, "v": v, "cat": rng.integers(0, 3, n).astype(np.int8)}
    return ds.DeviceFrame({k_: torch.from_numpy(a).cuda() for k_, a in cols.items()}, categories={"cat": ["a", "b", "c"]})


AGGS = ["count", "mean", "max", "where_max", "first", "by_count", "summary"]


def _agg(ds, name):
    return {"count": ds.count(), "mean": ds.mean("v"), "max": ds.max("v"), "where_max": ds.where(ds.max("v")), "first": ds.first("v"),
            "by_count": ds.by("cat", ds.count()), "summary": ds.summary(n=ds.count(), m=ds.min("v"))}[name]


def _same(a, b, name):
    if hasattr(a, "data"):
        assert tuple(a.dims) == tuple(b.dims), name
        for d in a.dims[:2]:
            np.testing.assert_array_equal(a.coords[d], b.coords[d], err_msg=name)
        assert a.attrs == b.attrs, name
        assert_agg_equal(a.data, b.data, name)
    else:
        for k in a:
            _same(a[k], b[k], f"{name}.{k}")


@pytest.mark.parametrize("l2_budget", [None, 1, 64 * 48 * 4 * 12 * 4], ids=["one-launch", "band-per-tile-row", "bands-of-rows"])
def test_tile_grid_equals_single_calls(l2_budget):
    """l2_budget: the stacked canvases 'exceed L2' and the level is done in bands of tile rows (one launch per band)."""
    import datashader_b200 as ds
    frame = _frame(ds, 300_000, 3)
    cvs = ds.Canvas(64, 48)
    views = [((float(ix), float(ix + 1)), (float(iy), float(iy + 1))) for iy in range(4) for ix in range(4)]
    old = ds.config.l2_budget_bytes
    for name in AGGS:
        if l2_budget is not None:
            ds.config.l2_budget_bytes = l2_budget
        try:
            got = cvs.points_batch(frame, "x", "y", _agg(ds, name), views, grid=(4, 4))
        finally:
            ds.config.l2_budget_bytes = old
        assert len(got) == 16
        for (xr, yr), g in zip(views, got):
            want = ds.Canvas(64, 48, x_range=xr, y_range=yr).points(frame, "x", "y", _agg(ds, name))
            _same(g, want, f"{name} tile {xr} {yr}")


def test_grid_views_must_form_the_grid():
    import datashader_b200 as ds
    frame = _frame(ds, 300_000, 3)
    views = [((float(ix), float(ix + 1)), (float(iy), float(iy + 1))) for iy in range(2) for ix in range(2)]
    views[3] = ((1.0, 2.5), (1.0, 2.0))
    with pytest.raises(ValueError, match="grid"):
        ds.Canvas(16, 16).points_batch(frame, "x", "y", ds.count(), views, grid=(2, 2))


def test_arbitrary_views_equal_single_calls():
    import datashader_b200 as ds
    frame = _frame(ds, 200_000, 4)
    cvs = ds.Canvas(50, 30)
    views = [((0.0, 4.0), (0.0, 4.0)), ((1.5, 2.5), (0.25, 3.75)), ((-1.0, 0.5), (3.0, 9.0)), ((2.0, 2.0625), (2.0, 2.0625))]
    for name in ("count", "mean", "where_max", "by_count"):
        got = cvs.points_batch(frame, "x", "y", _agg(ds, name), views)
        for (xr, yr), g in zip(views, got):
            want = ds.Canvas(50, 30, x_range=xr, y_range=yr).points(frame, "x", "y", _agg(ds, name))
            _same(g, want, f"{name} view {xr} {yr}")
    with pytest.raises(ValueError):
        cvs.points_batch(frame, "x", "y", ds.count(), views * 20)


def test_render_tiles_pyramid(tmp_path):
    """render_tiles (tiles.py:70-96): levels 0-2 of a small web-mercator extent; every written tile equals the matching
    256 x 256 block of the shaded super tile."""
    import torch
    from PIL import Image
    import datashader_b200 as ds
    from datashader_b200.tiles import MercatorTileDefinition, gen_super_tiles, render_tiles
    rng = np.random.default_rng(9)
    half = 20037508.34
    n = 400_000
    frame = ds.DeviceFrame({"x": torch.from_numpy((rng.normal(0, 0.3, n) * half).astype(np.float32)).cuda(),
                            "y": torch.from_numpy((rng.normal(0, 0.3, n) * half).astype(np.float32)).cuda()})
    extent = (-half, -half, half, half)

    def rasterize(df, x_range, y_range, height, width):
        return ds.Canvas(width, height, x_range=x_range, y_range=y_range).points(df, "x", "y", ds.count())

    def shader(agg, span=None):
        return ds.tf.shade(agg, how="linear", span=span)

    res = render_tiles(extent, range(3), lambda xr, yr: frame, rasterize, shader, None, str(tmp_path))
    assert [res[z]["supertile_count"] for z in range(3)] == [1, 1, 1] and [res[z]["tile_count"] for z in range(3)] == [1, 4, 16]
    level = 2
    st = next(gen_super_tiles(extent, level))
    img = np.asarray(shader(rasterize(frame, st["x_range"], st["y_range"], st["tile_size"], st["tile_size"]), span=res[level]["stats"]).data)
    td = MercatorTileDefinition(x_range=st["x_range"], y_range=st["y_range"], tile_size=256)
    for tx, ty, z, _e in td.get_tiles_by_extent(extent, level):
        tile = np.asarray(Image.open(tmp_path / str(z) / str(tx) / f"{ty}.png"))
        rows = slice((3 - ty) * 256, (4 - ty) * 256)           # Google rows run downwards, canvas rows upwards
        block = np.flip(img[rows, tx * 256:(tx + 1) * 256], 0)
        assert np.array_equal(tile, np.ascontiguousarray(block).view(np.uint8).reshape(256, 256, 4)), (tx, ty)
