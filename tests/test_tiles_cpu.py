"""datashader_b200.tiles against the reference's tile arithmetic (tests/golden/tiles.npz, generated from
datashader/tiles.py by tests/golden/make_golden.py): super tiles and 256-pixel tiles of several extents and levels."""
import numpy as np

from helpers import load


def test_tile_arithmetic_matches_reference():
    from datashader_b200.tiles import MercatorTileDefinition, gen_super_tiles
    g = load("tiles.npz")
    n = 0
    for key in g.files:
        if not key.startswith("tiles_"):
            continue
        _, name, level = key.split("_")
        ext = tuple(g[f"extent_{name}"])
        td = MercatorTileDefinition(x_range=(ext[0], ext[2]), y_range=(ext[1], ext[3]), tile_size=256)
        got = np.array([[t[0], t[1], t[2], *t[3]] for t in td.get_tiles_by_extent(ext, int(level))], dtype=np.float64).reshape(-1, 7)
        assert np.array_equal(got, g[key]), key
        sup = np.array([[s["level"], s["tile_size"], *s["x_range"], *s["y_range"]] for s in gen_super_tiles(ext, int(level))],
                       dtype=np.float64).reshape(-1, 6)
        assert np.array_equal(sup, g[f"super_{name}_{level}"]), key
        n += 1
    assert n >= 10


def test_tile_views_form_a_row_major_grid():
    from datashader_b200.tiles import tile_views
    ext = (-20037508.34, -20037508.34, 20037508.34, 20037508.34)
    views, (nx, ny) = tile_views(ext, 2)
    assert (nx, ny) == (4, 4) and len(views) == 16
    (x0, x1), (y0, y1) = views[0]
    for k, ((a, b), (c, d)) in enumerate(views):
        ix, iy = k % nx, k // nx
        assert np.isclose(a, x0 + ix * (x1 - x0)) and np.isclose(c, y0 + iy * (y1 - y0))
        assert np.isclose(b - a, x1 - x0) and np.isclose(d - c, y1 - y0)
