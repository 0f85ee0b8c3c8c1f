"""Tile-pyramid driver above the hot path: the same entry points as the reference's datashader/tiles.py
(`render_tiles`, `MercatorTileDefinition`, `gen_super_tiles`, `calculate_zoom_level_stats`), re-aggregating the data once
per SUPER TILE (up to 4096 x 4096 pixels, tiles.py:99-117) with this package's Canvas and cutting the shaded super tile
into 256 x 256 tiles (`{output_path}/{z}/{x}/{y}.png`, tiles.py:386-394).

Differences from the reference: no dask (the zoom-level span is a plain min / max over the super tiles, which the
reference delegates to dask.bag, tiles.py:57-62), tiles are cut by pixel block instead of `DataArray.loc` (the super tile's
pixel grid is the tile grid, so both select the same pixels), and `Canvas.points_batch` offers the other route - every
tile of a level as its own canvas in one pass over the data - for callers that need per-tile aggregates rather than images.
"""
from __future__ import annotations

import math
import os

import numpy as np

__all__ = ["render_tiles", "MercatorTileDefinition", "gen_super_tiles", "calculate_zoom_level_stats", "tile_views"]

_ORIGIN = 20037508.34            # half the web-mercator world, metres (tiles.py:179-181)
_RES0 = 156543.03392804097       # metres per pixel at zoom 0 with 256-pixel tiles


def invert_y_tile(y, z):
    """TMS <-> Google tile row (tiles.py:132-134)."""
    return (2 ** z) - 1 - y


class MercatorTileDefinition:
    """Web-mercator tile arithmetic (tiles.py:138-300): metres <-> pixels <-> tile indices for square tiles."""

    def __init__(self, x_range, y_range, tile_size=256, min_zoom=0, max_zoom=30, x_origin_offset=_ORIGIN,
                 y_origin_offset=_ORIGIN, initial_resolution=_RES0):
        self.x_range, self.y_range = x_range, y_range
        self.tile_size = tile_size
        self.min_zoom, self.max_zoom = min_zoom, max_zoom
        self.x_origin_offset, self.y_origin_offset = x_origin_offset, y_origin_offset
        self.initial_resolution = initial_resolution
        self._resolutions = [self._get_resolution(z) for z in range(min_zoom, max_zoom + 1)]

    def is_valid_tile(self, x, y, z):
        n = math.pow(2, z)
        return 0 <= x < n and 0 <= y < n

    def _get_resolution(self, z):
        return self.initial_resolution / (2 ** z)

    def get_resolution_by_extent(self, extent, height, width):
        return [(extent[2] - extent[0]) / width, (extent[3] - extent[1]) / height]

    def get_level_by_extent(self, extent, height, width):
        resolution = max(self.get_resolution_by_extent(extent, height, width))
        for i, r in enumerate(self._resolutions):
            if resolution > r:
                return max(i - 1, 0)
        return len(self._resolutions) - 1

    def pixels_to_meters(self, px, py, level):
        res = self._get_resolution(level)
        return (px * res) - self.x_origin_offset, (py * res) - self.y_origin_offset

    def meters_to_pixels(self, mx, my, level):
        res = self._get_resolution(level)
        return (mx + self.x_origin_offset) / res, (my + self.y_origin_offset) / res

    def pixels_to_tile(self, px, py, level):
        tx = math.ceil(px / self.tile_size)
        tx = tx if tx == 0 else tx - 1
        ty = max(math.ceil(py / self.tile_size) - 1, 0)
        return int(tx), invert_y_tile(int(ty), level)

    def pixels_to_raster(self, px, py, level):
        return px, (self.tile_size << level) - py

    def meters_to_tile(self, mx, my, level):
        return self.pixels_to_tile(*self.meters_to_pixels(mx, my, level), level)

    def get_tile_meters(self, tx, ty, level):
        ty = invert_y_tile(ty, level)
        xmin, ymin = self.pixels_to_meters(tx * self.tile_size, ty * self.tile_size, level)
        xmax, ymax = self.pixels_to_meters((tx + 1) * self.tile_size, (ty + 1) * self.tile_size, level)
        return xmin, ymin, xmax, ymax

    def get_tiles_by_extent(self, extent, level):
        xmin, ymin, xmax, ymax = extent
        txmin, tymax = self.meters_to_tile(xmin, ymin, level)          # tile rows run opposite to metres
        txmax, tymin = self.meters_to_tile(xmax, ymax, level)
        return [(tx, ty, level, self.get_tile_meters(tx, ty, level))
                for ty in range(tymin, tymax + 1) for tx in range(txmin, txmax + 1) if self.is_valid_tile(tx, ty, level)]


def gen_super_tiles(extent, zoom_level, span=None):
    """Super tiles of a zoom level: min(16, 2^z) x 256 pixels on a side (tiles.py:99-117)."""
    xmin, ymin, xmax, ymax = extent
    size = min(2 ** 4 * 256, (2 ** zoom_level) * 256)
    tile_def = MercatorTileDefinition(x_range=(xmin, xmax), y_range=(ymin, ymax), tile_size=size)
    for s in tile_def.get_tiles_by_extent(extent, zoom_level):
        e = s[3]
        yield {"level": zoom_level, "x_range": (e[0], e[2]), "y_range": (e[1], e[3]), "tile_size": size, "span": span}


def calculate_zoom_level_stats(super_tiles, load_data_func, rasterize_func, color_ranging_strategy="fullscan"):
    """Aggregate every super tile once (kept in tile['agg']) and return the level's colour span (tiles.py:40-67)."""
    if color_ranging_strategy != "fullscan":
        raise ValueError("Invalid color_ranging_strategy option")
    lo, hi, is_bool = math.inf, -math.inf, False
    for t in super_tiles:
        df = load_data_func(t["x_range"], t["y_range"])
        agg = rasterize_func(df, x_range=t["x_range"], y_range=t["y_range"], height=t["tile_size"], width=t["tile_size"])
        t["agg"] = agg
        data = np.asarray(agg.data)
        if data.dtype.kind == "b":
            is_bool = True
        elif data.size and not np.all(np.isnan(data.astype(np.float64))):
            lo, hi = min(lo, float(np.nanmin(data))), max(hi, float(np.nanmax(data)))
    return super_tiles, ((0, 1) if is_bool else (lo, hi))


def _tile_blocks(img, tile_info, level):
    """(tile image array, x, y, z) for each 256-pixel tile of a shaded super tile; rows flipped to run downwards like map
    tiles (tiles.py:318-346)."""
    tile_def = MercatorTileDefinition(x_range=tile_info["x_range"], y_range=tile_info["y_range"], tile_size=256)
    xmin, xmax = tile_info["x_range"]
    ymin, ymax = tile_info["y_range"]
    data = np.asarray(img.data)
    xs, ys = np.asarray(img.coords[img.dims[1]]), np.asarray(img.coords[img.dims[0]])
    for tx, ty, z, (dxmin, dymin, dxmax, dymax) in tile_def.get_tiles_by_extent((xmin, ymin, xmax, ymax), level):
        cols = np.nonzero((xs >= dxmin) & (xs <= dxmax))[0]
        rows = np.nonzero((ys >= dymin) & (ys <= dymax))[0]
        if cols.size == 0 or rows.size == 0:
            continue
        block = data[rows[0]:rows[-1] + 1, cols[0]:cols[-1] + 1]
        yield np.flip(block, 0), tx, ty, z


def render_super_tile(tile_info, span, output_path, shader_func, post_render_func, tile_format="PNG"):
    """Shade one super tile and write its tiles as {output_path}/{z}/{x}/{y}.png (tiles.py:120-129, 386-394)."""
    from PIL.Image import fromarray
    img = shader_func(tile_info["agg"], span=span)
    written = 0
    for block, x, y, z in _tile_blocks(img, tile_info, tile_info["level"]):
        if block.ndim == 2 and block.dtype == np.uint32:
            block = np.ascontiguousarray(block).view(np.uint8).reshape(block.shape + (4,))      # RGBA bytes
        pil = fromarray(block)
        if post_render_func:
            pil = post_render_func(pil, x=x, y=y, z=z)
        d = os.path.join(output_path, str(z), str(x))
        os.makedirs(d, exist_ok=True)
        pil.save(os.path.join(d, f"{y}.{tile_format.lower()}"), tile_format)
        written += 1
    return written


def render_tiles(full_extent, levels, load_data_func, rasterize_func, shader_func, post_render_func, output_path,
                 color_ranging_strategy="fullscan"):
    """tiles.py:70-96: for every zoom level aggregate the super tiles (fixing the level's colour span), then shade and cut
    them.  rasterize_func(df, x_range, y_range, height, width) -> aggregate; shader_func(agg, span=) -> image."""
    results = {}
    for level in levels:
        super_tiles, span = calculate_zoom_level_stats(list(gen_super_tiles(full_extent, level)), load_data_func,
                                                       rasterize_func, color_ranging_strategy=color_ranging_strategy)
        tiles = sum(render_super_tile(t, span, output_path, shader_func, post_render_func) for t in super_tiles)
        results[level] = dict(success=True, stats=span, supertile_count=len(super_tiles), tile_count=tiles)
    return results


def tile_views(full_extent, level, tile_size=256):
    """The (x_range, y_range) of every tile of a zoom level inside full_extent, row-major from the bottom-left, plus the grid
    shape - the `views` / `grid` arguments of Canvas.points_batch, which fills all of them in one pass over the data."""
    tile_def = MercatorTileDefinition(x_range=(full_extent[0], full_extent[2]), y_range=(full_extent[1], full_extent[3]),
                                      tile_size=tile_size)
    tiles = tile_def.get_tiles_by_extent(full_extent, level)
    txs, tys = sorted({t[0] for t in tiles}), sorted({t[1] for t in tiles}, reverse=True)      # Google rows run downwards
    by_index = {(t[0], t[1]): t[3] for t in tiles}
    views = []
    for ty in tys:
        for tx in txs:
            e = by_index[(tx, ty)]
            views.append(((e[0], e[2]), (e[1], e[3])))
    return views, (len(txs), len(tys))
