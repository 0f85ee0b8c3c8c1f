"""The user API of the hot path, kept call-compatible with datashader/core.py:
Canvas(plot_width, plot_height, x_range, y_range, x_axis_type, y_axis_type) with .points / .line / .area
returning the same DataArray aggregates (dims, coords, attrs, dtypes)."""
from __future__ import annotations

import warnings
from math import log10
from numbers import Number

import numpy as np

from . import reductions as rd
from .glyphs import (AreaGlyph, LineAxis0, LineAxis0Multi, LinesAxis1, LinesAxis1Ragged, LinesAxis1XConstant,
                     LinesAxis1YConstant, Point, _LineGlyph)


class Axis:
    """core.py:35-111"""
    is_log = False

    def compute_scale_and_translate(self, range, n):   # noqa: A002
        start, end = map(self.mapper, range)
        s = n / (end - start)
        t = -start * s
        return s, t

    def compute_index(self, st, n):
        px = np.arange(n) + 0.5
        s, t = st
        return self.inverse_mapper((px - t) / s)

    def validate(self, range):   # noqa: A002
        pass


class LinearAxis(Axis):
    """core.py:114-124"""
    @staticmethod
    def mapper(val):
        return val

    @staticmethod
    def inverse_mapper(val):
        return val


class LogAxis(Axis):
    """core.py:127-145"""
    is_log = True

    @staticmethod
    def mapper(val):
        return log10(float(val))

    @staticmethod
    def inverse_mapper(val):
        y = 10
        return y ** val

    def validate(self, range):   # noqa: A002
        if range is None:
            return
        if range[0] <= 0 or range[1] <= 0:
            raise ValueError('Range values must be >0 for logarithmic axes')


_axis_lookup = {'linear': LinearAxis(), 'log': LogAxis()}


def validate_xy_or_geometry(glyph, x, y, geometry):
    """core.py:151-160"""
    if (geometry is None and (x is None or y is None) or
            geometry is not None and (x is not None or y is not None)):
        raise ValueError(f"""
{glyph} coordinates may be specified by providing both the x and y arguments, or by
providing the geometry argument. Received:
    x: {x!r}
    y: {y!r}
    geometry: {geometry!r}
""")


def _broadcast_column_specifications(*args):
    """core.py:1433-1443"""
    lengths = {len(a) for a in args if isinstance(a, (list, tuple))}
    if len(lengths) != 1:
        return args
    n = lengths.pop()
    return tuple((arg,) * n if isinstance(arg, (Number, str)) else arg for arg in args)


class Canvas:
    """An abstract canvas representing the space in which to bin (core.py:163-185)."""

    def __init__(self, plot_width=600, plot_height=600, x_range=None, y_range=None,
                 x_axis_type='linear', y_axis_type='linear'):
        self.plot_width = plot_width
        self.plot_height = plot_height
        self.x_range = None if x_range is None else tuple(x_range)
        self.y_range = None if y_range is None else tuple(y_range)
        self.x_axis = _axis_lookup[x_axis_type]
        self.y_axis = _axis_lookup[y_axis_type]

    # ---------------------------------------------------------------------------- points
    def points(self, source, x=None, y=None, agg=None, geometry=None):
        """Compute a reduction by pixel, mapping data to pixels as points (core.py:187-232).

        source: pandas.DataFrame (columns are staged to the current CUDA device) or a
        datashader_b200.DeviceFrame (device-resident columns, optionally one shard of a multi-GPU job).
        """
        validate_xy_or_geometry('Point', x, y, geometry)
        if agg is None:
            agg = rd.count()
        if geometry is not None:
            raise NotImplementedError("geometry= sources (spatialpandas/geopandas) are outside the B200 hot path")
        glyph = Point(x, y)
        return bypixel(source, self, glyph, agg)

    def points_batch(self, source, x, y, agg=None, views=(), grid=None):
        """Canvas.points for many views of this canvas' size in ONE pass over the data: the batched-viewport form of the
        re-aggregation loops above the hot path (tiles.py:70-131 renders every zoom level tile by tile, pipeline.py:55-72
        re-aggregates on every zoom / pan).  views: [(x_range, y_range), ...]; grid=(nx, ny) when they form a row-major
        tile grid.  Returns a list with, per view, exactly what Canvas(w, h, x_range, y_range).points(...) returns."""
        from . import pipeline
        from .distributed import current_group
        validate_xy_or_geometry('Point', x, y, None)
        if agg is None:
            agg = rd.count()
        return pipeline.points_batch(source, self, Point(x, y), agg, views, grid=grid, dist=current_group(source))

    # ---------------------------------------------------------------------------- lines
    def line(self, source, x=None, y=None, agg=None, axis=0, geometry=None, line_width=0, antialias=False):
        """Compute a reduction by pixel, mapping data to pixels as one or more lines (core.py:234-478)."""
        validate_xy_or_geometry('Line', x, y, geometry)
        if agg is None:
            agg = rd.any()
        if line_width is None:
            line_width = 0
        if antialias and line_width != 0:
            raise ValueError(
                "Do not specify values for both the line_width and \n"
                "antialias keyword arguments; use line_width instead.")
        if antialias:
            line_width = 1.0
        if geometry is not None:
            raise NotImplementedError("geometry= sources (spatialpandas/geopandas) are outside the B200 hot path")
        orig_x, orig_y = x, y
        x, y = _broadcast_column_specifications(x, y)
        if axis == 0:
            if isinstance(x, (Number, str)) and isinstance(y, (Number, str)):
                glyph = LineAxis0(x, y)
            elif isinstance(x, (list, tuple)) and isinstance(y, (list, tuple)):
                glyph = LineAxis0Multi(tuple(x), tuple(y))
            else:
                raise ValueError(f"""
Invalid combination of x and y arguments to Canvas.line when axis=0.
    Received:
        x: {repr(orig_x)}
        y: {repr(orig_y)}
See docstring for more information on valid usage""")
        elif axis == 1:
            if isinstance(x, (list, tuple)) and isinstance(y, (list, tuple)):
                glyph = LinesAxis1(tuple(x), tuple(y))
            elif isinstance(x, np.ndarray) and isinstance(y, (list, tuple)):
                glyph = LinesAxis1XConstant(x, tuple(y))
            elif isinstance(x, (list, tuple)) and isinstance(y, np.ndarray):
                glyph = LinesAxis1YConstant(tuple(x), y)
            elif isinstance(x, (Number, str)) and isinstance(y, (Number, str)):
                glyph = LinesAxis1Ragged(x, y)
            else:
                raise ValueError(f"""
Invalid combination of x and y arguments to Canvas.line when axis=1.
    Received:
        x: {repr(orig_x)}
        y: {repr(orig_y)}
See docstring for more information on valid usage""")
        else:
            raise ValueError(f"""
The axis argument to Canvas.line must be 0 or 1
    Received: {axis}""")

        glyph.set_line_width(line_width)
        if glyph.antialiased:
            non_cat_agg = agg
            if isinstance(non_cat_agg, rd.by):
                non_cat_agg = non_cat_agg.reduction
            if not isinstance(non_cat_agg, (rd.any, rd.count, rd.max, rd.min, rd.sum, rd.summary, rd._first_or_last,
                                            rd.mean, rd.where)):
                raise NotImplementedError(
                    f"{type(non_cat_agg)} reduction not implemented for antialiased lines")
        return bypixel(source, self, glyph, agg, antialias=glyph.antialiased)

    # ---------------------------------------------------------------------------- areas
    def area(self, source, x, y, agg=None, axis=0, y_stack=None):
        """Compute a reduction by pixel, mapping data to pixels as a filled area region (core.py:480-709)."""
        if agg is None:
            agg = rd.any()
        orig_x, orig_y, orig_y_stack = x, y, y_stack
        x, y, y_stack = _broadcast_column_specifications(x, y, y_stack)
        scalar, seq = (Number, str), (list, tuple)
        if axis == 0:
            if y_stack is None:
                if isinstance(x, scalar) and isinstance(y, scalar):
                    glyph = AreaGlyph(LineAxis0(x, y))
                elif isinstance(x, seq) and isinstance(y, seq):
                    glyph = AreaGlyph(LineAxis0Multi(tuple(x), tuple(y)))
                else:
                    raise ValueError(f"""
Invalid combination of x and y arguments to Canvas.area when axis=0.
    Received:
        x: {repr(x)}
        y: {repr(y)}
See docstring for more information on valid usage""")
            else:
                if isinstance(x, scalar) and isinstance(y, scalar) and isinstance(y_stack, scalar):
                    glyph = AreaGlyph(LineAxis0(x, y), LineAxis0(x, y_stack))
                elif isinstance(x, seq) and isinstance(y, seq) and isinstance(y_stack, seq):
                    glyph = AreaGlyph(LineAxis0Multi(tuple(x), tuple(y)), LineAxis0Multi(tuple(x), tuple(y_stack)))
                else:
                    raise ValueError(f"""
Invalid combination of x, y, and y_stack arguments to Canvas.area when axis=0.
    Received:
        x: {repr(orig_x)}
        y: {repr(orig_y)}
        y_stack: {repr(orig_y_stack)}
See docstring for more information on valid usage""")
        elif axis == 1:
            if y_stack is None:
                if isinstance(x, seq) and isinstance(y, seq):
                    glyph = AreaGlyph(LinesAxis1(tuple(x), tuple(y)))
                elif isinstance(x, np.ndarray) and isinstance(y, seq):
                    glyph = AreaGlyph(LinesAxis1XConstant(x, tuple(y)))
                elif isinstance(x, seq) and isinstance(y, np.ndarray):
                    glyph = AreaGlyph(LinesAxis1YConstant(tuple(x), y))
                elif isinstance(x, scalar) and isinstance(y, scalar):
                    glyph = AreaGlyph(LinesAxis1Ragged(x, y))
                else:
                    raise ValueError(f"""
Invalid combination of x and y arguments to Canvas.area when axis=1.
    Received:
        x: {repr(x)}
        y: {repr(y)}
See docstring for more information on valid usage""")
            else:
                if isinstance(x, seq) and isinstance(y, seq) and isinstance(y_stack, seq):
                    glyph = AreaGlyph(LinesAxis1(tuple(x), tuple(y)), LinesAxis1(tuple(x), tuple(y_stack)))
                elif isinstance(x, np.ndarray) and isinstance(y, seq) and isinstance(y_stack, seq):
                    glyph = AreaGlyph(LinesAxis1XConstant(x, tuple(y)), LinesAxis1XConstant(x, tuple(y_stack)))
                elif isinstance(x, seq) and isinstance(y, np.ndarray) and isinstance(y_stack, np.ndarray):
                    glyph = AreaGlyph(LinesAxis1YConstant(tuple(x), y), LinesAxis1YConstant(tuple(x), y_stack))
                elif isinstance(x, scalar) and isinstance(y, scalar) and isinstance(y_stack, scalar):
                    glyph = AreaGlyph(LinesAxis1Ragged(x, y), LinesAxis1Ragged(x, y_stack))
                else:
                    raise ValueError(f"""
Invalid combination of x, y, and y_stack arguments to Canvas.area when axis=1.
    Received:
        x: {repr(orig_x)}
        y: {repr(orig_y)}
        y_stack: {repr(orig_y_stack)}
See docstring for more information on valid usage""")
        else:
            raise ValueError(f"""
The axis argument to Canvas.area must be 0 or 1
    Received: {axis}""")
        return bypixel(source, self, glyph, agg)

    # ---------------------------------------------------------------------------- validation
    def validate_ranges(self, x_range, y_range):
        self.x_axis.validate(x_range)
        self.y_axis.validate(y_range)

    def validate_size(self, width, height):
        if width <= 0 or height <= 0:
            raise ValueError("Invalid size: plot_width and plot_height must be bigger than 0")

    def validate(self):
        """core.py:1280-1283"""
        self.validate_ranges(self.x_range, self.y_range)
        self.validate_size(self.plot_width, self.plot_height)


def bypixel(source, canvas, glyph, agg, *, antialias=False):
    """Compute an aggregate grouped by pixel sized bins (core.py:1334-1359).  The reference dispatches on
    type(source) to a backend pipeline; here every source ends up as device columns and the glyph
    type selects the fused kernel."""
    from . import pipeline
    from .distributed import current_group
    from .frame import DeviceFrame, HostFrame, _is_arrow
    import pandas as pd

    if not (isinstance(source, (pd.DataFrame, DeviceFrame, HostFrame, dict)) or _is_arrow(source)):
        raise ValueError("source must be a pandas or dask DataFrame")
    dist = current_group(source)
    with warnings.catch_warnings():
        warnings.filterwarnings('ignore', r'All-NaN (slice|axis) encountered')
        if isinstance(glyph, Point):
            return pipeline.points(source, canvas, glyph, agg, dist=dist)
        if isinstance(glyph, AreaGlyph):
            return pipeline.areas(source, canvas, glyph, agg, dist=dist)
        if isinstance(glyph, _LineGlyph):
            return pipeline.lines(source, canvas, glyph, agg, antialias=antialias, dist=dist)
    raise NotImplementedError(f"glyph {type(glyph).__name__} is not supported")
