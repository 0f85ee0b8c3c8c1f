"""Glyph descriptions (host side only): which columns form the geometry and how bounds are found.
The rasterisation itself lives in csrc/points.cu and csrc/lines.cu."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def maybe_expand_bounds(bounds):
    """glyphs/glyph.py:42-49"""
    minval, maxval = bounds
    if not (np.isfinite(minval) and np.isfinite(maxval)):
        minval, maxval = -1.0, 1.0
    elif minval == maxval:
        minval, maxval = minval - 1, minval + 1
    return minval, maxval


def _column_bounds(ctx_stream_ptr, tensors):
    """NaN-skipping (min, max) over one or more device columns: Glyph._compute_bounds_numba
    (glyphs/glyph.py:66-78) as a device reduction (dsb_bounds)."""
    lo, hi = np.inf, -np.inf
    for t in tensors:
        out = torch.empty(2, dtype=torch.float64, device=t.device)
        dt = _lib.dsb_dtype(str(t.dtype).replace("torch.", ""))
        _lib.check(_lib.lib().dsb_bounds(t.data_ptr(), dt, t.numel(), out.data_ptr(), ctx_stream_ptr), "dsb_bounds")
        mn, mx = out.tolist()
        lo, hi = (mn if mn < lo else lo), (mx if mx > hi else hi)
    return lo, hi


class Glyph:
    antialiased = False

    def set_line_width(self, line_width):
        self._line_width = line_width
        if hasattr(self, "antialiased"):
            self.antialiased = line_width > 0


class Point(Glyph):
    """A point at (x, y); each record maps to one bin, points on the upper bounds fold into the last
    bin (glyphs/points.py:170-242)."""

    def __init__(self, x, y):
        self.x, self.y = x, y

    @property
    def x_label(self):
        return self.x

    @property
    def y_label(self):
        return self.y

    def required_columns(self):
        return [self.x, self.y]

    def validate(self, schema):
        if schema[str(self.x)][0] not in ("float", "int"):
            raise ValueError("x must be real")
        elif schema[str(self.y)][0] not in ("float", "int"):
            raise ValueError("y must be real")

    def compute_x_bounds(self, frame, stream_ptr):
        return maybe_expand_bounds(_column_bounds(stream_ptr, [frame[self.x]]))

    def compute_y_bounds(self, frame, stream_ptr):
        return maybe_expand_bounds(_column_bounds(stream_ptr, [frame[self.y]]))


class LinesAxis1(Glyph):
    """One line per row; vertex coordinates spread over columns x[0..k), y[0..k)
    (glyphs/line.py:198-306)."""
    antialiased = False
    _line_width = 0

    def __init__(self, x, y):
        self.x, self.y = tuple(x), tuple(y)

    x_label = "x"
    y_label = "y"

    def required_columns(self):
        return list(self.x) + list(self.y)

    def validate(self, schema):
        xk = {schema[str(c)][0] for c in self.x}
        yk = {schema[str(c)][0] for c in self.y}
        if not xk <= {"float", "int"}:
            raise ValueError("x columns must be real")
        elif not yk <= {"float", "int"}:
            raise ValueError("y columns must be real")
        if len(self.x) != len(self.y):
            raise ValueError(f"x and y coordinate lengths do not match: {len(self.x)} != {len(self.y)}")

    def compute_x_bounds(self, frame, stream_ptr):
        return maybe_expand_bounds(_column_bounds(stream_ptr, [frame[c] for c in self.x]))

    def compute_y_bounds(self, frame, stream_ptr):
        return maybe_expand_bounds(_column_bounds(stream_ptr, [frame[c] for c in self.y]))
