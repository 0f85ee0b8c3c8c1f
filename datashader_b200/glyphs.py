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


def _to_float_matrix(cols):
    dt = torch.float32 if all(c.dtype == torch.float32 for c in cols) else torch.float64
    return torch.stack([c.to(dt) for c in cols], dim=0).contiguous()


class _LineGlyph(Glyph):
    """Common host side of the line layouts: every layout is presented to the kernel as vertex vectors
    `xs, ys` of shape [nlines, nverts] (or one shared [nverts] vector), plus a dsb_line_layout."""
    antialiased = False
    ragged = False
    _line_width = 0
    x_label = "x"
    y_label = "y"
    value_per_vertex = False      # axis=0 layouts index values by vertex (row), axis=1 by line (row)

    def compute_x_bounds(self, frame, stream_ptr):
        return maybe_expand_bounds(_column_bounds(stream_ptr, self._x_tensors(frame)))

    def compute_y_bounds(self, frame, stream_ptr):
        return maybe_expand_bounds(_column_bounds(stream_ptr, self._y_tensors(frame)))


class LineAxis0(_LineGlyph):
    """One line through all rows: vertices (x[i], y[i]) (glyphs/line.py:59-111)."""
    value_per_vertex = True

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
            raise ValueError('x must be real')
        elif schema[str(self.y)][0] not in ("float", "int"):
            raise ValueError('y must be real')

    def _x_tensors(self, frame):
        return [frame[self.x]]

    def _y_tensors(self, frame):
        return [frame[self.y]]

    def vertices(self, frame):
        xs, ys = _to_float_matrix([frame[self.x]]), _to_float_matrix([frame[self.y]])
        return xs, ys, (xs.shape[1], ys.shape[1])


class LineAxis0Multi(_LineGlyph):
    """One line per (x_k, y_k) column pair, vertices along the rows (glyphs/line.py:114-195)."""
    value_per_vertex = True

    def __init__(self, x, y):
        self.x, self.y = tuple(x), tuple(y)

    def required_columns(self):
        return list(self.x) + list(self.y)

    def validate(self, schema):
        if not {schema[str(c)][0] for c in self.x} <= {"float", "int"}:
            raise ValueError('x columns must be real')
        elif not {schema[str(c)][0] for c in self.y} <= {"float", "int"}:
            raise ValueError('y columns must be real')
        if len(self.x) != len(self.y):
            raise ValueError(f"x and y coordinate lengths do not match: {len(self.x)} != {len(self.y)}")

    def _x_tensors(self, frame):
        return [frame[c] for c in self.x]

    def _y_tensors(self, frame):
        return [frame[c] for c in self.y]

    def vertices(self, frame):
        xs, ys = _to_float_matrix(self._x_tensors(frame)), _to_float_matrix(self._y_tensors(frame))   # [ncols, nrows]
        return xs, ys, (xs.shape[1], ys.shape[1])


class LinesAxis1(_LineGlyph):
    """One line per row; vertex coordinates spread over columns x[0..k), y[0..k)
    (glyphs/line.py:198-306)."""

    def __init__(self, x, y):
        self.x, self.y = tuple(x), tuple(y)

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

    def _x_tensors(self, frame):
        return [frame[c] for c in self.x]

    def _y_tensors(self, frame):
        return [frame[c] for c in self.y]

    def vertices(self, frame):
        # [nverts, nlines] stacked then transposed to the reference's [nlines, nverts] (line.py:298-299)
        xs = _to_float_matrix(self._x_tensors(frame)).t().contiguous()
        ys = _to_float_matrix(self._y_tensors(frame)).t().contiguous()
        return xs, ys, (xs.shape[1], ys.shape[1])


class LinesAxis1XConstant(LinesAxis1):
    """One line per row with a shared x vector (glyphs/line.py:309-381)."""

    def __init__(self, x, y):
        self.x = np.asarray(x)
        self.y = tuple(y)

    def required_columns(self):
        return list(self.y)

    def validate(self, schema):
        if not {schema[str(c)][0] for c in self.y} <= {"float", "int"}:
            raise ValueError('y columns must be real')
        if len(self.x) != len(self.y):
            raise ValueError(f"x and y coordinate lengths do not match: {len(self.x)} != {len(self.y)}")

    def _x_tensors(self, frame):
        return [torch.from_numpy(np.ascontiguousarray(self.x, dtype=np.float64)).to(frame.device)]

    def vertices(self, frame):
        ys = _to_float_matrix(self._y_tensors(frame)).t().contiguous()
        xs = torch.from_numpy(np.ascontiguousarray(self.x)).to(frame.device)
        if xs.dtype != ys.dtype or xs.dtype not in (torch.float32, torch.float64):
            xs, ys = xs.to(torch.float64), ys.to(torch.float64)
        return xs.reshape(1, -1).contiguous(), ys, (0, ys.shape[1])


class LinesAxis1YConstant(LinesAxis1):
    """One line per row with a shared y vector (glyphs/line.py:384-454)."""

    def __init__(self, x, y):
        self.x = tuple(x)
        self.y = np.asarray(y)

    def required_columns(self):
        return list(self.x)

    def validate(self, schema):
        if not {schema[str(c)][0] for c in self.x} <= {"float", "int"}:
            raise ValueError('x columns must be real')
        if len(self.x) != len(self.y):
            raise ValueError(f"x and y coordinate lengths do not match: {len(self.x)} != {len(self.y)}")

    def _y_tensors(self, frame):
        return [torch.from_numpy(np.ascontiguousarray(self.y, dtype=np.float64)).to(frame.device)]

    def vertices(self, frame):
        xs = _to_float_matrix(self._x_tensors(frame)).t().contiguous()
        ys = torch.from_numpy(np.ascontiguousarray(self.y)).to(frame.device)
        if xs.dtype != ys.dtype or ys.dtype not in (torch.float32, torch.float64):
            xs, ys = xs.to(torch.float64), ys.to(torch.float64)
        return xs, ys.reshape(1, -1).contiguous(), (xs.shape[1], 0)


class LinesAxis1Ragged(_LineGlyph):
    """One line per row, the row's vertices held by two RaggedArray columns (glyphs/line.py:457-523; extend :1538-1600).
    The kernels read the flat arrays in place: `vertices` returns them as [1, flat length] and `ragged_starts` the start
    index of every row (dsb_line_layout.x_starts / y_starts)."""
    ragged = True

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
        if schema[str(self.x)][0] != "ragged":
            raise ValueError('x must be a RaggedArray')
        elif schema[str(self.y)][0] != "ragged":
            raise ValueError('y must be a RaggedArray')

    def _x_tensors(self, frame):
        return [frame[self.x].flat]

    def _y_tensors(self, frame):
        return [frame[self.y].flat]

    def vertices(self, frame):
        xs, ys = _to_float_matrix([frame[self.x].flat]), _to_float_matrix([frame[self.y].flat])
        return xs, ys, (0, 0)

    def ragged_starts(self, frame):
        return [frame[self.x].starts, frame[self.y].starts]


class AreaGlyph(Glyph):
    """Filled area between the curve of `line` and y = 0 (stack is None) or the curve of `stack`
    (glyphs/area.py:60-900: AreaToZero* / AreaToLine* for the axis0, axis0-multi, axis1 and constant-x/y layouts).
    The vertex layout is the one of the wrapped line glyph."""

    def __init__(self, line, stack=None):
        self.line, self.stack = line, stack
        self.value_per_vertex = line.value_per_vertex

    @property
    def x_label(self):
        return self.line.x_label

    @property
    def y_label(self):
        return self.line.y_label

    def required_columns(self):
        cols = self.line.required_columns()
        if self.stack is not None:
            cols = cols + [c for c in self.stack.required_columns() if c not in cols]
        return cols

    def validate(self, schema):
        self.line.validate(schema)
        if self.stack is not None:
            self.stack.validate(schema)

    def _x_tensors(self, frame):
        return self.line._x_tensors(frame)

    def _y_tensors(self, frame):
        ts = self.line._y_tensors(frame)
        if self.stack is not None:
            ts = ts + self.stack._y_tensors(frame)
        return ts

    def y_bounds_include_zero(self):
        return self.stack is None      # area.py:71-79

    @property
    def ragged(self):
        return getattr(self.line, "ragged", False)

    def ragged_starts(self, frame):
        st = self.line.ragged_starts(frame)
        if self.stack is not None:
            st = st + [self.stack.ragged_starts(frame)[1]]
        return st

    def vertices(self, frame):
        xs, ys0, (xls, yls) = self.line.vertices(frame)
        ys1 = None
        if self.stack is not None:
            _, ys1, _ = self.stack.vertices(frame)
        ts = [t for t in (xs, ys0, ys1) if t is not None]
        if len({t.dtype for t in ts}) > 1:
            xs, ys0 = xs.to(torch.float64), ys0.to(torch.float64)
            ys1 = ys1.to(torch.float64) if ys1 is not None else None
        return xs, ys0, ys1, (xls, yls)
