"""Reduction objects with the reference's names, arguments and error behaviour
(datashader/reductions.py), re-expressed for the B200 path.

In the reference a reduction contributes numba `_append` functions that compiler.make_append
(compiler.py:321-475) stitches into generated Python.  Here each reduction instead declares the
*accumulators* it needs (`Acc`: commutative per-pixel ops implemented by libdsb200, see
include/dsb200.h) and a finishing pass; datashader_b200/pipeline.py fuses all accumulators of a
call into one kernel launch.  Because every accumulator is commutative, the same objects serve the
single-GPU and the sharded multi-GPU path (partials combine with elementwise sum/max/min), which is
what `uses_row_index(cuda or partitioned)` arranges in the reference (reductions.py:1346-1360).
"""
from __future__ import annotations

from enum import Enum

import numpy as np
import torch

from . import _lib

__all__ = ["count", "any", "sum", "mean", "min", "max", "first", "last", "where", "by", "count_cat", "summary",
           "category_codes", "category_modulo", "category_binning", "SpecialColumn"]


class SpecialColumn(Enum):
    """reductions.py:40-54"""
    RowIndex = 1


class Acc:
    """One accumulator canvas: (kind, value column, nan-check column[, aux accumulator])."""
    __slots__ = ("kind", "col", "chk", "aux")

    def __init__(self, kind, col=None, chk=None, aux=None):
        self.kind, self.col, self.chk, self.aux = kind, col, chk, aux

    @property
    def key(self):
        return (self.kind, self.col, self.chk, self.aux.key if self.aux is not None else None)

    def __repr__(self):
        return f"Acc{self.key}"


# accumulator kind -> (torch dtype of the canvas, NCCL combine op name)
ACC_INFO = {
    "count": (torch.int32, "sum"), "any": (torch.uint8, "max"), "sum": (torch.float64, "sum"),
    "max32": (torch.int32, "max"), "min32": (torch.int32, "min"),
    "max64": (torch.int64, "max"), "min64": (torch.int64, "min"),
    "maxrow": (torch.int64, "max"), "minrow": (torch.int64, "min"),
    "argmax32": (torch.int64, "max"), "argmin32": (torch.int64, "min"),
    "matchrow64": (torch.int64, "min"), "matchrow32": (torch.int64, "min"),
}
ACC_OP = {
    "count": _lib.OP_COUNT, "any": _lib.OP_ANY, "sum": _lib.OP_SUM, "max32": _lib.OP_MAX32, "min32": _lib.OP_MIN32,
    "max64": _lib.OP_MAX64, "min64": _lib.OP_MIN64, "maxrow": _lib.OP_MAXROW, "minrow": _lib.OP_MINROW,
    "argmax32": _lib.OP_ARGMAX32, "argmin32": _lib.OP_ARGMIN32, "matchrow64": _lib.OP_MATCHROW64,
    # initialised like a MINROW canvas; never part of a plan: pipeline._launch_points hands it to dsb_points_match32
    "matchrow32": _lib.OP_MINROW,
}


def _is_key32(np_dtype):
    """Columns whose max / min run on 32-bit order-preserving keys.  32-bit integers are excluded: their keys can
    equal the INT_MIN / INT_MAX "empty" sentinels of the key32 canvases (uint32 0, int32 INT32_MIN / INT32_MAX),
    so they take the 64-bit key path, where every value is far from the sentinels."""
    d = np.dtype(np_dtype)
    return d == np.float32 or (d.kind in "iub" and d.itemsize <= 2)


def _rows_fit_packed32(ctx):
    """The packed {key32, row} accumulators of where(max | min) keep the low 32 bits of the global row id and break ties on
    them (dsb_decode_arg restores the rest from the frame's row offset).  That is the reference's "first row of the extreme
    wins" (reductions.py:2009-2016) only while the rows of this frame - a shard of a bigger one, or a chunk-streamed host source -
    do not cross a multiple of 2^32; a frame that does takes the 64-bit path (extreme, then the first matching row)."""
    cached = getattr(ctx, "_packed32_ok", None)
    if cached is not None:
        return cached
    frame = getattr(ctx, "frame", None)
    if frame is None:
        return True
    lo, n = int(getattr(frame, "row_offset", 0)), len(frame)
    ok = n == 0 or (lo >> 32) == ((lo + n - 1) >> 32)
    if getattr(ctx, "dist", None) is not None:
        # every rank must declare the same accumulators (their canvases are all-reduced pairwise): one flag, agreed once per query
        flag = torch.tensor([0 if ok else 1], dtype=torch.int32, device=frame.device)
        ok = int(ctx.dist._all_reduce(flag, "max").item()) == 0
    ctx._packed32_ok = ok
    return ok


# ---------------------------------------------------------------------------------------------
# category preprocessors (reductions.py:123-260)
class CategoryPreprocess:
    def __init__(self, column):
        self.column = column

    @property
    def cat_column(self):
        return self.column


class category_codes(CategoryPreprocess):
    """Category codes of a categorical column (reductions.py:143-169)."""

    def categories(self, schema):
        return list(schema[self.column][1])

    def validate(self, schema):
        if self.column not in schema:
            raise ValueError("specified column not found")
        if schema[self.column][0] != "categorical":
            raise ValueError("input must be categorical")

    def codes(self, frame):
        return frame[self.column]


class category_modulo(category_codes):
    """(column_value - offset) % modulo for an integer column (reductions.py:171-205)."""

    def __init__(self, column, modulo, offset=0):
        super().__init__(column)
        self.offset = offset
        self.modulo = modulo

    def categories(self, schema):
        return list(range(self.modulo))

    def validate(self, schema):
        if self.column not in schema:
            raise ValueError("specified column not found")
        if schema[self.column][0] != "int":
            raise ValueError("input must be an integer column")

    def codes(self, frame):
        t = frame[self.column]
        return torch.remainder(t.to(torch.int64) - self.offset, self.modulo).to(torch.int32)


class category_binning(category_modulo):
    """Bin a continuous column into nbins (+1 for NaN / clipped) categories (reductions.py:208-260)."""

    def __init__(self, column, lower, upper, nbins, include_under=True, include_over=True):
        super().__init__(column, nbins + 1)
        self.bin0 = lower
        self.binsize = (upper - lower) / float(nbins)
        self.nbins = nbins
        self.bin_under = 0 if include_under else nbins
        self.bin_over = nbins - 1 if include_over else nbins

    def validate(self, schema):
        if self.column not in schema:
            raise ValueError("specified column not found")

    def codes(self, frame):
        v = frame[self.column]
        if not v.dtype.is_floating_point:
            v = v.to(torch.float64)
        nan = torch.isnan(v)
        idx_f = (v - self.bin0) / self.binsize          # same dtype arithmetic as numpy does on the column
        idx_f = torch.where(nan, torch.zeros_like(idx_f), idx_f)
        idx = idx_f.to(torch.int64)                       # truncation, like astype(int)
        idx = torch.where(idx < 0, torch.full_like(idx, self.bin_under), idx)
        idx = torch.where(idx >= self.nbins, torch.full_like(idx, self.bin_over), idx)
        idx = torch.where(nan, torch.full_like(idx, self.nbins), idx)
        return idx.to(torch.int32)


# ---------------------------------------------------------------------------------------------
class Reduction:
    """Base class for per-bin reductions (reductions.py:309-472)."""

    def __init__(self, column=None):
        self.column = column

    # -- reference-facing -------------------------------------------------------------------
    def validate(self, schema):
        if self.column == SpecialColumn.RowIndex:
            return
        if self.column not in schema:
            raise ValueError("specified column not found")
        if schema[self.column][0] not in ("float", "int"):
            raise ValueError("input must be numeric")

    def is_categorical(self):
        return False

    def is_where(self):
        return False

    @property
    def columns_needed(self):
        return [self.column] if isinstance(self.column, str) else []

    def _hashable_inputs(self):
        return (type(self).__name__, self.column)

    def __hash__(self):
        return hash(self._hashable_inputs())

    def __eq__(self, other):
        return type(self) is type(other) and self._hashable_inputs() == other._hashable_inputs()

    def __repr__(self):
        return f"{type(self).__name__}({self.column!r})"

    # -- device-facing ----------------------------------------------------------------------
    def _accs(self, ctx):
        raise NotImplementedError

    def _finalize(self, ctx, canv):
        """-> torch tensor on the device, already in the reference's output dtype/layout."""
        raise NotImplementedError

    def _exact_accs(self, ctx):
        """accumulators of the row-exact redo after DSB_NOTE_NEGZERO (max / min only)"""
        return []

    # antialiased-line support: (dsb_line_agg, needs value column)
    _line_agg = None


class OptionalFieldReduction(Reduction):
    def validate(self, schema):
        if self.column is not None:
            super().validate(schema)

    @property
    def columns_needed(self):
        return [self.column] if self.column is not None else []


class count(OptionalFieldReduction):
    """Count elements in each bin -> uint32 (float32 when antialiased). reductions.py:532-664"""
    _line_agg = _lib.LINE_COUNT

    def __init__(self, column=None, self_intersect=True):
        super().__init__(column)
        self.self_intersect = self_intersect

    def _hashable_inputs(self):
        return super()._hashable_inputs() + (self.self_intersect,)

    def _accs(self, ctx):
        return [Acc("count", self.column)]

    def _finalize(self, ctx, canv):
        return canv[Acc("count", self.column).key]     # int32 bits == uint32 bits; viewed at the host boundary

    _out_np_view = np.uint32


class any(OptionalFieldReduction):   # noqa: A001
    """Whether any element maps to each bin -> bool. reductions.py:825-890"""
    _line_agg = _lib.LINE_ANY

    def _accs(self, ctx):
        return [Acc("any", self.column)]

    def _finalize(self, ctx, canv):
        return canv[Acc("any", self.column).key]

    _out_np_view = np.bool_


class _FloatingReduction(Reduction):
    pass


class sum(_FloatingReduction):   # noqa: A001
    """Sum of `column` (NaN where nothing was added) -> float64. reductions.py:1030-1098.
    Built, like the reference's CUDA path (:1047-1051), from a zero-initialised sum and a "was anything added" mask;
    the mask here is the non-null count of the column, which mean() shares and the privatised count kernel serves."""
    _line_agg = _lib.LINE_SUM

    def __init__(self, column=None, self_intersect=True):
        super().__init__(column)
        self.self_intersect = self_intersect

    def _hashable_inputs(self):
        return super()._hashable_inputs() + (self.self_intersect,)

    def _accs(self, ctx):
        return [Acc("sum", self.column), Acc("count", self.column)]

    def _finalize(self, ctx, canv):
        s, m = canv[Acc("sum", self.column).key], canv[Acc("count", self.column).key]
        out = torch.empty_like(s)
        _lib.check(_lib.lib().dsb_finalize_sum_counted(s.data_ptr(), m.data_ptr(), out.data_ptr(), s.numel(), ctx.stream_ptr),
                   "dsb_finalize_sum_counted")
        return out


class mean(Reduction):
    """Mean of `column` -> float64. reductions.py:1280-1297"""

    _line_agg = _lib.LINE_MEAN      # antialiased lines: _sum_zero / _count_ignore_antialiasing (reductions.py:667-693, 967-975)

    def _accs(self, ctx):
        return [Acc("sum", self.column), Acc("count", self.column)]

    def _finalize(self, ctx, canv):
        s, c = canv[Acc("sum", self.column).key], canv[Acc("count", self.column).key]
        out = torch.empty_like(s)
        _lib.check(_lib.lib().dsb_finalize_mean(s.data_ptr(), c.data_ptr(), out.data_ptr(), s.numel(), ctx.stream_ptr),
                   "dsb_finalize_mean")
        return out


class _MinMax(_FloatingReduction):
    _which = None  # "max" / "min"

    def _acc(self, ctx):
        bits = "32" if _is_key32(ctx.np_dtype(self.column)) else "64"
        return Acc(self._which + bits, self.column)

    def _accs(self, ctx):
        return [self._acc(ctx)]

    def _as_where(self):
        """where(self, self.column): the selected row per pixel - the earliest row among the ties of the extreme, which
        is the row whose value the reference's strict compare keeps (reductions.py:1178-1183, 1222-1227)"""
        w = where.__new__(where)
        w.column, w.selector, w.columns = self.column, self, (self.column, self.column)
        return w

    def _exact_accs(self, ctx):
        return self._as_where()._row_accs(ctx)

    def _finalize(self, ctx, canv):
        if getattr(ctx, "exact_zero", False) and np.dtype(ctx.np_dtype(self.column)).kind == "f":
            return _gather(ctx, self._as_where()._rows(ctx, canv), self.column)
        acc = self._acc(ctx)
        k = canv[acc.key]
        out = torch.empty(k.shape, dtype=torch.float64, device=k.device)
        _lib.check(_lib.lib().dsb_decode_minmax(k.data_ptr(), ACC_OP[acc.kind], ctx.dsb_dtype(self.column),
                                                out.data_ptr(), k.numel(), ctx.stream_ptr), "dsb_decode_minmax")
        return out


class max(_MinMax):   # noqa: A001
    """Maximum of `column` -> float64. reductions.py:1208-1260"""
    _which = "max"
    _line_agg = _lib.LINE_MAX


class min(_MinMax):   # noqa: A001
    """Minimum of `column` -> float64. reductions.py:1160-1205"""
    _which = "min"
    _line_agg = _lib.LINE_MIN


def _gather(ctx, rows, column):
    """value canvas = column[row] per pixel (NaN where empty); cross-shard aware."""
    if ctx.resident is None:
        # host source streamed in chunks: the canvas-sized lookup is done against the host column
        if ctx.dist is not None:
            raise NotImplementedError("sharded + host-streamed sources: stage the shard in a DeviceFrame")
        r = rows.cpu()
        empty = (r < 0) | (r == torch.iinfo(torch.int64).max)
        idx = (r - ctx.frame.row_offset).clamp_(0, len(ctx.frame) - 1 if len(ctx.frame) else 0)
        vals = ctx.frame.columns[column][idx.reshape(-1)].to(torch.float64).reshape(r.shape)
        vals[empty] = float("nan")
        return vals.to(rows.device)
    out = torch.zeros(rows.shape, dtype=torch.float64, device=rows.device)
    col = ctx.resident[column]
    _lib.check(_lib.lib().dsb_gather_rows(rows.data_ptr(), ctx.resident.row_offset, len(ctx.resident), col.data_ptr(),
                                          ctx.dsb_dtype(column), out.data_ptr(), rows.numel(), ctx.stream_ptr),
               "dsb_gather_rows")
    if ctx.dist is not None:
        out = ctx.dist.sum_bits_f64(out, rows)
    return out


def _finish_rows(ctx, rows):
    _lib.check(_lib.lib().dsb_finish_minrow(rows.data_ptr(), rows.numel(), ctx.stream_ptr), "dsb_finish_minrow")
    return rows


class _first_or_last(Reduction):
    """first / last -> float64: min / max of the global row id among non-NaN rows, then a gather.
    This is the reference's own formulation whenever rows are processed in parallel
    (reductions.py:1346-1360: uses_row_index = cuda or partitioned)."""
    _row_kind = None

    def _row_acc(self):
        return Acc(self._row_kind, None, self.column)

    def _accs(self, ctx):
        return [self._row_acc()]

    def _finalize(self, ctx, canv):
        return _gather(ctx, canv[self._row_acc().key], self.column)


class first(_first_or_last):
    """reductions.py:1381-1416"""
    _row_kind = "minrow"


class last(_first_or_last):
    """reductions.py:1419-1454"""
    _row_kind = "maxrow"


class where(_FloatingReduction):
    """Values of `lookup_column` (or the row index when None -> int64, -1 = empty) at the row picked by
    `selector` (first, last, max or min). reductions.py:1842-2166."""

    def __init__(self, selector, lookup_column=None):
        if not isinstance(selector, (first, last, max, min)):
            raise TypeError(
                "selector can only be a first, first_n, last, last_n, "
                "max, max_n, min or min_n reduction")
        if lookup_column is None:
            lookup_column = SpecialColumn.RowIndex
        super().__init__(lookup_column)
        self.selector = selector
        self.columns = (selector.column, lookup_column)

    def _hashable_inputs(self):
        return super()._hashable_inputs() + (self.selector,)

    def is_where(self):
        return True

    @property
    def columns_needed(self):
        return [c for c in self.columns if isinstance(c, str)]

    def validate(self, schema):
        if self.column != SpecialColumn.RowIndex:
            super().validate(schema)
        self.selector.validate(schema)
        if self.column != SpecialColumn.RowIndex and self.column == self.selector.column:
            raise ValueError("where and its contained reduction cannot use the same column")

    def _row_accs(self, ctx):
        """accumulators whose result is (or decodes to) the selected global row per pixel"""
        sel = self.selector
        if isinstance(sel, _first_or_last):
            return [sel._row_acc()]
        which = sel._which
        if _is_key32(ctx.np_dtype(sel.column)):
            if getattr(ctx, "match32", False) and np.dtype(ctx.np_dtype(sel.column)) == np.float32 and len(ctx.shape) == 2:
                # canvases beyond L2 (pipeline.points sets ctx.match32): the packed {key, row} accumulator would be 8 bytes per
                # pixel, L2-banded, with a global RED per hit; two passes instead - the plain extreme (4 bytes per pixel, served
                # by the routed kernels), then "which row holds it" against the finished canvas (dsb_points_match32)
                value = Acc(which + "32", sel.column)
                return [value, Acc("matchrow32", sel.column, None, aux=value)]
            if _rows_fit_packed32(ctx):
                return [Acc("arg" + which + "32", sel.column)]
        value = Acc(which + "64", sel.column)
        return [value, Acc("matchrow64", sel.column, None, aux=value)]

    def _accs(self, ctx):
        return self._row_accs(ctx)

    def _rows(self, ctx, canv):
        accs = self._row_accs(ctx)
        last_acc = accs[-1]
        c = canv[last_acc.key]
        if last_acc.kind.startswith("arg"):
            rows = torch.empty(c.shape, dtype=torch.int64, device=c.device)
            if ctx.dist is not None:
                return ctx.dist.arg_rows(ctx, c, last_acc, self.selector.column, rows)
            _lib.check(_lib.lib().dsb_decode_arg(c.data_ptr(), ACC_OP[last_acc.kind], ctx.dsb_dtype(self.selector.column),
                                                 ctx.frame.row_offset, None, rows.data_ptr(), c.numel(), ctx.stream_ptr),
                       "dsb_decode_arg")
            return rows
        if last_acc.kind in ("minrow", "matchrow64", "matchrow32"):
            return _finish_rows(ctx, c.clone())
        return c

    def _finalize(self, ctx, canv):
        rows = self._rows(ctx, canv)
        if self.column == SpecialColumn.RowIndex:
            return rows
        return _gather(ctx, rows, self.column)

    def __repr__(self):
        return f"where(selector={self.selector!r}, lookup_column={self.column!r})"


class by(Reduction):
    """Apply `reduction` separately per category -> [H, W, ncat]. reductions.py:696-823"""

    def __init__(self, cat_column, reduction=None):
        super().__init__()
        if reduction is None:
            reduction = count()
        if isinstance(cat_column, CategoryPreprocess):
            self.categorizer = cat_column
        elif isinstance(cat_column, str):
            self.categorizer = category_codes(cat_column)
        else:
            raise TypeError("first argument must be a column name or a CategoryPreprocess instance")
        self.column = self.categorizer.column
        self.reduction = reduction

    def _hashable_inputs(self):
        c = self.categorizer
        return (type(self).__name__, type(c).__name__, tuple(sorted((k, repr(v)) for k, v in vars(c).items())),
                self.reduction)

    @property
    def cat_column(self):
        return self.categorizer.column

    @property
    def columns_needed(self):
        return [self.categorizer.column] + self.reduction.columns_needed

    def validate(self, schema):
        self.categorizer.validate(schema)
        self.reduction.validate(schema)

    def is_categorical(self):
        return True

    def is_where(self):
        return self.reduction.is_where()

    def _accs(self, ctx):
        return self.reduction._accs(ctx)

    def _exact_accs(self, ctx):
        return self.reduction._exact_accs(ctx)

    def _finalize(self, ctx, canv):
        return self.reduction._finalize(ctx, canv)

    @property
    def _out_np_view(self):
        return getattr(self.reduction, "_out_np_view", None)

    def __repr__(self):
        return f"{type(self).__name__}(column={self.column!r}, reduction={self.reduction!r})"


class count_cat(by):
    """Alias for by(column, count()). reductions.py:1263-1278"""

    def __init__(self, column):
        super().__init__(column, count())

    def __repr__(self):
        return f"count_cat(column={self.column!r})"


class summary:
    """A collection of named reductions -> Dataset. reductions.py:2169-2246"""

    def __init__(self, **kwargs):
        ks, vs = zip(*sorted(kwargs.items()))
        self.keys = ks
        self.values = vs

    def validate(self, schema):
        for v in self.values:
            v.validate(schema)

    @property
    def columns_needed(self):
        out = []
        for v in self.values:
            out += v.columns_needed
        return out

    def is_categorical(self):
        for v in self.values:
            if v.is_categorical():
                return True
        return False


