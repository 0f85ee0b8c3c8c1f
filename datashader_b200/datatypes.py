"""RaggedArray: a pandas column whose rows are variable-length numeric arrays - the input type of the ragged line and
area glyphs (Canvas.line(..., axis=1) / Canvas.area(..., axis=1) with scalar column names).

Same storage contract as the reference's datashader/datatypes.py (RaggedArray :209-300, RaggedDtype :91-170): one
`flat_array` holding every row's values back to back and `start_indices[i]` = the first element of row i (empty and missing
rows have start_indices[i] == start_indices[i + 1]).  Only the container is provided; the frame (frame.py RaggedColumn)
reads `flat_array` / `start_indices` by attribute, so the reference's own RaggedArray is accepted as well.
"""
from __future__ import annotations

import re

import numpy as np
from pandas.api.extensions import ExtensionArray, ExtensionDtype, register_extension_dtype


def _index_dtype(n):
    for dt in (np.uint8, np.uint16, np.uint32):
        if n <= np.iinfo(dt).max:
            return dt
    return np.uint64


@register_extension_dtype
class RaggedDtype(ExtensionDtype):
    """'ragged[float64]' and friends (datatypes.py:91-170)."""
    type = np.ndarray
    base = np.dtype('O')
    na_value = np.nan
    _metadata = ('_dtype',)
    _pattern = re.compile(r'^ragged\[(?P<subtype>\w+)\]$')

    def __init__(self, dtype=np.float64):
        if isinstance(dtype, RaggedDtype):
            dtype = dtype.subtype
        self._dtype = np.dtype(dtype)

    @property
    def name(self):
        return f'ragged[{self._dtype.name}]'

    @property
    def subtype(self):
        return self._dtype

    def __repr__(self):
        return self.name

    @classmethod
    def construct_array_type(cls):
        return RaggedArray

    @classmethod
    def construct_from_string(cls, string):
        if not isinstance(string, str):
            raise TypeError(f"'construct_from_string' expects a string, got {type(string)}")
        if string == 'ragged':
            return cls()
        m = cls._pattern.match(string)
        if m is None:
            raise TypeError(f"Cannot construct a 'RaggedDtype' from '{string}'")
        return cls(m.group('subtype'))


class RaggedArray(ExtensionArray):
    """RaggedArray(rows, dtype=None): rows = a list of 1-D array-likes (None / NaN = a missing row), another RaggedArray,
    or {'start_indices': ..., 'flat_array': ...} to wrap existing buffers without copying."""

    def __init__(self, data, dtype=None, copy=False):
        if isinstance(data, dict) or (hasattr(data, 'flat_array') and hasattr(data, 'start_indices')):
            get = data.__getitem__ if isinstance(data, dict) else (lambda k: getattr(data, k))
            starts, flat = np.asarray(get('start_indices')), np.asarray(get('flat_array'))
            if starts.ndim != 1 or flat.ndim != 1:
                raise ValueError('start_indices and flat_array must be 1-D')
            if starts.dtype.kind not in 'iu':
                raise ValueError('start_indices must be an integer array')
            if len(starts) and (np.any(np.diff(starts.astype(np.int64)) < 0) or int(starts[0]) < 0 or int(starts[-1]) > len(flat)):
                raise ValueError('start_indices must be non-decreasing and inside flat_array')
            if dtype is not None:
                flat = flat.astype(RaggedDtype(dtype).subtype if not isinstance(dtype, np.dtype) else dtype, copy=False)
            self._starts = starts.copy() if copy else starts
            self._flat = flat.copy() if copy else flat
        else:
            rows = [None if _missing(r) else np.asarray(r) for r in data]
            if dtype is None:
                present = [r.dtype for r in rows if r is not None and r.size]
                sub = np.result_type(*present) if present else np.dtype(np.float64)
            else:
                sub = RaggedDtype(dtype).subtype if not isinstance(dtype, np.dtype) else dtype
            lens = np.array([0 if r is None else len(r) for r in rows], dtype=np.int64)
            total = int(lens.sum())
            self._starts = (np.cumsum(lens) - lens).astype(_index_dtype(total))
            self._flat = (np.concatenate([r.astype(sub, copy=False) for r in rows if r is not None and len(r)])
                          if total else np.empty(0, dtype=sub))
        self._dtype = RaggedDtype(self._flat.dtype)

    # ---- the storage contract the glyphs read
    @property
    def flat_array(self):
        return self._flat

    @property
    def start_indices(self):
        return self._starts

    # ---- ExtensionArray
    @property
    def dtype(self):
        return self._dtype

    @property
    def nbytes(self):
        return self._flat.nbytes + self._starts.nbytes

    def __len__(self):
        return len(self._starts)

    def _bounds(self, i):
        lo = int(self._starts[i])
        hi = int(self._starts[i + 1]) if i + 1 < len(self._starts) else len(self._flat)
        return lo, hi

    def __getitem__(self, item):
        if isinstance(item, (int, np.integer)):
            n = len(self)
            if item < -n or item >= n:
                raise IndexError(item)
            lo, hi = self._bounds(item % n if n else 0)
            return self._flat[lo:hi] if hi > lo else np.nan
        if isinstance(item, slice):
            return self.take(np.arange(*item.indices(len(self))))
        item = np.asarray(item)
        if item.dtype == bool:
            if len(item) != len(self):
                raise IndexError('boolean mask has the wrong length')
            item = np.nonzero(item)[0]
        return self.take(item)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    @classmethod
    def _from_sequence(cls, scalars, *, dtype=None, copy=False):
        return cls(scalars, dtype=dtype)

    @classmethod
    def _from_factorized(cls, values, original):
        return cls(values, dtype=original.flat_array.dtype)

    def isna(self):
        stops = np.append(self._starts[1:], len(self._flat)).astype(np.int64)
        return stops == self._starts.astype(np.int64)

    def take(self, indices, allow_fill=False, fill_value=None):
        indices = np.asarray(indices, dtype=np.int64)
        n = len(self)
        if allow_fill:
            if np.any(indices < -1):
                raise ValueError('negative indices other than -1 are not allowed with allow_fill')
            rows = [None if i < 0 else self[int(i)] for i in indices]
        else:
            if len(indices) and (indices.min() < -n or indices.max() >= n):
                raise IndexError('out of bounds')
            rows = [self[int(i)] for i in indices]
        return RaggedArray(rows, dtype=self._flat.dtype)

    def copy(self):
        return RaggedArray({'start_indices': self._starts, 'flat_array': self._flat}, copy=True)

    @classmethod
    def _concat_same_type(cls, to_concat):
        to_concat = list(to_concat)
        flat = np.concatenate([a.flat_array for a in to_concat]) if to_concat else np.empty(0)
        offs = np.cumsum([0] + [len(a.flat_array) for a in to_concat[:-1]]) if to_concat else []
        starts = np.concatenate([a.start_indices.astype(np.int64) + o for a, o in zip(to_concat, offs)]) if to_concat else np.empty(0, np.int64)
        return cls({'start_indices': starts.astype(_index_dtype(len(flat))), 'flat_array': flat})

    def __eq__(self, other):
        if isinstance(other, RaggedArray):
            if len(other) != len(self):
                raise ValueError('lengths must match')
            return np.array([_rows_equal(a, b) for a, b in zip(self, other)], dtype=bool)
        other = np.asarray(other)
        return np.array([_rows_equal(a, other) for a in self], dtype=bool)

    def _formatter(self, boxed=False):
        return repr

    def _values_for_factorize(self):
        return np.array([None if _missing(r) else tuple(r) for r in self], dtype=object), None


def _missing(r):
    return r is None or (np.isscalar(r) and r != r) or (isinstance(r, float) and r != r)


def _rows_equal(a, b):
    if _missing(a) or _missing(b):
        return False
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.array_equal(a, b, equal_nan=True))
