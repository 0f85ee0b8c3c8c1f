"""Container-only import shim: a dumb DataArray/Dataset holder (no arithmetic, never copies data)
so the reference's finalize() has something to return when the real xarray is absent."""
import numpy as np


class DataArray:
    __slots__ = ("data", "coords", "dims", "attrs", "name")

    def __init__(self, data=None, coords=None, dims=None, attrs=None, name=None):
        self.data = data
        self.coords = dict(coords) if coords is not None else {}
        self.dims = tuple(dims) if dims is not None else ()
        self.attrs = dict(attrs) if attrs is not None else {}
        self.name = name

    @property
    def values(self):
        return np.asarray(self.data)

    @property
    def shape(self):
        return self.data.shape

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def ndim(self):
        return self.data.ndim

    def __getitem__(self, key):
        return self.coords[key] if isinstance(key, str) else self.data[key]

    @property
    def indexes(self):
        return {d: list(self.coords[d]) for d in self.dims if d in self.coords}

    def transpose(self, *dims):
        order = [self.dims.index(d) for d in dims]
        return type(self)(np.transpose(self.data, order), coords=self.coords, dims=dims, attrs=self.attrs,
                          name=self.name)

    def copy(self):
        return type(self)(self.data.copy(), coords=self.coords, dims=self.dims, attrs=self.attrs, name=self.name)


class Dataset(dict):
    def __init__(self, data_vars=None, coords=None, attrs=None):
        super().__init__(data_vars or {})
        self.attrs = dict(attrs) if attrs is not None else {}


def align(*objs, **kw):
    return objs
