"""xarray when it is installed; otherwise a minimal DataArray / Dataset holder with the attributes the
datashader API promises (data, coords, dims, attrs, name).  The reference returns xarray objects
(compiler.py:510-536); this image has no xarray, so the stand-in keeps the same field names."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    import xarray as _xr
    DataArray = _xr.DataArray
    Dataset = _xr.Dataset
    HAVE_XARRAY = True
except Exception:  # noqa: BLE001
    HAVE_XARRAY = False

    class DataArray:
        """Just enough of xarray.DataArray: never copies `data`."""

        def __init__(self, data=None, coords=None, dims=None, attrs=None, name=None):
            self.data = data
            self.dims = tuple(dims) if dims is not None else tuple(f"dim_{i}" for i in range(np.ndim(data)))
            self.coords = {k: np.asarray(v) for k, v in (coords or {}).items()}
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
            if isinstance(key, str):
                return self.coords[key]
            return self.data[key]

        def __array__(self, dtype=None, copy=None):
            a = np.asarray(self.data)
            return a.astype(dtype) if dtype is not None else a

        def __repr__(self):
            return (f"<datashader_b200.DataArray {self.name or ''} dims={self.dims} shape={self.shape} "
                    f"dtype={self.dtype}>\n{self.data!r}")

    class Dataset(dict):
        def __init__(self, data_vars=None, coords=None, attrs=None):
            super().__init__(data_vars or {})
            self.attrs = dict(attrs) if attrs is not None else {}

        @property
        def data_vars(self):
            return self
