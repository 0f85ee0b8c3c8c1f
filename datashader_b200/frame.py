"""Column staging at the boundary: pandas (host) columns -> device columns.

The reference borrows `df[col].values` (glyphs/points.py:234-235) and, for cudf, device columns
(reductions.py:97-108).  `DeviceFrame` is the resident-on-GPU equivalent of a cudf.DataFrame for this
path: a dict of 1-D CUDA tensors plus the categorical metadata the reductions need, and the global
row offset of the shard (data_libraries/dask.py:102-117)."""
from __future__ import annotations

import numpy as np
import torch

_TORCH_OK = {np.dtype(t) for t in ("float32", "float64", "int8", "uint8", "int16", "int32", "int64", "bool")}


def _to_tensor(arr: np.ndarray, device, pin=False, stream=None):
    arr = np.ascontiguousarray(arr)
    if not arr.flags.writeable:
        arr = arr.copy()           # torch refuses to alias read-only memory (pandas copy-on-write views)
    if arr.dtype not in _TORCH_OK:
        if arr.dtype.kind in "iu":      # uint16/32/64: widen (values are only read as numbers)
            arr = arr.astype(np.int64 if arr.dtype.itemsize < 8 else np.float64)
        elif arr.dtype.kind == "f":
            arr = arr.astype(np.float32 if arr.dtype.itemsize < 4 else np.float64)
        else:
            raise TypeError(f"unsupported column dtype {arr.dtype}")
    if arr.dtype == np.bool_:
        arr = arr.view(np.uint8)
    t = torch.from_numpy(arr)
    if pin:
        t = t.pin_memory()
    return t.to(device, non_blocking=True)


class DeviceFrame:
    """Device-resident columns for Canvas.points / Canvas.line.

    columns:     name -> 1-D CUDA tensor
    categories:  name -> list of category labels (the column tensor then holds integer codes)
    row_offset:  global id of row 0 (multi-GPU shards; reductions.py:87-113)
    """

    def __init__(self, columns, categories=None, row_offset=0, n_global=None):
        self.columns = dict(columns)
        self.categories = dict(categories or {})
        self.row_offset = int(row_offset)
        lens = {int(t.shape[0]) for t in self.columns.values()}
        if len(lens) > 1:
            raise ValueError("all columns must have the same length")
        self._len = lens.pop() if lens else 0
        self.n_global = n_global

    def __len__(self):
        return self._len

    def __contains__(self, name):
        return name in self.columns

    def __getitem__(self, name):
        return self.columns[name]

    @property
    def device(self):
        for t in self.columns.values():
            return t.device
        return torch.device("cuda")

    @classmethod
    def from_pandas(cls, df, columns=None, device=None, row_offset=0, pin=False):
        import pandas as pd
        device = torch.device(device if device is not None else "cuda")
        cols, cats = {}, {}
        for name in (columns if columns is not None else list(df.columns)):
            if name not in df.columns:
                raise ValueError("specified column not found")     # reductions.py:352-353
            s = df[name]
            if isinstance(s.dtype, pd.CategoricalDtype):
                cats[name] = list(s.cat.categories)
                cols[name] = _to_tensor(np.asarray(s.cat.codes.values), device, pin)
            else:
                cols[name] = _to_tensor(s.to_numpy(), device, pin)
        return cls(cols, cats, row_offset)

    def schema(self):
        out = {}
        for name, t in self.columns.items():
            if name in self.categories:
                out[name] = ("categorical", list(self.categories[name]))
            else:
                out[name] = ("float" if t.dtype.is_floating_point else "int", None)
        return out

    def np_dtype(self, name):
        return np.dtype(str(self.columns[name].dtype).replace("torch.", ""))


def as_device_frame(source, needed, device=None):
    """pandas.DataFrame | DeviceFrame | dict of tensors/arrays -> DeviceFrame with the needed columns."""
    import pandas as pd
    if isinstance(source, DeviceFrame):
        return source
    if isinstance(source, pd.DataFrame):
        # only the needed columns are staged, like _bypixel_sanitise (core.py:1384-1392)
        off = getattr(source, "_datashader_row_offset", 0)
        return DeviceFrame.from_pandas(source, columns=needed, device=device, row_offset=off)
    raise ValueError("source must be a pandas DataFrame or a datashader_b200.DeviceFrame")
