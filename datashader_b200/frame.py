"""Column staging at the boundary: pandas (host) columns -> device columns.

The reference borrows `df[col].values` (glyphs/points.py:234-235) and, for cudf, device columns
(reductions.py:97-108).  `DeviceFrame` is the resident-on-GPU equivalent of a cudf.DataFrame for this
path: a dict of 1-D CUDA tensors plus the categorical metadata the reductions need, and the global
row offset of the shard (data_libraries/dask.py:102-117)."""
from __future__ import annotations

import numpy as np
import torch

_TORCH_OK = {np.dtype(t) for t in ("float32", "float64", "int8", "uint8", "int16", "int32", "int64", "bool")}


def _from_numpy(a):
    """torch view of a numpy array without copying it: pandas hands out read-only (copy-on-write) views, which torch
    only warns about - the columns are never written to."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        return torch.from_numpy(a)


class _PinnedRing:
    """Pageable host memory -> device through a small ring of pinned staging buffers.

    A pageable-source cudaMemcpy is staged by the driver through one internal buffer, synchronously (measured here:
    3.6 GB/s for a pandas frame against 54 GB/s from pinned memory).  The ring does the staging explicitly: a
    multi-threaded CPU copy into a pinned slot, then an asynchronous H2D copy of that slot on the copy stream, so the
    CPU copy of sub-chunk k+1 overlaps the DMA of sub-chunk k (and the kernels of the previous chunk)."""
    SLOT_BYTES = 16 << 20
    NSLOTS = 4
    _rings = {}

    def __init__(self):
        self.slots = [torch.empty(self.SLOT_BYTES, dtype=torch.uint8).pin_memory() for _ in range(self.NSLOTS)]
        self.events = [None] * self.NSLOTS
        self.next = 0

    @classmethod
    def get(cls, device):
        key = str(device)
        if key not in cls._rings:
            cls._rings[key] = cls()
        return cls._rings[key]

    def copy(self, dst, src, stream):
        """dst: 1-D CUDA tensor, src: 1-D contiguous CPU tensor of the same dtype and length; enqueued on `stream`."""
        d8, s8 = dst.view(torch.uint8), src.view(torch.uint8)
        n = s8.numel()
        for off in range(0, n, self.SLOT_BYTES):
            m = min(self.SLOT_BYTES, n - off)
            k = self.next
            self.next = (k + 1) % self.NSLOTS
            if self.events[k] is not None:
                self.events[k].synchronize()              # the DMA that last read this slot has finished
            self.slots[k][:m].copy_(s8[off:off + m])      # pageable -> pinned, torch's parallel CPU copy
            with torch.cuda.stream(stream):
                d8[off:off + m].copy_(self.slots[k][:m], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
            self.events[k] = ev


def _h2d(dst, src, stream):
    """Asynchronous host -> device copy on `stream`: direct DMA from pinned memory, the staging ring otherwise."""
    if src.is_pinned() or src.numel() * src.element_size() < (1 << 20):
        with torch.cuda.stream(stream):
            dst.copy_(src, non_blocking=True)
    else:
        _PinnedRing.get(dst.device).copy(dst, src, stream)


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


class RaggedColumn:
    """A column of variable-length rows: `flat` holds every row's values back to back, `starts[i]` (int64) is the first
    element of row i - the storage of a RaggedArray (datatypes.py: flat_array / start_indices), which is what the ragged
    line and area glyphs read (line.py:1545-1552, area.py:1945-1952).  Host tensors inside a HostFrame, CUDA tensors inside
    a DeviceFrame."""

    def __init__(self, flat, starts):
        self.flat, self.starts = flat.contiguous(), starts.to(torch.int64).contiguous()

    @classmethod
    def from_array(cls, arr):
        """anything with flat_array / start_indices (this package's RaggedArray or the reference's)."""
        flat = np.ascontiguousarray(arr.flat_array)
        starts = np.ascontiguousarray(arr.start_indices).astype(np.int64)
        if flat.dtype not in _TORCH_OK:
            flat = flat.astype(np.float64)
        if len(starts) and (np.any(np.diff(starts) < 0) or starts[0] < 0 or starts[-1] > len(flat)):
            raise ValueError("start_indices must be non-decreasing and inside flat_array")
        return cls(_from_numpy(flat), _from_numpy(starts))

    @property
    def shape(self):
        return (int(self.starts.shape[0]),)

    @property
    def dtype(self):
        return self.flat.dtype

    @property
    def is_cuda(self):
        return self.flat.is_cuda

    @property
    def device(self):
        return self.flat.device

    def contiguous(self):
        return self


def _is_ragged_array(a):
    return hasattr(a, "flat_array") and hasattr(a, "start_indices")


class DeviceFrame:
    """Device-resident columns for Canvas.points / Canvas.line.

    columns:     name -> 1-D CUDA tensor
    categories:  name -> list of category labels (the column tensor then holds integer codes)
    row_offset:  global id of row 0 (multi-GPU shards; reductions.py:87-113)
    """

    def __init__(self, columns, categories=None, row_offset=0, n_global=None):
        self.columns = {}
        for name, t in dict(columns).items():
            if _is_ragged_array(t):
                t = RaggedColumn.from_array(t)
                t = RaggedColumn(t.flat.cuda(), t.starts.cuda())
            if not isinstance(t, (torch.Tensor, RaggedColumn)):
                # cupy / numba / cudf-column style device arrays (__cuda_array_interface__, DLPack): borrowed, not copied
                t = torch.as_tensor(t, device="cuda")
            if not t.is_cuda:
                raise ValueError("DeviceFrame columns must live on the GPU; use HostFrame for host arrays")
            self.columns[name] = t.contiguous()
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
            elif _is_ragged_array(s.array):
                r = RaggedColumn.from_array(s.array)
                cols[name] = RaggedColumn(r.flat.to(device), r.starts.to(device))
            else:
                cols[name] = _to_tensor(s.to_numpy(), device, pin)
        return cls(cols, cats, row_offset)

    def schema(self):
        out = {}
        for name, t in self.columns.items():
            if name in self.categories:
                out[name] = ("categorical", list(self.categories[name]))
            elif isinstance(t, RaggedColumn):
                out[name] = ("ragged", None)
            else:
                out[name] = ("float" if t.dtype.is_floating_point else "int", None)
        return out

    def np_dtype(self, name):
        return np.dtype(str(self.columns[name].dtype).replace("torch.", ""))

    CHUNK_ROWS = 1 << 31     # rows per slice: the ABI takes n <= 2^32 per call, the routed / privatised-any forms a little less - at 2^31
                             # every specialised kernel stays eligible; a multiple of 4, so the vector loads stay aligned

    def n_chunks(self):
        # one call while the frame fits one (n <= 2^32 - 2: every kernel form takes it - slicing config 5's 4e9 rows would route the
        # head of first / last once per slice), slices of CHUNK_ROWS beyond
        return 1 if self._len <= 2 * self.CHUNK_ROWS - 2 else -(-self._len // self.CHUNK_ROWS)

    def resident(self, needed):
        return self

    def chunks(self, needed):
        """Point glyphs walk a resident frame of more than 2^32 rows in slices of the same device columns (no copy), each with
        its own global row offset - like the row chunks of a HostFrame, minus the transfer."""
        if self.n_chunks() == 1:
            yield self
            return
        for lo in range(0, self._len, self.CHUNK_ROWS):
            hi = min(lo + self.CHUNK_ROWS, self._len)
            yield DeviceFrame({c: self.columns[c][lo:hi] for c in needed}, self.categories, self.row_offset + lo)


class HostFrame:
    """Host-resident columns (numpy arrays or CPU torch tensors, ideally pinned) that are streamed to
    the GPU in row chunks while the previous chunk is being aggregated.  This is the ingestion side
    of the boundary: what `df[col].values` (glyphs/points.py:234-235) is for the reference.

    A pandas.DataFrame passed to Canvas.points / Canvas.line is wrapped in a HostFrame without
    copying (only the needed columns, like _bypixel_sanitise, core.py:1384-1392)."""

    CHUNK_ROWS = 1 << 25     # 32 Mi rows: 128 MiB per f32 column per staging buffer

    def __init__(self, columns, categories=None, row_offset=0, device=None):
        self.columns = {}
        for name, a in columns.items():
            if _is_ragged_array(a):
                a = RaggedColumn.from_array(a)
            if isinstance(a, RaggedColumn):
                if a.is_cuda:
                    raise ValueError("HostFrame columns must live on the host; use DeviceFrame")
                self.columns[name] = a
            elif isinstance(a, torch.Tensor):
                if a.is_cuda:
                    raise ValueError("HostFrame columns must live on the host; use DeviceFrame")
                self.columns[name] = a.contiguous()
            else:
                a = np.ascontiguousarray(a)
                if a.dtype == np.bool_:
                    a = a.view(np.uint8)
                if a.dtype not in _TORCH_OK:
                    a = a.astype(np.int64 if a.dtype.kind in "iu" and a.dtype.itemsize < 8 else np.float64)
                self.columns[name] = _from_numpy(a)
        self.categories = dict(categories or {})
        self.row_offset = int(row_offset)
        lens = {int(t.shape[0]) for t in self.columns.values()}
        if len(lens) > 1:
            raise ValueError("all columns must have the same length")
        self._len = lens.pop() if lens else 0
        self.device = torch.device(device if device is not None else "cuda")

    def __len__(self):
        return self._len

    def __contains__(self, name):
        return name in self.columns

    @classmethod
    def from_pandas(cls, df, columns=None, device=None):
        import pandas as pd
        cols, cats = {}, {}
        for name in (columns if columns is not None else list(df.columns)):
            if name not in df.columns:
                raise ValueError("specified column not found")     # reductions.py:352-353
            s = df[name]
            if isinstance(s.dtype, pd.CategoricalDtype):
                cats[name] = list(s.cat.categories)
                cols[name] = np.asarray(s.cat.codes.values)
            elif _is_ragged_array(s.array):
                cols[name] = RaggedColumn.from_array(s.array)
            else:
                cols[name] = s.to_numpy()
        off = getattr(df, "_datashader_row_offset", 0)
        return cls(cols, cats, off, device)

    @classmethod
    def from_arrow(cls, table, columns=None, device=None, row_offset=0):
        """pyarrow.Table / RecordBatch -> HostFrame.  Numeric columns are taken zero-copy when they are a single chunk
        without nulls; float nulls become NaN (what pandas hands the reference), integer columns with nulls are
        widened to float64 with NaN, dictionary columns become category codes + labels (null -> -1, pandas' code)."""
        import pyarrow as pa
        if isinstance(table, pa.RecordBatch):
            table = pa.Table.from_batches([table])
        cols, cats = {}, {}
        for name in (columns if columns is not None else table.column_names):
            if name not in table.column_names:
                raise ValueError("specified column not found")     # reductions.py:352-353
            col = table.column(name)
            arr = col.chunk(0) if col.num_chunks == 1 else col.combine_chunks()
            if isinstance(arr, pa.ChunkedArray):
                arr = arr.combine_chunks()
            if pa.types.is_dictionary(arr.type):
                cats[name] = arr.dictionary.to_pylist()
                idx = arr.indices
                codes = idx.to_numpy(zero_copy_only=False)
                if idx.null_count:
                    codes = np.where(np.asarray(idx.is_null()), -1, np.nan_to_num(codes, nan=0)).astype(np.int64)
                small = np.int8 if len(cats[name]) <= 127 else (np.int16 if len(cats[name]) <= 32767 else np.int32)
                cols[name] = np.asarray(codes).astype(small)
            elif pa.types.is_floating(arr.type) or pa.types.is_integer(arr.type) or pa.types.is_boolean(arr.type):
                cols[name] = arr.to_numpy(zero_copy_only=False)     # nulls: float NaN (ints are promoted to float64)
            else:
                raise ValueError(f"input '{name}' must be a numeric or dictionary column")
        return cls(cols, cats, row_offset, device)

    @classmethod
    def from_parquet(cls, path, columns=None, device=None, row_offset=0):
        """Read only the needed columns of a Parquet file (the reference's performance guidance,
        examples/user_guide/10_Performance.ipynb) straight into a HostFrame."""
        import pyarrow.parquet as pq
        return cls.from_arrow(pq.read_table(path, columns=list(columns) if columns is not None else None),
                              columns=columns, device=device, row_offset=row_offset)

    def schema(self):
        out = {}
        for name, t in self.columns.items():
            if name in self.categories:
                out[name] = ("categorical", list(self.categories[name]))
            elif isinstance(t, RaggedColumn):
                out[name] = ("ragged", None)
            else:
                out[name] = ("float" if t.dtype.is_floating_point else "int", None)
        return out

    def np_dtype(self, name):
        return np.dtype(str(self.columns[name].dtype).replace("torch.", ""))

    def n_chunks(self):
        if any(isinstance(t, RaggedColumn) for t in self.columns.values()):
            return 1          # ragged rows are staged whole (lines and areas take the resident frame)
        return max(1, -(-self._len // self.CHUNK_ROWS))

    def resident(self, needed):
        """The whole frame on the device (single-chunk sources and gathers)."""
        stream = torch.cuda.current_stream(self.device)
        cols = {}
        def up(src):
            dst = torch.empty(src.shape, dtype=src.dtype, device=self.device)
            _h2d(dst, src, stream)
            return dst
        for c in needed:
            src = self.columns[c]
            cols[c] = RaggedColumn(up(src.flat), up(src.starts)) if isinstance(src, RaggedColumn) else up(src)
        return DeviceFrame(cols, self.categories, self.row_offset)

    def chunks(self, needed):
        """Yield DeviceFrame row chunks; H2D copies run on a side stream, double-buffered, so the copy of
        chunk k+1 overlaps the aggregation of chunk k."""
        if self.n_chunks() == 1:
            yield self.resident(needed)
            return
        dev = self.device
        compute = torch.cuda.current_stream(dev)
        copy = torch.cuda.Stream(dev)
        rows = self.CHUNK_ROWS
        bufs = [{c: torch.empty(rows, dtype=self.columns[c].dtype, device=dev) for c in needed} for _ in range(2)]
        free_ev = [None, None]
        for k, lo in enumerate(range(0, self._len, rows)):
            hi = min(lo + rows, self._len)
            b = bufs[k & 1]
            if free_ev[k & 1] is not None:
                copy.wait_event(free_ev[k & 1])           # the kernel that read this buffer has finished
            for c in needed:
                _h2d(b[c][:hi - lo], self.columns[c][lo:hi], copy)
            with torch.cuda.stream(copy):
                ready = torch.cuda.Event()
                ready.record(copy)
            compute.wait_event(ready)
            yield DeviceFrame({c: b[c][:hi - lo] for c in needed}, self.categories, self.row_offset + lo)
            ev = torch.cuda.Event()
            ev.record(compute)
            free_ev[k & 1] = ev
        compute.synchronize()


def to_host_array(t):
    """Device tensor -> numpy array through pinned memory.  A pageable-destination copy is staged by the driver at
    ~2 GB/s (measured: 33 ms for a 3840x2160 f64 aggregate).  Up to 128 MiB the array IS a pinned block (torch's caching
    host allocator recycles it once the array is garbage-collected, so only the first call of a given size pays
    cudaHostAlloc); larger results (8192^2 canvases) go into ordinary memory through the pinned staging ring, so that
    long-lived aggregates do not hold hundreds of MB of page-locked memory each."""
    nbytes = t.numel() * t.element_size()
    if not t.is_cuda or nbytes < (1 << 18):
        return t.cpu().numpy()
    stream = torch.cuda.current_stream(t.device)
    if nbytes <= (128 << 20):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        stream.synchronize()
        return h.numpy()
    t = t.contiguous()
    out = torch.empty(t.shape, dtype=t.dtype)
    ring = _PinnedRing.get(t.device)
    s8, d8 = t.view(-1).view(torch.uint8), out.view(-1).view(torch.uint8)
    pending = []                                         # (slot, offset, bytes, event) of copies in flight
    def drain(entry):
        k, off, m, ev = entry
        ev.synchronize()
        d8[off:off + m].copy_(ring.slots[k][:m])
    for off in range(0, nbytes, ring.SLOT_BYTES):
        m = min(ring.SLOT_BYTES, nbytes - off)
        if len(pending) == ring.NSLOTS:
            drain(pending.pop(0))
        k = ring.next
        ring.next = (k + 1) % ring.NSLOTS
        if ring.events[k] is not None:
            ring.events[k].synchronize()
        with torch.cuda.stream(stream):
            ring.slots[k][:m].copy_(s8[off:off + m], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
        ring.events[k] = ev
        pending.append((k, off, m, ev))
    for entry in pending:
        drain(entry)
    return out.numpy()


def _is_arrow(source):
    mod = type(source).__module__ or ""
    return mod.startswith("pyarrow") and type(source).__name__ in ("Table", "RecordBatch")


def as_frame(source, needed, device=None):
    """pandas.DataFrame | pyarrow.Table / RecordBatch | dict of columns | HostFrame | DeviceFrame -> a frame holding
    the needed columns.  Only those columns are touched (like _bypixel_sanitise, core.py:1384-1392)."""
    import pandas as pd
    if isinstance(source, (DeviceFrame, HostFrame)):
        return source
    if isinstance(source, pd.DataFrame):
        return HostFrame.from_pandas(source, columns=needed, device=device)
    if _is_arrow(source):
        return HostFrame.from_arrow(source, columns=needed, device=device)
    if isinstance(source, dict):
        missing = [c for c in needed if c not in source]
        if missing:
            raise ValueError("specified column not found")
        cols = {c: source[c] for c in needed}
        on_gpu = [hasattr(v, "__cuda_array_interface__") or (isinstance(v, torch.Tensor) and v.is_cuda) for v in cols.values()]
        if cols and all(on_gpu):
            return DeviceFrame(cols)
        if any(on_gpu):
            raise ValueError("columns of a dict source must be all on the host or all on the GPU")
        return HostFrame(cols, device=device)
    raise ValueError("source must be a pandas or dask DataFrame")
