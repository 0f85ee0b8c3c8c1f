"""Multi-GPU combine: one process per GPU, rows sharded contiguously, every rank aggregates its shard
into private accumulator canvases, and the partials are combined with NCCL all-reduces.

This is the reference's dask tree-reduction (data_libraries/dask.py:168-217 + make_combine,
compiler.py:478-507) with the per-reduction `_combine` functions mapped onto collective ops:

    count / _sum_zero       aggs.sum(axis=0)            -> all_reduce SUM   (reductions.py:652-654, 994-996)
    any                     aggs.sum(dtype=bool)        -> all_reduce MAX   (reductions.py:881-883)
    max / min               np.nanmax / np.nanmin       -> all_reduce MAX / MIN on order-preserving keys
                                                           (empty = INT_MIN / INT_MAX instead of NaN)
    _max_row_index          np.maximum                  -> all_reduce MAX   (reductions.py:2289-2297)
    _min_row_index          row_min_in_place            -> all_reduce MIN   (empty = INT64_MAX instead of -1)
    where(max/min)          combine_cpu_2d, strict compare, earlier partition wins ties (reductions.py:2009-2016)
                                                        -> MAX/MIN of the value key, then MIN of the global row id
    where(...) lookup values                            -> owner rank contributes the f64 bit pattern, others 0: SUM

Rows carry global ids (shard offset + local index), as the reference stamps `_datashader_row_offset`
on every partition (dask.py:102-117, reductions.py:87-113).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from . import reductions as rd

_I64_MAX = torch.iinfo(torch.int64).max

_OPS = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}


def arg_rows_from_parts(hi_local, hi_global, rows_local):
    """Per pixel: the earliest global row among the ranks holding the winning selector value."""
    cand = torch.where((hi_local == hi_global) & (rows_local >= 0), rows_local,
                       torch.full_like(rows_local, _I64_MAX))
    return cand


class ShardGroup:
    """Collective helpers bound to a torch.distributed process group."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def _all_reduce(self, t, op):
        dist.all_reduce(t, op=_OPS[op], group=self.group)
        return t

    # ---- accumulator canvases ------------------------------------------------------------------
    def combine(self, accs, canv):
        for a in accs:
            if a.kind.startswith("arg"):
                continue            # packed (value, local row) keys are combined in arg_rows()
            _, op = rd.ACC_INFO[a.kind]
            self._all_reduce(canv[a.key], op)

    def arg_rows(self, ctx, packed, acc, sel_col, rows_out):
        is_max = acc.kind == "argmax32"
        hi_local = packed >> 32                      # signed key32 of the selector value (sentinel preserved)
        hi_global = self._all_reduce(hi_local.clone(), "max" if is_max else "min")
        _lib.check(_lib.lib().dsb_decode_arg(packed.data_ptr(), rd.ACC_OP[acc.kind], ctx.dsb_dtype(sel_col),
                                             ctx.frame.row_offset, None, rows_out.data_ptr(), packed.numel(),
                                             ctx.stream_ptr), "dsb_decode_arg")
        cand = arg_rows_from_parts(hi_local, hi_global, rows_out)
        self._all_reduce(cand, "min")
        return torch.where(cand == _I64_MAX, torch.full_like(cand, -1), cand)

    def sum_bits_f64(self, out, rows):
        """`out` holds the looked-up value on the rank that owns the winning row and 0.0 elsewhere."""
        empty = (rows < 0) | (rows == _I64_MAX)
        bits = out.view(torch.int64)
        bits.masked_fill_(empty, 0)
        self._all_reduce(bits, "sum")
        out = bits.view(torch.float64)
        return torch.where(empty, torch.full_like(out, float("nan")), out)

    # ---- axis=0 glyphs sharded by rows (data_libraries/dask.py:244-266) ------------------------------
    def carry_last_row(self, frame, needed):
        """LineAxis0 / area axis=0 shards: the segment from the previous shard's last vertex to this shard's first
        vertex belongs to this shard.  The reference prepends the previous partition's last row and clears
        plot_start; here every rank all-gathers its last row (needed columns only, as f64) and does the same.
        Returns (frame, plot_start)."""
        from .frame import DeviceFrame
        dev = frame.device
        cols = list(needed)
        last = torch.zeros(len(cols) + 1, dtype=torch.float64, device=dev)
        if len(frame) > 0:
            last[0] = 1.0                                             # "this rank has rows"
            for k, c in enumerate(cols):
                last[k + 1] = frame[c][-1].to(torch.float64)
        gathered = [torch.empty_like(last) for _ in range(self.world)]
        dist.all_gather(gathered, last, group=self.group)
        prev = None
        for r in range(self.rank - 1, -1, -1):                         # nearest previous rank that has rows
            if float(gathered[r][0]) != 0.0:
                prev = gathered[r]
                break
        if prev is None or len(frame) == 0:
            return frame, prev is None
        new_cols = {c: torch.cat([prev[k + 1:k + 2].to(frame[c].dtype), frame[c]]) for k, c in enumerate(cols)}
        out = DeviceFrame(new_cols, frame.categories, frame.row_offset - 1, frame.n_global)
        out.sharded = True
        out.group = getattr(frame, "group", None)
        return out, False

    # ---- auto-ranging (compute_bounds_dask, glyphs/points.py:153-167) -----------------------------
    def global_bounds(self, lo, hi, device):
        t = torch.tensor([lo, -hi], dtype=torch.float64, device=device)
        self._all_reduce(t, "min")
        lo, nhi = t.tolist()
        return lo, -nhi

    # ---- line canvases ---------------------------------------------------------------------------
    def combine_lines(self, line_agg, aa, canvas, mask):
        if line_agg == _lib.LINE_ANY:
            self._all_reduce(canvas, "max")
        elif line_agg in (_lib.LINE_COUNT, _lib.LINE_SUM):
            self._all_reduce(canvas, "sum")
        elif line_agg == _lib.LINE_MAX:
            self._all_reduce(canvas, "max")
        else:
            self._all_reduce(canvas, "min")
        if mask is not None:
            self._all_reduce(mask, "max")
        return canvas, mask


def current_group(source):
    """A ShardGroup when `source` is a shard of a distributed DeviceFrame, else None."""
    if getattr(source, "sharded", False):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("sharded DeviceFrame used without an initialised torch.distributed process group")
        return ShardGroup(getattr(source, "group", None))
    return None


def shard_bounds(n_global, rank, world):
    """Contiguous row range of `rank`: [lo, hi)."""
    return n_global * rank // world, n_global * (rank + 1) // world
