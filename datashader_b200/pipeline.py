"""Host dispatch for the hot path: what data_libraries/pandas.py:26-66 (single pass) and
data_libraries/dask.py:86-219 (partition + combine) do in the reference, with compiler.py's
make_create / make_append / make_combine / make_finalize replaced by a plan handed to libdsb200.

    view  = ranges + Axis.compute_scale_and_translate                  (pandas.py:35-46)
    accs  = unique accumulators of all reductions                      (compiler.py:103-107)
    launch dsb_points / dsb_lines_axis1 once per <= 8 accumulators     (make_append + extend)
    [multi-GPU] all-reduce each accumulator canvas                     (make_combine, dask.py:175-217)
    finalize each reduction, wrap in DataArray / Dataset               (make_finalize, pandas.py:61-66)
"""
from __future__ import annotations

import ctypes as C
from builtins import any as builtins_any

import numpy as np
import torch

from . import _lib
from . import config
from . import reductions as rd
from .frame import to_host_array, DeviceFrame, as_frame
from .glyphs import Point, _column_bounds, maybe_expand_bounds
from .xr_compat import DataArray, Dataset


class _Ctx:
    """Everything a reduction needs to declare accumulators and finish them."""

    def __init__(self, frame, view, shape, dist=None):
        self.frame = frame            # the source (DeviceFrame or HostFrame): dtypes, length, row offset
        self.resident = frame if isinstance(frame, DeviceFrame) else None   # device columns for gathers
        self.view = view
        self.shape = shape            # (H, W) or (H, W, ncat)
        self.dist = dist
        self.stream_ptr = torch.cuda.current_stream(frame.device).cuda_stream

    def np_dtype(self, col):
        return self.frame.np_dtype(col)

    def dsb_dtype(self, col):
        return _lib.dsb_dtype(self.frame.np_dtype(col))


def _auto_range(tensors, stream_ptr, dist, device):
    """Glyph.compute_x_bounds / compute_bounds_dask (glyphs/points.py:145-167)."""
    lo, hi = _column_bounds(stream_ptr, tensors)
    if dist is not None:
        lo, hi = dist.global_bounds(lo, hi, device)
    return maybe_expand_bounds((lo, hi))


def make_view(canvas, x_range, y_range):
    """pandas.py:42-46 + core.py:62-81"""
    x_st = canvas.x_axis.compute_scale_and_translate(x_range, canvas.plot_width)
    y_st = canvas.y_axis.compute_scale_and_translate(y_range, canvas.plot_height)
    v = _lib.View(int(canvas.plot_width), int(canvas.plot_height), int(canvas.x_axis.is_log), int(canvas.y_axis.is_log),
                  float(x_st[0]), float(x_st[1]), float(y_st[0]), float(y_st[1]),
                  float(x_range[0]), float(x_range[1]), float(y_range[0]), float(y_range[1]))
    return v, x_st, y_st


def _alloc_canvas(acc, shape, device, stream_ptr):
    dtype, _ = rd.ACC_INFO[acc.kind]
    t = torch.empty(shape, dtype=dtype, device=device)
    _lib.check(_lib.lib().dsb_init_canvas(rd.ACC_OP[acc.kind], t.data_ptr(), t.numel(), stream_ptr), "dsb_init_canvas")
    return t


def _xy_columns(frame, xname, yname):
    x, y = frame[xname], frame[yname]
    # coordinates are consumed as f32 or f64; anything else is widened the way the reference's
    # f64 arithmetic would see it (glyphs/points.py:199-201)
    if x.dtype != y.dtype or x.dtype not in (torch.float32, torch.float64):
        x, y = x.to(torch.float64), y.to(torch.float64)
    return x.contiguous(), y.contiguous(), (_lib.F32 if x.dtype == torch.float32 else _lib.F64)


def _reductions_of(agg):
    return list(agg.values) if isinstance(agg, rd.summary) else [agg]


def _group_key(r):
    """Reductions that can share one pass share a canvas shape: every plain reduction -> None, by() -> its categorizer."""
    return r._hashable_inputs()[:3] if isinstance(r, rd.by) else None


def _categorical_setup(reds, schema):
    """by(): the categorizer + number of categories (compiler.py:379-390) of one group of reductions (_group_key)."""
    cats = [r for r in reds if isinstance(r, rd.by)]
    if not cats:
        return None, 0, None
    assert len(cats) == len(reds) and len({_group_key(c) for c in cats}) == 1
    categorizer = cats[0].categorizer
    labels = categorizer.categories(schema)
    return categorizer, len(labels), labels


def _plans(chunk, accs, canv, ctx, categorizer, ncat):
    """Yield (plan, keepalive) for groups of <= DSB_MAX_OPS accumulators over one resident row chunk."""
    codes = categorizer.codes(chunk).contiguous() if ncat else None
    for i in range(0, len(accs), _lib.DSB_MAX_OPS):
        group = accs[i:i + _lib.DSB_MAX_OPS]
        plan = _lib.Plan()
        plan.nops = len(group)
        keep = [codes]
        for k, acc in enumerate(group):
            b = plan.ops[k]
            b.op = rd.ACC_OP[acc.kind]
            b.agg = canv[acc.key].data_ptr()
            if acc.col is not None:
                t = chunk[acc.col]
                keep.append(t)
                b.val_dtype, b.val = ctx.dsb_dtype(acc.col), t.data_ptr()
            if acc.chk is not None:
                t = chunk[acc.chk]
                keep.append(t)
                b.chk_dtype, b.chk = ctx.dsb_dtype(acc.chk), t.data_ptr()
            if acc.aux is not None:
                b.aux = canv[acc.aux.key].data_ptr()
        if ncat:
            plan.cat = codes.data_ptr()
            plan.cat_dtype = _lib.dsb_dtype(str(codes.dtype).replace("torch.", ""))
            plan.ncat = ncat
        notes = getattr(ctx, "notes", None)
        if notes is not None:
            plan.notes = notes.data_ptr()
        yield plan, keep


# accumulators dsb_points_routed serves -> bytes per canvas cell
_ROUTED_OPS = {_lib.OP_MAX32: 4, _lib.OP_MIN32: 4, _lib.OP_MINROW: 8, _lib.OP_MAXROW: 8, _lib.OP_COUNT: 4}
_MONO_OPS = (_lib.OP_MAX32, _lib.OP_MIN32, _lib.OP_MINROW, _lib.OP_MAXROW, _lib.OP_ARGMAX32, _lib.OP_ARGMIN32)


def _sub_plan(plan, idx):
    sub = _lib.Plan()
    sub.nops = len(idx)
    for j, k in enumerate(idx):
        sub.ops[j] = plan.ops[k]
    sub.cat, sub.cat_dtype, sub.ncat, sub.notes = plan.cat, plan.cat_dtype, plan.ncat, plan.notes
    return sub


def _specialised_groups(plan):
    """summary()-style plans on a small canvas: route each accumulator to the specialised kernel that serves it best
    instead of one interpreted pass that pays a global RED per accumulator per point - [SUM(c), COUNT(c)] to the K2
    mean shape, a plain COUNT / ANY to K2 count, every monotone accumulator to k_points_mono (filtered).  The columns are
    read once per group; the REDs saved outweigh that (summary(count, mean, max): 18.0 -> 13 ms at 1e9 points)."""
    ops = [plan.ops[k] for k in range(plan.nops)]
    left = set(range(plan.nops))
    groups = []
    for k in sorted(left):
        if k in left and ops[k].op == _lib.OP_SUM and ops[k].val_dtype == _lib.F32 and not ops[k].chk:
            mate = [j for j in left if ops[j].op == _lib.OP_COUNT and ops[j].val == ops[k].val and not ops[j].chk]
            if mate:
                groups.append([k, mate[0]])
                left -= {k, mate[0]}
    for k in sorted(left):
        if ops[k].op in (_lib.OP_COUNT, _lib.OP_ANY) or ops[k].op in _MONO_OPS:
            groups.append([k])
            left.discard(k)
    if left:
        groups.append(sorted(left))
    return groups if len(groups) > 1 else None


def _launch_points(view, chunk, glyph, accs, canv, ctx, categorizer, ncat):
    """One fused launch per <= DSB_MAX_OPS accumulators over one resident row chunk."""
    x, y, xy_dtype = _xy_columns(chunk, glyph.x, glyph.y)
    n, row_offset = len(chunk), chunk.row_offset
    if n > (1 << 32):
        raise NotImplementedError("more than 2^32 rows per device chunk")
    for acc in [a for a in accs if a.kind == "matchrow32"]:
        _launch_match32(view, chunk, x, y, xy_dtype, acc, canv, ctx)
    accs = [a for a in accs if a.kind != "matchrow32"]
    for plan, _keep in _plans(chunk, accs, canv, ctx, categorizer, ncat):
        groups = None
        if (config.split_summary and plan.nops >= 3 and ncat == 0 and xy_dtype == _lib.F32 and n >= config.priv_min_rows
                and int(np.prod(ctx.shape)) <= 954_000):
            groups = _specialised_groups(plan)
        for sub in ([_sub_plan(plan, g) for g in groups] if groups else [plan]):
            _launch_points_plan(view, x, y, xy_dtype, n, row_offset, sub, ctx)


def _launch_match32(view, chunk, x, y, xy_dtype, acc, canv, ctx):
    """Second pass of the two-pass where(max | min): rows whose key equals the finished extreme vote their row id."""
    lib = _lib.lib()
    need = int(lib.dsb_points_match32_scratch_bytes(C.byref(view)))
    scratch = getattr(ctx, "_match32_scratch", None)
    if scratch is None or scratch.numel() < need:
        scratch = ctx._match32_scratch = torch.empty(need, dtype=torch.uint8, device=x.device)
    val = chunk[acc.col]
    _lib.check(lib.dsb_points_match32(C.byref(view), x.data_ptr(), y.data_ptr(), xy_dtype, len(chunk), chunk.row_offset,
                                      val.data_ptr(), ctx.dsb_dtype(acc.col), canv[acc.aux.key].data_ptr(),
                                      int(acc.aux.kind == "max32"), canv[acc.key].data_ptr(), scratch.data_ptr(), scratch.numel(),
                                      ctx.stream_ptr), "dsb_points_match32")


def _launch_points_plan(view, x, y, xy_dtype, n, row_offset, plan, ctx):
    """One plan over one resident chunk: K2 (privatised count) when it applies, the 16-bit packed count for canvases
    of 1-2x the L2 budget, otherwise dsb_points (which picks the mono / generic / banded / split forms itself)."""
    lib = _lib.lib()
    ncell = int(np.prod(ctx.shape))
    if config.priv_count and xy_dtype in (_lib.F32, _lib.F64) and n >= config.priv_min_rows:
        priv = [k for k in range(plan.nops) if plan.ops[k].op == _lib.OP_COUNT] or \
               [k for k in range(plan.nops) if plan.ops[k].op == _lib.OP_ANY and n < (1 << 32)]
        if priv and ncell <= 954_000:      # 226 KB of 2-bit fields / 0.97; beyond, dsb_points_priv returns UNSUPPORTED
            scratch = getattr(ctx, "_priv_scratch", None)
            if scratch is None:
                scratch = ctx._priv_scratch = torch.empty(ncell + 1, dtype=torch.int32, device=x.device)
            rc = lib.dsb_points_priv(C.byref(view), x.data_ptr(), y.data_ptr(), xy_dtype, n, row_offset, C.byref(plan),
                                     priv[0], scratch.data_ptr(), scratch.data_ptr() + 4 * ncell, ctx.stream_ptr)
            if rc == 0:
                return
            if rc != -3:
                _lib.check(rc, "dsb_points_priv")
    if (config.count16 and plan.nops == 1 and plan.ops[0].op == _lib.OP_COUNT and n >= config.count16_min_rows
            and 4 * ncell > config.l2_budget_bytes >= 2 * ncell):
        # u32 canvas beyond L2 but its 16-bit packed form fits: one L2-resident pass instead of two banded ones
        scratch = getattr(ctx, "_count16_scratch", None)
        if scratch is None:
            scratch = ctx._count16_scratch = torch.empty(4 * ((ncell + 1) // 2) + 48, dtype=torch.uint8, device=x.device)
        rc = lib.dsb_points_count16(C.byref(view), x.data_ptr(), y.data_ptr(), xy_dtype, n, row_offset, C.byref(plan),
                                    scratch.data_ptr(), scratch.numel(), ctx.stream_ptr)
        if rc == 0:
            return
        if rc != -3:
            _lib.check(rc, "dsb_points_count16")
    if (config.routed and plan.nops == 1 and plan.ncat == 0 and xy_dtype == _lib.F32 and n >= config.routed_min_rows
            and plan.ops[0].op in _ROUTED_OPS
            and (ncell * _ROUTED_OPS[plan.ops[0].op] > config.l2_budget_bytes
                 or (plan.ops[0].op in (_lib.OP_MINROW, _lib.OP_MAXROW) and n >= config.routed_rows_per_cell_for_first * ncell))):
        # the accumulator canvas is beyond L2: route the points to shared-memory-sized buckets instead of banding.  first / last
        # with many rows per pixel, on ANY canvas: dsb_points_routed routes only the head (tail) of the rows - 10 per canvas cell -
        # and drops the rest against a bitmap of settled pixels (k_rows_rest); the filtered mono kernel pays an L2 load per row
        n_scratch = n
        if ncell * _ROUTED_OPS[plan.ops[0].op] <= config.l2_budget_bytes:
            # first / last on an L2-resident canvas: only the head of the rows is routed.  The record buffer is sized for that; should
            # the device-side sample hand the rest to the routed kernels too (rows sorted in space), the records that do not fit are
            # applied with direct atomics - on a canvas that sits in L2
            n_scratch = min(n, 2 * config.routed_rows_per_cell_for_first * ncell)
        need = int(lib.dsb_points_routed_scratch_bytes(C.byref(view), n_scratch))
        if 0 < need <= config.routed_max_scratch_bytes:
            scratch = getattr(ctx, "_routed_scratch", None)
            if scratch is None or scratch.numel() < need:
                try:
                    scratch = ctx._routed_scratch = torch.empty(need, dtype=torch.uint8, device=x.device)
                except torch.OutOfMemoryError:           # no room for the records: the L2-banded kernels need no scratch
                    scratch = None
            if scratch is not None:
                rc = lib.dsb_points_routed(C.byref(view), x.data_ptr(), y.data_ptr(), xy_dtype, n, row_offset, C.byref(plan),
                                           scratch.data_ptr(), scratch.numel(), ctx.stream_ptr)
                if rc == 0:
                    return
                if rc != -3:
                    _lib.check(rc, "dsb_points_routed")
    op0 = plan.ops[0]
    if (config.minmax_split_rows_per_cell and plan.nops == 1 and plan.ncat == 0 and xy_dtype == _lib.F32
            and op0.op in (_lib.OP_MAX32, _lib.OP_MIN32, _lib.OP_ARGMAX32, _lib.OP_ARGMIN32) and op0.val_dtype == _lib.F32
            and op0.chk_dtype == _lib.NONE and (4 if op0.op in (_lib.OP_MAX32, _lib.OP_MIN32) else 8) * ncell <= config.l2_budget_bytes
            and n >= config.routed_min_rows and n >= config.minmax_split_rows_per_cell * ncell
            and (op0.op in (_lib.OP_MAX32, _lib.OP_MIN32) or (row_offset + n) < (1 << 32))):
        # many rows per pixel: after the head the extreme of a pixel rarely changes - the rest only needs a threshold test per row
        n_head = (config.minmax_head_rows_per_cell * ncell + 3) & ~3
        _lib.check(lib.dsb_points(C.byref(view), x.data_ptr(), y.data_ptr(), xy_dtype, n_head, row_offset, C.byref(plan),
                                  ctx.stream_ptr), "dsb_points")
        scratch = getattr(ctx, "_minmax_scratch", None)
        if scratch is None:
            scratch = ctx._minmax_scratch = torch.empty(1 << 18, dtype=torch.uint8, device=x.device)
        if op0.op in (_lib.OP_MAX32, _lib.OP_MIN32):
            rc = lib.dsb_points_minmax_rest(C.byref(view), x.data_ptr() + 4 * n_head, y.data_ptr() + 4 * n_head, xy_dtype, n - n_head,
                                            row_offset + n_head, op0.val + 4 * n_head, op0.val_dtype, op0.agg,
                                            int(op0.op == _lib.OP_MAX32), plan.notes, scratch.data_ptr(), scratch.numel(), ctx.stream_ptr)
        else:       # where(max | min): the packed {key32, row} accumulator
            rc = lib.dsb_points_argminmax_rest(C.byref(view), x.data_ptr() + 4 * n_head, y.data_ptr() + 4 * n_head, xy_dtype, n - n_head,
                                               row_offset + n_head, op0.val + 4 * n_head, op0.val_dtype, op0.agg,
                                               int(op0.op == _lib.OP_ARGMAX32), scratch.data_ptr(), scratch.numel(), ctx.stream_ptr)
        if rc == 0:
            return
        if rc != -3:
            _lib.check(rc, "dsb_points_minmax_rest / dsb_points_argminmax_rest")
        rest = _lib.Plan.from_buffer_copy(plan)          # not served (axes / alignment): the same rows through dsb_points
        rest.ops[0].val = op0.val + 4 * n_head
        _lib.check(lib.dsb_points(C.byref(view), x.data_ptr() + 4 * n_head, y.data_ptr() + 4 * n_head, xy_dtype, n - n_head,
                                  row_offset + n_head, C.byref(rest), ctx.stream_ptr), "dsb_points")
        return
    _lib.check(lib.dsb_points(C.byref(view), x.data_ptr(), y.data_ptr(), xy_dtype, n, row_offset, C.byref(plan),
                              ctx.stream_ptr), "dsb_points")


def _launch_lines(view, chunk, glyph, accs, canv, ctx, categorizer, ncat):
    """Bresenham lines: the same accumulator plans, one row of the frame = one line (i = line index)."""
    lib = _lib.lib()
    xs, ys, xy_dtype, nlines, nverts, layout = ctx.line_vertices
    for plan, _keep in _plans(chunk, accs, canv, ctx, categorizer, ncat):
        _lib.check(lib.dsb_lines_axis1_plan(C.byref(view), xs.data_ptr(), ys.data_ptr(), xy_dtype, nlines, nverts,
                                            C.byref(layout), chunk.row_offset, C.byref(plan), ctx.stream_ptr),
                   "dsb_lines_axis1_plan")


def _to_host(t, np_view=None):
    if config.device_results:
        if np_view is np.uint32:
            return t.view(torch.uint32)
        if np_view is np.bool_:
            return t.bool()
        return t
    a = to_host_array(t)
    if np_view is not None:
        a = a.view(np_view)
    return a


def _prepare(source, glyph, agg, canvas):
    needed = list(dict.fromkeys(glyph.required_columns() + agg.columns_needed))
    frame = as_frame(source, needed)
    schema = frame.schema()
    for c in needed:
        if c not in schema:
            raise ValueError("specified column not found")
    glyph.validate(schema)
    agg.validate(schema)
    canvas.validate()
    return needed, frame, schema


def _accumulate_and_finalize(frame, resident, needed, schema, view, canvas, glyph, agg, dist, launch, ctx_extra=None):
    """summary() composes any bases (compiler.py:103-107, reductions.py:2169-2246): reductions are grouped by canvas shape -
    the plain ones [H, W], every categorizer its own [H, W, C] - and each group runs as one accumulate-and-finalize.
    Returns (reductions, results, category labels per reduction)."""
    reds = _reductions_of(agg)
    groups = {}
    for i, r in enumerate(reds):
        groups.setdefault(_group_key(r), []).append(i)
    results, labels = [None] * len(reds), [None] * len(reds)
    for idx in groups.values():
        res, lab = _accumulate_group(frame, resident, needed, schema, view, canvas, glyph, [reds[i] for i in idx], dist, launch,
                                     ctx_extra)
        for i, t in zip(idx, res):
            results[i], labels[i] = t, lab
    return reds, results, labels


def _accumulate_group(frame, resident, needed, schema, view, canvas, glyph, reds, dist, launch, ctx_extra=None):
    """accumulators -> fused launches (per chunk, per stage) -> [all-reduce] -> finalize, for reductions of one shape."""
    device = frame.device
    stream_ptr = torch.cuda.current_stream(device).cuda_stream
    single = resident is not None
    categorizer, ncat, labels = _categorical_setup(reds, schema)
    nviews = (ctx_extra or {}).get("nviews")          # points_batch: canvases stacked [V, H, W(, C)]
    shape = ((nviews,) if nviews else ()) + (canvas.plot_height, canvas.plot_width) + ((ncat,) if ncat else ())
    ctx = _Ctx(frame, view, shape, dist)
    if single:
        ctx.resident = resident
    for k, v in (ctx_extra or {}).items():
        setattr(ctx, k, v)

    def unique_accs(lists):
        out, seen = [], set()
        for a in lists:
            if a.key not in seen:
                seen.add(a.key)
                out.append(a)
        return out

    canv = {}

    def run(accs):
        """allocate the canvases of `accs`, launch them (stage 0, then the accumulators that read a finished stage-0
        canvas) over every chunk and combine them across ranks"""
        accs = [a for a in accs if a.key not in canv]
        canv.update({a.key: _alloc_canvas(a, shape, device, stream_ptr) for a in accs})
        for stage in ([a for a in accs if a.aux is None], [a for a in accs if a.aux is not None]):
            if not stage:
                continue
            if config.time_kernels:
                ev0 = torch.cuda.Event(enable_timing=True)
                ev0.record()
            for ch in ([resident] if single else frame.chunks(needed)):
                launch(view, ch, glyph, stage, canv, ctx, categorizer, ncat)
            if config.time_kernels:
                ev1 = torch.cuda.Event(enable_timing=True)
                ev1.record()
                config.kernel_events.append((ev0, ev1))
            if dist is not None:
                dist.combine(stage, canv)

    accs = unique_accs([a for r in reds for a in r._accs(ctx)])
    # max / min of a float column: the keys fold -0.0 onto +0.0, so the pass reports whether it met a -0.0 at all
    # (DSB_NOTE_NEGZERO, include/dsb200.h); only then is the sign of a zero extreme in doubt
    zero_keyed = [a for a in accs if a.kind in ("max32", "min32", "max64", "min64") and a.col is not None
                  and np.dtype(ctx.np_dtype(a.col)).kind == "f"]
    ctx.notes = torch.zeros(1, dtype=torch.int32, device=device) if zero_keyed else None
    run(accs)
    if ctx.notes is not None:
        if dist is not None:
            dist._all_reduce(ctx.notes, "max")
        if int(ctx.notes.item()) & _lib.NOTE_NEGZERO:
            # rare: redo those reductions through the row-exact accumulators (the earliest row among the extreme's
            # ties, exactly the reference's strict compare) and gather that row's own bit pattern
            ctx.exact_zero = True
            ctx.notes = None
            run(unique_accs([a for r in reds for a in r._exact_accs(ctx)]))
    return [r._finalize(ctx, canv) for r in reds], labels


def points(source, canvas, glyph: Point, agg, dist=None):
    """bypixel for Point glyphs."""
    needed, frame, schema = _prepare(source, glyph, agg, canvas)
    device = frame.device
    with torch.cuda.device(device):
        stream_ptr = torch.cuda.current_stream(device).cuda_stream
        single = frame.n_chunks() == 1
        resident = frame.resident(needed) if single else None      # one H2D for small host sources
        if canvas.x_range is None or canvas.y_range is None:
            # the reference's separate bounds pass (pandas.py:35-36, glyph.py:66-78)
            bx, by = [], []
            for ch in ([resident] if single else frame.chunks([glyph.x, glyph.y])):
                bx.append(_column_bounds(stream_ptr, [ch[glyph.x]]))
                by.append(_column_bounds(stream_ptr, [ch[glyph.y]]))
            x_range = canvas.x_range or _merge_bounds(bx, dist, device)
            y_range = canvas.y_range or _merge_bounds(by, dist, device)
        else:
            x_range, y_range = canvas.x_range, canvas.y_range
        canvas.validate_ranges(x_range, y_range)
        view, x_st, y_st = make_view(canvas, x_range, y_range)
        # where(max | min) as two passes when the packed {key, row} canvas would not fit L2 and the routed kernels apply
        # (float32 coordinates, linear axes, enough rows): see where._row_accs
        match32 = (config.where_two_pass and config.routed and len(frame) >= config.routed_min_rows
                   and 8 * canvas.plot_width * canvas.plot_height > config.l2_budget_bytes
                   and frame.np_dtype(glyph.x) == np.float32 and frame.np_dtype(glyph.y) == np.float32
                   and not canvas.x_axis.is_log and not canvas.y_axis.is_log)
        reds, results, labels = _accumulate_and_finalize(frame, resident, needed, schema, view, canvas, glyph, agg, dist,
                                                         _launch_points, ctx_extra={"match32": match32})
    x_axis = canvas.x_axis.compute_index(x_st, canvas.plot_width)
    y_axis = canvas.y_axis.compute_index(y_st, canvas.plot_height)
    return _wrap(agg, reds, results, glyph, x_axis, y_axis, x_range, y_range, labels)


def points_batch(source, canvas, glyph: Point, agg, views, grid=None, dist=None):
    """Canvas.points for several views at once: ONE pass over the columns fills a canvas per view (dsb_points_views).
    views: [(x_range, y_range), ...]; grid=(nx, ny) declares them a row-major nx x ny grid of equal extents (a tile level),
    otherwise at most 64 arbitrary views.  Returns one DataArray / Dataset per view, each identical to what
    Canvas(..., x_range, y_range).points(source, ...) returns for that view."""
    needed, frame, schema = _prepare(source, glyph, agg, canvas)
    views = [(tuple(map(float, xr)), tuple(map(float, yr))) for xr, yr in views]
    nv = len(views)
    if nv == 0:
        return []
    gx = (0, 0, 0.0, 0.0, 1.0, 1.0)
    if grid is not None:
        nx, ny = int(grid[0]), int(grid[1])
        if nx * ny != nv:
            raise ValueError("grid=(nx, ny) must describe len(views) views")
        (x0, x1), (y0, y1) = views[0]
        gx = (nx, ny, x0, y0, x1 - x0, y1 - y0)
        # the kernel finds a point's tile from the grid and looks across an edge only within 1/1024 of a tile of it: the views
        # must BE that grid (up to rounding of the tile extents)
        tw, th = x1 - x0, y1 - y0
        for k, ((ax0, ax1), (ay0, ay1)) in enumerate(views):
            ix, iy = k % nx, k // nx
            if (abs(ax0 - (x0 + ix * tw)) > 1e-6 * tw or abs(ax1 - (x0 + (ix + 1) * tw)) > 1e-6 * tw
                    or abs(ay0 - (y0 + iy * th)) > 1e-6 * th or abs(ay1 - (y0 + (iy + 1) * th)) > 1e-6 * th):
                raise ValueError("grid=(nx, ny): the views are not a row-major grid of equal extents")
    elif nv > 64:
        raise ValueError("more than 64 views need grid=(nx, ny)")
    device = frame.device
    with torch.cuda.device(device):
        single = frame.n_chunks() == 1
        resident = frame.resident(needed) if single else None
        vstructs, sts = [], []
        for xr, yr in views:
            canvas.validate_ranges(xr, yr)
            v, x_st, y_st = make_view(canvas, xr, yr)
            vstructs.append(v)
            sts.append((x_st, y_st))
        varr = (_lib.View * nv)(*vstructs)
        dev_views = torch.frombuffer(bytearray(bytes(varr)), dtype=torch.uint8).to(device)

        def launch(view, chunk, glyph_, accs, canv, ctx, categorizer, ncat):
            x, y, xy_dtype = _xy_columns(chunk, glyph_.x, glyph_.y)
            cells = int(canvas.plot_height) * int(canvas.plot_width) * max(ncat, 1)
            lib = _lib.lib()
            for gi, (plan, _keep) in enumerate(_plans(chunk, accs, canv, ctx, categorizer, ncat)):
                group = accs[gi * _lib.DSB_MAX_OPS:(gi + 1) * _lib.DSB_MAX_OPS]
                sizes = [(canv[acc.key].element_size(), canv[acc.aux.key].element_size() if acc.aux is not None else 0) for acc in group]
                # a tile level whose stacked canvases exceed L2 is done in bands of tile ROWS (one launch per band over all the
                # points): random REDs into DRAM-resident canvases cost far more than re-reading the columns
                bands = [(0, gx[1] if gx[0] else 0)]
                per_row = sum(a_ + b_ for a_, b_ in sizes) * cells * max(gx[0], 1)
                if gx[0] and gx[1] > 1 and per_row * gx[1] > config.l2_budget_bytes:
                    rows = max(1, int(config.l2_budget_bytes // per_row))
                    bands = [(r0, min(r0 + rows, gx[1])) for r0 in range(0, gx[1], rows)]
                for r0, r1 in bands:
                    sub, v0 = plan, 0
                    if r0 or (gx[0] and r1 != gx[1]):
                        v0 = r0 * gx[0]
                        sub = _sub_plan(plan, list(range(plan.nops)))
                        for k, (sa, sx) in enumerate(sizes):
                            sub.ops[k].agg = plan.ops[k].agg + v0 * cells * sa
                            if sx:
                                sub.ops[k].aux = plan.ops[k].aux + v0 * cells * sx
                    nvb = (r1 - r0) * gx[0] if gx[0] else nv
                    _lib.check(lib.dsb_points_views(dev_views.data_ptr() + v0 * C.sizeof(_lib.View), nvb, gx[0], (r1 - r0) if gx[0] else 0,
                                                    gx[2], gx[3] + r0 * gx[5], gx[4], gx[5],
                                                    x.data_ptr(), y.data_ptr(), xy_dtype, len(chunk), chunk.row_offset,
                                                    C.byref(sub), cells, ctx.stream_ptr), "dsb_points_views")

        reds, results, labels = _accumulate_and_finalize(frame, resident, needed, schema, vstructs[0], canvas, glyph, agg, dist,
                                                         launch, ctx_extra={"nviews": nv})
    out = []
    xcache, ycache = {}, {}                    # a tile level has nx distinct x ranges and ny distinct y ranges, not nx * ny
    for k, ((xr, yr), (x_st, y_st)) in enumerate(zip(views, sts)):
        x_axis = xcache.get(xr)
        if x_axis is None:
            x_axis = xcache[xr] = canvas.x_axis.compute_index(x_st, canvas.plot_width)
        y_axis = ycache.get(yr)
        if y_axis is None:
            y_axis = ycache[yr] = canvas.y_axis.compute_index(y_st, canvas.plot_height)
        out.append(_wrap(agg, reds, [t[k] for t in results], glyph, x_axis, y_axis, xr, yr, labels))
    return out


def _merge_bounds(parts, dist, device):
    lo = np.min([p[0] for p in parts])
    hi = np.max([p[1] for p in parts])
    if dist is not None:
        lo, hi = dist.global_bounds(float(lo), float(hi), device)
    return maybe_expand_bounds((float(lo), float(hi)))


def _wrap(agg, reds, results, glyph, x_axis, y_axis, x_range, y_range, labels):
    """make_finalize (compiler.py:510-536) + by._build_finalize (reductions.py:809-820)."""
    def one(r, t, lab):
        coords = {glyph.y_label: y_axis, glyph.x_label: x_axis}
        dims = [glyph.y_label, glyph.x_label]
        if isinstance(r, rd.by):
            dims = dims + [r.cat_column]
            coords[r.cat_column] = lab
        data = _to_host(t, getattr(r, "_out_np_view", None))
        return DataArray(data, coords=coords, dims=dims, attrs=dict(x_range=x_range, y_range=y_range))

    if isinstance(agg, rd.summary):
        return Dataset({k: one(r, t, lab) for k, r, t, lab in zip(agg.keys, reds, results, labels)},
                       attrs=dict(x_range=x_range, y_range=y_range))
    return one(reds[0], results[0], labels[0])


# ------------------------------------------------------------------------------------------ lines
def _line_setup(frame, canvas, glyph, dist):
    """ranges, view and the vertex matrices + dsb_line_layout of any line layout."""
    device = frame.device
    stream_ptr = torch.cuda.current_stream(device).cuda_stream
    x_range = canvas.x_range or _auto_range(glyph._x_tensors(frame), stream_ptr, dist, device)
    y_range = canvas.y_range or _auto_range(glyph._y_tensors(frame), stream_ptr, dist, device)
    canvas.validate_ranges(x_range, y_range)
    view, x_st, y_st = make_view(canvas, x_range, y_range)
    xs, ys, (xls, yls) = glyph.vertices(frame)
    if xs.dtype != ys.dtype:
        xs, ys = xs.to(torch.float64), ys.to(torch.float64)
    xy_dtype = _lib.F32 if xs.dtype == torch.float32 else _lib.F64
    nlines, nverts = int(max(xs.shape[0], ys.shape[0])), int(xs.shape[1])
    layout = _lib.LineLayout(int(xls), int(yls), int(glyph.value_per_vertex), int(getattr(frame, "plot_start", True)))
    if glyph.ragged:
        nlines, nverts = _ragged_layout(layout, glyph.ragged_starts(frame), [xs, ys]), 2
    return x_range, y_range, view, x_st, y_st, (xs, ys, xy_dtype, nlines, nverts, layout)


def _ragged_layout(layout, starts, flats):
    """Fill the ragged half of a dsb_line_layout: a start index per row for every flat vertex array (x, y[, y stack])."""
    names = ("x", "y", "y1")
    for name, st, flat in zip(names, starts, flats):
        setattr(layout, name + "_starts", st.data_ptr())
        setattr(layout, name + "_flat_len", int(flat.numel()))
    if len({int(st.shape[0]) for st in starts}) != 1:
        raise ValueError("ragged columns must have the same number of rows")
    layout._keep = list(starts)          # ctypes holds the addresses only
    return int(starts[0].shape[0])


def lines(source, canvas, glyph, agg, antialias=False, dist=None):
    """bypixel for the line glyphs (LineAxis0, LineAxis0Multi, LinesAxis1, LinesAxis1X/YConstant)."""
    needed, frame, schema = _prepare(source, glyph, agg, canvas)
    frame = frame.resident(needed)      # lines are staged whole ([nlines, nverts] matrices)
    if dist is not None and glyph.value_per_vertex:
        frame, frame.plot_start = dist.carry_last_row(frame, needed)     # row shards of one long line (dask.py:244-266)
    line_width = float(glyph._line_width)
    simple = config.lines_simple_path and type(agg) in (rd.any, rd.count, rd.sum, rd.max, rd.min)
    if simple and line_width == 0 and type(agg) in (rd.max, rd.min) and frame.np_dtype(agg.column).kind == "f":
        v = frame[agg.column]
        if bool(((v == 0) & torch.signbit(v)).any()):
            simple = False       # a -0.0 value: the plan path resolves which zero arrived first (DSB_NOTE_NEGZERO)
    if line_width == 0 and not simple:
        return _lines_plan(frame, needed, schema, canvas, glyph, agg, dist)
    if line_width > 0:
        return _lines_antialiased(frame, schema, canvas, glyph, agg, line_width, dist)
    return _lines_single_stage(frame, schema, canvas, glyph, agg, line_width, dist)


def _aa_inner(r):
    return r.reduction if isinstance(r, rd.by) else r


def _aa_requires_2_stages(r):
    """Reduction._antialias_requires_2_stages (reductions.py:370-374, 506-507, 779-780, 1009-1010, 1169-1170, 1349-1350)."""
    r = _aa_inner(r)
    if isinstance(r, (rd.count, rd.sum)):
        return not r.self_intersect
    return isinstance(r, (rd.min, rd.first, rd.last))


def _lines_antialiased(frame, schema, canvas, glyph, agg, line_width, dist):
    """Antialiased lines.  make_antialias_stage_2 (compiler.py:539-554): if ANY requested reduction needs the 2-stage
    procedure, every count / sum of the call is computed without self-intersection; each reduction then runs through the
    kernel that implements its stage-2 combination (antialias.py:30-58) and a summary() is the Dataset of those."""
    reds = _reductions_of(agg)
    force = builtins_any(_aa_requires_2_stages(r) for r in reds)
    outs = []
    for r in reds:
        inner = _aa_inner(r)
        if force and isinstance(inner, (rd.count, rd.sum)) and inner.self_intersect:
            inner = type(inner)(inner.column, self_intersect=False)
            r = rd.by(r.categorizer, inner) if isinstance(r, rd.by) else inner
        if force and isinstance(inner, rd.mean):
            outs.append(_lines_aa2_by(frame, schema, canvas, glyph, r, line_width, dist) if isinstance(r, rd.by)
                        else _lines_aa_mean_2stage(frame, canvas, glyph, inner, line_width, dist))
            continue
        if isinstance(inner, rd.where):
            outs.append(_lines_aa2_by(frame, schema, canvas, glyph, r, line_width, dist) if isinstance(r, rd.by)
                        else _lines_aa_where(frame, canvas, glyph, inner, line_width, dist))
        elif _aa2_combo(inner) is None:
            if getattr(inner, "_line_agg", None) is None:
                raise NotImplementedError(f"{type(inner).__name__} is not implemented for antialiased datashader_b200 lines yet")
            outs.append(_lines_single_stage(frame, schema, canvas, glyph, r, line_width, dist))
        elif isinstance(r, rd.by):
            outs.append(_lines_aa2_by(frame, schema, canvas, glyph, r, line_width, dist))
        else:
            outs.append(_lines_aa2(frame, canvas, glyph, inner, _aa2_combo(inner), line_width, dist))
    if isinstance(agg, rd.summary):
        return Dataset(dict(zip(agg.keys, outs)), attrs=dict(outs[0].attrs))
    return outs[0]


def _lines_aa2_by(frame, schema, canvas, glyph, agg, line_width, dist):
    """by(cat, <2-stage reduction | where(...) | mean next to a 2-stage member>) on antialiased lines: the stage-2
    combination is per category plane (categorical=True, reductions.py:782-787), i.e. each category's lines are folded on
    their own - one run per category over that category's lines (line order, hence first / last and the row ids of
    where(), is preserved inside a category; where()'s row ids are mapped back to the rows of the whole frame)."""
    if glyph_per_vertex(glyph):
        raise NotImplementedError("by() over per-vertex line glyphs is not implemented for 2-stage antialiasing")
    categorizer, ncat, labels = _categorical_setup([agg], schema)
    codes = categorizer.codes(frame)
    inner = agg.reduction
    is_where = isinstance(inner, rd.where)
    if is_where and dist is not None:
        raise NotImplementedError("by(cat, where(...)) on antialiased lines over sharded frames")
    # ranges must be those of the whole frame, not of a category's subset
    x_range, y_range = _line_setup(frame, canvas, glyph, dist)[:2]
    import copy
    sub_canvas = copy.copy(canvas)
    sub_canvas.x_range, sub_canvas.y_range = tuple(x_range), tuple(y_range)
    planes, first = [], None
    for c in range(ncat):
        sel = (codes == c) | (codes == c - ncat)            # negative codes wrap like numba's agg[:, :, -1]
        sub = DeviceFrame({k: frame[k][sel].contiguous() for k in frame.columns}, frame.categories, 0, None)
        sub.sharded, sub.group = getattr(frame, "sharded", False), getattr(frame, "group", None)
        if is_where:
            part = _lines_aa_where(sub, sub_canvas, glyph, inner, line_width, dist)
        elif isinstance(inner, rd.mean):
            part = _lines_aa_mean_2stage(sub, sub_canvas, glyph, inner, line_width, dist)
        else:
            part = _lines_aa2(sub, sub_canvas, glyph, inner, _aa2_combo(inner), line_width, dist)
        first = first or part
        plane = torch.as_tensor(part.data, device=frame.device)
        if is_where and inner.column == rd.SpecialColumn.RowIndex:      # rows of the subset -> rows of the frame
            idx = torch.nonzero(sel).squeeze(1)
            if idx.numel():
                plane = torch.where(plane >= 0, idx[plane.clamp(min=0)] + frame.row_offset, torch.full_like(plane, -1))
        planes.append(plane)
    data = torch.stack(planes, dim=-1)
    data = data if config.device_results else data.cpu().numpy()
    coords = dict(first.coords)
    coords[agg.cat_column] = list(labels)
    return DataArray(data, coords=coords, dims=list(first.dims) + [agg.cat_column], attrs=dict(first.attrs))


def _lines_aa_mean_2stage(frame, canvas, glyph, agg, line_width, dist):
    """mean(col) in a summary that also holds a 2-stage reduction.  Its bases (_sum_zero, _count_ignore_antialiasing) are
    plain FloatingReductions: they keep their single-stage appends (reductions.py:965-973, 679-685), but everything is now
    drawn in overwrite mode (antialias.py:47-56, so prev_aa_factor is always 0) into a per-line canvas that is summed into the
    result (SUM_2AGG, :946-950, 673-677) - i.e. every (segment, pixel) touch adds value x coverage and counts 1."""
    device = frame.device
    with torch.cuda.device(device):
        stream_ptr = torch.cuda.current_stream(device).cuda_stream
        x_range, y_range, view, x_st, y_st, (xs, ys, xy_dtype, nlines, nverts, layout) = _line_setup(frame, canvas, glyph, dist)
        H, W = canvas.plot_height, canvas.plot_width
        sums = torch.zeros((H, W), dtype=torch.float64, device=device)
        counts = torch.zeros((H, W), dtype=torch.int32, device=device)
        val = frame[agg.column]
        _lib.check(_lib.lib().dsb_lines_axis1_cat(C.byref(view), xs.data_ptr(), ys.data_ptr(), xy_dtype, nlines, nverts, C.byref(layout),
                                                  val.data_ptr(), _lib.dsb_dtype(frame.np_dtype(agg.column)), _lib.LINE_MEAN_2STAGE,
                                                  line_width, sums.data_ptr(), counts.data_ptr(), None, _lib.NONE, 0, stream_ptr),
                   "dsb_lines_axis1_cat")
        if dist is not None:
            dist._all_reduce(sums, "sum")
            dist._all_reduce(counts, "sum")
        out = torch.where(counts > 0, sums / counts.to(torch.float64), torch.full_like(sums, float("nan")))
        data = _to_host(out)
    x_axis = canvas.x_axis.compute_index(x_st, canvas.plot_width)
    y_axis = canvas.y_axis.compute_index(y_st, canvas.plot_height)
    return DataArray(data, coords={glyph.y_label: y_axis, glyph.x_label: x_axis}, dims=[glyph.y_label, glyph.x_label],
                     attrs=dict(x_range=x_range, y_range=y_range))


def _lines_aa_where(frame, canvas, glyph, agg, line_width, dist):
    """where(first | last | max | min (col)[, other]) on antialiased lines.  first / last: the selector's 2-stage combination
    keeps, per pixel, the first / last LINE that covers it with a non-null value (nanfirst / nanlast over whole lines,
    compiler.py:198-268, reductions.py:1906-1914, 1992-2027) - exactly the line dsb_lines_aa2's first phase votes for.
    max / min: the line whose value x coverage is the pixel's maximum / minimum of the per-line maxima, the earlier line on
    ties (strict compare, reductions.py:2009-2016): value canvas first, then the lowest matching line index."""
    sel = agg.selector
    combo = _aa2_combo(sel) if isinstance(sel, (rd.first, rd.last)) else (_lib.AA2_ARGMAX if isinstance(sel, rd.max) else _lib.AA2_ARGMIN)
    device = frame.device
    with torch.cuda.device(device):
        stream_ptr = torch.cuda.current_stream(device).cuda_stream
        rows, (x_range, y_range, x_st, y_st) = _lines_aa2(frame, canvas, glyph, sel, combo, line_width, dist, rows_only=True)
        _lib.check(_lib.lib().dsb_finish_minrow(rows.data_ptr(), rows.numel(), stream_ptr), "dsb_finish_minrow")
        if agg.column == rd.SpecialColumn.RowIndex:
            out = rows
        else:
            col = frame[agg.column]
            local = rows - frame.row_offset
            mine = (local >= 0) & (local < len(frame)) & (rows >= 0)
            out = torch.where(mine, col[local.clamp(0, max(len(frame) - 1, 0))].to(torch.float64), torch.zeros((), dtype=torch.float64, device=device))
            if dist is not None:
                out = dist.sum_bits_f64(out, rows)
            else:
                out = torch.where(rows >= 0, out, torch.full_like(out, float("nan")))
        data = _to_host(out)
    x_axis = canvas.x_axis.compute_index(x_st, canvas.plot_width)
    y_axis = canvas.y_axis.compute_index(y_st, canvas.plot_height)
    return DataArray(data, coords={glyph.y_label: y_axis, glyph.x_label: x_axis}, dims=[glyph.y_label, glyph.x_label],
                     attrs=dict(x_range=x_range, y_range=y_range))


def _lines_single_stage(frame, schema, canvas, glyph, agg, line_width, dist):
    """dsb_lines_axis1[_cat]: Bresenham any / count / sum / max / min and the single-stage antialiased reductions
    (any / count / sum / max / mean), optionally per category (by)."""
    red = agg.reduction if isinstance(agg, rd.by) else agg
    if (isinstance(red, (rd.summary, rd.by)) or getattr(red, "_line_agg", None) is None
            or (line_width > 0 and _aa2_combo(red) is not None)):
        raise NotImplementedError(f"{type(agg).__name__} is not implemented for antialiased datashader_b200 lines yet")

    device = frame.device
    with torch.cuda.device(device):
        stream_ptr = torch.cuda.current_stream(device).cuda_stream
        x_range, y_range, view, x_st, y_st, (xs, ys, xy_dtype, nlines, nverts, layout) = _line_setup(frame, canvas, glyph, dist)
        H, W = canvas.plot_height, canvas.plot_width
        categorizer, ncat, labels = _categorical_setup([agg], schema)
        codes = categorizer.codes(frame).contiguous() if ncat else None
        shape = (H, W) + ((ncat,) if ncat else ())
        ncell = int(np.prod(shape))
        la = red._line_agg
        aa = line_width > 0
        val, val_dtype = None, _lib.NONE
        if red.column is not None:
            val = frame[red.column]
            val_dtype = _lib.dsb_dtype(frame.np_dtype(red.column))
        mask = None
        lib = _lib.lib()
        # accumulator canvases per dsb_lines_axis1's contract (include/dsb200.h)
        if la == _lib.LINE_ANY:
            canvas_t = torch.empty(shape, dtype=torch.int32 if aa else torch.uint8, device=device)
            _lib.check(lib.dsb_init_canvas(_lib.OP_MAX32 if aa else _lib.OP_ANY, canvas_t.data_ptr(), ncell, stream_ptr))
        elif la == _lib.LINE_COUNT:
            canvas_t = torch.zeros(shape, dtype=torch.float32 if aa else torch.int32, device=device)
            mask = torch.zeros(shape, dtype=torch.uint8, device=device) if aa else None
        elif la == _lib.LINE_SUM:
            canvas_t = torch.zeros(shape, dtype=torch.float64, device=device)
            mask = torch.zeros(shape, dtype=torch.uint8, device=device)
        elif la == _lib.LINE_MEAN:
            canvas_t = torch.zeros(shape, dtype=torch.float64, device=device)
            mask = torch.zeros(shape, dtype=torch.int32, device=device)        # the count canvas
        else:
            canvas_t = torch.empty(shape, dtype=torch.int64, device=device)
            _lib.check(lib.dsb_init_canvas(_lib.OP_MAX64 if la == _lib.LINE_MAX else _lib.OP_MIN64, canvas_t.data_ptr(),
                                           ncell, stream_ptr))
        if config.time_kernels:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        _lib.check(lib.dsb_lines_axis1_cat(C.byref(view), xs.data_ptr(), ys.data_ptr(), xy_dtype, nlines, nverts, C.byref(layout),
                                           val.data_ptr() if val is not None else None, val_dtype, la, line_width,
                                           canvas_t.data_ptr(), mask.data_ptr() if mask is not None else None,
                                           codes.data_ptr() if ncat else None,
                                           _lib.dsb_dtype(str(codes.dtype).replace("torch.", "")) if ncat else _lib.NONE, ncat,
                                           stream_ptr), "dsb_lines_axis1_cat")
        if config.time_kernels:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            config.kernel_events.append((ev0, ev1))
        if dist is not None:
            if la == _lib.LINE_MEAN:
                dist._all_reduce(canvas_t, "sum")
                dist._all_reduce(mask, "sum")
            else:
                canvas_t, mask = dist.combine_lines(la, aa, canvas_t, mask)
        # finishing (dtypes pinned by test_pandas.py:3257-3277: AA any/count -> f32, others f64)
        if la == _lib.LINE_ANY:
            if aa:
                out = torch.empty(shape, dtype=torch.float64, device=device)
                _lib.check(lib.dsb_decode_minmax(canvas_t.data_ptr(), _lib.OP_MAX32, _lib.F32, out.data_ptr(), ncell, stream_ptr))
                data = _to_host(out.to(torch.float32))
            else:
                data = _to_host(canvas_t, np.bool_)
        elif la == _lib.LINE_COUNT:
            if aa:
                out = torch.where(mask.bool(), canvas_t, torch.full_like(canvas_t, float("nan")))
                data = _to_host(out)
            else:
                data = _to_host(canvas_t, np.uint32)
        elif la == _lib.LINE_SUM:
            out = torch.empty_like(canvas_t)
            _lib.check(lib.dsb_finalize_sum(canvas_t.data_ptr(), mask.data_ptr(), out.data_ptr(), ncell, stream_ptr))
            data = _to_host(out)
        elif la == _lib.LINE_MEAN:
            out = torch.empty_like(canvas_t)
            _lib.check(lib.dsb_finalize_mean(canvas_t.data_ptr(), mask.data_ptr(), out.data_ptr(), ncell, stream_ptr))
            data = _to_host(out)
        else:
            out = torch.empty(shape, dtype=torch.float64, device=device)
            _lib.check(lib.dsb_decode_minmax(canvas_t.data_ptr(), _lib.OP_MAX64 if la == _lib.LINE_MAX else _lib.OP_MIN64,
                                             _lib.F64, out.data_ptr(), ncell, stream_ptr))
            data = _to_host(out)

    x_axis = canvas.x_axis.compute_index(x_st, canvas.plot_width)
    y_axis = canvas.y_axis.compute_index(y_st, canvas.plot_height)
    coords = {glyph.y_label: y_axis, glyph.x_label: x_axis}
    dims = [glyph.y_label, glyph.x_label]
    if ncat:                          # by._build_finalize, reductions.py:809-820
        coords[agg.cat_column] = list(labels)
        dims.append(agg.cat_column)
    return DataArray(data, coords=coords, dims=dims, attrs=dict(x_range=x_range, y_range=y_range))


def _aa2_combo(agg):
    """The reductions whose antialiased form needs the 2-stage combine (each reduction's _antialias_stage_2,
    reductions.py:545-549, 1041-1045, 1172-1173, 1395-1396, 1433-1434; antialias.py:30-58)."""
    if isinstance(agg, rd.min):
        return _lib.AA2_MIN
    if isinstance(agg, rd.first):
        return _lib.AA2_FIRST
    if isinstance(agg, rd.last):
        return _lib.AA2_LAST
    if isinstance(agg, rd.sum) and not agg.self_intersect:
        return _lib.AA2_SUM
    if isinstance(agg, rd.count) and not agg.self_intersect:
        return _lib.AA2_COUNT
    return None


_NAN_BITS = 0x7ff8000000000000          # numpy's / torch's default quiet NaN


def _lines_aa2(frame, canvas, glyph, agg, combo, line_width, dist, rows_only=False):
    """Antialiased lines, 2-stage reductions: dsb_lines_aa2 (one CTA per line, per-line max then the stage-2 fold).
    rows_only (first / last): return the voted global row canvas (+ ranges and scale / translate) instead of the values."""
    if dist is not None and glyph_per_vertex(glyph):
        # every rank holds a row shard of the SAME line(s): stage 1 (the per-line maximum coverage) would have to be
        # combined across ranks before stage 2, which the per-rank fold cannot do
        raise NotImplementedError("2-stage antialiased reductions over row-sharded per-vertex line glyphs (axis=0)")
    device = frame.device
    with torch.cuda.device(device):
        stream_ptr = torch.cuda.current_stream(device).cuda_stream
        x_range, y_range, view, x_st, y_st, (xs, ys, xy_dtype, nlines, nverts, layout) = _line_setup(frame, canvas, glyph, dist)
        H, W = canvas.plot_height, canvas.plot_width
        lib = _lib.lib()
        val, val_dtype = None, _lib.NONE
        if agg.column is not None:
            val = frame[agg.column]
            val_dtype = _lib.dsb_dtype(frame.np_dtype(agg.column))
        # scratch for the lines that overflow the shared-memory stage-1 table: a key64 canvas + a touched bitmap per CTA
        # (8.125 bytes per pixel), one CTA per SM if 1/4 of the free memory (<= 16 GiB) allows, plus the redo queue
        free, _total = torch.cuda.mem_get_info(device)
        per_cta = 8 * H * W + 4 * ((H * W + 31) // 32) + 512 * 512 * 4     # stage-1 canvas + bitmap + per-thread lists
        nctas = int(max(1, min(torch.cuda.get_device_properties(device).multi_processor_count, max(nlines, 1),
                               min(free // 4, 16 << 30) // per_cta)))
        scratch = torch.empty(nctas * per_cta + 4 * (nlines + 4), dtype=torch.uint8, device=device)
        row_offset = frame.row_offset if not glyph_per_vertex(glyph) else 0

        def launch(phase, out, aux):
            _lib.check(lib.dsb_lines_aa2(C.byref(view), xs.data_ptr(), ys.data_ptr(), xy_dtype, nlines, nverts, C.byref(layout),
                                         row_offset, val.data_ptr() if val is not None else None, val_dtype, combo, phase,
                                         line_width, out.data_ptr(), aux.data_ptr() if aux is not None else None,
                                         scratch.data_ptr(), scratch.numel(), stream_ptr), "dsb_lines_aa2")

        if combo in (_lib.AA2_SUM, _lib.AA2_COUNT):
            acc = torch.zeros((H, W), dtype=torch.float64 if combo == _lib.AA2_SUM else torch.float32, device=device)
            mask = torch.zeros((H, W), dtype=torch.uint8, device=device)
            launch(1, acc, mask)
            if dist is not None:
                dist._all_reduce(acc, "sum")
                dist._all_reduce(mask, "max")
            out = torch.where(mask.bool(), acc, torch.full_like(acc, float("nan")))
        elif combo == _lib.AA2_MIN:
            keys = torch.empty((H, W), dtype=torch.int64, device=device)
            _lib.check(lib.dsb_init_canvas(_lib.OP_MIN64, keys.data_ptr(), H * W, stream_ptr))
            launch(1, keys, None)
            if dist is not None:
                dist._all_reduce(keys, "min")
            out = torch.empty((H, W), dtype=torch.float64, device=device)
            _lib.check(lib.dsb_decode_minmax(keys.data_ptr(), _lib.OP_MIN64, _lib.F64, out.data_ptr(), H * W, stream_ptr))
        elif combo in (_lib.AA2_ARGMIN, _lib.AA2_ARGMAX):
            # where(min | max): {value key, line index} pairs in one rasterisation (phase 3, 128-bit compare-and-swap): the best
            # value, the lowest line index among the lines that reach it
            is_max = combo == _lib.AA2_ARGMAX
            i64 = torch.iinfo(torch.int64)
            pairs = torch.empty((H * W, 2), dtype=torch.int64, device=device)
            pairs[:, 0] = i64.min if is_max else i64.max
            pairs[:, 1] = i64.max
            launch(3, pairs, None)
            rows = pairs[:, 1].contiguous().view(H, W)
            if dist is not None:
                mine = pairs[:, 0].contiguous().view(H, W)
                keys = mine.clone()
                dist._all_reduce(keys, "max" if is_max else "min")
                rows = torch.where(mine == keys, rows, torch.full_like(rows, i64.max))
                dist._all_reduce(rows, "min")
            assert rows_only
            return rows, (x_range, y_range, x_st, y_st)
        else:
            # first / last in one rasterisation: {line index, value bits} pairs updated with a 128-bit compare-and-swap
            # (phase 3); the two-phase form (vote, then the winning line stores) rasterised every line twice
            first = combo == _lib.AA2_FIRST
            pairs = torch.empty((H * W, 2), dtype=torch.int64, device=device)
            pairs[:, 0] = torch.iinfo(torch.int64).max if first else -1
            pairs[:, 1] = _NAN_BITS
            launch(3, pairs, None)
            rows = pairs[:, 0].contiguous().view(H, W)
            mine = rows
            if dist is not None:
                rows = rows.clone()
                dist._all_reduce(rows, "min" if first else "max")
            if rows_only:
                return rows, (x_range, y_range, x_st, y_st)
            out = pairs[:, 1].contiguous().view(torch.float64).view(H, W)
            if dist is not None:      # exactly one rank owns each winning line: the others contribute 0 bits
                out = torch.where(mine == rows, out, torch.zeros((), dtype=torch.float64, device=device))
                out = dist.sum_bits_f64(torch.nan_to_num(out, nan=0.0), rows)
        data = _to_host(out)
    x_axis = canvas.x_axis.compute_index(x_st, canvas.plot_width)
    y_axis = canvas.y_axis.compute_index(y_st, canvas.plot_height)
    return DataArray(data, coords={glyph.y_label: y_axis, glyph.x_label: x_axis}, dims=[glyph.y_label, glyph.x_label],
                     attrs=dict(x_range=x_range, y_range=y_range))


def glyph_per_vertex(glyph):
    """axis=0 layouts hand append() the vertex row, not the line index (line.py:1104-1125, 1157-1180)."""
    from .glyphs import LineAxis0, LineAxis0Multi
    return isinstance(glyph, (LineAxis0, LineAxis0Multi))


def _lines_plan(frame, needed, schema, canvas, glyph, agg, dist):
    """line_width == 0: every reduction of the point path, applied per touched pixel with i = the row the
    reference passes to append (the line for axis=1 layouts, the segment's first vertex for axis=0)."""
    device = frame.device
    with torch.cuda.device(device):
        x_range, y_range, view, x_st, y_st, verts = _line_setup(frame, canvas, glyph, dist)
        reds, results, labels = _accumulate_and_finalize(frame, frame, needed, schema, view, canvas, glyph, agg, dist,
                                                         _launch_lines, ctx_extra={"line_vertices": verts})
    x_axis = canvas.x_axis.compute_index(x_st, canvas.plot_width)
    y_axis = canvas.y_axis.compute_index(y_st, canvas.plot_height)
    return _wrap(agg, reds, results, glyph, x_axis, y_axis, x_range, y_range, labels)


# ------------------------------------------------------------------------------------------ areas
def _launch_areas(view, chunk, glyph, accs, canv, ctx, categorizer, ncat):
    lib = _lib.lib()
    xs, ys0, ys1, xy_dtype, nlines, nverts, layout = ctx.area_vertices
    for plan, _keep in _plans(chunk, accs, canv, ctx, categorizer, ncat):
        _lib.check(lib.dsb_areas_plan(C.byref(view), xs.data_ptr(), ys0.data_ptr(), ys1.data_ptr() if ys1 is not None else None,
                                      xy_dtype, nlines, nverts, C.byref(layout), chunk.row_offset, C.byref(plan),
                                      ctx.stream_ptr), "dsb_areas_plan")


def areas(source, canvas, glyph, agg, dist=None):
    """bypixel for the area glyphs (Canvas.area, core.py:480-709)."""
    needed, frame, schema = _prepare(source, glyph, agg, canvas)
    frame = frame.resident(needed)
    if dist is not None and glyph.value_per_vertex:
        frame, frame.plot_start = dist.carry_last_row(frame, needed)     # row shards of one long curve (dask.py:244-266)
    device = frame.device
    with torch.cuda.device(device):
        stream_ptr = torch.cuda.current_stream(device).cuda_stream
        x_range = canvas.x_range or _auto_range(glyph._x_tensors(frame), stream_ptr, dist, device)
        if canvas.y_range:
            y_range = canvas.y_range
        else:
            lo, hi = _column_bounds(stream_ptr, glyph._y_tensors(frame))
            if dist is not None:
                lo, hi = dist.global_bounds(lo, hi, device)
            if glyph.y_bounds_include_zero():          # area.py:71-79
                lo, hi = (lo if lo < 0 else 0), (hi if hi > 0 else 0)
            y_range = maybe_expand_bounds((lo, hi))
        canvas.validate_ranges(x_range, y_range)
        view, x_st, y_st = make_view(canvas, x_range, y_range)
        xs, ys0, ys1, (xls, yls) = glyph.vertices(frame)
        xy_dtype = _lib.F32 if xs.dtype == torch.float32 else _lib.F64
        nlines, nverts = int(max(xs.shape[0], ys0.shape[0])), int(xs.shape[1])
        layout = _lib.LineLayout(int(xls), int(yls), int(glyph.value_per_vertex), int(getattr(frame, "plot_start", True)))
        if glyph.ragged:
            nlines, nverts = _ragged_layout(layout, glyph.ragged_starts(frame), [xs, ys0] + ([ys1] if ys1 is not None else [])), 2
        reds, results, labels = _accumulate_and_finalize(
            frame, frame, needed, schema, view, canvas, glyph, agg, dist, _launch_areas,
            ctx_extra={"area_vertices": (xs, ys0, ys1, xy_dtype, nlines, nverts, layout)})
    x_axis = canvas.x_axis.compute_index(x_st, canvas.plot_width)
    y_axis = canvas.y_axis.compute_index(y_st, canvas.plot_height)
    return _wrap(agg, reds, results, glyph, x_axis, y_axis, x_range, y_range, labels)
