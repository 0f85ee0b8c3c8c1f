"""Multi-rank parity worker (run under torchrun): every rank aggregates its contiguous row shard on its own
GPU, partial canvases are combined with NCCL all-reduces, and rank 0 checks the result against the
single-pass CPU oracle - the multi-GPU statement of tests/test_dask.py's npartitions sweep."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import datashader_b200 as ds
    from datashader_b200.distributed import shard_bounds
    from helpers import NCAT, SPECS, assert_agg_equal, make_agg
    from oracle import oracle as ora

    rng = np.random.default_rng(1234)
    n = 300_001
    cols = {"x": rng.random(n, dtype=np.float32), "y": rng.random(n, dtype=np.float32),
            "v32": np.round(rng.standard_normal(n), 1).astype(np.float32), "other": rng.random(n).astype(np.float32),
            "vi": rng.integers(-9, 9, n).astype(np.int32), "v64": np.round(rng.standard_normal(n), 1),
            "cat": rng.integers(0, NCAT, n).astype(np.int8), "cat__ncat": NCAT}
    cols["v32"][rng.integers(0, n, 300)] = np.nan
    lo, hi = shard_bounds(n, rank, world)
    frame = ds.DeviceFrame({k: torch.from_numpy(np.ascontiguousarray(v[lo:hi])).cuda() for k, v in cols.items()
                            if k != "cat__ncat"}, categories={"cat": [f"c{i}" for i in range(NCAT)]}, row_offset=lo)
    frame.sharded = True
    W, H = 97, 61
    cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
    names = ["count", "count_v32", "any", "sum_v32", "mean_v32", "max_v32", "min_v32", "max_vi", "max_v64", "first_v32",
             "last_v32", "where_max_v32_other", "where_min_v32_row", "where_max_vi_other", "where_first_v32_other",
             "where_last_v32_row", "by_count", "by_mean_v32", "by_max_v32"]
    bad = []
    for name in names + ["where_max_v64_other"]:
        spec = SPECS.get(name, ("where", ("max", "v64"), "other"))
        got = cvs.points(frame, "x", "y", make_agg(spec)).data
        if rank == 0:
            want = ora.points(cols, "x", "y", spec, view, npartitions=2 if ("first" in name or "last" in name) else 1)
            try:
                assert_agg_equal(got, want, f"{name} world={world}")
            except AssertionError as e:   # noqa: PERF203
                bad.append(f"{name}: {str(e)[:200]}")
    # the same shards as part of a frame whose global rows cross 2^32 INSIDE rank 1's shard: the packed {key32, row} accumulator
    # breaks ties on 32 row bits, so every rank must agree on the 64-bit form (reductions._rows_fit_packed32) - v32 is full of ties
    base = (1 << 32) - (n // 2) - 100
    shifted = ds.DeviceFrame(dict(frame.columns), categories=frame.categories, row_offset=base + lo)
    shifted.sharded = True
    for name in ("where_max_v32_other", "where_min_v32_row"):
        got = cvs.points(shifted, "x", "y", make_agg(SPECS[name])).data
        if rank == 0:
            if name.endswith("_row"):
                got = np.where(got >= 0, got - base, got)
            try:
                assert_agg_equal(got, ora.points(cols, "x", "y", SPECS[name], view), f"rows across 2^32 {name} world={world}")
            except AssertionError as e:
                bad.append(f"rows across 2^32 {name}: {str(e)[:200]}")
    # where(max | min) as two passes (config 5's form, forced on at this size): the key canvas is all-reduced between the passes,
    # so every rank matches its rows against the GLOBAL extreme, and the row canvas (min) after
    from datashader_b200 import _lib
    knobs = (ds.config.routed_min_rows, ds.config.l2_budget_bytes)
    ds.config.routed_min_rows, ds.config.l2_budget_bytes = 0, 1
    _lib.check(_lib.lib().dsb_routed_configure(0))
    for name in ("where_max_v32_other", "where_min_v32_row"):
        got = cvs.points(frame, "x", "y", make_agg(SPECS[name])).data
        if b"k_points_match32" not in _lib.lib().dsb_last_kernel():
            bad.append(f"two-pass {name}: took {_lib.lib().dsb_last_kernel()}")
        if rank == 0:
            try:
                assert_agg_equal(got, ora.points(cols, "x", "y", SPECS[name], view), f"two-pass {name} world={world}")
            except AssertionError as e:
                bad.append(f"two-pass {name}: {str(e)[:200]}")
    ds.config.routed_min_rows, ds.config.l2_budget_bytes = knobs
    _lib.check(_lib.lib().dsb_routed_configure(1 << 24))
    # LinesAxis1 sharded by line: each rank rasterises its lines, canvases all-reduced (sum / max / min per accumulator)
    nl, nv = 2001, 9
    lx = np.cumsum(rng.normal(0, 0.05, (nl, nv)), axis=1).astype(np.float32) + np.float32(0.5)
    ly = np.cumsum(rng.normal(0, 0.05, (nl, nv)), axis=1).astype(np.float32) + np.float32(0.5)
    lval = rng.normal(size=nl)
    llo, lhi = shard_bounds(nl, rank, world)
    lcols = {f"x{j}": lx[llo:lhi, j] for j in range(nv)}
    lcols.update({f"y{j}": ly[llo:lhi, j] for j in range(nv)})
    lcols["val"] = lval[llo:lhi]
    lframe = ds.DeviceFrame({k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in lcols.items()}, row_offset=llo)
    lframe.sharded = True
    xc, yc = [f"x{j}" for j in range(nv)], [f"y{j}" for j in range(nv)]
    lview = ora.make_view(W, H, (0.2, 0.9), (0.1, 0.8))
    lcvs = ds.Canvas(W, H, x_range=(0.2, 0.9), y_range=(0.1, 0.8))
    for name, agg, lw in [("any", ds.any(), 0), ("count", ds.count(), 0), ("sum", ds.sum("val"), 0), ("max", ds.max("val"), 0),
                          ("min", ds.min("val"), 0), ("max", ds.max("val"), 2.0), ("count", ds.count(), 1.0)]:
        got = lcvs.line(lframe, x=xc, y=yc, agg=agg, axis=1, line_width=lw).data
        if rank == 0:
            want = ora.lines_axis1(lx, ly, lview, name, None if name in ("any", "count") else lval, lw)
            ok = got.dtype == want.dtype and np.array_equal(np.isnan(got.astype("f8")), np.isnan(want.astype("f8")))
            ok = ok and np.allclose(got, want, rtol=1e-6, atol=1e-6, equal_nan=True)
            if lw == 0 and name != "sum":
                ok = ok and np.array_equal(got, want, equal_nan=got.dtype.kind == "f")
            if not ok:
                bad.append(f"lines {name} lw={lw}")
    # 2-stage antialiased reductions over lines sharded by line: min (key canvas all-reduced), first / last ({line, value}
    # pairs per rank, line canvas all-reduced, the owning rank contributes the value bits), sum(self_intersect=False); and
    # where(first) / where(max) row ids, which must be those of the single-rank run over all lines
    for name, agg in [("min", ds.min("val")), ("first", ds.first("val")), ("last", ds.last("val")),
                      ("sum", ds.sum("val", self_intersect=False))]:
        got = lcvs.line(lframe, x=xc, y=yc, agg=agg, axis=1, line_width=2.0).data
        if rank == 0:
            want = ora.lines_aa2(lx, ly, lview, name, lval, 2.0)
            ok = np.array_equal(np.isnan(got), np.isnan(want)) and np.allclose(got, want, rtol=1e-6, atol=1e-6, equal_nan=True)
            if not ok:
                bad.append(f"lines 2-stage antialiased {name}")
    full = ds.DeviceFrame({f"x{j}": torch.from_numpy(np.ascontiguousarray(lx[:, j])).cuda() for j in range(nv)}
                          | {f"y{j}": torch.from_numpy(np.ascontiguousarray(ly[:, j])).cuda() for j in range(nv)}
                          | {"val": torch.from_numpy(lval).cuda()})
    for name, agg in [("where(first)", ds.where(ds.first("val"))), ("where(last)", ds.where(ds.last("val"))),
                      ("where(max)", ds.where(ds.max("val"))), ("where(min)", ds.where(ds.min("val")))]:
        got = lcvs.line(lframe, x=xc, y=yc, agg=agg, axis=1, line_width=2.0).data
        want = lcvs.line(full, x=xc, y=yc, agg=agg, axis=1, line_width=2.0).data      # one rank, all the lines
        if rank == 0 and not np.array_equal(got, want):
            bad.append(f"lines antialiased {name}: sharded != single-rank row ids")
    # LineAxis0 / AreaToZeroAxis0: ONE long curve sharded by rows; each rank receives the previous shard's last vertex
    # (data_libraries/dask.py:244-266) so that the segment across the shard boundary is drawn exactly once
    nv0 = 4001
    x0 = np.cumsum(rng.normal(0, 0.01, nv0)).astype(np.float32) + np.float32(0.5)
    y0 = np.cumsum(rng.normal(0, 0.01, nv0)).astype(np.float32) + np.float32(0.5)
    y0[rng.integers(0, nv0, 15)] = np.nan
    y0[nv0 // 2 - 1] = np.nan                      # a break right at the 2-rank shard boundary
    v0 = rng.normal(size=nv0)
    alo, ahi = shard_bounds(nv0, rank, world)
    aframe = ds.DeviceFrame({"x": torch.from_numpy(x0[alo:ahi].copy()).cuda(), "y": torch.from_numpy(y0[alo:ahi].copy()).cuda(),
                             "val": torch.from_numpy(v0[alo:ahi].copy()).cuda()}, row_offset=alo)
    aframe.sharded = True
    for name, agg, lw in [("any", ds.any(), 0), ("count", ds.count(), 0), ("sum", ds.sum("val"), 0), ("max", ds.max("val"), 0),
                          ("max", ds.max("val"), 2.0)]:
        got = lcvs.line(aframe, "x", "y", agg=agg, line_width=lw).data
        if rank == 0:
            want = ora.lines(x0[None], y0[None], lview, name, None if name in ("any", "count") else v0, lw, per_vertex=True)
            ok = got.dtype == want.dtype and np.array_equal(np.isnan(got.astype("f8")), np.isnan(want.astype("f8")))
            ok = ok and np.allclose(got, want, rtol=1e-6, atol=1e-6, equal_nan=True)
            if lw == 0 and name != "sum":
                ok = ok and np.array_equal(got, want, equal_nan=got.dtype.kind == "f")
            if not ok:
                bad.append(f"line axis0 sharded {name} lw={lw}")
    for name, agg in [("any", ds.any()), ("count", ds.count()), ("max", ds.max("val"))]:
        got = lcvs.area(aframe, "x", "y", agg=agg).data
        if rank == 0:
            want = ora.areas(x0[None], y0[None], lview, None, name, None if name in ("any", "count") else v0, per_vertex=True)
            if not (got.dtype == want.dtype and np.array_equal(got, want, equal_nan=got.dtype.kind == "f")):
                bad.append(f"area axis0 sharded {name}")
    got = lcvs.line(aframe, "x", "y", agg=ds.first("val")).data      # global row ids survive the prepended vertex
    if rank == 0:
        full = ds.DeviceFrame({"x": torch.from_numpy(x0).cuda(), "y": torch.from_numpy(y0).cuda(), "val": torch.from_numpy(v0).cuda()})
    whole = lcvs.line(full, "x", "y", agg=ds.first("val")).data if rank == 0 else None
    if rank == 0 and not np.array_equal(got, whole, equal_nan=True):
        bad.append("line axis0 sharded first")
    # LinesAxis1Ragged / AreaToZeroAxis1Ragged sharded by row: every rank holds whole rows (its slice of the start indices,
    # re-based) and the global row offset; Bresenham / fill / antialiased max / 2-stage first against the oracle's row loop,
    # where(first) row ids against the single-rank run
    nr = 801
    rlen = rng.integers(0, 14, nr)
    rst = (np.cumsum(rlen) - rlen).astype(np.int64)
    rx = (np.cumsum(rng.normal(0, 0.04, int(rlen.sum()))) % 1.0).astype(np.float32)
    ry = rng.random(int(rlen.sum())).astype(np.float32)
    rval = rng.normal(size=nr)
    rlo, rhi = shard_bounds(nr, rank, world)
    flo, fhi = int(rst[rlo]) if rlo < nr else len(rx), (int(rst[rhi]) if rhi < nr else len(rx))

    def rcol(flat, a, b, fa, fb):
        return ds.RaggedColumn(torch.from_numpy(flat[fa:fb].copy()).cuda(), torch.from_numpy(rst[a:b] - fa).cuda())
    rframe = ds.DeviceFrame({"x": rcol(rx, rlo, rhi, flo, fhi), "y": rcol(ry, rlo, rhi, flo, fhi),
                             "val": torch.from_numpy(rval[rlo:rhi].copy()).cuda()}, row_offset=rlo)
    rframe.sharded = True
    rfull = ds.DeviceFrame({"x": rcol(rx, 0, nr, 0, len(rx)), "y": rcol(ry, 0, nr, 0, len(ry)), "val": torch.from_numpy(rval).cuda()})
    for name, agg, lw in [("count", ds.count(), 0), ("max", ds.max("val"), 0), ("max", ds.max("val"), 2.0), ("first", ds.first("val"), 2.0)]:
        got = lcvs.line(rframe, "x", "y", agg=agg, axis=1, line_width=lw).data
        if rank == 0:
            want = (ora.lines_ragged_aa2(rx, rst, ry, rst, lview, name, rval, lw) if name == "first"
                    else ora.lines_ragged(rx, rst, ry, rst, lview, name, None if name == "count" else rval, lw))
            ok = got.dtype == want.dtype and np.array_equal(np.isnan(got.astype("f8")), np.isnan(want.astype("f8")))
            ok = ok and np.allclose(got, want, rtol=1e-6, atol=1e-6, equal_nan=True)
            if lw == 0:
                ok = ok and np.array_equal(got, want, equal_nan=got.dtype.kind == "f")
            if not ok:
                bad.append(f"ragged lines sharded {name} lw={lw}")
    got = lcvs.area(rframe, "x", "y", agg=ds.count(), axis=1).data
    if rank == 0 and not np.array_equal(got, ora.areas_ragged(rx, rst, ry, rst, lview, None, None, "count", None)):
        bad.append("ragged area sharded count")
    got = lcvs.line(rframe, "x", "y", agg=ds.where(ds.first("val")), axis=1, line_width=2.0).data
    want = lcvs.line(rfull, "x", "y", agg=ds.where(ds.first("val")), axis=1, line_width=2.0).data
    if rank == 0 and not np.array_equal(got, want):
        bad.append("ragged lines antialiased where(first): sharded != single-rank row ids")
    # auto-ranging across shards
    got = ds.Canvas(31, 17).points(frame, "x", "y").data
    if rank == 0:
        v2 = ora.make_view(31, 17, ora.compute_bounds(cols["x"]), ora.compute_bounds(cols["y"]))
        if not np.array_equal(got, ora.points(cols, "x", "y", ("count",), v2)):
            bad.append("auto-range count")
        print("DIST_PARITY " + ("OK" if not bad else "FAIL " + "; ".join(bad)), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if bad:
        sys.exit(1)


if __name__ == "__main__":
    main()
