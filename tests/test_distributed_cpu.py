"""Host-side logic of the multi-GPU combine on CPU: world_size-2 gloo process group, numpy-emulated partial
accumulator canvases (same encodings as libdsb200: order-preserving keys, INT64_MAX / -1 sentinels)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
I64_MAX = np.iinfo(np.int64).max


def key32(f):
    b = (np.asarray(f, dtype=np.float32) + np.float32(0)).view(np.int32)
    return b ^ ((b >> 31) & 0x7fffffff)


def unkey32(k):
    k = np.asarray(k, dtype=np.int32)
    return (k ^ ((k >> 31) & 0x7fffffff)).view(np.float32)


def _partials(rank, world, n=20_000, ncell=64):
    rng = np.random.default_rng(7)
    cell = rng.integers(0, ncell, n)
    v = np.round(rng.standard_normal(n), 1).astype(np.float32)
    other = rng.random(n)
    lo, hi = n * rank // world, n * (rank + 1) // world
    c, vv, oo, rows = cell[lo:hi], v[lo:hi], other[lo:hi], np.arange(lo, hi)
    count = np.bincount(c, minlength=ncell).astype(np.int32)
    s = np.bincount(c, weights=vv.astype(np.float64), minlength=ncell)
    mx = np.full(ncell, np.iinfo(np.int32).min, np.int32)
    np.maximum.at(mx, c, key32(vv))
    minrow = np.full(ncell, I64_MAX, np.int64)
    np.minimum.at(minrow, c, rows)
    packed = np.full(ncell, np.iinfo(np.int64).min, np.int64)          # argmax32: key << 32 | ~u32(row)
    p = (key32(vv).astype(np.int64) << 32) | ((~rows.astype(np.uint32)).astype(np.int64) & 0xFFFFFFFF)
    np.maximum.at(packed, c, p)
    return dict(cell=cell, v=v, other=other, lo=lo, hi=hi, count=count, sum=s, mx=mx, minrow=minrow, packed=packed)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from datashader_b200 import reductions as rd
    from datashader_b200.distributed import ShardGroup, arg_rows_from_parts, shard_bounds
    g = ShardGroup()
    P = _partials(rank, world)
    accs = [rd.Acc("count"), rd.Acc("sum", "v"), rd.Acc("max32", "v"), rd.Acc("minrow", None, "v")]
    canv = {a.key: torch.from_numpy(P[k].copy()) for a, k in zip(accs, ("count", "sum", "mx", "minrow"))}
    g.combine(accs, canv)
    # where(max(v)): value key all-reduce, then earliest global row among the winners
    packed = torch.from_numpy(P["packed"].copy())
    hi_local = packed >> 32
    hi_global = g._all_reduce(hi_local.clone(), "max")
    lo_bits = (packed & 0xFFFFFFFF)
    rows_local = torch.where(packed == torch.iinfo(torch.int64).min, torch.full_like(packed, -1),
                             (~lo_bits) & 0xFFFFFFFF)
    cand = arg_rows_from_parts(hi_local, hi_global, rows_local)
    g._all_reduce(cand, "min")
    rows = torch.where(cand == I64_MAX, torch.full_like(cand, -1), cand)
    # lookup gather: owner contributes the f64 bits
    out = torch.zeros(rows.shape, dtype=torch.float64)
    own = (rows >= P["lo"]) & (rows < P["hi"])
    out[own] = torch.from_numpy(P["other"])[rows[own]]
    out = g.sum_bits_f64(out, rows)
    lo, hi = g.global_bounds(float(P["v"][P["lo"]:P["hi"]].min()), float(P["v"][P["lo"]:P["hi"]].max()), "cpu")
    assert shard_bounds(10, 0, 2) == (0, 5) and shard_bounds(10, 1, 2) == (5, 10)
    if rank == 0:
        q.put({k: v.numpy() for k, v in canv.items()} | {"rows": rows.numpy(), "lookup": out.numpy(), "bounds": (lo, hi)})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gloo_world2_combine_matches_single_pass():
    world, port = 2, 29641
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = _partials(0, 1)
    from datashader_b200 import reductions as rd
    np.testing.assert_array_equal(res[rd.Acc("count").key], full["count"])
    np.testing.assert_allclose(res[rd.Acc("sum", "v").key], full["sum"], rtol=1e-12, atol=1e-12)
    np.testing.assert_array_equal(res[rd.Acc("max32", "v").key], full["mx"])
    np.testing.assert_array_equal(res[rd.Acc("minrow", None, "v").key], full["minrow"])
    # reference semantics: where(max(v)) picks the earliest row holding the maximum (reductions.py:1224, 2014-2016)
    cell, v, other = full["cell"], full["v"], full["other"]
    want_rows = np.full(64, -1, np.int64)
    for c in range(64):
        idx = np.nonzero(cell == c)[0]
        if len(idx):
            want_rows[c] = idx[np.argmax(v[idx])]
    np.testing.assert_array_equal(res["rows"], want_rows)
    np.testing.assert_array_equal(res["lookup"], other[want_rows])
    assert res["bounds"] == (float(v.min()), float(v.max()))
