#!/usr/bin/env python
"""End-to-end ingestion: the same mean('value') aggregation from a pageable pandas.DataFrame, a pyarrow.Table, a pinned
HostFrame and a resident DeviceFrame.    python tools/bench_ingest.py [--n 200000000]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pandas as pd
import torch
import datashader_b200 as ds


def wall(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=200_000_000); a = ap.parse_args()
    n = a.n
    rng = np.random.default_rng(0)
    cols = {"x": rng.random(n, dtype=np.float32), "y": rng.random(n, dtype=np.float32), "value": rng.standard_normal(n, dtype=np.float32)}
    df = pd.DataFrame(cols)
    cvs = ds.Canvas(900, 525, x_range=(0, 1), y_range=(0, 1))
    agg = ds.mean("value")
    out = {"n": n, "bytes": 12 * n}
    t = wall(lambda: cvs.points(df, "x", "y", agg)); out["pandas_pageable"] = {"s": t, "gpts": n / t / 1e9, "GBps": 12 * n / t / 1e9}
    try:
        import pyarrow as pa
        tab = pa.Table.from_pandas(df)
        t = wall(lambda: cvs.points(tab, "x", "y", agg)); out["pyarrow_table"] = {"s": t, "gpts": n / t / 1e9, "GBps": 12 * n / t / 1e9}
    except ImportError:
        pass
    hf = ds.HostFrame({k: torch.from_numpy(v).pin_memory() for k, v in cols.items()})
    t = wall(lambda: cvs.points(hf, "x", "y", agg)); out["hostframe_pinned"] = {"s": t, "gpts": n / t / 1e9, "GBps": 12 * n / t / 1e9}
    dfr = ds.DeviceFrame({k: torch.from_numpy(v).cuda() for k, v in cols.items()})
    t = wall(lambda: cvs.points(dfr, "x", "y", agg)); out["deviceframe"] = {"s": t, "gpts": n / t / 1e9}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
