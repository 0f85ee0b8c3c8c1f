#!/usr/bin/env python
"""Benchmark of the hot path: Canvas(900x525).points over 1e9 float32 points per GPU, agg=mean('value')
(BASELINE.json configs[1]), plus the count() variant of the same pass.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference ...                      # the reference's CPU algorithm (oracle port,
                                                              #   all host threads) on a bounded sample

Prints ONE JSON line (rank 0).  `value` = Gpoints/s with the columns resident in HBM; `e2e` = the same
call fed from pinned HOST columns (H2D of every column + D2H of the aggregate inside the timed
region); `roofline` = algorithmic bytes (12 B/point for mean) / CUDA-event time of the
fused aggregation kernel against the measured HBM peak; `cpu_baseline` = the oracle port timed on
this box's host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H = 900, 525
BYTES_PER_POINT = {"mean": 12, "count": 8}
SEED = 20240917


def ncu_traffic_gb(workload, n):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu capture
    (profiles/r01_traffic.json, one `ncu --set full` launch at the same n); None when no capture matches."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        with open(p) as f:
            t = json.load(f)
        e = t.get(f"{workload}:{n}")
        return e["dram_gb"] if e else None
    except Exception:  # noqa: BLE001
        return None


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:  # noqa: BLE001
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def make_device_columns(n, rank, device):
    """uniform x, y in [0,1) f32; value ~ N(0,1) f32 with 0.1 % NaN, generated on the device."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(SEED + rank)
    x = torch.rand(n, generator=g, device=device, dtype=torch.float32)
    y = torch.rand(n, generator=g, device=device, dtype=torch.float32)
    v = torch.empty(n, device=device, dtype=torch.float32)
    step = 1 << 27
    for lo in range(0, n, step):      # chunked: randn + mask temporaries stay small
        hi = min(n, lo + step)
        v[lo:hi].normal_(generator=g)
        m = torch.rand(hi - lo, generator=g, device=device) < 1e-3
        v[lo:hi].masked_fill_(m, float("nan"))
        del m
    return x, y, v


def cpu_baseline(sample_n, workload, threads=None):
    """The oracle port (the reference's numba algorithm restated in C) on the host cores: `threads` row
    partitions aggregated into private canvases and combined, like dask's threaded scheduler."""
    import numpy as np
    from oracle import oracle as ora
    threads = threads or os.cpu_count() or 1
    rng = np.random.default_rng(SEED)
    cols = {"x": rng.random(sample_n, dtype=np.float32), "y": rng.random(sample_n, dtype=np.float32)}
    v = rng.standard_normal(sample_n, dtype=np.float32)
    v[rng.integers(0, sample_n, sample_n // 1000)] = np.nan
    cols["value"] = v
    view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
    spec = ("mean", "value") if workload == "mean" else ("count",)
    ora.points_mt(cols, "x", "y", spec, view, threads)      # warm-up (page faults, thread start)
    best = float("inf")
    for _ in range(3):
        t0 = time.perf_counter()
        ora.points_mt(cols, "x", "y", spec, view, threads)
        best = min(best, time.perf_counter() - t0)
    return sample_n / best / 1e9, best, threads


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for this path (oracle port; the reference is pure
    Python + numba and cannot travel to the GPU box), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as ora
    threads = os.cpu_count() or 1
    sample_n = args.cpu_sample
    rng = np.random.default_rng(SEED)
    cols = {"x": rng.random(sample_n, dtype=np.float32), "y": rng.random(sample_n, dtype=np.float32)}
    v = rng.standard_normal(sample_n, dtype=np.float32)
    v[rng.integers(0, sample_n, sample_n // 1000)] = np.nan
    cols["value"] = v
    view = ora.make_view(W, H, (0.0, 1.0), (0.0, 1.0))
    spec = ("mean", "value") if args.workload == "mean" else ("count",)
    for _ in range(args.warmup):
        ora.points_mt(cols, "x", "y", spec, view, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ora.points_mt(cols, "x", "y", spec, view, threads)
    dt = time.perf_counter() - t0
    val = sample_n * args.steps / dt / 1e9
    sample = f"{sample_n:.0e} of the {args.n:.0e} points per step (same generator), {threads} threads"
    line = {
        "impl": "reference", "metric": f"Canvas.points Gpoints/s ({args.workload}('value'), 900x525)", "value": val,
        "unit": "Gpoints/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"Canvas(900x525).points, agg={args.workload}('value'), float32 x/y/value, "
                               f"bounded CPU sample of {sample_n} rows per step", "points_per_gpu": args.n},
        "cpu_baseline": {"value": val, "unit": "Gpoints/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (datashader_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    import datashader_b200 as ds
    from datashader_b200 import _lib, config

    n = args.n
    x, y, v = make_device_columns(n, rank, device)
    frame = ds.DeviceFrame({"x": x, "y": y, "value": v}, row_offset=rank * n)
    frame.sharded = world > 1
    cvs = ds.Canvas(W, H, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
    agg = ds.mean("value") if args.workload == "mean" else ds.count()
    config.device_results = True

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return cvs.points(frame, "x", "y", agg)

    for _ in range(args.warmup):
        out = step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    config.time_kernels = True
    config.kernel_events.clear()
    launches0 = _lib.lib().dsb_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    barrier()
    config.time_kernels = False
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    launches = _lib.lib().dsb_launch_count() - launches0
    kernel_ms = sum(a.elapsed_time(b) for a, b in config.kernel_events) / max(1, len(config.kernel_events))
    t = torch.tensor([ms, kernel_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, kernel_ms = t.tolist()
    value = world * n * args.steps / (ms * 1e-3) / 1e9
    def total_u32(t):
        return int(t.view(torch.int32).to(torch.int64).sum().item())

    checksum = float(torch.nan_to_num(out.data.double()).sum().item()) if args.workload == "mean" else total_u32(out.data)

    # ---- the count() variant of the same pass (the other half of the north-star target), device-resident
    also = {}
    if args.also_count and args.workload == "mean":
        cagg = ds.count()
        for _ in range(2):
            cvs.points(frame, "x", "y", cagg)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(args.steps):
            cout = cvs.points(frame, "x", "y", cagg)
        c1.record()
        barrier()
        cms = torch.tensor([c0.elapsed_time(c1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(cms, op=dist.ReduceOp.MAX)
        cg = world * n * args.steps / (cms.item() * 1e-3) / 1e9
        peak, _ = measured_hbm_peak()
        also = {"count_gpoints_per_s": cg, "count_ms_per_step": cms.item() / args.steps,
                "count_hbm_frac": cg / world * 8 / peak, "count_total": total_u32(cout.data)}

    # ---- e2e: the same call fed from pinned host columns (H2D + D2H inside the timed region)
    config.device_results = False
    e2e = None
    if not args.no_e2e:
        n_e2e = min(n, args.e2e_n)
        host = {}
        for name, t_ in (("x", x), ("y", y), ("value", v)):
            h = torch.empty(n_e2e, dtype=torch.float32, pin_memory=True)
            h.copy_(t_[:n_e2e])
            host[name] = h
        torch.cuda.synchronize()
        hframe = ds.HostFrame(host, row_offset=rank * n_e2e, device=device)
        hframe.sharded = world > 1
        cvs.points(hframe, "x", "y", agg)     # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            res = cvs.points(hframe, "x", "y", agg)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n_e2e * args.e2e_steps / dt.item() / 1e9, "unit": "Gpoints/s",
               "h2d_bytes_per_step": int(n_e2e * BYTES_PER_POINT[args.workload]),
               "d2h_bytes_per_step": int(res.data.nbytes), "points_per_gpu_per_step": n_e2e,
               "ms_per_step": dt.item() / args.e2e_steps * 1e3}
        del host, hframe

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        bpp = BYTES_PER_POINT[args.workload]
        achieved = n * bpp / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else None
        cb = None
        if not args.no_cpu:
            cv, csec, cthreads = cpu_baseline(args.cpu_sample, args.workload)
            cb = {"value": cv, "unit": "Gpoints/s", "cores": cthreads, "kind": "port",
                  "sample": f"{args.cpu_sample:.0e} of the {n:.0e} points (same distribution), best of 3, {csec:.2f} s per pass"}
        line = {
            "metric": f"Canvas.points Gpoints/s ({args.workload}('value'), 900x525)" if args.workload == "mean"
                      else "Canvas.points count Gpoints/s (900x525)",
            "value": value, "unit": "Gpoints/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Canvas(900x525).points, {n:.0e} float32 points per GPU, agg={args.workload}"
                                   + ("('value')" if args.workload == "mean" else "()")
                                   + " (BASELINE.json configs[1]), uniform x/y in [0,1), 0.1% NaN values",
                       "points_per_gpu": n, "canvas": [W, H],
                       "l2": "inputs (12 GB per GPU) are far larger than the 126 MB L2; no flush needed",
                       "combine": "NCCL all-reduce of the f64 sum and u32 count canvases" if world > 1 else "none (1 GPU)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": ncu_traffic_gb(args.workload, n),
                         "kernel": ("k_points_priv_tight<4,mean> (K2: count privatised in shared memory, f64 sum via global RED)"
                                    if args.workload == "mean" else "k_points_priv_tight<3,count> (K2)"),
                         "kernel_ms": kernel_ms, "peak_source": peak_src,
                         "algorithmic_bytes_per_point": bpp, "traffic_unit": "GB per launch (ncu dram read+write)"},
            "cpu_baseline": cb,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "checksum": checksum,
            "also": also,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mean", choices=["mean", "count"])
    ap.add_argument("--n", type=float, default=1e9, help="points per GPU")
    ap.add_argument("--cpu-sample", type=float, default=1e8, help="rows of the CPU baseline sample")
    ap.add_argument("--e2e-n", type=float, default=1e9)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-also-count", dest="also_count", action="store_false")
    args = ap.parse_args()
    args.n, args.cpu_sample, args.e2e_n = int(args.n), int(args.cpu_sample), int(args.e2e_n)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
