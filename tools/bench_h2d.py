#!/usr/bin/env python
"""Host -> device ceiling of the box: every rank copies a pinned 1 GiB buffer to its GPU in a loop (no kernels), all ranks
at once.  The aggregate is what bench.py's `e2e` (pinned host columns -> Canvas.points) can reach at best at that N.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/bench_h2d.py
"""
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 1 << 30
host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
host.fill_(1)
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
for _ in range(3):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
reps = 20
t0 = time.perf_counter()
for _ in range(reps):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
mine = nbytes * reps / dt / 1e9
t = torch.tensor([mine, dt], device="cuda", dtype=torch.float64)
gbs = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(gbs, t)
else:
    gbs = [t]
if rank == 0:
    per = [float(g[0]) for g in gbs]
    slowest = max(float(g[1]) for g in gbs)
    topo = ""
    try:
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except Exception:  # noqa: BLE001
        pass
    print(json.dumps({"n_gpus": world, "h2d_gbs_per_rank": per, "h2d_gbs_aggregate": world * nbytes * reps / slowest / 1e9,
                      "points_per_s_ceiling_count_8B": world * nbytes * reps / slowest / 8 / 1e9,
                      "host_cpus": os.cpu_count()}))
    print(topo)
if world > 1:
    dist.destroy_process_group()
