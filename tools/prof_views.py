import sys, os, time, json
sys.path.insert(0, "/root/repo")
import torch
import datashader_b200 as ds
from datashader_b200 import config, _lib
n = 100_000_000
g = torch.Generator(device="cuda"); g.manual_seed(7)
x = torch.rand(n, generator=g, device="cuda") * 16
y = torch.rand(n, generator=g, device="cuda") * 16
v = torch.randn(n, generator=g, device="cuda")
frame = ds.DeviceFrame({"x": x, "y": y, "v": v})
views = [((float(ix), float(ix + 1)), (float(iy), float(iy + 1))) for iy in range(16) for ix in range(16)]
cvs = ds.Canvas(256, 256)
config.device_results = True
for name, agg in (("count", ds.count()), ("mean", ds.mean("v")), ("max", ds.max("v")), ("count", ds.count())):
    for _ in range(3):
        cvs.points_batch(frame, "x", "y", agg, views, grid=(16, 16))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    cvs.points_batch(frame, "x", "y", agg, views, grid=(16, 16))
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(name, "host ms", round((t1 - t0) * 1e3, 2), "wall ms", round((t2 - t0) * 1e3, 2), "gpu ms", round(e0.elapsed_time(e1), 2), _lib.lib().dsb_last_kernel())
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
cvs.points_batch(frame, "x", "y", ds.max("v"), views, grid=(16, 16)); torch.cuda.synchronize()
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
