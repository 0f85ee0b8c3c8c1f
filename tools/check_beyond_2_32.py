#!/usr/bin/env python
"""A resident frame with MORE than 2^32 rows on one B200 (4.5e9 float32 points, 54 GB of columns): Canvas.points walks it in
slices of 2^31 rows (DeviceFrame.chunks).  Checks that need no oracle: count conserves the rows, equals the sum of the two
slices aggregated separately; where(last) reports global row ids beyond 2^32 whose value is last('value'); where(max) picks
a row holding the pixel's max; first / last / max equal the combination of the two slices.
    python tools/check_beyond_2_32.py [n=4.5e9]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import datashader_b200 as ds

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_500_000_000
cut = 1 << 32
assert n > cut
g = torch.Generator(device="cuda")
g.manual_seed(7)
x = torch.rand(n, generator=g, device="cuda")
y = torch.rand(n, generator=g, device="cuda")
v = torch.randn(n, generator=g, device="cuda")
frame = ds.DeviceFrame({"x": x, "y": y, "value": v})
assert frame.n_chunks() == -(-n // ds.DeviceFrame.CHUNK_ROWS) >= 3
lo = ds.DeviceFrame({"x": x[:cut], "y": y[:cut], "value": v[:cut]})
hi = ds.DeviceFrame({"x": x[cut:], "y": y[cut:], "value": v[cut:]}, row_offset=cut)
assert lo.n_chunks() == 2 and hi.n_chunks() == 1      # 2^32 rows are one too many for a single call; 2e8 are not
cvs = ds.Canvas(900, 525, x_range=(0.0, 1.0), y_range=(0.0, 1.0))
run = lambda f, agg: cvs.points(f, "x", "y", agg).data   # noqa: E731
ok = True


def check(name, cond):
    global ok
    print(f"{name}: {'ok' if cond else 'FAILED'}", flush=True)
    ok = ok and bool(cond)


torch.cuda.synchronize()
t0 = time.perf_counter()
c = run(frame, ds.count())
torch.cuda.synchronize()
print(f"count over {n:.3e} rows: {1e3 * (time.perf_counter() - t0):.1f} ms (first call)")
check("count conserves the rows", int(c.sum(dtype=np.int64)) == n)
check("count == slice 0 + slice 1", np.array_equal(c, run(lo, ds.count()) + run(hi, ds.count())))
mx = run(frame, ds.max("value"))
check("max == max of the slices", np.array_equal(mx, np.fmax(run(lo, ds.max("value")), run(hi, ds.max("value")))))
rows = run(frame, ds.where(ds.max("value")))
check("where(max) rows hold the max", np.array_equal(v[torch.from_numpy(rows).cuda()].cpu().numpy(), mx))
first, last = run(frame, ds.first("value")), run(frame, ds.last("value"))
check("first == first of slice 0", np.array_equal(first, run(lo, ds.first("value"))))
check("last == last of slice 1", np.array_equal(last, run(hi, ds.last("value"))))
lrows = run(frame, ds.where(ds.last("value")))
check("where(last) rows lie beyond 2^32", int(lrows.min()) >= cut and int(lrows.max()) < n)
check("where(last) rows hold last('value')", np.array_equal(v[torch.from_numpy(lrows).cuda()].cpu().numpy(), last))
m = run(frame, ds.mean("value"))
s0, s1 = run(lo, ds.sum("value")), run(hi, ds.sum("value"))
check("mean == (sum 0 + sum 1) / count to 1e-12", np.allclose(m, (s0 + s1) / c, rtol=1e-12, atol=1e-15))
from datashader_b200 import _lib
for name, agg in (("count", ds.count()), ("max", ds.max("value")), ("first", ds.first("value")), ("last", ds.last("value"))):
    run(frame, agg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(frame, agg)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{name}: {1e3 * dt:.1f} ms = {n / dt / 1e9:.0f} Gpts/s (host clock, whole call)  [{_lib.lib().dsb_last_kernel().decode()[:70]}]")
print("beyond 2^32:", "all ok" if ok else "FAILED")
sys.exit(0 if ok else 1)
