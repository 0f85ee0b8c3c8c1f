"""The reference's own dispatcher driving the GPU: integration/b200.py registered with `bypixel.pipeline` of the reference
tree (oracle/_ref), called exactly as the reference's bypixel does (core.py:1334-1359), compared with the reference's pandas
backend on the same rows."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def test_reference_dispatcher_runs_on_libdsb200():
    if not os.path.isdir(os.path.join(REF, "datashader")):
        pytest.skip("oracle/_ref not staged")
    for p in (os.path.join(REF, "_shims"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import datashader as ds
    import pandas as pd
    import torch
    from datashader.glyphs import Point
    from datashader.utils import dshape_from_pandas
    import datashader_b200 as dsb
    from integration import b200
    b200.register(ds)
    rng = np.random.default_rng(21)
    n = 50_000
    df = pd.DataFrame({"x": rng.random(n, dtype=np.float32) * 1.2 - 0.1, "y": rng.random(n, dtype=np.float32) * 1.2 - 0.1,
                       "v": rng.standard_normal(n).astype(np.float32), "o": rng.random(n),
                       "c": pd.Categorical.from_codes(rng.integers(0, 3, n), categories=list("abc"))})
    df.loc[rng.integers(0, n, 500), "v"] = np.nan
    frame = dsb.DeviceFrame({"x": torch.from_numpy(df["x"].values).cuda(), "y": torch.from_numpy(df["y"].values).cuda(),
                             "v": torch.from_numpy(df["v"].values).cuda(), "o": torch.from_numpy(df["o"].values).cuda(),
                             "c": torch.from_numpy(df["c"].cat.codes.values.astype(np.int8)).cuda()},
                            categories={"c": list("abc")})
    cvs = ds.Canvas(plot_width=70, plot_height=45, x_range=(0, 1), y_range=(0, 1))
    glyph = Point("x", "y")
    schema = dshape_from_pandas(df)
    for agg in (ds.count(), ds.mean("v"), ds.max("v"), ds.where(ds.max("v"), "o"), ds.by("c", ds.count()), ds.first("v")):
        want = cvs.points(df, "x", "y", agg)                                       # the reference, pandas + numba
        got = ds.core.bypixel.pipeline(frame, schema, cvs, glyph, agg)            # the reference's dispatcher -> libdsb200
        assert type(got).__name__ == "DataArray" and list(got.dims) == list(want.dims), agg
        a, b = np.asarray(got.data), np.asarray(want.data)
        assert a.dtype == b.dtype and a.shape == b.shape, agg
        if type(agg).__name__ == "mean":
            np.testing.assert_allclose(a, b, rtol=1e-12, equal_nan=True)
        else:
            assert np.array_equal(a, b, equal_nan=a.dtype.kind == "f"), agg
        for d in want.dims[:2]:
            np.testing.assert_array_equal(np.asarray(got.coords[d]), np.asarray(want.coords[d]))
