"""Shared test vocabulary: golden-case names -> oracle specs / canvases."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
NCAT = 5

# name -> oracle spec (oracle/oracle.py tuple language); the same names key tests/golden/*.npz
SPECS = {
    "count": ("count",), "count_v32": ("count", "v32"), "any": ("any",), "any_v32": ("any", "v32"),
    "sum_v32": ("sum", "v32"), "sum_v64": ("sum", "v64"), "sum_vi": ("sum", "vi"),
    "mean_v32": ("mean", "v32"), "mean_v64": ("mean", "v64"),
    "min_v32": ("min", "v32"), "max_v32": ("max", "v32"), "min_v64": ("min", "v64"), "max_v64": ("max", "v64"),
    "max_vi": ("max", "vi"),
    "first_v32": ("first", "v32"), "last_v32": ("last", "v32"),
    "where_max_v32_other": ("where", ("max", "v32"), "other"), "where_min_v32_other": ("where", ("min", "v32"), "other"),
    "where_max_v32_row": ("where", ("max", "v32"), None), "where_min_v32_row": ("where", ("min", "v32"), None),
    "where_first_v32_other": ("where", ("first", "v32"), "other"),
    "where_last_v32_other": ("where", ("last", "v32"), "other"),
    "where_first_v32_row": ("where", ("first", "v32"), None), "where_last_v32_row": ("where", ("last", "v32"), None),
    "where_max_vi_other": ("where", ("max", "vi"), "other"),
    "by_count": ("by", "cat", ("count",)), "by_count_v32": ("by", "cat", ("count", "v32")),
    "by_sum_v32": ("by", "cat", ("sum", "v32")), "by_mean_v32": ("by", "cat", ("mean", "v32")),
    "by_max_v32": ("by", "cat", ("max", "v32")), "by_min_v32": ("by", "cat", ("min", "v32")),
    "by_any": ("by", "cat", ("any",)),
}

CANVASES = {
    "c2x2": dict(plot_width=2, plot_height=2, x_range=(0, 1), y_range=(0, 1)),
    "c37x23": dict(plot_width=37, plot_height=23, x_range=(-0.1, 1.05), y_range=(0.1, 0.9)),
    "c90x52": dict(plot_width=90, plot_height=52, x_range=(0, 1), y_range=(0, 1)),
    "cauto": dict(plot_width=31, plot_height=17),
}

LINE_CANVASES = {
    "c64x48": dict(plot_width=64, plot_height=48, x_range=(0, 1), y_range=(0, 1)),
    "c33x57": dict(plot_width=33, plot_height=57, x_range=(-0.3, 1.3), y_range=(-0.5, 1.5)),
}

# float reductions whose result depends on summation order: compared with a tolerance, everything
# else bit-exact (np.array_equal with NaNs equal).
FLOAT_SUM = ("sum", "mean")
RTOL_SUM = 1e-12   # f64 accumulation on both sides; only ordering differs


def is_float_sum(name):
    return any(k in name for k in FLOAT_SUM)


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def columns_from_golden(g, prefix):
    cols = {k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)}
    cols["cat__ncat"] = NCAT
    return cols


def assert_agg_equal(got, want, name, rtol=RTOL_SUM, atol=0):
    """Bit-exact except for float sums / means, whose f64 atomic adds run in a different order than the reference's
    loop: rtol (1e-12), plus `atol` where the values of a pixel may cancel to ~0 (stated by the caller)."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    assert got.dtype == want.dtype, (name, got.dtype, want.dtype)
    if is_float_sum(name):
        assert np.array_equal(np.isnan(got), np.isnan(want)), name
        np.testing.assert_allclose(got, want, rtol=rtol, atol=atol, equal_nan=True, err_msg=name)
    else:
        assert np.array_equal(got, want, equal_nan=(got.dtype.kind == "f")), name
        if got.dtype.kind == "f":
            # bit patterns, not just values: np.array_equal treats -0.0 == +0.0, the reference keeps the zero that
            # arrived first (NaN payloads are not compared)
            ok = ~np.isnan(want)
            bits = {4: np.uint32, 8: np.uint64}[got.dtype.itemsize]
            assert np.array_equal(np.ascontiguousarray(got)[ok].view(bits), np.ascontiguousarray(want)[ok].view(bits)), \
                f"{name}: sign of zero differs"


def make_agg(spec):
    """oracle spec tuple -> datashader_b200 reduction object"""
    import datashader_b200 as ds
    kind = spec[0]
    if kind == "by":
        return ds.by(spec[1], make_agg(spec[2]))
    if kind == "where":
        return ds.where(make_agg(spec[1]), spec[2])
    ctor = getattr(ds, kind)
    return ctor(*spec[1:])


def pandas_frame(cols, ncat=NCAT):
    import pandas as pd
    d = {k: v for k, v in cols.items() if not k.endswith("__ncat") and k != "cat"}
    if "cat" in cols:
        d["cat"] = pd.Categorical.from_codes(cols["cat"], categories=[f"c{i}" for i in range(int(cols.get("cat__ncat", ncat)))])
    return pd.DataFrame(d)
