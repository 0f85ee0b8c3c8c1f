"""tf.shade on the GPU against the real reference's images (tests/golden/shade.npz).
eq_hist and linear: bit-exact.  log / cbrt: CUDA's log1p / pow are not bit-identical to numpy's, so the
alpha (or colour) byte may differ by one level on a handful of pixels - tolerance written below."""
import numpy as np
import pytest

from helpers import load

pytestmark = pytest.mark.gpu

HOT = ["black", "darkred", "red", "orange", "yellow", "white"]


def _channels(img):
    return np.ascontiguousarray(img).view(np.uint8).reshape(img.shape + (4,)).astype(np.int16)


def _check(got, want, how, key, frac=0.02):
    assert got.dtype == np.uint32 and got.shape == want.shape, key
    if how in ("eq_hist", "linear"):
        np.testing.assert_array_equal(got, want, err_msg=key)
    else:
        diff = np.abs(_channels(got) - _channels(want))
        assert diff.max() <= 1, key
        assert (diff > 0).mean() < frac, key


def _cat_agg(ds, data):
    from datashader_b200.xr_compat import DataArray
    H, W, C = data.shape
    return DataArray(data, coords={"y": np.arange(H), "x": np.arange(W), "cat": [f"c{i}" for i in range(C)]},
                     dims=["y", "x", "cat"])


def _agg2(data):
    from datashader_b200.xr_compat import DataArray
    H, W = data.shape
    return DataArray(data, coords={"y": np.arange(H), "x": np.arange(W)}, dims=["y", "x"])


@pytest.mark.parametrize("name", ["poisson5", "pareto16", "dense3", "single4"])
def test_shade_categorical_golden(name):
    import datashader_b200 as ds
    g = load("shade.npz")
    data = g[f"cat_{name}_in"]
    agg = _cat_agg(ds, data)
    for how in ("eq_hist", "log", "cbrt", "linear"):
        _check(ds.tf.shade(agg, how=how).data, g[f"cat_{name}_{how}"], how, f"{name} {how}")
    _check(ds.tf.shade(agg, how="eq_hist", alpha=200, min_alpha=10).data, g[f"cat_{name}_eq_hist_a200_m10"], "eq_hist", name)
    _check(ds.tf.shade(agg, how="eq_hist", rescale_discrete_levels=True).data, g[f"cat_{name}_eq_hist_rescale"], "eq_hist", name)


@pytest.mark.parametrize("name", ["u32", "u32big", "f64", "f32"])
def test_shade_2d_golden(name):
    import datashader_b200 as ds
    g = load("shade.npz")
    agg = _agg2(g[f"d2_{name}_in"])
    for how in ("eq_hist", "log", "cbrt", "linear"):
        _check(ds.tf.shade(agg, how=how).data, g[f"d2_{name}_{how}_default"], how, f"{name} {how} default")
        _check(ds.tf.shade(agg, cmap=HOT, how=how).data, g[f"d2_{name}_{how}_hot"], how, f"{name} {how} hot")
        _check(ds.tf.shade(agg, cmap="#3070c0", how=how, min_alpha=20).data, g[f"d2_{name}_{how}_single"], how,
               f"{name} {how} single")


def test_shade_2d_f32_linear_log():
    """float32 canvases (antialiased any/count): offset subtraction happens in float32 like the reference."""
    import datashader_b200 as ds
    g = load("shade.npz")
    agg = _agg2(g["d2_f32_in"])
    for how in ("linear", "log", "cbrt"):
        _check(ds.tf.shade(agg, how=how).data, g[f"d2_f32_{how}_default"], how, f"f32 {how}")


def test_shade_reference_literal_table():
    """datashader/tests/test_transfer_functions.py:22-37, 109-111."""
    import datashader_b200 as ds
    a = np.arange(10, 19, dtype="u4").reshape((3, 3))
    a[[0, 1, 2], [0, 1, 2]] = 0
    sol = np.array([[0, 4291543295, 4288846335], [4286149631, 0, 4283518207], [4280821503, 4278190335, 0]], dtype="u4")
    np.testing.assert_array_equal(ds.tf.shade(_agg2(a), cmap=["pink", "red"], how="eq_hist").data, sol)


def test_points_by_then_shade_pipeline():
    """BASELINE config 3 in miniature: by('cat', count()) -> tf.shade(how='eq_hist') against oracle + shade oracle."""
    import pandas as pd
    import datashader_b200 as ds
    from oracle import oracle as ora, shade_oracle as so
    rng = np.random.default_rng(9)
    n, C = 400_000, 16
    cols = {"x": rng.normal(0.5, 0.15, n).astype(np.float32), "y": rng.normal(0.5, 0.15, n).astype(np.float32),
            "cat": rng.integers(0, C, n).astype(np.int8), "cat__ncat": C}
    df = pd.DataFrame({"x": cols["x"], "y": cols["y"]})
    df["cat"] = pd.Categorical.from_codes(cols["cat"], categories=[f"c{i}" for i in range(C)])
    agg = ds.Canvas(192, 108, x_range=(0.0, 1.0), y_range=(0.0, 1.0)).points(df, "x", "y", ds.by("cat", ds.count()))
    want_agg = ora.points(cols, "x", "y", ("by", "cat", ("count",)), ora.make_view(192, 108, (0.0, 1.0), (0.0, 1.0)))
    np.testing.assert_array_equal(agg.data, want_agg)
    img = ds.tf.shade(agg, how="eq_hist")
    colors = [ds.palette.rgb(c) for c in ds.palette.Sets1to3[:C]]
    np.testing.assert_array_equal(img.data, so.shade_categorical(want_agg, colors, how="eq_hist"))


SPAN_CASES_2D = {"u32": [(2, 9), (0.5, 7.5)], "f64": [(-5.0, 12.5)], "f32": [(0.2, 0.7)]}
SPAN_CASES_CAT = {"poisson5": [(3, 20), (2.5, 15.5)], "dense3": [(40, 90), (30.5, 100.25)]}


def test_shade_span_golden():
    """tf.shade(span=...) vs the reference: data clipped to the span, fixed normalisation range; eq_hist + span raises."""
    import datashader_b200 as ds
    g, gs = load("shade.npz"), load("shade_span.npz")
    for name, spans in SPAN_CASES_2D.items():
        agg = _agg2(g[f"d2_{name}_in"])
        for k, span in enumerate(spans):
            for how in ("log", "cbrt", "linear"):
                # float32 canvases: every pixel clipped to the upper bound carries the same value f(span[1] - span[0]),
                # which numpy evaluates with glibc's float32 log1p / pow (not correctly rounded: 1 ulp high here) and
                # the kernel with the correctly rounded value - one alpha level, on the whole clipped plateau
                frac = 0.08 if name == "f32" else 0.02
                _check(ds.tf.shade(agg, how=how, span=span).data, gs[f"d2_{name}_s{k}_{how}_default"], how, f"{name} {span} {how}", frac)
                _check(ds.tf.shade(agg, cmap="#3070c0", how=how, span=list(span), min_alpha=20).data,
                       gs[f"d2_{name}_s{k}_{how}_single"], how, f"{name} {span} {how} single", frac)
    for name, spans in SPAN_CASES_CAT.items():
        agg = _cat_agg(ds, g[f"cat_{name}_in"])
        for k, span in enumerate(spans):
            for how in ("log", "cbrt", "linear"):
                _check(ds.tf.shade(agg, how=how, span=span).data, gs[f"cat_{name}_s{k}_{how}"], how, f"{name} {span} {how}")
    with pytest.raises(ValueError, match="span is not"):
        ds.tf.shade(_agg2(g["d2_u32_in"]), how="eq_hist", span=(1, 5))


@pytest.mark.parametrize("name", ["f64", "f32", "i32"])
def test_shade_categorical_float_golden(name):
    """by(cat, mean | sum | max ...) aggregates - float [H, W, C] with NaN for empty cells, and signed integers - through
    tf.shade (_colorize :382-452) vs the real reference (tests/golden/shade_catfloat.npz).  Alpha: exact for eq_hist and
    linear, +-1 level for log / cbrt like the 2-D path.  Colour bytes: +-1 level on < 2 % of the pixels (the float32 sums
    over the category axis run in torch's order, numpy's are pairwise; a quotient that sits on an integer can flip)."""
    import datashader_b200 as ds
    g = load("shade_catfloat.npz")
    agg = _cat_agg(ds, g[f"catf_{name}_in"])
    cases = [(how, {}, f"catf_{name}_{how}") for how in ("eq_hist", "log", "cbrt", "linear")]
    cases.append(("linear", dict(color_baseline=0.5 if name != "i32" else 2), f"catf_{name}_linear_base"))
    cases.append(("linear", dict(span=(0, 20)), f"catf_{name}_linear_span"))
    for how, kw, key in cases:
        got, want = ds.tf.shade(agg, how=how, **kw).data, g[key]
        assert got.dtype == np.uint32 and got.shape == want.shape, key
        cg, cw = _channels(got), _channels(want)
        if how in ("eq_hist", "linear"):
            np.testing.assert_array_equal(cg[..., 3], cw[..., 3], err_msg=key + " alpha")
        else:
            assert np.abs(cg[..., 3] - cw[..., 3]).max() <= 1, key
        visible = cw[..., 3] > 0
        diff = np.abs(cg[..., :3] - cw[..., :3])[visible]
        assert diff.max() <= 1 and (diff > 0).mean() < 0.02, (key, int(diff.max()), float((diff > 0).mean()))


def _sqrt_how(d, m):
    return np.where(m, np.nan, np.sqrt(d))


def _fake_cmap(x, bytes=True):     # noqa: A002  (stand-in for a matplotlib colormap; identical to tests/golden/make_golden.py)
    v = np.nan_to_num(np.clip(x, 0, 1))
    out = np.empty(x.shape + (4,), dtype=np.uint8)
    out[..., 0] = (v * 255).astype(np.uint8)
    out[..., 1] = 255 - (v * 200).astype(np.uint8)
    out[..., 2] = 64
    out[..., 3] = 255
    return out


def test_shade_python_callables_and_discrete_keys_golden():
    """tf.shade with a callable `how`, a callable (matplotlib-style) cmap and a discrete colour key on a 2-D aggregate
    (transfer_functions/__init__.py:218-231, 340-349, 535-612) vs the real reference (tests/golden/shade_extra.npz)."""
    import datashader_b200 as ds
    g = load("shade_extra.npz")
    for name in ("f64", "u32"):
        agg = _agg2(g[f"x_{name}_in"])
        _check(ds.tf.shade(agg, cmap=["black", "red", "white"], how=_sqrt_how).data, g[f"x_{name}_callhow_list"], "linear", f"{name} callable how, list")
        _check(ds.tf.shade(agg, cmap="#3070c0", how=_sqrt_how, min_alpha=20).data, g[f"x_{name}_callhow_single"], "linear", f"{name} callable how, single")
        for how in ("linear", "log", "cbrt"):
            _check(ds.tf.shade(agg, cmap=_fake_cmap, how=how, alpha=200).data, g[f"x_{name}_callcmap_{how}"], "linear", f"{name} callable cmap {how}")
    cats = _agg2(g["x_cats_in"])
    key = {1: "red", 2: "#00ff00", 4: (0, 0, 255)}
    np.testing.assert_array_equal(ds.tf.shade(cats, color_key=key).data, g["x_cats_key"])
    np.testing.assert_array_equal(ds.tf.shade(cats, color_key=key, alpha=100).data, g["x_cats_key_a100"])
    np.testing.assert_array_equal(ds.tf.shade(cats, color_key=key, color_baseline=0.25).data, g["x_cats_key_base"])
