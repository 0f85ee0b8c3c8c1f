"""The numpy shade oracle (oracle/shade_oracle.py) against the real reference's tf.shade outputs
(tests/golden/shade.npz) and the literal tables of datashader/tests/test_transfer_functions.py."""
import numpy as np
import pytest

from helpers import load
from oracle import shade_oracle as so

SETS1TO3 = ['#e41a1c', '#377eb8', '#4daf4a', '#984ea3', '#ff7f00', '#ffff33', '#a65628', '#f781bf', '#999999', '#66c2a5',
            '#fc8d62', '#8da0cb', '#a6d854', '#ffd92f', '#e5c494', '#ffffb3', '#fb8072', '#fdb462', '#fccde5', '#d9d9d9',
            '#ccebc5', '#ffed6f']


def _rgb(h):
    return int(h[1:3], 16), int(h[3:5], 16), int(h[5:7], 16)


LIGHTBLUE, DARKBLUE = (173, 216, 230), (0, 0, 139)
HOT = [(0, 0, 0), (139, 0, 0), (255, 0, 0), (255, 165, 0), (255, 255, 0), (255, 255, 255)]


@pytest.mark.parametrize("name", ["poisson5", "pareto16", "dense3", "single4"])
def test_categorical_golden(name):
    g = load("shade.npz")
    data = g[f"cat_{name}_in"]
    colors = [_rgb(c) for c in SETS1TO3[:data.shape[2]]]
    for how in ("eq_hist", "log", "cbrt", "linear"):
        np.testing.assert_array_equal(so.shade_categorical(data, colors, how=how), g[f"cat_{name}_{how}"], err_msg=how)
    np.testing.assert_array_equal(so.shade_categorical(data, colors, how="eq_hist", alpha=200, min_alpha=10),
                                  g[f"cat_{name}_eq_hist_a200_m10"])
    np.testing.assert_array_equal(so.shade_categorical(data, colors, how="eq_hist", rescale_discrete_levels=True),
                                  g[f"cat_{name}_eq_hist_rescale"])


@pytest.mark.parametrize("name", ["u32", "u32big", "f64", "f32"])
def test_2d_golden(name):
    g = load("shade.npz")
    data = g[f"d2_{name}_in"]
    for how in ("eq_hist", "log", "cbrt", "linear"):
        np.testing.assert_array_equal(so.shade_2d(data, [LIGHTBLUE, DARKBLUE], how=how), g[f"d2_{name}_{how}_default"], err_msg=how)
        np.testing.assert_array_equal(so.shade_2d(data, HOT, how=how), g[f"d2_{name}_{how}_hot"], err_msg=how)
        np.testing.assert_array_equal(so.shade_2d(data, (0x30, 0x70, 0xc0), how=how, min_alpha=20),
                                      g[f"d2_{name}_{how}_single"], err_msg=how)


def test_reference_literal_tables():
    """datashader/tests/test_transfer_functions.py:22-37, 86-120 (3x3 fixture, pink->red cmap)."""
    a = np.arange(10, 19, dtype="u4").reshape((3, 3))
    a[[0, 1, 2], [0, 1, 2]] = 0
    sol_eq_hist = np.array([[0, 4291543295, 4288846335], [4286149631, 0, 4283518207], [4280821503, 4278190335, 0]], dtype="u4")
    sol_linear = np.array([[0, 4291543295, 4289306879], [4287070463, 0, 4282597631], [4280361215, 4278190335, 0]], dtype="u4")
    sol_log = np.array([[0, 4291543295, 4286741503], [4283978751, 0, 4280492543], [4279242751, 4278190335, 0]], dtype="u4")
    sol_cbrt = np.array([[0, 4291543295, 4284176127], [4282268415, 0, 4279834879], [4278914047, 4278190335, 0]], dtype="u4")
    cmap = [(255, 192, 203), (255, 0, 0)]
    np.testing.assert_array_equal(so.shade_2d(a, cmap, how="eq_hist"), sol_eq_hist)
    np.testing.assert_array_equal(so.shade_2d(a, cmap, how="linear"), sol_linear)
    np.testing.assert_array_equal(so.shade_2d(a, cmap, how="log"), sol_log)
    np.testing.assert_array_equal(so.shade_2d(a, cmap, how="cbrt"), sol_cbrt)


def test_post_shade_ops_golden():
    """composite operators, spread (image / float / int / uint32 kernels), density: oracle vs the reference's kernels."""
    g = load("spread.npz")
    img, img2 = g["img"], g["img2"]
    for how in ("over", "add", "saturate", "source"):
        np.testing.assert_array_equal(so.composite(how, img, img2), g[f"comp_{how}"], err_msg=how)
        np.testing.assert_array_equal(so.composite(how, img, np.uint32(0xff204060)), g[f"comp_bg_{how}"], err_msg=how)
    np.testing.assert_array_equal(so.circle_mask(2), g["mask_c2"])
    np.testing.assert_array_equal(so.square_mask(2), g["mask_s2"])
    for key in g.files:
        if not key.startswith("spread_"):
            continue
        _, kind, mname, how = key.split("_")
        mask = g[f"mask_{mname}"]
        src = {"img": img, "f64": g["f64"], "f32": g["f32"], "i32": g["i32"], "u32": g["u32"], "cat": g["cat"]}[kind]
        got = so.spread(src, how=how, is_image=(kind == "img"), mask=mask)
        assert got.dtype == g[key].dtype, key
        np.testing.assert_array_equal(got, g[key], err_msg=key)
    for px in (1, 2, 4, 6):
        assert so.density(img, px, True) == float(g[f"density_img_{px}"])
        assert so.density(g["f64"], px, False) == float(g[f"density_f64_{px}"])
        assert so.density(g["u32"], px, False) == float(g[f"density_u32_{px}"])
    for px in (2, 4, 6):
        assert so.density(g["sparse"], px, True) == float(g[f"density_sparse_{px}"])


SPAN_CASES_2D = {"u32": [(2, 9), (0.5, 7.5)], "f64": [(-5.0, 12.5)], "f32": [(0.2, 0.7)]}
SPAN_CASES_CAT = {"poisson5": [(3, 20), (2.5, 15.5)], "dense3": [(40, 90), (30.5, 100.25)]}


def test_span_golden():
    """tf.shade(span=...) (clip to the span, fixed normalisation range) vs the reference; eq_hist + span raises."""
    g, gs = load("shade.npz"), load("shade_span.npz")
    for name, spans in SPAN_CASES_2D.items():
        data = g[f"d2_{name}_in"]
        for k, span in enumerate(spans):
            for how in ("log", "cbrt", "linear"):
                np.testing.assert_array_equal(so.shade_2d(data, [LIGHTBLUE, DARKBLUE], how=how, span=span),
                                              gs[f"d2_{name}_s{k}_{how}_default"], err_msg=f"{name} {span} {how}")
                np.testing.assert_array_equal(so.shade_2d(data, (0x30, 0x70, 0xc0), how=how, min_alpha=20, span=span),
                                              gs[f"d2_{name}_s{k}_{how}_single"], err_msg=f"{name} {span} {how} single")
    for name, spans in SPAN_CASES_CAT.items():
        data = g[f"cat_{name}_in"]
        colors = [_rgb(c) for c in SETS1TO3[:data.shape[2]]]
        for k, span in enumerate(spans):
            for how in ("log", "cbrt", "linear"):
                np.testing.assert_array_equal(so.shade_categorical(data, colors, how=how, span=span), gs[f"cat_{name}_s{k}_{how}"],
                                              err_msg=f"{name} {span} {how}")
    with pytest.raises(ValueError, match="span is not"):
        so.shade_2d(g["d2_u32_in"], [LIGHTBLUE, DARKBLUE], how="eq_hist", span=(1, 5))
