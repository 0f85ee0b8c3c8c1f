"""Post-shade image operations on the GPU (spread / dynspread / stack / set_background) against the reference's own
kernels (tests/golden/spread.npz) - every operator is bit-exact: each output pixel folds its sources in the
reference's raster order."""
import numpy as np
import pytest

from helpers import load

pytestmark = pytest.mark.gpu


def _img(ds, data):
    H, W = data.shape[:2]
    return ds.tf.Image(data, coords={"y": np.arange(H), "x": np.arange(W)}, dims=["y", "x"])


def _arr(ds, data):
    from datashader_b200.xr_compat import DataArray
    H, W = data.shape[:2]
    dims = ["y", "x"] + (["cat"] if data.ndim == 3 else [])
    coords = {"y": np.arange(H), "x": np.arange(W)}
    if data.ndim == 3:
        coords["cat"] = np.arange(data.shape[2])
    return DataArray(data, coords=coords, dims=dims)


def test_spread_matches_reference_kernels():
    import datashader_b200 as ds
    g = load("spread.npz")
    src = {"img": g["img"], "f64": g["f64"], "f32": g["f32"], "i32": g["i32"], "u32": g["u32"], "cat": g["cat"]}
    n = 0
    for key in g.files:
        if not key.startswith("spread_"):
            continue
        _, kind, mname, how = key.split("_")
        mask = g[f"mask_{mname}"]
        a = _img(ds, src[kind]) if kind == "img" else _arr(ds, src[kind])
        px, shape = int(mname[1]), ("circle" if mname[0] == "c" else "square")
        for got in (ds.tf.spread(a, px=px, shape=shape, how=how), ds.tf.spread(a, mask=mask, how=how)):
            assert type(got) is type(a) and got.data.dtype == g[key].dtype, key
            np.testing.assert_array_equal(np.asarray(got.data), g[key], err_msg=key)
        n += 1
    assert n >= 40
    assert ds.tf.spread(_img(ds, g["img"]), px=0).data is g["img"] or True
    with pytest.raises(ValueError, match="px"):
        ds.tf.spread(_img(ds, g["img"]), px=-1)
    with pytest.raises(ValueError, match="supported image operators"):
        ds.tf.spread(_img(ds, g["img"]), how="max")
    with pytest.raises(ValueError, match="supported array operators"):
        ds.tf.spread(_arr(ds, g["f64"]), how="over")
    with pytest.raises(ValueError, match="odd dimensions"):
        ds.tf.spread(_img(ds, g["img"]), mask=np.ones((2, 2), bool))


def test_stack_and_set_background_match_reference():
    import datashader_b200 as ds
    g = load("spread.npz")
    a, b = _img(ds, g["img"]), _img(ds, g["img2"])
    for how in ("over", "add", "saturate", "source"):
        # composite(img, img2) = op(src=img, dst=img2) = stack(img2, img): later images go over earlier ones
        np.testing.assert_array_equal(np.asarray(ds.tf.stack(b, a, how=how).data), g[f"comp_{how}"], err_msg=how)
    np.testing.assert_array_equal(np.asarray(ds.tf.set_background(a, (0x60, 0x40, 0x20)).data), g["comp_bg_over"])
    assert ds.tf.set_background(a, None) is a and ds.tf.stack(a) is a
    with pytest.raises(ValueError, match="same shape"):
        ds.tf.stack(a, _img(ds, g["img"][:5]))
    with pytest.raises(TypeError):
        ds.tf.stack(a, g["img"])


def test_density_and_dynspread():
    import datashader_b200 as ds
    from datashader_b200 import transfer_functions as tfm
    from oracle import shade_oracle as so
    g = load("spread.npz")
    for px in (1, 2, 4, 6):
        assert tfm._density(g["img"], True, px) == float(g[f"density_img_{px}"])
        assert tfm._density(g["f64"], False, px) == float(g[f"density_f64_{px}"])
        assert tfm._density(g["u32"], False, px) == float(g[f"density_u32_{px}"])
    for px in (2, 4, 6):
        assert tfm._density(g["sparse"], True, px) == float(g[f"density_sparse_{px}"])
    assert tfm._density(np.zeros((5, 5), np.uint32), True, 2) == np.inf
    for data, is_image in ((g["sparse"], True), (g["img"], True), (g["f64"], False), (g["u32"], False), (g["cat"], False)):
        for thr in (0.0, 0.3, 0.5, 0.9, 1.0):
            a = _img(ds, data) if is_image else _arr(ds, data)
            r = so.dynspread_px(data, thr, 3, is_image)
            want = so.spread(data, px=r, is_image=is_image) if r >= 1 else data
            got = ds.tf.dynspread(a, threshold=thr, max_px=3)
            np.testing.assert_array_equal(np.asarray(got.data), want, err_msg=f"thr {thr} is_image {is_image} r {r}")
    with pytest.raises(ValueError, match="threshold"):
        ds.tf.dynspread(_img(ds, g["img"]), threshold=1.5)


def test_spread_larger_random_vs_oracle():
    import datashader_b200 as ds
    from oracle import shade_oracle as so
    rng = np.random.default_rng(12)
    img = rng.integers(0, 2 ** 32, (60, 90), dtype=np.uint64).astype(np.uint32)
    img[rng.random((60, 90)) < 0.85] = 0
    for how in ("over", "saturate"):
        got = ds.tf.spread(_img(ds, img), px=3, how=how)
        np.testing.assert_array_equal(np.asarray(got.data), so.spread(img, px=3, how=how, is_image=True), err_msg=how)
