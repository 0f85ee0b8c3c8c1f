"""Host-side logic that needs no GPU: how summary() plans are routed to the specialised kernels, the reference's argument
checks of the tf image operations, the spread masks."""
import numpy as np
import pytest


def _plan(ops):
    from datashader_b200 import _lib
    plan = _lib.Plan()
    plan.nops = len(ops)
    for k, (op, val_dtype, val, chk) in enumerate(ops):
        b = plan.ops[k]
        b.op, b.val_dtype, b.val, b.chk = op, val_dtype, val, chk
        b.agg = 0x1000 * (k + 1)
    return plan


def test_specialised_groups_routing():
    from datashader_b200 import _lib
    from datashader_b200.pipeline import _specialised_groups, _sub_plan
    V, W = 0x7000, 0x9000
    # summary(count(), mean(v), max(v), min(w)): [COUNT], [SUM v, COUNT v], [MAX32 v], [MIN32 w]
    plan = _plan([(_lib.OP_COUNT, _lib.NONE, None, None), (_lib.OP_SUM, _lib.F32, V, None), (_lib.OP_COUNT, _lib.F32, V, None),
                  (_lib.OP_MAX32, _lib.F32, V, None), (_lib.OP_MIN32, _lib.F32, W, None)])
    groups = _specialised_groups(plan)
    assert sorted(map(tuple, groups)) == [(0,), (1, 2), (3,), (4,)]
    sub = _sub_plan(plan, [1, 2])
    assert sub.nops == 2 and sub.ops[0].op == _lib.OP_SUM and sub.ops[1].op == _lib.OP_COUNT and sub.ops[1].agg == 0x3000
    # a float64 SUM has no K2 mean shape: it stays with the interpreted remainder, its COUNT goes to K2 count
    plan = _plan([(_lib.OP_SUM, _lib.F64, V, None), (_lib.OP_COUNT, _lib.F64, V, None), (_lib.OP_MATCHROW64, _lib.F64, V, None)])
    groups = _specialised_groups(plan)
    assert sorted(map(tuple, groups)) == [(0, 2), (1,)]
    # nothing to split: one group -> None (a single fused launch)
    assert _specialised_groups(_plan([(_lib.OP_SUM, _lib.F64, V, None), (_lib.OP_MATCHROW64, _lib.F64, V, None)])) is None


def test_spread_masks_and_argument_checks():
    import datashader_b200 as ds
    from datashader_b200 import transfer_functions as tfm
    from datashader_b200.xr_compat import DataArray
    assert tfm._circle_mask(1).tolist() == [[False, True, False], [True, True, True], [False, True, False]] or tfm._circle_mask(1).all()
    assert tfm._square_mask(2).shape == (5, 5) and tfm._square_mask(2).all()
    c3 = tfm._circle_mask(3)
    assert c3.shape == (7, 7) and c3[3].all() and not c3[0, 0] and np.array_equal(c3, c3.T)
    img = ds.tf.Image(np.zeros((3, 3), np.uint32), coords={"y": np.arange(3), "x": np.arange(3)}, dims=["y", "x"])
    arr = DataArray(np.zeros((3, 3)), coords={"y": np.arange(3), "x": np.arange(3)}, dims=["y", "x"])
    assert ds.tf.spread(img, px=0) is img                       # px == 0: returned untouched, before any device work
    with pytest.raises(ValueError, match="px"):
        ds.tf.spread(img, px=1.5)
    with pytest.raises(TypeError):
        ds.tf.spread(np.zeros((3, 3)))
    with pytest.raises(ValueError, match="supported image operators"):
        ds.tf.spread(img, how="max")
    with pytest.raises(ValueError, match="supported array operators"):
        ds.tf.spread(arr, how="saturate")
    with pytest.raises(ValueError, match="threshold"):
        ds.tf.dynspread(img, threshold=-0.1)
    with pytest.raises(ValueError, match="max_px"):
        ds.tf.dynspread(img, max_px=-1)
    with pytest.raises(ValueError, match="No images"):
        ds.tf.stack()
    with pytest.raises(TypeError):
        ds.tf.set_background(arr, "white")
    assert ds.tf.set_background(img, None) is img and ds.tf.stack(img) is img
    with pytest.raises(ValueError, match="span is not"):
        ds.tf.shade(arr, how="eq_hist", span=(0, 1))


def test_image_png_and_html_display():
    """Image.to_bytesio / _repr_png_ / _repr_html_ (transfer_functions/__init__.py:48-71): a PNG that decodes back to the
    same RGBA bytes (flipped for origin='lower'), and the data-URI <img> with the reference's 1 px border."""
    import base64
    import io
    import datashader_b200 as ds
    from PIL import Image as PILImage
    px = np.arange(12, dtype=np.uint32).reshape(3, 4) * np.uint32(0x01020304) + np.uint32(0xFF000000)
    img = ds.tf.Image(px, coords={"y": np.arange(3), "x": np.arange(4)}, dims=["y", "x"])
    fp = img.to_bytesio()
    assert fp.tell() == 0 and fp.getvalue()[:8] == b"\x89PNG\r\n\x1a\n"
    back = np.asarray(PILImage.open(fp).convert("RGBA"))
    assert np.array_equal(back, np.flipud(px).view(np.uint8).reshape(3, 4, 4))
    top = np.asarray(PILImage.open(img.to_bytesio(origin="upper")).convert("RGBA"))
    assert np.array_equal(top, px.view(np.uint8).reshape(3, 4, 4))
    assert img._repr_png_() == img.to_bytesio().getvalue()
    html = img._repr_html_()
    assert html.startswith("<img style=\"margin: auto; border:1px solid\" src='data:image/png;base64,")
    blob = html.split("base64,")[1].split("'")[0]
    assert np.array_equal(np.asarray(PILImage.open(io.BytesIO(base64.b64decode(blob))).convert("RGBA")), back)
