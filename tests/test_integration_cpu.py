"""INTEGRATION.md's backend module (integration/b200.py) executed against the reference tree staged in oracle/_ref:
registration with the reference's own dispatcher, the backend signature, dispatch by frame type, and the translation of
the reference's reduction objects.  No GPU needed (nothing is launched)."""
import inspect
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


@pytest.fixture(scope="module")
def ref():
    if not os.path.isdir(os.path.join(REF, "datashader")):
        pytest.skip("oracle/_ref not staged (python oracle/make_ref.py needs /root/reference)")
    for p in (os.path.join(REF, "_shims"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import datashader
    return datashader


def test_backend_registers_with_the_reference_dispatcher(ref):
    from integration import b200
    from datashader_b200.frame import DeviceFrame, HostFrame
    fn = b200.register(ref)
    lookup = ref.core.bypixel.pipeline._lookup
    assert lookup[DeviceFrame] is fn and lookup[HostFrame] is fn
    # the same signature as the reference's own backends (data_libraries/pandas.py:26-28)
    import pandas as pd
    want = inspect.signature(lookup[pd.DataFrame])
    got = inspect.signature(fn)
    assert list(got.parameters) == list(want.parameters) == ["df", "schema", "canvas", "glyph", "summary", "antialias"]
    assert got.parameters["antialias"].kind is inspect.Parameter.KEYWORD_ONLY and got.parameters["antialias"].default is False


def test_reference_reductions_translate_one_to_one(ref):
    from integration import b200
    import datashader_b200 as dsb
    ds = ref
    pairs = [
        (ds.count(), dsb.count()), (ds.count("v"), dsb.count("v")), (ds.any(), dsb.any()), (ds.sum("v"), dsb.sum("v")),
        (ds.mean("v"), dsb.mean("v")), (ds.max("v"), dsb.max("v")), (ds.min("v"), dsb.min("v")),
        (ds.first("v"), dsb.first("v")), (ds.last("v"), dsb.last("v")),
        (ds.where(ds.max("v"), "o"), dsb.where(dsb.max("v"), "o")), (ds.where(ds.first("v")), dsb.where(dsb.first("v"))),
        (ds.by("c", ds.mean("v")), dsb.by("c", dsb.mean("v"))), (ds.count_cat("c"), dsb.count_cat("c")),
        (ds.sum("v", self_intersect=False), dsb.sum("v", self_intersect=False)),
    ]
    for r, want in pairs:
        got = b200.translate_reduction(r)
        assert got._hashable_inputs()[1:] == want._hashable_inputs()[1:], r          # count_cat is by(col, count())
    s = b200.translate_reduction(ds.summary(a=ds.count(), b=ds.by("c", ds.max("v"))))
    assert s.keys == ("a", "b") and isinstance(s.values[1], dsb.by)
    with pytest.raises(NotImplementedError):
        b200.translate_reduction(ds.var("v"))


def test_canvas_translation_keeps_axes_and_ranges(ref):
    from integration import b200
    c = ref.Canvas(plot_width=40, plot_height=30, x_range=(1, 1000), y_range=(0, 5), x_axis_type="log")
    t = b200.translate_canvas(c)
    assert (t.plot_width, t.plot_height, t.x_range, t.y_range) == (40, 30, (1, 1000), (0, 5))
    assert t.x_axis.is_log and not t.y_axis.is_log
