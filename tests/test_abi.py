"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol include/dsb200.h
declares, and the host mirror of the reference interface validates arguments before any launch."""
import ctypes
import os

import numpy as np
import pytest

from datashader_b200 import _lib


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "libdsb200.so not built: run python -m datashader_b200._build"
    L = ctypes.CDLL(_lib.LIB_PATH)
    declared = _lib.declared_symbols()
    assert len(declared) >= 12
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert L.dsb_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header():
    # dsb_view: 4 x i32 + 8 x f64; dsb_base: 2 ints+ptr ... natural alignment
    assert ctypes.sizeof(_lib.View) == 16 + 64
    assert ctypes.sizeof(_lib.Base) == 48
    assert ctypes.sizeof(_lib.Plan) == 8 + 48 * 8 + 8 + 8 + 8
    assert ctypes.sizeof(_lib.LineLayout) == 24 + 3 * 8 + 3 * 8      # strides, flags, ragged starts + flat lengths


def test_argument_errors_do_not_need_a_gpu():
    L = _lib.lib()
    v = _lib.View(0, 0, 0, 0, 1, 0, 1, 0, 0, 1, 0, 1)
    p = _lib.Plan()
    rc = L.dsb_points(ctypes.byref(v), None, None, _lib.F32, 10, 0, ctypes.byref(p), None)
    assert rc == -1 and b"bad view" in L.dsb_last_error()
    v.width = v.height = 4
    rc = L.dsb_points(ctypes.byref(v), None, None, _lib.F32, 10, 0, ctypes.byref(p), None)
    assert rc == -1 and b"nops" in L.dsb_last_error()
    with pytest.raises(_lib.Dsb200Error):
        _lib.check(rc, "dsb_points")


def test_reduction_objects_mirror_reference_api():
    import datashader_b200 as ds
    assert repr(ds.count()) == "count(None)" or "count" in repr(ds.count())
    assert ds.mean("v") == ds.mean("v") and hash(ds.mean("v")) == hash(ds.mean("v"))
    assert ds.mean("v") != ds.mean("w")
    w = ds.where(ds.max("v"))
    assert w.column is ds.reductions.SpecialColumn.RowIndex
    b = ds.by("c", ds.sum("v"))
    assert b.cat_column == "c" and b.is_categorical()
    with pytest.raises(TypeError):
        ds.by(3)
    s = ds.summary(b=ds.count(), a=ds.max("v"))
    assert s.keys == ("a", "b")


def test_axis_math_matches_reference_formulas():
    from datashader_b200.core import LinearAxis, LogAxis
    s, t = LinearAxis().compute_scale_and_translate((-0.1, 1.05), 37)
    assert s == 37 / (1.05 - -0.1) and t == 0.1 * s
    idx = LinearAxis().compute_index((s, t), 37)
    np.testing.assert_array_equal(idx, ((np.arange(37) + 0.5) - t) / s)
    ls, lt = LogAxis().compute_scale_and_translate((1, 1000), 40)
    assert ls == 40 / 3.0 and lt == -0.0
