"""The backend module a datashader maintainer adds as `datashader/data_libraries/b200.py`: it registers
libdsb200's fused pipeline with the reference's own dispatcher, `bypixel.pipeline` (datashader/core.py:1446,
utils.py:95-122), for the device-frame type, with the reference backends' signature
`pipeline(df, schema, canvas, glyph, summary, *, antialias=False)` (data_libraries/pandas.py:26-28).

Nothing in the reference's Canvas, reductions or glyph classes changes: the reference's reduction and glyph OBJECTS arrive
here, are restated as this package's objects (same names and arguments), and the aggregation runs in libdsb200 through
datashader_b200.pipeline.  The result is wrapped in the reference's xarray types.

    import datashader, integration.b200 as b200
    b200.register(datashader)                      # what `from . import b200` in data_libraries/__init__.py would do
    datashader.core.bypixel.pipeline(frame, schema, canvas, glyph, agg)

tests/test_integration_cpu.py executes this file against the reference tree (signature, registration, dispatch, object
translation); tests/test_gpu_integration.py runs the reference's dispatcher into the GPU and compares with the reference's
own pandas backend.
"""
from __future__ import annotations

import datashader_b200 as dsb
from datashader_b200 import pipeline as _pipeline
from datashader_b200 import reductions as _rd
from datashader_b200.frame import DeviceFrame, HostFrame
from datashader_b200.glyphs import Point as _Point


def translate_reduction(red):
    """reference reduction object -> the same reduction of datashader_b200 (names and arguments are identical by design)."""
    name = type(red).__name__
    if name == "summary":
        return _rd.summary(**{k: translate_reduction(v) for k, v in zip(red.keys, red.values)})
    if name in ("by", "count_cat"):
        pre = red.categorizer
        pname = type(pre).__name__
        if pname == "category_codes":
            cat = pre.column
        elif pname == "category_modulo":
            cat = _rd.category_modulo(pre.column, pre.modulo, pre.offset)
        elif pname == "category_binning":
            cat = _rd.category_binning(pre.column, pre.bin0, pre.bin0 + pre.binsize * pre.nbins, pre.nbins,
                                       pre.bin_under == 0, pre.bin_over == pre.nbins - 1)
        else:
            raise NotImplementedError(f"categorizer {pname}")
        return _rd.by(cat, translate_reduction(red.reduction))
    if name == "where":
        lookup = red.column if isinstance(red.column, str) else None
        return _rd.where(translate_reduction(red.selector), lookup)
    if name in ("count", "sum"):
        return getattr(_rd, name)(red.column, self_intersect=getattr(red, "self_intersect", True))
    if name in ("any", "mean", "min", "max", "first", "last"):
        return getattr(_rd, name)(red.column)
    raise NotImplementedError(f"reduction {name} is outside libdsb200's hot path")


def translate_canvas(canvas):
    x_log = type(canvas.x_axis).__name__ == "LogAxis"
    y_log = type(canvas.y_axis).__name__ == "LogAxis"
    return dsb.Canvas(canvas.plot_width, canvas.plot_height, x_range=canvas.x_range, y_range=canvas.y_range,
                      x_axis_type="log" if x_log else "linear", y_axis_type="log" if y_log else "linear")


def make_pipeline(ref):
    """The function registered with bypixel.pipeline; `ref` is the reference package (for its xarray result types)."""
    import xarray as xr

    def b200_pipeline(df, schema, canvas, glyph, summary, *, antialias=False):
        if type(glyph).__name__ != "Point":
            raise NotImplementedError("integration/b200.py binds the Point glyph; lines and areas follow the same pattern")
        out = _pipeline.points(df, translate_canvas(canvas), _Point(glyph.x, glyph.y), translate_reduction(summary))

        def wrap(a):
            return xr.DataArray(a.data, coords=dict(a.coords), dims=list(a.dims), attrs=dict(a.attrs))

        if hasattr(out, "data"):
            return wrap(out)
        return xr.Dataset({k: wrap(out[k]) for k in out}, attrs=dict(out.attrs))

    return b200_pipeline


def register(ref):
    """Register the backend for DeviceFrame / HostFrame with the reference's dispatcher and return the function."""
    fn = make_pipeline(ref)
    ref.core.bypixel.pipeline.register((DeviceFrame, HostFrame), fn)
    return fn
