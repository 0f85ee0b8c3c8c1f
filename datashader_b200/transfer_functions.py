"""tf.shade and the post-shade image operations on the GPU (the reference's datashader/transfer_functions/__init__.py:148-745,
748-1051; composite.py).

shade: 3-D categorical aggregates (uint32 counts such as by('cat', count()), and float aggregates such as by('cat', mean()))
-> colour mix + alpha; 2-D aggregates with a list colormap, a single colour or a discrete colour key; how in {'eq_hist',
'log', 'cbrt', 'linear'} with span=None or an explicit span, alpha / min_alpha / rescale_discrete_levels.  The canvas-sized
work (totals, histogram, scan / CDF, per-pixel lookup and colour mixing) runs in libdsb200 (csrc/shade.cu); the host only
picks scalars (offset, histogram range) exactly the way _interpolate_alpha / eq_hist do.  A callable `how` or a callable
(matplotlib-style) cmap is applied on the host like the reference does (_interpolate_host); a callable cmap with
how='eq_hist' raises NotImplementedError.  spread / dynspread / stack / set_background run on the GPU (csrc/imageops.cu).
"""
from __future__ import annotations

import ctypes as C
from collections.abc import Iterator

import numpy as np
import torch

from . import _lib
from .palette import Sets1to3, rgb
from .frame import to_host_array
from .xr_compat import DataArray

__all__ = ["shade", "Image", "spread", "dynspread", "stack", "set_background"]

_HOW = {"eq_hist": 0, "log": 1, "cbrt": 2, "linear": 3}
_NBINS = 256 * 256

_p, _i32, _i64, _f64, _u32, _u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_uint32, C.c_uint64
_SIGS = {
    "dsb_shade_cat_totals": [_p, _i64, _i32, _p, _p, _p],
    "dsb_eqhist_hist_u64": [_p, _i64, _u64, _i32, _f64, _f64, _i32, _i32, _p, _p],
    "dsb_eqhist_hist_f64": [_p, _i64, _f64, _f64, _f64, _i32, _i32, _p, _p],
    "dsb_eqhist_scan": [_p, _i32, _i32, _f64, _f64, _p, _p, _p, _p],
    "dsb_shade_norm_span": [_i32, _f64, _f64, _p, _p, _p, _i32, _p, _p],
    "dsb_shade_cat_colorize": [_p, _p, _i64, _i32, _p, _u32, _u32, _u64, _i32, _i32, _p, _p, _p, _p, _f64, _f64, _i32, _f64, _f64,
                               _p, _p],
    "dsb_shade_map2d": [_p, _i64, _i32, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _f64, _f64, _p, _p],
}
_bound = False


def _lib_shade():
    global _bound
    L = _lib.lib()
    if not _bound:
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _bound = True
    return L


class Image(DataArray):
    """An RGBA image stored as uint32 (transfer_functions/__init__.py:30-78)."""
    __slots__ = ()
    __array_priority__ = 70
    border = 1

    def to_pil(self, origin="lower"):
        from PIL import Image as PILImage
        data = np.asarray(self.data)
        arr = np.flipud(data) if origin == "lower" else data
        return PILImage.fromarray(np.ascontiguousarray(arr).view(np.uint8).reshape(arr.shape + (4,)), "RGBA")

    def to_bytesio(self, format="png", origin="lower"):   # noqa: A002 - the reference's argument name
        """The encoded image in a rewound in-memory file (transfer_functions/__init__.py:48-52)."""
        import io
        buf = io.BytesIO()
        self.to_pil(origin).save(buf, format)
        buf.seek(0)
        return buf

    def _repr_png_(self):
        """Notebook PNG display (transfer_functions/__init__.py:54-56)."""
        return self.to_bytesio().getvalue()

    def _repr_html_(self):
        """Notebook HTML display: the PNG inlined as a data URI with the reference's border (:58-71)."""
        import base64
        blob = base64.b64encode(self.to_bytesio().getvalue()).decode("ascii")
        return f"<img style=\"margin: auto; border:{self.border}px solid\" src='data:image/png;base64,{blob}'/>"


def _device_tensor(data, device):
    if isinstance(data, torch.Tensor):
        return data.to(device)
    a = np.ascontiguousarray(data)
    if a.dtype == np.uint32:
        return torch.from_numpy(a.view(np.int32)).to(device).view(torch.uint32)
    if a.dtype == np.uint64:
        return torch.from_numpy(a.view(np.int64)).to(device)
    return torch.from_numpy(a).to(device)


def _host_transform(how, d, f32=False):
    """The reference's analytic transfer functions on a scalar, evaluated with numpy like the reference (in float32
    for a float32 canvas: `data ** (1/3.)` keeps the array dtype)."""
    d = np.float32(d) if f32 else np.float64(d)
    if how == "log":
        return float(np.log1p(d))
    if how == "cbrt":
        return float(d ** (np.float32(1 / 3.) if f32 else (1 / 3.)))
    return float(d)


def _eq_hist_tables(L, s, kind, values, npix, offset, mask_zero, dmax, integer_mode, rescale, device):
    """histogram -> scan -> (xp, cdf, meta, span) on the device.  kind: 'u64' (totals) or 'f64'."""
    nbins = int(dmax) + 1 if integer_mode else _NBINS
    hist = torch.empty(nbins, dtype=torch.int32, device=device)
    xp = torch.empty(nbins, dtype=torch.float64, device=device)
    cdf = torch.empty(nbins, dtype=torch.float64, device=device)
    meta = torch.zeros(2, dtype=torch.int32, device=device)
    span = torch.empty(2, dtype=torch.float64, device=device)
    first, last = 0.0, float(dmax)
    if kind == "u64":
        _lib.check(L.dsb_eqhist_hist_u64(values.data_ptr(), npix, int(offset), int(mask_zero), first, last, nbins,
                                         int(integer_mode), hist.data_ptr(), s), "dsb_eqhist_hist_u64")
    else:
        _lib.check(L.dsb_eqhist_hist_f64(values.data_ptr(), npix, float(offset), first, last, nbins, int(integer_mode),
                                         hist.data_ptr(), s), "dsb_eqhist_hist_f64")
    _lib.check(L.dsb_eqhist_scan(hist.data_ptr(), nbins, int(integer_mode), first, last, xp.data_ptr(), cdf.data_ptr(),
                                 meta.data_ptr(), s), "dsb_eqhist_scan")
    _lib.check(L.dsb_shade_norm_span(0, 0.0, float(dmax), xp.data_ptr(), cdf.data_ptr(), meta.data_ptr(), int(rescale),
                                     span.data_ptr(), s), "dsb_shade_norm_span")
    return xp, cdf, meta, span


def _colorize(agg, color_key, how, alpha, span, min_alpha, name, color_baseline, rescale_discrete_levels, device):
    """3-D categorical path: _colorize (:359-463) + _interpolate_alpha (:466-532)."""
    data = agg.data
    cats = list(np.asarray(agg.coords[agg.dims[-1]]))
    H, W = int(data.shape[0]), int(data.shape[1])
    coords = {agg.dims[1]: agg.coords[agg.dims[1]], agg.dims[0]: agg.coords[agg.dims[0]]}
    if not len(cats):
        return Image(np.zeros((H, W), dtype=np.uint32), dims=agg.dims[:-1], coords=coords, name=name)
    if color_key is None:
        raise ValueError("Color key must be provided, with at least as many " +
                         "colors as there are categorical fields")
    if not isinstance(color_key, dict):
        color_key = dict(zip(cats, color_key))
    if len(color_key) < len(cats):
        raise ValueError(f"Insufficient colors provided ({len(color_key)}) for the categorical "
                         f"fields available ({len(cats)})")
    dt = str(data.dtype).replace("torch.", "")
    colors = [rgb(color_key[c]) for c in cats]
    if dt != "uint32":
        return _colorize_float(agg, colors, how, alpha, span, min_alpha, name, color_baseline, rescale_discrete_levels, device,
                               coords)
    ncat = len(cats)
    RGB = np.array(colors, dtype=np.float32)                       # (C, 3)
    rgb2 = ((np.ones((1, ncat), np.float32) @ RGB) / np.float32(ncat)).astype(np.uint8)[0]   # :432-442
    fallback = int(rgb2[0]) | (int(rgb2[1]) << 8) | (int(rgb2[2]) << 16)

    L = _lib_shade()
    with torch.cuda.device(device):
        s = torch.cuda.current_stream(device).cuda_stream
        counts = _device_tensor(data, device).contiguous()
        npix = H * W
        total = torch.empty(npix, dtype=torch.int64, device=device)
        stats = torch.empty(4, dtype=torch.int64, device=device)
        _lib.check(L.dsb_shade_cat_totals(counts.data_ptr(), npix, ncat, total.data_ptr(), stats.data_ptr(), s),
                   "dsb_shade_cat_totals")
        min_entry, min_total, min_nz, max_total = [int(v) & 0xFFFFFFFFFFFFFFFF for v in stats.tolist()]
        baseline = min_entry if color_baseline is None else int(color_baseline)
        # _interpolate_alpha (:475-522)
        mask_zero = min_total == 0
        all_masked = mask_zero and max_total == 0
        clip_mode, clip_lo, clip_hi = 0, 0.0, 0.0
        if span is None:
            offset = (min_nz if not all_masked else 0) if mask_zero else min_total
            dmax = max_total - offset if not all_masked else 0
        else:       # explicit span: clip the totals, fixed normalisation range (:507-522)
            offset = int(np.array(span, dtype=np.uint32)[0])
            clip_mode, clip_lo, clip_hi = (1 if mask_zero else 2), float(span[0]), float(span[1])
            dmax = span[1] - span[0]
        rgbt = torch.from_numpy(RGB).to(device)
        out = torch.empty(npix, dtype=torch.int32, device=device)
        xp = cdf = meta = None
        if all_masked:
            span_t = torch.tensor([0.0, 1.0], dtype=torch.float64, device=device)
            how_code = 3
        elif how == "eq_hist":
            integer_mode = (not mask_zero) and dmax < _NBINS          # eq_hist :194-196 (totals stay uint64)
            xp, cdf, meta, span_t = _eq_hist_tables(L, s, "u64", total, npix, offset, mask_zero, dmax, integer_mode,
                                                  rescale_discrete_levels, device)
            how_code = 0
        else:
            span_t = torch.tensor([_host_transform(how, 0), _host_transform(how, dmax)], dtype=torch.float64, device=device)
            how_code = _HOW[how]
        _lib.check(L.dsb_shade_cat_colorize(
            counts.data_ptr(), total.data_ptr(), npix, ncat, rgbt.data_ptr(), fallback, baseline & 0xFFFFFFFF, offset,
            int(mask_zero), how_code, xp.data_ptr() if xp is not None else None, cdf.data_ptr() if cdf is not None else None,
            meta.data_ptr() if meta is not None else None, span_t.data_ptr(), float(min_alpha), float(alpha), clip_mode, clip_lo,
            clip_hi, out.data_ptr(), s), "dsb_shade_cat_colorize")
        img = to_host_array(out).view(np.uint32).reshape(H, W)
    return Image(img, dims=agg.dims[:-1], coords=coords, name=name)


def _colorize_float(agg, colors, how, alpha, span, min_alpha, name, color_baseline, rescale_discrete_levels, device, coords):
    """Categorical aggregates that are not uint32 counts - by(cat, mean | sum | max | min ...) -> float64 [H, W, C] with
    NaN for empty cells, or signed integers (_colorize :382-452).  Canvas-sized float32 arithmetic in the reference's
    order of operations: baseline subtraction on the present cells, NaN -> 0, colour = (data @ RGB) / total with the
    average colour of the present categories where the total is 0; alpha from the per-pixel nansum through the same
    transfer function as a 2-D aggregate (_interpolate_alpha :466-532).  The float32 sums over the category axis run in
    torch's order, not numpy's: colour bytes may differ from the reference's by 1 level where a quotient sits on an
    integer (tests state the bar); the alpha channel is exact."""
    t = _device_tensor(agg.data, device)
    if t.dtype == torch.bool:
        t = t.to(torch.uint8)
    H, W, ncat = (int(v) for v in t.shape)
    is_float = t.dtype.is_floating_point
    signed = is_float or t.dtype in (torch.int8, torch.int16, torch.int32, torch.int64)
    with torch.cuda.device(device):
        nan_mask = torch.isnan(t) if is_float else torch.zeros_like(t, dtype=torch.bool)
        present = ~nan_mask
        if color_baseline is None:
            if is_float:
                vals = t[present]
                baseline = vals.min() if vals.numel() else torch.tensor(float("nan"), device=device, dtype=t.dtype)
            else:
                baseline = t.min()
        else:
            baseline = torch.tensor(color_baseline, device=device).to(t.dtype)
        color_data = torch.where(present, t - baseline, t)
        if color_baseline is not None and signed:
            color_data = torch.where(color_data < 0, torch.zeros_like(color_data), color_data)      # np.maximum(color_data, 0)
        color_data = torch.nan_to_num(color_data.to(torch.float32), nan=0.0)
        color_total = color_data.sum(dim=2)
        RGB = torch.tensor(colors, dtype=torch.float32, device=device)                              # (C, 3)
        rgb_sum = color_data.reshape(-1, ncat) @ RGB
        present_f = present.to(torch.float32)
        rgb_avg_present = present_f.reshape(-1, ncat) @ RGB
        rgb_array = (rgb_sum / color_total.reshape(-1, 1))
        rgb2 = rgb_avg_present / present_f.sum(dim=2).reshape(-1, 1)
        missing = (color_total == 0).reshape(-1, 1)
        rgbv = torch.where(missing, rgb2, rgb_array)
        rgbv = torch.nan_to_num(rgbv, nan=0.0, posinf=0.0, neginf=0.0).to(torch.int64) & 0xFF     # .astype(np.uint8) of in-range values
        # total = nansum_missing(data, axis=2): NaN where every category is missing; alpha through the 2-D machinery
        if is_float:
            total = torch.where(present.any(dim=2), torch.nan_to_num(t, nan=0.0).sum(dim=2), torch.full((H, W), float("nan"), dtype=t.dtype, device=device))
        else:
            total = t.sum(dim=2)
        total_agg = DataArray(total, coords=coords, dims=agg.dims[:-1])
        if bool(torch.isnan(total).all()) if is_float else False:
            a = torch.zeros((H, W), dtype=torch.int64, device=device)
        else:
            alpha_img = _interpolate(total_agg, "#000000", how, alpha, span, min_alpha, name, rescale_discrete_levels, device)
            a = (torch.from_numpy(np.asarray(alpha_img.data).view(np.int32).astype(np.int64)).to(device) >> 24) & 0xFF
        packed = rgbv[:, 0] | (rgbv[:, 1] << 8) | (rgbv[:, 2] << 16) | (a.reshape(-1) << 24)
        img = packed.to(torch.int32) if False else (packed & 0xFFFFFFFF)
        out = img.cpu().numpy().astype(np.uint32).reshape(H, W)
    return Image(out, dims=agg.dims[:-1], coords=coords, name=name)


def _apply_discrete_colorkey(agg, color_key, alpha, name, color_baseline, device):
    """2-D aggregate of category values + {value: colour}: _apply_discrete_colorkey (:535-612).  A pixel whose value is a key
    gets that key's colour at `alpha`; the baseline arithmetic of the reference is kept as is (with an explicit
    color_baseline that leaves 1 - baseline != 0 the colour stays black, as it does there)."""
    if len(agg.data.shape) != 2:
        raise ValueError("agg must be 2D")
    if color_key is None or not isinstance(color_key, dict):
        raise ValueError("Color key must be provided as a dictionary")
    t = _device_tensor(agg.data, device)
    if t.dtype == torch.uint32:
        t = t.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    H, W = (int(v) for v in t.shape)
    with torch.cuda.device(device):
        matched = torch.zeros((H, W), dtype=torch.bool, device=device)
        rgb2 = torch.zeros((H, W, 3), dtype=torch.int64, device=device)
        for c, col in color_key.items():
            m = t == c
            matched |= m
            rgb2[m] = torch.tensor(rgb(col), dtype=torch.int64, device=device)
        data = torch.where(matched, 1.0, float("nan")).to(torch.float64)
        baseline = (1.0 if bool(matched.any()) else float("nan")) if color_baseline is None else color_baseline
        color_data = data.clone()
        if baseline > 0:
            color_data -= baseline
        elif baseline < 0:
            color_data += -baseline
        if color_baseline is not None:
            color_data = torch.where(color_data < 0, torch.zeros_like(color_data), color_data)
        color_data = torch.where(torch.isnan(data), torch.zeros_like(color_data), color_data)
        missing = (color_data == 0).unsqueeze(-1)
        rgbv = torch.where(missing, rgb2, torch.zeros_like(rgb2))
        a = torch.where(matched, int(np.uint8(alpha)), 0)
        packed = rgbv[..., 0] | (rgbv[..., 1] << 8) | (rgbv[..., 2] << 16) | (a << 24)
        out = packed.cpu().numpy().astype(np.uint32)
    return Image(out, dims=agg.dims, coords=agg.coords, name=name)


def _interpolate_host(agg, cmap, how, alpha, span, min_alpha, name, rescale_discrete_levels, device):
    """_interpolate (:251-357) for the two arguments that are Python callables and therefore run on the host: a callable
    `how(data, mask)` (the transformed canvas then goes through the device's linear colour mapping) and a callable `cmap`
    (a matplotlib colormap: `cmap(scaled, bytes=True)` - applied to the transformed canvas on the host, numpy on both sides)."""
    data = np.array(to_host_array(_device_tensor(agg.data, device)))
    if data.dtype == np.bool_:
        mask = ~data
        data = data.astype(np.int8)
    elif data.dtype.kind == "u":
        mask = data == 0
    else:
        mask = np.isnan(data)
    if mask.all():
        return Image(np.zeros(data.shape, dtype=np.uint32), coords=agg.coords, dims=agg.dims, attrs=agg.attrs, name=name)
    if how == "eq_hist":
        raise NotImplementedError("a callable cmap with how='eq_hist' is not supported by datashader_b200.tf.shade")
    fn = how if callable(how) else {"log": lambda d, m: np.log1p(np.where(m, np.nan, d)),
                                    "cbrt": lambda d, m: np.where(m, np.nan, d) ** (1 / 3.),
                                    "linear": lambda d, m: np.where(m, np.nan, d)}[how]
    if span is None:
        offset = np.nanmin(data[~mask])
    else:
        offset = np.array(span, dtype=data.dtype)[0]
        lo, hi = np.array(span).astype(data.dtype) if data.dtype.kind in "iu" else span
        sel = ~mask
        data[sel & (data < span[0])] = lo
        data[sel & ~(data < span[0]) & (data > span[1])] = hi
    data = data - offset
    with np.errstate(invalid="ignore", divide="ignore"):
        t = fn(data, mask)
        if isinstance(t, (list, tuple)):
            t = t[0]
        if span is None:
            md = np.where(~mask, t, np.nan)
            nspan = (np.nanmin(md), np.nanmax(md))
        else:
            nspan = fn([0, span[1] - span[0]], 0)
            if isinstance(nspan, (list, tuple)) and len(nspan) == 2 and not np.isscalar(nspan[0]):
                nspan = nspan[0]
    t = np.asarray(t, dtype=np.float64)
    if callable(cmap):
        scaled = (t - nspan[0]) / (nspan[1] - nspan[0])
        rgba = np.ascontiguousarray(cmap(scaled, bytes=True))
        rgba[:, :, 3] = np.where(np.isnan(scaled), 0, alpha).astype(np.uint8)
        return Image(rgba.view(np.uint32).reshape(t.shape), coords=agg.coords, dims=agg.dims, name=name)
    # callable `how`, ordinary cmap: the device's linear mapping of the transformed canvas over [nspan0, nspan1]
    shifted = DataArray(torch.from_numpy(np.where(np.isnan(t), np.nan, t - nspan[0])).to(device), coords=agg.coords, dims=agg.dims)
    return _interpolate(shifted, cmap, "linear", alpha, (0.0, float(nspan[1] - nspan[0])), min_alpha, name, False, device)


def _interpolate(agg, cmap, how, alpha, span, min_alpha, name, rescale_discrete_levels, device):
    """2-D path: _interpolate (:251-357) with a list / single-colour cmap."""
    data = agg.data
    if len(data.shape) != 2:
        raise ValueError("agg must be 2D")
    if isinstance(cmap, Iterator):
        cmap = list(cmap)
    if isinstance(cmap, tuple) and isinstance(cmap[0], str):
        cmap = list(cmap)
    if callable(cmap) or callable(how):
        return _interpolate_host(agg, cmap, how, alpha, span, min_alpha, name, rescale_discrete_levels, device)
    if not isinstance(cmap, (list, str, tuple)):
        raise TypeError("Expected `cmap` of `matplotlib.colors.Colormap`, "
                        f"`list`, `str`, or `tuple`; got: '{type(cmap)}'")
    H, W = int(data.shape[0]), int(data.shape[1])
    npix = H * W
    L = _lib_shade()
    with torch.cuda.device(device):
        s = torch.cuda.current_stream(device).cuda_stream
        t = _device_tensor(data, device).contiguous()
        integer = not t.dtype.is_floating_point
        is_f32 = t.dtype == torch.float32
        if t.dtype == torch.bool:
            mask = ~t
            v = t.to(torch.int64)
        elif t.dtype in (torch.uint32, torch.uint8, torch.uint16, torch.uint64):
            v = t.view(torch.int32).to(torch.int64) & 0xFFFFFFFF if t.dtype == torch.uint32 else t.to(torch.int64)
            mask = v == 0
        elif integer:
            v = t.to(torch.int64)
            mask = torch.zeros_like(v, dtype=torch.bool)      # np.isnan of signed ints is never true
        else:
            v = t
            mask = torch.isnan(t)
        if bool(mask.all()):
            return Image(np.zeros((H, W), dtype=np.uint32), coords=agg.coords, dims=agg.dims, attrs=agg.attrs, name=name)
        if span is None:
            valid = v[~mask]
            offset = valid.min()
            dmax = float((valid.max() - offset).item())
        else:
            # explicit span (:287-289, 310-315): clip the unmasked data to it (the bound is cast to the canvas dtype when
            # stored, masked_clip_2d), offset = span[0] in the canvas dtype, fixed normalisation range
            np_dt = np.dtype(np.int8) if t.dtype == torch.bool else _np_dtype_of(t)
            lo_c, hi_c = (x.item() for x in np.array(span).astype(np_dt)) if np_dt.kind in "iu" else (float(span[0]), float(span[1]))
            offset = np.array(span, dtype=np_dt)[0].item()
            below, above = (~mask) & (v < span[0]), (~mask) & ~(v < span[0]) & (v > span[1])
            v = torch.where(below, torch.full_like(v, lo_c), torch.where(above, torch.full_like(v, hi_c), v))
            dmax = span[1] - span[0]
        # data -= offset in the canvas dtype (f32 stays f32), then everything downstream is float64
        d = (v - offset).to(torch.float64)
        d = torch.where(mask, torch.full_like(d, float("nan")), d).reshape(-1).contiguous()
        xp = cdf = meta = None
        if how == "eq_hist":
            integer_mode = integer and dmax < _NBINS
            xp, cdf, meta, span_t = _eq_hist_tables(L, s, "f64", d, npix, 0.0, False, dmax, integer_mode,
                                                  rescale_discrete_levels, device)
            span_host = span_t.tolist()
        elif span is None:
            # span = nanmin / nanmax of the transformed data: evaluated in the canvas dtype
            span_host = [_host_transform(how, 0, is_f32), _host_transform(how, dmax, is_f32)]
            span_t = torch.tensor(span_host, dtype=torch.float64, device=device)
        else:
            # span = interpolater([0, span[1] - span[0]], 0): a python list, hence float64 (:315)
            span_host = [_host_transform(how, 0), _host_transform(how, dmax)]
            span_t = torch.tensor(span_host, dtype=torch.float64, device=device)
        if isinstance(cmap, list):
            cols = np.array([rgb(c) for c in cmap], dtype=np.float64)
            ncolors = len(cmap)
            cspan = torch.from_numpy(np.linspace(span_host[0], span_host[1], ncolors)).to(device)
        else:
            cols = np.array([rgb(cmap)], dtype=np.float64)
            ncolors, cspan = 1, None
        rs, gs, bs = (torch.from_numpy(np.ascontiguousarray(cols[:, k])).to(device) for k in range(3))
        out = torch.empty(npix, dtype=torch.int32, device=device)
        _lib.check(L.dsb_shade_map2d(
            d.data_ptr(), npix, _HOW[how] | (0x100 if is_f32 else 0), xp.data_ptr() if xp is not None else None,
            cdf.data_ptr() if cdf is not None else None, meta.data_ptr() if meta is not None else None, span_t.data_ptr(),
            ncolors, cspan.data_ptr() if cspan is not None else None, rs.data_ptr(), gs.data_ptr(), bs.data_ptr(),
            float(min_alpha), float(alpha), out.data_ptr(), s), "dsb_shade_map2d")
        img = to_host_array(out).view(np.uint32).reshape(H, W)
    return Image(img, coords=agg.coords, dims=agg.dims, name=name)


def shade(agg, cmap=["lightblue", "darkblue"], color_key=Sets1to3, how='eq_hist', alpha=255, min_alpha=40, span=None,  # noqa: B006
          name=None, color_baseline=None, rescale_discrete_levels=False):
    """Convert a DataArray to an RGBA image (same signature as the reference's tf.shade, :616-745)."""
    if not isinstance(agg, DataArray):
        raise TypeError("agg must be instance of DataArray")
    name = agg.name if name is None else name
    if not ((0 <= min_alpha <= 255) and (0 <= alpha <= 255)):
        raise ValueError(f"min_alpha ({min_alpha}) and alpha ({alpha}) must be between 0 and 255")
    if not callable(how) and how not in _HOW:
        raise ValueError(f"Unknown interpolation method: {how}")
    if span is not None:
        if how == "eq_hist":
            raise ValueError("span is not (yet) valid to use with eq_hist")
        span = (span[0], span[1])
    if rescale_discrete_levels and how != 'eq_hist':
        rescale_discrete_levels = False
    device = agg.data.device if isinstance(agg.data, torch.Tensor) and agg.data.is_cuda else torch.device("cuda")
    ndim = len(agg.data.shape)
    if ndim == 2:
        if color_key is not None and isinstance(color_key, dict):
            return _apply_discrete_colorkey(agg, color_key, alpha, name, color_baseline, device)
        return _interpolate(agg, cmap, how, alpha, span, min_alpha, name, rescale_discrete_levels, device)
    elif ndim == 3:
        return _colorize(agg, color_key, how, alpha, span, min_alpha, name, color_baseline, rescale_discrete_levels, device)
    raise ValueError("agg must use 2D or 3D coordinates")


# ---------------------------------------------------------------------------------------------------------------
# Post-shade image operations (transfer_functions/__init__.py:115-145, 748-1051; composite.py): the step after
# shade() in every real pipeline (pipeline.py:43-72 defaults to spread_fn=tf.dynspread).
# ---------------------------------------------------------------------------------------------------------------
_IMAGE_OPS = {"over": 0, "add": 1, "saturate": 2, "source": 3}
_ARRAY_OPS = {"add": 0, "max": 1, "min": 2, "source": 3}
_NP_DSB = {"float32": _lib.F32, "float64": _lib.F64, "int32": _lib.I32, "int64": _lib.I64, "uint32": _lib.U32}


def _validate_operator(how, is_image):
    """composite.py:18-29"""
    if is_image:
        if how not in _IMAGE_OPS:
            image_repr = ', '.join(repr(el) for el in _IMAGE_OPS)
            raise ValueError(f'Operator {how!r} not one of the supported image operators: {image_repr}')
    elif how not in _ARRAY_OPS:
        array_repr = ', '.join(repr(el) for el in _ARRAY_OPS)
        raise ValueError(f'Operator {how!r} not one of the supported array operators: {array_repr}')


def _np_dtype_of(data):
    return np.dtype(str(data.dtype).replace("torch.", ""))


def _result_like(data, out_t, np_dtype):
    """Same residency as the input: torch in -> torch out (device results), numpy in -> numpy out."""
    if isinstance(data, torch.Tensor):
        return out_t
    a = to_host_array(out_t)
    return a.view(np_dtype) if a.dtype != np_dtype else a


def _device():
    return torch.device("cuda", torch.cuda.current_device())


def _square_mask(px):
    """transfer_functions/__init__.py:918-922"""
    w = 2 * int(px) + 1
    return np.ones((w, w), dtype='bool')


def _circle_mask(r):
    """transfer_functions/__init__.py:925-928"""
    x = np.arange(-r, r + 1, dtype='i4')
    return np.where(np.sqrt(x**2 + x[:, None]**2) <= r+0.5, True, False)


_mask_lookup = {'square': _square_mask, 'circle': _circle_mask}


def set_background(img, color=None, name=None):
    """Return a new image, with the background set to `color` (transfer_functions/__init__.py:748-768)."""
    if not isinstance(img, Image):
        raise TypeError(f"Expected `Image`, got: `{type(img)}`")
    name = img.name if name is None else name
    if color is None:
        return img
    background = int(np.uint8(rgb(color) + (255,)).view('uint32')[0])
    dev = _device()
    L = _lib.lib()
    with torch.cuda.device(dev):
        src = _device_tensor(img.data, dev).contiguous()
        out = torch.empty_like(src)
        _lib.check(L.dsb_composite(src.data_ptr(), None, background, src.numel(), _IMAGE_OPS["over"], out.data_ptr(),
                                   torch.cuda.current_stream(dev).cuda_stream), "dsb_composite")
        data = _result_like(img.data, out, np.dtype("uint32"))
    return Image(data, coords=img.coords, dims=img.dims, name=name)


def stack(*imgs, **kwargs):
    """Combine images together, overlaying later images onto earlier ones (transfer_functions/__init__.py:115-145)."""
    if not imgs:
        raise ValueError("No images passed in")
    shapes = []
    for i in imgs:
        if not isinstance(i, Image):
            raise TypeError(f"Expected `Image`, got: `{type(i)}`")
        elif not shapes:
            shapes.append(tuple(i.shape))
        elif shapes and tuple(i.shape) not in shapes:
            raise ValueError("The stacked images must have the same shape.")
    name = kwargs.get('name', None)
    how = kwargs.get('how', 'over')
    if how not in _IMAGE_OPS:
        raise KeyError(how)
    if len(imgs) == 1:
        return imgs[0]
    dev = _device()
    L = _lib.lib()
    with torch.cuda.device(dev):
        s = torch.cuda.current_stream(dev).cuda_stream
        acc = _device_tensor(imgs[0].data, dev).contiguous()
        for nxt in imgs[1:]:                          # reduce(flip(op)): op(next, accumulated)
            src = _device_tensor(nxt.data, dev).contiguous()
            out = torch.empty_like(acc)
            _lib.check(L.dsb_composite(src.data_ptr(), acc.data_ptr(), 0, acc.numel(), _IMAGE_OPS[how], out.data_ptr(), s),
                       "dsb_composite")
            acc = out
        data = _result_like(imgs[0].data, acc, np.dtype("uint32"))
    return Image(data, coords=imgs[0].coords, dims=imgs[0].dims, name=name)


def spread(img, px=1, shape='circle', how=None, mask=None, name=None):
    """Spread pixels in an image or aggregate (transfer_functions/__init__.py:771-822): each pixel is expanded `px` pixels
    on all sides according to `shape` ('circle' or 'square') or an explicit odd square `mask`, merging with the
    compositing operator `how` ('over' for Images, 'add' otherwise by default)."""
    if not isinstance(img, DataArray):
        raise TypeError(f"Expected `xr.DataArray`, got: `{type(img)}`")
    is_image = isinstance(img, Image)
    name = img.name if name is None else name
    if mask is None:
        if not isinstance(px, int) or px < 0:
            raise ValueError("``px`` must be an integer >= 0")
        if px == 0:
            return img
        mask = _mask_lookup[shape](px)
    elif not (isinstance(mask, np.ndarray) and mask.ndim == 2 and
              mask.shape[0] == mask.shape[1] and mask.shape[0] % 2 == 1):
        raise ValueError("mask must be a square 2 dimensional ndarray with "
                         "odd dimensions.")
    if how is None:
        how = 'over' if is_image else 'add'
    _validate_operator(how, is_image)
    np_dtype = _np_dtype_of(img.data)
    if not is_image and np_dtype.name not in _NP_DSB:
        raise TypeError(f"spread: unsupported aggregate dtype {np_dtype}")
    dev = _device()
    L = _lib.lib()
    with torch.cuda.device(dev):
        s = torch.cuda.current_stream(dev).cuda_stream
        src = _device_tensor(img.data, dev).contiguous()
        m = torch.from_numpy(np.ascontiguousarray(mask != 0).view(np.uint8)).to(dev)
        out = torch.empty_like(src)
        H, W = int(src.shape[0]), int(src.shape[1])
        if is_image:
            _lib.check(L.dsb_spread_image(src.data_ptr(), H, W, m.data_ptr(), int(mask.shape[0]), _IMAGE_OPS[how], out.data_ptr(), s),
                       "dsb_spread_image")
        else:
            ncat = int(src.shape[2]) if src.ndim == 3 else 1
            _lib.check(L.dsb_spread_array(src.data_ptr(), _NP_DSB[np_dtype.name], H, W, ncat, m.data_ptr(), int(mask.shape[0]),
                                          _ARRAY_OPS[how], out.data_ptr(), s), "dsb_spread_array")
        data = _result_like(img.data, out, np_dtype)
    return img.__class__(data, dims=img.dims, coords=img.coords, name=name)


def _density(data, is_image, px):
    """_rgb_density / _array_density (transfer_functions/__init__.py:1004-1051) on the device."""
    dev = _device()
    L = _lib.lib()
    with torch.cuda.device(dev):
        src = _device_tensor(data, dev).contiguous()
        out2 = torch.zeros(2, dtype=torch.int64, device=dev)
        np_dtype = _np_dtype_of(src) if not isinstance(data, np.ndarray) else data.dtype
        _lib.check(L.dsb_density(src.data_ptr(), _NP_DSB[np.dtype(np_dtype).name], int(is_image), int(src.shape[0]), int(src.shape[1]),
                                 int(px), out2.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "dsb_density")
        cnt, has = (int(v) for v in out2.tolist())
    return has / cnt if cnt else np.inf


def dynspread(img, threshold=0.5, max_px=3, shape='circle', how=None, name=None):
    """Spread pixels dynamically based on the image density (transfer_functions/__init__.py:931-1001): spreading starts
    at 1 pixel and stops when the fraction of non-empty pixels with a non-empty neighbour exceeds `threshold`, or at
    `max_px`."""
    is_image = isinstance(img, Image)
    if not 0 <= threshold <= 1:
        raise ValueError("threshold must be in [0, 1]")
    if not isinstance(max_px, int) or max_px < 0:
        raise ValueError("max_px must be >= 0")
    np_dtype = _np_dtype_of(img.data)
    float_type = np_dtype in (np.float32, np.float64)
    flat = None
    px_ = 0
    for px in range(1, max_px + 1):
        px_ = px
        if is_image or len(img.shape) == 2:
            density = _density(img.data, is_image, px * 2)
        else:
            if flat is None:           # number of non-empty categories per pixel
                d = _device_tensor(img.data, _device())
                flat = ((~torch.isnan(d)) if float_type else (d != 0)).sum(dim=2).to(torch.int32)
            density = _density(flat, False, px * 2)
        if density > threshold:
            px_ = px_ - 1
            break
    if px_ >= 1:
        return spread(img, px_, shape=shape, how=how, name=name)
    else:
        return img
