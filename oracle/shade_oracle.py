"""TEST INFRASTRUCTURE ONLY - numpy restatement of the reference's tf.shade for the paths the B200 build
covers (datashader/transfer_functions/__init__.py): eq_hist (:148-215), _interpolate (:251-357, list and
single-colour cmaps), _colorize (:359-463) and _interpolate_alpha (:466-532).

The arithmetic the reference delegates to numpy (unique / histogram / cumsum / interp / matmul) is
delegated to numpy here too.  Pinned against tests/golden/shade.npz, which the real reference produced
(tests/golden/make_golden.py).  Colours arrive as (r, g, b) tuples; no datashader_b200 import.
"""
from __future__ import annotations

import numpy as np


def eq_hist(data, mask=None, nbins=256 * 256):
    """transfer_functions/__init__.py:148-215"""
    if mask is not None and np.all(mask):
        return np.full_like(data, np.nan), 0
    data2 = data if mask is None else data[~mask]
    if data2.dtype == bool or (np.issubdtype(data2.dtype, np.integer) and np.ptp(data2) < nbins):
        values, counts = np.unique(data2, return_counts=True)
        vmin, vmax = values[0].item(), values[-1].item()
        interval = vmax - vmin
        bin_centers = np.arange(vmin, vmax + 1)
        hist = np.zeros(interval + 1, dtype="uint64")
        hist[values - vmin] = counts
        discrete_levels = len(values)
    else:
        hist, bin_edges = np.histogram(data2, bins=nbins)
        bin_centers = (bin_edges[:-1] + bin_edges[1:]) / 2
        keep_mask = (hist > 0)
        discrete_levels = np.count_nonzero(keep_mask)
        if discrete_levels != len(hist):
            hist = hist[keep_mask]
            bin_centers = bin_centers[keep_mask]
    cdf = hist.cumsum()
    cdf = cdf / float(cdf[-1])
    out = np.interp(data, bin_centers, cdf).reshape(data.shape)
    return out if mask is None else np.where(mask, np.nan, out), discrete_levels


_HOW = {
    "log": lambda d, m: np.log1p(np.where(m, np.nan, d)),
    "cbrt": lambda d, m: np.where(m, np.nan, d) ** (1 / 3.),
    "linear": lambda d, m: np.where(m, np.nan, d),
    "eq_hist": eq_hist,
}


def _rescale_discrete_levels(discrete_levels, span):
    """:232-248"""
    m = -0.5 / 98.0
    c = 1.5 - 2 * m
    multiple = m * discrete_levels + c
    if multiple > 1:
        lower_span = max(span[1] - multiple * (span[1] - span[0]), 0)
        span = (lower_span, 1)
    return span


def _masked_clip_2d(data, mask, lower, upper):
    """_cpu_utils.masked_clip_2d: in place, the bound is cast to the array dtype on assignment"""
    lo_sel = (~mask) & (data < lower)
    hi_sel = (~mask) & ~(data < lower) & (data > upper)
    data[lo_sel] = np.array(lower).astype(data.dtype)
    data[hi_sel] = np.array(upper).astype(data.dtype)


def interpolate_alpha(data, total, mask, how, alpha, min_alpha, rescale_discrete_levels=False, span=None):
    """_interpolate_alpha, :466-532"""
    if span is not None:
        if how == "eq_hist":
            raise ValueError("span is not (yet) valid to use with eq_hist")
        with np.errstate(invalid="ignore", divide="ignore"):
            offset = np.array(span, dtype=data.dtype)[0]
            if total.dtype.kind == "u" and np.nanmin(total) == 0:
                mask = mask | (total <= 0)
                total = np.where(~mask, total, np.nan)
            total = total.copy()
            _masked_clip_2d(total, mask, *span)
            a_scaled = _HOW[how](total - offset, mask)
            norm_span = np.hstack(_HOW[how]([0, span[1] - span[0]], 0))
            a_float = np.interp(a_scaled, norm_span, np.array([min_alpha, alpha]), left=0, right=255)
            return np.nan_to_num(a_float, copy=False).astype(np.uint8)
    with np.errstate(invalid="ignore", divide="ignore"):
        offset = np.nanmin(total)
        if total.dtype.kind == "u" and offset == 0:
            mask = mask | (total == 0)
            if not np.all(mask):
                offset = total[total > 0].min()
            total = np.where(~mask, total, np.nan)
        a_scaled = _HOW[how](total - offset, mask)
        discrete_levels = None
        if isinstance(a_scaled, (list, tuple)):
            a_scaled, discrete_levels = a_scaled
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            norm_span = [np.nanmin(a_scaled).item(), np.nanmax(a_scaled).item()]
        if rescale_discrete_levels and discrete_levels is not None:
            norm_span = _rescale_discrete_levels(discrete_levels, norm_span)
        norm_span = np.hstack(norm_span)
        a_float = np.interp(a_scaled, norm_span, np.array([min_alpha, alpha]), left=0, right=255)
        return np.nan_to_num(a_float, copy=False).astype(np.uint8)


def shade_categorical(data, colors, how="eq_hist", alpha=255, min_alpha=40, color_baseline=None,
                      rescale_discrete_levels=False, span=None):
    """_colorize for a [H, W, C] aggregate, :359-463.  colors: list of (r, g, b)."""
    rs, gs, bs = map(np.array, zip(*colors))
    color_data = np.array(data, order="C", copy=True)
    with np.errstate(invalid="ignore", divide="ignore"):
        nan_mask = np.isnan(data) if data.dtype.kind == "f" else np.zeros(data.shape, bool)
        color_mask = ~nan_mask
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            baseline = np.nanmin(color_data) if color_baseline is None else color_baseline
        if baseline > 0:
            np.subtract(color_data, baseline, out=color_data, where=color_mask)
        else:
            np.add(color_data, -baseline, out=color_data, where=color_mask, casting="unsafe")
        if (color_baseline is not None) and (color_data.dtype.kind != "u"):
            np.maximum(color_data, 0, out=color_data)
        color_data = color_data.astype(np.float32)
        np.nan_to_num(color_data, copy=False)
        color_total = np.sum(color_data, axis=2)
        color_mask_f = color_mask.astype(np.float32)
        RGB = np.stack([rs, gs, bs], axis=1).astype(np.float32)
        rgb_sum = color_data @ RGB
        rgb_avg_present = color_mask_f @ RGB
        rgb_array = (rgb_sum / color_total[..., None]).astype(np.uint8)
        cmask_sum = np.sum(color_mask_f, axis=2)
        rgb2 = (rgb_avg_present / cmask_sum[..., None]).astype(np.uint8)
        missing_colors = (color_total == 0)
        if np.any(missing_colors):
            rgb_array = np.where(missing_colors[..., None], rgb2, rgb_array)
        # nansum_missing (utils.py:161-181)
        if data.dtype.kind == "f":
            missing = np.isnan(data)
            all_empty = np.all(missing, axis=2)
            total = np.where(missing & ~all_empty[..., None], 0, data).sum(axis=2)
        else:
            total = data.sum(axis=2)
        mask = np.isnan(total) if total.dtype.kind == "f" else np.zeros(total.shape, bool)
        a = interpolate_alpha(data, total, mask, how, alpha, min_alpha, rescale_discrete_levels, span)
    rgba = np.empty((a.shape[0], a.shape[1], 4), dtype=np.uint8)
    rgba[..., :3] = rgb_array
    rgba[..., 3] = a
    return rgba.view(np.uint32).reshape(a.shape)


def shade_2d(data, cmap, how="eq_hist", alpha=255, min_alpha=40, rescale_discrete_levels=False, span=None):
    """_interpolate, :251-357.  cmap: list of (r, g, b) tuples, or one (r, g, b) tuple."""
    data = data.copy()
    if np.issubdtype(data.dtype, np.bool_):
        mask = ~data
        data = data.astype(np.int8)
    elif data.dtype.kind == "u":
        mask = data == 0
    else:
        mask = np.isnan(data)
    if mask.all():
        return np.zeros(data.shape, dtype=np.uint32)
    if span is None:
        offset = np.nanmin(data[~mask])
    else:
        offset = np.array(span, dtype=data.dtype)[0]
        _masked_clip_2d(data, mask, *span)
    data -= offset
    with np.errstate(invalid="ignore", divide="ignore"):
        data = _HOW[how](data, mask)
        discrete_levels = None
        if isinstance(data, (list, tuple)):
            data, discrete_levels = data
        if span is None:
            masked_data = np.where(~mask, data, np.nan)
            span = np.nanmin(masked_data), np.nanmax(masked_data)
            if rescale_discrete_levels and discrete_levels is not None:
                span = _rescale_discrete_levels(discrete_levels, span)
        else:
            if how == "eq_hist":
                raise ValueError("span is not (yet) valid to use with eq_hist")
            span = _HOW[how]([0, span[1] - span[0]], 0)
        if isinstance(cmap, list):
            rspan, gspan, bspan = np.array(list(zip(*cmap)))
            span = np.linspace(span[0], span[1], len(cmap))
            r = np.nan_to_num(np.interp(data, span, rspan, left=255), copy=False).astype(np.uint8)
            g = np.nan_to_num(np.interp(data, span, gspan, left=255), copy=False).astype(np.uint8)
            b = np.nan_to_num(np.interp(data, span, bspan, left=255), copy=False).astype(np.uint8)
            a = np.where(np.isnan(data), 0, alpha).astype(np.uint8)
        else:
            color = cmap
            aspan = np.arange(min_alpha, alpha + 1)
            span = np.linspace(span[0], span[1], len(aspan))
            r = np.full(data.shape, color[0], dtype=np.uint8)
            g = np.full(data.shape, color[1], dtype=np.uint8)
            b = np.full(data.shape, color[2], dtype=np.uint8)
            a = np.nan_to_num(np.interp(data, span, aspan, left=0, right=255), copy=False).astype(np.uint8)
    rgba = np.dstack([r, g, b, a])
    return rgba.view(np.uint32).reshape(data.shape)


# ---------------------------------------------------------------------------------------------------------------
# Post-shade image operations: composite operators, spread, dynspread's density heuristic, stack, set_background.
# TEST INFRASTRUCTURE ONLY - restates datashader/composite.py and transfer_functions/__init__.py:115-145, 748-1060.
# Pinned by tests/golden/spread.npz (generated from the reference's own kernels) in tests/test_shade_oracle.py.
# ---------------------------------------------------------------------------------------------------------------
def _extract_scaled(x):
    """composite.py:32-38"""
    x = np.asarray(x, dtype=np.uint32)
    return tuple(((x >> s) & 255).astype(np.float64) / 255 for s in (0, 8, 16, 24))


def _combine_scaled(r, g, b, a):
    """composite.py:43-49: truncating uint32 casts, clamped at 255"""
    with np.errstate(invalid="ignore"):
        ch = [np.minimum(255, (np.nan_to_num(c) * 255).astype(np.uint32)) for c in (r, g, b, a)]
    return (ch[3] << 24) | (ch[2] << 16) | (ch[1] << 8) | ch[0]


def composite(how, src, dst):
    """The image operators of composite.py:72-125 on uint32 RGBA arrays (broadcasting): op(src, dst)."""
    src, dst = np.broadcast_arrays(np.asarray(src, np.uint32), np.asarray(dst, np.uint32))
    if how == "source":
        return np.where(src & np.uint32(0xff000000), src, dst).astype(np.uint32)
    sr, sg, sb, sa = _extract_scaled(src)
    dr, dg, db, da = _extract_scaled(dst)
    with np.errstate(divide="ignore", invalid="ignore"):
        if how == "over":
            factor = 1 - sa
            a = sa + da * factor
            r, g, b = ((s * sa + d * da * factor) / a for s, d in ((sr, dr), (sg, dg), (sb, db)))
        elif how == "add":
            a = np.minimum(1, sa + da)
            r, g, b = ((s * sa + d * da) / a for s, d in ((sr, dr), (sg, dg), (sb, db)))
        elif how == "saturate":
            a = np.minimum(1, sa + da)
            factor = np.minimum(sa, 1 - da)
            r, g, b = ((factor * s + d * da) / a for s, d in ((sr, dr), (sg, dg), (sb, db)))
        else:
            raise ValueError(how)
    out = _combine_scaled(r, g, b, a)
    return np.where(a == 0, np.uint32(0), out).astype(np.uint32)


def circle_mask(r):
    """transfer_functions/__init__.py:925-928"""
    x = np.arange(-r, r + 1, dtype="i4")
    return np.where(np.sqrt(x ** 2 + x[:, None] ** 2) <= r + 0.5, True, False)


def square_mask(px):
    """transfer_functions/__init__.py:918-922"""
    w = 2 * int(px) + 1
    return np.ones((w, w), dtype=bool)


def _arr_op(how, el, out):
    """composite.py:150-168 (scalars)"""
    if how == "add":
        return el + out
    if how == "max":
        return max(el, out)
    if how == "min":
        return min(el, out)
    if how == "source":
        return el if el else out
    raise ValueError(how)


def spread_plane(arr, mask, how, is_image):
    """One 2-D layer of tf.spread (transfer_functions/__init__.py:771-915): the reference's serial scatter, pixel by
    pixel in raster order, through the image / float / int kernel that spread() would pick.  Small inputs only."""
    arr = np.asarray(arr)
    w = mask.shape[0]
    extra = w // 2
    M, N = arr.shape
    float_type = arr.dtype in (np.float32, np.float64)
    out = np.full((M + 2 * extra, N + 2 * extra), np.nan if float_type else 0, dtype=arr.dtype)
    ignore_zeros = (not is_image) and (not float_type) and arr.dtype == np.uint32
    for y in range(M):
        for x in range(N):
            el = arr[y, x]
            if is_image:                                     # _build_spread_kernel :880-915
                if not ((int(el) >> 24) & 255):
                    continue
            for i in range(w):
                for j in range(w):
                    if not mask[i, j]:
                        continue
                    o = out[i + y, j + x]
                    if is_image:
                        res = el if o == 0 else composite(how, el, o)[()]
                    elif float_type:                         # _build_float_kernel :852-877
                        if np.isnan(el):
                            res = o
                        elif np.isnan(o):
                            res = el
                        else:
                            res = _arr_op(how, el, o)
                    else:                                    # _build_int_kernel :825-849
                        if ignore_zeros and el == 0:
                            res = o
                        elif ignore_zeros and o == 0:
                            res = el
                        else:
                            res = _arr_op(how, el, o)
                    out[i + y, j + x] = res
    return out[extra:extra + M, extra:extra + N].copy()


def spread(arr, px=1, shape="circle", how=None, is_image=False, mask=None):
    """tf.spread on a 2-D or 3-D ([H, W, C], per category plane) array."""
    arr = np.asarray(arr)
    if mask is None:
        if px == 0:
            return arr
        mask = circle_mask(px) if shape == "circle" else square_mask(px)
    how = how or ("over" if is_image else "add")
    if arr.ndim == 2:
        return spread_plane(arr, mask, how, is_image)
    return np.dstack([spread_plane(arr[:, :, c], mask, how, is_image) for c in range(arr.shape[2])])


def density(arr, px, is_image):
    """_rgb_density / _array_density (transfer_functions/__init__.py:1004-1051)."""
    arr = np.asarray(arr)
    if is_image:
        occ = ((arr >> 24) & 255) != 0
    elif arr.dtype in (np.float32, np.float64):
        occ = ~np.isnan(arr)
    else:
        occ = arr != 0
    M, N = occ.shape
    cnt = has = 0
    for y in range(M):
        for x in range(N):
            if occ[y, x]:
                cnt += 1
                if occ[max(0, y - px):min(y + px + 1, M), max(0, x - px):min(x + px + 1, N)].sum() > 1:
                    has += 1
    return has / cnt if cnt else np.inf


def dynspread_px(arr, threshold=0.5, max_px=3, is_image=False):
    """The radius dynspread settles on (transfer_functions/__init__.py:972-994)."""
    arr = np.asarray(arr)
    float_type = arr.dtype in (np.float32, np.float64)
    px_ = 0
    for px in range(1, max_px + 1):
        px_ = px
        if is_image or arr.ndim == 2:
            d = density(arr, px * 2, is_image)
        else:
            masked = ~np.isnan(arr) if float_type else (arr != 0)
            d = density(np.sum(masked, axis=2, dtype="uint32"), px * 2, False)
        if d > threshold:
            px_ -= 1
            break
    return px_


def stack(imgs, how="over"):
    """tf.stack: later images over earlier ones, reduce(flip(op)) (transfer_functions/__init__.py:139-144)."""
    out = np.asarray(imgs[0], np.uint32)
    for nxt in imgs[1:]:
        out = composite(how, nxt, out)
    return out
