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


def interpolate_alpha(data, total, mask, how, alpha, min_alpha, rescale_discrete_levels=False):
    """_interpolate_alpha with span=None, :466-532"""
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
                      rescale_discrete_levels=False):
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
        a = interpolate_alpha(data, total, mask, how, alpha, min_alpha, rescale_discrete_levels)
    rgba = np.empty((a.shape[0], a.shape[1], 4), dtype=np.uint8)
    rgba[..., :3] = rgb_array
    rgba[..., 3] = a
    return rgba.view(np.uint32).reshape(a.shape)


def shade_2d(data, cmap, how="eq_hist", alpha=255, min_alpha=40, rescale_discrete_levels=False):
    """_interpolate with span=None, :251-357.  cmap: list of (r, g, b) tuples, or one (r, g, b) tuple."""
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
    offset = np.nanmin(data[~mask])
    data -= offset
    with np.errstate(invalid="ignore", divide="ignore"):
        data = _HOW[how](data, mask)
        discrete_levels = None
        if isinstance(data, (list, tuple)):
            data, discrete_levels = data
        masked_data = np.where(~mask, data, np.nan)
        span = np.nanmin(masked_data), np.nanmax(masked_data)
        if rescale_discrete_levels and discrete_levels is not None:
            span = _rescale_discrete_levels(discrete_levels, span)
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
