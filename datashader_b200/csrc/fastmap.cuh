// float32 fast path of the reference's pixel mapping (glyphs/points.py:193-203), shared by the point kernels.
#pragma once
#include "common.cuh"

// xf = fmaf(x, sx32, tx32) differs from the reference's exact value by at most `ex` (host-side bound:
// 2^-23 * (W + 1 + max|x| * |sx| + |tx|)); whenever xf is further than that from an integer its floor IS the
// reference's truncated f64 result, otherwise the exact f64 mapping is evaluated.  The bounds test is exact: xlo/xhi
// are the float32 values that bracket the f64 bounds from inside.
struct FastMap {
  float sx, tx, sy, ty, xlo, xhi, ylo, yhi, ex, ey, omex, omey;   // omex = 1 - ex
  int enabled;
};

static inline float f32_at_least(double v) { float f = (float)v; return ((double)f < v) ? nextafterf(f, INFINITY) : f; }
static inline float f32_at_most(double v) { float f = (float)v; return ((double)f > v) ? nextafterf(f, -INFINITY) : f; }

static inline FastMap make_fast_map(const dsb_view* v) {
  FastMap f;
  f.sx = (float)v->sx; f.tx = (float)v->tx; f.sy = (float)v->sy; f.ty = (float)v->ty;
  f.xlo = f32_at_least(v->xmin); f.xhi = f32_at_most(v->xmax); f.ylo = f32_at_least(v->ymin); f.yhi = f32_at_most(v->ymax);
  const double ax = fmax(fabs(v->xmin), fabs(v->xmax)), ay = fmax(fabs(v->ymin), fabs(v->ymax));
  // |x * sx| is at most max(ax * |sx|, |tx| + W + 2) for any point, in or out of bounds, whose fast image lands in
  // [-1, W + 1]; the bound covers the float32 roundings of sx, tx and of the fused multiply-add.
  const double mx = fmax(ax * fabs(v->sx), fabs(v->tx) + v->width + 2.0), my = fmax(ay * fabs(v->sy), fabs(v->ty) + v->height + 2.0);
  const double ex = ldexp(1.0, -23) * (v->width + 1.0 + mx + fabs(v->tx));
  const double ey = ldexp(1.0, -23) * (v->height + 1.0 + my + fabs(v->ty));
  f.ex = (float)ex; f.ey = (float)ey; f.omex = 1.0f - f.ex; f.omey = 1.0f - f.ey;
  f.enabled = !v->x_log && !v->y_log && ex < 0.125 && ey < 0.125 && isfinite(ex) && isfinite(ey) && v->sx > 0 && v->sy > 0;
  return f;
}

// the reference mapping, bounds test included, kept out of line (taken by ~0.1 % of the points); -1 = not on the canvas
static __device__ __noinline__ int map_exact_linear(const dsb_view& v, float xr, float yr) {
  return (int)map_to_cell<float>(v, xr, yr);
}
