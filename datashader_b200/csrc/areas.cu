// Area glyphs: one thread per trapezoid, x-driven double-Bresenham scan fill fused with the accumulator plan.
// Replaces _build_draw_trapezoid_y (glyphs/area.py:1076-1320), _skip_or_clip_trapezoid_y (:1323-1380) and the
// extend kernels of the twelve area layouts (:1383-2083; the ragged ones read flat arrays + start indices).  "to zero" areas pass ys1 == NULL
// (y1 = y2 = 0.0, stacked = False); "to line" areas pass the second curve (stacked = True).
#include "common.cuh"
#include "accum.cuh"

struct AreaArgs {
  dsb_view v;
  const void* xs;
  const void* ys0;
  const void* ys1;          // NULL: fill to y = 0
  long long nlines, nverts;
  long long x_line_stride, y_line_stride;
  int value_per_vertex, plot_start;
  long long row_offset;
  long long xxmax, yymax;   // round(mapper(max) * s + t), map_onto_pixel_snap (line.py:714-715)
  long long xmaxi, ymaxi;   // map_onto_pixel(xmax, ymax) (area.py:1178-1180)
  const long long* rg_x;    // ragged layouts (AreaToZeroAxis1Ragged / AreaToLineAxis1Ragged, area.py:1939-2083): first flat vertex
  const long long* rg_y;    //   of every row in xs / ys0 / ys1; NULL = dense
  const long long* rg_y1;
  long long rg_xlen, rg_ylen, rg_y1len;
  dsb_plan plan;
};

struct AreaCtx {
  const dsb_plan* plan;
  long long width, idx, row;
  int cat;
};

__device__ __forceinline__ void area_append(const AreaCtx& c, long long x, long long y) {
  long long cell = y * c.width + x;
  if (c.plan->ncat > 0) {
    if (c.cat < 0) return;
    cell = cell * c.plan->ncat + c.cat;
  }
  for (int k = 0; k < c.plan->nops; k++) apply_base(c.plan->ops[k], cell, c.idx, c.row, c.plan->notes);
}

__device__ __forceinline__ double mul64(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add64(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub64(double a, double b) { return __dadd_rn(a, -b); }

// line.py:783-800
__device__ __forceinline__ bool area_clipt(double p, double q, double& t0, double& t1) {
  if (p < 0 && q < 0) {
    double r = __ddiv_rn(q, p);
    if (r > t1) return false;
    else if (r > t0) t0 = r;
  } else if (p > 0 && q < p) {
    double r = __ddiv_rn(q, p);
    if (r < t0) return false;
    else if (r < t1) t1 = r;
  } else if (q < 0) {
    return false;
  }
  return true;
}

// numba keeps float32 operands in float32: differences of two float32 coordinates round to float32
template <typename XY> __device__ __forceinline__ double delta_xy(double a1, double a0);
template <> __device__ __forceinline__ double delta_xy<float>(double a1, double a0) { return (double)__fsub_rn((float)a1, (float)a0); }
template <> __device__ __forceinline__ double delta_xy<double>(double a1, double a0) { return sub64(a1, a0); }

// area.py:1083-1100
__device__ __forceinline__ bool clamp_y_indices(long long ystarti, long long ystopi, long long ymaxi, long long& cs,
                                                long long& ce) {
  bool oob = (ystarti < 0 && ystopi <= 0) || (ystarti > ymaxi && ystopi >= ymaxi);
  cs = max(0LL, min(ymaxi, ystarti));
  ce = max(-1LL, min(ymaxi + 1, ystopi));
  return oob;
}

// fill one pixel column between y_start and y_stop (area.py:1205-1222 and the two repeats)
__device__ __forceinline__ void fill_column(const AreaCtx& c, long long x, long long y_start, long long y_stop, bool stacked,
                                            long long ymaxi) {
  if (y_start == y_stop && !stacked) { area_append(c, x, y_start); return; }
  long long y = y_start;
  long long iy = (y_start < y_stop) - (y_stop < y_start);
  if (!stacked && -1 <= y_stop + iy && y_stop + iy <= ymaxi + 1) y_stop += iy;
  while (y != y_stop) { area_append(c, x, y); y += iy; }
}

template <typename XY>
__device__ void draw_trapezoid_y(const AreaArgs& a, const AreaCtx& c, double x0, double x1, double y0, double y1, double y2,
                                 double y3, bool trapezoid_start, bool stacked, bool second_is_xy) {
  const dsb_view& v = a.v;
  // _skip_or_clip_trapezoid_y, area.py:1323-1380
  bool skip = (x0 != x0) || (x1 != x1) || (y0 != y0) || (y1 != y1) || (y2 != y2) || (y3 != y3);
  if ((y0 > v.ymax && y1 > v.ymax && y2 > v.ymax && y3 > v.ymax) || (y0 < v.ymin && y1 < v.ymin && y2 < v.ymin && y3 < v.ymin)) return;
  double t0 = 0.0, t1 = 1.0;
  const double dx = delta_xy<XY>(x1, x0);
  const double dy0 = delta_xy<XY>(y3, y0);
  // the second curve is float64 zeros for "to zero" areas, column dtype otherwise
  const double dy1 = second_is_xy ? delta_xy<XY>(y2, y1) : sub64(y2, y1);
  if (!area_clipt(-dx, sub64(x0, v.xmin), t0, t1)) skip = true;
  if (!area_clipt(dx, sub64(v.xmax, x0), t0, t1)) skip = true;
  bool clipped_start = false, clipped_end = false;
  if (t1 < 1) { clipped_end = true; x1 = add64(x0, mul64(t1, dx)); y2 = add64(y1, mul64(t1, dy1)); y3 = add64(y0, mul64(t1, dy0)); }
  if (t0 > 0) { clipped_start = true; x0 = add64(x0, mul64(t0, dx)); y0 = add64(y0, mul64(t0, dy0)); y1 = add64(y1, mul64(t0, dy1)); }
  if (skip) return;

  // map_onto_pixel_snap, line.py:689-720
  auto mapx = [&](double x) { long long xx = __double2ll_rz(add64(mul64(v.x_log ? log10(x) : x, v.sx), v.tx)); return xx == a.xxmax ? xx - 1 : xx; };
  auto mapy = [&](double y) { long long yy = __double2ll_rz(add64(mul64(v.y_log ? log10(y) : y, v.sy), v.ty)); return yy == a.yymax ? yy - 1 : yy; };
  long long x0i = mapx(x0), y0i = mapy(y0), y1i = mapy(y1), x1i = mapx(x1), y2i = mapy(y2), y3i = mapy(y3);
  const long long xmaxi = a.xmaxi, ymaxi = a.ymaxi;

  long long dxi = x1i - x0i;
  const long long ix = (dxi > 0) - (dxi < 0);
  long long dy0i = y3i - y0i;
  const long long iy0 = (dy0i > 0) - (dy0i < 0);
  long long dy1i = y2i - y1i;
  const long long iy1 = (dy1i > 0) - (dy1i < 0);

  trapezoid_start = trapezoid_start || clipped_start;
  long long ys, ye;
  if (trapezoid_start) {
    bool y_oob = clamp_y_indices(y0i, y1i, ymaxi, ys, ye);
    bool x_oob = x0i < 0 || x0i > xmaxi;
    if (!(y_oob || x_oob)) fill_column(c, x0i, ys, ye, stacked, ymaxi);
  }
  const bool clipped = clipped_start || clipped_end;
  if (dxi == 0 && !clipped) {
    bool y_oob = clamp_y_indices(y3i, y2i, ymaxi, ys, ye);
    bool x_oob = x1i < 0 || x1i > xmaxi;
    if (!(y_oob || x_oob)) fill_column(c, x1i, ys, ye, stacked, ymaxi);
    return;
  }
  dxi = llabs(dxi) * 2;
  dy0i = llabs(dy0i) * 2;
  dy1i = llabs(dy1i) * 2;
  long long error0 = 2 * dy0i - dxi, error1 = 2 * dy1i - dxi;
  while (x0i != x1i) {
    while (error0 >= 0 && (error0 || ix > 0)) { error0 -= 2 * dxi; y0i += iy0; }
    error0 += 2 * dy0i;
    while (error1 >= 0 && (error1 || ix > 0)) { error1 -= 2 * dxi; y1i += iy1; }
    error1 += 2 * dy1i;
    x0i += ix;
    if (x0i < 0 || x0i > xmaxi) continue;
    bool y_oob = clamp_y_indices(y0i, y1i, ymaxi, ys, ye);
    if (!y_oob) fill_column(c, x0i, ys, ye, stacked, ymaxi);
  }
}

template <typename XY, bool RG>
__global__ void __launch_bounds__(128) k_areas(const AreaArgs a) {
  const XY* __restrict__ xs = (const XY*)a.xs;
  const XY* __restrict__ ys0 = (const XY*)a.ys0;
  const XY* __restrict__ ys1 = (const XY*)a.ys1;
  const long long nseg = a.nverts - 1;
  const long long total = RG ? a.rg_xlen - a.rg_x[0] : a.nlines * nseg;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool to_line = ys1 != nullptr;
  // warp-uniform trip count + __syncwarp() per trapezoid: keeps the lanes together over the rounds (see k_lines_axis1)
  for (long long s0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); s0 < total; s0 += stride, __syncwarp()) {
    const long long s = s0 + (threadIdx.x & 31);
    if (s >= total) continue;
    long long i, j, ox, oy, oy1;
    if (!RG) {
      i = s / nseg; j = s - i * nseg;
      ox = i * a.x_line_stride + j; oy = oy1 = i * a.y_line_stride + j;
    } else {
      // perform_extend_area_to_{zero,line}_axis1_ragged, area.py:1959-2004, 2033-2081: the slots are the flat x vertices; the row
      // is the last one that starts at or before the slot and it draws min(x, y0 (, y1) lengths) vertices
      const long long p = a.rg_x[0] + s;
      long long lo = 0, hi = a.nlines - 1;
      while (lo < hi) { const long long mid = (lo + hi + 1) >> 1; if (a.rg_x[mid] <= p) lo = mid; else hi = mid - 1; }
      const bool last = lo + 1 >= a.nlines;
      const long long xb = a.rg_x[lo], yb = a.rg_y[lo];
      long long nv = (last ? a.rg_xlen : a.rg_x[lo + 1]) - xb;
      const long long yn = (last ? a.rg_ylen : a.rg_y[lo + 1]) - yb;
      nv = nv < yn ? nv : yn;
      long long y1b = 0;
      if (to_line) {
        y1b = a.rg_y1[lo];
        const long long y1n = (last ? a.rg_y1len : a.rg_y1[lo + 1]) - y1b;
        nv = nv < y1n ? nv : y1n;
      }
      i = lo; j = p - xb;
      if (j + 1 >= nv) continue;
      ox = p; oy = yb + j; oy1 = y1b + j;
    }
    const double x0 = (double)xs[ox], x1 = (double)xs[ox + 1];
    const double y0 = (double)ys0[oy], y3 = (double)ys0[oy + 1];
    const double y1 = to_line ? (double)ys1[oy1] : 0.0, y2 = to_line ? (double)ys1[oy1 + 1] : 0.0;
    bool trapezoid_start;
    if (j == 0) trapezoid_start = RG ? true : a.plot_start != 0;
    else {
      const double xm = (double)xs[ox - 1], ym = (double)ys0[oy - 1];
      trapezoid_start = (xm != xm) || (ym != ym);
      // the ragged to-line form tests `isnull(y1_flat[y1_start_i + j] - 1)` (area.py:2069): the CURRENT vertex of the second curve
      if (to_line) { const double ym1 = RG ? y1 : (double)ys1[oy1 - 1]; trapezoid_start = trapezoid_start || (ym1 != ym1); }
    }
    AreaCtx c;
    c.plan = &a.plan; c.width = a.v.width;
    c.idx = a.value_per_vertex ? j : i;
    c.row = a.row_offset + c.idx;
    c.cat = 0;
    if (a.plan.ncat > 0) {
      int cc = load_cat(a.plan.cat, a.plan.cat_dtype, c.idx);
      if (cc < 0) cc += a.plan.ncat;
      c.cat = (cc < 0 || cc >= a.plan.ncat) ? -1 : cc;
    }
    draw_trapezoid_y<XY>(a, c, x0, x1, y0, y1, y2, y3, trapezoid_start, /*stacked=*/to_line, /*second_is_xy=*/to_line);
  }
}

static long long area_py_round(double v) { return (long long)nearbyint(v); }

extern "C" int dsb_areas_plan(const dsb_view* view, const void* xs, const void* ys0, const void* ys1, int32_t xy_dtype,
                              int64_t nlines, int64_t nverts, const dsb_line_layout* layout, int64_t row_offset,
                              const dsb_plan* plan, void* stream) {
  if (!view || view->width <= 0 || view->height <= 0) { dsb_set_error("dsb_areas_plan: bad view"); return DSB_ERR_ARG; }
  if (!plan || plan->nops < 1 || plan->nops > DSB_MAX_OPS) { dsb_set_error("dsb_areas_plan: bad plan"); return DSB_ERR_ARG; }
  for (int k = 0; k < plan->nops; k++)
    if (!plan->ops[k].agg || plan->ops[k].op < DSB_OP_COUNT || plan->ops[k].op > DSB_OP_MATCHROW64) {
      dsb_set_error("dsb_areas_plan: bad op %d", k); return DSB_ERR_ARG;
    }
  if (nlines <= 0 || nverts < 2) return DSB_OK;
  if (!xs || !ys0) { dsb_set_error("dsb_areas_plan: null vertex arrays"); return DSB_ERR_ARG; }
  if (xy_dtype != DSB_F32 && xy_dtype != DSB_F64) { dsb_set_error("dsb_areas_plan: xy_dtype must be f32 or f64"); return DSB_ERR_ARG; }
  AreaArgs a;
  a.v = *view; a.xs = xs; a.ys0 = ys0; a.ys1 = ys1; a.nlines = nlines; a.nverts = nverts; a.row_offset = row_offset; a.plan = *plan;
  a.rg_x = a.rg_y = a.rg_y1 = nullptr; a.rg_xlen = a.rg_ylen = a.rg_y1len = 0;
  if (layout && layout->x_starts) {
    if (!layout->y_starts || (ys1 && !layout->y1_starts) || layout->x_flat_len < 0 || layout->y_flat_len < 0 || layout->y1_flat_len < 0) {
      dsb_set_error("dsb_areas_plan: ragged layout needs the start indices and flat lengths of every vertex array"); return DSB_ERR_ARG;
    }
    a.rg_x = (const long long*)layout->x_starts; a.rg_y = (const long long*)layout->y_starts; a.rg_y1 = (const long long*)layout->y1_starts;
    a.rg_xlen = layout->x_flat_len; a.rg_ylen = layout->y_flat_len; a.rg_y1len = layout->y1_flat_len;
    a.x_line_stride = a.y_line_stride = 0; a.value_per_vertex = 0; a.plot_start = 1;
  } else if (layout) {
    if (layout->x_line_stride < 0 || layout->y_line_stride < 0) { dsb_set_error("dsb_areas_plan: negative line stride"); return DSB_ERR_ARG; }
    a.x_line_stride = layout->x_line_stride; a.y_line_stride = layout->y_line_stride;
    a.value_per_vertex = layout->value_per_vertex; a.plot_start = layout->plot_start;
  } else {
    a.x_line_stride = nverts; a.y_line_stride = nverts; a.value_per_vertex = 0; a.plot_start = 1;
  }
  const double mx = view->x_log ? log10(view->xmax) : view->xmax, my = view->y_log ? log10(view->ymax) : view->ymax;
  a.xxmax = area_py_round(mx * view->sx + view->tx);
  a.yymax = area_py_round(my * view->sy + view->ty);
  long long xi = (long long)(mx * view->sx + view->tx), yi = (long long)(my * view->sy + view->ty);   // int() truncation
  a.xmaxi = (xi == a.xxmax) ? xi - 1 : xi;
  a.ymaxi = (yi == a.yymax) ? yi - 1 : yi;
  const bool rg = a.rg_x != nullptr;
  const long long total = rg ? a.rg_xlen : nlines * (nverts - 1);
  if (total <= 0) return DSB_OK;
  const int threads = 128;
  long long want = (total + threads - 1) / threads, cap = (long long)dsb_num_sms() * 16;
  int grid = (int)(want < cap ? want : cap);
  cudaStream_t s = (cudaStream_t)stream;
  dsb_note_kernel("k_areas<%s%s>", xy_dtype == DSB_F32 ? "f32" : "f64", rg ? ", ragged" : "");
  if (xy_dtype == DSB_F32) { if (rg) k_areas<float, true><<<grid, threads, 0, s>>>(a); else k_areas<float, false><<<grid, threads, 0, s>>>(a); }
  else { if (rg) k_areas<double, true><<<grid, threads, 0, s>>>(a); else k_areas<double, false><<<grid, threads, 0, s>>>(a); }
  DSB_CUDA_CHECK_LAUNCH("dsb_areas_plan");
  return DSB_OK;
}
