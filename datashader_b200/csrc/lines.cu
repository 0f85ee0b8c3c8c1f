// Line glyph (LinesAxis1): per-segment Liang-Barsky clip + snapped Bresenham or the full antialiased
// rasteriser, fused with the reduction.  Replaces _build_extend_line_axis1_none_constant.extend_cuda
// (glyphs/line.py:1321-1332), _build_draw_segment (:1033-1097), _build_bresenham (:986-1031) and
// _build_full_antialias (:826-983).  The reference has NO antialiasing on its CUDA path
// (core.py:454-459); here the antialiased single-stage combinations run on the GPU.
//
// All geometry is f64 with unfused multiply/add (-fmad=false semantics via _rn intrinsics where the
// result feeds a floor/ceil/compare), matching the numba CPU arithmetic that the oracle pins.
#include "common.cuh"
#include "accum.cuh"
#include <limits.h>

static int g_lines_balanced = 1;       // antialiased single-stage lines: rows balanced over the warp (k_lines_aa_balanced)
extern "C" int dsb_lines_configure(int balanced) { g_lines_balanced = balanced != 0; return DSB_OK; }

struct LineArgs {
  dsb_view v;
  const void* xs;
  const void* ys;
  long long nlines, nverts;
  const void* val;
  int val_dtype;
  int agg;
  double line_width;
  void* canvas;
  uint8_t* mask;
  long long xxmax, yymax;   // round(mapper(xmax)*sx+tx): map_onto_pixel_snap, line.py:714-715
  long long nx, ny;         // round((xmax-xmin)*sx): draw_segment, line.py:1087-1088
  int overwrite;
  int use_plan;             // line_width == 0 with an accumulator plan (any reduction the point path supports)
  long long row_offset;
  long long x_line_stride, y_line_stride;   // elements between consecutive lines (0: one shared vertex vector)
  int value_per_vertex;     // axis=0 layouts: append(i = row of the segment's first vertex); axis=1: i = line
  int plot_start;           // axis=0: whether vertex 0 starts a line (False for a continued dask partition)
  const void* cat;          // antialiased by(): category code per line / vertex row, canvases are [H, W, ncat]
  int cat_dtype, ncat;
  const long long* rg_x;    // ragged layouts (LinesAxis1Ragged): first flat vertex of every row in xs / ys; NULL = dense
  const long long* rg_y;
  long long rg_xlen, rg_ylen;
  dsb_plan plan;
};

// Which (row i, vertex j) a work slot is, where its vertices live and how many vertices the row has.  Dense layouts: slot t of
// the rows [i0, i0 + nl) is segment t % (nverts - 1) of row i0 + t / (nverts - 1).  Ragged layouts (extend_cpu_numba,
// line.py:1559-1600): the slots are the flat x vertices of those rows; the row is found by bisection of the start indices, it
// draws min(x length, y length) vertices and a slot past its last segment is idle.
struct SegLoc { long long i, j, ox, oy, nv; };

template <bool RG>
__device__ __forceinline__ long long seg_slots(const LineArgs& a, long long i0, long long nl) {
  if (!RG) return nl * (a.nverts - 1);
  return (i0 + nl < a.nlines ? a.rg_x[i0 + nl] : a.rg_xlen) - a.rg_x[i0];
}

template <bool RG>
__device__ __forceinline__ bool seg_locate(const LineArgs& a, long long i0, long long nl, long long t, SegLoc& q) {
  if (!RG) {
    const long long nseg = a.nverts - 1, g = t / nseg;
    q.i = i0 + g; q.j = t - g * nseg; q.nv = a.nverts;
    q.ox = q.i * a.x_line_stride + q.j; q.oy = q.i * a.y_line_stride + q.j;
    return true;
  }
  const long long p = a.rg_x[i0] + t;
  long long lo = i0, hi = i0 + nl - 1;                    // the last row that starts at or before p (empty rows sort before it)
  while (lo < hi) { const long long mid = (lo + hi + 1) >> 1; if (a.rg_x[mid] <= p) lo = mid; else hi = mid - 1; }
  const long long x0 = a.rg_x[lo], y0 = a.rg_y[lo];
  const long long xn = (lo + 1 < a.nlines ? a.rg_x[lo + 1] : a.rg_xlen) - x0, yn = (lo + 1 < a.nlines ? a.rg_y[lo + 1] : a.rg_ylen) - y0;
  q.i = lo; q.j = p - x0; q.nv = xn < yn ? xn : yn; q.ox = p; q.oy = y0 + q.j;
  return q.j + 1 < q.nv;
}

__device__ __forceinline__ double fmul64(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double fadd64(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double fsub64(double a, double b) { return __dadd_rn(a, -b); }

struct LineCtx {
  int agg;
  bool has_field;
  double field;
  bool field_nan;
  long long width;
  void* canvas;
  uint8_t* mask;
  const dsb_plan* plan;     // non-null: run the accumulator plan for every touched pixel
  long long line, row;      // line index within the frame / global row id
  int cat;                  // category of this line (by()), -1 = skip
  unsigned int* touched_n;  // 2-stage antialiasing, global stage-1 canvas: `touched` = the CTA's bitmap of touched cells,
  uint32_t* touched;        //   bbox = this thread's {ymin, ymax} of touched rows; or, when hkeys != nullptr, the
  uint32_t* tlist;          //   tlist / tn: this thread's own list of touched cells (entry k at tlist[k * AA2_THREADS + tid]);
  int* tn;                  //   the bitmap is only used once that list is full
  int* bbox;                //   shared-memory hash table:
  uint32_t* hkeys;          //   [cap] cell ids (0xffffffff = free), hvals [cap] key64 maxima, touched_n = entries in use
  long long* hvals;
  uint32_t hmask;           //   cap - 1
  uint32_t hgroup;          //   (index of the line within its group) << 27, or-ed into the table key
  int ncat;                 // antialiased by(): cell = (y * W + x) * ncat + cat
};

// internal agg codes of the 2-stage path (stage 1 = per-line max of field * aa_factor / of aa_factor)
#define AA2_THREADS 512
#define AA2_LIST_CAP 512       // touched cells a thread can remember per line before it falls back to the bitmap
#define DSB_LINE_AA2_VALUE 101
#define DSB_LINE_AA2_COVER 102

// ---- appends, line_width == 0 (reductions.py _append / _append_no_field) ------------------------
__device__ __forceinline__ void append_px(const LineCtx& c, long long x, long long y) {
  long long cell = y * c.width + x;
  if (c.plan) {
    if (c.plan->ncat > 0) {
      if (c.cat < 0) return;
      cell = cell * c.plan->ncat + c.cat;
    }
    for (int k = 0; k < c.plan->nops; k++) apply_base(c.plan->ops[k], cell, c.line, c.row, c.plan->notes);
    return;
  }
  switch (c.agg) {
    case DSB_LINE_ANY:
      if (c.has_field && c.field_nan) return;
      ((uint8_t*)c.canvas)[cell] = 1;
      return;
    case DSB_LINE_COUNT:
      if (c.has_field && c.field_nan) return;
      atomicAdd((unsigned int*)c.canvas + cell, 1u);
      return;
    case DSB_LINE_SUM:
      if (c.field_nan) return;
      atomicAdd((double*)c.canvas + cell, c.field);
      c.mask[cell] = 1;
      return;
    case DSB_LINE_MAX:
      if (c.field_nan) return;
      atomicMax((long long*)c.canvas + cell, key64_from_f64(c.field));
      return;
    case DSB_LINE_MIN:
      if (c.field_nan) return;
      atomicMin((long long*)c.canvas + cell, key64_from_f64(c.field));
      return;
  }
}

// ---- appends, antialiased (reductions.py _append_antialias / _append_no_field_antialias) --------
__device__ __forceinline__ void append_aa(const LineCtx& c, long long x, long long y, double aa, double prev_aa) {
  long long cell = y * c.width + x;
  if (c.ncat > 0) {                      // by(): the reduction runs on agg[:, :, cat] (compiler.py:379-390)
    if (c.cat < 0) return;
    cell = cell * c.ncat + c.cat;
  }
  switch (c.agg) {
    case DSB_LINE_ANY:     // max of aa_factor, stored as f32 (reductions.py:850-869)
      if (c.has_field && c.field_nan) return;
      atomicMax((int*)c.canvas + cell, key32_from_f32((float)aa));
      return;
    case DSB_LINE_COUNT:   // SUM_1AGG: += aa - prev (reductions.py:560-568, 586-592)
      if (c.has_field && c.field_nan) return;
      atomicAdd((float*)c.canvas + cell, (float)fsub64(aa, prev_aa));
      c.mask[cell] = 1;
      return;
    case DSB_LINE_SUM: {   // reductions.py:1065-1075
      double v = fmul64(c.field, fsub64(aa, prev_aa));
      if (v != v) return;
      atomicAdd((double*)c.canvas + cell, v);
      c.mask[cell] = 1;
      return;
    }
    case DSB_LINE_MAX: {   // reductions.py:1229-1236
      double v = fmul64(c.field, aa);
      if (v != v) return;
      atomicMax((long long*)c.canvas + cell, key64_from_f64(v));
      return;
    }
    case DSB_LINE_MEAN: {  // mean = _sum_zero / _count_ignore_antialiasing (reductions.py:967-975, 681-686, 1289-1297);
                           // `mask` is the u32 count canvas here
      double v = fmul64(c.field, fsub64(aa, prev_aa));
      if (v == v) atomicAdd((double*)c.canvas + cell, v);
      if (!c.field_nan && prev_aa == 0.0) atomicAdd((unsigned int*)c.mask + cell, 1u);
      return;
    }
    case DSB_LINE_AA2_VALUE:   // stage 1 of min / first / last / sum(self_intersect=False): reductions.py:1186-1191,
    case DSB_LINE_AA2_COVER: { // 1408-1413, 1446-1451, 1079-1085; of count(self_intersect=False): :570-578, 594-600
      double v;
      if (c.agg == DSB_LINE_AA2_COVER) {
        if (c.has_field && c.field_nan) return;
        v = (double)(float)aa;             // the reference's stage-1 canvas is float32; rounding is monotone, so max commutes
      } else {
        v = fmul64(c.field, aa);
        if (v != v) return;
      }
      const long long key = key64_from_f64(v);
      if (c.hkeys) {                        // open addressing, linear probing; at most 3/4 full, else the line is redone
        const uint32_t id = c.hgroup | (uint32_t)cell;
        uint32_t h = (id * 2654435761u) >> 7 & c.hmask;
        for (;;) {
          uint32_t k = c.hkeys[h];
          if (k == 0xffffffffu) {
            if (*(volatile unsigned int*)c.touched_n > (c.hmask >> 2) * 3u) { atomicOr(c.touched_n, 0x80000000u); return; }
            k = atomicCAS(c.hkeys + h, 0xffffffffu, id);
            if (k == 0xffffffffu) { atomicAdd(c.touched_n, 1u); k = id; }
          }
          if (k == id) { atomicMax((long long*)c.hvals + h, key); return; }
          h = (h + 1) & c.hmask;
        }
      }
      atomicMax((long long*)c.canvas + cell, key);                 // fire-and-forget RED
      if (*c.tn < AA2_LIST_CAP) {           // remember the cell in this thread's own list: no atomics, no shared state
        c.tlist[(long long)(*c.tn) * AA2_THREADS + threadIdx.x] = (uint32_t)cell;
        (*c.tn)++;
      } else {                              // list full (very long / wide segments): fall back to the CTA's bitmap
        atomicOr(c.touched + (cell >> 5), 1u << (cell & 31));
        c.bbox[0] = min(c.bbox[0], (int)y); c.bbox[1] = max(c.bbox[1], (int)y);
      }
      return;
    }
  }
}

// line.py:783-800
__device__ __forceinline__ bool clipt(double p, double q, double& t0, double& t1) {
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

template <typename XY> __device__ __forceinline__ double seg_delta(double a1, double a0);
// numba types the vertex values as float32 inside _liang_barsky, so x1 - x0 rounds to float32
template <> __device__ __forceinline__ double seg_delta<float>(double a1, double a0) { return (double)__fsub_rn((float)a1, (float)a0); }
template <> __device__ __forceinline__ double seg_delta<double>(double a1, double a0) { return fsub64(a1, a0); }

__device__ __forceinline__ double clampd(double x, double lo, double hi) { return fmax(lo, fmin(x, hi)); }
__device__ __forceinline__ double linearstep(double e0, double e1, double x) {
  const double den = fsub64(e1, e0);
  // line_width <= 1 gives den == 1.0 exactly: x / 1.0 == x, so the f64 division (a ~30-instruction sequence) is skipped
  const double t = (den == 1.0) ? fsub64(x, e0) : __ddiv_rn(fsub64(x, e0), den);
  return clampd(t, 0.0, 1.0);
}
__device__ __forceinline__ double x_intercept(double y, double cx0, double cy0, double cx1, double cy1) {
  if (cy0 == cy1) return cx1;
  double frac = __ddiv_rn(fsub64(y, cy0), fsub64(cy1, cy0));
  return fadd64(cx0, fmul64(frac, fsub64(cx1, cx0)));
}

// line.py:986-1031
__device__ void bresenham(const LineCtx& c, bool segment_start, long long x0, long long x1, long long y0, long long y1,
                          bool clipped) {
  long long dx = x1 - x0;
  long long ix = (dx > 0) - (dx < 0);
  dx = llabs(dx) * 2;
  long long dy = y1 - y0;
  long long iy = (dy > 0) - (dy < 0);
  dy = llabs(dy) * 2;
  if (!clipped && !(dx | dy)) { append_px(c, x0, y0); return; }
  if (segment_start) append_px(c, x0, y0);
  if (dx >= dy) {
    long long error = 2 * dy - dx;
    while (x0 != x1) {
      if (error >= 0 && (error || ix > 0)) { error -= 2 * dx; y0 += iy; }
      error += 2 * dy;
      x0 += ix;
      append_px(c, x0, y0);
    }
  } else {
    long long error = 2 * dx - dy;
    while (y0 != y1) {
      if (error >= 0 && (error || iy > 0)) { error -= 2 * dy; x0 += ix; }
      error += 2 * dx;
      y0 += iy;
      append_px(c, x0, y0);
    }
  }
}

// line.py:830-981
__device__ void full_antialias(const LineCtx& c, double line_width, bool overwrite, double x0, double x1, double y0,
                               double y1, bool segment_start, bool segment_end, double xm, double ym, long long nx,
                               long long ny) {
  if (x0 == x1 && y0 == y1) return;
  const bool flip_xy = fabs(fsub64(x0, x1)) < fabs(fsub64(y0, y1));
  if (flip_xy) {
    double t;
    t = x0; x0 = y0; y0 = t;
    t = x1; x1 = y1; y1 = t;
    t = xm; xm = ym; ym = t;
  }
  double scale = 1.0;
  if (line_width < 1.0) { scale = fmul64(scale, line_width); line_width = 1.0; }
  const double aa = 1.0;
  const double halfwidth = fmul64(0.5, fadd64(line_width, aa));
  const bool flip_order = y1 < y0 || (y1 == y0 && x1 < x0);
  double alongx = fsub64(x1, x0), alongy = fsub64(y1, y0);
  const double length = __dsqrt_rn(fadd64(fmul64(alongx, alongx), fmul64(alongy, alongy)));
  alongx = __ddiv_rn(alongx, length);
  alongy = __ddiv_rn(alongy, length);
  const double rightx = alongy, righty = -alongx;
  double bx[4], by[4];
  if (flip_order) {
    bx[0] = fsub64(x1, fmul64(halfwidth, fsub64(rightx, alongx)));
    bx[1] = fsub64(x1, fmul64(halfwidth, fsub64(-rightx, alongx)));
    bx[2] = fsub64(x0, fmul64(halfwidth, fadd64(-rightx, alongx)));
    bx[3] = fsub64(x0, fmul64(halfwidth, fadd64(rightx, alongx)));
    by[0] = fsub64(y1, fmul64(halfwidth, fsub64(righty, alongy)));
    by[1] = fsub64(y1, fmul64(halfwidth, fsub64(-righty, alongy)));
    by[2] = fsub64(y0, fmul64(halfwidth, fadd64(-righty, alongy)));
    by[3] = fsub64(y0, fmul64(halfwidth, fadd64(righty, alongy)));
  } else {
    bx[0] = fadd64(x0, fmul64(halfwidth, fsub64(rightx, alongx)));
    bx[1] = fadd64(x0, fmul64(halfwidth, fsub64(-rightx, alongx)));
    bx[2] = fadd64(x1, fmul64(halfwidth, fadd64(-rightx, alongx)));
    bx[3] = fadd64(x1, fmul64(halfwidth, fadd64(rightx, alongx)));
    by[0] = fadd64(y0, fmul64(halfwidth, fsub64(righty, alongy)));
    by[1] = fadd64(y0, fmul64(halfwidth, fsub64(-righty, alongy)));
    by[2] = fadd64(y1, fmul64(halfwidth, fadd64(-righty, alongy)));
    by[3] = fadd64(y1, fmul64(halfwidth, fadd64(righty, alongy)));
  }
  long long xmax = nx - 1, ymax = ny - 1;
  if (flip_xy) { long long t = xmax; xmax = ymax; ymax = t; }
  int lowindex;
  if (flip_order) lowindex = x0 > x1 ? 0 : 1;
  else lowindex = x1 > x0 ? 0 : 1;
  double prev_alongx = 0, prev_alongy = 0, prev_length = 0, prev_rightx = 0, prev_righty = 0;
  if (!overwrite && !segment_start) {
    prev_alongx = fsub64(x0, xm);
    prev_alongy = fsub64(y0, ym);
    prev_length = __dsqrt_rn(fadd64(fmul64(prev_alongx, prev_alongx), fmul64(prev_alongy, prev_alongy)));
    if (prev_length > 0.0) {
      prev_alongx = __ddiv_rn(prev_alongx, prev_length);
      prev_alongy = __ddiv_rn(prev_alongy, prev_length);
      prev_rightx = prev_alongy;
      prev_righty = -prev_alongx;
    } else {
      overwrite = true;
    }
  }
  const long long ystart = (long long)clampd(ceil(by[lowindex]), 0.0, (double)ymax);
  const long long yend = (long long)clampd(floor(by[(lowindex + 2) & 3]), 0.0, (double)ymax);
  int ll = lowindex, lu = (ll + 1) & 3, rl = lowindex, ru = (rl + 3) & 3;
  const double e0 = fmul64(0.5, fsub64(line_width, aa));
  for (long long y = ystart; y <= yend; y++) {
    const double yd = (double)y;
    if (ll == lowindex && yd > by[lu]) { ll = lu; lu = (ll + 1) & 3; }
    if (rl == lowindex && yd > by[ru]) { rl = ru; ru = (rl + 3) & 3; }
    const long long xleft = (long long)clampd(ceil(x_intercept(yd, bx[ll], by[ll], bx[lu], by[lu])), 0.0, (double)xmax);
    const long long xright = (long long)clampd(floor(x_intercept(yd, bx[rl], by[rl], bx[ru], by[ru])), 0.0, (double)xmax);
    const double ry0 = fsub64(yd, y0), ry1 = fsub64(yd, y1);
    for (long long x = xleft; x <= xright; x++) {
      const double rx0 = fsub64((double)x, x0);
      const double along = fadd64(fmul64(rx0, alongx), fmul64(ry0, alongy));
      bool prev_correction = false;
      double distance;
      if (along < 0.0) {
        if (overwrite || segment_start || fadd64(fmul64(rx0, prev_alongx), fmul64(ry0, prev_alongy)) > 0.0)
          distance = __dsqrt_rn(fadd64(fmul64(rx0, rx0), fmul64(ry0, ry0)));
        else continue;
      } else if (along > length) {
        if (overwrite || segment_end) {
          const double rx1 = fsub64((double)x, x1);
          distance = __dsqrt_rn(fadd64(fmul64(rx1, rx1), fmul64(ry1, ry1)));
        } else continue;
      } else {
        distance = fabs(fadd64(fmul64(rx0, rightx), fmul64(ry0, righty)));
        if (!overwrite && !segment_start) {
          const double pa = fadd64(fmul64(rx0, prev_alongx), fmul64(ry0, prev_alongy));
          if (-prev_length <= pa && pa <= 0.0 && fabs(fadd64(fmul64(rx0, prev_rightx), fmul64(ry0, prev_righty))) <= halfwidth)
            prev_correction = true;
        }
      }
      double value = fmul64(fsub64(1.0, linearstep(e0, halfwidth, distance)), scale);
      double prev_value = 0.0;
      if (prev_correction) {
        const double prev_distance = fabs(fadd64(fmul64(rx0, prev_rightx), fmul64(ry0, prev_righty)));
        prev_value = fmul64(fsub64(1.0, linearstep(e0, halfwidth, prev_distance)), scale);
        if (value <= prev_value) value = 0.0;
      }
      if (value > 0.0) {
        if (flip_xy) append_aa(c, y, x, value, prev_value);
        else append_aa(c, x, y, value, prev_value);
      }
    }
  }
}

template <typename XY>
__device__ __forceinline__ double map_axis(bool is_log, double v) {
  // after clipping the coordinate is float64 (the clipped value is computed in f64), so the log
  // mapper sees a float64 unless the vertex is an unclipped float32 value: the reference's numba
  // unifies x0 to float64 on return from _liang_barsky, so log10 is always the f64 one here.
  return is_log ? log10(v) : v;
}

// line.py:1045-1097
template <typename XY, bool AA>       // AA: line_width > 0 (a compile-time split keeps the antialiased code's registers out of Bresenham)
__device__ void draw_segment(const LineArgs& a, const LineCtx& c, bool segment_start, bool segment_end, double x0,
                             double x1, double y0, double y1, double xm, double ym) {
  const dsb_view& v = a.v;
  bool skip = (x0 != x0) || (y0 != y0) || (x1 != x1) || (y1 != y1);
  // _liang_barsky, line.py:734-780
  if (x0 < v.xmin && x1 < v.xmin) skip = true;
  else if (x0 > v.xmax && x1 > v.xmax) skip = true;
  else if (y0 < v.ymin && y1 < v.ymin) skip = true;
  else if (y0 > v.ymax && y1 > v.ymax) skip = true;
  double t0 = 0.0, t1 = 1.0;
  const double dx1 = seg_delta<XY>(x1, x0);
  if (!clipt(-dx1, fsub64(x0, v.xmin), t0, t1)) skip = true;
  if (!clipt(dx1, fsub64(v.xmax, x0), t0, t1)) skip = true;
  const double dy1 = seg_delta<XY>(y1, y0);
  if (!clipt(-dy1, fsub64(y0, v.ymin), t0, t1)) skip = true;
  if (!clipt(dy1, fsub64(v.ymax, y0), t0, t1)) skip = true;
  if (skip) return;
  bool clipped_start = false, clipped_end = false;
  if (t1 < 1) { clipped_end = true; x1 = fadd64(x0, fmul64(t1, dx1)); y1 = fadd64(y0, fmul64(t1, dy1)); }
  if (t0 > 0) { clipped_start = true; x0 = fadd64(x0, fmul64(t0, dx1)); y0 = fadd64(y0, fmul64(t0, dy1)); }
  const bool clipped = clipped_start || clipped_end;
  segment_start = segment_start || clipped_start;
  if (AA) {
    // map_onto_pixel_no_snap, line.py:722-726
    const double x0p = fsub64(fadd64(fmul64(map_axis<XY>(v.x_log, x0), v.sx), v.tx), 0.5);
    const double y0p = fsub64(fadd64(fmul64(map_axis<XY>(v.y_log, y0), v.sy), v.ty), 0.5);
    const double x1p = fsub64(fadd64(fmul64(map_axis<XY>(v.x_log, x1), v.sx), v.tx), 0.5);
    const double y1p = fsub64(fadd64(fmul64(map_axis<XY>(v.y_log, y1), v.sy), v.ty), 0.5);
    double xmp = 0.0, ymp = 0.0;
    if (!segment_start) {
      xmp = fsub64(fadd64(fmul64(map_axis<XY>(v.x_log, xm), v.sx), v.tx), 0.5);
      ymp = fsub64(fadd64(fmul64(map_axis<XY>(v.y_log, ym), v.sy), v.ty), 0.5);
    }
    full_antialias(c, a.line_width, a.overwrite != 0, x0p, x1p, y0p, y1p, segment_start, segment_end, xmp, ymp, a.nx, a.ny);
  } else {
    // map_onto_pixel_snap, line.py:689-720
    long long x0i = (long long)__double2ll_rz(fadd64(fmul64(map_axis<XY>(v.x_log, x0), v.sx), v.tx));
    long long y0i = (long long)__double2ll_rz(fadd64(fmul64(map_axis<XY>(v.y_log, y0), v.sy), v.ty));
    long long x1i = (long long)__double2ll_rz(fadd64(fmul64(map_axis<XY>(v.x_log, x1), v.sx), v.tx));
    long long y1i = (long long)__double2ll_rz(fadd64(fmul64(map_axis<XY>(v.y_log, y1), v.sy), v.ty));
    if (x0i == a.xxmax) x0i--;
    if (y0i == a.yymax) y0i--;
    if (x1i == a.xxmax) x1i--;
    if (y1i == a.yymax) y1i--;
    bresenham(c, segment_start, x0i, x1i, y0i, y1i, clipped);
  }
}

// ---- antialiased segments with the rows balanced over the warp ------------------------------------------------------------
// ncu on k_lines_axis1 at config 4 (profiles/r02_lines_aa.md): 6.8 of 32 lanes active on average - a thread scans its own
// segment row by row, and the rows per segment (2 .. 30, |dy| of a random walk) and the pixels per row differ from lane to lane.
// The rows of a scan conversion are independent once the segment's invariants are known (the left / right edge index of a
// row depends only on whether its y lies above the corner that follows the lowest one), so a warp first prepares its 32
// segments (clip, map, corners, unit vectors: uniform work, AaSeg in shared memory), then hands the ROWS of all 32 segments out
// evenly, 32 at a time.  Same pixels, same arithmetic per pixel, same appends as full_antialias.
struct AaSeg {                 // 192 bytes: eight 128-thread CTAs per SM next to 64 registers per thread
  double x0, y0, x1, y1, alongx, alongy, length;         // right = (alongy, -alongx), prev_right likewise: not stored
  double bx[4], by[4];
  double prev_alongx, prev_alongy, prev_length;
  double field;
  long long line;                                         // the global row id is row_offset + line
  int xmax, ymax, ystart, nrows, lowindex, cat;
  unsigned char flip_xy, overwrite, segment_start, segment_end, field_nan, pad[3];
};
static_assert(sizeof(AaSeg) == 192, "AaSeg layout");

// the per-call constants of full_antialias (line.py:836-842): the same for every segment
struct AaConst { double halfwidth, e0, scale; };
__device__ __forceinline__ AaConst aa_const(double line_width) {
  AaConst k;
  k.scale = 1.0;
  if (line_width < 1.0) { k.scale = fmul64(k.scale, line_width); line_width = 1.0; }
  const double aa = 1.0;
  k.halfwidth = fmul64(0.5, fadd64(line_width, aa));
  k.e0 = fmul64(0.5, fsub64(line_width, aa));
  return k;
}

// the invariants of full_antialias (line.py:830-905); g.nrows = 0 when nothing is drawn
__device__ __forceinline__ void aa_prepare(AaSeg& g, double line_width, bool overwrite, double x0, double x1, double y0, double y1,
                                           bool segment_start, bool segment_end, double xm, double ym, long long nx, long long ny) {
  g.nrows = 0;
  if (x0 == x1 && y0 == y1) return;
  const bool flip_xy = fabs(fsub64(x0, x1)) < fabs(fsub64(y0, y1));
  if (flip_xy) {
    double t;
    t = x0; x0 = y0; y0 = t;
    t = x1; x1 = y1; y1 = t;
    t = xm; xm = ym; ym = t;
  }
  double scale = 1.0;
  if (line_width < 1.0) { scale = fmul64(scale, line_width); line_width = 1.0; }
  const double aa = 1.0;
  const double halfwidth = fmul64(0.5, fadd64(line_width, aa));
  const bool flip_order = y1 < y0 || (y1 == y0 && x1 < x0);
  double alongx = fsub64(x1, x0), alongy = fsub64(y1, y0);
  const double length = __dsqrt_rn(fadd64(fmul64(alongx, alongx), fmul64(alongy, alongy)));
  alongx = __ddiv_rn(alongx, length);
  alongy = __ddiv_rn(alongy, length);
  const double rightx = alongy, righty = -alongx;
  if (flip_order) {
    g.bx[0] = fsub64(x1, fmul64(halfwidth, fsub64(rightx, alongx)));
    g.bx[1] = fsub64(x1, fmul64(halfwidth, fsub64(-rightx, alongx)));
    g.bx[2] = fsub64(x0, fmul64(halfwidth, fadd64(-rightx, alongx)));
    g.bx[3] = fsub64(x0, fmul64(halfwidth, fadd64(rightx, alongx)));
    g.by[0] = fsub64(y1, fmul64(halfwidth, fsub64(righty, alongy)));
    g.by[1] = fsub64(y1, fmul64(halfwidth, fsub64(-righty, alongy)));
    g.by[2] = fsub64(y0, fmul64(halfwidth, fadd64(-righty, alongy)));
    g.by[3] = fsub64(y0, fmul64(halfwidth, fadd64(righty, alongy)));
  } else {
    g.bx[0] = fadd64(x0, fmul64(halfwidth, fsub64(rightx, alongx)));
    g.bx[1] = fadd64(x0, fmul64(halfwidth, fsub64(-rightx, alongx)));
    g.bx[2] = fadd64(x1, fmul64(halfwidth, fadd64(-rightx, alongx)));
    g.bx[3] = fadd64(x1, fmul64(halfwidth, fadd64(rightx, alongx)));
    g.by[0] = fadd64(y0, fmul64(halfwidth, fsub64(righty, alongy)));
    g.by[1] = fadd64(y0, fmul64(halfwidth, fsub64(-righty, alongy)));
    g.by[2] = fadd64(y1, fmul64(halfwidth, fadd64(-righty, alongy)));
    g.by[3] = fadd64(y1, fmul64(halfwidth, fadd64(righty, alongy)));
  }
  long long xmax = nx - 1, ymax = ny - 1;
  if (flip_xy) { long long t = xmax; xmax = ymax; ymax = t; }
  int lowindex;
  if (flip_order) lowindex = x0 > x1 ? 0 : 1;
  else lowindex = x1 > x0 ? 0 : 1;
  g.prev_alongx = g.prev_alongy = g.prev_length = 0.0;
  if (!overwrite && !segment_start) {
    double pax = fsub64(x0, xm), pay = fsub64(y0, ym);
    const double pl = __dsqrt_rn(fadd64(fmul64(pax, pax), fmul64(pay, pay)));
    if (pl > 0.0) {
      pax = __ddiv_rn(pax, pl);
      pay = __ddiv_rn(pay, pl);
      g.prev_alongx = pax; g.prev_alongy = pay; g.prev_length = pl;
    } else {
      g.prev_alongx = pax; g.prev_alongy = pay; g.prev_length = pl;
      overwrite = true;
    }
  }
  const long long ystart = (long long)clampd(ceil(g.by[lowindex]), 0.0, (double)ymax);
  const long long yend = (long long)clampd(floor(g.by[(lowindex + 2) & 3]), 0.0, (double)ymax);
  g.x0 = x0; g.y0 = y0; g.x1 = x1; g.y1 = y1; g.alongx = alongx; g.alongy = alongy; g.length = length;
  g.xmax = (int)xmax; g.ymax = (int)ymax; g.ystart = (int)ystart; g.lowindex = lowindex;
  g.flip_xy = flip_xy; g.overwrite = overwrite; g.segment_start = segment_start; g.segment_end = segment_end;
  g.nrows = yend >= ystart ? (int)(yend - ystart + 1) : 0;
}

// one row of the scan conversion (line.py:906-981): its pixel span.  The edge indices are those the reference's running ll / rl
// hold at row y: each advances once, at the first row above the corner that follows the lowest one.
__device__ __forceinline__ void aa_row_span(const AaSeg& g, long long y, long long& xleft, long long& xright) {
  const double yd = (double)y;
  const int lowindex = g.lowindex;
  int ll = lowindex, lu = (ll + 1) & 3, rl = lowindex, ru = (rl + 3) & 3;
  if (yd > g.by[lu]) { ll = lu; lu = (ll + 1) & 3; }
  if (yd > g.by[ru]) { rl = ru; ru = (rl + 3) & 3; }
  xleft = (long long)clampd(ceil(x_intercept(yd, g.bx[ll], g.by[ll], g.bx[lu], g.by[lu])), 0.0, (double)g.xmax);
  xright = (long long)clampd(floor(x_intercept(yd, g.bx[rl], g.by[rl], g.bx[ru], g.by[ru])), 0.0, (double)g.xmax);
}

// one pixel of that span: the body of the reference's inner loop (line.py:925-981)
__device__ __forceinline__ void aa_pixel(const AaSeg& g, const AaConst& k, const LineCtx& c, long long y, long long x) {
  const double yd = (double)y;
  const double rightx = g.alongy, righty = -g.alongx, prev_rightx = g.prev_alongy, prev_righty = -g.prev_alongx;
  const double x0 = g.x0, y0 = g.y0, x1 = g.x1, y1 = g.y1;
  const double ry0 = fsub64(yd, y0), ry1 = fsub64(yd, y1);
  const bool overwrite = g.overwrite, segment_start = g.segment_start, segment_end = g.segment_end;
  const double rx0 = fsub64((double)x, x0);
  const double along = fadd64(fmul64(rx0, g.alongx), fmul64(ry0, g.alongy));
  bool prev_correction = false;
  double distance;
  if (along < 0.0) {
    if (overwrite || segment_start || fadd64(fmul64(rx0, g.prev_alongx), fmul64(ry0, g.prev_alongy)) > 0.0)
      distance = __dsqrt_rn(fadd64(fmul64(rx0, rx0), fmul64(ry0, ry0)));
    else return;
  } else if (along > g.length) {
    if (overwrite || segment_end) {
      const double rx1 = fsub64((double)x, x1);
      distance = __dsqrt_rn(fadd64(fmul64(rx1, rx1), fmul64(ry1, ry1)));
    } else return;
  } else {
    distance = fabs(fadd64(fmul64(rx0, rightx), fmul64(ry0, righty)));
    if (!overwrite && !segment_start) {
      const double pa = fadd64(fmul64(rx0, g.prev_alongx), fmul64(ry0, g.prev_alongy));
      if (-g.prev_length <= pa && pa <= 0.0 && fabs(fadd64(fmul64(rx0, prev_rightx), fmul64(ry0, prev_righty))) <= k.halfwidth)
        prev_correction = true;
    }
  }
  double value = fmul64(fsub64(1.0, linearstep(k.e0, k.halfwidth, distance)), k.scale);
  double prev_value = 0.0;
  if (prev_correction) {
    const double prev_distance = fabs(fadd64(fmul64(rx0, prev_rightx), fmul64(ry0, prev_righty)));
    prev_value = fmul64(fsub64(1.0, linearstep(k.e0, k.halfwidth, prev_distance)), k.scale);
    if (value <= prev_value) value = 0.0;
  }
  if (value > 0.0) {
    if (g.flip_xy) append_aa(c, y, x, value, prev_value);
    else append_aa(c, x, y, value, prev_value);
  }
}

// clip + map of draw_segment (line.py:1045-1085) for the antialiased form: false = nothing to draw
template <typename XY>
__device__ __forceinline__ bool aa_clip_map(const LineArgs& a, bool& segment_start, double& x0, double& x1, double& y0, double& y1,
                                            double& xm, double& ym) {
  const dsb_view& v = a.v;
  bool skip = (x0 != x0) || (y0 != y0) || (x1 != x1) || (y1 != y1);
  if (x0 < v.xmin && x1 < v.xmin) skip = true;
  else if (x0 > v.xmax && x1 > v.xmax) skip = true;
  else if (y0 < v.ymin && y1 < v.ymin) skip = true;
  else if (y0 > v.ymax && y1 > v.ymax) skip = true;
  double t0 = 0.0, t1 = 1.0;
  const double dx1 = seg_delta<XY>(x1, x0);
  if (!clipt(-dx1, fsub64(x0, v.xmin), t0, t1)) skip = true;
  if (!clipt(dx1, fsub64(v.xmax, x0), t0, t1)) skip = true;
  const double dy1 = seg_delta<XY>(y1, y0);
  if (!clipt(-dy1, fsub64(y0, v.ymin), t0, t1)) skip = true;
  if (!clipt(dy1, fsub64(v.ymax, y0), t0, t1)) skip = true;
  if (skip) return false;
  bool clipped_start = false;
  if (t1 < 1) { x1 = fadd64(x0, fmul64(t1, dx1)); y1 = fadd64(y0, fmul64(t1, dy1)); }
  if (t0 > 0) { clipped_start = true; x0 = fadd64(x0, fmul64(t0, dx1)); y0 = fadd64(y0, fmul64(t0, dy1)); }
  segment_start = segment_start || clipped_start;
  x0 = fsub64(fadd64(fmul64(map_axis<XY>(v.x_log, x0), v.sx), v.tx), 0.5);
  y0 = fsub64(fadd64(fmul64(map_axis<XY>(v.y_log, y0), v.sy), v.ty), 0.5);
  x1 = fsub64(fadd64(fmul64(map_axis<XY>(v.x_log, x1), v.sx), v.tx), 0.5);
  y1 = fsub64(fadd64(fmul64(map_axis<XY>(v.y_log, y1), v.sy), v.ty), 0.5);
  if (!segment_start) {
    xm = fsub64(fadd64(fmul64(map_axis<XY>(v.x_log, xm), v.sx), v.tx), 0.5);
    ym = fsub64(fadd64(fmul64(map_axis<XY>(v.y_log, ym), v.sy), v.ty), 0.5);
  } else { xm = 0.0; ym = 0.0; }
  return true;
}

template <typename XY, bool RG>
__global__ void __launch_bounds__(128, 8) k_lines_aa_balanced(const LineArgs a) {
  __shared__ AaSeg segs[128];
  __shared__ int prefix[128];
  const AaConst kaa = aa_const(a.line_width);
  const XY* __restrict__ xs = (const XY*)a.xs;
  const XY* __restrict__ ys = (const XY*)a.ys;
  const long long total = seg_slots<RG>(a, 0, a.nlines);
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31, w0 = threadIdx.x & ~31;
  AaSeg* const wseg = segs + w0;
  int* const wpre = prefix + w0;
  // every warp walks whole batches of 32 consecutive segments; the loop bound is warp-uniform
  for (long long s0 = (long long)blockIdx.x * blockDim.x + w0; s0 < total; s0 += stride) {
    const long long s = s0 + lane;
    AaSeg& g = wseg[lane];
    g.nrows = 0;
    SegLoc q;
    if (s < total && seg_locate<RG>(a, 0, a.nlines, s, q)) {
      const long long i = q.i, j = q.j, ox = q.ox, oy = q.oy;
      double x0 = (double)xs[ox], y0 = (double)ys[oy], x1 = (double)xs[ox + 1], y1 = (double)ys[oy + 1];
      bool segment_start = (j == 0) ? (a.plot_start != 0) : false;
      double xm = 0.0, ym = 0.0;
      if (j > 0) {
        xm = (double)xs[ox - 1]; ym = (double)ys[oy - 1];
        segment_start = (xm != xm) || (ym != ym);
        if (segment_start) { xm = 0.0; ym = 0.0; }
      }
      bool segment_end = (j == q.nv - 2);
      if (!segment_end) {
        const double xn = (double)xs[ox + 2], yn = (double)ys[oy + 2];
        segment_end = (xn != xn) || (yn != yn);
      }
      const long long vi = a.value_per_vertex ? j : i;
      if (aa_clip_map<XY>(a, segment_start, x0, x1, y0, y1, xm, ym))
        aa_prepare(g, a.line_width, a.overwrite != 0, x0, x1, y0, y1, segment_start, segment_end, xm, ym, a.nx, a.ny);
      g.field = a.val_dtype != DSB_NONE ? load_f64(a.val, a.val_dtype, vi) : 0.0;
      g.field_nan = a.val_dtype != DSB_NONE && (g.field != g.field);
      g.line = vi; g.cat = 0;
      if (a.ncat > 0) {
        int cc = load_cat(a.cat, a.cat_dtype, vi);
        if (cc < 0) cc += a.ncat;
        g.cat = (cc < 0 || cc >= a.ncat) ? -1 : cc;
      }
    }
    // inclusive prefix of the row counts over the warp
    int incl = g.nrows;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    wpre[lane] = incl;
    const int rows_total = __shfl_sync(0xffffffffu, incl, 31);
    __syncwarp();
    // Rounds of 32 rows, warp-uniform trip counts, a __syncwarp() per round (without it the lanes drift apart over the rounds:
    // ncu showed 5.9 active lanes).  Inside a round every lane works out the pixel span of ITS row; a second prefix sum over
    // the 32 span lengths numbers the pixels of the round, and lane i draws pixels i, i + 32, ... whatever row they belong to
    // (a row of a steep segment is 2-3 pixels, of a flat one 10+: per-row loops left 12.7 of 32 lanes active).
    for (int r0 = 0; r0 < rows_total; r0 += 32) {
      const int r = r0 + lane;
      int segi = 0, yrow = 0, xl = 0, cnt = 0;
      if (r < rows_total) {
        int lo = 0, hi = 31;                     // first lane whose inclusive prefix exceeds r
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (wpre[mid] > r) hi = mid; else lo = mid + 1; }
        const AaSeg& q = wseg[lo];
        const long long y = (long long)q.ystart + (r - (wpre[lo] - q.nrows));
        long long xleft, xright;
        aa_row_span(q, y, xleft, xright);
        segi = lo; yrow = (int)y; xl = (int)xleft; cnt = xright >= xleft ? (int)(xright - xleft + 1) : 0;
      }
      int incl2 = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl2, o); if (lane >= o) incl2 += t; }
      const int npix = __shfl_sync(0xffffffffu, incl2, 31);
      const int excl2 = incl2 - cnt;
      for (int p0 = 0; p0 < npix; p0 += 32) {
        const int p = min(p0 + lane, npix - 1);
        int lo = 0, hi = 31;                     // first lane whose inclusive pixel prefix exceeds p: exactly five halvings
#pragma unroll
        for (int it = 0; it < 5; it++) {
          const int mid = (lo + hi) >> 1;
          if (__shfl_sync(0xffffffffu, incl2, mid) > p) hi = mid; else lo = mid + 1;
        }
        const int ssegi = __shfl_sync(0xffffffffu, segi, lo), sy = __shfl_sync(0xffffffffu, yrow, lo);
        const int sxl = __shfl_sync(0xffffffffu, xl, lo), sex = __shfl_sync(0xffffffffu, excl2, lo);
        if (p0 + lane < npix) {
          const AaSeg& q = wseg[ssegi];
          LineCtx c;
          c.agg = a.agg; c.has_field = a.val_dtype != DSB_NONE; c.width = a.v.width; c.canvas = a.canvas; c.mask = a.mask;
          c.field = q.field; c.field_nan = q.field_nan != 0; c.plan = nullptr; c.line = q.line; c.row = a.row_offset + q.line;
          c.cat = q.cat; c.ncat = a.ncat; c.hkeys = nullptr; c.touched = nullptr;
          aa_pixel(q, kaa, c, (long long)sy, (long long)(sxl + (p - sex)));
        }
      }
      __syncwarp();
    }
  }
}

// one thread per (line, segment): extend_cuda, line.py:1321-1332 + perform_extend_line :1250-1275
template <typename XY, bool AA, bool RG>
__global__ void __launch_bounds__(128) k_lines_axis1(const LineArgs a) {
  const XY* __restrict__ xs = (const XY*)a.xs;
  const XY* __restrict__ ys = (const XY*)a.ys;
  const long long total = seg_slots<RG>(a, 0, a.nlines);
  const long long stride = (long long)gridDim.x * blockDim.x;
  // warp-uniform trip count + __syncwarp() per segment: the pixel loops of the 32 segments differ in length, and without a
  // reconvergence point per round the lanes drift apart over the rounds (ncu: 6.8 active lanes, profiles/r02_lines_aa.md)
  for (long long s0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); s0 < total; s0 += stride, __syncwarp()) {
    const long long s = s0 + (threadIdx.x & 31);
    if (s >= total) continue;
    long long i, j, ox, oy, nv;
    if (!RG) {
      const long long nseg = a.nverts - 1;
      i = s / nseg; j = s - i * nseg; ox = i * a.x_line_stride + j; oy = i * a.y_line_stride + j; nv = a.nverts;
    } else {
      SegLoc q;
      if (!seg_locate<true>(a, 0, a.nlines, s, q)) continue;
      i = q.i; j = q.j; ox = q.ox; oy = q.oy; nv = q.nv;
    }
    const double x0 = (double)xs[ox], y0 = (double)ys[oy], x1 = (double)xs[ox + 1], y1 = (double)ys[oy + 1];
    bool segment_start = (j == 0) ? (a.plot_start != 0) : false;
    double xm = 0.0, ym = 0.0;
    if (j > 0) {
      xm = (double)xs[ox - 1]; ym = (double)ys[oy - 1];
      segment_start = (xm != xm) || (ym != ym);
      if (segment_start) { xm = 0.0; ym = 0.0; }
    }
    bool segment_end = (j == nv - 2);
    if (!segment_end) {
      const double xn = (double)xs[ox + 2], yn = (double)ys[oy + 2];
      segment_end = (xn != xn) || (yn != yn);
    }
    const long long vi = a.value_per_vertex ? j : i;      // the index the reference hands to append()
    LineCtx c;
    c.agg = a.agg; c.has_field = a.val_dtype != DSB_NONE; c.width = a.v.width; c.canvas = a.canvas; c.mask = a.mask;
    c.field = c.has_field ? load_f64(a.val, a.val_dtype, vi) : 0.0;
    c.field_nan = c.has_field && (c.field != c.field);
    c.plan = a.use_plan ? &a.plan : nullptr;
    c.line = vi; c.row = a.row_offset + vi; c.cat = 0; c.ncat = 0;
    c.hkeys = nullptr; c.touched = nullptr;
    if (!a.use_plan && a.ncat > 0) {
      int cc = load_cat(a.cat, a.cat_dtype, vi);
      if (cc < 0) cc += a.ncat;
      c.cat = (cc < 0 || cc >= a.ncat) ? -1 : cc;
      c.ncat = a.ncat;
    }
    if (a.use_plan && a.plan.ncat > 0) {
      int cc = load_cat(a.plan.cat, a.plan.cat_dtype, vi);
      if (cc < 0) cc += a.plan.ncat;
      c.cat = (cc < 0 || cc >= a.plan.ncat) ? -1 : cc;
    }
    draw_segment<XY, AA>(a, c, segment_start, segment_end, x0, x1, y0, y1, xm, ym);
  }
}


// ---- 2-stage antialiased lines -------------------------------------------------------------------------------------
// compiler.py:198-268 + line.py:1291-1319: for min / first / last and count / sum with self_intersect=False the
// reference renders every line on its own into a cleared canvas with a max() combination (stage 1) and folds that
// canvas into the result with nansum / nanmin / nanfirst / nanlast (stage 2), one line after the other.
// Here: one CTA per line.  Stage 1 goes into the CTA's private full-size key64 canvas (`temp`, all EMPTY between
// lines) with atomicMax; the first touch of a cell appends it to the CTA's touched list.  Stage 2 walks that list: it
// resets the cell and applies the line's value to the shared result with the commutative form of the combine
// (sum: atomic add; min: atomicMin on key64; first / last: atomicMin / atomicMax of the global line index in phase 1,
// and in phase 2 - after every line has voted - the winning line alone stores its value).
struct Aa2Args {
  int combo, phase;
  void* out;
  void* aux;              // sum / count: u8 mask; first / last: i64 line-index canvas
  long long* temp;        // [nctas][ncell]  (global stage-1 canvases, used for the lines that overflow the hash table)
  uint32_t* touched;      // [nctas][nwords] bitmaps of the cells the CTA's current line has touched (nwords = ceil(ncell / 32))
  uint32_t* tlists;       // [nctas][AA2_LIST_CAP][AA2_THREADS] per-thread lists of touched cells
  unsigned int* stats;    // [2] groups tried / groups overflowed by the hash pass (it gives up when most overflow)
  unsigned int* redo_n;   // number of / indices of the lines whose touched-pixel set overflowed the shared-memory table
  unsigned int* redo;     // [nlines]
};

#define AA2_HASH_CAP 16384   // 16384 x (4 + 8) B = 192 KB of shared memory: one 512-thread CTA per SM
#define AA2_CELL_BITS 27      // table key = (line within its group) << 27 | cell: canvases below 2^27 pixels, groups <= 32

// first / last in ONE rasterisation (phase 3): out holds {line index, value bits} pairs, replaced as a whole with a
// 128-bit compare-and-swap when this line precedes (follows) the one recorded - the two-phase form rasterises every line twice
__device__ __forceinline__ void aa2_pair_update(void* out, uint32_t cell, long long line, long long key, bool first) {
  unsigned long long* p = (unsigned long long*)out + 2 * (size_t)cell;
  unsigned long long cl, cv;
  asm volatile("ld.global.relaxed.gpu.v2.u64 {%0, %1}, [%2];" : "=l"(cl), "=l"(cv) : "l"(p) : "memory");
  const unsigned long long nv = (unsigned long long)__double_as_longlong(f64_from_key64(key));
  for (;;) {
    if (first ? line >= (long long)cl : line <= (long long)cl) return;
    unsigned long long ol, ov;
    asm volatile("{\n.reg .b128 c, n, o;\nmov.b128 c, {%2, %3};\nmov.b128 n, {%4, %5};\n"
                 "atom.global.relaxed.gpu.cas.b128 o, [%6], c, n;\nmov.b128 {%0, %1}, o;\n}"
                 : "=l"(ol), "=l"(ov) : "l"(cl), "l"(cv), "l"((unsigned long long)line), "l"(nv), "l"(p) : "memory");
    if (ol == cl && ov == cv) return;
    cl = ol; cv = ov;
  }
}

// where(min | max) in one rasterisation: {key64, line index} pairs; a line replaces the pair when its value is better, or
// equal with a lower line index (the reference's strict compare keeps the earlier line, reductions.py:2009-2016)
__device__ __forceinline__ void aa2_pair_update_arg(void* out, uint32_t cell, long long line, long long key, bool is_max) {
  unsigned long long* p = (unsigned long long*)out + 2 * (size_t)cell;
  unsigned long long ck, cl;
  asm volatile("ld.global.relaxed.gpu.v2.u64 {%0, %1}, [%2];" : "=l"(ck), "=l"(cl) : "l"(p) : "memory");
  for (;;) {
    const long long k0 = (long long)ck;
    const bool better = is_max ? key > k0 : key < k0;
    if (!(better || (key == k0 && line < (long long)cl))) return;
    unsigned long long ok, ol;
    asm volatile("{\n.reg .b128 c, n, o;\nmov.b128 c, {%2, %3};\nmov.b128 n, {%4, %5};\n"
                 "atom.global.relaxed.gpu.cas.b128 o, [%6], c, n;\nmov.b128 {%0, %1}, o;\n}"
                 : "=l"(ok), "=l"(ol) : "l"(ck), "l"(cl), "l"((unsigned long long)key), "l"((unsigned long long)line), "l"(p) : "memory");
    if (ok == ck && ol == cl) return;
    ck = ok; cl = ol;
  }
}

__device__ __forceinline__ void aa2_stage2(const Aa2Args& b, uint32_t cell, long long key, long long line) {
  switch (b.combo) {
    case DSB_AA2_SUM:      // nansum_in_place, utils.py:885-897 (the NaN start is restored from the mask afterwards)
      atomicAdd((double*)b.out + cell, f64_from_key64(key));
      ((uint8_t*)b.aux)[cell] = 1;
      break;
    case DSB_AA2_COUNT:
      atomicAdd((float*)b.out + cell, (float)f64_from_key64(key));
      ((uint8_t*)b.aux)[cell] = 1;
      break;
    case DSB_AA2_MIN:      // nanmin_in_place, utils.py:655-668
      atomicMin((long long*)b.out + cell, key);
      break;
    case DSB_AA2_ARGMIN:   // where(min(col), ...): the selector's combine keeps the earlier line on ties (strict compare)
      if (b.phase == 3) aa2_pair_update_arg(b.out, cell, line, key, false);
      else if (b.phase == 1) atomicMin((long long*)b.out + cell, key);
      else if (((const long long*)b.out)[cell] == key) atomicMin((long long*)b.aux + cell, line);
      break;
    case DSB_AA2_ARGMAX:   // where(max(col), ...)
      if (b.phase == 3) aa2_pair_update_arg(b.out, cell, line, key, true);
      else if (b.phase == 1) atomicMax((long long*)b.out + cell, key);
      else if (((const long long*)b.out)[cell] == key) atomicMin((long long*)b.aux + cell, line);
      break;
    case DSB_AA2_FIRST:    // nanfirst_in_place, utils.py:615-623: the lowest line index that touches the pixel wins
      if (b.phase == 3) aa2_pair_update(b.out, cell, line, key, true);
      else if (b.phase == 1) atomicMin((long long*)b.aux + cell, line);
      else if (((const long long*)b.aux)[cell] == line) ((double*)b.out)[cell] = f64_from_key64(key);
      break;
    case DSB_AA2_LAST:     // nanlast_in_place, utils.py:627-635
      if (b.phase == 3) aa2_pair_update(b.out, cell, line, key, false);
      else if (b.phase == 1) atomicMax((long long*)b.aux + cell, line);
      else if (((const long long*)b.aux)[cell] == line) ((double*)b.out)[cell] = f64_from_key64(key);
      break;
  }
}

// HASH: stage 1 in a shared-memory hash table ((line, cell) -> max key64).  A CTA takes a group of G consecutive lines
// at a time (G = 1 for long lines; short lines are batched so that all threads have a segment) and flushes the table
// after each group.  Groups that fill more than 3/4 of the table are queued in b.redo and handled, line by line, by a
// second launch with HASH = false, whose stage 1 is the CTA's private full-size global canvas.
template <typename XY, bool HASH, bool RG>
__global__ void __launch_bounds__(AA2_THREADS, 1) k_lines_aa2(const LineArgs a, const Aa2Args b, const int G) {
  extern __shared__ long long aa2_smem[];
  __shared__ unsigned int s_touched;
  long long* hvals = aa2_smem;
  uint32_t* hkeys = (uint32_t*)(aa2_smem + AA2_HASH_CAP);
  const XY* __restrict__ xs = (const XY*)a.xs;
  const XY* __restrict__ ys = (const XY*)a.ys;
  const long long ncell = (long long)a.v.width * a.v.height;
  const long long nwords = (ncell + 31) >> 5;
  long long* temp = HASH ? nullptr : b.temp + (long long)blockIdx.x * ncell;
  uint32_t* touched = HASH ? nullptr : b.touched + (long long)blockIdx.x * nwords;
  uint32_t* tlist = HASH ? nullptr : b.tlists + (long long)blockIdx.x * AA2_LIST_CAP * AA2_THREADS;
  __shared__ int s_bbox[2];
  __shared__ int s_skip;
  __shared__ int s_prefix[AA2_THREADS];
  const long long nwork = HASH ? (a.nlines + G - 1) / G : (long long)*b.redo_n;
  if (HASH) {
    for (int k = threadIdx.x; k < AA2_HASH_CAP; k += blockDim.x) { hkeys[k] = 0xffffffffu; hvals[k] = LLONG_MIN; }
  }
  for (long long w = blockIdx.x; w < nwork; w += gridDim.x) {
    const long long i0 = HASH ? w * G : (long long)b.redo[w];
    const long long glines = HASH ? (a.nlines - i0 < G ? a.nlines - i0 : (long long)G) : 1;
    if (threadIdx.x == 0) {
      s_touched = 0; s_bbox[0] = INT_MAX; s_bbox[1] = -1;
      if (HASH) {
        const unsigned int tried = *(volatile unsigned int*)b.stats, failed = *(volatile unsigned int*)(b.stats + 1);
        s_skip = tried >= 64u && failed * 2u > tried;
      }
    }
    __syncthreads();
    if (HASH && s_skip) {
      // most groups so far did not fit the table (long lines): stop trying, hand the rest to the global path
      if (threadIdx.x < glines) b.redo[atomicAdd(b.redo_n, 1u)] = (unsigned int)(i0 + threadIdx.x);
      __syncthreads();
      continue;
    }
    int bbox[2] = {INT_MAX, -1};
    int tn = 0;
    const long long nslots = seg_slots<RG>(a, i0, glines);
    if (!HASH) {
      // global stage-1 path (long lines): the rows of the warp's 32 segments are handed out evenly, as in k_lines_aa_balanced
      const AaConst kaa = aa_const(a.line_width);
      AaSeg* const wseg = (AaSeg*)aa2_smem + (threadIdx.x & ~31);
      int* const wpre = s_prefix + (threadIdx.x & ~31);
      const int lane = threadIdx.x & 31;
      for (long long t0 = threadIdx.x & ~31; t0 < nslots; t0 += blockDim.x) {
        const long long t = t0 + lane;
        AaSeg& g = wseg[lane];
        g.nrows = 0;
        SegLoc q;
        if (t < nslots && seg_locate<RG>(a, i0, glines, t, q)) {
          const long long i = q.i, j = q.j, ox = q.ox, oy = q.oy;
          double x0 = (double)xs[ox], y0 = (double)ys[oy], x1 = (double)xs[ox + 1], y1 = (double)ys[oy + 1];
          bool segment_start = (j == 0) ? (a.plot_start != 0) : false;
          if (j > 0) {
            const double xm = (double)xs[ox - 1], ym = (double)ys[oy - 1];
            segment_start = (xm != xm) || (ym != ym);
          }
          bool segment_end = (j == q.nv - 2);
          if (!segment_end) {
            const double xn = (double)xs[ox + 2], yn = (double)ys[oy + 2];
            segment_end = (xn != xn) || (yn != yn);
          }
          const long long vi = a.value_per_vertex ? j : i;
          double xm0 = 0.0, ym0 = 0.0;          // xm = ym = 0 in 2-stage mode (line.py:1266-1268); unused: overwrite is True
          if (aa_clip_map<XY>(a, segment_start, x0, x1, y0, y1, xm0, ym0))
            aa_prepare(g, a.line_width, true, x0, x1, y0, y1, segment_start, segment_end, 0.0, 0.0, a.nx, a.ny);
          g.field = a.val_dtype != DSB_NONE ? load_f64(a.val, a.val_dtype, vi) : 0.0;
          g.field_nan = a.val_dtype != DSB_NONE && (g.field != g.field);
          g.line = vi; g.cat = 0;
        }
        int incl = g.nrows;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int tt = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += tt; }
        wpre[lane] = incl;
        const int rows_total = __shfl_sync(0xffffffffu, incl, 31);
        __syncwarp();
        for (int r0 = 0; r0 < rows_total; r0 += 32) {            // rows, then the pixels of those rows, spread over the warp
          const int r = r0 + lane;                                 //   (see k_lines_aa_balanced)
          int segi = 0, yrow = 0, xl = 0, cnt = 0;
          if (r < rows_total) {
            int lo = 0, hi = 31;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (wpre[mid] > r) hi = mid; else lo = mid + 1; }
            const AaSeg& q = wseg[lo];
            const long long y = (long long)q.ystart + (r - (wpre[lo] - q.nrows));
            long long xleft, xright;
            aa_row_span(q, y, xleft, xright);
            segi = lo; yrow = (int)y; xl = (int)xleft; cnt = xright >= xleft ? (int)(xright - xleft + 1) : 0;
          }
          int incl2 = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const int tt = __shfl_up_sync(0xffffffffu, incl2, o); if (lane >= o) incl2 += tt; }
          const int npix = __shfl_sync(0xffffffffu, incl2, 31);
          const int excl2 = incl2 - cnt;
          for (int p0 = 0; p0 < npix; p0 += 32) {
            const int p = min(p0 + lane, npix - 1);
            int lo = 0, hi = 31;
#pragma unroll
            for (int it = 0; it < 5; it++) {
              const int mid = (lo + hi) >> 1;
              if (__shfl_sync(0xffffffffu, incl2, mid) > p) hi = mid; else lo = mid + 1;
            }
            const int ssegi = __shfl_sync(0xffffffffu, segi, lo), sy = __shfl_sync(0xffffffffu, yrow, lo);
            const int sxl = __shfl_sync(0xffffffffu, xl, lo), sex = __shfl_sync(0xffffffffu, excl2, lo);
            if (p0 + lane < npix) {
              const AaSeg& q = wseg[ssegi];
              LineCtx c;
              c.agg = a.agg; c.has_field = a.val_dtype != DSB_NONE; c.width = a.v.width; c.canvas = temp; c.mask = nullptr;
              c.field = q.field; c.field_nan = q.field_nan != 0; c.plan = nullptr; c.line = q.line; c.row = a.row_offset + q.line; c.cat = 0; c.ncat = 0;
              c.touched_n = &s_touched; c.touched = touched; c.bbox = bbox; c.tlist = tlist; c.tn = &tn;
              c.hkeys = nullptr; c.hvals = hvals; c.hmask = AA2_HASH_CAP - 1; c.hgroup = 0;
              aa_pixel(q, kaa, c, (long long)sy, (long long)(sxl + (p - sex)));
            }
          }
          __syncwarp();
        }
      }
    } else
    for (long long t = threadIdx.x; t < nslots; t += blockDim.x) {
      SegLoc q;
      if (!seg_locate<RG>(a, i0, glines, t, q)) continue;
      const long long i = q.i, j = q.j, ox = q.ox, oy = q.oy, g = i - i0;
      const double x0 = (double)xs[ox], y0 = (double)ys[oy], x1 = (double)xs[ox + 1], y1 = (double)ys[oy + 1];
      bool segment_start = (j == 0) ? (a.plot_start != 0) : false;
      if (j > 0) {
        const double xm = (double)xs[ox - 1], ym = (double)ys[oy - 1];
        segment_start = (xm != xm) || (ym != ym);
      }
      bool segment_end = (j == q.nv - 2);
      if (!segment_end) {
        const double xn = (double)xs[ox + 2], yn = (double)ys[oy + 2];
        segment_end = (xn != xn) || (yn != yn);
      }
      const long long vi = a.value_per_vertex ? j : i;
      LineCtx c;
      c.agg = a.agg; c.has_field = a.val_dtype != DSB_NONE; c.width = a.v.width; c.canvas = temp; c.mask = nullptr;
      c.field = c.has_field ? load_f64(a.val, a.val_dtype, vi) : 0.0;
      c.field_nan = c.has_field && (c.field != c.field);
      c.plan = nullptr; c.line = vi; c.row = a.row_offset + vi; c.cat = 0; c.ncat = 0;
      c.touched_n = &s_touched; c.touched = touched; c.bbox = bbox; c.tlist = tlist; c.tn = &tn;
      c.hkeys = HASH ? hkeys : nullptr; c.hvals = hvals; c.hmask = AA2_HASH_CAP - 1; c.hgroup = (uint32_t)g << AA2_CELL_BITS;
      // xm = ym = 0 in 2-stage mode (line.py:1266-1268); unused because overwrite is True
      draw_segment<XY, true>(a, c, segment_start, segment_end, x0, x1, y0, y1, 0.0, 0.0);
    }
    if (!HASH && bbox[1] >= 0) { atomicMin(&s_bbox[0], bbox[0]); atomicMax(&s_bbox[1], bbox[1]); }
    __syncthreads();
    const unsigned int n = s_touched;
    if (HASH) {
      const bool overflow = (n & 0x80000000u) != 0;
      if (overflow && threadIdx.x < glines) b.redo[atomicAdd(b.redo_n, 1u)] = (unsigned int)(i0 + threadIdx.x);
      if (threadIdx.x == 0) { atomicAdd(b.stats, 1u); if (overflow) atomicAdd(b.stats + 1, 1u); }
      if (n != 0) {
        for (int k = threadIdx.x; k < AA2_HASH_CAP; k += blockDim.x) {
          const uint32_t id = hkeys[k];
          if (id == 0xffffffffu) continue;
          const long long key = hvals[k];
          hkeys[k] = 0xffffffffu; hvals[k] = LLONG_MIN;
          // global line index: what "first" / "last" order by
          if (!overflow) aa2_stage2(b, id & ((1u << AA2_CELL_BITS) - 1u), key, a.row_offset + i0 + (id >> AA2_CELL_BITS));
        }
      }
    } else {
      // this thread's own touched cells: whoever exchanges a cell's stage-1 value out first folds it in (adjacent segments of
      // a line share pixels, so a cell can sit in several lists); four exchanges in flight
      for (int k = 0; k < tn; k += 4) {
        uint32_t cells[4];
        long long keys[4];
#pragma unroll
        for (int u = 0; u < 4; u++) cells[u] = (k + u < tn) ? tlist[(long long)(k + u) * AA2_THREADS + threadIdx.x] : 0xffffffffu;
#pragma unroll
        for (int u = 0; u < 4; u++)
          keys[u] = cells[u] != 0xffffffffu ? (long long)atomicExch((unsigned long long*)(temp + cells[u]), (unsigned long long)LLONG_MIN) : LLONG_MIN;
#pragma unroll
        for (int u = 0; u < 4; u++) if (keys[u] != LLONG_MIN) aa2_stage2(b, cells[u], keys[u], a.row_offset + i0);
      }
      __syncthreads();
    }
    if (!HASH && s_bbox[1] >= 0) {
      // some thread's list overflowed: walk the bitmap words of those rows; a set bit is a cell of this line unless a
      // list already took its value
      const long long w0 = ((long long)s_bbox[0] * a.v.width) >> 5, w1 = (((long long)s_bbox[1] + 1) * a.v.width - 1) >> 5;
      for (long long wi = w0 + threadIdx.x; wi <= w1; wi += blockDim.x) {
        uint32_t bits = __ldcg(touched + wi);          // written by L2 REDs: do not trust a stale L1 line
        if (!bits) continue;
        touched[wi] = 0;
        while (bits) {
          const int bit = __ffs(bits) - 1;
          bits &= bits - 1;
          const uint32_t cell = (uint32_t)(wi << 5) + bit;
          const long long key = __ldcg(temp + cell);
          if (key == LLONG_MIN) continue;
          temp[cell] = LLONG_MIN;
          aa2_stage2(b, cell, key, a.row_offset + i0);
        }
      }
    }
    __syncthreads();
  }
}

// canvases of 2^27 pixels and more do not fit the table key: every line is queued for the global stage-1 path
__global__ void k_aa2_queue_all(unsigned int* redo_n, unsigned int* redo, unsigned int nlines) {
  for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < nlines; i += gridDim.x * blockDim.x) redo[i] = i;
  if (blockIdx.x == 0 && threadIdx.x == 0) *redo_n = nlines;
}

__global__ void k_fill_i64(long long* p, long long v, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

static long long py_round(double v) { return (long long)nearbyint(v); }   // Python round(): half to even

static int launch_lines(LineArgs& a, int32_t xy_dtype, void* stream, const char* what) {
  const dsb_view* view = &a.v;
  const double mx = view->x_log ? log10(view->xmax) : view->xmax, my = view->y_log ? log10(view->ymax) : view->ymax;
  a.xxmax = py_round(mx * view->sx + view->tx);
  a.yymax = py_round(my * view->sy + view->ty);
  a.nx = py_round((view->xmax - view->xmin) * view->sx);
  a.ny = py_round((view->ymax - view->ymin) * view->sy);
  const bool rg = a.rg_x != nullptr;
  const long long total = rg ? a.rg_xlen : a.nlines * (a.nverts - 1);
  if (total <= 0) return DSB_OK;
  const int threads = 128;
  long long want = (total + threads - 1) / threads;
  long long cap = (long long)dsb_num_sms() * 16;
  int grid = (int)(want < cap ? want : cap);
  cudaStream_t s = (cudaStream_t)stream;
  const bool balanced = g_lines_balanced && a.line_width > 0.0 && !a.use_plan;
  if (xy_dtype != DSB_F32 && xy_dtype != DSB_F64) { dsb_set_error("%s: xy_dtype must be f32 or f64", what); return DSB_ERR_ARG; }
  dsb_note_kernel(balanced ? "k_lines_aa_balanced<%s%s>" : "k_lines_axis1<%s%s>", xy_dtype == DSB_F32 ? "f32" : "f64",
                  rg ? ", ragged" : "");
  const bool f32 = xy_dtype == DSB_F32, aa = a.line_width > 0.0;
#define DSB_LINES_GO(RG)                                                                              \
  do {                                                                                                \
    if (balanced && f32) k_lines_aa_balanced<float, RG><<<grid, threads, 0, s>>>(a);                  \
    else if (balanced) k_lines_aa_balanced<double, RG><<<grid, threads, 0, s>>>(a);                   \
    else if (f32 && aa) k_lines_axis1<float, true, RG><<<grid, threads, 0, s>>>(a);                   \
    else if (f32) k_lines_axis1<float, false, RG><<<grid, threads, 0, s>>>(a);                        \
    else if (aa) k_lines_axis1<double, true, RG><<<grid, threads, 0, s>>>(a);                         \
    else k_lines_axis1<double, false, RG><<<grid, threads, 0, s>>>(a);                                \
  } while (0)
  if (rg) DSB_LINES_GO(true); else DSB_LINES_GO(false);
#undef DSB_LINES_GO
  DSB_CUDA_CHECK_LAUNCH(what);
  return DSB_OK;
}

// Bresenham lines with a full accumulator plan: any reduction Canvas.points supports, applied to every pixel a
// line touches, with i = the line's row (the reference passes the row index i to append, line.py:1046-1097).
static int apply_layout(LineArgs& a, const dsb_line_layout* L, const char* what) {
  a.rg_x = a.rg_y = nullptr; a.rg_xlen = a.rg_ylen = 0;
  if (!L) {   // LinesAxis1 default: dense [nlines, nverts]
    a.x_line_stride = a.nverts; a.y_line_stride = a.nverts; a.value_per_vertex = 0; a.plot_start = 1;
    return DSB_OK;
  }
  if (L->x_starts) {   // LinesAxis1Ragged: flat vertex arrays + a start index per row
    if (!L->y_starts || L->x_flat_len < 0 || L->y_flat_len < 0) { dsb_set_error("%s: ragged layout needs y_starts and the flat lengths", what); return DSB_ERR_ARG; }
    a.rg_x = (const long long*)L->x_starts; a.rg_y = (const long long*)L->y_starts;
    a.rg_xlen = L->x_flat_len; a.rg_ylen = L->y_flat_len;
    a.x_line_stride = a.y_line_stride = 0; a.value_per_vertex = 0; a.plot_start = 1;
    return DSB_OK;
  }
  if (L->x_line_stride < 0 || L->y_line_stride < 0) { dsb_set_error("%s: negative line stride", what); return DSB_ERR_ARG; }
  a.x_line_stride = L->x_line_stride; a.y_line_stride = L->y_line_stride;
  a.value_per_vertex = L->value_per_vertex; a.plot_start = L->plot_start;
  return DSB_OK;
}

extern "C" int dsb_lines_axis1_plan(const dsb_view* view, const void* xs, const void* ys, int32_t xy_dtype,
                                    int64_t nlines, int64_t nverts, const dsb_line_layout* layout, int64_t row_offset,
                                    const dsb_plan* plan, void* stream) {
  if (!view || view->width <= 0 || view->height <= 0) { dsb_set_error("dsb_lines_axis1_plan: bad view"); return DSB_ERR_ARG; }
  if (!plan || plan->nops < 1 || plan->nops > DSB_MAX_OPS) { dsb_set_error("dsb_lines_axis1_plan: bad plan"); return DSB_ERR_ARG; }
  for (int k = 0; k < plan->nops; k++)
    if (!plan->ops[k].agg || plan->ops[k].op < DSB_OP_COUNT || plan->ops[k].op > DSB_OP_MATCHROW64) {
      dsb_set_error("dsb_lines_axis1_plan: bad op %d", k); return DSB_ERR_ARG;
    }
  if (nlines <= 0 || nverts < 2) return DSB_OK;
  if (!xs || !ys) { dsb_set_error("dsb_lines_axis1_plan: null vertex arrays"); return DSB_ERR_ARG; }
  LineArgs a;
  a.v = *view; a.xs = xs; a.ys = ys; a.nlines = nlines; a.nverts = nverts; a.val = nullptr; a.val_dtype = DSB_NONE;
  a.agg = 0; a.line_width = 0.0; a.canvas = nullptr; a.mask = nullptr; a.overwrite = 1;
  a.use_plan = 1; a.row_offset = row_offset; a.plan = *plan; a.cat = nullptr; a.cat_dtype = DSB_NONE; a.ncat = 0;
  int rc = apply_layout(a, layout, "dsb_lines_axis1_plan");
  if (rc != DSB_OK) return rc;
  return launch_lines(a, xy_dtype, stream, "dsb_lines_axis1_plan");
}

extern "C" int dsb_lines_axis1(const dsb_view* view, const void* xs, const void* ys, int32_t xy_dtype, int64_t nlines,
                               int64_t nverts, const dsb_line_layout* layout, const void* val, int32_t val_dtype,
                               int32_t agg, double line_width, void* canvas, uint8_t* mask, void* stream) {
  return dsb_lines_axis1_cat(view, xs, ys, xy_dtype, nlines, nverts, layout, val, val_dtype, agg, line_width, canvas, mask,
                             nullptr, DSB_NONE, 0, stream);
}

extern "C" int dsb_lines_axis1_cat(const dsb_view* view, const void* xs, const void* ys, int32_t xy_dtype, int64_t nlines,
                                   int64_t nverts, const dsb_line_layout* layout, const void* val, int32_t val_dtype,
                                   int32_t agg, double line_width, void* canvas, uint8_t* mask, const void* cat,
                                   int32_t cat_dtype, int32_t ncat, void* stream) {
  if (ncat < 0 || (ncat > 0 && (!cat || cat_dtype == DSB_NONE))) { dsb_set_error("dsb_lines_axis1_cat: bad category column"); return DSB_ERR_ARG; }
  if (ncat > 0 && !(line_width > 0.0)) { dsb_set_error("dsb_lines_axis1_cat: categories are for the antialiased form (use dsb_lines_axis1_plan)"); return DSB_ERR_UNSUPPORTED; }
  if (!view || view->width <= 0 || view->height <= 0 || !canvas) { dsb_set_error("dsb_lines_axis1: bad view/canvas"); return DSB_ERR_ARG; }
  if (agg < DSB_LINE_ANY || agg > DSB_LINE_MEAN_2STAGE) { dsb_set_error("dsb_lines_axis1: unknown agg %d", agg); return DSB_ERR_ARG; }
  const bool aa = line_width > 0.0;
  // mean's bases next to a 2-stage reduction: the same appends, drawn in overwrite mode (antialias.py:47-56: no SUM_1AGG left)
  const bool mean_overwrite = agg == DSB_LINE_MEAN_2STAGE;
  if (mean_overwrite) agg = DSB_LINE_MEAN;
  if (aa && agg == DSB_LINE_MIN) { dsb_set_error("dsb_lines_axis1: antialiased min needs the 2-stage combine"); return DSB_ERR_UNSUPPORTED; }
  if (agg == DSB_LINE_MEAN && !aa) { dsb_set_error("dsb_lines_axis1: mean is the antialiased form only (use dsb_lines_axis1_plan)"); return DSB_ERR_UNSUPPORTED; }
  if ((agg == DSB_LINE_SUM || agg == DSB_LINE_MAX || agg == DSB_LINE_MIN || agg == DSB_LINE_MEAN) && (val_dtype == DSB_NONE || !val)) {
    dsb_set_error("dsb_lines_axis1: this reduction needs a value column"); return DSB_ERR_ARG;
  }
  if ((agg == DSB_LINE_SUM || agg == DSB_LINE_MEAN || (aa && agg == DSB_LINE_COUNT)) && !mask) { dsb_set_error("dsb_lines_axis1: mask canvas required"); return DSB_ERR_ARG; }
  if (nlines <= 0 || nverts < 2) return DSB_OK;
  if (!xs || !ys) { dsb_set_error("dsb_lines_axis1: null vertex arrays"); return DSB_ERR_ARG; }
  LineArgs a;
  a.v = *view; a.xs = xs; a.ys = ys; a.nlines = nlines; a.nverts = nverts; a.val = val; a.val_dtype = val_dtype;
  a.agg = agg; a.line_width = line_width; a.canvas = canvas; a.mask = mask;
  a.overwrite = mean_overwrite || !(agg == DSB_LINE_COUNT || agg == DSB_LINE_SUM || agg == DSB_LINE_MEAN);   // antialias.py:47-56
  a.use_plan = 0; a.row_offset = 0; a.cat = cat; a.cat_dtype = cat_dtype; a.ncat = ncat;
  int rc = apply_layout(a, layout, "dsb_lines_axis1");
  if (rc != DSB_OK) return rc;
  return launch_lines(a, xy_dtype, stream, "dsb_lines_axis1");
}

extern "C" int dsb_lines_aa2(const dsb_view* view, const void* xs, const void* ys, int32_t xy_dtype, int64_t nlines,
                             int64_t nverts, const dsb_line_layout* layout, int64_t row_offset, const void* val,
                             int32_t val_dtype, int32_t combo, int32_t phase, double line_width, void* out, void* aux,
                             void* scratch, int64_t scratch_bytes, void* stream) {
  if (!view || view->width <= 0 || view->height <= 0 || !out) { dsb_set_error("dsb_lines_aa2: bad view/canvas"); return DSB_ERR_ARG; }
  if (combo < DSB_AA2_SUM || combo > DSB_AA2_ARGMAX) { dsb_set_error("dsb_lines_aa2: unknown combination %d", combo); return DSB_ERR_ARG; }
  if (!(line_width > 0.0)) { dsb_set_error("dsb_lines_aa2: line_width must be > 0"); return DSB_ERR_ARG; }
  if (combo != DSB_AA2_COUNT && (val_dtype == DSB_NONE || !val)) { dsb_set_error("dsb_lines_aa2: this reduction needs a value column"); return DSB_ERR_ARG; }
  const bool two_phase = combo == DSB_AA2_FIRST || combo == DSB_AA2_LAST || combo == DSB_AA2_ARGMIN || combo == DSB_AA2_ARGMAX;
  const bool fused = two_phase && phase == 3;
  if (combo != DSB_AA2_MIN && !((combo == DSB_AA2_ARGMIN || combo == DSB_AA2_ARGMAX) && phase == 1) && !fused && !aux) { dsb_set_error("dsb_lines_aa2: aux canvas required"); return DSB_ERR_ARG; }
  if (two_phase && phase != 1 && phase != 2 && !fused) { dsb_set_error("dsb_lines_aa2: phase must be 1 or 2 (3: first / last in one pass)"); return DSB_ERR_ARG; }
  if (fused && (((uintptr_t)out) & 15) != 0) { dsb_set_error("dsb_lines_aa2: the pair canvas must be 16-byte aligned"); return DSB_ERR_ARG; }
  if (nlines <= 0 || nverts < 2) return DSB_OK;
  if (!xs || !ys) { dsb_set_error("dsb_lines_aa2: null vertex arrays"); return DSB_ERR_ARG; }
  const long long ncell = (long long)view->width * view->height;
  if (ncell >= (1LL << 32)) { dsb_set_error("dsb_lines_aa2: canvas too large"); return DSB_ERR_UNSUPPORTED; }
  if (!scratch) { dsb_set_error("dsb_lines_aa2: scratch required"); return DSB_ERR_ARG; }
  const long long per_cta = ncell * 8 + ((ncell + 31) >> 5) * 4 + (long long)AA2_LIST_CAP * AA2_THREADS * 4;
  const long long cap = (long long)dsb_num_sms();
  long long nctas;
  LineArgs a;
  a.v = *view; a.xs = xs; a.ys = ys; a.nlines = nlines; a.nverts = nverts; a.val = val; a.val_dtype = val_dtype;
  a.agg = combo == DSB_AA2_COUNT ? DSB_LINE_AA2_COVER : DSB_LINE_AA2_VALUE;
  a.line_width = line_width; a.canvas = nullptr; a.mask = nullptr; a.overwrite = 1; a.use_plan = 0; a.row_offset = row_offset;
  a.cat = nullptr; a.cat_dtype = DSB_NONE; a.ncat = 0;
  int rc = apply_layout(a, layout, "dsb_lines_aa2");
  if (rc != DSB_OK) return rc;
  const double mx = view->x_log ? log10(view->xmax) : view->xmax, my = view->y_log ? log10(view->ymax) : view->ymax;
  a.xxmax = py_round(mx * view->sx + view->tx);
  a.yymax = py_round(my * view->sy + view->ty);
  a.nx = py_round((view->xmax - view->xmin) * view->sx);
  a.ny = py_round((view->ymax - view->ymin) * view->sy);
  if (nlines >= (1LL << 31)) { dsb_set_error("dsb_lines_aa2: too many lines"); return DSB_ERR_UNSUPPORTED; }
  Aa2Args b;
  b.combo = combo; b.phase = phase; b.out = out; b.aux = aux;
  // scratch layout: [nctas][ncell] i64 stage-1 canvases, [nctas][nwords] u32 touched bitmaps, [nctas][512][512] u32
  // per-thread touched lists, [4 + nlines] u32 queue
  const long long nwords = (ncell + 31) >> 5;
  const long long redo_bytes = 4 * (4 + nlines);
  nctas = (scratch_bytes - redo_bytes) / per_cta;
  if (nctas < 1) { dsb_set_error("dsb_lines_aa2: scratch must hold at least %lld bytes", per_cta + redo_bytes); return DSB_ERR_ARG; }
  if (nctas > cap) nctas = cap;
  if (nctas > nlines) nctas = nlines;
  b.temp = (long long*)scratch;
  b.touched = (uint32_t*)((char*)scratch + nctas * ncell * 8);
  b.tlists = b.touched + nctas * nwords;
  b.redo_n = (unsigned int*)((char*)scratch + nctas * (ncell * 8 + nwords * 4 + (long long)AA2_LIST_CAP * AA2_THREADS * 4));
  b.stats = b.redo_n + 1;
  b.redo = b.redo_n + 4;
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(b.touched, 0, (size_t)(nctas * nwords * 4), s);           // bitmaps
  cudaMemsetAsync(b.redo_n, 0, 16, s);                                        // queue header
  dsb_note_kernel("k_lines_aa2<%s>", xy_dtype == DSB_F32 ? "f32" : "f64");
  k_fill_i64<<<dsb_num_sms() * 8, 256, 0, s>>>(b.temp, LLONG_MIN, nctas * ncell);
  const size_t smem = (size_t)AA2_HASH_CAP * 12;
  // short lines are batched G to a group so that every thread of the CTA has a segment (at most 32 lines per group)
  const bool rg = a.rg_x != nullptr;
  long long segs_per_line = rg ? a.rg_xlen / nlines : nverts - 1;       // ragged: the average row
  long long G = AA2_THREADS / (segs_per_line < 1 ? 1 : segs_per_line);
  G = G < 1 ? 1 : (G > 32 ? 32 : G);
  const bool use_hash = ncell < (1LL << AA2_CELL_BITS) - 1;
  if (!use_hash) G = 1;
  long long hgrid = (long long)dsb_num_sms();
  if (hgrid > (nlines + G - 1) / G) hgrid = (nlines + G - 1) / G;
  if (xy_dtype != DSB_F32 && xy_dtype != DSB_F64) { dsb_set_error("dsb_lines_aa2: xy_dtype must be f32 or f64"); return DSB_ERR_ARG; }
#define DSB_AA2_GO(XY, RG)                                                                                                          \
  do {                                                                                                                              \
    cudaFuncSetAttribute(k_lines_aa2<XY, true, RG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                        \
    if (use_hash) k_lines_aa2<XY, true, RG><<<(int)hgrid, AA2_THREADS, smem, s>>>(a, b, (int)G);                                    \
    else k_aa2_queue_all<<<dsb_num_sms(), 256, 0, s>>>(b.redo_n, b.redo, (unsigned int)nlines);                                     \
    cudaFuncSetAttribute(k_lines_aa2<XY, false, RG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(AA2_THREADS * sizeof(AaSeg))); \
    k_lines_aa2<XY, false, RG><<<(int)nctas, AA2_THREADS, AA2_THREADS * sizeof(AaSeg), s>>>(a, b, 1);                               \
  } while (0)
  if (xy_dtype == DSB_F32) { if (rg) DSB_AA2_GO(float, true); else DSB_AA2_GO(float, false); }
  else { if (rg) DSB_AA2_GO(double, true); else DSB_AA2_GO(double, false); }
#undef DSB_AA2_GO
  DSB_CUDA_CHECK_LAUNCH("dsb_lines_aa2");
  return DSB_OK;
}
