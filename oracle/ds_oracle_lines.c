/* TEST INFRASTRUCTURE ONLY - see ds_oracle.h.  Line glyph restated from the reference's numba CPU
 * path: Liang-Barsky clip, snapped Bresenham, full antialiased rasteriser, LinesAxis1 layout.
 * Build with -ffp-contract=off.
 */
#include "ds_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { ORA_AA2_VALUE = 101, ORA_AA2_COVER = 102,     /* internal: stage-1 appends of the 2-stage antialias path */
       ORA_AA_COUNT_IGNORE = 103 };                  /* _count_ignore_antialiasing (mean's denominator) */

typedef struct {
  int32_t agg_op;     /* ORA_ANY / ORA_COUNT / ORA_SUM / ORA_MAX / ORA_MIN */
  int antialias;
  int has_field;
  double field;       /* value of the current line */
  int64_t width;
  void* agg;
} line_ctx;

/* append for line_width == 0: the plain reduction appends (same as points) */
static inline void append_px(const line_ctx* c, int64_t x, int64_t y) {
  int64_t cell = y * c->width + x;
  double f = c->field;
  switch (c->agg_op) {
    case ORA_ANY:   /* reductions.py:843-848 / 858-862 */
      if (c->has_field && isnan(f)) return;
      ((uint8_t*)c->agg)[cell] = 1; return;
    case ORA_COUNT: /* reductions.py:552-558 / 580-584 */
      if (c->has_field && isnan(f)) return;
      ((uint32_t*)c->agg)[cell] += 1u; return;
    case ORA_SUM: { /* reductions.py:1054-1063 */
      if (isnan(f)) return;
      double* a = (double*)c->agg + cell;
      if (isnan(*a)) *a = f; else *a += f;
      return;
    }
    case ORA_MAX: { /* reductions.py:1222-1227 */
      double* a = (double*)c->agg + cell;
      if (!isnan(f) && (isnan(*a) || *a < f)) *a = f;
      return;
    }
    case ORA_MIN: { /* reductions.py:1178-1183 */
      double* a = (double*)c->agg + cell;
      if (!isnan(f) && (isnan(*a) || *a > f)) *a = f;
      return;
    }
  }
}

/* append for antialiased lines (single-stage combinations only) */
static inline void append_aa(const line_ctx* c, int64_t x, int64_t y, double aa, double prev_aa) {
  int64_t cell = y * c->width + x;
  double f = c->field;
  switch (c->agg_op) {
    case ORA_ANY: { /* reductions.py:850-856 / 864-869, f32 canvas */
      if (c->has_field && isnan(f)) return;
      float* a = (float*)c->agg + cell;
      if (isnan(*a) || aa > *a) *a = (float)aa;
      return;
    }
    case ORA_COUNT: { /* reductions.py:560-568 / 586-592, f32 canvas, SUM_1AGG */
      if (c->has_field && isnan(f)) return;
      float* a = (float*)c->agg + cell;
      if (isnan(*a)) *a = (float)(aa - prev_aa); else *a = (float)(*a + (aa - prev_aa));
      return;
    }
    case ORA_SUM: { /* reductions.py:1065-1075 */
      double v = f * (aa - prev_aa);
      if (isnan(v)) return;
      double* a = (double*)c->agg + cell;
      if (isnan(*a)) *a = v; else *a += v;
      return;
    }
    case ORA_MAX:   /* reductions.py:1229-1236 */
    case ORA_AA2_VALUE: { /* stage 1 of min / first / last / sum(self_intersect=False): the same max of field * aa_factor
                             (reductions.py:1186-1191, 1408-1413, 1446-1451, 1079-1085) */
      double v = f * aa;
      double* a = (double*)c->agg + cell;
      if (!isnan(v) && (isnan(*a) || v > *a)) *a = v;
      return;
    }
    case ORA_AA_COUNT_IGNORE: { /* reductions.py:681-686: u32, += 1 where this segment is the first of its line to touch */
      if (!isnan(f) && prev_aa == 0.0) ((uint32_t*)c->agg)[cell] += 1u;
      return;
    }
    case ORA_AA2_COVER: { /* stage 1 of count(self_intersect=False): max of aa_factor in a float32 canvas
                             (reductions.py:570-578, 594-600) */
      if (c->has_field && isnan(f)) return;
      float* a = (float*)c->agg + cell;
      if (isnan(*a) || aa > *a) *a = (float)aa;
      return;
    }
  }
}

/* line.py:783-800 */
static inline int clipt(double p, double q, double* t0, double* t1) {
  if (p < 0 && q < 0) {
    double r = q / p;
    if (r > *t1) return 0;
    else if (r > *t0) *t0 = r;
  } else if (p > 0 && q < p) {
    double r = q / p;
    if (r < *t0) return 0;
    else if (r < *t1) *t1 = r;
  } else if (q < 0) {
    return 0;
  }
  return 1;
}

/* line.py:734-780; returns skip */
/* With float32 vertex arrays numba types x0..y1 as float32 inside _liang_barsky, so dx1 = x1 - x0 and
   dy1 = y1 - y0 are rounded to float32 before being widened (checked against the reference:
   tests/golden/lines.npz, clipped f32 segments); everything else is float64. */
static inline int liang_barsky(double xmin, double xmax, double ymin, double ymax, double* x0, double* x1,
                               double* y0, double* y1, int skip, int is_f32, int* clipped_start, int* clipped_end) {
  if (*x0 < xmin && *x1 < xmin) skip = 1;
  else if (*x0 > xmax && *x1 > xmax) skip = 1;
  else if (*y0 < ymin && *y1 < ymin) skip = 1;
  else if (*y0 > ymax && *y1 > ymax) skip = 1;
  double t0 = 0, t1 = 1;
  double dx1 = is_f32 ? (double)((float)*x1 - (float)*x0) : *x1 - *x0;
  if (!clipt(-dx1, *x0 - xmin, &t0, &t1)) skip = 1;
  if (!clipt(dx1, xmax - *x0, &t0, &t1)) skip = 1;
  double dy1 = is_f32 ? (double)((float)*y1 - (float)*y0) : *y1 - *y0;
  if (!clipt(-dy1, *y0 - ymin, &t0, &t1)) skip = 1;
  if (!clipt(dy1, ymax - *y0, &t0, &t1)) skip = 1;
  if (t1 < 1) { *clipped_end = 1; *x1 = *x0 + t1 * dx1; *y1 = *y0 + t1 * dy1; }
  else *clipped_end = 0;
  if (t0 > 0) { *clipped_start = 1; *x0 = *x0 + t0 * dx1; *y0 = *y0 + t0 * dy1; }
  else *clipped_start = 0;
  return skip;
}

static inline double axmap(int is_log, double v) { return is_log ? log10(v) : v; }
static inline double clampd(double x, double lo, double hi) { return fmax(lo, fmin(x, hi)); }       /* line.py:803-806 */
static inline double linearstep(double e0, double e1, double x) { return clampd((x - e0) / (e1 - e0), 0.0, 1.0); }
static inline double x_intercept(double y, double cx0, double cy0, double cx1, double cy1) {      /* line.py:815-823 */
  if (cy0 == cy1) return cx1;
  double frac = (y - cy0) / (cy1 - cy0);
  return cx0 + frac * (cx1 - cx0);
}

/* line.py:986-1031 */
static void bresenham(const line_ctx* c, int segment_start, int64_t x0, int64_t x1, int64_t y0, int64_t y1, int clipped) {
  int64_t dx = x1 - x0;
  int64_t ix = (dx > 0) - (dx < 0);
  dx = llabs(dx) * 2;
  int64_t dy = y1 - y0;
  int64_t iy = (dy > 0) - (dy < 0);
  dy = llabs(dy) * 2;
  if (!clipped && !(dx | dy)) { append_px(c, x0, y0); return; }
  if (segment_start) append_px(c, x0, y0);
  if (dx >= dy) {
    int64_t error = 2 * dy - dx;
    while (x0 != x1) {
      if (error >= 0 && (error || ix > 0)) { error -= 2 * dx; y0 += iy; }
      error += 2 * dy;
      x0 += ix;
      append_px(c, x0, y0);
    }
  } else {
    int64_t error = 2 * dx - dy;
    while (y0 != y1) {
      if (error >= 0 && (error || iy > 0)) { error -= 2 * dy; x0 += ix; }
      error += 2 * dx;
      y0 += iy;
      append_px(c, x0, y0);
    }
  }
}

/* line.py:830-981 */
static void full_antialias(const line_ctx* c, double line_width, int overwrite, double x0, double x1, double y0,
                           double y1, int segment_start, int segment_end, double xm, double ym, int64_t nx, int64_t ny) {
  if (x0 == x1 && y0 == y1) return;
  int flip_xy = fabs(x0 - x1) < fabs(y0 - y1);
  if (flip_xy) {
    double t;
    t = x0; x0 = y0; y0 = t;
    t = x1; x1 = y1; y1 = t;
    t = xm; xm = ym; ym = t;
  }
  double scale = 1.0;
  if (line_width < 1.0) { scale *= line_width; line_width = 1.0; }
  double aa = 1.0;
  double halfwidth = 0.5 * (line_width + aa);
  int flip_order = y1 < y0 || (y1 == y0 && x1 < x0);
  double alongx = x1 - x0, alongy = y1 - y0;
  double length = sqrt(alongx * alongx + alongy * alongy);
  alongx /= length; alongy /= length;
  double rightx = alongy, righty = -alongx;
  double b[8];
  if (flip_order) {
    b[0] = x1 - halfwidth * (rightx - alongx);
    b[1] = x1 - halfwidth * (-rightx - alongx);
    b[2] = x0 - halfwidth * (-rightx + alongx);
    b[3] = x0 - halfwidth * (rightx + alongx);
    b[4] = y1 - halfwidth * (righty - alongy);
    b[5] = y1 - halfwidth * (-righty - alongy);
    b[6] = y0 - halfwidth * (-righty + alongy);
    b[7] = y0 - halfwidth * (righty + alongy);
  } else {
    b[0] = x0 + halfwidth * (rightx - alongx);
    b[1] = x0 + halfwidth * (-rightx - alongx);
    b[2] = x1 + halfwidth * (-rightx + alongx);
    b[3] = x1 + halfwidth * (rightx + alongx);
    b[4] = y0 + halfwidth * (righty - alongy);
    b[5] = y0 + halfwidth * (-righty - alongy);
    b[6] = y1 + halfwidth * (-righty + alongy);
    b[7] = y1 + halfwidth * (righty + alongy);
  }
  int64_t xmax = nx - 1, ymax = ny - 1;
  if (flip_xy) { int64_t t = xmax; xmax = ymax; ymax = t; }
  int lowindex;
  if (flip_order) lowindex = x0 > x1 ? 0 : 1;
  else lowindex = x1 > x0 ? 0 : 1;
  double prev_alongx = 0, prev_alongy = 0, prev_length = 0, prev_rightx = 0, prev_righty = 0;
  if (!overwrite && !segment_start) {
    prev_alongx = x0 - xm;
    prev_alongy = y0 - ym;
    prev_length = sqrt(prev_alongx * prev_alongx + prev_alongy * prev_alongy);
    if (prev_length > 0.0) {
      prev_alongx /= prev_length; prev_alongy /= prev_length;
      prev_rightx = prev_alongy; prev_righty = -prev_alongx;
    } else {
      overwrite = 1;
    }
  }
  int64_t ystart = (int64_t)clampd(ceil(b[4 + lowindex]), 0, (double)ymax);
  int64_t yend = (int64_t)clampd(floor(b[4 + (lowindex + 2) % 4]), 0, (double)ymax);
  int ll = lowindex, lu = (ll + 1) % 4, rl = lowindex, ru = (rl + 3) % 4;
  for (int64_t y = ystart; y <= yend; y++) {
    if (ll == lowindex && y > b[4 + lu]) { ll = lu; lu = (ll + 1) % 4; }
    if (rl == lowindex && y > b[4 + ru]) { rl = ru; ru = (rl + 3) % 4; }
    int64_t xleft = (int64_t)clampd(ceil(x_intercept((double)y, b[ll], b[4 + ll], b[lu], b[4 + lu])), 0, (double)xmax);
    int64_t xright = (int64_t)clampd(floor(x_intercept((double)y, b[rl], b[4 + rl], b[ru], b[4 + ru])), 0, (double)xmax);
    for (int64_t x = xleft; x <= xright; x++) {
      double along = (x - x0) * alongx + (y - y0) * alongy;
      int prev_correction = 0;
      double distance;
      if (along < 0.0) {
        if (overwrite || segment_start || (x - x0) * prev_alongx + (y - y0) * prev_alongy > 0.0)
          distance = sqrt((x - x0) * (x - x0) + (y - y0) * (y - y0));
        else continue;
      } else if (along > length) {
        if (overwrite || segment_end) distance = sqrt((x - x1) * (x - x1) + (y - y1) * (y - y1));
        else continue;
      } else {
        distance = fabs((x - x0) * rightx + (y - y0) * righty);
        if (!overwrite && !segment_start) {
          double pa = (x - x0) * prev_alongx + (y - y0) * prev_alongy;
          if (-prev_length <= pa && pa <= 0.0 && fabs((x - x0) * prev_rightx + (y - y0) * prev_righty) <= halfwidth)
            prev_correction = 1;
        }
      }
      double value = 1.0 - linearstep(0.5 * (line_width - aa), halfwidth, distance);
      value *= scale;
      double prev_value = 0.0;
      if (prev_correction) {
        double prev_distance = fabs((x - x0) * prev_rightx + (y - y0) * prev_righty);
        prev_value = 1.0 - linearstep(0.5 * (line_width - aa), halfwidth, prev_distance);
        prev_value *= scale;
        if (value <= prev_value) value = 0.0;
      }
      if (value > 0.0) {
        if (flip_xy) append_aa(c, y, x, value, prev_value);
        else append_aa(c, x, y, value, prev_value);
      }
    }
  }
}

/* line.py:1045-1097 */
static void draw_segment(const ora_view* v, const line_ctx* c, double line_width, int overwrite, int segment_start,
                         int segment_end, double x0, double x1, double y0, double y1, double xm, double ym, int is_f32) {
  int skip = 0;
  if (isnan(x0) || isnan(y0) || isnan(x1) || isnan(y1)) skip = 1;
  int clipped_start, clipped_end;
  skip = liang_barsky(v->xmin, v->xmax, v->ymin, v->ymax, &x0, &x1, &y0, &y1, skip, is_f32, &clipped_start, &clipped_end);
  if (skip) return;
  int clipped = clipped_start || clipped_end;
  segment_start = segment_start || clipped_start;
  if (line_width > 0.0) {
    /* map_onto_pixel_no_snap, line.py:722-726 */
    double x0p = axmap(v->x_log, x0) * v->sx + v->tx - 0.5, y0p = axmap(v->y_log, y0) * v->sy + v->ty - 0.5;
    double x1p = axmap(v->x_log, x1) * v->sx + v->tx - 0.5, y1p = axmap(v->y_log, y1) * v->sy + v->ty - 0.5;
    double xmp = 0.0, ymp = 0.0;
    if (!segment_start) {
      xmp = axmap(v->x_log, xm) * v->sx + v->tx - 0.5;
      ymp = axmap(v->y_log, ym) * v->sy + v->ty - 0.5;
    }
    int64_t nx = (int64_t)nearbyint((v->xmax - v->xmin) * v->sx);   /* Python round(): half to even */
    int64_t ny = (int64_t)nearbyint((v->ymax - v->ymin) * v->sy);
    full_antialias(c, line_width, overwrite, x0p, x1p, y0p, y1p, segment_start, segment_end, xmp, ymp, nx, ny);
  } else {
    /* map_onto_pixel_snap, line.py:689-720 */
    int64_t xxmax = (int64_t)nearbyint(axmap(v->x_log, v->xmax) * v->sx + v->tx);
    int64_t yymax = (int64_t)nearbyint(axmap(v->y_log, v->ymax) * v->sy + v->ty);
    int64_t x0i = (int64_t)(axmap(v->x_log, x0) * v->sx + v->tx), y0i = (int64_t)(axmap(v->y_log, y0) * v->sy + v->ty);
    int64_t x1i = (int64_t)(axmap(v->x_log, x1) * v->sx + v->tx), y1i = (int64_t)(axmap(v->y_log, y1) * v->sy + v->ty);
    if (x0i == xxmax) x0i--;
    if (y0i == yymax) y0i--;
    if (x1i == xxmax) x1i--;
    if (y1i == yymax) y1i--;
    bresenham(c, segment_start, x0i, x1i, y0i, y1i, clipped);
  }
}

static inline double ldxy(const void* p, int32_t dt, int64_t i) {
  return dt == ORA_F32 ? (double)((const float*)p)[i] : ((const double*)p)[i];
}

/* All line layouts through one driver: vertex (line i, vertex j) = xs[i * x_line_stride + j], ys[i * y_line_stride + j]
 * (stride 0 = a vertex vector shared by every line: LinesAxis1XConstant / YConstant, line.py:1340-1535).
 * value_per_vertex = 1 for the axis=0 layouts (LineAxis0, LineAxis0Multi, line.py:1099-1242), where append() receives
 * the row of the segment's first vertex; otherwise the line index (LinesAxis1, line.py:1244-1337). */
void ora_lines(const ora_view* v, const void* xs, const void* ys, int32_t xy_dtype, int64_t nlines, int64_t nverts,
               int64_t x_line_stride, int64_t y_line_stride, int32_t value_per_vertex, const void* val,
               int32_t val_dtype, int32_t agg_op, double line_width, void* agg) {
  line_ctx c;
  c.agg_op = agg_op; c.antialias = line_width > 0.0; c.has_field = val_dtype != ORA_NONE;
  c.width = v->width; c.agg = agg; c.field = 0.0;
  /* antialias.py:30-58: overwrite unless a SUM_1AGG combination (count / sum) is present */
  int overwrite = !(agg_op == ORA_COUNT || agg_op == ORA_SUM || agg_op == ORA_AA_COUNT_IGNORE);
  for (int64_t i = 0; i < nlines; i++) {          /* extend_cpu, line.py:1277-1289 / 1127-1136 / 1200-1210 */
    for (int64_t j = 0; j + 1 < nverts; j++) {    /* perform_extend_line, line.py:1250-1275 / 1104-1125 */
      const int64_t ox = i * x_line_stride + j, oy = i * y_line_stride + j;
      if (c.has_field) c.field = ldxy(val, val_dtype, value_per_vertex ? j : i);
      double x0 = ldxy(xs, xy_dtype, ox), y0 = ldxy(ys, xy_dtype, oy);
      double x1 = ldxy(xs, xy_dtype, ox + 1), y1 = ldxy(ys, xy_dtype, oy + 1);
      int segment_start = (j == 0) || isnan(ldxy(xs, xy_dtype, ox - 1)) || isnan(ldxy(ys, xy_dtype, oy - 1));
      int segment_end = (j == nverts - 2) || isnan(ldxy(xs, xy_dtype, ox + 2)) || isnan(ldxy(ys, xy_dtype, oy + 2));
      double xm = 0.0, ym = 0.0;
      if (!segment_start) { xm = ldxy(xs, xy_dtype, ox - 1); ym = ldxy(ys, xy_dtype, oy - 1); }
      draw_segment(v, &c, line_width, overwrite, segment_start, segment_end, x0, x1, y0, y1, xm, ym, xy_dtype == ORA_F32);
    }
  }
}

void ora_lines_axis1(const ora_view* v, const void* xs, const void* ys, int32_t xy_dtype, int64_t nlines,
                     int64_t nverts, const void* val, int32_t val_dtype, int32_t agg_op, double line_width, void* agg) {
  ora_lines(v, xs, ys, xy_dtype, nlines, nverts, nverts, nverts, 0, val, val_dtype, agg_op, line_width, agg);
}

/* ---- areas (glyphs/area.py) ------------------------------------------------------------------------------- */
/* clamp_y_indices, area.py:1083-1100 */
static inline int clamp_y_indices(int64_t ystarti, int64_t ystopi, int64_t ymaxi, int64_t* cs, int64_t* ce) {
  int oob = (ystarti < 0 && ystopi <= 0) || (ystarti > ymaxi && ystopi >= ymaxi);
  int64_t a = ystarti < ymaxi ? ystarti : ymaxi;
  *cs = a > 0 ? a : 0;
  int64_t b = ystopi < ymaxi + 1 ? ystopi : ymaxi + 1;
  *ce = b > -1 ? b : -1;
  return oob;
}

/* one pixel column of the scan fill, area.py:1205-1222 (repeated at :1236-1253 and :1300-1318) */
static void fill_column(const line_ctx* c, int64_t x, int64_t y_start, int64_t y_stop, int stacked, int64_t ymaxi) {
  if (y_start == y_stop && !stacked) { append_px(c, x, y_start); return; }
  int64_t y = y_start;
  int64_t iy = (y_start < y_stop) - (y_stop < y_start);
  if (!stacked && -1 <= y_stop + iy && y_stop + iy <= ymaxi + 1) y_stop += iy;
  while (y != y_stop) { append_px(c, x, y); y += iy; }
}

static inline int64_t snap(const ora_view* v, int is_y, double val) {   /* map_onto_pixel_snap, line.py:689-720 */
  double s = is_y ? v->sy : v->sx, t = is_y ? v->ty : v->tx, mx = is_y ? v->ymax : v->xmax;
  int lg = is_y ? v->y_log : v->x_log;
  int64_t p = (int64_t)(axmap(lg, val) * s + t);
  int64_t pmax = (int64_t)nearbyint(axmap(lg, mx) * s + t);
  return p == pmax ? p - 1 : p;
}

/* draw_trapezoid_y, area.py:1102-1320 with _skip_or_clip_trapezoid_y, area.py:1323-1380.  f32: the coordinate
 * differences of float32 columns are rounded to float32 by numba (the "to zero" curve is a float64 0.0). */
static void draw_trapezoid_y(const ora_view* v, const line_ctx* c, double x0, double x1, double y0, double y1, double y2,
                             double y3, int trapezoid_start, int stacked, int is_f32, int second_is_f32) {
  int skip = isnan(x0) || isnan(x1) || isnan(y0) || isnan(y1) || isnan(y2) || isnan(y3);
  if ((y0 > v->ymax && y1 > v->ymax && y2 > v->ymax && y3 > v->ymax) ||
      (y0 < v->ymin && y1 < v->ymin && y2 < v->ymin && y3 < v->ymin)) return;
  double t0 = 0, t1 = 1;
  double dx = is_f32 ? (double)((float)x1 - (float)x0) : x1 - x0;
  double dy0 = is_f32 ? (double)((float)y3 - (float)y0) : y3 - y0;
  double dy1 = second_is_f32 ? (double)((float)y2 - (float)y1) : y2 - y1;
  if (!clipt(-dx, x0 - v->xmin, &t0, &t1)) skip = 1;
  if (!clipt(dx, v->xmax - x0, &t0, &t1)) skip = 1;
  int clipped_start = 0, clipped_end = 0;
  if (t1 < 1) { clipped_end = 1; x1 = x0 + t1 * dx; y2 = y1 + t1 * dy1; y3 = y0 + t1 * dy0; }
  if (t0 > 0) { clipped_start = 1; x0 = x0 + t0 * dx; y0 = y0 + t0 * dy0; y1 = y1 + t0 * dy1; }
  if (skip) return;
  int64_t x0i = snap(v, 0, x0), y0i = snap(v, 1, y0), y1i = snap(v, 1, y1);
  int64_t x1i = snap(v, 0, x1), y2i = snap(v, 1, y2), y3i = snap(v, 1, y3);
  int64_t xmaxi = snap(v, 0, v->xmax), ymaxi = snap(v, 1, v->ymax);
  int64_t dxi = x1i - x0i, ix = (dxi > 0) - (dxi < 0);
  int64_t dy0i = y3i - y0i, iy0 = (dy0i > 0) - (dy0i < 0);
  int64_t dy1i = y2i - y1i, iy1 = (dy1i > 0) - (dy1i < 0);
  trapezoid_start = trapezoid_start || clipped_start;
  int64_t ys, ye;
  if (trapezoid_start) {
    int y_oob = clamp_y_indices(y0i, y1i, ymaxi, &ys, &ye);
    int x_oob = x0i < 0 || x0i > xmaxi;
    if (!(y_oob || x_oob)) fill_column(c, x0i, ys, ye, stacked, ymaxi);
  }
  int clipped = clipped_start || clipped_end;
  if (dxi == 0 && !clipped) {
    int y_oob = clamp_y_indices(y3i, y2i, ymaxi, &ys, &ye);
    int x_oob = x1i < 0 || x1i > xmaxi;
    if (!(y_oob || x_oob)) fill_column(c, x1i, ys, ye, stacked, ymaxi);
    return;
  }
  dxi = llabs(dxi) * 2; dy0i = llabs(dy0i) * 2; dy1i = llabs(dy1i) * 2;
  int64_t error0 = 2 * dy0i - dxi, error1 = 2 * dy1i - dxi;
  while (x0i != x1i) {
    while (error0 >= 0 && (error0 || ix > 0)) { error0 -= 2 * dxi; y0i += iy0; }
    error0 += 2 * dy0i;
    while (error1 >= 0 && (error1 || ix > 0)) { error1 -= 2 * dxi; y1i += iy1; }
    error1 += 2 * dy1i;
    x0i += ix;
    if (x0i < 0 || x0i > xmaxi) continue;
    if (!clamp_y_indices(y0i, y1i, ymaxi, &ys, &ye)) fill_column(c, x0i, ys, ye, stacked, ymaxi);
  }
}

/* The ten non-ragged area layouts (area.py:1383-2083): ys1 == NULL fills to y = 0 (AreaToZero*, stacked = False),
 * otherwise between the curves (AreaToLine*, stacked = True).  Layout parameters as for ora_lines. */
void ora_areas(const ora_view* v, const void* xs, const void* ys0, const void* ys1, int32_t xy_dtype, int64_t nlines,
               int64_t nverts, int64_t x_line_stride, int64_t y_line_stride, int32_t value_per_vertex, const void* val,
               int32_t val_dtype, int32_t agg_op, void* agg) {
  line_ctx c;
  c.agg_op = agg_op; c.antialias = 0; c.has_field = val_dtype != ORA_NONE; c.width = v->width; c.agg = agg; c.field = 0.0;
  const int to_line = ys1 != NULL;
  for (int64_t i = 0; i < nlines; i++) {
    for (int64_t j = 0; j + 1 < nverts; j++) {
      const int64_t ox = i * x_line_stride + j, oy = i * y_line_stride + j;
      if (c.has_field) c.field = ldxy(val, val_dtype, value_per_vertex ? j : i);
      double x0 = ldxy(xs, xy_dtype, ox), x1 = ldxy(xs, xy_dtype, ox + 1);
      double y0 = ldxy(ys0, xy_dtype, oy), y3 = ldxy(ys0, xy_dtype, oy + 1);
      double y1 = to_line ? ldxy(ys1, xy_dtype, oy) : 0.0, y2 = to_line ? ldxy(ys1, xy_dtype, oy + 1) : 0.0;
      int trapezoid_start = (j == 0) || isnan(ldxy(xs, xy_dtype, ox - 1)) || isnan(ldxy(ys0, xy_dtype, oy - 1)) ||
                            (to_line && isnan(ldxy(ys1, xy_dtype, oy - 1)));
      draw_trapezoid_y(v, &c, x0, x1, y0, y1, y2, y3, trapezoid_start, to_line, xy_dtype == ORA_F32,
                       to_line && xy_dtype == ORA_F32);
    }
  }
}

/* AreaToZeroAxis1Ragged / AreaToLineAxis1Ragged (perform_extend_area_to_zero_axis1_ragged, area.py:1959-2004;
 * perform_extend_area_to_line_axis1_ragged, :2033-2081): flat vertex arrays + a start index per row; a row draws the shortest of
 * its x / y0 (/ y1) lengths; append() receives the row.  ys1 == NULL: to zero.  The to-line form tests
 * `isnull(y1_flat[y1_start_i + j] - 1)` (:2069) - the CURRENT vertex of the second curve, not the previous one. */
void ora_areas_ragged(const ora_view* v, const void* xs, const int64_t* x_starts, int64_t x_len, const void* ys0,
                      const int64_t* y0_starts, int64_t y0_len, const void* ys1, const int64_t* y1_starts, int64_t y1_len,
                      int32_t xy_dtype, int64_t nrows, const void* val, int32_t val_dtype, int32_t agg_op, void* agg) {
  line_ctx c;
  c.agg_op = agg_op; c.antialias = 0; c.has_field = val_dtype != ORA_NONE; c.width = v->width; c.agg = agg; c.field = 0.0;
  const int to_line = ys1 != NULL;
  for (int64_t i = 0; i < nrows; i++) {
    const int64_t xb = x_starts[i], xe = i < nrows - 1 ? x_starts[i + 1] : x_len;
    const int64_t yb = y0_starts[i], ye = i < nrows - 1 ? y0_starts[i + 1] : y0_len;
    int64_t n = xe - xb < ye - yb ? xe - xb : ye - yb, sb = 0;
    if (to_line) {
      sb = y1_starts[i];
      const int64_t se = i < nrows - 1 ? y1_starts[i + 1] : y1_len;
      if (se - sb < n) n = se - sb;
    }
    if (c.has_field) c.field = ldxy(val, val_dtype, i);
    for (int64_t j = 0; j + 1 < n; j++) {
      double x0 = ldxy(xs, xy_dtype, xb + j), x1 = ldxy(xs, xy_dtype, xb + j + 1);
      double y0 = ldxy(ys0, xy_dtype, yb + j), y3 = ldxy(ys0, xy_dtype, yb + j + 1);
      double y1 = to_line ? ldxy(ys1, xy_dtype, sb + j) : 0.0, y2 = to_line ? ldxy(ys1, xy_dtype, sb + j + 1) : 0.0;
      int trapezoid_start = (j == 0) || isnan(ldxy(xs, xy_dtype, xb + j - 1)) || isnan(ldxy(ys0, xy_dtype, yb + j - 1)) ||
                            (to_line && isnan(y1));
      draw_trapezoid_y(v, &c, x0, x1, y0, y1, y2, y3, trapezoid_start, to_line, xy_dtype == ORA_F32,
                       to_line && xy_dtype == ORA_F32);
    }
  }
}

/* ---- 2-stage antialiased lines (compiler.py:198-268; line.py:1291-1319 and its per-layout twins) ------------------
 * For min / first / last and count / sum with self_intersect=False every line is first rendered on its own into a
 * cleared canvas with a max() combination (stage 1: overlapping segments of ONE line do not accumulate), then that
 * canvas is folded into the running result with the reduction's stage-2 combine: nansum_in_place, nanmin_in_place,
 * nanfirst_in_place, nanlast_in_place (utils.py:615-668, 885-897).  A single line returns after stage 1.
 * combo: ORA_SUM (f64 canvas), ORA_COUNT (f32 canvas), ORA_MIN, ORA_FIRST, ORA_LAST (f64).  agg: NaN-initialised. */
void ora_lines_aa2(const ora_view* v, const void* xs, const void* ys, int32_t xy_dtype, int64_t nlines, int64_t nverts,
                   int64_t x_line_stride, int64_t y_line_stride, int32_t value_per_vertex, const void* val,
                   int32_t val_dtype, int32_t combo, double line_width, void* agg) {
  const int64_t ncell = (int64_t)v->width * v->height;
  const int is_f32_canvas = combo == ORA_COUNT;
  void* stage1 = malloc((size_t)ncell * (is_f32_canvas ? 4 : 8));
  line_ctx c;
  c.agg_op = is_f32_canvas ? ORA_AA2_COVER : ORA_AA2_VALUE;
  c.antialias = 1; c.has_field = val_dtype != ORA_NONE; c.width = v->width; c.agg = stage1; c.field = 0.0;
  for (int64_t i = 0; i < nlines; i++) {
    if (is_f32_canvas) for (int64_t k = 0; k < ncell; k++) ((float*)stage1)[k] = NAN;      /* aa_stage_2_clear */
    else for (int64_t k = 0; k < ncell; k++) ((double*)stage1)[k] = NAN;
    for (int64_t j = 0; j + 1 < nverts; j++) {
      const int64_t ox = i * x_line_stride + j, oy = i * y_line_stride + j;
      if (c.has_field) c.field = ldxy(val, val_dtype, value_per_vertex ? j : i);
      double x0 = ldxy(xs, xy_dtype, ox), y0 = ldxy(ys, xy_dtype, oy);
      double x1 = ldxy(xs, xy_dtype, ox + 1), y1 = ldxy(ys, xy_dtype, oy + 1);
      int segment_start = (j == 0) || isnan(ldxy(xs, xy_dtype, ox - 1)) || isnan(ldxy(ys, xy_dtype, oy - 1));
      int segment_end = (j == nverts - 2) || isnan(ldxy(xs, xy_dtype, ox + 2)) || isnan(ldxy(ys, xy_dtype, oy + 2));
      /* xm = ym = 0 in 2-stage mode (line.py:1266-1268); unused because overwrite is True (antialias.py:47-56) */
      draw_segment(v, &c, line_width, 1, segment_start, segment_end, x0, x1, y0, y1, 0.0, 0.0, xy_dtype == ORA_F32);
    }
    for (int64_t k = 0; k < ncell; k++) {          /* aa_stage_2_accumulate; the first pass is a copy */
      if (is_f32_canvas) {
        float* r = (float*)agg + k; const float o = ((float*)stage1)[k];
        if (i == 0) { *r = o; continue; }
        if (isnan(*r)) { if (!isnan(o)) *r = o; } else if (!isnan(o)) *r += o;             /* nansum_in_place */
        continue;
      }
      double* r = (double*)agg + k; const double o = ((double*)stage1)[k];
      if (i == 0) { *r = o; continue; }
      switch (combo) {
        case ORA_SUM: if (isnan(*r)) { if (!isnan(o)) *r = o; } else if (!isnan(o)) *r += o; break;
        case ORA_MIN: if (isnan(*r)) { if (!isnan(o)) *r = o; } else if (!isnan(o) && o < *r) *r = o; break;
        case ORA_FIRST: if (isnan(*r) && !isnan(o)) *r = o; break;
        case ORA_LAST: if (!isnan(o)) *r = o; break;
      }
    }
  }
  free(stage1);
}
