/* TEST INFRASTRUCTURE ONLY - see ds_oracle.h.  Points + reductions restated from the reference's
 * numba CPU path.  Build with -ffp-contract=off: numba/LLVM does not fuse x*sx+tx.
 */
#include "ds_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

static inline double ld(const void* p, int32_t dt, int64_t i) {
  return dt == ORA_F32 ? (double)((const float*)p)[i] : ((const double*)p)[i];
}

/* glyphs/points.py:193-203: bounds test in f64, unfused multiply-add, truncating cast, upper-edge fold */
/* LogAxis.mapper = log10(float(val)) (core.py:129-132).  Under numba float(float32) stays float32, so
   an f32 coordinate goes through the single-precision log10f and is widened afterwards (checked
   against the reference: tests/golden/points.npz pts_log_*). */
static inline double log_mapper(double v, int is_f32) {
  return is_f32 ? (double)log10f((float)v) : log10(v);
}

static inline int map_point(const ora_view* v, double x, double y, int is_f32, int64_t* xi, int64_t* yi) {
  if (!((v->xmin <= x && x <= v->xmax) && (v->ymin <= y && y <= v->ymax))) return 0;
  double xm = v->x_log ? log_mapper(x, is_f32) : x;   /* core.py:116-119 */
  double ym = v->y_log ? log_mapper(y, is_f32) : y;
  int64_t xx = (int64_t)(xm * v->sx + v->tx);   /* unfused: built with -ffp-contract=off */
  int64_t yy = (int64_t)(ym * v->sy + v->ty);
  *xi = xx >= v->width ? v->width - 1 : xx;
  *yi = yy >= v->height ? v->height - 1 : yy;
  return 1;
}

/* one reduction append; returns the reference's code (0 updated / -1 not) */
static inline int append_op(const ora_op* o, int64_t cell, int64_t i, int64_t row, const int* rets) {
  double f = 0.0;
  int has = o->val_dtype != ORA_NONE;
  if (has) f = ld(o->val, o->val_dtype, i);
  switch (o->op) {
    case ORA_COUNT:  /* reductions.py:552-558 / 580-584 */
      if (has && isnan(f)) return -1;
      ((uint32_t*)o->agg)[cell] += 1u;
      return 0;
    case ORA_ANY:    /* reductions.py:843-848 / 858-862 */
      if (has && isnan(f)) return -1;
      ((uint8_t*)o->agg)[cell] = 1;
      return 0;
    case ORA_SUM_ZERO:  /* reductions.py:956-963 */
      if (isnan(f)) return -1;
      ((double*)o->agg)[cell] += f;
      return 0;
    case ORA_SUM: {     /* reductions.py:1054-1063 */
      if (isnan(f)) return -1;
      double* a = (double*)o->agg + cell;
      if (isnan(*a)) *a = f; else *a += f;
      return 0;
    }
    case ORA_MIN: {     /* reductions.py:1178-1183 */
      double* a = (double*)o->agg + cell;
      if (!isnan(f) && (isnan(*a) || *a > f)) { *a = f; return 0; }
      return -1;
    }
    case ORA_MAX: {     /* reductions.py:1222-1227 */
      double* a = (double*)o->agg + cell;
      if (!isnan(f) && (isnan(*a) || *a < f)) { *a = f; return 0; }
      return -1;
    }
    case ORA_FIRST: {   /* reductions.py:1398-1404 */
      double* a = (double*)o->agg + cell;
      if (!isnan(f) && isnan(*a)) { *a = f; return 0; }
      return -1;
    }
    case ORA_LAST: {    /* reductions.py:1436-1442 */
      if (isnan(f)) return -1;
      ((double*)o->agg)[cell] = f;
      return 0;
    }
    case ORA_MIN_ROW: { /* reductions.py:2318-2324, field = row index */
      int64_t* a = (int64_t*)o->agg + cell;
      if (row != -1 && (*a == -1 || row < *a)) { *a = row; return 0; }
      return -1;
    }
    case ORA_MAX_ROW: { /* reductions.py:2263-2269 */
      int64_t* a = (int64_t*)o->agg + cell;
      if (row > *a) { *a = row; return 0; }
      return -1;
    }
    case ORA_WHERE: {   /* reductions.py:1921-1928 called only when selector updated (compiler.py:448-449) */
      if (rets[o->selector] < 0) return -1;
      if (o->lookup_is_row) ((int64_t*)o->agg)[cell] = row;
      else ((double*)o->agg)[cell] = f;
      return rets[o->selector];
    }
  }
  return -1;
}

void ora_points(const ora_view* v, const void* x, const void* y, int32_t xy_dtype, int64_t n,
                int64_t row_offset, const ora_plan* plan) {
  int rets[8];
  for (int64_t i = 0; i < n; i++) {   /* extend_cpu, points.py:206-212 */
    int64_t xi, yi;
    if (!map_point(v, ld(x, xy_dtype, i), ld(y, xy_dtype, i), xy_dtype == ORA_F32, &xi, &yi)) continue;
    int64_t cell = yi * v->width + xi;
    if (plan->ncat > 0) {             /* compiler.py:379-390: agg = agg[:, :, int(cat[i])] */
      int64_t c = plan->cat[i];
      if (c < 0) c += plan->ncat;     /* numba wraparound indexing for code -1 (missing category) */
      if (c < 0 || c >= plan->ncat) continue;
      cell = cell * plan->ncat + c;
    }
    for (int k = 0; k < plan->nops; k++) {
      const ora_op* o = &plan->ops[k];
      rets[k] = -1;
      /* nan_check_column guards the selector that precedes a where as well (compiler.py:439-446):
         the guarded where names it, so look ahead one op. */
      const ora_op* g = o;
      if (o->op != ORA_WHERE && k + 1 < plan->nops && plan->ops[k + 1].op == ORA_WHERE &&
          plan->ops[k + 1].selector == k && plan->ops[k + 1].nan_check_dtype != ORA_NONE)
        g = &plan->ops[k + 1];
      if (g->nan_check_dtype != ORA_NONE && isnan(ld(g->nan_check, g->nan_check_dtype, i))) continue;
      rets[k] = append_op(o, cell, i, row_offset + i, rets);
    }
  }
}

void ora_init(int32_t op, int32_t lookup_is_row, void* agg, int64_t ncell) {
  switch (op) {
    case ORA_COUNT: memset(agg, 0, ncell * 4); break;
    case ORA_ANY: memset(agg, 0, ncell); break;
    case ORA_SUM_ZERO: memset(agg, 0, ncell * 8); break;
    case ORA_MIN_ROW: case ORA_MAX_ROW:
      for (int64_t i = 0; i < ncell; i++) ((int64_t*)agg)[i] = -1;
      break;
    case ORA_WHERE:
      if (lookup_is_row) { for (int64_t i = 0; i < ncell; i++) ((int64_t*)agg)[i] = -1; break; }
      /* fallthrough */
    default:
      for (int64_t i = 0; i < ncell; i++) ((double*)agg)[i] = NAN;
  }
}

void ora_combine(int32_t op, void* a_, const void* b_, int64_t ncell) {
  switch (op) {
    case ORA_COUNT: {   /* reductions.py:652-654 */
      uint32_t* a = a_; const uint32_t* b = b_;
      for (int64_t i = 0; i < ncell; i++) a[i] += b[i];
      break;
    }
    case ORA_ANY: {     /* reductions.py:881-883 */
      uint8_t* a = a_; const uint8_t* b = b_;
      for (int64_t i = 0; i < ncell; i++) a[i] = a[i] | b[i];
      break;
    }
    case ORA_SUM_ZERO: {  /* reductions.py:994-996 */
      double* a = a_; const double* b = b_;
      for (int64_t i = 0; i < ncell; i++) a[i] += b[i];
      break;
    }
    case ORA_SUM: {     /* nansum_missing, utils.py:161-181 */
      double* a = a_; const double* b = b_;
      for (int64_t i = 0; i < ncell; i++) {
        if (isnan(a[i])) a[i] = b[i];
        else if (!isnan(b[i])) a[i] += b[i];
      }
      break;
    }
    case ORA_MIN: {     /* np.nanmin, reductions.py:1203-1205 */
      double* a = a_; const double* b = b_;
      for (int64_t i = 0; i < ncell; i++) if (isnan(a[i]) || b[i] < a[i]) a[i] = b[i];
      break;
    }
    case ORA_MAX: {     /* np.nanmax, reductions.py:1258-1260 */
      double* a = a_; const double* b = b_;
      for (int64_t i = 0; i < ncell; i++) if (isnan(a[i]) || b[i] > a[i]) a[i] = b[i];
      break;
    }
    case ORA_MAX_ROW: { /* np.maximum, reductions.py:2289-2297 */
      int64_t* a = a_; const int64_t* b = b_;
      for (int64_t i = 0; i < ncell; i++) if (b[i] > a[i]) a[i] = b[i];
      break;
    }
    case ORA_MIN_ROW: { /* row_min_in_place, utils.py:913-923 */
      int64_t* a = a_; const int64_t* b = b_;
      for (int64_t i = 0; i < ncell; i++) if (b[i] != -1 && (a[i] == -1 || b[i] < a[i])) a[i] = b[i];
      break;
    }
  }
}

void ora_combine_where(int32_t selector_op, void* sel_a, const void* sel_b, void* where_a,
                       const void* where_b, int32_t where_is_i64, int64_t ncell) {
  /* combine_cpu_2d, reductions.py:2009-2016: value = selector_b; if valid and
     selector._append(selector_a, value) >= 0: where_a = where_b */
  for (int64_t i = 0; i < ncell; i++) {
    int upd = 0;
    if (selector_op == ORA_MIN_ROW || selector_op == ORA_MAX_ROW) {
      int64_t* a = (int64_t*)sel_a + i; int64_t b = ((const int64_t*)sel_b)[i];
      if (b == -1) continue;
      if (selector_op == ORA_MIN_ROW) { if (*a == -1 || b < *a) { *a = b; upd = 1; } }
      else if (b > *a) { *a = b; upd = 1; }
    } else {
      double* a = (double*)sel_a + i; double b = ((const double*)sel_b)[i];
      if (isnan(b)) continue;
      if (selector_op == ORA_MIN) { if (isnan(*a) || *a > b) { *a = b; upd = 1; } }
      else if (isnan(*a) || *a < b) { *a = b; upd = 1; }
    }
    if (upd) {
      if (where_is_i64) ((int64_t*)where_a)[i] = ((const int64_t*)where_b)[i];
      else ((double*)where_a)[i] = ((const double*)where_b)[i];
    }
  }
}

static int64_t op_cell_bytes(int32_t op) {
  switch (op) {
    case ORA_COUNT: return 4;
    case ORA_ANY: return 1;
    default: return 8;
  }
}

typedef struct {
  const ora_view* v; const void* x; const void* y; int32_t xy_dtype; int64_t n;
  const ora_plan* plan; int32_t t, nthreads; int64_t ncell; void** priv;
} mt_arg;

/* one partition: create + extend (dask.py:168-173) */
static void* mt_worker(void* arg_) {
  mt_arg* a = (mt_arg*)arg_;
  const ora_plan* plan = a->plan;
  int64_t lo = a->n * a->t / a->nthreads, hi = a->n * (a->t + 1) / a->nthreads;
  int64_t xyb = a->xy_dtype == ORA_F32 ? 4 : 8;
  ora_plan p = *plan;
  for (int k = 0; k < plan->nops; k++) {
    ora_op* o = &p.ops[k];
    if (a->t > 0) {
      o->agg = malloc((size_t)(a->ncell * op_cell_bytes(o->op)));
      ora_init(o->op, 0, o->agg, a->ncell);
    }
    a->priv[(size_t)a->t * plan->nops + k] = o->agg;
    if (o->val_dtype != ORA_NONE) o->val = (const char*)o->val + lo * (o->val_dtype == ORA_F32 ? 4 : 8);
  }
  if (p.ncat > 0) p.cat = plan->cat + lo;
  ora_points(a->v, (const char*)a->x + lo * xyb, (const char*)a->y + lo * xyb, a->xy_dtype, hi - lo, lo, &p);
  return NULL;
}

int ora_points_mt(const ora_view* v, const void* x, const void* y, int32_t xy_dtype, int64_t n,
                  const ora_plan* plan, int32_t nthreads) {
  for (int k = 0; k < plan->nops; k++) {
    int op = plan->ops[k].op;
    if (!(op == ORA_COUNT || op == ORA_SUM_ZERO || op == ORA_ANY || op == ORA_MIN || op == ORA_MAX)) return -1;
  }
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  int64_t ncell = (int64_t)v->width * v->height * (plan->ncat > 0 ? plan->ncat : 1);
  void** priv = calloc((size_t)nthreads * plan->nops, sizeof(void*));
  mt_arg* args = calloc((size_t)nthreads, sizeof(mt_arg));
  pthread_t* th = calloc((size_t)nthreads, sizeof(pthread_t));
  for (int t = 0; t < nthreads; t++) {
    mt_arg a = {v, x, y, xy_dtype, n, plan, t, nthreads, ncell, priv};
    args[t] = a;
    if (t > 0) pthread_create(&th[t], NULL, mt_worker, &args[t]);
  }
  mt_worker(&args[0]);
  for (int t = 1; t < nthreads; t++) pthread_join(th[t], NULL);
  /* combine (compiler.py:492-505), partition 0 owns the caller's canvases */
  for (int t = 1; t < nthreads; t++)
    for (int k = 0; k < plan->nops; k++) {
      void* b = priv[(size_t)t * plan->nops + k];
      ora_combine(plan->ops[k].op, plan->ops[k].agg, b, ncell);
      free(b);
    }
  free(priv); free(args); free(th);
  return 0;
}
