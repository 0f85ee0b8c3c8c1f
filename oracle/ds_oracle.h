/* TEST INFRASTRUCTURE ONLY - the CPU oracle for the datashader projection + aggregation hot path.
 *
 * A plain-C restatement of the reference's numba CPU algorithm (holoviz/datashader 0.19.1).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this; the product (datashader_b200/) never does.
 *
 * Parity status: PINNED.  tests/golden/make_golden.py imports the real reference from
 * /root/reference (with container-only import shims for xarray/toolz/multipledispatch, which hold
 * no arithmetic) and writes tests/golden/\*.npz; tests/test_oracle_golden.py checks this oracle
 * against those vectors and against the literal known-answer tables of the reference's own tests
 * (datashader/tests/test_pandas.py).
 */
#ifndef DS_ORACLE_H
#define DS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Canvas view: what data_libraries/pandas.py:35-48 computes before calling extend(). */
typedef struct {
  int32_t width, height;           /* plot_width, plot_height (aggs are [height, width]) */
  int32_t x_log, y_log;            /* 0 LinearAxis, 1 LogAxis (core.py:114-132) */
  double sx, tx, sy, ty;           /* Axis.compute_scale_and_translate, core.py:62-81 */
  double xmin, xmax, ymin, ymax;   /* x_range + y_range, inclusive (points.py:199) */
} ora_view;

enum {
  ORA_COUNT = 1,     /* reductions.py:552-558, 580-584  u32 */
  ORA_ANY = 2,       /* reductions.py:843-848, 858-862  u8 (bool) */
  ORA_SUM_ZERO = 3,  /* reductions.py:956-963           f64, zero-initialised */
  ORA_SUM = 4,       /* reductions.py:1054-1063         f64, NaN-initialised */
  ORA_MIN = 5,       /* reductions.py:1178-1183         f64 */
  ORA_MAX = 6,       /* reductions.py:1222-1227         f64 */
  ORA_FIRST = 7,     /* reductions.py:1398-1404         f64 */
  ORA_LAST = 8,      /* reductions.py:1436-1442         f64 */
  ORA_MIN_ROW = 9,   /* reductions.py:2318-2324         i64, -1 = empty */
  ORA_MAX_ROW = 10,  /* reductions.py:2263-2269         i64, -1 = empty */
  ORA_WHERE = 11     /* reductions.py:1921-1928         f64 (lookup column) or i64 (row index) */
};

enum { ORA_NONE = 0, ORA_F32 = 1, ORA_F64 = 2 };

/* One call of the generated append() (compiler.py:321-475).  Ops run in order for every in-bounds
 * row; each yields the reference's return code (>=0 updated, -1 not), which a later ORA_WHERE op
 * reads through `selector`. */
typedef struct {
  int32_t op;
  int32_t val_dtype;        /* ORA_NONE / ORA_F32 / ORA_F64: the `field` argument */
  const void* val;
  int32_t nan_check_dtype;  /* optional nan_check_column (compiler.py:430-446, 461-466) */
  const void* nan_check;
  void* agg;                /* [H, W] or [H, W, ncat] when the plan is categorical */
  int32_t selector;         /* ORA_WHERE: index of the selector op in the plan (must precede) */
  int32_t lookup_is_row;    /* ORA_WHERE with lookup_column=None: store the int64 row index */
} ora_op;

typedef struct {
  int32_t nops;
  ora_op ops[8];
  const int32_t* cat;       /* optional category index per row (by / count_cat), may be negative */
  int32_t ncat;             /* 0 when not categorical */
} ora_plan;

/* Point._build_extend.extend_cpu, glyphs/points.py:188-212, rows [0, n) with global row ids
 * row_offset + i (reductions.py:87-113). xy_dtype is ORA_F32 or ORA_F64. */
void ora_points(const ora_view* v, const void* x, const void* y, int32_t xy_dtype, int64_t n,
                int64_t row_offset, const ora_plan* plan);

/* Canvas initial values, reductions.py:450-472 and 952-953. */
void ora_init(int32_t op, int32_t lookup_is_row, void* agg, int64_t ncell);

/* Combine of two partial canvases `a <- combine(a, b)` (compiler.py:478-507 + each reduction's
 * _combine): count/sum_zero add, sum nansum_missing, min/max nanmin/nanmax, any or,
 * max_row maximum, min_row row_min_in_place.  ORA_WHERE pairs use ora_combine_where. */
void ora_combine(int32_t op, void* a, const void* b, int64_t ncell);

/* where._combine_callback.combine_cpu_2d/3d (reductions.py:2009-2027): selector op is one of
 * ORA_MIN/ORA_MAX (f64 aggs, NaN empty) or ORA_MIN_ROW/ORA_MAX_ROW (i64 aggs, -1 empty). */
void ora_combine_where(int32_t selector_op, void* sel_a, const void* sel_b, void* where_a,
                       const void* where_b, int32_t where_is_i64, int64_t ncell);

/* Threaded baseline: the row range is cut into `nthreads` contiguous partitions, each aggregated
 * into a private canvas set and combined pairwise, i.e. what data_libraries/dask.py:168-217 does
 * with the threaded scheduler.  Supports plans made of COUNT / SUM_ZERO / ANY / MIN / MAX ops
 * (enough for the count, mean, by-count and max baselines). Returns 0, or -1 if unsupported. */
int ora_points_mt(const ora_view* v, const void* x, const void* y, int32_t xy_dtype, int64_t n,
                  const ora_plan* plan, int32_t nthreads);

/* ---- lines (ds_oracle_lines.c) ------------------------------------------------------------ */
/* LinesAxis1 (line.py:1244-1337) over xs, ys [nlines, nverts] (row-major, xy_dtype), one value per
 * line.  line_width == 0: Bresenham (line.py:986-1031); > 0: antialiased (line.py:826-983), which
 * supports the single-stage combinations only (any / max / count / sum with self_intersect, see
 * antialias.py:30-58).  agg_op: ORA_ANY, ORA_COUNT, ORA_SUM, ORA_MAX, ORA_MIN (non-AA only).
 * Non-AA aggs: any u8, count u32, others f64.  AA aggs: any/count f32, others f64. */
void ora_lines_axis1(const ora_view* v, const void* xs, const void* ys, int32_t xy_dtype,
                     int64_t nlines, int64_t nverts, const void* val, int32_t val_dtype,
                     int32_t agg_op, double line_width, void* agg);

/* Every line layout (LineAxis0, LineAxis0Multi, LinesAxis1, LinesAxis1XConstant/YConstant): see ds_oracle_lines.c. */
void ora_lines(const ora_view* v, const void* xs, const void* ys, int32_t xy_dtype, int64_t nlines, int64_t nverts,
               int64_t x_line_stride, int64_t y_line_stride, int32_t value_per_vertex, const void* val,
               int32_t val_dtype, int32_t agg_op, double line_width, void* agg);

/* Area glyphs, the ten non-ragged layouts (glyphs/area.py:1076-2083). ys1 == NULL: fill to y = 0. */
void ora_areas(const ora_view* v, const void* xs, const void* ys0, const void* ys1, int32_t xy_dtype, int64_t nlines,
               int64_t nverts, int64_t x_line_stride, int64_t y_line_stride, int32_t value_per_vertex, const void* val,
               int32_t val_dtype, int32_t agg_op, void* agg);
/* ragged area layouts (area.py:1939-2083): flat vertex arrays + int64 start index per row; ys1 == NULL: to zero */
void ora_areas_ragged(const ora_view* v, const void* xs, const int64_t* x_starts, int64_t x_len, const void* ys0,
                      const int64_t* y0_starts, int64_t y0_len, const void* ys1, const int64_t* y1_starts, int64_t y1_len,
                      int32_t xy_dtype, int64_t nrows, const void* val, int32_t val_dtype, int32_t agg_op, void* agg);

/* Antialiased lines whose reduction needs the 2-stage combine (compiler.py:198-268): combo = ORA_SUM / ORA_COUNT
 * (self_intersect=False; f64 / f32 canvas), ORA_MIN, ORA_FIRST, ORA_LAST (f64).  agg must be NaN-initialised. */
void ora_lines_aa2(const ora_view* v, const void* xs, const void* ys, int32_t xy_dtype, int64_t nlines, int64_t nverts,
                   int64_t x_line_stride, int64_t y_line_stride, int32_t value_per_vertex, const void* val,
                   int32_t val_dtype, int32_t combo, double line_width, void* agg);

#ifdef __cplusplus
}
#endif
#endif
