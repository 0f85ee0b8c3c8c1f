/* dsb200 - B200-native (sm_100a) kernels for datashader's projection + aggregation hot path.
 *
 * C ABI of libdsb200.so.  Plain pointers and sizes only; every buffer is a DEVICE pointer owned by
 * the caller (the library allocates nothing and never copies to the host).  All entry points are
 * asynchronous on `stream` (a cudaStream_t passed as void*), return 0 on success or a negative
 * dsb_status, and leave a message for dsb_last_error().
 *
 * Each entry point names the reference interface (holoviz/datashader 0.19.1, paths relative to
 * datashader/) it replaces; INTEGRATION.md shows the ctypes binding a maintainer would add.
 */
#ifndef DSB200_H
#define DSB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSB_ABI_VERSION 3
#define DSB_MAX_OPS 8

typedef enum {
  DSB_OK = 0,
  DSB_ERR_ARG = -1,       /* bad argument (dtype, op, null pointer, size) */
  DSB_ERR_CUDA = -2,      /* CUDA runtime error, text in dsb_last_error() */
  DSB_ERR_UNSUPPORTED = -3
} dsb_status;

/* column dtypes (numpy kinds the reference accepts for coordinates / value columns) */
typedef enum {
  DSB_NONE = 0, DSB_F32 = 1, DSB_F64 = 2, DSB_I8 = 3, DSB_U8 = 4, DSB_I16 = 5, DSB_U16 = 6,
  DSB_I32 = 7, DSB_U32 = 8, DSB_I64 = 9, DSB_U64 = 10
} dsb_dtype;

/* What data_libraries/pandas.py:35-48 computes before calling extend(): canvas shape, the
 * Axis.compute_scale_and_translate pairs (core.py:62-81) and the inclusive data bounds. */
typedef struct {
  int32_t width, height;          /* canvases are row-major [height, width] (pandas.py:48) */
  int32_t x_log, y_log;           /* 0 LinearAxis, 1 LogAxis (core.py:114-132) */
  double sx, tx, sy, ty;
  double xmin, xmax, ymin, ymax;
} dsb_view;

/* Accumulator ops: the device-side state of the reference's base reductions (compiler.py:103-107).
 * Every accumulator is commutative, so partial canvases combine with a plain elementwise
 * sum / max / min (NCCL) in any order.  Canvas element types are given per op; `row` is the global
 * row id row_offset + i (reductions.py:87-113). */
typedef enum {
  DSB_OP_COUNT = 1,    /* i32/u32 canvas (init 0): += 1                      count._append[_no_field] reductions.py:552-584 */
  DSB_OP_ANY = 2,      /* u8 canvas (init 0): = 1                            any._append              reductions.py:843-862 */
  DSB_OP_SUM = 3,      /* f64 canvas (init 0): += val                        _sum_zero._append        reductions.py:956-963 */
  DSB_OP_MAX32 = 4,    /* i32 key canvas (init INT32_MIN): max of key32(val) max._append              reductions.py:1222-1227 */
  DSB_OP_MIN32 = 5,    /* i32 key canvas (init INT32_MAX): min of key32(val) min._append              reductions.py:1178-1183 */
  DSB_OP_MAX64 = 6,    /* i64 key canvas (init INT64_MIN): max of key64((double)val) */
  DSB_OP_MIN64 = 7,    /* i64 key canvas (init INT64_MAX) */
  DSB_OP_MAXROW = 8,   /* i64 canvas (init -1): max of row                   _max_row_index._append   reductions.py:2263-2269 */
  DSB_OP_MINROW = 9,   /* i64 canvas (init INT64_MAX): min of row            _min_row_index._append   reductions.py:2318-2324 */
  DSB_OP_ARGMAX32 = 10,/* i64 canvas (init INT64_MIN): max of key32(val)<<32 | ~u32(row)  -> where(max(val)) with the
                          reference's earliest-row tie rule (strict compare, reductions.py:1224, 2014-2016) */
  DSB_OP_ARGMIN32 = 11,/* i64 canvas (init INT64_MAX): min of key32(val)<<32 | u32(row) */
  DSB_OP_MATCHROW64 = 12 /* second pass for 64-bit selectors: `aux` is a finished MAX64/MIN64 key canvas; rows whose
                          key equals it contribute min(row) into an i64 canvas (init INT64_MAX) */
} dsb_op;

/* key32: order-preserving signed 32-bit key of an f32 / (u)int8-32 value; key64: the same for the
 * value widened to f64 (what the reference stores: agg[y, x] = field).  NaN values never reach a key. */

typedef struct {
  int32_t op;               /* dsb_op */
  int32_t val_dtype;        /* dsb_dtype of `val`, DSB_NONE if the op has no field */
  const void* val;          /* [n] value column (the reference's `field`); NaN rows are skipped */
  int32_t chk_dtype;        /* optional nan_check_column (compiler.py:439-446, 461-466) */
  const void* chk;
  void* agg;                /* accumulator canvas [H*W] or [H*W*ncat] */
  const void* aux;          /* DSB_OP_MATCHROW64 only */
} dsb_base;

typedef struct {
  int32_t nops;
  dsb_base ops[DSB_MAX_OPS];
  const void* cat;          /* optional [n] integer category codes: by()/count_cat (compiler.py:379-390) */
  int32_t cat_dtype;        /* DSB_I8 / DSB_I16 / DSB_I32 / DSB_I64, DSB_NONE when not categorical */
  int32_t ncat;             /* canvases become [H, W, ncat]; negative codes wrap like numba's agg[:, :, -1] */
  uint32_t* notes;          /* optional device word (may be NULL), OR-ed with DSB_NOTE_* bits by the pass */
} dsb_plan;

/* dsb_plan.notes bits.  DSB_NOTE_NEGZERO: a MAX32/MIN32/MAX64/MIN64 accumulator consumed a float -0.0.  The keys fold
 * -0.0 onto +0.0 (the reference compares with < / >, for which the zeros tie), so a pixel whose extreme is a zero
 * decodes as +0.0 while the reference keeps whichever zero ARRIVED FIRST (strict compare, reductions.py:1178-1183,
 * 1222-1227).  When the bit is set the caller re-runs those reductions through the row-exact accumulators
 * (ARGMAX32 / ARGMIN32, or MAX64 / MIN64 + MATCHROW64) and gathers the winning row's own bit pattern. */
#define DSB_NOTE_NEGZERO 1u

int dsb_abi_version(void);
const char* dsb_last_error(void);
/* number of kernel-launching API calls made by this process so far (diagnostics / bench.py gpu_launches) */
int64_t dsb_launch_count(void);
/* name (template arguments included) of the aggregation kernel the most recent dsb_points* / dsb_lines* / dsb_areas* call
 * of this thread launched - diagnostics: bench.py labels its roofline with it */
const char* dsb_last_kernel(void);

/* Runtime knobs: "l2_band_bytes" - accumulator bytes one dsb_points launch may touch before the rows are re-read
 * once per band of canvas rows (default 96 MiB, 0 disables); "band_min_rows" - smallest n that is banded;
 * "priv_smem_kb" / "priv_smem_kb_mean" - shared memory dsb_points_priv may take for the privatised canvas (defaults
 * 192 / 226 KB, see points.cu); "priv_tight" - 0 disables the specialised count()/mean() kernels (A/B testing);
 * "mono" / "mono_banded" - 0 disables k_points_mono (single monotone accumulator) / its use in banded passes;
 * "split_bytes" - plans whose canvases total more than this run one pass per accumulator (default 48 MiB, 0 = never);
 * "count16_band_bytes" - optional banding of dsb_points_count16's packed canvas (default 0 = off). */
int dsb_configure(const char* key, int64_t value);

/* Initialise an accumulator canvas of `ncell` elements to the op's identity (see dsb_op).
 * Replaces Reduction._build_create / make_create (reductions.py:397-472, compiler.py:294-301). */
int dsb_init_canvas(int32_t op, void* agg, int64_t ncell, void* stream);

/* Fused projection + reduction over n points: replaces Point._build_extend.extend_cuda
 * (glyphs/points.py:188-221) and the generated append() (compiler.py:321-475).
 * x, y: [n] device columns of xy_dtype (DSB_F32 or DSB_F64); n <= 2^32 per call. */
int dsb_points(const dsb_view* view, const void* x, const void* y, int32_t xy_dtype, int64_t n,
               int64_t row_offset, const dsb_plan* plan, void* stream);

/* K2 - the same contract as dsb_points for a plan that contains a COUNT or ANY accumulator (`priv_op` = its index), with
 * that accumulator's whole canvas privatised per SM in shared memory as packed guard-bit counters.
 * Requires float32 coordinates (any plan) or float64 coordinates (count / any, or SUM + COUNT of one float64 column;
 * 16-byte aligned columns, no categories) and width*height*ncat <= ~950 000 cells; otherwise returns
 * DSB_ERR_UNSUPPORTED and the caller uses dsb_points.  scratch: [cells] u32 work canvas, flag: 1 u32 (both device, contents ignored on entry).
 * The result is always exact: a detected counter carry makes the library redo the count with global REDs.
 * An ANY accumulator is counted in `scratch` and committed as canvas |= (scratch > 0); it takes n < 2^32 rows per call. */
int dsb_points_priv(const dsb_view* view, const void* x, const void* y, int32_t xy_dtype, int64_t n,
                    int64_t row_offset, const dsb_plan* plan, int32_t priv_op, uint32_t* scratch, uint32_t* flag,
                    void* stream);

/* Routed aggregation for a single-accumulator plan on a canvas that fits neither shared memory nor L2 (BASELINE config 5,
 * 8192 x 8192): the points are binned into buckets of <= 45 056 canvas cells as 8-byte records (pass 1) and every bucket
 * is then accumulated with its cells in shared memory and folded into the canvas with plain loads and stores (pass 2) -
 * instead of dsb_points' L2-banded re-reads with a global RED per hit.  Same contract and result as dsb_points for
 * MAX32 / MIN32 (float32 value column), MINROW / MAXROW (float32 nan-check column) and COUNT (optional float32 column)
 * with float32 coordinates on linear axes, 16-byte aligned columns, no categories, n in [2^24, 2^32 - 1); anything else
 * returns DSB_ERR_UNSUPPORTED and the caller uses dsb_points.  Exact for every distribution: records that do not fit the
 * sampled capacity of their bucket are applied to the canvas directly with atomics.
 * scratch: device memory of at least dsb_points_routed_scratch_bytes(view, n) bytes (contents ignored on entry). */
int64_t dsb_points_routed_scratch_bytes(const dsb_view* view, int64_t n);
int dsb_points_routed(const dsb_view* view, const void* x, const void* y, int32_t xy_dtype, int64_t n, int64_t row_offset,
                      const dsb_plan* plan, void* scratch, int64_t scratch_bytes, void* stream);
/* A/B switch of the antialiased single-stage line kernel: 1 (default) = rows of the scan conversion balanced over the warp */
int dsb_lines_configure(int balanced);
/* smallest n dsb_points_routed accepts (default 2^24; tests lower it) */
int dsb_routed_configure(int64_t min_rows);

/* Batched viewports: the same contract as dsb_points for `nviews` views of identical canvas size at once, in ONE pass
 * over the columns (a zoom level of a tile pyramid, tiles.py:70-131; the viewports of an interaction, pipeline.py:55-72).
 * views: DEVICE array of dsb_view.  The plan's canvases are stacked [nviews, H, W(, ncat)], view_cells = H * W * max(ncat, 1);
 * every view keeps its own scale / translate / bounds, so each canvas equals the one dsb_points produces for that view.
 * grid_nx > 0: the views form a row-major grid_nx x grid_ny grid of equal extents (gtw x gth) starting at (gx0, gy0) and a
 * point is tested against its grid cell, and against the neighbour(s) across an edge when it lies within 1/1024 of a tile of that
 * edge (the views' own bounds decide; they must match the grid to 1e-6 of a tile); grid_nx == 0: at most 64 arbitrary views, all tested. */
int dsb_points_views(const dsb_view* views, int32_t nviews, int32_t grid_nx, int32_t grid_ny, double gx0, double gy0,
                     double gtw, double gth, const void* x, const void* y, int32_t xy_dtype, int64_t n, int64_t row_offset,
                     const dsb_plan* plan, int64_t view_cells, void* stream);

/* NaN-skipping min/max of a column: Glyph._compute_bounds_numba (glyphs/glyph.py:66-78).
 * out_minmax: 2 doubles on the device, (+inf, -inf) when no finite-or-inf value exists. */
int dsb_bounds(const void* col, int32_t dtype, int64_t n, double* out_minmax, void* stream);

/* ---- canvas-sized finishing passes (replace each reduction's _finalize, reductions.py:491-493,
 *      1091-1098, 1292-1297, 1374-1378, 2150-2161) -------------------------------------------- */
/* key canvas -> f64 with NaN for empty cells. val_dtype selects the key decoding. */
int dsb_decode_minmax(const void* keys, int32_t op, int32_t val_dtype, double* out, int64_t ncell, void* stream);
/* ARGMAX32/ARGMIN32 packed canvas -> selector value (f64, may be NULL) and global row id (i64, -1 empty).
 * row_offset is the first global row that contributed (rows span < 2^32), used to widen the 32-bit row field. */
int dsb_decode_arg(const void* packed, int32_t op, int32_t val_dtype, int64_t row_offset, double* out_sel,
                   int64_t* out_row, int64_t ncell, void* stream);
/* row canvas (i64; -1 or INT64_MAX = empty) -> f64 lookup[row - row_offset] (NaN for empty); rows outside
 * [row_offset, row_offset + n) are left untouched in `out` (used by the multi-GPU combine). */
int dsb_gather_rows(const int64_t* rows, int64_t row_offset, int64_t n, const void* lookup, int32_t lookup_dtype,
                    double* out, int64_t ncell, void* stream);
/* MINROW canvas -> reference form (-1 for empty). */
int dsb_finish_minrow(int64_t* rows, int64_t ncell, void* stream);
/* mean = where(count > 0, sum / count, nan)  (reductions.py:1292-1297) */
int dsb_finalize_mean(const double* sum, const void* count_u32, double* out, int64_t ncell, void* stream);
/* sum  = where(mask, sum, nan); mask is a u8 any-canvas (reductions.py:1091-1096) */
int dsb_finalize_sum(const double* sum, const uint8_t* mask, double* out, int64_t ncell, void* stream);
/* the same with the non-null count of the column as the mask: sum = where(count > 0, sum, nan).  Lets sum() share
 * its two accumulators with mean() and ride the privatised count kernel (dsb_points_priv). */
int dsb_finalize_sum_counted(const double* sum, const void* count_u32, double* out, int64_t ncell, void* stream);

/* where(max | min) of a float32 selector as TWO passes, for canvases beyond L2 (csrc/match.cu): pass 1 is the plain
 * DSB_OP_MAX32 / MIN32 accumulator of the selector column (dsb_points_routed serves it without global atomics); this is pass 2.
 * `keys` = the finished (all ranks combined) key32 canvas [H, W]; every row whose key equals its pixel's entry votes its
 * global row row_offset + i into `rows` (i64 [H, W], dsb_init_canvas(DSB_OP_MINROW)) with an atomic min: the earliest row among
 * the ties, the one the reference's strict compare keeps (reductions.py:1178-1183, 1222-1227, 2009-2016).  A coarse map in
 * `scratch` (dsb_points_match32_scratch_bytes: the least extreme of every 16 x 16 pixel block) drops the rows that cannot
 * match after one L2 hit.  float32 x / y / val; any axes (linear axes inside the float32 mapping's error bound take the
 * filtered kernel, everything else the exact mapping per row). */
int64_t dsb_points_match32_scratch_bytes(const dsb_view* view);
int dsb_points_match32(const dsb_view* view, const void* x, const void* y, int32_t xy_dtype, int64_t n, int64_t row_offset,
                       const void* val, int32_t val_dtype, const void* keys, int32_t is_max, void* rows, void* scratch,
                       int64_t scratch_bytes, void* stream);

/* max / min of a float32 column on a canvas that fits L2, for frames with hundreds of rows per pixel: the caller runs dsb_points
 * (DSB_OP_MAX32 / MIN32) over the HEAD of the rows and this entry over the rest.  `keys` is that live accumulator; the least extreme
 * of every small block of pixels after the head (16 bits each, shared memory) drops the rows that cannot win any more - ~98 % - and
 * the others are compared with their pixel's key and replace it if they beat it (max._append / min._append, reductions.py:1222-1227,
 * 1178-1183: the same strict compare).  `notes`: DSB_NOTE_NEGZERO is set when a zero takes or ties an extreme (its sign is then
 * resolved by the caller as for dsb_points).  scratch: 4 bytes per block (<= 256 KB).  float32 x / y / val, linear axes inside the
 * float32 mapping's error bound, 16-byte aligned columns; DSB_ERR_UNSUPPORTED otherwise (run dsb_points over these rows as well). */
int dsb_points_minmax_rest(const dsb_view* view, const void* x, const void* y, int32_t xy_dtype, int64_t n, int64_t row_offset,
                           const void* val, int32_t val_dtype, void* keys, int32_t is_max, unsigned int* notes, void* scratch,
                           int64_t scratch_bytes, void* stream);
/* The same for where(max | min) of a float32 selector: `packed` is the live DSB_OP_ARGMAX32 / ARGMIN32 accumulator (i64 {key32, row}
 * per pixel, the head of the rows already in it); a row that reaches its block's threshold swaps itself in if its packed value beats
 * the pixel's - value first, the earliest row on ties (reductions.py:2009-2016). */
int dsb_points_argminmax_rest(const dsb_view* view, const void* x, const void* y, int32_t xy_dtype, int64_t n, int64_t row_offset,
                              const void* val, int32_t val_dtype, void* packed, int32_t is_max, void* scratch, int64_t scratch_bytes,
                              void* stream);

/* ---- lines ---------------------------------------------------------------------------------- */
typedef enum { DSB_LINE_ANY = 1, DSB_LINE_COUNT = 2, DSB_LINE_SUM = 3, DSB_LINE_MAX = 4, DSB_LINE_MIN = 5,
               DSB_LINE_MEAN = 6 /* antialiased only: canvas f64 sum (zeroed), `mask` = u32 count canvas (zeroed) */,
               DSB_LINE_MEAN_2STAGE = 7 /* DSB_LINE_MEAN drawn in overwrite mode: mean when a 2-stage reduction shares its summary (its
                                           bases are FloatingReductions: every touch adds value x coverage and counts 1,
                                           reductions.py:965-973, 687-693; compiler.py:539-554) */ } dsb_line_agg;

/* Vertex addressing of the other line layouts.  NULL = LinesAxis1: dense [nlines, nverts] matrices, one value per
 * line.  x_line_stride / y_line_stride = elements between consecutive lines (0 = one vertex vector shared by all
 * lines: LinesAxis1XConstant / YConstant, line.py:1340-1535).  value_per_vertex = 1 for the axis=0 layouts
 * (LineAxis0, LineAxis0Multi, line.py:1099-1242): value / category / row index are those of the segment's first
 * vertex.  plot_start: whether vertex 0 starts a line (line.py:1112-1113).
 * Ragged layouts (LinesAxis1Ragged, line.py:457-523 + 1538-1600; AreaToZeroAxis1Ragged / AreaToLineAxis1Ragged,
 * area.py:1939-2083; the columns are datatypes.py RaggedArrays): x_starts != NULL.  xs / ys (/ ys1 of dsb_areas_plan) are
 * then the FLAT vertex arrays of x_flat_len / y_flat_len / y1_flat_len elements and row i owns the vertices
 * [starts[i], starts[i + 1]) of each (the last row up to the flat length); starts are int64 device arrays of nlines
 * non-decreasing entries.  A row draws min(its x, y (, y1) lengths) vertices; value / category / row index are per row.
 * The strides, value_per_vertex and nverts are ignored (pass nverts >= 2). */
typedef struct {
  int64_t x_line_stride, y_line_stride;
  int32_t value_per_vertex;
  int32_t plot_start;
  const int64_t* x_starts;
  const int64_t* y_starts;
  const int64_t* y1_starts;
  int64_t x_flat_len, y_flat_len, y1_flat_len;
} dsb_line_layout;

/* LinesAxis1 (glyphs/line.py:1244-1337): xs, ys are [nlines, nverts] row-major of xy_dtype, `val`
 * one value per line.  line_width == 0 -> Liang-Barsky clip + snapped Bresenham (line.py:734-780,
 * 986-1031); line_width > 0 -> the antialiased rasteriser (line.py:826-983) for the single-stage
 * combinations (any, max: overwrite; count, sum: previous-segment correction; antialias.py:30-58).
 * Canvases: line_width == 0: any u8, count i32, sum f64 (+ `mask` u8), max/min i64 key64.
 *           line_width  > 0: any i32 key32 of f32, count f32 (+mask), sum f64 (+mask), max i64 key64. */
int dsb_lines_axis1(const dsb_view* view, const void* xs, const void* ys, int32_t xy_dtype, int64_t nlines,
                    int64_t nverts, const dsb_line_layout* layout, const void* val, int32_t val_dtype, int32_t agg,
                    double line_width, void* canvas, uint8_t* mask, void* stream);

/* The same with a category column (antialiased by(cat, any | count | sum | max | mean), compiler.py:379-390): canvases
 * are [H, W, ncat], every line (or vertex row, axis=0 layouts) updates the plane of its category; negative codes wrap
 * like numba's agg[:, :, -1].  ncat == 0: identical to dsb_lines_axis1. */
int dsb_lines_axis1_cat(const dsb_view* view, const void* xs, const void* ys, int32_t xy_dtype, int64_t nlines,
                        int64_t nverts, const dsb_line_layout* layout, const void* val, int32_t val_dtype, int32_t agg,
                        double line_width, void* canvas, uint8_t* mask, const void* cat, int32_t cat_dtype, int32_t ncat,
                        void* stream);

/* LinesAxis1 with line_width == 0 and a full accumulator plan (every reduction dsb_points supports): the plan runs
 * for every pixel a line touches with i = the line's row, exactly how the reference hands the row index to append()
 * from _bresenham (line.py:1006-1031).  Value / nan-check / category columns are per line ([nlines]). */
int dsb_lines_axis1_plan(const dsb_view* view, const void* xs, const void* ys, int32_t xy_dtype, int64_t nlines,
                         int64_t nverts, const dsb_line_layout* layout, int64_t row_offset, const dsb_plan* plan,
                         void* stream);

/* Antialiased lines whose reduction needs the reference's 2-stage combine (compiler.py:198-268, line.py:1291-1319,
 * antialias.py:30-58): every line is rendered on its own with a max() combination (stage 1) and folded into the result
 * with nansum / nanmin / nanfirst / nanlast (stage 2).  One CTA per line; stage 1 lives in a shared-memory hash table
 * (12288 touched pixels per group of lines); longer lines are queued and redone with a private full-size stage-1
 * canvas + touched bitmap per CTA in `scratch`: 8 * pixels + 4 * ceil(pixels / 32) + 1 MiB bytes per CTA (as many CTAs as fit,
 * at most one per SM) + 4 * (nlines + 4) bytes of queue.
 *   DSB_AA2_SUM   sum(self_intersect=False):   out f64 zero-initialised, aux u8 mask (zeroed)
 *   DSB_AA2_COUNT count(self_intersect=False): out f32 zero-initialised, aux u8 mask (zeroed); val optional (NaN check)
 *   DSB_AA2_MIN   min:                         out i64 key64 canvas (dsb_init_canvas(DSB_OP_MIN64)), aux unused
 *   DSB_AA2_ARGMIN / DSB_AA2_ARGMAX (where(min | max)): phase 1 folds the per-line values into out (i64 key64 canvas,
 *                 DSB_OP_MIN64 / DSB_OP_MAX64 init); phase 2 - on the finished (all-reduced) out - records in aux (i64,
 *                 DSB_OP_MINROW init) the lowest line index whose per-line value equals it: the line the reference's
 *                 strict compare keeps (reductions.py:2009-2016)
 *   DSB_AA2_FIRST / DSB_AA2_LAST:              aux i64 line-index canvas (DSB_OP_MINROW / DSB_OP_MAXROW init); call
 *                 with phase = 1 (votes the global line index row_offset + i into aux; all-reduce aux across GPUs
 *                 here), then phase = 2 (the winning line stores its value into out, f64, NaN-initialised).  Or ONE call
 *                 with phase = 3: out = [pixels] pairs {i64 line index, f64 value bits} (16-byte aligned, initialised to
 *                 {INT64_MAX | -1, NaN}), updated with a 128-bit compare-and-swap; aux unused.  ARGMIN / ARGMAX with
 *                 phase = 3: out = pairs {i64 key64, i64 line index} initialised to {INT64_MAX | INT64_MIN, INT64_MAX}. */
typedef enum { DSB_AA2_SUM = 1, DSB_AA2_COUNT = 2, DSB_AA2_MIN = 3, DSB_AA2_FIRST = 4, DSB_AA2_LAST = 5,
               DSB_AA2_ARGMIN = 6, DSB_AA2_ARGMAX = 7 } dsb_aa2_combo;
int dsb_lines_aa2(const dsb_view* view, const void* xs, const void* ys, int32_t xy_dtype, int64_t nlines,
                  int64_t nverts, const dsb_line_layout* layout, int64_t row_offset, const void* val,
                  int32_t val_dtype, int32_t combo, int32_t phase, double line_width, void* out, void* aux,
                  void* scratch, int64_t scratch_bytes, void* stream);

/* ---- areas ----------------------------------------------------------------------------------------- */
/* Filled areas (glyphs/area.py): one trapezoid per vertex pair, x-driven double-Bresenham scan fill
 * (_build_draw_trapezoid_y :1076-1320, _skip_or_clip_trapezoid_y :1323-1380) with the accumulator plan applied to
 * every filled pixel.  ys1 == NULL: fill to y = 0 (AreaToZero*, stacked = False); otherwise fill between the two
 * curves (AreaToLine*, stacked = True).  Layouts as for lines (dsb_line_layout; y_line_stride applies to both ys). */
int dsb_areas_plan(const dsb_view* view, const void* xs, const void* ys0, const void* ys1, int32_t xy_dtype,
                   int64_t nlines, int64_t nverts, const dsb_line_layout* layout, int64_t row_offset,
                   const dsb_plan* plan, void* stream);

/* ---- shade: tf.shade / eq_hist (transfer_functions/__init__.py) --------------------------------- */
/* how codes */
#define DSB_HOW_EQ_HIST 0
#define DSB_HOW_LOG 1
#define DSB_HOW_CBRT 2
#define DSB_HOW_LINEAR 3
/* or-ed into `how` of dsb_shade_map2d for float32 canvases: numpy evaluates log1p / ** (1/3.) in float32 there */
#define DSB_HOW_F32 0x100

/* Per-pixel totals of a categorical count canvas counts[npix, ncat] (u32) and global statistics:
 * stats[0] min entry (colour baseline, __init__.py:398), [1] min total, [2] min non-zero total, [3] max total.
 * Replaces nansum_missing (utils.py:161-181) + the nanmin scans of _colorize / _interpolate_alpha. */
int dsb_shade_cat_totals(const uint32_t* counts, int64_t npix, int32_t ncat, uint64_t* total, uint64_t* stats, void* stream);

/* eq_hist step 1 (__init__.py:194-211): histogram of d = total - offset over unmasked pixels.  integer_mode = the
 * exact unique-value path (one bin per integer in [first, last]); otherwise numpy.histogram's uniform-bin rule with
 * `nbins` bins over [first, last]. */
int dsb_eqhist_hist_u64(const uint64_t* total, int64_t npix, uint64_t offset, int32_t mask_zero, double first, double last,
                        int32_t nbins, int32_t integer_mode, uint32_t* hist, void* stream);
int dsb_eqhist_hist_f64(const double* vals /* NaN = masked */, int64_t npix, double offset, double first, double last,
                        int32_t nbins, int32_t integer_mode, uint32_t* hist, void* stream);
/* eq_hist step 2 (__init__.py:205-213): drop empty bins (unless integer_mode), cumulative sum, cdf / cdf[-1].
 * xp, cdf: [nbins] outputs; meta[0] = entries written, meta[1] = discrete_levels.  One CTA, block scan. */
int dsb_eqhist_scan(const uint32_t* hist, int32_t nbins, int32_t integer_mode, double first, double last, double* xp,
                    double* cdf, int32_t* meta, void* stream);
/* norm_span = [f(dmin), f(dmax)] (+ _rescale_discrete_levels, __init__.py:232-248) written to span[2] on the device */
int dsb_shade_norm_span(int32_t how, double dmin, double dmax, const double* xp, const double* cdf, const int32_t* meta,
                        int32_t rescale, double* span, void* stream);
/* _colorize + _interpolate_alpha (__init__.py:359-532): weighted colour mix over categories (f32) and alpha from the
 * transfer function of the total; out[npix] = r | g<<8 | b<<16 | a<<24.  clip_mode != 0: an explicit span - the totals
 * are clipped to [clip_lo, clip_hi] first (masked_clip_2d, :507-514), as float64 (1: zeros were masked) or uint64 (2). */
int dsb_shade_cat_colorize(const uint32_t* counts, const uint64_t* total, int64_t npix, int32_t ncat, const float* rgb,
                           uint32_t fallback_rgb, uint32_t baseline, uint64_t offset, int32_t mask_zero, int32_t how,
                           const double* xp, const double* cdf, const int32_t* meta, const double* span, double min_alpha,
                           double alpha, int32_t clip_mode, double clip_lo, double clip_hi, uint32_t* out, void* stream);
/* _interpolate (__init__.py:251-357) for list colormaps (ncolors >= 2) and single colours (ncolors == 1):
 * data[npix] = value - offset as f64 with NaN for masked pixels. */
int dsb_shade_map2d(const double* data, int64_t npix, int32_t how, const double* xp, const double* cdf, const int32_t* meta,
                    const double* span, int32_t ncolors, const double* cspan, const double* rs, const double* gs,
                    const double* bs, double min_alpha, double alpha, uint32_t* out, void* stream);

/* count() / by(cat, count()) on a u32 canvas of 1x..2x the L2 budget (config 3: 133 MB): one pass into packed counters in
 * `scratch` (2 bytes per cell + 48) - 8-bit fields first (a quarter of the u32 footprint), verified by a checksum (sum of
 * the fields == accepted hits) and then added into the canvas; on a mismatch (some cell took more than 255 hits) the pass
 * is redone with 16-bit fields, and past 65 535 hits with u32 REDs, each by flag-gated launches (no host round trip).  Same contract as dsb_points for a plan of exactly one COUNT accumulator; otherwise
 * DSB_ERR_UNSUPPORTED. */
int dsb_points_count16(const dsb_view* view, const void* x, const void* y, int32_t xy_dtype, int64_t n, int64_t row_offset,
                       const dsb_plan* plan, void* scratch, int64_t scratch_bytes, void* stream);

/* ---- post-shade image operations ------------------------------------------------------------------- */
/* Binary compositing operators on uint32 RGBA (datashader/composite.py:72-125) and on plain arrays (:150-168). */
typedef enum { DSB_COMP_OVER = 0, DSB_COMP_ADD = 1, DSB_COMP_SATURATE = 2, DSB_COMP_SOURCE = 3 } dsb_composite_op;
typedef enum { DSB_ARR_ADD = 0, DSB_ARR_MAX = 1, DSB_ARR_MIN = 2, DSB_ARR_SOURCE = 3 } dsb_array_op;
/* out[i] = op(src[i], dst ? dst[i] : dst_scalar): tf.stack (transfer_functions/__init__.py:139-144) and
 * tf.set_background (:766, over(img, background)). */
int dsb_composite(const uint32_t* src, const uint32_t* dst, uint32_t dst_scalar, int64_t n, int32_t how, uint32_t* out,
                  void* stream);
/* tf.spread of an Image (transfer_functions/__init__.py:771-915): mask is [w, w] u8, w odd.  Each output pixel folds
 * its sources in the reference's raster order, so non-commutative operators give bit-identical results. */
int dsb_spread_image(const uint32_t* img, int32_t H, int32_t W, const uint8_t* mask, int32_t w, int32_t how, uint32_t* out,
                     void* stream);
/* tf.spread of an aggregate [H, W] or [H, W, C] (every category layer on its own): the float kernel (NaN = empty)
 * for f32 / f64, the int kernel for i32 / i64, the zero-ignoring kernel for u32 (:825-877). */
int dsb_spread_array(const void* arr, int32_t dtype, int32_t H, int32_t W, int32_t C, const uint8_t* mask, int32_t w,
                     int32_t how, void* out, void* stream);
/* dynspread's density heuristic (_rgb_density / _array_density, :1004-1051): out2[0] = non-empty pixels, out2[1] =
 * those with another non-empty pixel within px; density = out2[1] / out2[0] (inf when out2[0] == 0). */
int dsb_density(const void* arr, int32_t dtype, int32_t is_image, int32_t H, int32_t W, int32_t px, uint64_t* out2,
                void* stream);

#ifdef __cplusplus
}
#endif
#endif
