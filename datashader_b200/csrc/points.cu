// Fused glyph projection + reduction for points (K1: the general path, L2-resident global atomics).
//
// Replaces Point._build_extend.extend_cuda (glyphs/points.py:188-221) together with the generated
// append() (compiler.py:321-475).  One pass over the x / y / value columns; every base reduction is a
// commutative accumulator (see dsb_op in include/dsb200.h), so a row costs one RED per base and no
// per-pixel mutex is needed (the reference spin-locks one for where/first, _cuda_utils.py:177-199).
#include "common.cuh"
#include "accum.cuh"
#include <stdlib.h>

struct PointsArgs {
  dsb_view v;
  const void* x;
  const void* y;
  long long n;
  long long row_offset;
  long long band_lo, band_hi;   // only pixels [band_lo, band_hi) are updated in this launch (L2 banding)
  dsb_plan plan;
};

// One point per thread per step, 4 independent steps in flight (coalesced 4-byte loads; the path is
// bound by the RED rate, not by load issue - see profiles/r01_ubench.md).
template <typename XY>
__global__ void __launch_bounds__(256) k_points_generic(const PointsArgs a) {
  const XY* __restrict__ x = (const XY*)a.x;
  const XY* __restrict__ y = (const XY*)a.y;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int ncat = a.plan.ncat;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < a.n; i0 += 4 * stride) {
    XY xs[4], ys[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      long long i = i0 + u * stride;
      if (i < a.n) { xs[u] = __ldcs(x + i); ys[u] = __ldcs(y + i); }
      else { xs[u] = (XY)NAN; ys[u] = (XY)NAN; }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      long long i = i0 + u * stride;
      long long cell = map_to_cell<XY>(a.v, xs[u], ys[u]);
      if (cell < a.band_lo || cell >= a.band_hi) continue;     // also drops cell == -1 (out of bounds / NaN)
      if (ncat > 0) {
        int c = load_cat(a.plan.cat, a.plan.cat_dtype, i);
        if (c < 0) c += ncat;                 // numba wraparound for agg[:, :, -1]
        if (c < 0 || c >= ncat) continue;
        cell = cell * ncat + c;
      }
      for (int k = 0; k < a.plan.nops; k++) apply_base(a.plan.ops[k], cell, i, a.row_offset + i);
    }
  }
}

static long long op_cell_bytes(int op) {
  switch (op) {
    case DSB_OP_COUNT: case DSB_OP_MAX32: case DSB_OP_MIN32: return 4;
    case DSB_OP_ANY: return 1;
    case DSB_OP_MATCHROW64: return 16;   // reads the finished key canvas as well
    default: return 8;
  }
}

// Bytes of accumulator canvas one launch may touch before banding kicks in.  Default 96 MB of the 126 MB L2
// (swept 16..144 MB on configs 3 and 5, profiles/r01b_l2_banding.md; the streamed input is read with
// ld.global.cs so it does not displace the canvas).  DSB_L2_BAND_MB overrides (0 disables).
static long long l2_band_budget_bytes() {
  static long long cached = -1;
  if (cached < 0) {
    const char* e = getenv("DSB_L2_BAND_MB");
    long long mb = e ? atoll(e) : 96;
    cached = mb * (1LL << 20);
  }
  return cached;
}

static int validate_plan(const dsb_plan* p) {
  if (!p || p->nops < 1 || p->nops > DSB_MAX_OPS) { dsb_set_error("dsb_points: bad plan (nops)"); return DSB_ERR_ARG; }
  for (int k = 0; k < p->nops; k++) {
    const dsb_base& b = p->ops[k];
    if (!b.agg) { dsb_set_error("dsb_points: op %d has no canvas", k); return DSB_ERR_ARG; }
    bool needs_val = !(b.op == DSB_OP_COUNT || b.op == DSB_OP_ANY || b.op == DSB_OP_MAXROW || b.op == DSB_OP_MINROW);
    if (needs_val && (b.val_dtype == DSB_NONE || !b.val)) { dsb_set_error("dsb_points: op %d needs a value column", k); return DSB_ERR_ARG; }
    if (b.val_dtype != DSB_NONE && !b.val) { dsb_set_error("dsb_points: op %d value pointer is null", k); return DSB_ERR_ARG; }
    if ((b.op == DSB_OP_MAX32 || b.op == DSB_OP_MIN32 || b.op == DSB_OP_ARGMAX32 || b.op == DSB_OP_ARGMIN32) &&
        !(b.val_dtype == DSB_F32 || (b.val_dtype >= DSB_I8 && b.val_dtype <= DSB_U32))) {
      dsb_set_error("dsb_points: op %d needs a <=32-bit value column", k); return DSB_ERR_ARG;
    }
    if (b.op == DSB_OP_MATCHROW64 && !b.aux) { dsb_set_error("dsb_points: MATCHROW64 needs aux"); return DSB_ERR_ARG; }
    if (b.op < DSB_OP_COUNT || b.op > DSB_OP_MATCHROW64) { dsb_set_error("dsb_points: unknown op %d", b.op); return DSB_ERR_ARG; }
  }
  if (p->ncat < 0 || (p->ncat > 0 && (!p->cat || p->cat_dtype == DSB_NONE))) { dsb_set_error("dsb_points: bad categorical plan"); return DSB_ERR_ARG; }
  return DSB_OK;
}

extern "C" int dsb_points(const dsb_view* view, const void* x, const void* y, int32_t xy_dtype, int64_t n,
                          int64_t row_offset, const dsb_plan* plan, void* stream) {
  if (!view || view->width <= 0 || view->height <= 0) { dsb_set_error("dsb_points: bad view"); return DSB_ERR_ARG; }
  if (n < 0 || n > (1LL << 32)) { dsb_set_error("dsb_points: n must be in [0, 2^32] per call"); return DSB_ERR_ARG; }
  if (n == 0) return (plan && plan->nops >= 1 && plan->nops <= DSB_MAX_OPS) ? DSB_OK : (dsb_set_error("dsb_points: bad plan (nops)"), DSB_ERR_ARG);
  int rc = validate_plan(plan);
  if (rc != DSB_OK) return rc;
  if (!x || !y) { dsb_set_error("dsb_points: null coordinate column"); return DSB_ERR_ARG; }
  if ((int64_t)view->width * view->height * (plan->ncat > 0 ? plan->ncat : 1) > (1LL << 40)) { dsb_set_error("dsb_points: canvas too large"); return DSB_ERR_ARG; }
  if (xy_dtype != DSB_F32 && xy_dtype != DSB_F64) { dsb_set_error("dsb_points: xy_dtype must be f32 or f64"); return DSB_ERR_ARG; }
  PointsArgs a;
  a.v = *view; a.x = x; a.y = y; a.n = n; a.row_offset = row_offset; a.plan = *plan;
  const int threads = 256;
  long long want = (n + (long long)threads * 4 - 1) / ((long long)threads * 4);
  long long cap = (long long)dsb_num_sms() * 8;   // 8 CTAs of 256 threads per SM: full occupancy, whole waves
  int grid = (int)(want < cap ? want : cap);
  cudaStream_t s = (cudaStream_t)stream;

  // L2 banding.  REDs into an L2-resident canvas run at ~170 G/s; once the canvases outgrow L2 every RED
  // becomes a DRAM read-modify-write (27-47 G/s measured, profiles/r01a_configs.md).  HBM read bandwidth is
  // nearly idle on this path, so when the accumulator footprint exceeds the budget the rows are re-read once
  // per band of canvas rows and only the points of that band are scattered: K passes, each L2-resident.
  long long bytes_per_pixel = 0;
  for (int k = 0; k < plan->nops; k++) bytes_per_pixel += op_cell_bytes(plan->ops[k].op);
  bytes_per_pixel *= (plan->ncat > 0 ? plan->ncat : 1);
  const long long npixels = (long long)view->width * view->height;
  const long long budget = l2_band_budget_bytes();
  long long nbands = 1;
  if (budget > 0 && bytes_per_pixel * npixels > budget && n >= (1LL << 22)) {
    nbands = (bytes_per_pixel * npixels + budget - 1) / budget;
    if (nbands > view->height) nbands = view->height;
    if (nbands > 64) nbands = 64;
  }
  const long long rows_per_band = (view->height + nbands - 1) / nbands;
  for (long long b = 0; b < nbands; b++) {
    a.band_lo = b * rows_per_band * view->width;
    a.band_hi = (b + 1) * rows_per_band * view->width;
    if (a.band_hi > npixels) a.band_hi = npixels;
    if (a.band_lo >= a.band_hi) break;
    if (xy_dtype == DSB_F32) k_points_generic<float><<<grid, threads, 0, s>>>(a);
    else k_points_generic<double><<<grid, threads, 0, s>>>(a);
    DSB_CUDA_CHECK_LAUNCH("dsb_points");
  }
  return DSB_OK;
}
