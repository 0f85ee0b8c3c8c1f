// Batched viewports: one pass over resident columns aggregates into V canvases of the same size at once - a tile level
// of a pyramid (tiles.py:70-131: every zoom level re-aggregates the same data into many 256 x 256 tiles) or the viewports
// of a zoom / pan interaction (pipeline.py:55-72).  Each view keeps its OWN scale / translate / bounds, so every canvas is
// bit-identical to the one Canvas.points would produce for that view alone; the canvases are stacked [V, H, W(, C)] and a
// hit in view t updates cell t * view_cells + cell of the plan's accumulators (the ops of accum.cuh, unchanged).
//
// Regular grids (nx * ny views, row-major, equal extents) find their view from the coarse grid index; a point within 1/1024 of
// a tile from a tile edge is also tested against the neighbour(s) across that edge with the exact mapping (a point ON a
// shared edge belongs to both tiles, as it does for separate calls; the margin covers the rounding of the coarse index and of the
// caller's tile extents, which the host checks against the grid).  Short lists of arbitrary views (<= 64) are all tested.
#include "common.cuh"
#include "accum.cuh"

struct ViewsArgs {
  const dsb_view* views;
  int nviews, grid_nx, grid_ny;
  double gx0, gy0, inv_gtw, inv_gth;
  const void* x; const void* y;
  long long n, row_offset, view_cells;
  dsb_plan plan;
};

template <typename XY>
__global__ void __launch_bounds__(256) k_points_views(const ViewsArgs a) {
  const XY* __restrict__ x = (const XY*)a.x;
  const XY* __restrict__ y = (const XY*)a.y;
  const int ncat = a.plan.ncat;
  auto hit = [&](int t, XY xv, XY yv, long long i) {
    long long cell = map_to_cell<XY>(a.views[t], xv, yv);
    if (cell < 0) return;
    if (ncat > 0) {
      int c = load_cat(a.plan.cat, a.plan.cat_dtype, i);
      if (c < 0) c += ncat;
      if (c < 0 || c >= ncat) return;
      cell = cell * ncat + c;
    }
    cell += (long long)t * a.view_cells;
    for (int k = 0; k < a.plan.nops; k++) apply_base<false>(a.plan.ops[k], cell, i, a.row_offset + i, a.plan.notes);
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (long long)gridDim.x * blockDim.x) {
    const XY xv = __ldcs(x + i), yv = __ldcs(y + i);
    if (a.grid_nx > 0) {
      const double fx = ((double)xv - a.gx0) * a.inv_gtw, fy = ((double)yv - a.gy0) * a.inv_gth;
      if (!(fx >= -1.0 && fx <= a.grid_nx + 1.0 && fy >= -1.0 && fy <= a.grid_ny + 1.0)) continue;    // NaN or far outside
      const double flx = floor(fx), fly = floor(fy);
      const int ix = (int)flx, iy = (int)fly;
      constexpr double D = 1.0 / 1024.0;
      const int x_lo = max(fx - flx < D ? ix - 1 : ix, 0), x_hi = min(fx - flx > 1.0 - D ? ix + 1 : ix, a.grid_nx - 1);
      const int y_lo = max(fy - fly < D ? iy - 1 : iy, 0), y_hi = min(fy - fly > 1.0 - D ? iy + 1 : iy, a.grid_ny - 1);
      for (int ty = y_lo; ty <= y_hi; ty++)
        for (int tx = x_lo; tx <= x_hi; tx++) hit(ty * a.grid_nx + tx, xv, yv, i);
    } else {
      for (int t = 0; t < a.nviews; t++) hit(t, xv, yv, i);
    }
  }
}

extern "C" int dsb_points_views(const dsb_view* views, int32_t nviews, int32_t grid_nx, int32_t grid_ny, double gx0, double gy0,
                                double gtw, double gth, const void* x, const void* y, int32_t xy_dtype, int64_t n, int64_t row_offset,
                                const dsb_plan* plan, int64_t view_cells, void* stream) {
  if (!views || nviews < 1) { dsb_set_error("dsb_points_views: no views"); return DSB_ERR_ARG; }
  if (!plan || plan->nops < 1 || plan->nops > DSB_MAX_OPS) { dsb_set_error("dsb_points_views: bad plan (nops)"); return DSB_ERR_ARG; }
  if (grid_nx > 0 ? ((int64_t)grid_nx * grid_ny != nviews || !(gtw > 0) || !(gth > 0)) : nviews > 64) {
    dsb_set_error("dsb_points_views: views must form the stated nx x ny grid, or be at most 64"); return DSB_ERR_ARG;
  }
  if (xy_dtype != DSB_F32 && xy_dtype != DSB_F64) { dsb_set_error("dsb_points_views: xy_dtype must be f32 or f64"); return DSB_ERR_ARG; }
  if (n < 0 || view_cells <= 0) { dsb_set_error("dsb_points_views: bad sizes"); return DSB_ERR_ARG; }
  for (int k = 0; k < plan->nops; k++) if (!plan->ops[k].agg) { dsb_set_error("dsb_points_views: op %d has no canvas", k); return DSB_ERR_ARG; }
  if (n == 0) return DSB_OK;
  if (!x || !y) { dsb_set_error("dsb_points_views: null coordinate column"); return DSB_ERR_ARG; }
  ViewsArgs a;
  a.views = views; a.nviews = nviews; a.grid_nx = grid_nx; a.grid_ny = grid_ny; a.gx0 = gx0; a.gy0 = gy0; a.inv_gtw = grid_nx > 0 ? 1.0 / gtw : 0.0; a.inv_gth = grid_nx > 0 ? 1.0 / gth : 0.0;
  a.x = x; a.y = y; a.n = n; a.row_offset = row_offset; a.view_cells = view_cells; a.plan = *plan;
  const long long want = (n + 255) / 256, cap = (long long)dsb_num_sms() * 8;
  const int grid = (int)(want < cap ? want : cap);
  dsb_note_kernel("k_points_views<%s> views=%d", xy_dtype == DSB_F32 ? "f32" : "f64", nviews);
  if (xy_dtype == DSB_F32) k_points_views<float><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  else k_points_views<double><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  DSB_CUDA_CHECK_LAUNCH("dsb_points_views");
  return DSB_OK;
}
