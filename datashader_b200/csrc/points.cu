// Fused glyph projection + reduction for points (K1: the general path, L2-resident global atomics).
//
// Replaces Point._build_extend.extend_cuda (glyphs/points.py:188-221) together with the generated
// append() (compiler.py:321-475).  One pass over the x / y / value columns; every base reduction is a
// commutative accumulator (see dsb_op in include/dsb200.h), so a row costs one RED per base and no
// per-pixel mutex is needed (the reference spin-locks one for where/first, _cuda_utils.py:177-199).
#include "common.cuh"
#include "accum.cuh"
#include "fastmap.cuh"
#include <stdlib.h>
#include <string.h>

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
template <typename XY, bool FILTER>
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
      for (int k = 0; k < a.plan.nops; k++) apply_base<FILTER>(a.plan.ops[k], cell, i, a.row_offset + i, a.plan.notes);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// K2: count() with the WHOLE canvas privatised per SM in shared memory.
//
// Shared-memory atomics run 6x faster than global REDs (profiles/r01_ubench.md) but a 900x525 u32 canvas is
// 1.89 MB.  Here every pixel gets a SLOT-bit field ((SLOT-1)-bit counter + 1 guard bit, 32/SLOT fields per word:
// SLOT=3 -> 472 500 pixels in 189 KB).  A hit is one ATOMS.ADD with return; the thread that sees the counter wrap
// into its guard bit owns the spill: it clears the guard (ATOMS) and adds 2^(SLOT-1) to the global scratch canvas
// (one RED per 2^(SLOT-1) hits).  At the end each CTA flushes its residual counters.  The guard bit absorbs hits
// that arrive between a wrap and its spill; a field can only be corrupted if 2^(SLOT-1) MORE hits land on the same
// pixel of the same SM inside that window, and the thread whose add overflows an all-ones field sees it (old ==
// all ones) and raises `flag`.  flag == 0 therefore proves exactness; otherwise the scratch is discarded and the
// count is redone with global REDs (k_points_generic gated on the flag), so the result is always exact.
// Other accumulators of the plan (e.g. the f64 sum of mean) run as global REDs in the same pass.
struct PrivArgs {
  PointsArgs p;
  int priv_op;              // index of the COUNT op held in shared memory
  const float* vcol;        // the float32 value column shared by the plan's ops (vector-loaded with x, y), or NULL
  long long npriv;          // cells [0, npriv) live in shared memory; the (few) cells beyond go straight to REDs
  unsigned int* scratch;    // [ncell] u32, zeroed: spills + flush land here
  unsigned int* flag;       // set to 1 on a carry event
};

template <int SLOT>
__device__ __forceinline__ void priv_hit(uint32_t* sh, long long cell, unsigned int* scratch, unsigned int& bad) {
  constexpr uint32_t PER = 32 / SLOT;
  constexpr uint32_t CNT_MASK = (1u << (SLOT - 1)) - 1u, GUARD = 1u << (SLOT - 1), FIELD = (1u << SLOT) - 1u;
  const uint32_t b = (uint32_t)cell;
  const uint32_t w = b / PER, sft = (b - w * PER) * SLOT;
  const uint32_t old = atomicAdd(sh + w, 1u << sft);
  const uint32_t f = (old >> sft) & FIELD;
  if (f == CNT_MASK) {
    atomicSub(sh + w, GUARD << sft);
    atomicAdd(scratch + b, GUARD);
  } else if (f == FIELD) {
    bad = 1;
  }
}

// apply_base for an op whose value column is the vector-loaded float32 column (value already in a register)
__device__ __forceinline__ void apply_base_f32(const dsb_base& b, long long cell, long long i, long long row, float v, unsigned int* notes) {
  if (b.chk_dtype != DSB_NONE) { apply_base<false>(b, cell, i, row, notes); return; }
  if (v != v) return;                       // every op below skips NaN fields
  switch (b.op) {
    case DSB_OP_COUNT: atomicAdd((unsigned int*)b.agg + cell, 1u); return;
    case DSB_OP_ANY: ((uint8_t*)b.agg)[cell] = 1; return;
    case DSB_OP_SUM: atomicAdd((double*)b.agg + cell, (double)v); return;
    // no load-before-RED filter here: with 32 warps per SM and 36 KB of L1 the dependent load costs K2 4x (measured)
    case DSB_OP_MAX32: if (notes && is_negzero(v)) *notes = DSB_NOTE_NEGZERO; atomicMax((int*)b.agg + cell, key32_from_f32(v)); return;
    case DSB_OP_MIN32: if (notes && is_negzero(v)) *notes = DSB_NOTE_NEGZERO; atomicMin((int*)b.agg + cell, key32_from_f32(v)); return;
    default: apply_base<false>(b, cell, i, row, notes); return;
  }
}

__device__ __forceinline__ long long map_to_cell_fast(const FastMap& f, const dsb_view& v, float x, float y) {
  if (!(x >= f.xlo && x <= f.xhi && y >= f.ylo && y <= f.yhi)) return -1;
  const float xf = fmaf(x, f.sx, f.tx), yf = fmaf(y, f.sy, f.ty);
  const float xr = floorf(xf), yr = floorf(yf);
  const float dx = xf - xr, dy = yf - yr;
  if (dx < f.ex || dx > 1.0f - f.ex || dy < f.ey || dy > 1.0f - f.ey) return map_to_cell<float>(v, x, y);
  return (long long)((int)yr * v.width + (int)xr);
}

// MODE 0: count() only; MODE 1: SUM + COUNT of one vector-loaded float32 column (mean); MODE 2: any plan
template <int SLOT, int MODE, bool VEC>
__global__ void __launch_bounds__(1024, 1) k_points_priv(const PrivArgs a, const FastMap fm) {
  extern __shared__ uint32_t sh[];
  constexpr uint32_t PER = 32 / SLOT;
  constexpr uint32_t FIELD = (1u << SLOT) - 1u;
  const PointsArgs& p = a.p;
  const long long ncell = a.npriv;
  const int nwords = (int)((ncell + PER - 1) / PER);
  for (int j = threadIdx.x; j < nwords; j += blockDim.x) sh[j] = 0;
  __syncthreads();
  const float* __restrict__ x = (const float*)p.x;
  const float* __restrict__ y = (const float*)p.y;
  const dsb_base& cop = p.plan.ops[a.priv_op];
  const int ncat = p.plan.ncat;
  double* __restrict__ sum_canvas = (MODE == 1) ? (double*)p.plan.ops[1 - a.priv_op].agg : nullptr;
  unsigned int bad = 0;
  constexpr bool have_v = VEC && MODE >= 1;

  auto one = [&](float xv, float yv, float vv, long long i) {
    long long cell = fm.enabled ? map_to_cell_fast(fm, p.v, xv, yv) : map_to_cell<float>(p.v, xv, yv);
    if (cell < 0) return;
    if (MODE == 0) {
      if (cell < a.npriv) priv_hit<SLOT>(sh, cell, a.scratch, bad);
      else atomicAdd(a.scratch + cell, 1u);
      return;
    }
    if (MODE == 1) {
      if (vv != vv) return;
      atomicAdd(sum_canvas + cell, (double)vv);
      if (cell < a.npriv) priv_hit<SLOT>(sh, cell, a.scratch, bad);
      else atomicAdd(a.scratch + cell, 1u);
      return;
    }
    if (ncat > 0) {
      int c = load_cat(p.plan.cat, p.plan.cat_dtype, i);
      if (c < 0) c += ncat;
      if (c < 0 || c >= ncat) return;
      cell = cell * ncat + c;
    }
    for (int k = 0; k < p.plan.nops; k++) {
      const dsb_base& b = p.plan.ops[k];
      const bool reg_v = have_v && a.vcol != nullptr && b.val == (const void*)a.vcol;
      if (k == a.priv_op) {
        if (cop.chk_dtype != DSB_NONE && col_isnan(cop.chk, cop.chk_dtype, i)) continue;
        if (reg_v) { if (vv != vv) continue; }
        else if (cop.val_dtype != DSB_NONE && col_isnan(cop.val, cop.val_dtype, i)) continue;
        if (cell < a.npriv) priv_hit<SLOT>(sh, cell, a.scratch, bad);
        else atomicAdd(a.scratch + cell, 1u);
      } else if (reg_v) {
        apply_base_f32(b, cell, i, p.row_offset + i, vv, p.plan.notes);
      } else {
        apply_base<false>(b, cell, i, p.row_offset + i, p.plan.notes);
      }
    }
  };

  if (VEC) {
    const float4* __restrict__ x4 = (const float4*)p.x;
    const float4* __restrict__ y4 = (const float4*)p.y;
    const float4* __restrict__ v4 = (const float4*)a.vcol;
    const bool lv = have_v && a.vcol != nullptr;
    const long long n4 = p.n >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float4 nan4 = make_float4(NAN, NAN, NAN, NAN);
    for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += 2 * stride) {
      const bool two = i4 + stride < n4;
      float4 xa = __ldcs(x4 + i4), ya = __ldcs(y4 + i4);
      float4 va = lv ? __ldcs(v4 + i4) : nan4;
      float4 xb = two ? __ldcs(x4 + i4 + stride) : nan4;
      float4 yb = two ? __ldcs(y4 + i4 + stride) : nan4;
      float4 vb = (two && lv) ? __ldcs(v4 + i4 + stride) : nan4;
      one(xa.x, ya.x, va.x, 4 * i4 + 0); one(xa.y, ya.y, va.y, 4 * i4 + 1);
      one(xa.z, ya.z, va.z, 4 * i4 + 2); one(xa.w, ya.w, va.w, 4 * i4 + 3);
      if (two) {
        const long long j4 = i4 + stride;
        one(xb.x, yb.x, vb.x, 4 * j4 + 0); one(xb.y, yb.y, vb.y, 4 * j4 + 1);
        one(xb.z, yb.z, vb.z, 4 * j4 + 2); one(xb.w, yb.w, vb.w, 4 * j4 + 3);
      }
    }
    if (blockIdx.x == 0 && threadIdx.x < (p.n & 3)) {           // tail rows
      const long long i = (n4 << 2) + threadIdx.x;
      one(x[i], y[i], lv ? a.vcol[i] : NAN, i);
    }
  } else {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) one(__ldcs(x + i), __ldcs(y + i), NAN, i);
  }

  __syncthreads();
  // flush the residual counters (guard bits are clear unless a spill was lost, which `bad` reports)
  for (long long j = threadIdx.x; j < ncell; j += blockDim.x) {
    const uint32_t w = (uint32_t)j / PER, sft = ((uint32_t)j - w * PER) * SLOT;
    const uint32_t c = (sh[w] >> sft) & FIELD;
    if (c) atomicAdd(a.scratch + j, c);
  }
  if (bad) atomicOr(a.flag, 1u);
}

// ---- K2 "tight": the two headline shapes (count(), mean(f32 column)) when the float32 fast mapping applies, there is
// no category axis and the whole canvas fits the packed shared-memory fields.  Same algorithm as k_points_priv with the
// per-point instruction stream cut down: 32-bit cell arithmetic, shared-window address formed once, the exact f64
// mapping kept out of line (taken by ~0.1 % of the points), no plan interpretation.

// One hit.  `ok` (0 / 1) gates it without a branch: the shared-memory add is unconditional, of ok << shift at a clamped
// address.  The every-2^(SLOT-1)-th-hit spill is the only branch; with 32 lanes some lane of nearly every warp takes
// it, so its operands are formed outside and it holds just the two memory instructions.
template <int SLOT, bool ALLP>
__device__ __forceinline__ void priv_hit_tight(uint32_t sh_addr, uint32_t cell, bool ok, uint32_t npriv, unsigned int* scratch,
                                               uint32_t& bad) {
  constexpr uint32_t PER = 32 / SLOT;
  constexpr uint32_t CNT_MASK = (1u << (SLOT - 1)) - 1u, GUARD = 1u << (SLOT - 1), FIELD = (1u << SLOT) - 1u;
  if (ALLP) {
    cell = min(cell, npriv - 1u);                  // only matters when !ok
  } else if (cell >= npriv) {                      // the few cells that did not fit shared memory: plain REDs
    if (ok) atomicAdd(scratch + cell, 1u);
    return;
  }
  const uint32_t w = cell / PER, sft = (cell - w * PER) * SLOT;
  const uint32_t addr = sh_addr + 4u * w;
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"((uint32_t)ok << sft) : "memory");
  const uint32_t f = (old >> sft) & FIELD;
  const uint32_t unguard = 0u - (GUARD << sft);
  unsigned int* const spill_to = scratch + cell;
  if (ok && f == CNT_MASK) {                       // this hit wrapped the counter into its guard bit
    asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(unguard) : "memory");
    asm volatile("red.global.add.u32 [%0], %1;" :: "l"(spill_to), "r"(GUARD) : "memory");
  }
  bad |= (uint32_t)(ok && f == FIELD);             // the add carried into the neighbouring field: redo exactly
}

template <int SLOT, bool MEAN, bool ALLP>
__global__ void __launch_bounds__(1024, 1) k_points_priv_tight(const __grid_constant__ PrivArgs a, const __grid_constant__ FastMap fm) {
  extern __shared__ uint32_t sh[];
  constexpr uint32_t PER = 32 / SLOT;
  constexpr uint32_t FIELD = (1u << SLOT) - 1u;
  const PointsArgs& p = a.p;
  const int ncell = (int)a.npriv;
  const int nwords = (ncell + (int)PER - 1) / (int)PER;
  for (int j = threadIdx.x; j < nwords; j += blockDim.x) sh[j] = 0;
  __syncthreads();
  uint32_t sh_addr;                                // opaque to the compiler so that it is formed once, not per hit
  asm volatile("mov.u32 %0, %1;" : "=r"(sh_addr) : "r"((uint32_t)__cvta_generic_to_shared(sh)));
  const float* __restrict__ x = (const float*)p.x;
  const float* __restrict__ y = (const float*)p.y;
  double* __restrict__ sum_canvas = MEAN ? (double*)p.plan.ops[1 - a.priv_op].agg : nullptr;
  const uint32_t W = (uint32_t)p.v.width, H = (uint32_t)p.v.height;
  uint32_t bad = 0;

  auto hit = [&](uint32_t cell, bool ok, float vv) {
    if (MEAN) {
      ok = ok && vv == vv;
      if (ok) atomicAdd(sum_canvas + cell, (double)vv);
    }
    priv_hit_tight<SLOT, ALLP>(sh_addr, cell, ok, (uint32_t)ncell, a.scratch, bad);
  };

  // xf is within fm.ex of the real-number value of the reference mapping for every point whose xf lands in [-1, W + 1]
  // (make_fast_map).  If its fractional part is in [ex, 1 - ex], floor(xf) is therefore the reference's pixel column
  // when 0 <= floor(xf) < W - and the point is strictly inside (xmin, xmax) - and the point is outside [xmin, xmax]
  // otherwise.  Everything else (near a pixel edge, NaN, inf) is deferred to the exact f64 mapping: returns 1.
  auto one = [&](float xv, float yv, float vv) -> uint32_t {
    const float xf = fmaf(xv, fm.sx, fm.tx), yf = fmaf(yv, fm.sy, fm.ty);
    const int xi = __float2int_rd(xf), yi = __float2int_rd(yf);
    const float dx = xf - (float)xi, dy = yf - (float)yi;
    const bool sure = dx >= fm.ex && dx <= fm.omex && dy >= fm.ey && dy <= fm.omey;
    const bool inside = (uint32_t)xi < W && (uint32_t)yi < H;
    hit((uint32_t)(yi * (int)W + xi), sure && inside, vv);
    return (uint32_t)!sure;
  };
  auto exact = [&](float xv, float yv, float vv) {
    const int cell = map_exact_linear(p.v, xv, yv);
    hit((uint32_t)cell, cell >= 0, vv);
  };

  const float4* __restrict__ x4 = (const float4*)p.x;
  const float4* __restrict__ y4 = (const float4*)p.y;
  const float4* __restrict__ v4 = (const float4*)a.vcol;
  const long long n4 = p.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float4 nan4 = make_float4(NAN, NAN, NAN, NAN);
  // two vectors (8 points) per thread per step, both loads issued before the first use; the loop carries no
  // predicates - the odd vector left at the end is handled after it
  long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i4 + stride < n4; i4 += 2 * stride) {
    const float4 xa = __ldcs(x4 + i4), ya = __ldcs(y4 + i4);
    const float4 va = MEAN ? __ldcs(v4 + i4) : nan4;
    const float4 xb = __ldcs(x4 + i4 + stride), yb = __ldcs(y4 + i4 + stride);
    const float4 vb = MEAN ? __ldcs(v4 + i4 + stride) : nan4;
    uint32_t slow = one(xa.x, ya.x, va.x) | one(xa.y, ya.y, va.y) << 1 | one(xa.z, ya.z, va.z) << 2 | one(xa.w, ya.w, va.w) << 3 |
                    one(xb.x, yb.x, vb.x) << 4 | one(xb.y, yb.y, vb.y) << 5 | one(xb.z, yb.z, vb.z) << 6 | one(xb.w, yb.w, vb.w) << 7;
    if (slow) {                                    // ~0.07 % of the points
      if (slow & 1) exact(xa.x, ya.x, va.x);
      if (slow & 2) exact(xa.y, ya.y, va.y);
      if (slow & 4) exact(xa.z, ya.z, va.z);
      if (slow & 8) exact(xa.w, ya.w, va.w);
      if (slow & 16) exact(xb.x, yb.x, vb.x);
      if (slow & 32) exact(xb.y, yb.y, vb.y);
      if (slow & 64) exact(xb.z, yb.z, vb.z);
      if (slow & 128) exact(xb.w, yb.w, vb.w);
    }
  }
  if (i4 < n4) {
    const float4 xa = __ldcs(x4 + i4), ya = __ldcs(y4 + i4);
    const float4 va = MEAN ? __ldcs(v4 + i4) : nan4;
    exact(xa.x, ya.x, va.x); exact(xa.y, ya.y, va.y); exact(xa.z, ya.z, va.z); exact(xa.w, ya.w, va.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (p.n & 3)) {           // tail rows
    const long long i = (n4 << 2) + threadIdx.x;
    exact(x[i], y[i], MEAN ? a.vcol[i] : 0.0f);
  }

  __syncthreads();
  for (int j = threadIdx.x; j < ncell; j += blockDim.x) {
    const uint32_t w = (uint32_t)j / PER, sft = ((uint32_t)j - w * PER) * SLOT;
    const uint32_t c = (sh[w] >> sft) & FIELD;
    if (c) atomicAdd(a.scratch + j, c);
  }
  if (bad) atomicOr(a.flag, 1u);
}

// ---- K1 "mono": single monotone accumulator (max / min of a float32 column, first / last, where(max / min)) ----------
// The plans of max(col), min(col), first(col), last(col) and where(max / min (col), ...) hold ONE accumulator.  With
// float32 coordinates on linear axes and an L2-resident canvas this kernel replaces the plan interpreter of
// k_points_generic by the K2-tight front end (vector loads, float32 fast mapping with the exact f64 mapping deferred
// to the ~0.1 % of points near a pixel edge) and batches the load-before-RED filter: the eight current values of a
// batch are fetched first (eight independent L2 loads in flight), then compared, and only winners issue a RED.
enum { MONO_MAX32 = 0, MONO_MIN32 = 1, MONO_MINROW = 2, MONO_MAXROW = 3, MONO_ARGMAX32 = 4, MONO_ARGMIN32 = 5,
       MONO_COUNT = 6 };   // count([col]) rides along for the banded passes of big canvases (plain RED, no filter)

template <int OP> struct MonoT { typedef long long cell_t; };
template <> struct MonoT<MONO_MAX32> { typedef int cell_t; };
template <> struct MonoT<MONO_MIN32> { typedef int cell_t; };
template <> struct MonoT<MONO_COUNT> { typedef unsigned int cell_t; };

template <int OP, bool BANDED>
__global__ void __launch_bounds__(256, 3) k_points_mono(const __grid_constant__ PointsArgs a, const __grid_constant__ FastMap fm,
                                                        const float* __restrict__ vcol) {
  typedef typename MonoT<OP>::cell_t T;
  T* __restrict__ canvas = (T*)a.plan.ops[0].agg;
  const float* __restrict__ x = (const float*)a.x;
  const float* __restrict__ y = (const float*)a.y;
  const uint32_t W = (uint32_t)a.v.width, H = (uint32_t)a.v.height;
  constexpr bool IS_MAX = OP == MONO_MAX32 || OP == MONO_MAXROW || OP == MONO_ARGMAX32;
  // banded passes of big canvases: measured again in round 2 at 8192^2, 1e9 points, where(max): 24.1 ms unfiltered, 52.5 ms with a
  // filter load per point, 26.0 ms with the load predicated on band membership - the filter stays off
  constexpr bool FILTERED = !BANDED && OP != MONO_COUNT;
  // "last" = the largest row id: walking the rows forwards every hit wins and pays a RED; walked BACKWARDS the first
  // hit of a pixel is final and the filter removes the rest, as for "first"
  constexpr bool REVERSE = OP == MONO_MAXROW;

  bool negzero = false;          // a max / min candidate was -0.0: see DSB_NOTE_NEGZERO
  auto key_of = [&](float vv, long long i) -> T {
    const long long row = a.row_offset + i;
    if (OP == MONO_COUNT) return (T)1;
    if (OP == MONO_MAX32 || OP == MONO_MIN32) { negzero |= is_negzero(vv); return (T)key32_from_f32(vv); }
    if (OP == MONO_MINROW || OP == MONO_MAXROW) return (T)row;
    const long long k = (long long)key32_from_f32(vv) << 32;          // see apply_base: value first, earliest row on ties
    return (T)(OP == MONO_ARGMAX32 ? (k | (long long)(uint32_t)(~(uint32_t)row)) : (k | (long long)(uint32_t)row));
  };
  auto commit = [&](int cell, T key, T cur) {
    if constexpr (OP == MONO_COUNT) { atomicAdd(canvas + cell, (T)1); return; }
    else if (IS_MAX) { if (!FILTERED || key > cur) atomicMax(canvas + cell, key); }
    else { if (!FILTERED || key < cur) atomicMin(canvas + cell, key); }
  };
  auto exact = [&](float xv, float yv, float vv, long long i) {
    if (vv != vv) return;
    const int cell = map_exact_linear(a.v, xv, yv);
    if (cell < 0) return;
    if (BANDED && (cell < a.band_lo || cell >= a.band_hi)) return;
    commit(cell, key_of(vv, i), FILTERED ? __ldcg(canvas + cell) : (T)0);
  };

  const float4* __restrict__ x4 = (const float4*)a.x;
  const float4* __restrict__ y4 = (const float4*)a.y;
  const float4* __restrict__ v4 = (const float4*)vcol;
  const long long n4 = a.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float4 nan4 = make_float4(NAN, NAN, NAN, NAN);
  for (long long t4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; t4 < n4; t4 += 2 * stride) {
    const bool two = t4 + stride < n4;
    const long long i4 = REVERSE ? n4 - 1 - t4 : t4;
    const long long j4 = REVERSE ? i4 - stride : i4 + stride;
    const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);      // count() without a column: nothing to NaN-check
    const bool hasv = OP != MONO_COUNT || vcol != nullptr;      // compile-time true for everything but count()
    const float4 xa = __ldcs(x4 + i4), ya = __ldcs(y4 + i4), va = hasv ? __ldcs(v4 + i4) : one4;
    const float4 xb = two ? __ldcs(x4 + j4) : nan4, yb = two ? __ldcs(y4 + j4) : nan4;
    const float4 vb = two ? (hasv ? __ldcs(v4 + j4) : one4) : nan4;
    const float xs[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
    const float ys[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
    const float vs[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
    int cell[8];
    bool ok[8];
    uint32_t slow = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {                     // the K2-tight mapping: see k_points_priv_tight
      const float xf = fmaf(xs[k], fm.sx, fm.tx), yf = fmaf(ys[k], fm.sy, fm.ty);
      const int xi = __float2int_rd(xf), yi = __float2int_rd(yf);
      const float dx = xf - (float)xi, dy = yf - (float)yi;
      const bool sure = dx >= fm.ex && dx <= fm.omex && dy >= fm.ey && dy <= fm.omey;
      ok[k] = sure && (uint32_t)xi < W && (uint32_t)yi < H && vs[k] == vs[k];
      cell[k] = ok[k] ? yi * (int)W + xi : 0;
      if (BANDED) {                                   // L2 banding: this launch owns the canvas rows [band_lo, band_hi)
        ok[k] = ok[k] && cell[k] >= (int)a.band_lo && cell[k] < (int)a.band_hi;
        cell[k] = ok[k] ? cell[k] : (int)a.band_lo;
      }
      slow |= (uint32_t)(!sure && vs[k] == vs[k]) << k;
    }
    T cur[8];
    if (FILTERED) {
#pragma unroll
      for (int k = 0; k < 8; k++) cur[k] = __ldcg(canvas + cell[k]);   // eight independent L2 loads in flight
    }
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (ok[k]) commit(cell[k], key_of(vs[k], 4 * (k < 4 ? i4 : j4) + (k & 3)), FILTERED ? cur[k] : (T)0);
    if (slow) {
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (slow & (1u << k)) exact(xs[k], ys[k], vs[k], 4 * (k < 4 ? i4 : j4) + (k & 3));
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {           // tail rows
    const long long i = (n4 << 2) + threadIdx.x;
    exact(x[i], y[i], vcol ? vcol[i] : 1.f, i);
  }
  if ((OP == MONO_MAX32 || OP == MONO_MIN32) && negzero && a.plan.notes) *a.plan.notes = DSB_NOTE_NEGZERO;
}


// Returns true when the launch was taken by k_points_mono.
template <int OP>
static void launch_mono(const PointsArgs& a, const FastMap& fm, const float* vcol, bool banded, cudaStream_t s) {
  const int grid = dsb_num_sms() * 3;
  if (banded) k_points_mono<OP, true><<<grid, 256, 0, s>>>(a, fm, vcol);
  else k_points_mono<OP, false><<<grid, 256, 0, s>>>(a, fm, vcol);
}

static long long g_mono_min_rows = 1LL << 20;   // below this the generic kernel's launch is as fast

static bool try_launch_mono(const PointsArgs& a, int32_t xy_dtype, bool banded, cudaStream_t s) {
  const dsb_plan& p = a.plan;
  if (xy_dtype != DSB_F32 || p.nops != 1 || p.ncat != 0 || a.n < g_mono_min_rows) return false;
  const dsb_base& b = p.ops[0];
  int op = -1;
  const float* vcol = nullptr;
  switch (b.op) {
    case DSB_OP_MAX32: op = MONO_MAX32; break;
    case DSB_OP_MIN32: op = MONO_MIN32; break;
    case DSB_OP_ARGMAX32: op = MONO_ARGMAX32; break;
    case DSB_OP_ARGMIN32: op = MONO_ARGMIN32; break;
    case DSB_OP_MINROW: op = MONO_MINROW; break;
    case DSB_OP_MAXROW: op = MONO_MAXROW; break;
    case DSB_OP_COUNT: if (!banded) return false; op = MONO_COUNT; break;   // unbanded counts are RED-bound either way
    default: return false;
  }
  if (op == MONO_COUNT) {                                  // count(): optional float32 column to NaN-check
    if (b.chk_dtype != DSB_NONE || (b.val_dtype != DSB_NONE && b.val_dtype != DSB_F32)) return false;
    vcol = b.val_dtype == DSB_F32 ? (const float*)b.val : nullptr;
  } else if (op == MONO_MINROW || op == MONO_MAXROW) {     // first / last: the row id, gated by the nan-check column
    if (b.chk_dtype != DSB_F32 || !b.chk || b.val_dtype != DSB_NONE) return false;
    vcol = (const float*)b.chk;
  } else {
    if (b.val_dtype != DSB_F32 || !b.val || b.chk_dtype != DSB_NONE) return false;
    vcol = (const float*)b.val;
  }
  if ((((uintptr_t)a.x | (uintptr_t)a.y | (uintptr_t)vcol) & 15) != 0) return false;
  if ((long long)a.v.width * a.v.height >= (1LL << 31)) return false;
  const FastMap fm = make_fast_map(&a.v);
  if (!fm.enabled) return false;
  static const char* const names[] = {"max32", "min32", "minrow", "maxrow", "argmax32", "argmin32", "count"};
  dsb_note_kernel("k_points_mono<%s,%s>", names[op], banded ? "banded" : "filtered");
  switch (op) {
    case MONO_MAX32: launch_mono<MONO_MAX32>(a, fm, vcol, banded, s); break;
    case MONO_MIN32: launch_mono<MONO_MIN32>(a, fm, vcol, banded, s); break;
    case MONO_MINROW: launch_mono<MONO_MINROW>(a, fm, vcol, banded, s); break;
    case MONO_MAXROW: launch_mono<MONO_MAXROW>(a, fm, vcol, banded, s); break;
    case MONO_ARGMAX32: launch_mono<MONO_ARGMAX32>(a, fm, vcol, banded, s); break;
    case MONO_COUNT: launch_mono<MONO_COUNT>(a, fm, vcol, banded, s); break;
    default: launch_mono<MONO_ARGMIN32>(a, fm, vcol, banded, s); break;
  }
  return true;
}


// ---- K1 mono for float64 frames (pandas' default dtype): max / min of a float64 column, first / last gated by one ----------
// Exact f64 mapping (linear or log axes), two points per 16-byte load, four points per batch with their four filter
// loads in flight; key64 canvases (MAX64 / MIN64) or row-index canvases (MINROW / MAXROW; last walks the rows backwards).
enum { M64_MAX = 0, M64_MIN = 1, M64_MINROW = 2, M64_MAXROW = 3 };

template <int OP>
__global__ void __launch_bounds__(256, 3) k_points_mono_f64(const __grid_constant__ PointsArgs a, const double* __restrict__ vcol) {
  long long* __restrict__ canvas = (long long*)a.plan.ops[0].agg;
  constexpr bool IS_MAX = OP == M64_MAX || OP == M64_MAXROW;
  constexpr bool REVERSE = OP == M64_MAXROW;
  const double* __restrict__ x = (const double*)a.x;
  const double* __restrict__ y = (const double*)a.y;
  bool negzero = false;
  auto key_of = [&](double vv, long long i) -> long long {
    if (OP == M64_MAX || OP == M64_MIN) { negzero |= is_negzero(vv); return (long long)key64_from_f64(vv); }
    return a.row_offset + i;
  };
  auto commit = [&](long long cell, long long key, long long cur) {
    if (IS_MAX) { if (key > cur) atomicMax(canvas + cell, key); }
    else { if (key < cur) atomicMin(canvas + cell, key); }
  };
  const double2* __restrict__ x2 = (const double2*)a.x;
  const double2* __restrict__ y2 = (const double2*)a.y;
  const double2* __restrict__ v2 = (const double2*)vcol;
  const long long n2 = a.n >> 1;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const double2 nan2 = make_double2(NAN, NAN);
  for (long long t2 = (long long)blockIdx.x * blockDim.x + threadIdx.x; t2 < n2; t2 += 2 * stride) {
    const bool two = t2 + stride < n2;
    const long long i2 = REVERSE ? n2 - 1 - t2 : t2;
    const long long j2 = REVERSE ? i2 - stride : i2 + stride;
    const double2 xa = __ldcs(x2 + i2), ya = __ldcs(y2 + i2), va = __ldcs(v2 + i2);
    const double2 xb = two ? __ldcs(x2 + j2) : nan2, yb = two ? __ldcs(y2 + j2) : nan2, vb = two ? __ldcs(v2 + j2) : nan2;
    const double xs[4] = {xa.x, xa.y, xb.x, xb.y}, ys[4] = {ya.x, ya.y, yb.x, yb.y}, vs[4] = {va.x, va.y, vb.x, vb.y};
    long long cell[4], cur[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      cell[k] = (vs[k] == vs[k]) ? map_to_cell<double>(a.v, xs[k], ys[k]) : -1;
      if (cell[k] >= 0 && (cell[k] < a.band_lo || cell[k] >= a.band_hi)) cell[k] = -1;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) cur[k] = cell[k] >= 0 ? __ldcg(canvas + cell[k]) : 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (cell[k] >= 0) commit(cell[k], key_of(vs[k], 2 * (k < 2 ? i2 : j2) + (k & 1)), cur[k]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && (a.n & 1)) {
    const long long i = a.n - 1;
    const double vv = vcol[i];
    const long long c = (vv == vv) ? map_to_cell<double>(a.v, x[i], y[i]) : -1;
    if (c >= 0 && c >= a.band_lo && c < a.band_hi) commit(c, key_of(vv, i), __ldcg(canvas + c));
  }
  if ((OP == M64_MAX || OP == M64_MIN) && negzero && a.plan.notes) *a.plan.notes = DSB_NOTE_NEGZERO;
}

static bool try_launch_mono_f64(const PointsArgs& a, int32_t xy_dtype, cudaStream_t s) {
  const dsb_plan& p = a.plan;
  if (xy_dtype != DSB_F64 || p.nops != 1 || p.ncat != 0 || a.n < g_mono_min_rows) return false;
  const dsb_base& b = p.ops[0];
  const double* vcol = nullptr;
  int op;
  if (b.op == DSB_OP_MAX64 || b.op == DSB_OP_MIN64) {
    if (b.val_dtype != DSB_F64 || !b.val || b.chk_dtype != DSB_NONE) return false;
    vcol = (const double*)b.val; op = b.op == DSB_OP_MAX64 ? M64_MAX : M64_MIN;
  } else if (b.op == DSB_OP_MINROW || b.op == DSB_OP_MAXROW) {
    if (b.chk_dtype != DSB_F64 || !b.chk || b.val_dtype != DSB_NONE) return false;
    vcol = (const double*)b.chk; op = b.op == DSB_OP_MINROW ? M64_MINROW : M64_MAXROW;
  } else return false;
  if ((((uintptr_t)a.x | (uintptr_t)a.y | (uintptr_t)vcol) & 15) != 0) return false;
  const int grid = dsb_num_sms() * 3;
  static const char* const names64[] = {"max64", "min64", "minrow", "maxrow"};
  dsb_note_kernel("k_points_mono_f64<%s>", names64[op]);
  switch (op) {
    case M64_MAX: k_points_mono_f64<M64_MAX><<<grid, 256, 0, s>>>(a, vcol); break;
    case M64_MIN: k_points_mono_f64<M64_MIN><<<grid, 256, 0, s>>>(a, vcol); break;
    case M64_MINROW: k_points_mono_f64<M64_MINROW><<<grid, 256, 0, s>>>(a, vcol); break;
    default: k_points_mono_f64<M64_MAXROW><<<grid, 256, 0, s>>>(a, vcol); break;
  }
  return true;
}


// canvas += scratch when the privatised pass was exact (flag == 0)
__global__ void k_priv_commit(unsigned int* __restrict__ canvas, const unsigned int* __restrict__ scratch,
                              const unsigned int* __restrict__ flag, long long n) {
  if (*flag) return;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) { unsigned int v = scratch[i]; if (v) canvas[i] += v; }
}

// any(): the privatised pass counted hits (fewer than 2^32 per call, so no wrap); canvas |= (scratch > 0)
__global__ void k_priv_commit_any(uint8_t* __restrict__ canvas, const unsigned int* __restrict__ scratch,
                                  const unsigned int* __restrict__ flag, long long n) {
  if (*flag) return;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) if (scratch[i]) canvas[i] = 1;
}

// the exact redo of the privatised op with global REDs, only when the flag was raised
template <typename XY>
__global__ void __launch_bounds__(256) k_points_generic_if(const PointsArgs a, const unsigned int* __restrict__ flag) {
  if (*flag == 0) return;
  const XY* __restrict__ x = (const XY*)a.x;
  const XY* __restrict__ y = (const XY*)a.y;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int ncat = a.plan.ncat;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    long long cell = map_to_cell<XY>(a.v, __ldcs(x + i), __ldcs(y + i));
    if (cell < 0) continue;
    if (ncat > 0) {
      int c = load_cat(a.plan.cat, a.plan.cat_dtype, i);
      if (c < 0) c += ncat;
      if (c < 0 || c >= ncat) continue;
      cell = cell * ncat + c;
    }
    for (int k = 0; k < a.plan.nops; k++) apply_base(a.plan.ops[k], cell, i, a.row_offset + i, a.plan.notes);
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
static long long g_band_budget = -1;        // bytes; -1 = not initialised
static long long g_band_min_rows = 1LL << 22;
static long long g_priv_smem_kb = 192;        // shared memory the privatised canvas may take (see dsb_points_priv)
static long long g_priv_smem_kb_mean = 226;   // the same for the mean() shape
static long long g_split_bytes = 48LL << 20;  // plans whose canvases total more than this run one pass per accumulator
static long long g_count16_band_bytes = 0;   // dsb_points_count16: optional banding of the packed canvas (0 = off; measured on
                                             // config 3: 1 pass 8.5 ms, 2 bands 9.0 ms, 3 bands 13.4 ms - each pass pays the generic front end)
static int g_count8 = 1;                     // dsb_points_count16: try 8-bit packed counters first (tight front end only)
static int g_mono = 1;                       // use k_points_mono for single monotone accumulators
static int g_mono_banded = 1;                //   ... also for the L2-banded passes of big canvases (without the filter)
static int g_priv_threads = 1024;            // threads per CTA of the tight K2 kernels (one CTA per SM)
static int g_priv_tight = 1;                 // use k_points_priv_tight for the count() / mean(f32) shapes
static long long l2_band_budget_bytes() {
  if (g_band_budget < 0) {
    const char* e = getenv("DSB_L2_BAND_MB");
    long long mb = e ? atoll(e) : 96;
    g_band_budget = mb * (1LL << 20);
  }
  return g_band_budget;
}

// runtime knobs (tests and tuning): "l2_band_bytes" (0 disables banding), "band_min_rows"
void dsb_routed_set_head(long long rows_per_cell);   // routed.cu: first / last route this many rows per canvas cell and only filter the rest (0 = off)
void dsb_match_set_queue(bool on);   // match.cu: the queued shared-memory form of dsb_points_match32 (default) or its first form (A/B arm)
void dsb_routed_set_tma(bool on);   // routed.cu: pass 2 through the TMA ring (default) or plain loads (A/B arm)

extern "C" int dsb_configure(const char* key, int64_t value) {
  if (!key) { dsb_set_error("dsb_configure: null key"); return DSB_ERR_ARG; }
  if (!strcmp(key, "l2_band_bytes")) { g_band_budget = value; return DSB_OK; }
  if (!strcmp(key, "band_min_rows")) { g_band_min_rows = value; return DSB_OK; }
  if (!strcmp(key, "priv_tight")) { g_priv_tight = value != 0; return DSB_OK; }
  if (!strcmp(key, "priv_threads")) { if (value < 128 || value > 1024 || (value & 31)) { dsb_set_error("dsb_configure: priv_threads must be a multiple of 32 in [128, 1024]"); return DSB_ERR_ARG; } g_priv_threads = (int)value; return DSB_OK; }
  if (!strcmp(key, "mono")) { g_mono = value != 0; return DSB_OK; }
  if (!strcmp(key, "split_bytes")) { g_split_bytes = value; return DSB_OK; }
  if (!strcmp(key, "count16_band_bytes")) { g_count16_band_bytes = value; return DSB_OK; }
  if (!strcmp(key, "count8")) { g_count8 = value != 0; return DSB_OK; }
  if (!strcmp(key, "routed_tma")) { dsb_routed_set_tma(value != 0); return DSB_OK; }
  if (!strcmp(key, "match_queue")) { dsb_match_set_queue(value != 0); return DSB_OK; }
  if (!strcmp(key, "routed_head_per_cell")) { if (value < 0) { dsb_set_error("dsb_configure: routed_head_per_cell must be >= 0"); return DSB_ERR_ARG; } dsb_routed_set_head(value); return DSB_OK; }
  if (!strcmp(key, "mono_banded")) { g_mono_banded = value != 0; return DSB_OK; }
  if (!strcmp(key, "mono_min_rows")) { g_mono_min_rows = value; return DSB_OK; }
  if (!strcmp(key, "priv_smem_kb")) { if (value < 16 || value > 226) { dsb_set_error("dsb_configure: priv_smem_kb must be in [16, 226]"); return DSB_ERR_ARG; } g_priv_smem_kb = value; return DSB_OK; }
  if (!strcmp(key, "priv_smem_kb_mean")) { if (value < 16 || value > 226) { dsb_set_error("dsb_configure: priv_smem_kb_mean must be in [16, 226]"); return DSB_ERR_ARG; } g_priv_smem_kb_mean = value; return DSB_OK; }
  dsb_set_error("dsb_configure: unknown key %s", key);
  return DSB_ERR_ARG;
}

// Optional L2 persistence window for banded passes (DSB_L2_PERSIST=1 enables; off by default: measured gains were
// marginal and inconsistent - 10.8 -> 10.6 ms on config 3, profiles/r01b_l2_banding.md - because a banded pass is
// bound by K1's per-point issue cost and the RED rate, not by L2 misses).  The carve-out is set once per device.
static size_t g_l2_window = 0;
static bool l2_persist_enabled() {
  static thread_local int state = -1, state_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (state < 0 || state_dev != dev) {
    const char* e = getenv("DSB_L2_PERSIST");
    state = (e && atoi(e) == 1) ? 1 : 0;
    state_dev = dev;
    if (state) {
      int maxp = 0, maxw = 0;
      cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev);
      cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxAccessPolicyWindowSize, dev);
      if (maxp <= 0 || maxw <= 0 || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp) != cudaSuccess) { state = 0; cudaGetLastError(); }
      else g_l2_window = (size_t)maxw;
    }
  }
  return state == 1;
}
static size_t l2_max_window_bytes() { return g_l2_window; }

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
  // Accumulator splitting.  Several canvases that together outgrow what stays resident in L2 (by('cat', mean()): 60 MB
  // of f64 sums + 30 MB of counts) thrash it when they are updated in one pass.  The input columns are cheap to read
  // again (HBM idles on this path), so such a plan runs as one pass per accumulator, each with its canvas L2-resident.
  // MATCHROW64 reads another op's finished canvas and is already planned on its own by the host.
  if (plan->nops > 1 && g_split_bytes > 0 && n >= g_band_min_rows) {
    long long total = 0, biggest = 0;
    for (int k = 0; k < plan->nops; k++) {
      const long long b = op_cell_bytes(plan->ops[k].op) * (long long)view->width * view->height * (plan->ncat > 0 ? plan->ncat : 1);
      total += b;
      if (b > biggest) biggest = b;
    }
    if (total > g_split_bytes && biggest < total) {
      for (int k = 0; k < plan->nops; k++) {
        dsb_plan one = *plan;
        one.nops = 1;
        one.ops[0] = plan->ops[k];
        rc = dsb_points(view, x, y, xy_dtype, n, row_offset, &one, stream);
        if (rc != DSB_OK) return rc;
      }
      return DSB_OK;
    }
  }
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
  long long budget = l2_band_budget_bytes();
  // Measured on 8192^2 canvases, 1e9 points (profiles/r01b_l2_banding.md, sweep of 48..160 MB): 4-byte accumulators
  // (count, max / min keys) are fastest with bands of 2/3 of the budget (64 MB: 14.8 / 16.6 ms vs 19.4 / 18.8 at 96),
  // first / last row-index canvases with 4/3 (128 MB: 12.7 vs 18.8 ms - a min/max RED that does not change the value
  // leaves its sector clean, so larger bands do not pay write-backs); the packed where() keys stay at the budget.
  {
    bool all4 = true, rowonly = plan->nops == 1 && (plan->ops[0].op == DSB_OP_MINROW || plan->ops[0].op == DSB_OP_MAXROW);
    for (int k = 0; k < plan->nops; k++) all4 = all4 && op_cell_bytes(plan->ops[k].op) == 4;
    if (all4) budget = budget * 2 / 3;
    else if (rowonly) budget = budget * 4 / 3;
  }
  long long nbands = 1;
  if (budget > 0 && bytes_per_pixel * npixels > budget && n >= g_band_min_rows) {
    nbands = (bytes_per_pixel * npixels + budget - 1) / budget;
    if (nbands > view->height) nbands = view->height;
    if (nbands > 64) nbands = 64;
  }
  const long long rows_per_band = (view->height + nbands - 1) / nbands;
  int biggest_op = 0;
  for (int k = 1; k < plan->nops; k++) if (op_cell_bytes(plan->ops[k].op) > op_cell_bytes(plan->ops[biggest_op].op)) biggest_op = k;
  for (long long b = 0; b < nbands; b++) {
    a.band_lo = b * rows_per_band * view->width;
    a.band_hi = (b + 1) * rows_per_band * view->width;
    if (a.band_hi > npixels) a.band_hi = npixels;
    if (a.band_lo >= a.band_hi) break;
    if (g_mono && g_mono_banded && nbands > 1 && try_launch_mono(a, xy_dtype, true, s)) { DSB_CUDA_CHECK_LAUNCH("dsb_points(mono, banded)"); continue; }
    if (nbands > 1 && l2_persist_enabled()) {
      // pin this band of the (largest) accumulator canvas in L2: the launch carries an access-policy window whose
      // hits persist while everything else (the streamed columns) is treated as streaming
      const dsb_base& big = plan->ops[biggest_op];
      const long long cb = op_cell_bytes(big.op) * (plan->ncat > 0 ? plan->ncat : 1);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = s;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeAccessPolicyWindow;
      at[0].val.accessPolicyWindow.base_ptr = (char*)big.agg + a.band_lo * cb;
      size_t nb = (size_t)((a.band_hi - a.band_lo) * cb);
      const size_t maxw = l2_max_window_bytes();
      at[0].val.accessPolicyWindow.num_bytes = nb < maxw ? nb : maxw;
      at[0].val.accessPolicyWindow.hitRatio = 1.0f;
      at[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      at[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (xy_dtype == DSB_F32) cudaLaunchKernelEx(&cfg, k_points_generic<float, false>, a);
      else cudaLaunchKernelEx(&cfg, k_points_generic<double, false>, a);
    } else {
      // the load-before-RED filter of the monotone accumulators pays only while the canvases are L2-resident
      const bool filter = nbands == 1 && bytes_per_pixel * npixels <= (96LL << 20);
      dsb_note_kernel("k_points_generic<%s,%s> nops=%d bands=%lld", xy_dtype == DSB_F32 ? "f32" : "f64", filter ? "filtered" : "plain",
                      plan->nops, nbands);
      if (filter && g_mono && try_launch_mono(a, xy_dtype, false, s)) { DSB_CUDA_CHECK_LAUNCH("dsb_points(mono)"); continue; }
      if (filter && g_mono && try_launch_mono_f64(a, xy_dtype, s)) { DSB_CUDA_CHECK_LAUNCH("dsb_points(mono f64)"); continue; }
      if (xy_dtype == DSB_F32) { if (filter) k_points_generic<float, true><<<grid, threads, 0, s>>>(a); else k_points_generic<float, false><<<grid, threads, 0, s>>>(a); }
      else { if (filter) k_points_generic<double, true><<<grid, threads, 0, s>>>(a); else k_points_generic<double, false><<<grid, threads, 0, s>>>(a); }
    }
    DSB_CUDA_CHECK_LAUNCH("dsb_points");
  }
  return DSB_OK;
}


// ---- K2 entry point ------------------------------------------------------------------------------------
template <int SLOT, int MODE, bool VEC>
static void launch_priv_one(const PrivArgs& a, const FastMap& fm, size_t smem, cudaStream_t s) {
  // per launch: the attribute is per device, and one process may drive several devices
  cudaFuncSetAttribute(k_points_priv<SLOT, MODE, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  k_points_priv<SLOT, MODE, VEC><<<dsb_num_sms(), 1024, smem, s>>>(a, fm);
}

template <int SLOT, bool MEAN>
static void launch_priv_tight(const PrivArgs& a, const FastMap& fm, bool allp, size_t smem, cudaStream_t s) {
  if (allp) {
    cudaFuncSetAttribute(k_points_priv_tight<SLOT, MEAN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    k_points_priv_tight<SLOT, MEAN, true><<<dsb_num_sms(), g_priv_threads, smem, s>>>(a, fm);
  } else {
    cudaFuncSetAttribute(k_points_priv_tight<SLOT, MEAN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    k_points_priv_tight<SLOT, MEAN, false><<<dsb_num_sms(), g_priv_threads, smem, s>>>(a, fm);
  }
}


// ---- K2 for float64 coordinates (pandas' default dtype): count() and mean(float64 column) ----------------------------
// Same privatised counters; the mapping is the exact f64 one (there is no cheaper form for doubles), two points per
// 16-byte load, four loads in flight per column.  Streams 16 B / point (count) or 24 B / point (mean).
template <int SLOT, bool MEAN, bool ALLP>
__global__ void __launch_bounds__(1024, 1) k_points_priv_f64(const __grid_constant__ PrivArgs a, const double* __restrict__ vcol) {
  extern __shared__ uint32_t sh[];
  constexpr uint32_t PER = 32 / SLOT;
  constexpr uint32_t FIELD = (1u << SLOT) - 1u;
  const PointsArgs& p = a.p;
  const int ncell = (int)a.npriv;
  const int nwords = (ncell + (int)PER - 1) / (int)PER;
  for (int j = threadIdx.x; j < nwords; j += blockDim.x) sh[j] = 0;
  __syncthreads();
  uint32_t sh_addr;
  asm volatile("mov.u32 %0, %1;" : "=r"(sh_addr) : "r"((uint32_t)__cvta_generic_to_shared(sh)));
  const double* __restrict__ x = (const double*)p.x;
  const double* __restrict__ y = (const double*)p.y;
  double* __restrict__ sum_canvas = MEAN ? (double*)p.plan.ops[1 - a.priv_op].agg : nullptr;
  uint32_t bad = 0;
  auto one = [&](double xv, double yv, double vv) {
    const long long c = map_to_cell<double>(p.v, xv, yv);
    bool ok = c >= 0;
    if (MEAN) {
      ok = ok && vv == vv;
      if (ok) atomicAdd(sum_canvas + c, vv);
    }
    priv_hit_tight<SLOT, ALLP>(sh_addr, (uint32_t)c, ok, (uint32_t)ncell, a.scratch, bad);
  };
  const double2* __restrict__ x2 = (const double2*)p.x;
  const double2* __restrict__ y2 = (const double2*)p.y;
  const double2* __restrict__ v2 = (const double2*)vcol;
  const long long n2 = p.n >> 1;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const double2 nan2 = make_double2(NAN, NAN);
  long long i2 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i2 + stride < n2; i2 += 2 * stride) {
    const double2 xa = __ldcs(x2 + i2), ya = __ldcs(y2 + i2), va = MEAN ? __ldcs(v2 + i2) : nan2;
    const double2 xb = __ldcs(x2 + i2 + stride), yb = __ldcs(y2 + i2 + stride), vb = MEAN ? __ldcs(v2 + i2 + stride) : nan2;
    one(xa.x, ya.x, va.x); one(xa.y, ya.y, va.y); one(xb.x, yb.x, vb.x); one(xb.y, yb.y, vb.y);
  }
  if (i2 < n2) {
    const double2 xa = __ldcs(x2 + i2), ya = __ldcs(y2 + i2), va = MEAN ? __ldcs(v2 + i2) : nan2;
    one(xa.x, ya.x, va.x); one(xa.y, ya.y, va.y);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && (p.n & 1)) one(x[p.n - 1], y[p.n - 1], MEAN ? vcol[p.n - 1] : 0.0);
  __syncthreads();
  for (int j = threadIdx.x; j < ncell; j += blockDim.x) {
    const uint32_t w = (uint32_t)j / PER, sft = ((uint32_t)j - w * PER) * SLOT;
    const uint32_t c = (sh[w] >> sft) & FIELD;
    if (c) atomicAdd(a.scratch + j, c);
  }
  if (bad) atomicOr(a.flag, 1u);
}

template <int SLOT>
static void launch_priv_f64(const PrivArgs& a, bool mean, const double* vcol, size_t smem, cudaStream_t s) {
  const bool allp = a.npriv == a.p.band_hi;
#define DSB_F64_LAUNCH(M, A) do { cudaFuncSetAttribute(k_points_priv_f64<SLOT, M, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
    k_points_priv_f64<SLOT, M, A><<<dsb_num_sms(), 1024, smem, s>>>(a, vcol); } while (0)
  if (mean) { if (allp) DSB_F64_LAUNCH(true, true); else DSB_F64_LAUNCH(true, false); }
  else { if (allp) DSB_F64_LAUNCH(false, true); else DSB_F64_LAUNCH(false, false); }
#undef DSB_F64_LAUNCH
}

template <int SLOT>
static void launch_priv(const PrivArgs& a, const FastMap& fm, int mode, bool vec, bool tight, size_t smem, cudaStream_t s) {
  const bool allp = a.npriv == a.p.band_hi;          // band_hi = number of cells
  if (tight && mode == 0) launch_priv_tight<SLOT, false>(a, fm, allp, smem, s);
  else if (tight && mode == 1) launch_priv_tight<SLOT, true>(a, fm, allp, smem, s);
  else if (!vec) launch_priv_one<SLOT, 2, false>(a, fm, smem, s);      // unaligned columns: scalar loads, generic plan
  else if (mode == 0) launch_priv_one<SLOT, 0, true>(a, fm, smem, s);
  else if (mode == 1) launch_priv_one<SLOT, 1, true>(a, fm, smem, s);
  else launch_priv_one<SLOT, 2, true>(a, fm, smem, s);
}

extern "C" int dsb_points_priv(const dsb_view* view, const void* x, const void* y, int32_t xy_dtype, int64_t n,
                               int64_t row_offset, const dsb_plan* plan, int32_t priv_op, uint32_t* scratch,
                               uint32_t* flag, void* stream) {
  if (!view || view->width <= 0 || view->height <= 0) { dsb_set_error("dsb_points_priv: bad view"); return DSB_ERR_ARG; }
  if (!scratch || !flag) { dsb_set_error("dsb_points_priv: scratch/flag required"); return DSB_ERR_ARG; }
  if (n < 0 || n > (1LL << 32)) { dsb_set_error("dsb_points_priv: n must be in [0, 2^32] per call"); return DSB_ERR_ARG; }
  if (n == 0) return DSB_OK;
  int rc = validate_plan(plan);
  if (rc != DSB_OK) return rc;
  if (priv_op < 0 || priv_op >= plan->nops || !(plan->ops[priv_op].op == DSB_OP_COUNT || plan->ops[priv_op].op == DSB_OP_ANY)) {
    dsb_set_error("dsb_points_priv: priv_op must name a COUNT or ANY op"); return DSB_ERR_ARG;
  }
  const bool priv_any = plan->ops[priv_op].op == DSB_OP_ANY;
  if (priv_any && n == (1LL << 32)) { dsb_set_error("dsb_points_priv: an ANY op takes fewer than 2^32 rows per call"); return DSB_ERR_ARG; }
  if (!x || !y) { dsb_set_error("dsb_points_priv: null coordinate column"); return DSB_ERR_ARG; }
  const long long ncell = (long long)view->width * view->height * (plan->ncat > 0 ? plan->ncat : 1);
  if (xy_dtype != DSB_F32 && xy_dtype != DSB_F64) { dsb_set_error("dsb_points_priv: needs float32 or float64 coordinates"); return DSB_ERR_UNSUPPORTED; }
  const bool f64 = xy_dtype == DSB_F64;
  cudaStream_t s = (cudaStream_t)stream;
  PrivArgs a;
  a.p.v = *view; a.p.x = x; a.p.y = y; a.p.n = n; a.p.row_offset = row_offset; a.p.band_lo = 0; a.p.band_hi = ncell; a.p.plan = *plan;
  a.priv_op = priv_op; a.scratch = scratch; a.flag = flag;
  a.vcol = nullptr;       // the first 16-byte aligned float32 value column of the plan rides along with x, y
  for (int k = 0; k < plan->nops && !a.vcol; k++)
    if (plan->ops[k].val_dtype == DSB_F32 && plan->ops[k].val && ((uintptr_t)plan->ops[k].val & 15) == 0)
      a.vcol = (const float*)plan->ops[k].val;
  // compile-time specialisations of the two headline shapes
  int mode = 2;
  const dsb_base& c0 = plan->ops[priv_op];
  if (plan->ncat == 0 && plan->nops == 1 && c0.val_dtype == DSB_NONE && c0.chk_dtype == DSB_NONE) mode = 0;
  else if (plan->ncat == 0 && plan->nops == 2 && a.vcol) {
    const dsb_base& o = plan->ops[1 - priv_op];
    if (o.op == DSB_OP_SUM && o.val == (const void*)a.vcol && o.val_dtype == DSB_F32 && o.chk_dtype == DSB_NONE &&
        c0.val == (const void*)a.vcol && c0.val_dtype == DSB_F32 && c0.chk_dtype == DSB_NONE) mode = 1;
  }
  const double* vcol64 = nullptr;
  if (f64) {                  // float64 coordinates: count(), or SUM + COUNT of one aligned float64 column; nothing else
    if (plan->ncat != 0 || ((((uintptr_t)x | (uintptr_t)y)) & 15) != 0) { dsb_set_error("dsb_points_priv: float64 path needs aligned columns and no categories"); return DSB_ERR_UNSUPPORTED; }
    if (mode != 0) {
      mode = 2;
      if (plan->nops == 2) {
        const dsb_base& o = plan->ops[1 - priv_op];
        if (o.op == DSB_OP_SUM && o.val_dtype == DSB_F64 && o.val && o.chk_dtype == DSB_NONE && c0.val == o.val && c0.val_dtype == DSB_F64 &&
            c0.chk_dtype == DSB_NONE && (((uintptr_t)o.val) & 15) == 0) { mode = 1; vcol64 = (const double*)o.val; }
      }
      if (mode != 1) { dsb_set_error("dsb_points_priv: float64 path serves count() and mean(float64 column) only"); return DSB_ERR_UNSUPPORTED; }
    }
  }
  // Shared-memory budget.  count() streams 8 B / point and needs the rest of the 228 KB L1/shared array as L1 for its
  // loads in flight: 192 KB (measured: 226 KB drops count from 484 to 438 Gpts/s).  mean() is bound by the global REDs
  // instead (one f64 sum RED per point + one spill RED per 2^(SLOT-1) points), so a wider slot is worth more than
  // the L1: 226 KB (900x525: SLOT 3 -> 4, mean 134 -> 144 Gpts/s).
  const long long kb = (mode == 1 && g_priv_smem_kb_mean > g_priv_smem_kb) ? g_priv_smem_kb_mean : g_priv_smem_kb;
  const size_t max_words = (size_t)(kb * 1024) / 4;
  // widest slot (fewest spills) that keeps at least 97 % of the cells in shared memory; the remainder, if any,
  // is scattered with plain REDs
  int slot = 0;
  const int slots[5] = {8, 5, 4, 3, 2};
  for (int k = 0; k < 5; k++) {
    const long long per = 32 / slots[k];
    if ((long long)max_words * per * 100 >= ncell * 97) { slot = slots[k]; break; }
  }
  if (slot == 0) {
    dsb_set_error("dsb_points_priv: needs a canvas of at most %lld cells", (long long)max_words * 16);
    return DSB_ERR_UNSUPPORTED;
  }
  cudaMemsetAsync(scratch, 0, (size_t)ncell * 4, s);
  cudaMemsetAsync(flag, 0, 4, s);
  const bool vec = (((uintptr_t)x | (uintptr_t)y) & 15) == 0;
  const long long per = 32 / slot;
  a.npriv = ncell < (long long)max_words * per ? ncell : (long long)max_words * per;
  const size_t smem = (size_t)((a.npriv + per - 1) / per) * 4;
  const FastMap fm = make_fast_map(view);
  const bool vvec = mode != 1 || ((uintptr_t)a.vcol & 15) == 0;
  const bool tight = g_priv_tight && vec && vvec && mode <= 1 && fm.enabled && ncell < (1LL << 31);
  if (f64) dsb_note_kernel("k_points_priv_f64<%d,%s>", slot, mode == 1 ? "mean" : "count");
  else if (tight) dsb_note_kernel("k_points_priv_tight<%d,%s>", slot, mode == 1 ? "mean" : "count");
  else dsb_note_kernel("k_points_priv<%d,mode%d,%s>", slot, mode, vec ? "vec" : "scalar");
  if (f64) {
    if (ncell >= (1LL << 31)) { dsb_set_error("dsb_points_priv: canvas too large"); return DSB_ERR_UNSUPPORTED; }
    switch (slot) {
      case 8: launch_priv_f64<8>(a, mode == 1, vcol64, smem, s); break;
      case 5: launch_priv_f64<5>(a, mode == 1, vcol64, smem, s); break;
      case 4: launch_priv_f64<4>(a, mode == 1, vcol64, smem, s); break;
      case 3: launch_priv_f64<3>(a, mode == 1, vcol64, smem, s); break;
      default: launch_priv_f64<2>(a, mode == 1, vcol64, smem, s); break;
    }
  } else
  switch (slot) {
    case 8: launch_priv<8>(a, fm, mode, vec, tight, smem, s); break;
    case 5: launch_priv<5>(a, fm, mode, vec, tight, smem, s); break;
    case 4: launch_priv<4>(a, fm, mode, vec, tight, smem, s); break;
    case 3: launch_priv<3>(a, fm, mode, vec, tight, smem, s); break;
    default: launch_priv<2>(a, fm, mode, vec, tight, smem, s); break;
  }
  DSB_CUDA_CHECK_LAUNCH("dsb_points_priv");
  long long g = (ncell + 255) / 256, cap = (long long)dsb_num_sms() * 8;
  if (priv_any) k_priv_commit_any<<<(int)(g < cap ? g : cap), 256, 0, s>>>((uint8_t*)plan->ops[priv_op].agg, scratch, flag, ncell);
  else k_priv_commit<<<(int)(g < cap ? g : cap), 256, 0, s>>>((unsigned int*)plan->ops[priv_op].agg, scratch, flag, ncell);
  // exact redo of the count / any with global REDs / stores, executed only if a carry event was seen
  PointsArgs r = a.p;
  r.plan.nops = 1; r.plan.ops[0] = plan->ops[priv_op];
  long long want = (n + 255) / 256;
  int grid = (int)(want < cap ? want : cap);
  if (f64) k_points_generic_if<double><<<grid, 256, 0, s>>>(r, flag);
  else k_points_generic_if<float><<<grid, 256, 0, s>>>(r, flag);
  DSB_CUDA_CHECK_LAUNCH("dsb_points_priv(commit)");
  return DSB_OK;
}

// ---- count() on canvases between 1x and 2x the L2 budget: 16-bit packed counters ------------------------------------
// A u32 count canvas of 96..192 MB (config 3: by('cat', count()) with 16 categories at 1920x1080 = 133 MB) does not
// stay in L2, so dsb_points bands it and reads the rows twice.  Here the pass counts into 16-bit halves of a u32
// scratch canvas (half the footprint: L2-resident, ONE pass) with plain REDs of 1 or 1 << 16.  A half that wraps
// loses 65 536 (low half: carries into its neighbour) or 65 536 (high half: carries out), so the sum of all halves
// falls short of the number of accepted hits - which the pass counts on the side.  Equal sums prove that no half
// wrapped; then the halves are added into the real canvas.  Otherwise the scratch is discarded and the pass is redone
// with u32 REDs on the real canvas (flag-gated launch, no host round trip).  Always exact.
template <typename XY>
__global__ void __launch_bounds__(256) k_points_count16(const PointsArgs a, unsigned int* __restrict__ packed,
                                                        unsigned long long* __restrict__ accepted) {
  const XY* __restrict__ x = (const XY*)a.x;
  const XY* __restrict__ y = (const XY*)a.y;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int ncat = a.plan.ncat;
  const dsb_base& b = a.plan.ops[0];
  unsigned long long mine = 0;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < a.n; i0 += 4 * stride) {
    XY xs[4], ys[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long i = i0 + u * stride;
      if (i < a.n) { xs[u] = __ldcs(x + i); ys[u] = __ldcs(y + i); } else { xs[u] = (XY)NAN; ys[u] = (XY)NAN; }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const long long i = i0 + u * stride;
      long long cell = map_to_cell<XY>(a.v, xs[u], ys[u]);
      if (cell < 0) continue;
      if (ncat > 0) {
        int c = load_cat(a.plan.cat, a.plan.cat_dtype, i);
        if (c < 0) c += ncat;
        if (c < 0 || c >= ncat) continue;
        cell = cell * ncat + c;
      }
      if (cell < a.band_lo || cell >= a.band_hi) continue;        // packed-canvas banding (cells, category axis included)
      if (b.chk_dtype != DSB_NONE && col_isnan(b.chk, b.chk_dtype, i)) continue;
      if (b.val_dtype != DSB_NONE && col_isnan(b.val, b.val_dtype, i)) continue;
      atomicAdd(packed + (cell >> 1), (cell & 1) ? 0x10000u : 1u);
      mine++;
    }
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(accepted, mine);
}


// count16 with the tight front end: float32 coordinates on linear axes, optional 1-byte category codes (by('cat',
// count()): four codes per 32-bit load beside the float4 of x and y), optional float32 column to NaN-check.  Same
// contract as k_points_count16.
// BITS = 16: two counters per word; BITS = 8: four (half the footprint again: 33 MB at config 3, comfortably L2-resident).
// gate != nullptr: the launch runs only if *gate != 0 (the 16-bit stage after an 8-bit stage whose checksum failed).
template <bool CAT, int BITS>
__global__ void __launch_bounds__(256, 3) k_points_count16_tight(const __grid_constant__ PointsArgs a, const __grid_constant__ FastMap fm,
                                                                 const float* __restrict__ vcol, unsigned int* __restrict__ packed,
                                                                 unsigned long long* __restrict__ accepted, const unsigned int* __restrict__ gate) {
  if (gate && *gate == 0) return;
  const float* __restrict__ x = (const float*)a.x;
  const float* __restrict__ y = (const float*)a.y;
  const uint32_t W = (uint32_t)a.v.width, H = (uint32_t)a.v.height;
  const int ncat = a.plan.ncat;
  const bool cat_signed = a.plan.cat_dtype == DSB_I8;
  const uint8_t* __restrict__ cat = (const uint8_t*)a.plan.cat;
  unsigned long long mine = 0;
  auto code_of = [&](uint32_t byte) -> int {
    int c = cat_signed ? (int)(int8_t)byte : (int)byte;
    if (c < 0) c += ncat;                                 // numba wraparound for agg[:, :, -1]
    return (c < 0 || c >= ncat) ? -1 : c;
  };
  auto put = [&](long long cell) {
    if (BITS == 16) atomicAdd(packed + (cell >> 1), (cell & 1) ? 0x10000u : 1u);
    else atomicAdd(packed + (cell >> 2), 1u << (8 * (int)(cell & 3)));
    mine++;
  };
  auto exact = [&](float xv, float yv, float vv, int code) {
    if (vv != vv || (CAT && code < 0)) return;
    const int cell = map_exact_linear(a.v, xv, yv);
    if (cell < 0) return;
    put(CAT ? (long long)cell * ncat + code : (long long)cell);
  };
  const float4* __restrict__ x4 = (const float4*)a.x;
  const float4* __restrict__ y4 = (const float4*)a.y;
  const float4* __restrict__ v4 = (const float4*)vcol;
  const uint32_t* __restrict__ c4 = (const uint32_t*)a.plan.cat;
  const long long n4 = a.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);
  for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += stride) {
    const float4 xa = __ldcs(x4 + i4), ya = __ldcs(y4 + i4), va = vcol ? __ldcs(v4 + i4) : one4;
    const uint32_t cw = CAT ? __ldcs(c4 + i4) : 0u;
    const float xs[4] = {xa.x, xa.y, xa.z, xa.w}, ys[4] = {ya.x, ya.y, ya.z, ya.w}, vs[4] = {va.x, va.y, va.z, va.w};
    uint32_t slow = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float xf = fmaf(xs[k], fm.sx, fm.tx), yf = fmaf(ys[k], fm.sy, fm.ty);
      const int xi = __float2int_rd(xf), yi = __float2int_rd(yf);
      const float dx = xf - (float)xi, dy = yf - (float)yi;
      const bool sure = dx >= fm.ex && dx <= fm.omex && dy >= fm.ey && dy <= fm.omey;
      const int code = CAT ? code_of((cw >> (8 * k)) & 255u) : 0;
      const bool ok = sure && (uint32_t)xi < W && (uint32_t)yi < H && vs[k] == vs[k] && code >= 0;
      if (ok) {
        const long long cell = (long long)(yi * (int)W + xi);
        put(CAT ? cell * ncat + code : cell);
      }
      slow |= (uint32_t)(!sure) << k;
    }
    if (slow) {
#pragma unroll
      for (int k = 0; k < 4; k++)
        if (slow & (1u << k)) exact(xs[k], ys[k], vs[k], CAT ? code_of((cw >> (8 * k)) & 255u) : 0);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {           // tail rows
    const long long i = (n4 << 2) + threadIdx.x;
    exact(x[i], y[i], vcol ? vcol[i] : 1.f, CAT ? code_of(cat[i]) : 0);
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(accepted, mine);
}

template <int BITS>
__global__ void k_sum16(const unsigned int* __restrict__ packed, long long nwords, unsigned long long* __restrict__ total,
                        const unsigned int* __restrict__ gate) {
  if (gate && *gate == 0) return;
  unsigned long long t = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (long long)gridDim.x * blockDim.x) {
    const unsigned int w = packed[i];
    if (BITS == 16) t += (w & 0xffffu) + (w >> 16);
    else t += (w & 0xffu) + ((w >> 8) & 0xffu) + ((w >> 16) & 0xffu) + (w >> 24);
  }
  for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
  if ((threadIdx.x & 31) == 0 && t) atomicAdd(total, t);
}

// st[0] = accepted hits, st[1] = sum of the fields, st[2] = flag (1: a field wrapped, redo with the next wider counters)
template <int BITS>
__global__ void k_unpack16_if(unsigned int* __restrict__ canvas, const unsigned int* __restrict__ packed, long long ncell,
                              unsigned long long* __restrict__ st, const unsigned int* __restrict__ gate) {
  if (gate && *gate == 0) return;
  if (st[0] != st[1]) { if (blockIdx.x == 0 && threadIdx.x == 0) *(unsigned int*)(st + 2) = 1u; return; }
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (long long)gridDim.x * blockDim.x) {
    unsigned int h;
    if (BITS == 16) { const unsigned int w = packed[c >> 1]; h = (c & 1) ? (w >> 16) : (w & 0xffffu); }
    else h = (packed[c >> 2] >> (8 * (int)(c & 3))) & 0xffu;
    if (h) canvas[c] += h;
  }
}

extern "C" int dsb_points_count16(const dsb_view* view, const void* x, const void* y, int32_t xy_dtype, int64_t n,
                                  int64_t row_offset, const dsb_plan* plan, void* scratch, int64_t scratch_bytes, void* stream) {
  if (!view || view->width <= 0 || view->height <= 0) { dsb_set_error("dsb_points_count16: bad view"); return DSB_ERR_ARG; }
  if (n < 0 || n > (1LL << 32)) { dsb_set_error("dsb_points_count16: n must be in [0, 2^32] per call"); return DSB_ERR_ARG; }
  int rc = validate_plan(plan);
  if (rc != DSB_OK) return rc;
  if (plan->nops != 1 || plan->ops[0].op != DSB_OP_COUNT) { dsb_set_error("dsb_points_count16: the plan must be one COUNT accumulator"); return DSB_ERR_UNSUPPORTED; }
  if (xy_dtype != DSB_F32 && xy_dtype != DSB_F64) { dsb_set_error("dsb_points_count16: xy_dtype must be f32 or f64"); return DSB_ERR_ARG; }
  if (n == 0) return DSB_OK;
  if (!x || !y) { dsb_set_error("dsb_points_count16: null coordinate column"); return DSB_ERR_ARG; }
  const long long ncell = (long long)view->width * view->height * (plan->ncat > 0 ? plan->ncat : 1);
  const long long nwords = (ncell + 1) >> 1;
  if (!scratch || scratch_bytes < nwords * 4 + 48) { dsb_set_error("dsb_points_count16: scratch must hold %lld bytes", nwords * 4 + 48); return DSB_ERR_ARG; }
  unsigned long long* st8 = (unsigned long long*)scratch;              // [3] u64 state of the 8-bit stage, [3] of the 16-bit stage,
  unsigned long long* st = st8 + 3;                                    //   then the packed canvas (shared by both stages)
  unsigned int* packed = (unsigned int*)(st + 3);
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(scratch, 0, (size_t)(nwords * 4 + 48), s);
  PointsArgs a;
  a.v = *view; a.x = x; a.y = y; a.n = n; a.row_offset = row_offset; a.plan = *plan; a.band_lo = 0; a.band_hi = ncell;
  const int threads = 256;
  long long want = (n + (long long)threads * 4 - 1) / ((long long)threads * 4);
  const long long cap = (long long)dsb_num_sms() * 8;
  const int grid = (int)(want < cap ? want : cap);
  // the tight front end when it applies: float32 coordinates on linear axes, 1-byte category codes (or none), an optional
  // float32 NaN-check column, everything 16-byte (codes: 4-byte) aligned
  const dsb_base& cb = plan->ops[0];
  const FastMap fm = make_fast_map(view);
  const bool cat1 = plan->ncat > 0 && (plan->cat_dtype == DSB_I8 || plan->cat_dtype == DSB_U8) && (((uintptr_t)plan->cat) & 3) == 0;
  const bool tight_ok = g_priv_tight && xy_dtype == DSB_F32 && fm.enabled && ncell < (1LL << 31) && (plan->ncat == 0 || cat1) &&
                        cb.chk_dtype == DSB_NONE && (cb.val_dtype == DSB_NONE || (cb.val_dtype == DSB_F32 && (((uintptr_t)cb.val) & 15) == 0)) &&
                        ((((uintptr_t)x | (uintptr_t)y)) & 15) == 0 && g_count16_band_bytes == 0 && !l2_persist_enabled();
  dsb_note_kernel(tight_ok ? "k_points_count16_tight<%s>" : "k_points_count16<%s>", tight_ok ? (plan->ncat > 0 ? "cat" : "nocat") : (xy_dtype == DSB_F32 ? "f32" : "f64"));
  const unsigned int* gate16 = nullptr;          // non-null: the 16-bit stage runs only if the 8-bit stage's checksum failed
  if (tight_ok) {
    const float* vcol = cb.val_dtype == DSB_F32 ? (const float*)cb.val : nullptr;
    const int g3 = dsb_num_sms() * 3;
    if (g_count8) {
      // stage 0: 8-bit counters, a quarter of the u32 footprint (config 3: 33 MB, L2-resident with room to spare).  A cell
      // with more than 255 hits wraps, the sum of the fields then falls short of the accepted hits, and the 16-bit stage
      // below (gated on that flag) redoes the pass.
      const long long nwords8 = (ncell + 3) >> 2;
      if (plan->ncat > 0) k_points_count16_tight<true, 8><<<g3, 256, 0, s>>>(a, fm, vcol, packed, st8, nullptr);
      else k_points_count16_tight<false, 8><<<g3, 256, 0, s>>>(a, fm, vcol, packed, st8, nullptr);
      k_sum16<8><<<(int)cap, 256, 0, s>>>(packed, nwords8, st8 + 1, nullptr);
      k_unpack16_if<8><<<(int)cap, 256, 0, s>>>((unsigned int*)plan->ops[0].agg, packed, ncell, st8, nullptr);
      gate16 = (const unsigned int*)(st8 + 2);
      cudaMemsetAsync(packed, 0, (size_t)nwords * 4, s);
      dsb_note_kernel("k_points_count16_tight<%s,8 bit>", plan->ncat > 0 ? "cat" : "nocat");
    }
    if (plan->ncat > 0) k_points_count16_tight<true, 16><<<g3, 256, 0, s>>>(a, fm, vcol, packed, st, gate16);
    else k_points_count16_tight<false, 16><<<g3, 256, 0, s>>>(a, fm, vcol, packed, st, gate16);
  } else if (l2_persist_enabled()) {
    // opt-in (DSB_L2_PERSIST=1): pin the packed canvas in L2, everything else streams
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeAccessPolicyWindow;
    at[0].val.accessPolicyWindow.base_ptr = (void*)packed;
    const size_t nb = (size_t)nwords * 4, maxw = l2_max_window_bytes();
    at[0].val.accessPolicyWindow.num_bytes = nb < maxw ? nb : maxw;
    at[0].val.accessPolicyWindow.hitRatio = 1.0f;
    at[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    at[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (xy_dtype == DSB_F32) cudaLaunchKernelEx(&cfg, k_points_count16<float>, a, packed, st);
    else cudaLaunchKernelEx(&cfg, k_points_count16<double>, a, packed, st);
  } else {
    // optional banding of the packed canvas (off by default, see g_count16_band_bytes)
    long long nb16 = g_count16_band_bytes > 0 ? (nwords * 4 + g_count16_band_bytes - 1) / g_count16_band_bytes : 1;
    if (nb16 < 1) nb16 = 1;
    const long long cells_per_band = ((ncell + nb16 - 1) / nb16 + 1) & ~1LL;
    for (long long bnd = 0; bnd < nb16; bnd++) {
      a.band_lo = bnd * cells_per_band;
      a.band_hi = (bnd + 1) * cells_per_band < ncell ? (bnd + 1) * cells_per_band : ncell;
      if (a.band_lo >= a.band_hi) break;
      if (xy_dtype == DSB_F32) k_points_count16<float><<<grid, threads, 0, s>>>(a, packed, st);
      else k_points_count16<double><<<grid, threads, 0, s>>>(a, packed, st);
    }
    a.band_lo = 0; a.band_hi = ncell;
  }
  k_sum16<16><<<(int)cap, 256, 0, s>>>(packed, nwords, st + 1, gate16);
  k_unpack16_if<16><<<(int)cap, 256, 0, s>>>((unsigned int*)plan->ops[0].agg, packed, ncell, st, gate16);
  want = (n + 255) / 256;
  const int g2 = (int)(want < cap ? want : cap);
  if (xy_dtype == DSB_F32) k_points_generic_if<float><<<g2, 256, 0, s>>>(a, (const unsigned int*)(st + 2));
  else k_points_generic_if<double><<<g2, 256, 0, s>>>(a, (const unsigned int*)(st + 2));
  DSB_CUDA_CHECK_LAUNCH("dsb_points_count16");
  return DSB_OK;
}
