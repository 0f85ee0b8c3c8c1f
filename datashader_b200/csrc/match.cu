// where(max | min) of a float32 selector on canvases beyond L2, second pass ("which row holds the extreme").
//
// dsb_points answers where(max('v')) in one pass with packed {key32, row} 64-bit atomics (DSB_OP_ARGMAX32): on an 8192 x 8192
// canvas that is a 537 MB accumulator, five L2-banded re-reads of the columns and a global RED per hit (95 ms for 4e9 points).
// Split in two the extreme itself is a single 4-byte accumulator, which the routed path (routed.cu) computes without global
// atomics; what remains is this pass: every row whose key equals its pixel's finished extreme votes its global row with
// atomicMin - the earliest row among the ties, the row the reference's strict compare keeps (reductions.py:1178-1183,
// 1222-1227, 2009-2016).  Only ~1 row per pixel can match, but finding them naively gathers the key of every row's pixel from
// a 268 MB canvas.  A COARSE map removes that: the least extreme of every 16 x 16 pixel block (1 MB at 8192^2, L2-resident).
// A row that does not reach its block's entry cannot be the extreme of its own pixel and is dropped after one L2 hit; the
// survivors (~10 % for 60 uniform rows per pixel) pay the DRAM gather and, if equal, the RED.  Exact for any distribution:
// the filter only ever drops rows that provably lose.
#include "common.cuh"
#include "fastmap.cuh"
#include <limits.h>
#include <stdio.h>
#include <string.h>

constexpr int MB_SHIFT = 4;       // 16 x 16 pixels per coarse entry.  Measured at 8192^2, 4e9 rows (tools/bench_match.py, profiles/r02_where_two_pass.md):
                                  // 4 x 4 ... 32 x 32 all within 2 % - the pass is bound by its dependent loads, not by the survivors

static bool g_match_queue = true;      // the shared-memory / queued form of the second pass (k_points_match32q); false: the first form
void dsb_match_set_queue(bool on) { g_match_queue = on; }

struct MatchArgs {
  dsb_view v;
  FastMap fm;
  const float* x; const float* y; const float* val;
  long long n, row_offset;
  const int* keys;        // finished MAX32 / MIN32 key canvas [H, W]
  long long* rows;        // i64 canvas, DSB_OP_MINROW-initialised
  int* coarse;            // [ch, cw]
  int cw, ch, shift;
  unsigned int* notes;    // UPDATE form: DSB_NOTE_NEGZERO
  int kstride, koff;      // k_match_coarse reads keys[pixel * kstride + koff]: 1, 0 for a key32 canvas; 2, 1 for the key half of a packed
                          //   {key32, row} canvas (DSB_OP_ARGMAX32 / ARGMIN32)
};

// coarse[by][bx] = the least extreme key of the block: min of the maxima (IS_MAX) / max of the minima
template <bool IS_MAX>
__global__ void __launch_bounds__(256) k_match_coarse(const MatchArgs a) {
  __shared__ int part[8];
  const int B = 1 << a.shift;
  int k = IS_MAX ? INT_MAX : INT_MIN;
  // one warp per coarse entry when the block is small (<= 8 x 8), one CTA otherwise
  const bool per_warp = a.shift <= 3;
  const long long e = per_warp ? (long long)blockIdx.x * 8 + (threadIdx.x >> 5) : blockIdx.x;
  const int nthr = per_warp ? 32 : 256, t = per_warp ? (threadIdx.x & 31) : threadIdx.x;
  if (e < (long long)a.cw * a.ch) {
    const int bx = (int)(e % a.cw), by = (int)(e / a.cw);
    for (int p = t; p < B * B; p += nthr) {
      const int px = (bx << a.shift) + (p & (B - 1)), py = (by << a.shift) + (p >> a.shift);
      if (px < a.v.width && py < a.v.height) {
        const int q = a.keys[((long long)py * a.v.width + px) * a.kstride + a.koff];
        k = IS_MAX ? min(k, q) : max(k, q);
      }
    }
  }
  k = IS_MAX ? __reduce_min_sync(0xffffffffu, k) : __reduce_max_sync(0xffffffffu, k);
  if (per_warp) {
    if ((threadIdx.x & 31) == 0 && e < (long long)a.cw * a.ch) a.coarse[e] = k;
    return;
  }
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = k;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++) k = IS_MAX ? min(k, part[w]) : max(k, part[w]);
    a.coarse[blockIdx.x] = k;
  }
}

template <bool IS_MAX, int PPT, int CTAS>
__global__ void __launch_bounds__(256, CTAS) k_points_match32(const __grid_constant__ MatchArgs a) {
  const uint32_t W = (uint32_t)a.v.width, H = (uint32_t)a.v.height;
  const FastMap& fm = a.fm;
  const int* __restrict__ keys = a.keys;
  const int* __restrict__ coarse = a.coarse;
  const int sh = a.shift;
  auto exact = [&](float xv, float yv, float vv, long long i) {      // rows near a pixel edge: exact f64 mapping, no filter
    if (vv != vv) return;
    const int cell = map_exact_linear(a.v, xv, yv);
    if (cell >= 0 && __ldcg(keys + cell) == key32_from_f32(vv)) atomicMin(a.rows + cell, a.row_offset + i);
  };
  constexpr int NV = PPT / 4;                                        // 16-byte loads per column per step
  const float4* __restrict__ x4 = (const float4*)a.x;
  const float4* __restrict__ y4 = (const float4*)a.y;
  const float4* __restrict__ v4 = (const float4*)a.val;
  const long long n4 = a.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float4 nan4 = make_float4(NAN, NAN, NAN, NAN);
  for (long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += NV * stride) {
    float xs[PPT], ys[PPT], vs[PPT];
#pragma unroll
    for (int u = 0; u < NV; u++) {
      const long long q = i4 + u * stride;
      const bool in = q < n4;
      const float4 xa = in ? __ldcs(x4 + q) : nan4, ya = in ? __ldcs(y4 + q) : nan4, va = in ? __ldcs(v4 + q) : nan4;
      xs[4 * u] = xa.x; xs[4 * u + 1] = xa.y; xs[4 * u + 2] = xa.z; xs[4 * u + 3] = xa.w;
      ys[4 * u] = ya.x; ys[4 * u + 1] = ya.y; ys[4 * u + 2] = ya.z; ys[4 * u + 3] = ya.w;
      vs[4 * u] = va.x; vs[4 * u + 1] = va.y; vs[4 * u + 2] = va.z; vs[4 * u + 3] = va.w;
    }
    int cell[PPT], key[PPT], thr[PPT];
    uint32_t okm = 0, slow = 0;
#pragma unroll
    for (int k = 0; k < PPT; k++) {                   // the K2-tight mapping: see k_points_priv_tight
      const float xf = fmaf(xs[k], fm.sx, fm.tx), yf = fmaf(ys[k], fm.sy, fm.ty);
      const int xi = __float2int_rd(xf), yi = __float2int_rd(yf);
      const float dx = xf - (float)xi, dy = yf - (float)yi;
      const bool sure = dx >= fm.ex && dx <= fm.omex && dy >= fm.ey && dy <= fm.omey;
      const bool ok = sure && (uint32_t)xi < W && (uint32_t)yi < H && vs[k] == vs[k];
      cell[k] = ok ? yi * (int)W + xi : 0;
      thr[k] = __ldg(coarse + (ok ? (yi >> sh) * a.cw + (xi >> sh) : 0));      // independent L2 / L1 hits in flight
      key[k] = key32_from_f32(vs[k]);
      okm |= (uint32_t)ok << k;
      slow |= (uint32_t)(!sure && vs[k] == vs[k]) << k;
    }
    uint32_t live = 0;
#pragma unroll
    for (int k = 0; k < PPT; k++) live |= (uint32_t)(IS_MAX ? key[k] >= thr[k] : key[k] <= thr[k]) << k;
    live &= okm;
    if (live) {
      int cur[PPT];
#pragma unroll
      for (int k = 0; k < PPT; k++) cur[k] = (live >> k) & 1u ? __ldcg(keys + cell[k]) : (IS_MAX ? INT_MAX : INT_MIN);
#pragma unroll
      for (int k = 0; k < PPT; k++)
        if (((live >> k) & 1u) && cur[k] == key[k]) atomicMin(a.rows + cell[k], a.row_offset + 4 * (i4 + (k >> 2) * stride) + (k & 3));
    }
    if (slow) {
#pragma unroll
      for (int k = 0; k < PPT; k++)
        if (slow & (1u << k)) exact(xs[k], ys[k], vs[k], 4 * (i4 + (k >> 2) * stride) + (k & 3));
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {           // tail rows
    const long long i = (n4 << 2) + threadIdx.x;
    exact(a.x[i], a.y[i], a.val[i], i);
  }
}

// The shipped form of the second pass.  What the first form (k_points_match32 above, kept as the A/B arm: dsb_configure("match_queue", 0))
// taught: a global lookup per row costs an L1 tag cycle per LANE (281 G rows/s at best), and whatever only some rows need costs the
// whole warp each time one lane has it.  So (1) the coarse thresholds live in SHARED memory - 16-bit, blocks of 32 x 32 pixels, 128 KB at
// 8192^2; more rows pass than with the 16 x 16 map (ncu: DRAM 70 % busy with 64 x 64 blocks); (2) the rows that pass are not handled where they are found: each warp appends
// them to its own queue in shared memory ({pixel, key, row}) and gathers the pixels' keys 32 queue entries at a time, every lane
// busy; (3) the loop is k_points_priv_tight's: two vectors per thread per step, nothing kept per row but its queue entry.
constexpr int MQ_CAP = 64;                 // queue entries per warp (drained at 32): {pixel, value bits, row}
constexpr int MQ_THR_MAX = 65536;          // thresholds: the upper 16 bits of the keys (128 KB) - 32 x 32 pixel blocks at 8192^2

// UPDATE = true is the same loop serving max / min THEMSELVES on a small canvas with many rows per pixel (dsb_points_minmax_rest):
// `keys` then is the live accumulator, already holding the extreme of the head of the rows; a queued row replaces its pixel's key if
// it beats it.  The thresholds were taken from the head and only get staler - a row below a stale bound still cannot win.
// MODE 2: the same for where(max | min) on such a canvas - `keys` is the live packed {key32, row} accumulator (DSB_OP_ARGMAX32 /
// ARGMIN32: the earliest row among the ties), a queued row swaps itself in if its packed value beats the pixel's.
template <bool IS_MAX, int MODE>
__global__ void __launch_bounds__(1024, 1) k_points_match32q(const __grid_constant__ MatchArgs a) {
  constexpr bool UPDATE = MODE == 1;
  extern __shared__ int msm[];
  short* thr = (short*)msm;                                              // [cw * ch] key >> 16: order-preserving, so the test stays conservative
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* q = (uint32_t*)(msm + MQ_THR_MAX / 2) + warp * 3 * MQ_CAP;   // pixel[MQ_CAP], key[MQ_CAP], row[MQ_CAP] (row within the call)
  for (int j = threadIdx.x; j < a.cw * a.ch; j += blockDim.x) thr[j] = (short)(a.coarse[j] >> 16);
  __syncthreads();
  const uint32_t W = (uint32_t)a.v.width, H = (uint32_t)a.v.height;
  const FastMap& fm = a.fm;
  const int* __restrict__ keys = a.keys;
  const int sh = a.shift, cw = a.cw;
  int qn = 0;                                                            // warp-uniform
  bool negzero = false;                                                  // UPDATE: a candidate was -0.0 (DSB_NOTE_NEGZERO)
  auto settle = [&](uint32_t cell, uint32_t vbits, uint32_t row) {      // vbits: the row's float32 value as stored
    const int key = key32_from_f32(__uint_as_float(vbits));
    if (MODE == 2) {           // value first, the earliest row on ties: the row field of a max is complemented (accum.cuh, DSB_OP_ARGMAX32)
      const uint32_t grow = (uint32_t)(a.row_offset + (long long)row);
      const long long p = ((long long)key << 32) | (long long)(IS_MAX ? ~grow : grow);
      long long* canvas = (long long*)a.rows;
      const long long curp = __ldcg(canvas + cell);
      if (IS_MAX ? p > curp : p < curp) { if (IS_MAX) atomicMax(canvas + cell, p); else atomicMin(canvas + cell, p); }
      return;
    }
    const int cur = __ldcg(keys + cell);
    if (UPDATE) {
      if (IS_MAX ? key > cur : key < cur) { if (IS_MAX) atomicMax((int*)keys + cell, key); else atomicMin((int*)keys + cell, key); }
      // the keys fold -0.0 onto +0.0: a -0.0 that takes or ties the extreme puts the sign of that pixel's zero in doubt
      // (one that loses here cannot be the extreme)
      if (vbits == 0x80000000u && !(IS_MAX ? key < cur : key > cur)) negzero = true;
    } else if (cur == key) atomicMin(a.rows + cell, a.row_offset + (long long)row);
  };
  // (measured and dropped: consulting the finer 16 x 16 L2 map for queued rows before their DRAM gather - a third fewer gathers,
  // but a second dependent load per drain: 21.0 vs 19.6 ms)
  auto settle_queued = [&](int e) { settle(q[e], q[MQ_CAP + e], q[2 * MQ_CAP + e]); };
  auto drain32 = [&]() {                                                 // the first 32 entries, one per lane; the rest moves to the front
    __syncwarp();
    settle_queued(lane);
    const int rest = qn - 32;
    uint32_t c = 0, k = 0, r = 0;
    if (lane < rest) { c = q[32 + lane]; k = q[MQ_CAP + 32 + lane]; r = q[2 * MQ_CAP + 32 + lane]; }
    __syncwarp();
    if (lane < rest) { q[lane] = c; q[MQ_CAP + lane] = k; q[2 * MQ_CAP + lane] = r; }
    qn = rest;
  };
  // one row: 1 = the fast pixel is not certain (exact mapping, out of line); rows that reach their block's threshold are queued
  auto one = [&](float xv, float yv, float vv, uint32_t row) -> uint32_t {
    const float xf = fmaf(xv, fm.sx, fm.tx), yf = fmaf(yv, fm.sy, fm.ty);
    const int xi = __float2int_rd(xf), yi = __float2int_rd(yf);
    const float dx = xf - (float)xi, dy = yf - (float)yi;
    const bool sure = dx >= fm.ex && dx <= fm.omex && dy >= fm.ey && dy <= fm.omey;
    const bool inside = (uint32_t)xi < W && (uint32_t)yi < H;
    const int key = key32_from_f32(vv);
    const int t = thr[inside ? (yi >> sh) * cw + (xi >> sh) : 0];
    const bool live = sure && inside && vv == vv && (IS_MAX ? (key >> 16) >= t : (key >> 16) <= t);
    const unsigned bal = __ballot_sync(0xffffffffu, live);
    if (bal) {
      if (live) {
        const int pos = qn + __popc(bal & ((1u << lane) - 1u));
        q[pos] = (uint32_t)(yi * (int)W + xi); q[MQ_CAP + pos] = __float_as_uint(vv); q[2 * MQ_CAP + pos] = row;
      }
      qn += __popc(bal);
      if (qn >= 32) drain32();
    }
    return (uint32_t)(!sure && vv == vv);
  };
  auto exact = [&](float xv, float yv, float vv, long long i) {
    if (vv != vv) return;
    const int cell = map_exact_linear(a.v, xv, yv);
    if (cell >= 0) settle((uint32_t)cell, __float_as_uint(vv), (uint32_t)i);
  };
  const float4* __restrict__ x4 = (const float4*)a.x;
  const float4* __restrict__ y4 = (const float4*)a.y;
  const float4* __restrict__ v4 = (const float4*)a.val;
  const long long n4 = a.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // warp-uniform trip count (the ballots need every lane): whole steps only; what is left takes the exact path below
  long long w4 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
  for (; w4 + stride + 31 < n4; w4 += 2 * stride) {
    const long long i4 = w4 + lane;
    const float4 xa = __ldcs(x4 + i4), ya = __ldcs(y4 + i4), va = __ldcs(v4 + i4);
    const float4 xb = __ldcs(x4 + i4 + stride), yb = __ldcs(y4 + i4 + stride), vb = __ldcs(v4 + i4 + stride);
    const uint32_t ra = (uint32_t)(4 * i4), rb = (uint32_t)(4 * (i4 + stride));
    const uint32_t m = one(xa.x, ya.x, va.x, ra) | one(xa.y, ya.y, va.y, ra + 1) << 1 | one(xa.z, ya.z, va.z, ra + 2) << 2 |
                       one(xa.w, ya.w, va.w, ra + 3) << 3 | one(xb.x, yb.x, vb.x, rb) << 4 | one(xb.y, yb.y, vb.y, rb + 1) << 5 |
                       one(xb.z, yb.z, vb.z, rb + 2) << 6 | one(xb.w, yb.w, vb.w, rb + 3) << 7;
    if (m) {
      if (m & 1) exact(xa.x, ya.x, va.x, ra);
      if (m & 2) exact(xa.y, ya.y, va.y, ra + 1);
      if (m & 4) exact(xa.z, ya.z, va.z, ra + 2);
      if (m & 8) exact(xa.w, ya.w, va.w, ra + 3);
      if (m & 16) exact(xb.x, yb.x, vb.x, rb);
      if (m & 32) exact(xb.y, yb.y, vb.y, rb + 1);
      if (m & 64) exact(xb.z, yb.z, vb.z, rb + 2);
      if (m & 128) exact(xb.w, yb.w, vb.w, rb + 3);
    }
  }
  __syncwarp();
  if (lane < qn) settle_queued(lane);                                                      // what is left in the queue (< 32 entries)
  for (long long i4 = w4 + lane; i4 < n4; i4 += stride) {                                   // the last, partial steps
    const float4 xa = __ldcs(x4 + i4), ya = __ldcs(y4 + i4), va = __ldcs(v4 + i4);
    exact(xa.x, ya.x, va.x, 4 * i4); exact(xa.y, ya.y, va.y, 4 * i4 + 1); exact(xa.z, ya.z, va.z, 4 * i4 + 2); exact(xa.w, ya.w, va.w, 4 * i4 + 3);
  }
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {           // tail rows
    const long long i = (n4 << 2) + threadIdx.x;
    exact(a.x[i], a.y[i], a.val[i], i);
  }
  if (UPDATE && negzero && a.notes) *a.notes = DSB_NOTE_NEGZERO;
}

// any axes / alignment: the exact mapping for every row, no filter (log axes, or ranges beyond the float32 mapping's error bound)
template <int DUMMY>
__global__ void __launch_bounds__(256) k_points_match32_exact(const MatchArgs a) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    const float vv = a.val[i];
    if (vv != vv) continue;
    const long long cell = map_to_cell<float>(a.v, a.x[i], a.y[i]);
    if (cell >= 0 && __ldcg(a.keys + cell) == key32_from_f32(vv)) atomicMin(a.rows + cell, a.row_offset + i);
  }
}

extern "C" int64_t dsb_points_match32_scratch_bytes(const dsb_view* view) {
  if (!view || view->width <= 0 || view->height <= 0) return 0;
  const long long cw = (view->width + (1 << MB_SHIFT) - 1) >> MB_SHIFT, ch = (view->height + (1 << MB_SHIFT) - 1) >> MB_SHIFT;
  return cw * ch * 4;      // the 16 x 16-block map of the first form; the shared-memory form's map (<= 64 KB) is never larger
}

extern "C" int dsb_points_match32(const dsb_view* view, const void* x, const void* y, int32_t xy_dtype, int64_t n,
                                  int64_t row_offset, const void* val, int32_t val_dtype, const void* keys, int32_t is_max,
                                  void* rows, void* scratch, int64_t scratch_bytes, void* stream) {
  if (!view || view->width <= 0 || view->height <= 0) { dsb_set_error("dsb_points_match32: bad view"); return DSB_ERR_ARG; }
  if (!keys || !rows) { dsb_set_error("dsb_points_match32: null canvas"); return DSB_ERR_ARG; }
  if (n < 0 || n > (1LL << 32)) { dsb_set_error("dsb_points_match32: n must be in [0, 2^32] per call"); return DSB_ERR_ARG; }
  if (n == 0) return DSB_OK;
  if (!x || !y || !val) { dsb_set_error("dsb_points_match32: null column"); return DSB_ERR_ARG; }
  if (xy_dtype != DSB_F32 || val_dtype != DSB_F32) { dsb_set_error("dsb_points_match32: float32 columns only"); return DSB_ERR_UNSUPPORTED; }
  if ((long long)view->width * view->height >= (1LL << 31)) { dsb_set_error("dsb_points_match32: canvas too large"); return DSB_ERR_UNSUPPORTED; }
  MatchArgs a;
  a.v = *view;
  a.fm = make_fast_map(view);
  a.shift = MB_SHIFT;
  a.cw = (view->width + (1 << a.shift) - 1) >> a.shift; a.ch = (view->height + (1 << a.shift) - 1) >> a.shift;
  if (!scratch || scratch_bytes < dsb_points_match32_scratch_bytes(view)) { dsb_set_error("dsb_points_match32: scratch too small"); return DSB_ERR_ARG; }
  a.x = (const float*)x; a.y = (const float*)y; a.val = (const float*)val; a.n = n; a.row_offset = row_offset;
  a.keys = (const int*)keys; a.rows = (long long*)rows; a.coarse = (int*)scratch; a.notes = nullptr; a.kstride = 1; a.koff = 0;
  cudaStream_t s = (cudaStream_t)stream;
  const bool fast = a.fm.enabled && (((uintptr_t)x | (uintptr_t)y | (uintptr_t)val) & 15) == 0;
  {   // the label names both passes when the extreme was computed just before (bench.py reads it back as roofline.kernel)
    char prev[96];
    snprintf(prev, sizeof(prev), "%s", dsb_last_kernel());
    const bool chained = strstr(prev, is_max ? "max32" : "min32") != nullptr && strstr(prev, "k_points_match32") == nullptr;
    dsb_note_kernel(chained ? "%s<%s> after %s" : "%s<%s>", fast ? "k_points_match32" : "k_points_match32_exact", is_max ? "max" : "min", prev);
  }
  if (!fast) {
    k_points_match32_exact<0><<<dsb_num_sms() * 8, 256, 0, s>>>(a);
  } else {
    if (g_match_queue)        // thresholds in shared memory: the finest power-of-two blocks (>= 16 x 16) whose 16-bit map fits 128 KB (32 x 32 at 8192^2)
      while ((long long)((view->width + (1 << a.shift) - 1) >> a.shift) * ((view->height + (1 << a.shift) - 1) >> a.shift) > MQ_THR_MAX) a.shift++;
    a.cw = (view->width + (1 << a.shift) - 1) >> a.shift; a.ch = (view->height + (1 << a.shift) - 1) >> a.shift;
    const int cgrid = a.shift <= 3 ? (int)(((long long)a.cw * a.ch + 7) / 8) : a.cw * a.ch;
    if (is_max) k_match_coarse<true><<<cgrid, 256, 0, s>>>(a); else k_match_coarse<false><<<cgrid, 256, 0, s>>>(a);
    if (g_match_queue) {
      const size_t smem = (size_t)MQ_THR_MAX * 2 + 32 * 3 * MQ_CAP * 4;
      cudaFuncSetAttribute(k_points_match32q<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      cudaFuncSetAttribute(k_points_match32q<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (is_max) k_points_match32q<true, 0><<<dsb_num_sms(), 1024, smem, s>>>(a);
      else k_points_match32q<false, 0><<<dsb_num_sms(), 1024, smem, s>>>(a);
    } else {
      // first form: 4 rows per thread per step, 6 CTAs of 256 threads per SM: 23.4 ms for 4e9 rows against 28.0 ms with 8 rows per
      // step and 3 CTAs - the step is a chain of three dependent loads (columns, coarse entry, key) and wants warps, not registers
      if (is_max) k_points_match32<true, 4, 6><<<dsb_num_sms() * 6, 256, 0, s>>>(a);
      else k_points_match32<false, 4, 6><<<dsb_num_sms() * 6, 256, 0, s>>>(a);
    }
  }
  DSB_CUDA_CHECK_LAUNCH("dsb_points_match32");
  return DSB_OK;
}

// max / min of a float32 column on a canvas that fits L2, the REST of the rows after the head went through dsb_points: `keys` is the
// live DSB_OP_MAX32 / MIN32 accumulator.  Block thresholds (the least extreme of every block after the head, 16 bits, shared memory)
// drop the rows that cannot win - 98 % of them once a pixel has seen a few hundred rows; the others are queued per warp and compared
// with their pixel's key 32 at a time.  Linear axes inside the float32 mapping's error bound and 16-byte aligned columns only
// (DSB_ERR_UNSUPPORTED otherwise: the caller runs dsb_points over these rows as well).
static int minmax_rest(const dsb_view* view, const void* x, const void* y, int32_t xy_dtype, int64_t n, int64_t row_offset,
                       const void* val, int32_t val_dtype, void* keys, int32_t is_max, bool packed, unsigned int* notes, void* scratch,
                       int64_t scratch_bytes, void* stream) {
  if (!view || view->width <= 0 || view->height <= 0) { dsb_set_error("dsb_points_minmax_rest: bad view"); return DSB_ERR_ARG; }
  if (!keys) { dsb_set_error("dsb_points_minmax_rest: null canvas"); return DSB_ERR_ARG; }
  if (n < 0 || n > (1LL << 32)) { dsb_set_error("dsb_points_minmax_rest: n must be in [0, 2^32] per call"); return DSB_ERR_ARG; }
  if (n == 0) return DSB_OK;
  if (!x || !y || !val) { dsb_set_error("dsb_points_minmax_rest: null column"); return DSB_ERR_ARG; }
  if (xy_dtype != DSB_F32 || val_dtype != DSB_F32) { dsb_set_error("dsb_points_minmax_rest: float32 columns only"); return DSB_ERR_UNSUPPORTED; }
  if ((long long)view->width * view->height >= (1LL << 31)) { dsb_set_error("dsb_points_minmax_rest: canvas too large"); return DSB_ERR_UNSUPPORTED; }
  if (((uintptr_t)x | (uintptr_t)y | (uintptr_t)val) & 15) { dsb_set_error("dsb_points_minmax_rest: columns must be 16-byte aligned"); return DSB_ERR_UNSUPPORTED; }
  MatchArgs a;
  a.v = *view;
  a.fm = make_fast_map(view);
  if (!a.fm.enabled) { dsb_set_error("dsb_points_minmax_rest: linear axes inside the float32 mapping's error bound only"); return DSB_ERR_UNSUPPORTED; }
  a.shift = 1;              // the finest power-of-two blocks whose 16-bit map fits 128 KB: 4 x 4 pixels at 900 x 525
  while ((long long)((view->width + (1 << a.shift) - 1) >> a.shift) * ((view->height + (1 << a.shift) - 1) >> a.shift) > MQ_THR_MAX) a.shift++;
  a.cw = (view->width + (1 << a.shift) - 1) >> a.shift; a.ch = (view->height + (1 << a.shift) - 1) >> a.shift;
  if (!scratch || scratch_bytes < (long long)a.cw * a.ch * 4) { dsb_set_error("dsb_points_minmax_rest: scratch too small (%lld bytes needed)", (long long)a.cw * a.ch * 4); return DSB_ERR_ARG; }
  a.x = (const float*)x; a.y = (const float*)y; a.val = (const float*)val; a.n = n; a.row_offset = row_offset;
  a.keys = (const int*)keys; a.rows = nullptr; a.coarse = (int*)scratch; a.notes = notes; a.kstride = 1; a.koff = 0;
  if (packed) { a.rows = (long long*)keys; a.kstride = 2; a.koff = 1; a.notes = nullptr; }      // little endian: the key is the high word
  cudaStream_t s = (cudaStream_t)stream;
  {   // the head's kernel, copied first: dsb_note_kernel formats into the buffer dsb_last_kernel() returns
    char prev[104];
    snprintf(prev, sizeof(prev), "%s", dsb_last_kernel());
    dsb_note_kernel(packed ? "k_points_argminmax_rest<%s> after %.100s" : "k_points_minmax_rest<%s> after %.100s", is_max ? "max" : "min", prev);
  }
  const int cgrid = a.shift <= 3 ? (int)(((long long)a.cw * a.ch + 7) / 8) : a.cw * a.ch;
  if (is_max) k_match_coarse<true><<<cgrid, 256, 0, s>>>(a); else k_match_coarse<false><<<cgrid, 256, 0, s>>>(a);
  const size_t smem = (size_t)MQ_THR_MAX * 2 + 32 * 3 * MQ_CAP * 4;
  cudaFuncSetAttribute(k_points_match32q<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_points_match32q<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_points_match32q<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_points_match32q<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (packed) { if (is_max) k_points_match32q<true, 2><<<dsb_num_sms(), 1024, smem, s>>>(a); else k_points_match32q<false, 2><<<dsb_num_sms(), 1024, smem, s>>>(a); }
  else if (is_max) k_points_match32q<true, 1><<<dsb_num_sms(), 1024, smem, s>>>(a);
  else k_points_match32q<false, 1><<<dsb_num_sms(), 1024, smem, s>>>(a);
  DSB_CUDA_CHECK_LAUNCH("dsb_points_minmax_rest");
  return DSB_OK;
}

extern "C" int dsb_points_minmax_rest(const dsb_view* view, const void* x, const void* y, int32_t xy_dtype, int64_t n,
                                      int64_t row_offset, const void* val, int32_t val_dtype, void* keys, int32_t is_max,
                                      unsigned int* notes, void* scratch, int64_t scratch_bytes, void* stream) {
  return minmax_rest(view, x, y, xy_dtype, n, row_offset, val, val_dtype, keys, is_max, false, notes, scratch, scratch_bytes, stream);
}

// the same for where(max | min): `packed` is the live DSB_OP_ARGMAX32 / ARGMIN32 accumulator (i64 {key32, row} per pixel)
extern "C" int dsb_points_argminmax_rest(const dsb_view* view, const void* x, const void* y, int32_t xy_dtype, int64_t n,
                                         int64_t row_offset, const void* val, int32_t val_dtype, void* packed, int32_t is_max,
                                         void* scratch, int64_t scratch_bytes, void* stream) {
  return minmax_rest(view, x, y, xy_dtype, n, row_offset, val, val_dtype, packed, is_max, true, nullptr, scratch, scratch_bytes, stream);
}
