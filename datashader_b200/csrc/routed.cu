// Routed aggregation ("bin, then accumulate in shared memory") for single-accumulator plans on canvases that fit neither
// shared memory nor L2 (BASELINE config 5: 8192 x 8192).
//
// dsb_points L2-bands such canvases: the columns are re-read once per band (5 passes = 60 B / point at 8192^2) and every
// hit is still a global RED.  Here the points are routed instead (profiles/r02_routed.md):
//
//   sample   every 64th block of 1024 points is mapped and histogrammed per BUCKET (a contiguous range of <= 45 056 canvas
//            cells); a one-CTA planning kernel turns the estimate into per-bucket record capacities and offsets
//   pass 1   every CTA maps a tile of points (float32 fast mapping, exact f64 mapping for points near a pixel edge),
//            counting-sorts the tile by bucket in shared memory and appends each bucket's run to its region as coalesced
//            8-byte records (bucket << 16 | cell in bucket, payload); one global atomic per bucket per tile reserves the space
//   pass 2   a CTA takes one bucket at a time, keeps the bucket's cells in shared memory (u32 keys / rows / counts:
//            shared-memory atomics, ~500 Gpts/s) and folds the finished tile into the canvas with plain loads and stores
//
// Traffic 12 + 8 + 8 B / point, no global atomics at all on the fast path.  Records that do not fit their bucket's region
// (the sample underestimated it) are applied to the canvas directly with the same atomics dsb_points would use, so the
// result is exact for any distribution; pass 2 hands out buckets dynamically, largest shares first come first served.
#include "common.cuh"
#include "fastmap.cuh"
#include <limits.h>
#include <stdlib.h>

enum { R_MAX32 = 0, R_MIN32 = 1, R_MINROW = 2, R_MAXROW = 3, R_COUNT = 4 };

constexpr int RT = 512;                 // threads per CTA of pass 1 (two CTAs per SM)
constexpr int RPPT = 16;                // points per thread per tile
constexpr int RTILE = RT * RPPT;        // 8192 points per tile
constexpr uint32_t R_CPB_MAX = 45056;   // cells per bucket: 176 KB of u32 in pass 2

struct RouteArgs {
  dsb_view v;
  FastMap fm;
  const float* x; const float* y; const float* vcol;
  long long n, row_offset;
  uint32_t cpb, inv, sh, kmul, nb, ncell;     // inv, sh, kmul: route_key
  unsigned long long* recs;             // records; bucket b owns [off[b], off[b] + cap[b])
  const unsigned long long* off;        // [nb]
  const uint32_t* cap;                  // [nb]
  uint32_t* cursor;                     // [nb] records offered so far (beyond cap: applied to the canvas directly)
  uint32_t* queue;                      // pass 2: next bucket to hand out
  uint32_t* slow; uint32_t* slow_n;     // points whose fast pixel is not certain (near a pixel edge) as {x bits, y bits, payload}:
  uint32_t slow_cap;                    //   mapped exactly by k_route_slow (12-byte entries: it streams them instead of gathering rows)
  void* canvas;
  unsigned int* notes;
  const unsigned long long* gate;       // first / last, rest of the rows (dsb_points_routed): {sampled rows, sampled rows whose pixel is still open};
};                                      //   the routed kernels run only if the sample says the filter would not pay (route_gate_open)

// first / last: the routed pass over the REST of the rows is the fallback of the filtered pass (k_rows_rest); both are launched and
// the sample decides on the device which of them does the work - no host round trip
__device__ __forceinline__ bool rest_wants_route(const unsigned long long* g) { return g[1] * 8ull > g[0]; }
__device__ __forceinline__ bool route_gate_closed(const RouteArgs& a) { return a.gate != nullptr && !rest_wants_route(a.gate); }

// bucket << 16 | cell in bucket.  The bucket is cell / cpb by multiplication with m = ceil(2^(32 + sh) / cpb): exact for every
// cell below ncell because ncell * (m * cpb - 2^(32 + sh)) < 2^(32 + sh) (checked on the host, dsb_points_routed); the key
// is then cell + bucket * (65536 - cpb): three instructions.
__device__ __forceinline__ uint32_t route_key(const RouteArgs& a, uint32_t cell) {
  const uint32_t b = __umulhi(cell, a.inv) >> a.sh;
  return cell + b * a.kmul;
}

// the accumulator op on the canvas itself (overflow records of pass 1)
template <int OP>
__device__ __forceinline__ void route_direct(const RouteArgs& a, uint32_t cell, uint32_t payload) {
  if (OP == R_MAX32) atomicMax((int*)a.canvas + cell, key32_from_f32(__uint_as_float(payload)));
  else if (OP == R_MIN32) atomicMin((int*)a.canvas + cell, key32_from_f32(__uint_as_float(payload)));
  else if (OP == R_MINROW) atomicMin((long long*)a.canvas + cell, a.row_offset + (long long)payload);
  else if (OP == R_MAXROW) atomicMax((long long*)a.canvas + cell, a.row_offset + (long long)payload);
  else atomicAdd((unsigned int*)a.canvas + cell, 1u);
}

// ---- sample + plan ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_route_sample(const RouteArgs a, long long stride_blocks, uint32_t* __restrict__ hist) {
  // every stride_blocks-th block of 256 vectors (4 KB per column, fully used sectors) is mapped and histogrammed
  extern __shared__ uint32_t sh[];
  if (route_gate_closed(a)) return;
  for (uint32_t b = threadIdx.x; b < a.nb; b += blockDim.x) sh[b] = 0;
  __syncthreads();
  const float4* __restrict__ x4 = (const float4*)a.x;
  const float4* __restrict__ y4 = (const float4*)a.y;
  const long long n4 = a.n >> 2;
  for (long long blk = blockIdx.x; blk * stride_blocks * 256 < n4; blk += gridDim.x) {
    const long long i4 = blk * stride_blocks * 256 + threadIdx.x;
    if (i4 >= n4) continue;
    const float4 xa = __ldg(x4 + i4), ya = __ldg(y4 + i4);
    const float xs[4] = {xa.x, xa.y, xa.z, xa.w}, ys[4] = {ya.x, ya.y, ya.z, ya.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float xf = fmaf(xs[k], a.fm.sx, a.fm.tx), yf = fmaf(ys[k], a.fm.sy, a.fm.ty);
      const int xi = __float2int_rd(xf), yi = __float2int_rd(yf);
      if ((uint32_t)xi < (uint32_t)a.v.width && (uint32_t)yi < (uint32_t)a.v.height)
        atomicAdd(sh + (route_key(a, (uint32_t)(yi * a.v.width + xi)) >> 16), 1u);      // an estimate: the fast pixel is enough
    }
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < a.nb; b += blockDim.x) if (sh[b]) atomicAdd(hist + b, sh[b]);
}

// capacities from the sampled histogram (estimate + 8 standard deviations + slack, scaled down if the scratch is too
// small - the overflow path keeps the result exact), exclusive scan -> offsets; cursors and the queue are cleared
__global__ void __launch_bounds__(1024) k_route_plan(const uint32_t* __restrict__ hist, uint32_t nb, uint32_t stride_pts, unsigned long long capacity,
                                                     unsigned long long* __restrict__ off, uint32_t* __restrict__ cap,
                                                     uint32_t* __restrict__ cursor, uint32_t* __restrict__ queue, uint32_t* __restrict__ slow_n) {
  __shared__ unsigned long long part[1024];
  __shared__ double scale_sh;
  const int tid = threadIdx.x;
  const uint32_t per = (nb + 1023) / 1024;
  const uint32_t lo = tid * per, hi = min(lo + per, nb);
  auto want = [&](uint32_t b) -> unsigned long long {
    const double est = (double)hist[b] * stride_pts;
    return (unsigned long long)(est + 8.0 * sqrt(est * stride_pts) + 2048.0) & ~1ull;
  };
  unsigned long long mine = 0;
  for (uint32_t b = lo; b < hi; b++) mine += want(b);
  part[tid] = mine;
  __syncthreads();
  if (tid == 0) {
    unsigned long long run = 0;
    for (int i = 0; i < 1024; i++) { const unsigned long long t = part[i]; part[i] = run; run += t; }
    scale_sh = run > capacity ? (double)capacity / (double)run : 1.0;
  }
  __syncthreads();
  const double scale = scale_sh;
  unsigned long long run = ((unsigned long long)((double)part[tid] * scale) + 3) & ~1ull;     // rounded UP: regions never overlap
  for (uint32_t b = lo; b < hi; b++) {
    unsigned long long c = want(b);
    if (scale < 1.0) c = (unsigned long long)((double)c * scale) & ~1ull;
    if (c > 0xfffffff0ull) c = 0xfffffff0ull;
    off[b] = run; cap[b] = (uint32_t)c; cursor[b] = 0;
    run += c;
  }
  if (tid == 0) { *queue = 0; *slow_n = 0; }
}

// ---- pass 1 ----------------------------------------------------------------------------------------------------------
template <int OP>
__global__ void __launch_bounds__(RT, 2) k_route_bin(const __grid_constant__ RouteArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  if (route_gate_closed(a)) return;
  const uint32_t nbp = (a.nb + 2) & ~1u;                               // + the dummy bucket a.nb: filtered rows, no branch per point
  unsigned long long* rec = (unsigned long long*)smem;                 // [RTILE]
  uint4* run_of = (uint4*)(rec + RTILE);                               // [nbp] per bucket: {&recs[global offset of this tile's run
                                                                       //   minus base (lo, hi), base, records that fit}
  uint32_t* hist = (uint32_t*)(run_of + nbp);                          // [nbp]
  uint32_t* wsum = hist + nbp;                                         // [64]
  const int tid = threadIdx.x;
  const uint32_t W = (uint32_t)a.v.width, H = (uint32_t)a.v.height;
  const float4* __restrict__ x4 = (const float4*)a.x;
  const float4* __restrict__ y4 = (const float4*)a.y;
  const float4* __restrict__ v4 = (const float4*)a.vcol;
  const long long n4 = a.n >> 2;
  constexpr long long TILE4 = RTILE / 4;
  bool negzero = false;
  for (uint32_t b = tid; b < nbp; b += RT) hist[b] = 0;
  __syncthreads();

  for (long long t4 = (long long)blockIdx.x * TILE4; t4 < n4; t4 += (long long)gridDim.x * TILE4) {
    uint32_t key[RPPT], rank[RPPT], pay[RPPT];
    uint32_t unsure = 0;                                         // bit u * 4 + k: that point is left to k_route_slow
#pragma unroll
    for (int u = 0; u < RPPT / 4; u++) {
      const long long i4 = t4 + (long long)u * RT + tid;
      const bool inr = i4 < n4;                                // rows past the end re-read the last vector and are masked
      const long long i4c = inr ? i4 : n4 - 1;
      const uint32_t row0 = (uint32_t)i4 << 2;                 // the row within this call (n < 2^32)
      const float4 xa = __ldcs(x4 + i4c), ya = __ldcs(y4 + i4c);
      float4 va = make_float4(1.f, 1.f, 1.f, 1.f);
      if (v4) va = __ldcs(v4 + i4c);
      const float xs[4] = {xa.x, xa.y, xa.z, xa.w}, ys[4] = {ya.x, ya.y, ya.z, ya.w}, vs[4] = {va.x, va.y, va.z, va.w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        // the K2-tight mapping (k_points_priv_tight): a sure fractional part proves the pixel and the bounds decision
        const float xf = fmaf(xs[k], a.fm.sx, a.fm.tx), yf = fmaf(ys[k], a.fm.sy, a.fm.ty);
        const int xi = __float2int_rd(xf), yi = __float2int_rd(yf);
        const float dx = xf - (float)xi, dy = yf - (float)yi;
        const bool sure = dx >= a.fm.ex && dx <= a.fm.omex && dy >= a.fm.ey && dy <= a.fm.omey;
        const bool live = inr && vs[k] == vs[k];               // NaN rows are skipped by every op
        unsure |= (uint32_t)(!sure && live) << (u * 4 + k);    // ~0.1 % of the points
        const bool ok = sure && live && (uint32_t)xi < W && (uint32_t)yi < H;
        const uint32_t kk = ok ? route_key(a, (uint32_t)(yi * (int)W + xi)) : (a.nb << 16);
        key[u * 4 + k] = kk;
        if (OP == R_MAX32 || OP == R_MIN32) { pay[u * 4 + k] = __float_as_uint(vs[k]); negzero |= ok && is_negzero(vs[k]); }
        else pay[u * 4 + k] = row0 + k;
        rank[u * 4 + k] = atomicAdd(hist + (kk >> 16), 1u);
      }
    }
    while (unsure) {                                           // out of the unrolled body: queue the points for k_route_slow
      const int j = __ffs(unsure) - 1;
      unsure &= unsure - 1;
      const uint32_t row = (uint32_t)(4 * (t4 + (long long)(j >> 2) * RT + tid) + (j & 3));
      const float xr = a.x[row], yr = a.y[row];                // re-read (the streaming loads above do not stay in L2: +1.1 GB of DRAM
      const float vv = a.vcol ? a.vcol[row] : 1.f;             //   reads per 1e9 points at 8192^2, against 3 GB of row gathers in k_route_slow)
      const uint32_t payload = (OP == R_MAX32 || OP == R_MIN32) ? __float_as_uint(vv) : row;
      // one counter update per warp iteration, not per lane (0.8 % of 1e9 points on ONE address otherwise)
      const unsigned m = __activemask();
      const int leader = __ffs(m) - 1, lane = tid & 31;
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(a.slow_n, (uint32_t)__popc(m));
      const uint32_t pos = __shfl_sync(m, base, leader) + (uint32_t)__popc(m & ((1u << lane) - 1u));
      if (pos < a.slow_cap) {
        uint32_t* e = a.slow + 3 * (size_t)pos;
        e[0] = __float_as_uint(xr); e[1] = __float_as_uint(yr); e[2] = payload;
      } else {                                                 // list full (adversarial data): map it here, straight to the canvas
        const int cell = map_exact_linear(a.v, xr, yr);
        if (cell >= 0) route_direct<OP>(a, (uint32_t)cell, payload);
        if (OP == R_MAX32 || OP == R_MIN32) negzero |= cell >= 0 && is_negzero(vv);
      }
    }
    __syncthreads();
    // block-wide exclusive scan of hist; every thread owns the entries [tid * per, tid * per + per)
    const uint32_t per = (a.nb + RT - 1) / RT;
    const uint32_t lo = tid * per, hi = min(lo + per, a.nb);
    uint32_t mine = 0;
    for (uint32_t b = lo; b < hi; b++) mine += hist[b];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += t; }
    if ((tid & 31) == 31) wsum[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
      const uint32_t w = tid < RT / 32 ? wsum[tid] : 0u;
      uint32_t wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (tid >= o) wi += t; }
      wsum[32 + tid] = wi - w;
    }
    __syncthreads();
    uint32_t run = wsum[32 + (tid >> 5)] + incl - mine;
    for (uint32_t b = lo; b < hi; b++) {
      const uint32_t h = hist[b];
      uint4 r = make_uint4(0u, 0u, run, 0u);
      if (h) {
        const uint32_t g = atomicAdd(a.cursor + b, h);         // one global atomic per non-empty bucket per tile
        const uint32_t c = a.cap[b];
        const unsigned long long o = (unsigned long long)(a.recs + (a.off[b] + g - run));
        r.x = (uint32_t)o; r.y = (uint32_t)(o >> 32); r.w = g >= c ? 0u : min(h, c - g);
      }
      run_of[b] = r;
      hist[b] = 0;                                             // ready for the next tile
      run += h;
    }
    if (tid == RT - 1) {                                       // records of the tile; the dummy bucket's run lies behind them
      wsum[31] = run;
      run_of[a.nb] = make_uint4(0u, 0u, run, 0u);
      hist[a.nb] = 0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RPPT; k++)
      rec[run_of[key[k] >> 16].z + rank[k]] = ((unsigned long long)pay[k] << 32) | key[k];
    __syncthreads();
    const uint32_t total = wsum[31];
#pragma unroll
    for (int u = 0; u < RPPT; u += 4) {                        // four records in flight per thread
      unsigned long long rr[4];
#pragma unroll
      for (int k = 0; k < 4; k++) { const uint32_t q = (u + k) * RT + tid; rr[k] = q < total ? rec[q] : 0ull; }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint32_t q = (u + k) * RT + tid;
        if (q >= total) continue;
        const uint32_t b = ((uint32_t)rr[k]) >> 16;
        const uint4 r = run_of[b];
        if (q - r.z < r.w) __stcs((unsigned long long*)(((unsigned long long)r.y << 32) | r.x) + q, rr[k]);
        else route_direct<OP>(a, b * a.cpb + ((uint32_t)rr[k] & 0xffffu), (uint32_t)(rr[k] >> 32));   // the region is full
      }
    }
    __syncthreads();
  }
  // the last n & 3 rows: straight to the canvas
  if (blockIdx.x == 0 && tid < (int)(a.n & 3)) {
    const long long i = (n4 << 2) + tid;
    const float vv = a.vcol ? a.vcol[i] : 1.f;
    const int cell = vv == vv ? map_exact_linear(a.v, a.x[i], a.y[i]) : -1;
    if (cell >= 0) {
      if (OP == R_MAX32 || OP == R_MIN32) { negzero |= is_negzero(vv); route_direct<OP>(a, (uint32_t)cell, __float_as_uint(vv)); }
      else route_direct<OP>(a, (uint32_t)cell, (uint32_t)i);
    }
  }
  if ((OP == R_MAX32 || OP == R_MIN32) && negzero && a.notes) *a.notes = DSB_NOTE_NEGZERO;
}

// the rows pass 1 could not place with the float32 mapping: exact f64 mapping, straight to the canvas
template <int OP>
__global__ void __launch_bounds__(256) k_route_slow(const __grid_constant__ RouteArgs a) {
  if (route_gate_closed(a)) return;
  const uint32_t n = min(*a.slow_n, a.slow_cap);
  bool negzero = false;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const uint32_t* e = a.slow + 3 * (size_t)j;
    const uint32_t payload = e[2];
    const int cell = map_exact_linear(a.v, __uint_as_float(e[0]), __uint_as_float(e[1]));
    if (cell < 0) continue;
    if (OP == R_MAX32 || OP == R_MIN32) negzero |= is_negzero(__uint_as_float(payload));
    route_direct<OP>(a, (uint32_t)cell, payload);
  }
  if ((OP == R_MAX32 || OP == R_MIN32) && negzero && a.notes) *a.notes = DSB_NOTE_NEGZERO;
}

// ---- pass 2 ----------------------------------------------------------------------------------------------------------
// TMA bulk-copy helpers (cp.async.bulk + mbarrier, the 1-D form: a bucket's records are one contiguous run)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
               :: "r"(bar), "r"(parity) : "memory");
}

constexpr uint32_t RE_CHUNK = 2048;        // records per TMA chunk (16 KB): two per thread
constexpr uint32_t RE_STAGES = 3;          // chunks in flight per SM (48 KB) next to the 176 KB tile

// One CTA per SM takes buckets off a queue.  TMA = true: thread 0 streams the bucket's records through a 3-stage shared-memory
// ring with bulk copies (48 KB in flight per SM without holding registers; the LDG form has 16 KB in flight), and the first chunks of
// the NEXT bucket are already on their way while the tile is folded into the canvas.  TMA = false is the plain-load A/B arm.
template <int OP, bool TMA>
__global__ void __launch_bounds__(1024, 1) k_route_eat(const __grid_constant__ RouteArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  if (route_gate_closed(a)) return;
  uint32_t* tile = (uint32_t*)smem;
  __shared__ uint32_t next;
  __shared__ __align__(8) unsigned long long bars[RE_STAGES];
  constexpr uint32_t IDENT = OP == R_MAX32 ? (uint32_t)INT_MIN : OP == R_MIN32 ? (uint32_t)INT_MAX : OP == R_MINROW ? 0xffffffffu : 0u;
  auto eat = [&](uint32_t key, uint32_t pay) {
    const uint32_t l = key & 0xffffu;
    if (OP == R_MAX32) atomicMax((int*)tile + l, key32_from_f32(__uint_as_float(pay)));
    else if (OP == R_MIN32) atomicMin((int*)tile + l, key32_from_f32(__uint_as_float(pay)));
    else if (OP == R_MINROW) atomicMin(tile + l, pay);
    else if (OP == R_MAXROW) atomicMax(tile + l, pay + 1u);          // 0 = no row yet
    else atomicAdd(tile + l, 1u);
  };
  // ring: [RE_STAGES][RE_CHUNK] records behind the tile (the tile is padded to 128 bytes by the launcher)
  const uint32_t tile_bytes = (a.cpb * 4 + 127) & ~127u;
  const uint4* ring = (const uint4*)(smem + tile_bytes);
  const uint32_t ring_s = smem_u32(ring), bar_s = smem_u32(bars);
  uint32_t g = 0;                                                     // chunks consumed so far: stage g % 3, parity (g / 3) & 1
  auto issue = [&](uint32_t b, uint32_t c, uint32_t gi) {            // thread 0: chunk c of bucket b goes into stage gi % 3
    const uint32_t nrec = min(a.cursor[b], a.cap[b]);
    const uint32_t m = min(RE_CHUNK, nrec - c * RE_CHUNK);
    const uint32_t bytes = ((m + 1) & ~1u) * 8;                       // cap[b] is even: the odd record's neighbour is inside the region
    const uint32_t st = gi % RE_STAGES;
    mbar_expect_tx(bar_s + 8 * st, bytes);
    tma_load_1d(ring_s + st * (RE_CHUNK * 8), a.recs + a.off[b] + (size_t)c * RE_CHUNK, bytes, bar_s + 8 * st);
  };
  auto prefetch = [&](uint32_t b) {                                  // thread 0: the first chunks of a bucket
    if (b >= a.nb) return;
    const uint32_t nrec = min(a.cursor[b], a.cap[b]);
    const uint32_t nch = (nrec + RE_CHUNK - 1) / RE_CHUNK;
    for (uint32_t c = 0; c < nch && c < RE_STAGES; c++) issue(b, c, g + c);
  };
  if (TMA && threadIdx.x == 0) {
    for (uint32_t i = 0; i < RE_STAGES; i++) mbar_init(bar_s + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x == 0) { next = atomicAdd(a.queue, 1u); }
  __syncthreads();
  if (TMA && threadIdx.x == 0) prefetch(next);
  for (;;) {
    const uint32_t b = next;
    if (b >= a.nb) break;
    for (uint32_t q = threadIdx.x; q < a.cpb; q += 1024) tile[q] = IDENT;
    __syncthreads();                                                  // tile ready; everyone has read `next`
    const uint32_t nrec = min(a.cursor[b], a.cap[b]);
    if (TMA) {
      const uint32_t nch = (nrec + RE_CHUNK - 1) / RE_CHUNK;
      for (uint32_t c = 0; c < nch; c++, g++) {
        const uint32_t st = g % RE_STAGES;
        mbar_wait(bar_s + 8 * st, (g / RE_STAGES) & 1u);
        const uint32_t m = min(RE_CHUNK, nrec - c * RE_CHUNK);
        if (2 * threadIdx.x < m) {
          const uint4 q = ring[st * (RE_CHUNK / 2) + threadIdx.x];
          eat(q.x, q.y);
          if (2 * threadIdx.x + 1 < m) eat(q.z, q.w);
        }
        __syncthreads();                                              // the stage is free again
        if (threadIdx.x == 0) {
          if (c + RE_STAGES < nch) issue(b, c + RE_STAGES, g + RE_STAGES);
          else if (c + 1 == nch) { next = atomicAdd(a.queue, 1u); g++; prefetch(next); g--; }   // all stages are free here
        }
      }
      if (nch == 0) {
        if (threadIdx.x == 0) { next = atomicAdd(a.queue, 1u); prefetch(next); }
      }
    } else {
      const unsigned long long* base = a.recs + a.off[b];            // off and cap are even: 16-byte aligned
      const uint4* r4 = (const uint4*)base;
      const uint32_t n2 = nrec >> 1;
      for (uint32_t i = threadIdx.x; i < n2; i += 1024) {
        const uint4 q = __ldcs(r4 + i);
        eat(q.x, q.y);
        eat(q.z, q.w);
      }
      if ((nrec & 1) && threadIdx.x == 0) { const unsigned long long q = base[nrec - 1]; eat((uint32_t)q, (uint32_t)(q >> 32)); }
      __syncthreads();
      if (threadIdx.x == 0) next = atomicAdd(a.queue, 1u);
    }
    // fold the tile into the canvas: this CTA is the only writer of these cells now (pass 1 has finished)
    const uint32_t c0 = b * a.cpb;
    for (uint32_t q = threadIdx.x; q < a.cpb && c0 + q < a.ncell; q += 1024) {
      const uint32_t t = tile[q];
      if (t == IDENT) continue;
      if (OP == R_MAX32) { int* c = (int*)a.canvas + c0 + q; if ((int)t > *c) *c = (int)t; }
      else if (OP == R_MIN32) { int* c = (int*)a.canvas + c0 + q; if ((int)t < *c) *c = (int)t; }
      else if (OP == R_MINROW) { long long* c = (long long*)a.canvas + c0 + q; const long long r = a.row_offset + (long long)t; if (r < *c) *c = r; }
      else if (OP == R_MAXROW) { long long* c = (long long*)a.canvas + c0 + q; const long long r = a.row_offset + (long long)t - 1; if (r > *c) *c = r; }
      else ((unsigned int*)a.canvas)[c0 + q] += t;
    }
    __syncthreads();                                                  // `next` is visible; the tile may be cleared
  }
}

// ---- entry points ------------------------------------------------------------------------------------------------------
// ---- first / last: only the head (tail) of the rows can win ---------------------------------------------------------------
// first = the smallest row id of a pixel (reductions.py:2318-2324, 1346-1360).  Once a pixel holds a row of the HEAD of the call
// (its first ~10 rows per canvas cell, routed as above), no later row can replace it - and with 60 rows per pixel (config 5) that is
// every pixel but a handful.  The REST of the rows is therefore only tested against a bitmap of the pixels the head left open
// (1 bit per pixel: 8 MB at 8192^2, L2-resident): x and y are streamed (8 B/row), the row is dropped after one L2 hit, and the few
// rows of open pixels check their NaN column and vote with an atomic.  last = the largest row id: the same from the tail.
// Exact for any data; data whose head does not cover the canvas (rows sorted in space) would send most of the rest to DRAM
// atomics, so a sample of the rest decides ON THE DEVICE between this pass and the routed pass over the rest (RouteArgs::gate).
struct RestArgs {
  dsb_view v;
  FastMap fm;
  const float* x; const float* y; const float* chk;
  long long n, row_offset;          // the rest: rows [0, n) of these pointers, global row = row_offset + i
  long long* canvas;
  uint32_t* bits;                   // [ceil(ncell / 32)] 1 = the pixel is settled
  uint32_t* blocks;                 // [ceil(cw * ch / 32)] 1 = every pixel of the 2^bs x 2^bs block is settled (8 x 8 blocks, 128 KB at 8192^2:
  int cw, ch, bs;                   //   every CTA of k_rows_rest keeps a copy in shared memory)
  unsigned long long* gate;         // {sampled rows, sampled rows of open pixels}
  long long limit;                  // first: rows below it settle a pixel; last: rows at or above it
  long long ncell;
};

template <bool FIRST>
__global__ void __launch_bounds__(256) k_rows_settled(const RestArgs a) {
  const long long nw = (a.ncell + 31) >> 5;
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nw; w += (long long)gridDim.x * (blockDim.x >> 5)) {
    const long long c = (w << 5) + (threadIdx.x & 31);
    bool done = false;
    if (c < a.ncell) { const long long r = a.canvas[c]; done = FIRST ? r < a.limit : r >= a.limit; }
    const uint32_t m = __ballot_sync(0xffffffffu, done);
    if ((threadIdx.x & 31) == 0) a.bits[w] = m;
  }
}

// a 2^bs x 2^bs block is settled when all of its pixels are (pixels beyond the canvas edge count as settled); one warp per block
__global__ void __launch_bounds__(256) k_rows_settled_blocks(const RestArgs a) {
  const long long nblk = (long long)a.cw * a.ch, nbw = (nblk + 31) >> 5;
  const int lane = threadIdx.x & 31;
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nbw; w += (long long)gridDim.x * (blockDim.x >> 5)) {
    uint32_t word = 0;
    for (int j = 0; j < 32; j++) {
      const long long b = (w << 5) + j;
      bool all = true;
      if (b < nblk) {
        const int bx = (int)(b % a.cw), by = (int)(b / a.cw);
        for (int p = lane; p < (1 << (2 * a.bs)); p += 32) {
          const int px = (bx << a.bs) + (p & ((1 << a.bs) - 1)), py = (by << a.bs) + (p >> a.bs);
          if (px < a.v.width && py < a.v.height) {
            const long long c = (long long)py * a.v.width + px;
            all = all && ((a.bits[c >> 5] >> (c & 31)) & 1u);
          }
        }
      }
      if (__all_sync(0xffffffffu, all) && b < nblk) word |= 1u << j;
    }
    if (lane == 0) a.blocks[w] = word;
  }
}

// every 64th block of 1024 rows of the rest: how many land on the canvas, how many of those on a pixel that is still open
__global__ void __launch_bounds__(256) k_rows_rest_sample(const RestArgs a, long long stride_blocks) {
  const float4* __restrict__ x4 = (const float4*)a.x;
  const float4* __restrict__ y4 = (const float4*)a.y;
  const long long n4 = a.n >> 2;
  uint32_t tot = 0, open = 0;
  for (long long blk = blockIdx.x; blk * stride_blocks * 256 < n4; blk += gridDim.x) {
    const long long i4 = blk * stride_blocks * 256 + threadIdx.x;
    if (i4 >= n4) continue;
    const float4 xa = __ldg(x4 + i4), ya = __ldg(y4 + i4);
    const float xs[4] = {xa.x, xa.y, xa.z, xa.w}, ys[4] = {ya.x, ya.y, ya.z, ya.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float xf = fmaf(xs[k], a.fm.sx, a.fm.tx), yf = fmaf(ys[k], a.fm.sy, a.fm.ty);
      const int xi = __float2int_rd(xf), yi = __float2int_rd(yf);
      if ((uint32_t)xi < (uint32_t)a.v.width && (uint32_t)yi < (uint32_t)a.v.height) {        // an estimate: the fast pixel is enough
        const uint32_t cell = (uint32_t)(yi * a.v.width + xi);
        tot++;
        open += ((__ldg(a.bits + (cell >> 5)) >> (cell & 31)) & 1u) ^ 1u;
      }
    }
  }
  tot = __reduce_add_sync(0xffffffffu, tot); open = __reduce_add_sync(0xffffffffu, open);
  if ((threadIdx.x & 31) == 0 && tot) { atomicAdd(a.gate, (unsigned long long)tot); atomicAdd(a.gate + 1, (unsigned long long)open); }
}

constexpr int RQ_CAP = 64;                 // k_rows_rest: queue entries per warp (drained at 32)

template <bool FIRST, int THREADS, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS) k_rows_rest(const __grid_constant__ RestArgs a) {
  extern __shared__ uint32_t blocks[];                         // the block map: a random 4-byte read per row costs a few bank conflicts
  if (rest_wants_route(a.gate)) return;                        // the routed kernels take the rest
  for (int w = threadIdx.x; w < (a.cw * a.ch + 31) >> 5; w += blockDim.x) blocks[w] = a.blocks[w];
  __syncthreads();
  const uint32_t W = (uint32_t)a.v.width, H = (uint32_t)a.v.height;
  const FastMap& fm = a.fm;
  const uint32_t* __restrict__ bits = a.bits;
  const int bs = a.bs, cw = a.cw;
  // The loop of k_points_priv_tight: two vectors (8 rows) per thread per step, every row finished as it is mapped.  The instruction
  // count of this kernel is its bound (ncu: 65 % of the issue slots busy), and the rows that need a closer look - the fast pixel is
  // not certain (0.8 % at 8192^2), or its block still holds an open pixel (0.3 %) - cost a whole warp each time ONE lane has one when
  // they are handled in place (23 % of the instructions executed).  They go to a per-warp queue in shared memory instead ({x, y, row})
  // and are looked at 32 at a time, every lane busy: the exact mapping, the pixel's own bit (L2), the NaN check, the vote.
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* q = blocks + ((a.cw * a.ch + 31) >> 5) + warp * 3 * RQ_CAP;          // x[RQ_CAP], y[RQ_CAP], row[RQ_CAP] (row within the rest)
  int qn = 0;                                                                      // warp-uniform
  auto closer = [&](float xv, float yv, long long i) {
    const int cell = map_exact_linear(a.v, xv, yv);
    if (cell < 0 || ((__ldg(bits + (cell >> 5)) >> (cell & 31)) & 1u)) return;
    const float c = a.chk[i];
    if (c != c) return;
    if (FIRST) atomicMin(a.canvas + cell, a.row_offset + i); else atomicMax(a.canvas + cell, a.row_offset + i);
  };
  auto drain32 = [&]() {                                                 // the first 32 entries, one per lane; the rest moves to the front
    __syncwarp();
    closer(__uint_as_float(q[lane]), __uint_as_float(q[RQ_CAP + lane]), (long long)q[2 * RQ_CAP + lane]);
    const int rest = qn - 32;
    uint32_t c = 0, k = 0, r = 0;
    if (lane < rest) { c = q[32 + lane]; k = q[RQ_CAP + 32 + lane]; r = q[2 * RQ_CAP + 32 + lane]; }
    __syncwarp();
    if (lane < rest) { q[lane] = c; q[RQ_CAP + lane] = k; q[2 * RQ_CAP + lane] = r; }
    qn = rest;
  };
  auto one = [&](float xv, float yv) -> uint32_t {            // 1 = this row needs a closer look
    const float xf = fmaf(xv, fm.sx, fm.tx), yf = fmaf(yv, fm.sy, fm.ty);
    const int xi = __float2int_rd(xf), yi = __float2int_rd(yf);
    const float dx = xf - (float)xi, dy = yf - (float)yi;
    const bool sure = dx >= fm.ex && dx <= fm.omex && dy >= fm.ey && dy <= fm.omey;
    const bool inside = (uint32_t)xi < W && (uint32_t)yi < H;
    const int b = inside ? (yi >> bs) * cw + (xi >> bs) : 0;
    const uint32_t open = ((blocks[b >> 5] >> (b & 31)) & 1u) ^ 1u;
    return sure ? (uint32_t)inside & open : 1u;
  };
  auto push = [&](bool look, float xv, float yv, uint32_t row) {
    const unsigned bal = __ballot_sync(0xffffffffu, look);
    if (bal) {
      if (look) {
        const int pos = qn + __popc(bal & ((1u << lane) - 1u));
        q[pos] = __float_as_uint(xv); q[RQ_CAP + pos] = __float_as_uint(yv); q[2 * RQ_CAP + pos] = row;
      }
      qn += __popc(bal);
      if (qn >= 32) drain32();
    }
  };
  const float4* __restrict__ x4 = (const float4*)a.x;
  const float4* __restrict__ y4 = (const float4*)a.y;
  const long long n4 = a.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // warp-uniform trip count (the ballots need every lane): whole steps only; what is left takes the closer look directly
  long long w4 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
  for (; w4 + stride + 31 < n4; w4 += 2 * stride) {
    const long long i4 = w4 + lane;
    const float4 xa = __ldcs(x4 + i4), ya = __ldcs(y4 + i4);
    const float4 xb = __ldcs(x4 + i4 + stride), yb = __ldcs(y4 + i4 + stride);
    const uint32_t m = one(xa.x, ya.x) | one(xa.y, ya.y) << 1 | one(xa.z, ya.z) << 2 | one(xa.w, ya.w) << 3 |
                       one(xb.x, yb.x) << 4 | one(xb.y, yb.y) << 5 | one(xb.z, yb.z) << 6 | one(xb.w, yb.w) << 7;
    if (__any_sync(0xffffffffu, m != 0)) {         // one vote per step on small canvases (0.07 % of the rows), the eight ballots only then
      const uint32_t ra = (uint32_t)(4 * i4), rb = (uint32_t)(4 * (i4 + stride));
      push(m & 1, xa.x, ya.x, ra); push(m & 2, xa.y, ya.y, ra + 1); push(m & 4, xa.z, ya.z, ra + 2); push(m & 8, xa.w, ya.w, ra + 3);
      push(m & 16, xb.x, yb.x, rb); push(m & 32, xb.y, yb.y, rb + 1); push(m & 64, xb.z, yb.z, rb + 2); push(m & 128, xb.w, yb.w, rb + 3);
    }
  }
  __syncwarp();
  if (lane < qn) closer(__uint_as_float(q[lane]), __uint_as_float(q[RQ_CAP + lane]), (long long)q[2 * RQ_CAP + lane]);
  for (long long i4 = w4 + lane; i4 < n4; i4 += stride) {                                   // the last, partial steps
    const float4 xa = __ldcs(x4 + i4), ya = __ldcs(y4 + i4);
    closer(xa.x, ya.x, 4 * i4); closer(xa.y, ya.y, 4 * i4 + 1); closer(xa.z, ya.z, 4 * i4 + 2); closer(xa.w, ya.w, 4 * i4 + 3);
  }
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {           // tail rows
    const long long i = (n4 << 2) + threadIdx.x;
    closer(a.x[i], a.y[i], i);
  }
}

static long long g_routed_min_rows = 1LL << 24;
static long long g_head_per_cell = 10;          // first / last: rows per canvas cell that are routed before the rest is only filtered (0 = off);
                                                // 6 / 8 / 10 / 12 / 16 at 8192^2, 4e9 rows: 25.7 / 19.5 / 17.7 / 17.5 / 19.0 ms (tools/bench_first.py)
void dsb_routed_set_head(long long rows_per_cell) { g_head_per_cell = rows_per_cell; }
static bool g_routed_tma = true;
void dsb_routed_set_tma(bool on) { g_routed_tma = on; }

extern "C" int dsb_routed_configure(int64_t min_rows) { g_routed_min_rows = min_rows; return DSB_OK; }

static uint32_t route_nb(long long ncell, uint32_t* cpb) {
  const long long nb = (ncell + R_CPB_MAX - 1) / R_CPB_MAX;
  *cpb = (uint32_t)((ncell + nb - 1) / nb);
  return (uint32_t)nb;
}
static size_t route_header_bytes(uint32_t nb) { return (((size_t)nb * (8 + 4 + 4 + 4) + 64) + 255) & ~(size_t)255; }
static size_t route_slow_entries(int64_t n) { return (size_t)(n / 32) + 65536; }        // ~3 % of the rows (0.8 % at 8192^2, less on smaller canvases)
static size_t route_fixed_bytes(uint32_t nb, int64_t n) { return route_header_bytes(nb) + ((route_slow_entries(n) * 12 + 255) & ~(size_t)255); }

static size_t route_rest_cell_bits_bytes(long long ncell) { return (size_t)((((ncell + 31) >> 5) * 4 + 255) & ~255LL); }
// blocks of 8 x 8 pixels, or the smallest power of two whose bitmap fits REST_MAP_BYTES of shared memory (one CTA per SM)
constexpr long long REST_MAP_BYTES = 160 * 1024;
static int route_rest_block_shift(long long W, long long H) {
  int bs = 3;
  while ((((W + (1LL << bs) - 1) >> bs) * ((H + (1LL << bs) - 1) >> bs)) > REST_MAP_BYTES * 8) bs++;
  return bs;
}
// gate counters (256 B) + a bit per pixel + a bit per block
static size_t route_rest_bytes(long long W, long long H) {
  return 256 + route_rest_cell_bits_bytes(W * H) + (size_t)REST_MAP_BYTES + 256;
}

extern "C" int64_t dsb_points_routed_scratch_bytes(const dsb_view* view, int64_t n) {
  if (!view || view->width <= 0 || view->height <= 0 || n < 0) return 0;
  uint32_t cpb;
  const uint32_t nb = route_nb((long long)view->width * view->height, &cpb);
  // records: 1.10 n + 4096 per bucket (k_route_plan scales the capacities down to whatever it is given)
  // + the floor dsb_points_routed asks for, + first / last: a bit per pixel and the gate counters
  return (int64_t)(route_fixed_bytes(nb, n) + ((size_t)((double)n * 1.10) + (size_t)nb * 4096 + 1024) * 8 + route_rest_bytes(view->width, view->height));
}

template <int OP>
static void route_launch(const RouteArgs& a, size_t smem1, size_t smem2, cudaStream_t s) {
  cudaFuncSetAttribute(k_route_bin<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
  const size_t smem2t = ((smem2 + 127) & ~(size_t)127) + (size_t)RE_STAGES * RE_CHUNK * 8;
  cudaFuncSetAttribute(k_route_eat<OP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2t);
  cudaFuncSetAttribute(k_route_eat<OP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
  k_route_bin<OP><<<dsb_num_sms() * 2, RT, smem1, s>>>(a);
  k_route_slow<OP><<<dsb_num_sms() * 2, 256, 0, s>>>(a);
  if (g_routed_tma) k_route_eat<OP, true><<<dsb_num_sms(), 1024, smem2t, s>>>(a);
  else k_route_eat<OP, false><<<dsb_num_sms(), 1024, smem2, s>>>(a);
}

// one routed aggregation (sample, plan, bin, slow, eat) of rows [0, n); gate != NULL: the device decides whether it runs
static int route_one(const dsb_view* view, const void* x, const void* y, int32_t xy_dtype, int64_t n, int64_t row_offset,
                     const dsb_plan* plan, void* scratch, int64_t scratch_bytes, void* stream, const unsigned long long* gate) {
  if (!view || view->width <= 0 || view->height <= 0) { dsb_set_error("dsb_points_routed: bad view"); return DSB_ERR_ARG; }
  if (!plan || plan->nops != 1 || plan->ncat != 0) { dsb_set_error("dsb_points_routed: one accumulator, no categories"); return DSB_ERR_UNSUPPORTED; }
  if (n <= 0 || n >= (1LL << 32) - 1) { dsb_set_error("dsb_points_routed: n out of range"); return DSB_ERR_UNSUPPORTED; }
  if (!x || !y || !scratch) { dsb_set_error("dsb_points_routed: null pointer"); return DSB_ERR_ARG; }
  const dsb_base& b = plan->ops[0];
  if (!b.agg) { dsb_set_error("dsb_points_routed: op has no canvas"); return DSB_ERR_ARG; }
  int op = -1;
  const float* vcol = nullptr;
  switch (b.op) {
    case DSB_OP_MAX32: case DSB_OP_MIN32:
      if (b.val_dtype != DSB_F32 || !b.val || b.chk_dtype != DSB_NONE) break;
      vcol = (const float*)b.val; op = b.op == DSB_OP_MAX32 ? R_MAX32 : R_MIN32; break;
    case DSB_OP_MINROW: case DSB_OP_MAXROW:
      if (b.chk_dtype != DSB_F32 || !b.chk || b.val_dtype != DSB_NONE) break;
      vcol = (const float*)b.chk; op = b.op == DSB_OP_MINROW ? R_MINROW : R_MAXROW; break;
    case DSB_OP_COUNT:
      if (b.chk_dtype != DSB_NONE || (b.val_dtype != DSB_NONE && b.val_dtype != DSB_F32)) break;
      vcol = b.val_dtype == DSB_F32 ? (const float*)b.val : nullptr; op = R_COUNT; break;
    default: break;
  }
  if (op < 0) { dsb_set_error("dsb_points_routed: serves max / min / first / last of a float32 column and count"); return DSB_ERR_UNSUPPORTED; }
  const long long ncell = (long long)view->width * view->height;
  if (xy_dtype != DSB_F32 || ncell >= (1LL << 31) || ((((uintptr_t)x | (uintptr_t)y | (uintptr_t)vcol) & 15) != 0)) {
    dsb_set_error("dsb_points_routed: needs aligned float32 columns and fewer than 2^31 cells"); return DSB_ERR_UNSUPPORTED;
  }
  RouteArgs a;
  a.v = *view;
  a.fm = make_fast_map(view);
  if (!a.fm.enabled) { dsb_set_error("dsb_points_routed: needs linear axes within the float32 fast mapping's error bound"); return DSB_ERR_UNSUPPORTED; }
  a.x = (const float*)x; a.y = (const float*)y; a.vcol = vcol; a.n = n; a.row_offset = row_offset;
  a.nb = route_nb(ncell, &a.cpb);
  {                                           // the exact division by cpb of route_key
    uint32_t lg = 0;
    while ((2u << lg) <= a.cpb) lg++;         // floor(log2 cpb)
    if ((1u << lg) == a.cpb && lg > 0) lg--;  // a power of two: keep m below 2^32
    const unsigned __int128 one = (unsigned __int128)1 << (32 + lg);
    const unsigned __int128 m = (one + a.cpb - 1) / a.cpb;
    const unsigned __int128 e = m * a.cpb - one;
    if (m >> 32 || (unsigned __int128)ncell * e >= one) { dsb_set_error("dsb_points_routed: no exact bucket divisor for this canvas"); return DSB_ERR_UNSUPPORTED; }
    a.inv = (uint32_t)m; a.sh = lg; a.kmul = 65536u - a.cpb;
  }
  a.ncell = (uint32_t)ncell;
  // pass 1 keeps a 20-byte entry per bucket next to its 64 KB tile: canvases beyond ~8000 buckets (360 M cells) stay with the banded kernels
  if ((size_t)RTILE * 8 + (size_t)((a.nb + 2) & ~1u) * (16 + 4) + 64 * 4 > 226 * 1024) {
    dsb_set_error("dsb_points_routed: too many buckets for the shared-memory tables"); return DSB_ERR_UNSUPPORTED;
  }
  // the list of rows for k_route_slow shrinks to what a smaller scratch leaves (the caller sized it for fewer rows: first / last on an
  // L2-resident canvas, whose rest is routed only if the device-side sample says so); a full list maps its rows in place
  size_t slow_entries = route_slow_entries(n);
  {
    const size_t fixed = route_header_bytes(a.nb) + ((size_t)a.nb * 4096 + 1024) * 8 + 4096;
    if ((size_t)scratch_bytes > fixed) {
      const size_t room = ((size_t)scratch_bytes - fixed) / 4 / 12;
      if (slow_entries > room) slow_entries = room > 1024 ? room : 1024;
    }
  }
  const size_t hdr = route_header_bytes(a.nb) + ((slow_entries * 12 + 255) & ~(size_t)255);
  if (a.nb > 65535 || scratch_bytes < (int64_t)(hdr + ((size_t)a.nb * 4096 + 1024) * 8)) { dsb_set_error("dsb_points_routed: scratch too small"); return DSB_ERR_ARG; }
  unsigned char* p = (unsigned char*)scratch;
  unsigned long long* off = (unsigned long long*)p;
  uint32_t* cap = (uint32_t*)(off + a.nb);
  uint32_t* cursor = cap + a.nb;
  uint32_t* hist = cursor + a.nb;
  uint32_t* queue = hist + a.nb;
  a.slow_n = queue + 1;
  a.slow = (uint32_t*)(p + route_header_bytes(a.nb));
  a.slow_cap = (uint32_t)slow_entries;
  a.recs = (unsigned long long*)(p + hdr);
  a.off = off; a.cap = cap; a.cursor = cursor; a.queue = queue; a.canvas = b.agg; a.notes = plan->notes; a.gate = gate;
  const unsigned long long capacity = ((unsigned long long)scratch_bytes - hdr) / 8 - 4096;   // k_route_plan rounds every thread's start up
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(hist, 0, (size_t)a.nb * 4, s);
  const long long stride_blocks = n >= (1LL << 28) ? 64 : 16;   // every 64th (16th) block of 1024 points is sampled
  k_route_sample<<<dsb_num_sms() * 4, 256, (size_t)a.nb * 4, s>>>(a, stride_blocks, hist);
  k_route_plan<<<1, 1024, 0, s>>>(hist, a.nb, (uint32_t)stride_blocks, capacity, off, cap, cursor, queue, a.slow_n);
  const uint32_t nbp = (a.nb + 2) & ~1u;
  const size_t smem1 = (size_t)RTILE * 8 + (size_t)nbp * (16 + 4) + 64 * 4;
  const size_t smem2 = (size_t)a.cpb * 4;
  static const char* const names[] = {"max32", "min32", "minrow", "maxrow", "count"};
  dsb_note_kernel("k_route_bin<%s> + k_route_eat<%s> buckets=%u", names[op], names[op], a.nb);
  switch (op) {
    case R_MAX32: route_launch<R_MAX32>(a, smem1, smem2, s); break;
    case R_MIN32: route_launch<R_MIN32>(a, smem1, smem2, s); break;
    case R_MINROW: route_launch<R_MINROW>(a, smem1, smem2, s); break;
    case R_MAXROW: route_launch<R_MAXROW>(a, smem1, smem2, s); break;
    default: route_launch<R_COUNT>(a, smem1, smem2, s); break;
  }
  DSB_CUDA_CHECK_LAUNCH("dsb_points_routed");
  return DSB_OK;
}

extern "C" int dsb_points_routed(const dsb_view* view, const void* x, const void* y, int32_t xy_dtype, int64_t n, int64_t row_offset,
                                 const dsb_plan* plan, void* scratch, int64_t scratch_bytes, void* stream) {
  if (n < g_routed_min_rows) { dsb_set_error("dsb_points_routed: n out of range"); return DSB_ERR_UNSUPPORTED; }
  const bool rowop = plan && plan->nops == 1 && (plan->ops[0].op == DSB_OP_MINROW || plan->ops[0].op == DSB_OP_MAXROW);
  const long long ncell = view ? (long long)view->width * view->height : 0;
  const long long head = (g_head_per_cell * ncell + 3) & ~3LL;
  const size_t rest_bytes = view ? route_rest_bytes(view->width, view->height) : 0;
  if (!rowop || g_head_per_cell <= 0 || head < 4 || n < head + head / 4 ||      // worth it from 1.25 x the head on (the rest costs a quarter of the routed pass per row)
      !scratch || scratch_bytes < (int64_t)rest_bytes + (1 << 20) ||
      plan->ops[0].chk_dtype != DSB_F32 || !plan->ops[0].chk || xy_dtype != DSB_F32 || !x || !y)
    return route_one(view, x, y, xy_dtype, n, row_offset, plan, scratch, scratch_bytes, stream, nullptr);

  // first / last with many rows per pixel: route the head (tail) of the rows, only filter the rest (see k_rows_rest)
  const bool first = plan->ops[0].op == DSB_OP_MINROW;
  const long long n_rest = first ? n - head : (n - head) & ~3LL;          // both parts start 16-byte aligned
  const long long n_head = n - n_rest;
  const long long off_head = first ? 0 : n_rest, off_rest = first ? n_head : 0;
  const int64_t routed_bytes = (scratch_bytes - (int64_t)rest_bytes) & ~(int64_t)255;
  unsigned char* tailp = (unsigned char*)scratch + routed_bytes;
  const float* xf = (const float*)x; const float* yf = (const float*)y; const float* cf = (const float*)plan->ops[0].chk;
  dsb_plan ph = *plan, pr = *plan;
  ph.ops[0].chk = cf + off_head; pr.ops[0].chk = cf + off_rest;
  int rc = route_one(view, xf + off_head, yf + off_head, xy_dtype, n_head, row_offset + off_head, &ph, scratch, routed_bytes, stream, nullptr);
  if (rc != DSB_OK) return rc;                          // nothing was launched: the caller takes the banded kernels
  RestArgs r;
  r.v = *view; r.fm = make_fast_map(view);
  r.x = xf + off_rest; r.y = yf + off_rest; r.chk = cf + off_rest; r.n = n_rest; r.row_offset = row_offset + off_rest;
  r.canvas = (long long*)plan->ops[0].agg; r.ncell = ncell;
  r.gate = (unsigned long long*)tailp; r.bits = (uint32_t*)(tailp + 256);
  r.blocks = (uint32_t*)(tailp + 256 + route_rest_cell_bits_bytes(ncell));
  r.bs = route_rest_block_shift(view->width, view->height);
  r.cw = (int)((view->width + (1LL << r.bs) - 1) >> r.bs); r.ch = (int)((view->height + (1LL << r.bs) - 1) >> r.bs);
  const size_t rest_smem = (size_t)(((long long)r.cw * r.ch + 31) >> 5) * 4 + 32 * 3 * RQ_CAP * 4;
  r.limit = first ? row_offset + n_head : row_offset + off_head;
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(r.gate, 0, 16, s);
  const long long stride_blocks = n_rest >= (1LL << 28) ? 64 : 16;
  if (first) k_rows_settled<true><<<dsb_num_sms() * 8, 256, 0, s>>>(r); else k_rows_settled<false><<<dsb_num_sms() * 8, 256, 0, s>>>(r);
  k_rows_settled_blocks<<<dsb_num_sms() * 8, 256, 0, s>>>(r);
  k_rows_rest_sample<<<dsb_num_sms() * 4, 256, 0, s>>>(r, stride_blocks);
  cudaFuncSetAttribute(k_rows_rest<true, 1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(REST_MAP_BYTES + 32 * 3 * RQ_CAP * 4));
  cudaFuncSetAttribute(k_rows_rest<false, 1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(REST_MAP_BYTES + 32 * 3 * RQ_CAP * 4));
  if (first) k_rows_rest<true, 1024, 1><<<dsb_num_sms(), 1024, rest_smem, s>>>(r); else k_rows_rest<false, 1024, 1><<<dsb_num_sms(), 1024, rest_smem, s>>>(r);
  rc = route_one(view, r.x, r.y, xy_dtype, n_rest, r.row_offset, &pr, scratch, routed_bytes, stream, r.gate);   // the gated fallback
  if (rc != DSB_OK) return rc;
  dsb_note_kernel("k_rows_rest<%s> after k_route_bin + k_route_eat<%s> of %lld head rows", first ? "first" : "last", first ? "minrow" : "maxrow", n_head);
  DSB_CUDA_CHECK_LAUNCH("dsb_points_routed(rest)");
  return DSB_OK;
}
