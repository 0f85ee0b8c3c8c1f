// Micro-benchmark 6: routed design, second iteration of the BIN pass (ubench5 measured bin 165-183 Gpts/s: the load phase,
// the shared-memory sort phases and the write-out ran one after the other with nothing in flight in between).
//
//   * x / y tiles arrive through a 2-stage TMA ring (cp.async.bulk + mbarrier): the loads of tile k+2 are in flight
//     during the whole of tile k+1; the value column is prefetched into registers at the top of the tile
//   * block-wide scan (2 entries per thread) instead of one warp walking all bands
//   * every CTA appends to its OWN sub-region of each bucket (cursor in shared memory): no global atomics
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o ubench6 ubench6.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <climits>
#include <vector>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__global__ void gen_uniform(float* x, float* y, float* v, size_t n, uint32_t seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    uint32_t a = hash32((uint32_t)i * 2654435761U + seed), b = hash32(a ^ 0x9e3779b9U), c = hash32(b ^ 0x85ebca6bU);
    x[i] = (a >> 8) * (1.0f / 16777216.0f);
    y[i] = (b >> 8) * (1.0f / 16777216.0f);
    v[i] = ((c >> 8) * (1.0f / 16777216.0f) - 0.5f) * 8.0f;
  }
}

struct Map { float sx, tx, sy, ty; uint32_t W, H; };
struct Route { uint32_t cpb, inv, nb; };

__device__ __forceinline__ int cell_of(const Map& m, float x, float y) {
  const float xf = fmaf(x, m.sx, m.tx), yf = fmaf(y, m.sy, m.ty);
  const int xi = __float2int_rd(xf), yi = __float2int_rd(yf);
  return ((uint32_t)xi < m.W && (uint32_t)yi < m.H) ? yi * (int)m.W + xi : -1;
}
__device__ __forceinline__ uint32_t key_of(const Route& r, uint32_t cell) {
  uint32_t b = __umulhi(cell, r.inv);
  uint32_t l = cell - b * r.cpb;
  if (l >= r.cpb) { b++; l -= r.cpb; }
  return (b << 16) | l;
}
__device__ __forceinline__ int key32(float f) { int b = __float_as_int(f + 0.0f); return b ^ ((b >> 31) & 0x7fffffff); }

// ---- mbarrier / TMA bulk helpers ------------------------------------------------------------------------------------
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

// ---- the bin pass ---------------------------------------------------------------------------------------------------
constexpr int BT = 1024;          // threads per CTA; PPT = points per thread per tile (template), TILE = BT * PPT

struct BinSmem {                  // offsets into dynamic shared memory
  float* in;                      // x stage 0, x stage 1, y stage 0, y stage 1 (TILE floats each)
  unsigned long long* rec;
  uint32_t *hist, *base, *gdel, *scur, *wsum;
  uint32_t bar0;                  // mbarrier of stage 0; stage 1 follows at + 8
};
__host__ __device__ inline size_t bin_smem_bytes(uint32_t nb, int TILE) {
  const uint32_t nbp = (nb + 1) & ~1u;
  return 4 * (size_t)TILE * 4 + (size_t)TILE * 8 + (size_t)nbp * 16 + 64 * 4 + 32;
}
__device__ __forceinline__ BinSmem carve(unsigned char* smem, uint32_t nb, int TILE) {
  BinSmem s;
  const uint32_t nbp = (nb + 1) & ~1u;
  float* f = (float*)smem;
  s.in = f;
  s.rec = (unsigned long long*)(f + 4 * TILE);
  s.hist = (uint32_t*)(s.rec + TILE);
  s.base = s.hist + nbp; s.gdel = s.base + nbp; s.scur = s.gdel + nbp; s.wsum = s.scur + nbp;
  unsigned long long* bars = (unsigned long long*)(s.wsum + 64);
  s.bar0 = smem_u32(bars);
  return s;
}

struct BinJob {
  const float* x; const float* y; const float* v;
  long long n;                    // points (multiple of 4)
  Map m; Route r;
  unsigned long long* out;        // records: sub-region (b, cta) starts at ((size_t)b * nctas + cta) * cap_sub
  uint32_t cap_sub;
  uint32_t* overflow;
};

// issue the TMA loads of the tile that starts at point t0 into stage `st`
__device__ __forceinline__ void issue_tile(const BinJob& j, const BinSmem& s, int st, long long t0, int TILE) {
  const long long left = j.n - t0;
  const uint32_t pts = (uint32_t)(left < TILE ? left : TILE);
  const uint32_t bar = s.bar0 + 8 * st;
  mbar_expect_tx(bar, pts * 8);
  tma_load_1d(smem_u32(s.in + st * TILE), j.x + t0, pts * 4, bar);
  tma_load_1d(smem_u32(s.in + (2 + st) * TILE), j.y + t0, pts * 4, bar);
}

// one tile whose x / y are (about to be) in stage `st`; next_t0 >= 0: prefetch that tile into the same stage once the stage is free
// the value column of one tile, straight into registers (issued one tile ahead of its use)
template <int PPT>
__device__ __forceinline__ void load_values(const BinJob& j, long long t0, float4 (&va)[PPT / 4]) {
  constexpr int TILE = BT * PPT;
  const long long left = j.n - t0;
  const uint32_t pts = t0 < 0 ? 0u : (uint32_t)(left < TILE ? left : TILE);
#pragma unroll
  for (int u = 0; u < PPT / 4; u++) {
    const uint32_t p = (u * BT + threadIdx.x) * 4;
    va[u] = p < pts ? __ldcs((const float4*)(j.v + t0) + u * BT + threadIdx.x) : make_float4(NAN, NAN, NAN, NAN);
  }
}

// va: this tile's values on entry, the values of the tile after it (vnext_t0, -1 = none) on exit
template <int PPT>
__device__ __forceinline__ void bin_tile2(const BinJob& j, const BinSmem& s, int st, uint32_t parity, long long t0, long long next_t0,
                                          long long vnext_t0, float4 (&va)[PPT / 4],
                                          uint32_t region0 /* = cta * cap_sub */, uint32_t region_stride /* = nctas * cap_sub */) {
  constexpr int TILE = BT * PPT;
  const int tid = threadIdx.x;
  const long long left = j.n - t0;
  const uint32_t pts = (uint32_t)(left < TILE ? left : TILE);
  mbar_wait(s.bar0 + 8 * st, parity);
  const float4* sx4 = (const float4*)(s.in + st * TILE);
  const float4* sy4 = (const float4*)(s.in + (2 + st) * TILE);
  uint32_t key[PPT], rank[PPT];
  float val[PPT];
#pragma unroll
  for (int u = 0; u < PPT / 4; u++) {
    const uint32_t p = (u * BT + tid) * 4;
    const float4 xa = sx4[u * BT + tid], ya = sy4[u * BT + tid];
    const float xs[4] = {xa.x, xa.y, xa.z, xa.w}, ys[4] = {ya.x, ya.y, ya.z, ya.w}, vs[4] = {va[u].x, va[u].y, va[u].z, va[u].w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int c = cell_of(j.m, xs[k], ys[k]);
      const bool ok = p < pts && c >= 0 && vs[k] == vs[k];
      const uint32_t kk = ok ? key_of(j.r, (uint32_t)c) : 0xffffffffu;
      key[u * 4 + k] = kk;
      val[u * 4 + k] = vs[k];
      rank[u * 4 + k] = ok ? atomicAdd(s.hist + (kk >> 16), 1u) : 0u;
    }
  }
  __syncthreads();                                            // B1: hist complete, the stage has been read
  if (tid == 0 && next_t0 >= 0) issue_tile(j, s, st, next_t0, TILE);
  load_values<PPT>(j, vnext_t0, va);                          // lands while this tile is sorted and written out
  // block-wide exclusive scan of hist (2 entries per thread)
  const uint32_t nb = j.r.nb;
  const uint32_t e0 = 2 * tid, e1 = 2 * tid + 1;
  const uint32_t h0 = e0 < nb ? s.hist[e0] : 0u, h1 = e1 < nb ? s.hist[e1] : 0u;
  const uint32_t mine = h0 + h1;
  uint32_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += t; }
  if ((tid & 31) == 31) s.wsum[tid >> 5] = incl;
  __syncthreads();                                            // B2
  if (tid < 32) {
    const uint32_t w = s.wsum[tid];
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (tid >= o) wi += t; }
    s.wsum[32 + tid] = wi - w;
  }
  __syncthreads();                                            // B3
  const uint32_t excl = s.wsum[32 + (tid >> 5)] + incl - mine;
  if (e0 < nb) {
    const uint32_t c0 = s.scur[e0];
    if (c0 + h0 > j.cap_sub) *j.overflow = 1u;
    s.base[e0] = excl; s.gdel[e0] = e0 * region_stride + region0 + c0 - excl; s.scur[e0] = c0 + h0; s.hist[e0] = 0;
  }
  if (e1 < nb) {
    const uint32_t c1 = s.scur[e1];
    if (c1 + h1 > j.cap_sub) *j.overflow = 1u;
    s.base[e1] = excl + h0; s.gdel[e1] = e1 * region_stride + region0 + c1 - (excl + h0); s.scur[e1] = c1 + h1; s.hist[e1] = 0;
  }
  if (tid == BT - 1) s.wsum[31] = excl + mine;                // total records of the tile (wsum[31] is free after B3)
  __syncthreads();                                            // B4
#pragma unroll
  for (int k = 0; k < PPT; k++) {
    if (key[k] != 0xffffffffu)
      s.rec[s.base[key[k] >> 16] + rank[k]] = ((unsigned long long)__float_as_uint(val[k]) << 32) | key[k];
  }
  __syncthreads();                                            // B5
  const uint32_t total = s.wsum[31];
#pragma unroll
  for (int u = 0; u < PPT; u += 4) {                          // four records in flight per thread
    unsigned long long rr[4];
    uint32_t gd[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { const uint32_t q = (u + k) * BT + tid; rr[k] = q < total ? s.rec[q] : 0ull; }
#pragma unroll
    for (int k = 0; k < 4; k++) gd[k] = s.gdel[((uint32_t)rr[k]) >> 16];
#pragma unroll
    for (int k = 0; k < 4; k++) { const uint32_t q = (u + k) * BT + tid; if (q < total) j.out[(size_t)gd[k] + q] = rr[k]; }
  }
}

// stand-alone bin pass (records to DRAM), tiles cta, cta + grid, ...
template <int PPT>
__global__ void __launch_bounds__(BT, 1) k_bin2(const BinJob j, uint32_t* __restrict__ cnt) {
  constexpr int TILE = BT * PPT;
  extern __shared__ __align__(128) unsigned char smem[];
  const BinSmem s = carve(smem, j.r.nb, TILE);
  const int tid = threadIdx.x;
  for (uint32_t b = tid; b < ((j.r.nb + 1) & ~1u); b += BT) { s.hist[b] = 0; s.scur[b] = 0; }
  const long long stride = (long long)gridDim.x * TILE;
  long long t0 = (long long)blockIdx.x * TILE;
  if (tid == 0) {
    mbar_init(s.bar0, 1); mbar_init(s.bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    if (t0 < j.n) issue_tile(j, s, 0, t0, TILE);
    if (t0 + stride < j.n) issue_tile(j, s, 1, t0 + stride, TILE);
  }
  const uint32_t region0 = blockIdx.x * j.cap_sub, region_stride = gridDim.x * j.cap_sub;
  float4 va[PPT / 4];
  load_values<PPT>(j, t0 < j.n ? t0 : -1, va);
  for (uint32_t k = 0; t0 < j.n; k++, t0 += stride) {
    const long long nx = t0 + 2 * stride, nv = t0 + stride;
    bin_tile2<PPT>(j, s, k & 1, (k >> 1) & 1, t0, nx < j.n ? nx : -1, nv < j.n ? nv : -1, va, region0, region_stride);
  }
  __syncthreads();
  for (uint32_t b = tid; b < j.r.nb; b += BT) cnt[(size_t)b * gridDim.x + blockIdx.x] = s.scur[b];
}

// pass 2 (max): bucket b = records of nsub sub-regions; u32 key tile in shared memory
__global__ void __launch_bounds__(1024) k_eat_max2(const unsigned long long* __restrict__ recs, uint32_t cap_sub, uint32_t nsub,
                                                   const uint32_t* __restrict__ cnt, Route r, uint32_t ncell, int* __restrict__ canvas) {
  extern __shared__ __align__(16) unsigned char smem[];
  int* tile = (int*)smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t b = blockIdx.x; b < r.nb; b += gridDim.x) {
    for (uint32_t q = threadIdx.x; q < r.cpb; q += 1024) tile[q] = INT_MIN;
    __syncthreads();
    for (uint32_t sreg = warp; sreg < nsub; sreg += 32) {
      const uint32_t nrec = min(cnt[(size_t)b * nsub + sreg], cap_sub);
      const unsigned long long* base = recs + ((size_t)b * nsub + sreg) * cap_sub;
      const uint4* r4 = (const uint4*)base;
      const uint32_t n2 = nrec >> 1;
      for (uint32_t i = lane; i < n2; i += 32) {
        const uint4 q = __ldcs(r4 + i);
        atomicMax(tile + (q.x & 0xffffu), key32(__uint_as_float(q.y)));
        atomicMax(tile + (q.z & 0xffffu), key32(__uint_as_float(q.w)));
      }
      if ((nrec & 1) && lane == 0) {
        const unsigned long long q = base[nrec - 1];
        atomicMax(tile + ((uint32_t)q & 0xffffu), key32(__uint_as_float((uint32_t)(q >> 32))));
      }
    }
    __syncthreads();
    const uint32_t c0 = b * r.cpb;
    for (uint32_t q = threadIdx.x; q < r.cpb && c0 + q < ncell; q += 1024) canvas[c0 + q] = tile[q];
    __syncthreads();
  }
}

// ---- fused mean: persistent cooperative kernel, one band per CTA, L2-resident double-buffered record chunks ----------
// mode: 3 = both passes, 1 = bin only, 2 = eat only
template <int PPT>
__global__ void __launch_bounds__(BT, 1) k_routed_mean2(BinJob j, int tiles_per_cta, unsigned long long* buf0, unsigned long long* buf1,
                                                        uint32_t* cnt0, uint32_t* cnt1, double* __restrict__ sum_canvas,
                                                        uint32_t* __restrict__ cnt_canvas, int mode) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int TILE = BT * PPT;
  cg::grid_group grid = cg::this_grid();
  const BinSmem s = carve(smem, j.r.nb, TILE);
  double* s_sum = (double*)(smem + bin_smem_bytes(j.r.nb, TILE));
  uint32_t* s_cnt = (uint32_t*)(s_sum + j.r.cpb);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t G = gridDim.x;
  for (uint32_t q = tid; q < j.r.cpb; q += BT) { s_sum[q] = 0.0; s_cnt[q] = 0; }
  for (uint32_t b = tid; b < ((j.r.nb + 1) & ~1u); b += BT) { s.hist[b] = 0; s.scur[b] = 0; }
  if (tid == 0) {
    mbar_init(s.bar0, 1); mbar_init(s.bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // tile k of this CTA (k = ph * tiles_per_cta + t) starts at point ((ph * tiles_per_cta + t) * G + cta) * TILE
  auto tile_start = [&](long long k) { return (k * G + blockIdx.x) * (long long)TILE; };
  const long long ntiles_all = (j.n + TILE - 1) / TILE;
  const long long my_tiles = (ntiles_all > blockIdx.x) ? (ntiles_all - blockIdx.x + G - 1) / G : 0;
  const int nchunks = (int)((ntiles_all + (long long)G * tiles_per_cta - 1) / ((long long)G * tiles_per_cta));
  if (tid == 0 && (mode & 1)) {
    if (my_tiles > 0) issue_tile(j, s, 0, tile_start(0), TILE);
    if (my_tiles > 1) issue_tile(j, s, 1, tile_start(1), TILE);
  }
  const uint32_t region0 = blockIdx.x * j.cap_sub, region_stride = G * j.cap_sub;
  float4 va[PPT / 4];
  load_values<PPT>(j, (my_tiles > 0 && (mode & 1)) ? tile_start(0) : -1, va);
  long long k = 0;
  for (int ph = 0; ph <= nchunks; ph++) {
    if (ph < nchunks && (mode & 1)) {
      BinJob jj = j;
      jj.out = (ph & 1) ? buf1 : buf0;
      for (int t = 0; t < tiles_per_cta && k < my_tiles; t++, k++)
        bin_tile2<PPT>(jj, s, (int)(k & 1), (uint32_t)((k >> 1) & 1), tile_start(k), k + 2 < my_tiles ? tile_start(k + 2) : -1,
                       k + 1 < my_tiles ? tile_start(k + 1) : -1, va, region0, region_stride);
      __syncthreads();
      uint32_t* cnt = (ph & 1) ? cnt1 : cnt0;
      for (uint32_t b = tid; b < j.r.nb; b += BT) { cnt[(size_t)b * G + blockIdx.x] = s.scur[b]; s.scur[b] = 0; }
    }
    if (ph > 0 && (mode & 2)) {
      const unsigned long long* in = ((ph - 1) & 1) ? buf1 : buf0;
      const uint32_t* cnt = ((ph - 1) & 1) ? cnt1 : cnt0;
      const uint32_t b = blockIdx.x;
      for (uint32_t sreg = warp; sreg < G; sreg += 32) {
        const uint32_t nrec = min(__ldcg(cnt + (size_t)b * G + sreg), j.cap_sub);
        const unsigned long long* base = in + ((size_t)b * G + sreg) * j.cap_sub;
        const uint4* r4 = (const uint4*)base;
        const uint32_t n2 = nrec >> 1;
        for (uint32_t i = lane; i < n2; i += 32) {
          const uint4 q = __ldcg(r4 + i);
          const uint32_t l0 = q.x & 0xffffu, l1 = q.z & 0xffffu;
          atomicAdd(s_sum + l0, (double)__uint_as_float(q.y));
          atomicAdd(s_cnt + l0, 1u);
          atomicAdd(s_sum + l1, (double)__uint_as_float(q.w));
          atomicAdd(s_cnt + l1, 1u);
        }
        if ((nrec & 1) && lane == 0) {
          const unsigned long long q = __ldcg(base + nrec - 1);
          const uint32_t l0 = (uint32_t)q & 0xffffu;
          atomicAdd(s_sum + l0, (double)__uint_as_float((uint32_t)(q >> 32)));
          atomicAdd(s_cnt + l0, 1u);
        }
      }
    }
    grid.sync();
  }
  const uint32_t c0 = blockIdx.x * j.r.cpb;
  const uint32_t ncell = j.m.W * j.m.H;
  for (uint32_t q = tid; q < j.r.cpb && c0 + q < ncell; q += BT)
    if (s_cnt[q]) { sum_canvas[c0 + q] += s_sum[q]; cnt_canvas[c0 + q] += s_cnt[q]; }
}

// ---- references ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ref_mean(const float4* __restrict__ x4, const float4* __restrict__ y4, const float4* __restrict__ v4,
                                                  long long n4, Map m, double* sum_canvas, uint32_t* cnt_canvas) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 xa = __ldcs(x4 + i), ya = __ldcs(y4 + i), va = __ldcs(v4 + i);
    const float xs[4] = {xa.x, xa.y, xa.z, xa.w}, ys[4] = {ya.x, ya.y, ya.z, ya.w}, vs[4] = {va.x, va.y, va.z, va.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int c = cell_of(m, xs[k], ys[k]);
      if (c >= 0 && vs[k] == vs[k]) { atomicAdd(sum_canvas + c, (double)vs[k]); atomicAdd(cnt_canvas + c, 1u); }
    }
  }
}
__global__ void __launch_bounds__(256) k_ref_max(const float4* __restrict__ x4, const float4* __restrict__ y4, const float4* __restrict__ v4,
                                                 long long n4, Map m, int* canvas) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 xa = __ldcs(x4 + i), ya = __ldcs(y4 + i), va = __ldcs(v4 + i);
    const float xs[4] = {xa.x, xa.y, xa.z, xa.w}, ys[4] = {ya.x, ya.y, ya.z, ya.w}, vs[4] = {va.x, va.y, va.z, va.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int c = cell_of(m, xs[k], ys[k]);
      if (c >= 0 && vs[k] == vs[k]) atomicMax(canvas + c, key32(vs[k]));
    }
  }
}

static Route make_route(uint32_t ncell, uint32_t nb) {
  Route r;
  r.nb = nb;
  r.cpb = (ncell + nb - 1) / nb;
  r.inv = (uint32_t)((1ull << 32) / r.cpb);
  return r;
}

template <int PPT>
static void run_max(const float* x, const float* y, const float* v, long long n, uint32_t nb, const int* ref) {
  constexpr int TILE = BT * PPT;
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const uint32_t W = 8192, H = 8192, ncell = W * H;
  BinJob j;
  j.x = x; j.y = y; j.v = v; j.n = n;
  j.m = {(float)W, 0.f, (float)H, 0.f, W, H};
  j.r = make_route(ncell, nb);
  if (j.r.cpb > 65536) { printf("max: %u cells per band does not fit 16 bits\n", j.r.cpb); return; }
  const uint32_t G = sms;
  j.cap_sub = ((uint32_t)((double)n / nb / G * 1.08) + 256) & ~1u;
  uint32_t *cnt, *ovf;
  int* canvas;
  CK(cudaMalloc(&j.out, (size_t)j.cap_sub * nb * G * 8));
  CK(cudaMalloc(&cnt, (size_t)nb * G * 4 + 4));
  CK(cudaMalloc(&canvas, (size_t)ncell * 4));
  ovf = cnt + (size_t)nb * G;
  j.overflow = ovf;
  const size_t smem1 = bin_smem_bytes(nb, TILE), smem2 = (size_t)j.r.cpb * 4;
  CK(cudaFuncSetAttribute(k_bin2<PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  CK(cudaFuncSetAttribute(k_eat_max2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  printf("max2  8192^2 tile %d bands %u cells/band %u sub-region %u records (%.2f GB) smem bin %zu eat %zu\n", TILE, nb, j.r.cpb, j.cap_sub,
         (double)j.cap_sub * nb * G * 8 / 1e9, smem1, smem2);
  cudaEvent_t e0, e1, e2;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
  float b1 = 1e30f, b2 = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaMemset(cnt, 0, (size_t)nb * G * 4 + 4));
    CK(cudaEventRecord(e0));
    k_bin2<PPT><<<G, BT, smem1>>>(j, cnt);
    CK(cudaEventRecord(e1));
    k_eat_max2<<<sms, 1024, smem2>>>(j.out, j.cap_sub, G, cnt, j.r, ncell, canvas);
    CK(cudaEventRecord(e2));
    CK(cudaEventSynchronize(e2));
    CK(cudaGetLastError());
    float m1, m2;
    CK(cudaEventElapsedTime(&m1, e0, e1)); CK(cudaEventElapsedTime(&m2, e1, e2));
    if (rep > 0) { b1 = fminf(b1, m1); b2 = fminf(b2, m2); }
  }
  uint32_t hov = 0;
  CK(cudaMemcpy(&hov, ovf, 4, cudaMemcpyDeviceToHost));
  printf("  bin %.3f ms (%.1f Gpts/s, %.2f TB/s of 20 B/pt)  eat %.3f ms (%.1f Gpts/s)  total %.3f ms = %.1f Gpts/s  overflow %u\n",
         b1, n / b1 / 1e6, n * 20.0 / b1 / 1e9, b2, n / b2 / 1e6, b1 + b2, n / (b1 + b2) / 1e6, hov);
  if (ref) {
    std::vector<int> a(ncell), b(ncell);
    CK(cudaMemcpy(a.data(), canvas, (size_t)ncell * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), ref, (size_t)ncell * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (uint32_t i = 0; i < ncell; i++) bad += a[i] != b[i];
    printf("  check vs global-atomicMax reference: %zu mismatches\n", bad);
  }
  CK(cudaFree(j.out)); CK(cudaFree(cnt)); CK(cudaFree(canvas));
}

template <int PPT>
static void run_mean(const float* x, const float* y, const float* v, long long n, int tiles_per_cta, const double* ref_sum, const uint32_t* ref_cnt) {
  constexpr int TILE = BT * PPT;
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const uint32_t W = 900, H = 525, ncell = W * H, G = sms;
  BinJob j;
  j.x = x; j.y = y; j.v = v; j.n = n;
  j.m = {(float)W, 0.f, (float)H, 0.f, W, H};
  j.r = make_route(ncell, G);
  j.cap_sub = ((uint32_t)((double)TILE * tiles_per_cta / G * 1.3) + 64) & ~1u;
  unsigned long long *buf0, *buf1;
  uint32_t *cnt, *cntc;
  double* sum;
  const size_t bufb = (size_t)j.cap_sub * G * G * 8;
  CK(cudaMalloc(&buf0, bufb)); CK(cudaMalloc(&buf1, bufb));
  CK(cudaMemset(buf0, 0, bufb)); CK(cudaMemset(buf1, 0, bufb));
  CK(cudaMalloc(&cnt, (size_t)G * G * 8 + 4));
  CK(cudaMalloc(&sum, (size_t)ncell * 8));
  CK(cudaMalloc(&cntc, (size_t)ncell * 4));
  uint32_t* cnt0 = cnt; uint32_t* cnt1 = cnt + (size_t)G * G;
  j.overflow = cnt + 2 * (size_t)G * G;
  j.out = buf0;
  const size_t smem = bin_smem_bytes(G, TILE) + (size_t)j.r.cpb * 12 + 16;
  auto kern = k_routed_mean2<PPT>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, BT, smem));
  printf("mean2 tile %d bands %u cells/band %u chunk %.2f Mpts sub-region %u records (%.1f MB x2) smem %zu B occupancy %d\n", TILE, G, j.r.cpb,
         (double)TILE * tiles_per_cta * G / 1e6, j.cap_sub, bufb / 1e6, smem, occ);
  if (occ < 1) { printf("  -> does not fit\n"); return; }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int mode : {3, 1, 2}) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      CK(cudaMemset(cnt, 0, (size_t)G * G * 8 + 4));
      CK(cudaMemset(sum, 0, (size_t)ncell * 8));
      CK(cudaMemset(cntc, 0, (size_t)ncell * 4));
      if (mode == 2) {
        std::vector<uint32_t> h(2 * (size_t)G * G, (uint32_t)(TILE * tiles_per_cta / G));
        CK(cudaMemcpy(cnt, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
      }
      void* args[] = {&j, &tiles_per_cta, &buf0, &buf1, &cnt0, &cnt1, &sum, &cntc, &mode};
      CK(cudaEventRecord(e0));
      CK(cudaLaunchCooperativeKernel((void*)kern, dim3(G), dim3(BT), args, smem, 0));
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    uint32_t hov = 0;
    CK(cudaMemcpy(&hov, j.overflow, 4, cudaMemcpyDeviceToHost));
    printf("  mode %d (%s): %.3f ms  %.1f Gpts/s  overflow %u\n", mode, mode == 3 ? "bin+eat" : mode == 1 ? "bin only" : "eat only",
           best, n / best / 1e6, hov);
    if (mode == 3 && ref_sum) {
      std::vector<double> hs(ncell), rs(ncell);
      std::vector<uint32_t> hc(ncell), rc(ncell);
      CK(cudaMemcpy(hs.data(), sum, ncell * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hc.data(), cntc, ncell * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(rs.data(), ref_sum, ncell * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(rc.data(), ref_cnt, ncell * 4, cudaMemcpyDeviceToHost));
      size_t badc = 0; double worst = 0;
      for (uint32_t i = 0; i < ncell; i++) {
        if (hc[i] != rc[i]) badc++;
        double d = fabs(hs[i] - rs[i]) / fmax(1.0, fabs(rs[i]));
        if (d > worst) worst = d;
      }
      printf("  check vs global-RED reference: %zu count mismatches, worst relative sum difference %.3g\n", badc, worst);
    }
  }
  CK(cudaFree(buf0)); CK(cudaFree(buf1)); CK(cudaFree(cnt)); CK(cudaFree(sum)); CK(cudaFree(cntc));
}

int main(int argc, char** argv) {
  long long n = argc > 1 ? atoll(argv[1]) : 1000000000LL;
  n &= ~3LL;
  const char* what = argc > 2 ? argv[2] : "all";
  float *x, *y, *v;
  CK(cudaMalloc(&x, n * 4)); CK(cudaMalloc(&y, n * 4)); CK(cudaMalloc(&v, n * 4));
  gen_uniform<<<148 * 8, 256>>>(x, y, v, (size_t)n, 12345u);
  CK(cudaDeviceSynchronize());
  const float4 *x4 = (const float4*)x, *y4 = (const float4*)y, *v4 = (const float4*)v;
  printf("n = %lld points\n", n);
  if (!strcmp(what, "all") || !strcmp(what, "max")) {
    const uint32_t ncell = 8192u * 8192u;
    int* ref;
    CK(cudaMalloc(&ref, (size_t)ncell * 4));
    Map m = {8192.f, 0.f, 8192.f, 0.f, 8192, 8192};
    {
      std::vector<int> init(ncell, INT_MIN);
      CK(cudaMemcpy(ref, init.data(), (size_t)ncell * 4, cudaMemcpyHostToDevice));
    }
    k_ref_max<<<148 * 8, 256>>>(x4, y4, v4, n / 4, m, ref);
    CK(cudaDeviceSynchronize());
    run_max<8>(x, y, v, n, 1480, ref);
    run_max<8>(x, y, v, n, 1184, ref);
    run_max<4>(x, y, v, n, 1480, ref);
    CK(cudaFree(ref));
  }
  if (!strcmp(what, "all") || !strcmp(what, "mean")) {
    const uint32_t ncell = 900 * 525;
    double* rs; uint32_t* rc;
    CK(cudaMalloc(&rs, ncell * 8)); CK(cudaMalloc(&rc, ncell * 4));
    Map m = {900.f, 0.f, 525.f, 0.f, 900, 525};
    CK(cudaMemset(rs, 0, ncell * 8)); CK(cudaMemset(rc, 0, ncell * 4));
    k_ref_mean<<<148 * 8, 256>>>(x4, y4, v4, n / 4, m, rs, rc);
    CK(cudaDeviceSynchronize());
    run_mean<4>(x, y, v, n, 4, rs, rc);
    run_mean<4>(x, y, v, n, 8, rs, rc);
    run_mean<4>(x, y, v, n, 16, rs, rc);
    CK(cudaFree(rs)); CK(cudaFree(rc));
  }
  return 0;
}
