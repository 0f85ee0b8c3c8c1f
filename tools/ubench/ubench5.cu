// Micro-benchmark 5: the ROUTED design - "bin, then accumulate in shared memory" (VERDICT r01, next-round item 1).
//
//   pass 1  every CTA streams a tile of points, maps them, and counting-sorts the tile by BAND (a contiguous range of
//           canvas cells) in shared memory; each band's run leaves as coalesced 8-byte records (band << 16 | cell in band,
//           f32 value) into that band's bucket (space reserved with one global atomic per band per tile).
//   pass 2  the CTA that owns a band consumes its bucket with the band's accumulators in shared memory
//           (mean: f64 sums via CAS + u32 counts; max: u32 keys via atomicMax) - no global atomics at all.
//
// Two shapes:
//   "mean"  900x525, fused persistent cooperative kernel, records staged in an L2-resident double buffer chunk by chunk
//           (DRAM traffic stays 12 B / point); one grid barrier per chunk.
//   "max"   8192x8192 (canvas beyond L2), pass 1 writes all records to DRAM, pass 2 walks 1480 buckets with a u32 key tile
//           in shared memory (traffic 12 + 8 + 8 B / point instead of 5 banded re-reads).
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o ubench5 ubench5.cu
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
struct Route { uint32_t cpb, inv, nb; };     // cells per band, floor(2^32 / cpb), number of bands

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

// ---------------------------------------------------------------------------------------------------------------------
// pass 1 over one tile: THREADS * PPT points starting at float4 index t4.  Shared: hist[nb], base[nb], gdel[nb], rec[].
template <int THREADS, int PPT>
__device__ __forceinline__ void bin_tile(const float4* __restrict__ x4, const float4* __restrict__ y4, const float4* __restrict__ v4,
                                         long long t4, long long n4, const Map& m, const Route& r,
                                         uint32_t* hist, uint32_t* base, uint32_t* gdel, unsigned long long* rec,
                                         unsigned long long* __restrict__ out, uint32_t cap, uint32_t* cursor, uint32_t* overflow) {
  const int tid = threadIdx.x;
  for (int b = tid; b < (int)r.nb; b += THREADS) hist[b] = 0;
  __syncthreads();
  uint32_t key[PPT], rank[PPT];
  float val[PPT];
#pragma unroll
  for (int u = 0; u < PPT / 4; u++) {
    const long long i4 = t4 + (long long)u * THREADS + tid;
    float4 xa, ya, va;
    if (i4 < n4) { xa = __ldcs(x4 + i4); ya = __ldcs(y4 + i4); va = __ldcs(v4 + i4); }
    else { xa = ya = va = make_float4(NAN, NAN, NAN, NAN); }
    const float xs[4] = {xa.x, xa.y, xa.z, xa.w}, ys[4] = {ya.x, ya.y, ya.z, ya.w}, vs[4] = {va.x, va.y, va.z, va.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int c = cell_of(m, xs[k], ys[k]);
      const bool ok = c >= 0 && vs[k] == vs[k];
      const uint32_t kk = ok ? key_of(r, (uint32_t)c) : 0xffffffffu;
      key[u * 4 + k] = kk;
      val[u * 4 + k] = vs[k];
      rank[u * 4 + k] = ok ? atomicAdd(hist + (kk >> 16), 1u) : 0u;
    }
  }
  __syncthreads();
  // exclusive scan of hist -> base (warp 0), then one global atomic per non-empty band reserves the run's space
  if (tid < 32) {
    const int per = ((int)r.nb + 31) / 32;
    const int lo = tid * per, hi = min(lo + per, (int)r.nb);
    uint32_t s = 0;
    for (int b = lo; b < hi; b++) s += hist[b];
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += t; }
    uint32_t run = incl - s;
    for (int b = lo; b < hi; b++) { base[b] = run; run += hist[b]; }
  }
  __syncthreads();
  for (int b = tid; b < (int)r.nb; b += THREADS) {
    const uint32_t h = hist[b];
    uint32_t g = 0;
    if (h) {
      g = atomicAdd(cursor + b, h);
      if (g + h > cap) { atomicOr(overflow, 1u); g = 0; }      // prototype: flag it (the real path falls back to REDs)
    }
    gdel[b] = (uint32_t)b * cap + g - base[b];
  }
#pragma unroll
  for (int k = 0; k < PPT; k++) {
    if (key[k] != 0xffffffffu)
      rec[base[key[k] >> 16] + rank[k]] = ((unsigned long long)__float_as_uint(val[k]) << 32) | key[k];
  }
  __syncthreads();
  const uint32_t total = base[r.nb - 1] + hist[r.nb - 1];
  for (uint32_t j = tid; j < total; j += THREADS) {
    const unsigned long long rr = rec[j];
    const uint32_t b = ((uint32_t)rr) >> 16;
    out[(size_t)gdel[b] + j] = rr;
  }
  __syncthreads();
}

// pass 2 (mean): consume nrec records of one band into shared f64 sums + u32 counts
template <int THREADS>
__device__ __forceinline__ void eat_mean(const unsigned long long* __restrict__ recs, uint32_t nrec, double* s_sum, uint32_t* s_cnt) {
  const uint4* r4 = (const uint4*)recs;
  const uint32_t n2 = nrec >> 1;
  for (uint32_t i = threadIdx.x; i < n2; i += THREADS) {
    const uint4 q = __ldcg(r4 + i);
    const uint32_t l0 = q.x & 0xffffu, l1 = q.z & 0xffffu;
    atomicAdd(s_sum + l0, (double)__uint_as_float(q.y));
    atomicAdd(s_cnt + l0, 1u);
    atomicAdd(s_sum + l1, (double)__uint_as_float(q.w));
    atomicAdd(s_cnt + l1, 1u);
  }
  if ((nrec & 1) && threadIdx.x == 0) {
    const unsigned long long q = __ldcg(recs + nrec - 1);
    const uint32_t l0 = (uint32_t)q & 0xffffu;
    atomicAdd(s_sum + l0, (double)__uint_as_float((uint32_t)(q >> 32)));
    atomicAdd(s_cnt + l0, 1u);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// fused mean: persistent cooperative kernel, one band per CTA, L2-resident double-buffered record chunks
// mode: 3 = both passes, 1 = pass 1 only, 2 = pass 2 only (re-eats whatever the buffers hold)
template <int THREADS, int PPT, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_routed_mean(const float4* __restrict__ x4, const float4* __restrict__ y4, const float4* __restrict__ v4,
                                                         long long n4, Map m, Route r, int tiles_per_cta,
                                                         unsigned long long* buf0, unsigned long long* buf1, uint32_t cap,
                                                         uint32_t* cur0, uint32_t* cur1, uint32_t* overflow,
                                                         double* __restrict__ sum_canvas, uint32_t* __restrict__ cnt_canvas, int mode) {
  extern __shared__ __align__(16) unsigned char smem[];
  cg::grid_group grid = cg::this_grid();
  double* s_sum = (double*)smem;
  uint32_t* s_cnt = (uint32_t*)(s_sum + r.cpb);
  uint32_t* hist = s_cnt + ((r.cpb + 3) & ~3u);
  uint32_t* base = hist + r.nb;
  uint32_t* gdel = base + r.nb;
  unsigned long long* rec = (unsigned long long*)(((uintptr_t)(gdel + r.nb) + 15) & ~(uintptr_t)15);
  for (uint32_t j = threadIdx.x; j < r.cpb; j += THREADS) { s_sum[j] = 0.0; s_cnt[j] = 0; }
  __syncthreads();
  constexpr long long TILE4 = (long long)THREADS * PPT / 4;
  const long long chunk4 = TILE4 * tiles_per_cta * gridDim.x;
  const int nchunks = (int)((n4 + chunk4 - 1) / chunk4);
  for (int ph = 0; ph <= nchunks; ph++) {
    if (ph < nchunks && (mode & 1)) {
      unsigned long long* out = (ph & 1) ? buf1 : buf0;
      uint32_t* cursor = (ph & 1) ? cur1 : cur0;
      for (int t = 0; t < tiles_per_cta; t++) {
        const long long t4 = (long long)ph * chunk4 + ((long long)t * gridDim.x + blockIdx.x) * TILE4;
        if (t4 >= n4) break;
        bin_tile<THREADS, PPT>(x4, y4, v4, t4, n4, m, r, hist, base, gdel, rec, out, cap, cursor, overflow);
      }
    }
    if (ph > 0 && (mode & 2)) {
      const unsigned long long* in = ((ph - 1) & 1) ? buf1 : buf0;
      uint32_t* cursor = ((ph - 1) & 1) ? cur1 : cur0;
      const uint32_t nrec = min(__ldcg(cursor + blockIdx.x), cap);
      eat_mean<THREADS>(in + (size_t)blockIdx.x * cap, nrec, s_sum, s_cnt);
      __syncthreads();
      if (threadIdx.x == 0 && (mode & 1)) cursor[blockIdx.x] = 0;
    } else if (ph > 0 && threadIdx.x == 0) {
      (((ph - 1) & 1) ? cur1 : cur0)[blockIdx.x] = 0;
    }
    grid.sync();
  }
  const uint32_t c0 = blockIdx.x * r.cpb;
  const uint32_t ncell = m.W * m.H;
  for (uint32_t j = threadIdx.x; j < r.cpb && c0 + j < ncell; j += THREADS) {
    if (s_cnt[j]) { sum_canvas[c0 + j] += s_sum[j]; cnt_canvas[c0 + j] += s_cnt[j]; }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// reference: one f64 RED + one u32 RED per point (K1-style), for the correctness check and as the slow baseline
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
__device__ __forceinline__ int key32(float f) { int b = __float_as_int(f + 0.0f); return b ^ ((b >> 31) & 0x7fffffff); }
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

// ---------------------------------------------------------------------------------------------------------------------
// "max" shape: pass 1 to DRAM buckets (grid-stride over tiles), pass 2 with a u32 key tile per bucket in shared memory
template <int THREADS, int PPT, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_bin_all(const float4* __restrict__ x4, const float4* __restrict__ y4, const float4* __restrict__ v4,
                                                     long long n4, Map m, Route r, unsigned long long* out, uint32_t cap,
                                                     uint32_t* cursor, uint32_t* overflow) {
  extern __shared__ __align__(16) unsigned char smem[];
  uint32_t* hist = (uint32_t*)smem;
  uint32_t* base = hist + r.nb;
  uint32_t* gdel = base + r.nb;
  unsigned long long* rec = (unsigned long long*)(((uintptr_t)(gdel + r.nb) + 15) & ~(uintptr_t)15);
  constexpr long long TILE4 = (long long)THREADS * PPT / 4;
  for (long long t4 = blockIdx.x * TILE4; t4 < n4; t4 += gridDim.x * TILE4)
    bin_tile<THREADS, PPT>(x4, y4, v4, t4, n4, m, r, hist, base, gdel, rec, out, cap, cursor, overflow);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_eat_max(const unsigned long long* __restrict__ recs, uint32_t cap, const uint32_t* __restrict__ cursor,
                                                     Route r, uint32_t ncell, int* __restrict__ canvas) {
  extern __shared__ __align__(16) unsigned char smem[];
  int* tile = (int*)smem;
  for (uint32_t b = blockIdx.x; b < r.nb; b += gridDim.x) {
    for (uint32_t j = threadIdx.x; j < r.cpb; j += THREADS) tile[j] = INT_MIN;
    __syncthreads();
    const uint32_t nrec = min(cursor[b], cap);
    const uint4* r4 = (const uint4*)(recs + (size_t)b * cap);
    const uint32_t n2 = nrec >> 1;
    for (uint32_t i = threadIdx.x; i < n2; i += THREADS) {
      const uint4 q = __ldcs(r4 + i);
      atomicMax(tile + (q.x & 0xffffu), key32(__uint_as_float(q.y)));
      atomicMax(tile + (q.z & 0xffffu), key32(__uint_as_float(q.w)));
    }
    if ((nrec & 1) && threadIdx.x == 0) {
      const unsigned long long q = recs[(size_t)b * cap + nrec - 1];
      atomicMax(tile + ((uint32_t)q & 0xffffu), key32(__uint_as_float((uint32_t)(q >> 32))));
    }
    __syncthreads();
    const uint32_t c0 = b * r.cpb;
    for (uint32_t j = threadIdx.x; j < r.cpb && c0 + j < ncell; j += THREADS) canvas[c0 + j] = tile[j];   // exclusive owner: plain store
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
static Route make_route(uint32_t ncell, uint32_t nb) {
  Route r;
  r.nb = nb;
  r.cpb = (ncell + nb - 1) / nb;
  r.inv = (uint32_t)((1ull << 32) / r.cpb);
  return r;
}

template <int THREADS, int PPT, int ctas_per_sm>
static void run_mean(const float4* x4, const float4* y4, const float4* v4, long long n, int tiles_per_cta,
                     const double* ref_sum, const uint32_t* ref_cnt) {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const uint32_t W = 900, H = 525, ncell = W * H;
  const int grid = sms * ctas_per_sm;
  Map m = {(float)W, 0.f, (float)H, 0.f, W, H};
  Route r = make_route(ncell, (uint32_t)grid);
  const long long chunk = (long long)THREADS * PPT * tiles_per_cta * grid;
  uint32_t cap = (uint32_t)((chunk / grid) * 5 / 4 + 64) & ~1u;
  unsigned long long *buf0, *buf1;
  uint32_t *cur, *ovf, *cnt;
  double* sum;
  CK(cudaMalloc(&buf0, (size_t)cap * grid * 8));
  CK(cudaMalloc(&buf1, (size_t)cap * grid * 8));
  CK(cudaMemset(buf0, 0, (size_t)cap * grid * 8));
  CK(cudaMemset(buf1, 0, (size_t)cap * grid * 8));
  CK(cudaMalloc(&cur, (size_t)grid * 8 + 4));
  CK(cudaMalloc(&sum, (size_t)ncell * 8));
  CK(cudaMalloc(&cnt, (size_t)ncell * 4));
  ovf = cur + 2 * grid;
  size_t smem = (size_t)r.cpb * 8 + ((r.cpb + 3) & ~3u) * 4 + (size_t)r.nb * 12 + 16 + (size_t)THREADS * PPT * 8;
  auto kern = k_routed_mean<THREADS, PPT, ctas_per_sm>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
  printf("mean  THREADS %d PPT %d ctas/SM %d (occupancy %d) bands %u cells/band %u chunk %.2f Mpts (records %.1f MB x2) smem %zu B\n",
         THREADS, PPT, ctas_per_sm, occ, r.nb, r.cpb, chunk / 1e6, (double)cap * grid * 8 / 1e6, smem);
  if (occ < ctas_per_sm) { printf("  -> does not fit, skipped\n"); return; }
  long long n4 = n / 4;
  uint32_t* cur0 = cur; uint32_t* cur1 = cur + grid;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int mode : {3, 1, 2}) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      CK(cudaMemset(cur, 0, (size_t)grid * 8 + 4));
      CK(cudaMemset(sum, 0, (size_t)ncell * 8));
      CK(cudaMemset(cnt, 0, (size_t)ncell * 4));
      if (mode == 2) {   // leave something to eat: fill the cursors as a real chunk would
        std::vector<uint32_t> h(2 * grid, (uint32_t)(chunk / grid));
        CK(cudaMemcpy(cur, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
      }
      void* args[] = {&x4, &y4, &v4, &n4, &m, &r, &tiles_per_cta, &buf0, &buf1, &cap, &cur0, &cur1, &ovf, &sum, &cnt, &mode};
      CK(cudaEventRecord(e0));
      CK(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(THREADS), args, smem, 0));
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    uint32_t hov = 0;
    CK(cudaMemcpy(&hov, ovf, 4, cudaMemcpyDeviceToHost));
    printf("  mode %d (%s): %.3f ms  %.1f Gpts/s  overflow %u\n", mode, mode == 3 ? "bin+eat" : mode == 1 ? "bin only" : "eat only",
           best, n / best / 1e6, hov);
    if (mode == 3 && ref_sum) {
      std::vector<double> hs(ncell), rs(ncell);
      std::vector<uint32_t> hc(ncell), rc(ncell);
      CK(cudaMemcpy(hs.data(), sum, ncell * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hc.data(), cnt, ncell * 4, cudaMemcpyDeviceToHost));
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
  CK(cudaFree(buf0)); CK(cudaFree(buf1)); CK(cudaFree(cur)); CK(cudaFree(sum)); CK(cudaFree(cnt));
}

template <int THREADS, int PPT, int ctas_per_sm>
static void run_max(const float4* x4, const float4* y4, const float4* v4, long long n, uint32_t nb, const int* ref) {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const uint32_t W = 8192, H = 8192, ncell = W * H;
  Map m = {(float)W, 0.f, (float)H, 0.f, W, H};
  Route r = make_route(ncell, nb);
  if (r.cpb > 65536) { printf("max: %u cells per band does not fit 16 bits\n", r.cpb); return; }
  uint32_t cap = (uint32_t)((n / nb) * 21 / 20 + 4096) & ~1u;
  unsigned long long* recs;
  uint32_t* cur;
  int* canvas;
  CK(cudaMalloc(&recs, (size_t)cap * nb * 8));
  CK(cudaMalloc(&cur, (size_t)nb * 4 + 4));
  CK(cudaMalloc(&canvas, (size_t)ncell * 4));
  uint32_t* ovf = cur + nb;
  const size_t smem1 = (size_t)nb * 12 + 16 + (size_t)THREADS * PPT * 8;
  const size_t smem2 = (size_t)r.cpb * 4;
  auto k1 = k_bin_all<THREADS, PPT, ctas_per_sm>;
  auto k2 = k_eat_max<1024>;
  CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  CK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  printf("max   8192^2 THREADS %d PPT %d ctas/SM %d bands %u cells/band %u records %.2f GB smem1 %zu smem2 %zu\n",
         THREADS, PPT, ctas_per_sm, nb, r.cpb, (double)cap * nb * 8 / 1e9, smem1, smem2);
  long long n4 = n / 4;
  cudaEvent_t e0, e1, e2;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
  float b1 = 1e30f, b2 = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaMemset(cur, 0, (size_t)nb * 4 + 4));
    CK(cudaEventRecord(e0));
    k1<<<sms * ctas_per_sm, THREADS, smem1>>>(x4, y4, v4, n4, m, r, recs, cap, cur, ovf);
    CK(cudaEventRecord(e1));
    k2<<<sms, 1024, smem2>>>(recs, cap, cur, r, ncell, canvas);
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
  CK(cudaFree(recs)); CK(cudaFree(cur)); CK(cudaFree(canvas));
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
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  printf("n = %lld points\n", n);

  if (!strcmp(what, "all") || !strcmp(what, "mean")) {
    const uint32_t ncell = 900 * 525;
    double* rs; uint32_t* rc;
    CK(cudaMalloc(&rs, ncell * 8)); CK(cudaMalloc(&rc, ncell * 4));
    Map m = {900.f, 0.f, 525.f, 0.f, 900, 525};
    float best = 1e30f;
    for (int rep = 0; rep < 2; rep++) {
      CK(cudaMemset(rs, 0, ncell * 8)); CK(cudaMemset(rc, 0, ncell * 4));
      CK(cudaEventRecord(e0));
      k_ref_mean<<<148 * 8, 256>>>(x4, y4, v4, n / 4, m, rs, rc);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms);
    }
    printf("reference mean (2 global REDs / point): %.3f ms  %.1f Gpts/s\n", best, n / best / 1e6);
    run_mean<512, 8, 2>(x4, y4, v4, n, 4, rs, rc);
    run_mean<512, 8, 2>(x4, y4, v4, n, 8, rs, rc);
    run_mean<512, 16, 2>(x4, y4, v4, n, 2, rs, rc);
    run_mean<512, 16, 2>(x4, y4, v4, n, 4, rs, rc);
    run_mean<1024, 8, 1>(x4, y4, v4, n, 4, rs, rc);
    run_mean<1024, 8, 1>(x4, y4, v4, n, 8, rs, rc);
    run_mean<1024, 16, 1>(x4, y4, v4, n, 4, rs, rc);
    run_mean<256, 16, 4>(x4, y4, v4, n, 4, rs, rc);
    run_mean<256, 16, 4>(x4, y4, v4, n, 2, rs, rc);
    CK(cudaFree(rs)); CK(cudaFree(rc));
  }
  if (!strcmp(what, "all") || !strcmp(what, "max")) {
    const uint32_t ncell = 8192u * 8192u;
    int* ref;
    CK(cudaMalloc(&ref, (size_t)ncell * 4));
    Map m = {8192.f, 0.f, 8192.f, 0.f, 8192, 8192};
    {
      std::vector<int> init(ncell, INT_MIN);
      CK(cudaMemcpy(ref, init.data(), (size_t)ncell * 4, cudaMemcpyHostToDevice));
    }
    CK(cudaEventRecord(e0));
    k_ref_max<<<148 * 8, 256>>>(x4, y4, v4, n / 4, m, ref);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("reference max 8192^2 (1 global atomicMax / point, unbanded): %.3f ms  %.1f Gpts/s\n", ms, n / ms / 1e6);
    run_max<1024, 16, 1>(x4, y4, v4, n, 1480, ref);
    run_max<1024, 8, 1>(x4, y4, v4, n, 1480, ref);
    run_max<512, 16, 2>(x4, y4, v4, n, 1480, ref);
    run_max<1024, 16, 1>(x4, y4, v4, n, 1184, ref);
    CK(cudaFree(ref));
  }
  return 0;
}
