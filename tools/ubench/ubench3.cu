// Micro-benchmark 3: SM-side cost of the band-privatised design (K2) with no global traffic at all.
// Every CTA owns canvas band `blockIdx.x % B` (rows [b*Hb,(b+1)*Hb)) as packed 15+1-bit counters in
// shared memory and examines ALL candidate points of a shared-memory tile (re-read with rotating
// offsets): y-band pre-filter on the float bits -> ballot compaction into a per-warp queue -> exact f64
// mapping + ATOMS for the accepted 1/B.  Reported: candidate points/s per SM and the implied whole-chip
// rate 148 * per_SM / B (each point is examined by B CTAs of a cluster).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o ubench3 ubench3.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

struct Map { double sx, tx, sy, ty, xmin, xmax, ymin, ymax; int W, H; };

constexpr int TILE = 4096;          // points per staged tile
constexpr int QCAP = 96;            // per-warp compaction queue capacity (entries of 8 B)

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k2sim(Map m, int B, int Hb, long long iters, unsigned long long* out,
                                                 unsigned int* gcanvas) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* tx = (float*)smem;                       // TILE x
  float* ty = tx + TILE;                          // TILE y
  float2* queue = (float2*)(ty + TILE);           // [warps][QCAP]
  constexpr int WARPS = THREADS / 32;
  uint32_t* band = (uint32_t*)(queue + WARPS * QCAP);   // packed counters: 2 x (15-bit count + flag) per word
  const int b = blockIdx.x % B;
  const int row0 = b * Hb;
  const int rows = min(Hb, m.H - row0);
  const int nbins = rows * m.W;
  const int nwords = (nbins + 1) / 2;
  for (int j = threadIdx.x; j < nwords; j += THREADS) band[j] = 0;
  for (int j = threadIdx.x; j < TILE; j += THREADS) {
    uint32_t h = hash32(j * 2654435761u + 17u + blockIdx.x / B);
    tx[j] = (h >> 8) * (1.0f / 16777216.0f);
    ty[j] = (hash32(h) >> 8) * (1.0f / 16777216.0f);
  }
  __syncthreads();
  // conservative f32 band limits on y (pixel rows [row0, row0+rows)): y in [ylo, yhi)
  const float ylo = (float)((row0 - 0.01) / m.sy), yhi = (float)((row0 + rows + 0.01) / m.sy);
  const uint32_t lo_bits = __float_as_uint(fmaxf(ylo, 0.0f)), span = __float_as_uint(yhi) - lo_bits;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* q = queue + warp * QCAP;
  int qn = 0;                                      // warp-uniform queue fill
  unsigned long long accepted = 0, overflow_events = 0;
  const uint32_t lt_mask = (1u << lane) - 1u;

  auto process32 = [&](int start) {
    float2 p = q[start + lane];
    double xd = (double)p.x, yd = (double)p.y;
    if (!(xd >= m.xmin && xd <= m.xmax && yd >= m.ymin && yd <= m.ymax)) return;
    int xx = __double2int_rz(__dadd_rn(__dmul_rn(xd, m.sx), m.tx));
    int yy = __double2int_rz(__dadd_rn(__dmul_rn(yd, m.sy), m.ty));
    if (xx >= m.W) xx = m.W - 1;
    if (yy >= m.H) yy = m.H - 1;
    int ry = yy - row0;
    if (ry < 0 || ry >= rows) return;               // the f32 pre-filter is conservative: exact test here
    int bin = ry * m.W + xx;
    uint32_t inc = (bin & 1) ? 0x10000u : 1u;
    uint32_t old = atomicAdd(band + (bin >> 1), inc);
    uint32_t f = (bin & 1) ? (old >> 16) : (old & 0xffffu);
    accepted++;
    if (f == 0x7fffu) {                             // 15-bit counter wrapped into its flag bit: spill 32768
      atomicSub(band + (bin >> 1), (bin & 1) ? 0x80000000u : 0x8000u);
      atomicAdd(gcanvas + (size_t)row0 * m.W + bin, 32768u);
      overflow_events++;
    }
  };

  for (long long it = 0; it < iters; it++) {
    // one pass over the tile: each thread takes 4 consecutive candidates per step
    for (int base = (threadIdx.x * 4 + (int)(it & 3) * 4 * THREADS) % TILE, step = 0; step < TILE / (4 * THREADS) + (TILE < 4 * THREADS); step++) {
      int o = (base + step * 4 * THREADS) % TILE;
      float4 y4 = *(const float4*)(ty + o);
      float4 x4 = *(const float4*)(tx + o);
      float ys[4] = {y4.x, y4.y, y4.z, y4.w}, xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        bool acc = (__float_as_uint(ys[k]) - lo_bits) < span;
        uint32_t mask = __ballot_sync(0xffffffffu, acc);
        if (acc) q[qn + __popc(mask & lt_mask)] = make_float2(xs[k], ys[k]);
        qn += __popc(mask);
        if (qn >= 32) {                           // warp-uniform
          __syncwarp();
          process32(qn - 32);
          qn -= 32;
          __syncwarp();
        }
      }
    }
  }
  // drain
  __syncwarp();
  if (lane < qn) {
    float2 p = q[lane];
    (void)p;
  }
  __syncthreads();
  // flush: unpack and add to the global canvas (plain REDs here; the product uses TMA bulk reduce)
  unsigned long long total = 0;
  for (int j = threadIdx.x; j < nbins; j += THREADS) {
    uint32_t w = band[j >> 1];
    uint32_t c = (j & 1) ? (w >> 16) : (w & 0xffffu);
    total += c;
    if (c) atomicAdd(gcanvas + (size_t)row0 * m.W + j, c);
  }
  atomicAdd(out, accepted);
  atomicAdd(out + 1, total);
  atomicAdd(out + 2, overflow_events);
}

int main() {
  CK(cudaSetDevice(0));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device %s sms=%d\n", p.name, sms);
  unsigned long long* out; unsigned int* gcanvas;
  CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&gcanvas, 900 * 525 * 4));
  Map m; m.W = 900; m.H = 525; m.xmin = 0; m.xmax = 1; m.ymin = 0; m.ymax = 1; m.sx = 900; m.tx = 0; m.sy = 525; m.ty = 0;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int B = 3; B <= 8; B++) {
    int Hb = (m.H + B - 1) / B;
    size_t band_bytes = ((size_t)Hb * m.W + 1) / 2 * 4;
    for (int threads = 512; threads <= 1024; threads *= 2) {
      size_t sb = TILE * 8 + (size_t)(threads / 32) * QCAP * 8 + band_bytes;
      if (sb > 227 * 1024) { printf("[B=%d thr=%d] smem %zu too large\n", B, threads, sb); continue; }
      long long iters = 2000;
      CK(cudaMemset(out, 0, 64)); CK(cudaMemset(gcanvas, 0, 900 * 525 * 4));
      float ms = 0;
      for (int rep = 0; rep < 2; rep++) {
        CK(cudaMemset(out, 0, 64));
        CK(cudaEventRecord(e0));
        if (threads == 512) {
          CK(cudaFuncSetAttribute(k2sim<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
          k2sim<512><<<sms, 512, sb>>>(m, B, Hb, iters, out, gcanvas);
        } else {
          CK(cudaFuncSetAttribute(k2sim<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
          k2sim<1024><<<sms, 1024, sb>>>(m, B, Hb, iters, out, gcanvas);
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      unsigned long long h[3]; CK(cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost));
      double cand_per_sm = (double)TILE * iters / (ms * 1e-3);           // candidates examined per SM per second
      printf("[B=%d thr=%4d smem=%6zu] %7.3f ms  %6.2f Gcand/s/SM  -> chip-equivalent %7.1f Gpts/s   accepted %llu flushed %llu ovf %llu\n",
             B, threads, sb, ms, cand_per_sm * 1e-9, cand_per_sm * sms / B * 1e-9, h[0], h[1], h[2]);
    }
  }
  return 0;
}
