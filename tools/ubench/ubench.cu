// Micro-benchmarks that size the scatter primitives on B200 before the design is fixed.
// Not part of the product; results are summarised in profiles/r01_ubench.md.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o ubench ubench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// ---------------------------------------------------------------- data generation
__global__ void gen_uniform(float* x, float* y, float* v, size_t n, uint32_t seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    uint32_t a = hash32((uint32_t)i * 2654435761U + seed);
    uint32_t b = hash32(a ^ 0x9e3779b9U);
    uint32_t c = hash32(b ^ 0x85ebca6bU);
    x[i] = (a >> 8) * (1.0f / 16777216.0f);
    y[i] = (b >> 8) * (1.0f / 16777216.0f);
    v[i] = (c >> 8) * (1.0f / 16777216.0f) - 0.5f;
  }
}
// 8 gaussian blobs, sigma 0.03 (clustered / contention case)
__global__ void gen_cluster(float* x, float* y, size_t n, uint32_t seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    uint32_t a = hash32((uint32_t)i * 2654435761U + seed);
    uint32_t b = hash32(a ^ 0x9e3779b9U);
    uint32_t c = hash32(b ^ 0x85ebca6bU);
    uint32_t d = hash32(c ^ 0xc2b2ae35U);
    int k = a & 7;
    float cx = 0.15f + 0.1f * k, cy = 0.2f + 0.08f * ((k * 5) & 7);
    float u1 = ((b >> 8) + 1) * (1.0f / 16777217.0f), u2 = (c >> 8) * (1.0f / 16777216.0f);
    float r = sqrtf(-2.0f * logf(u1)) * 0.03f;
    x[i] = cx + r * cospif(2.0f * u2);
    y[i] = cy + r * sinpif(2.0f * u2);
    (void)d;
  }
}
__global__ void gen_idx(uint32_t* idx, size_t n, uint32_t nbins, uint32_t seed) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    uint64_t h = hash32((uint32_t)i * 2654435761U + seed);
    h = (h << 20) ^ hash32((uint32_t)h ^ 0x1234567U);
    idx[i] = (uint32_t)(h % nbins);
  }
}

struct Map { double sx, tx, sy, ty, xmin, xmax, ymin, ymax; int W, H; };

__device__ __forceinline__ int map_exact(float x, float y, const Map& m) {
  double xd = (double)x, yd = (double)y;
  if (!(xd >= m.xmin && xd <= m.xmax && yd >= m.ymin && yd <= m.ymax)) return -1;
  int xx = __double2int_rz(__dadd_rn(__dmul_rn(xd, m.sx), m.tx));
  int yy = __double2int_rz(__dadd_rn(__dmul_rn(yd, m.sy), m.ty));
  if (xx >= m.W) xx = m.W - 1;
  if (yy >= m.H) yy = m.H - 1;
  return yy * m.W + xx;
}

// f32 fast path with exact f64 fallback near pixel edges
struct MapF { float sx, tx, sy, ty, xmin, xmax, ymin, ymax, ex, ey; };
__device__ __forceinline__ int map_fast(float x, float y, const Map& m, const MapF& f) {
  // bounds in f32 are only a pre-filter here (benchmark); product code treats edges exactly
  if (!(x >= f.xmin && x <= f.xmax && y >= f.ymin && y <= f.ymax)) return -1;
  float xf = fmaf(x, f.sx, f.tx), yf = fmaf(y, f.sy, f.ty);
  float xr = rintf(xf), yr = rintf(yf);
  bool near = (fabsf(xf - xr) < f.ex) | (fabsf(yf - yr) < f.ey);
  if (near) return map_exact(x, y, m);
  int xx = (int)floorf(xf), yy = (int)floorf(yf);
  if (xx >= m.W) xx = m.W - 1;
  if (yy >= m.H) yy = m.H - 1;
  return yy * m.W + xx;
}

// ---------------------------------------------------------------- kernels
template <int MODE>  // 0 exact+red, 1 exact no atomics, 2 fast+red, 3 fast no atomics, 4 read only
__global__ void __launch_bounds__(512) k_points(const float4* __restrict__ x4, const float4* __restrict__ y4,
                                                size_t n4, Map m, MapF f, uint32_t* __restrict__ canvas,
                                                uint32_t* __restrict__ sink) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (; i < n4; i += stride) {
    float4 xv = __ldcs(x4 + i), yv = __ldcs(y4 + i);
    float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (MODE == 4) { acc += __float_as_uint(xs[k]) ^ __float_as_uint(ys[k]); continue; }
      int b = (MODE == 0 || MODE == 1) ? map_exact(xs[k], ys[k], m) : map_fast(xs[k], ys[k], m, f);
      if (MODE == 1 || MODE == 3) acc += (uint32_t)b;
      else if (b >= 0) atomicAdd(canvas + b, 1u);
    }
  }
  if (MODE == 1 || MODE == 3 || MODE == 4) if (acc == 0x12345678u) *sink = acc;
}

// mean: f64 sum + u32 count, separate canvases (AOS=0) or interleaved 16B struct (AOS=1)
template <int AOS>
__global__ void __launch_bounds__(512) k_mean(const float4* __restrict__ x4, const float4* __restrict__ y4,
                                              const float4* __restrict__ v4, size_t n4, Map m,
                                              double* __restrict__ sum, uint32_t* __restrict__ cnt) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 xv = __ldcs(x4 + i), yv = __ldcs(y4 + i), vv = __ldcs(v4 + i);
    float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ys[4] = {yv.x, yv.y, yv.z, yv.w}, vs[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      int b = map_exact(xs[k], ys[k], m);
      if (b >= 0 && vs[k] == vs[k]) {
        if (AOS) { atomicAdd(sum + 2 * (size_t)b, (double)vs[k]); atomicAdd((uint32_t*)(sum + 2 * (size_t)b + 1), 1u); }
        else { atomicAdd(sum + b, (double)vs[k]); atomicAdd(cnt + b, 1u); }
      }
    }
  }
}

// pure atomic rate from precomputed indices. OP: 0 red.add.u32, 1 red.add.f64, 2 red.max.u64, 3 red.add.u64
template <int OP>
__global__ void __launch_bounds__(512) k_idx(const uint4* __restrict__ idx4, size_t n4, void* canvas) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    uint4 q = __ldcs(idx4 + i);
    uint32_t b[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (OP == 0) atomicAdd((uint32_t*)canvas + b[k], 1u);
      if (OP == 1) atomicAdd((double*)canvas + b[k], 1.0);
      if (OP == 2) atomicMax((unsigned long long*)canvas + b[k], (unsigned long long)(i * 4 + k));
      if (OP == 3) atomicAdd((unsigned long long*)canvas + b[k], 1ull);
    }
  }
}

// shared-memory scatter rate: indices in [0, nb) with nb*4 bytes of dynamic smem.
// OP: 0 atomicAdd u32 (no return), 1 racy LDS+STS, 2 atomicAdd with return used, 3 packed u16 add via u32 atomic
template <int OP>
__global__ void __launch_bounds__(1024) k_smem(const uint4* __restrict__ idx4, size_t n4, uint32_t nb,
                                               uint32_t* __restrict__ out) {
  extern __shared__ uint32_t sh[];
  for (uint32_t j = threadIdx.x; j < nb; j += blockDim.x) sh[j] = 0;
  __syncthreads();
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (; i < n4; i += stride) {
    uint4 q = __ldcs(idx4 + i);
    uint32_t b[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (OP == 0) atomicAdd(sh + b[k], 1u);
      if (OP == 1) { uint32_t t = sh[b[k]]; sh[b[k]] = t + 1; }
      if (OP == 2) acc += atomicAdd(sh + b[k], 1u);
      if (OP == 3) atomicAdd(sh + (b[k] >> 1), (b[k] & 1) ? 0x10000u : 1u);
    }
  }
  __syncthreads();
  uint32_t s = acc;
  for (uint32_t j = threadIdx.x; j < nb; j += blockDim.x) s += sh[j];
  if (s == 0x12345678u) out[blockIdx.x] = s;
}

template <typename F>
static float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  size_t n = (argc > 1) ? strtoull(argv[1], 0, 10) : (size_t)1 << 28;   // 268M points
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  printf("device %s sms=%d smem/blk optin=%zu l2=%d MB\n", p.name, p.multiProcessorCount,
         p.sharedMemPerBlockOptin, p.l2CacheSize >> 20);
  int sms = p.multiProcessorCount;
  float *x, *y, *v; uint32_t* idx; void* canvas; uint32_t* sink;
  CK(cudaMalloc(&x, n * 4)); CK(cudaMalloc(&y, n * 4)); CK(cudaMalloc(&v, n * 4)); CK(cudaMalloc(&idx, n * 4));
  size_t canvas_bytes = (size_t)8192 * 8192 * 16;
  CK(cudaMalloc(&canvas, canvas_bytes)); CK(cudaMalloc(&sink, 4096 * 4));
  CK(cudaMemset(canvas, 0, canvas_bytes));
  size_t n4 = n / 4;

  for (int dist = 0; dist < 2; dist++) {
    if (dist == 0) gen_uniform<<<sms * 8, 512>>>(x, y, v, n, 12345u);
    else gen_cluster<<<sms * 8, 512>>>(x, y, n, 777u);
    CK(cudaDeviceSynchronize());
    const char* dn = dist ? "cluster" : "uniform";
    int Ws[3] = {900, 1920, 8192}, Hs[3] = {525, 1080, 8192};
    for (int c = 0; c < 3; c++) {
      int W = Ws[c], H = Hs[c];
      Map m; m.W = W; m.H = H; m.xmin = 0; m.xmax = 1; m.ymin = 0; m.ymax = 1;
      m.sx = W / 1.0; m.tx = 0; m.sy = H / 1.0; m.ty = 0;
      MapF f; f.sx = (float)m.sx; f.tx = 0; f.sy = (float)m.sy; f.ty = 0; f.xmin = 0; f.xmax = 1; f.ymin = 0; f.ymax = 1;
      f.ex = W * 4e-7f; f.ey = H * 4e-7f;
      for (int bps = 2; bps <= 4; bps += 2) {
        int grid = sms * bps;
        float t;
        if (c == 0 && dist == 0) {
          t = timeit([&] { k_points<4><<<grid, 512>>>((float4*)x, (float4*)y, n4, m, f, (uint32_t*)canvas, sink); });
          printf("[%s %dx%d bps=%d] read-only        : %8.3f ms  %7.1f Gpts/s  %7.1f GB/s\n", dn, W, H, bps, t, n / t * 1e-6, n * 8 / t * 1e-6);
          t = timeit([&] { k_points<1><<<grid, 512>>>((float4*)x, (float4*)y, n4, m, f, (uint32_t*)canvas, sink); });
          printf("[%s %dx%d bps=%d] exact map, no atom: %8.3f ms  %7.1f Gpts/s  %7.1f GB/s\n", dn, W, H, bps, t, n / t * 1e-6, n * 8 / t * 1e-6);
          t = timeit([&] { k_points<3><<<grid, 512>>>((float4*)x, (float4*)y, n4, m, f, (uint32_t*)canvas, sink); });
          printf("[%s %dx%d bps=%d] fast map, no atom : %8.3f ms  %7.1f Gpts/s  %7.1f GB/s\n", dn, W, H, bps, t, n / t * 1e-6, n * 8 / t * 1e-6);
        }
        t = timeit([&] { k_points<0><<<grid, 512>>>((float4*)x, (float4*)y, n4, m, f, (uint32_t*)canvas, sink); });
        printf("[%s %dx%d bps=%d] exact map + REDG  : %8.3f ms  %7.1f Gpts/s  %7.1f GB/s\n", dn, W, H, bps, t, n / t * 1e-6, n * 8 / t * 1e-6);
        t = timeit([&] { k_points<2><<<grid, 512>>>((float4*)x, (float4*)y, n4, m, f, (uint32_t*)canvas, sink); });
        printf("[%s %dx%d bps=%d] fast map + REDG   : %8.3f ms  %7.1f Gpts/s  %7.1f GB/s\n", dn, W, H, bps, t, n / t * 1e-6, n * 8 / t * 1e-6);
        if (dist == 0 || c == 0) {
          t = timeit([&] { k_mean<0><<<grid, 512>>>((float4*)x, (float4*)y, (float4*)v, n4, m, (double*)canvas, (uint32_t*)((char*)canvas + (size_t)W * H * 8)); });
          printf("[%s %dx%d bps=%d] mean SoA f64+u32  : %8.3f ms  %7.1f Gpts/s  %7.1f GB/s(12B)\n", dn, W, H, bps, t, n / t * 1e-6, n * 12 / t * 1e-6);
          if (c < 2) {
            t = timeit([&] { k_mean<1><<<grid, 512>>>((float4*)x, (float4*)y, (float4*)v, n4, m, (double*)canvas, nullptr); });
            printf("[%s %dx%d bps=%d] mean AoS 16B      : %8.3f ms  %7.1f Gpts/s  %7.1f GB/s(12B)\n", dn, W, H, bps, t, n / t * 1e-6, n * 12 / t * 1e-6);
          }
        }
      }
    }
  }
  // pure atomic rates with precomputed random indices
  uint32_t nbs[4] = {900u * 525u, 1920u * 1080u * 16u, 8192u * 8192u, 65536u};
  for (int c = 0; c < 4; c++) {
    gen_idx<<<sms * 8, 512>>>(idx, n, nbs[c], 99u + c); CK(cudaDeviceSynchronize());
    int grid = sms * 4;
    float t;
    t = timeit([&] { k_idx<0><<<grid, 512>>>((uint4*)idx, n4, canvas); });
    printf("[idx nb=%u] red.add.u32 : %8.3f ms %7.1f Gupd/s\n", nbs[c], t, n / t * 1e-6);
    t = timeit([&] { k_idx<1><<<grid, 512>>>((uint4*)idx, n4, canvas); });
    printf("[idx nb=%u] red.add.f64 : %8.3f ms %7.1f Gupd/s\n", nbs[c], t, n / t * 1e-6);
    t = timeit([&] { k_idx<2><<<grid, 512>>>((uint4*)idx, n4, canvas); });
    printf("[idx nb=%u] red.max.u64 : %8.3f ms %7.1f Gupd/s\n", nbs[c], t, n / t * 1e-6);
    t = timeit([&] { k_idx<3><<<grid, 512>>>((uint4*)idx, n4, canvas); });
    printf("[idx nb=%u] red.add.u64 : %8.3f ms %7.1f Gupd/s\n", nbs[c], t, n / t * 1e-6);
  }
  // shared-memory scatter
  uint32_t snb[3] = {51200u, 8192u, 1024u};
  CK(cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 51200 * 4));
  CK(cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 51200 * 4));
  CK(cudaFuncSetAttribute(k_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 51200 * 4));
  CK(cudaFuncSetAttribute(k_smem<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 51200 * 4));
  for (int c = 0; c < 3; c++) {
    gen_idx<<<sms * 8, 512>>>(idx, n, snb[c], 5u + c); CK(cudaDeviceSynchronize());
    for (int th = 512; th <= 1024; th *= 2) {
      float t;
      size_t sb = 51200 * 4;
      t = timeit([&] { k_smem<0><<<sms, th, sb>>>((uint4*)idx, n4, snb[c], sink); });
      printf("[smem nb=%u th=%d] atoms.add (red) : %8.3f ms %7.1f Gupd/s\n", snb[c], th, t, n / t * 1e-6);
      t = timeit([&] { k_smem<2><<<sms, th, sb>>>((uint4*)idx, n4, snb[c], sink); });
      printf("[smem nb=%u th=%d] atoms.add (ret) : %8.3f ms %7.1f Gupd/s\n", snb[c], th, t, n / t * 1e-6);
      t = timeit([&] { k_smem<3><<<sms, th, sb>>>((uint4*)idx, n4, snb[c], sink); });
      printf("[smem nb=%u th=%d] atoms packed u16: %8.3f ms %7.1f Gupd/s\n", snb[c], th, t, n / t * 1e-6);
      t = timeit([&] { k_smem<1><<<sms, th, sb>>>((uint4*)idx, n4, snb[c], sink); });
      printf("[smem nb=%u th=%d] racy LDS+STS    : %8.3f ms %7.1f Gupd/s\n", snb[c], th, t, n / t * 1e-6);
    }
  }
  printf("done\n");
  return 0;
}
