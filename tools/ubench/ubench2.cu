// Micro-benchmarks, part 2: cluster-distributed shared memory (DSMEM) atomics, f64 shared atomics,
// match.any, hybrid SMEM-band + global RED, TMA bulk-reduce flush.  Not part of the product.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o ubench2 ubench2.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
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

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void red_cluster_add_u32(uint32_t local_addr, uint32_t cta, uint32_t v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(cta));
  asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" :: "r"(remote), "r"(v) : "memory");
}

// canvas of CL*nb_local bins distributed over the cluster's shared memories
__global__ void __launch_bounds__(1024) k_dsmem(const uint4* __restrict__ idx4, size_t n4, uint32_t nb_local,
                                                uint32_t* __restrict__ out) {
  extern __shared__ uint32_t sh[];
  cg::cluster_group cl = cg::this_cluster();
  for (uint32_t j = threadIdx.x; j < nb_local; j += blockDim.x) sh[j] = 0;
  cl.sync();
  uint32_t base = smem_u32(sh);
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    uint4 q = __ldcs(idx4 + i);
    uint32_t b[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t cta = b[k] / nb_local, loc = b[k] - cta * nb_local;
      red_cluster_add_u32(base + loc * 4, cta, 1u);
    }
  }
  cl.sync();
  uint32_t s = 0;
  for (uint32_t j = threadIdx.x; j < nb_local; j += blockDim.x) s += sh[j];
  atomicAdd(out, s);
}

// OP 0: f64 atomicAdd in shared; OP 1: u64 atomicAdd in shared; OP 2: match.any only; OP 3: u32 shared atomicMax
template <int OP>
__global__ void __launch_bounds__(1024) k_smem2(const uint4* __restrict__ idx4, size_t n4, uint32_t nb,
                                                uint32_t* __restrict__ out) {
  extern __shared__ double shd[];
  unsigned long long* shu = (unsigned long long*)shd;
  uint32_t* sh32 = (uint32_t*)shd;
  for (uint32_t j = threadIdx.x; j < nb; j += blockDim.x) shd[j] = 0;
  __syncthreads();
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (; i < n4; i += stride) {
    uint4 q = __ldcs(idx4 + i);
    uint32_t b[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (OP == 0) atomicAdd(shd + b[k], 1.25);
      if (OP == 1) atomicAdd(shu + b[k], 3ull);
      if (OP == 2) acc += __match_any_sync(0xffffffffu, b[k]);
      if (OP == 3) atomicMax(sh32 + b[k], (uint32_t)i);
    }
  }
  __syncthreads();
  double s = acc;
  for (uint32_t j = threadIdx.x; j < nb; j += blockDim.x) s += shd[j];
  if (s == 1234.5) out[blockIdx.x] = 1;
}

// hybrid: each CTA keeps band [lo, lo+nb_band) of the canvas in shared memory as packed counters of BITS bits
// (no overflow handling here - throughput probe only), everything else goes to global REDs.
template <int BITS>
__global__ void __launch_bounds__(1024) k_hybrid(const uint4* __restrict__ idx4, size_t n4, uint32_t nbins,
                                                 uint32_t nb_band, uint32_t* __restrict__ canvas) {
  extern __shared__ uint32_t sh[];
  const uint32_t per = 32 / BITS;
  uint32_t words = (nb_band + per - 1) / per;
  for (uint32_t j = threadIdx.x; j < words; j += blockDim.x) sh[j] = 0;
  __syncthreads();
  uint32_t nbands = (nbins + nb_band - 1) / nb_band;
  uint32_t lo = (blockIdx.x % nbands) * nb_band;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    uint4 q = __ldcs(idx4 + i);
    uint32_t b[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t loc = b[k] - lo;
      if (loc < nb_band) {
        if (BITS == 32) atomicAdd(sh + loc, 1u);
        else atomicAdd(sh + loc / per, 1u << (BITS * (loc % per)));
      } else {
        atomicAdd(canvas + b[k], 1u);
      }
    }
  }
  __syncthreads();
  for (uint32_t j = threadIdx.x; j < nb_band; j += blockDim.x) {
    uint32_t v = (BITS == 32) ? sh[j] : ((sh[j / per] >> (BITS * (j % per))) & ((1u << BITS) - 1u));
    if (v && lo + j < nbins) atomicAdd(canvas + lo + j, v);
  }
}

// TMA bulk reduce: flush `bytes` of shared memory into global with add.u32, `reps` times
__global__ void __launch_bounds__(256) k_bulkred(uint32_t* __restrict__ canvas, uint32_t bytes, int reps) {
  extern __shared__ __align__(128) uint32_t shb[];
  for (uint32_t j = threadIdx.x; j < bytes / 4; j += blockDim.x) shb[j] = 1;
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int r = 0; r < reps; r++) {
      uint32_t* dst = canvas + (size_t)((blockIdx.x + r) % 16) * (bytes / 4);
      asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u32 [%0], [%1], %2;"
                   :: "l"(dst), "r"(smem_u32(shb)), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
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
  size_t n = (argc > 1) ? strtoull(argv[1], 0, 10) : (size_t)1 << 28;
  CK(cudaSetDevice(0));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("device %s sms=%d\n", p.name, sms);
  uint32_t* idx; uint32_t* canvas; uint32_t* sink;
  CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&canvas, (size_t)64 << 20)); CK(cudaMalloc(&sink, 4096 * 4));
  CK(cudaMemset(canvas, 0, (size_t)64 << 20)); CK(cudaMemset(sink, 0, 4096 * 4));
  size_t n4 = n / 4;

  // ---- DSMEM distributed canvas
  for (int CL = 2; CL <= 16; CL *= 2) {
    uint32_t nb_local = 29532;  // 472500/16 rounded up: 118 KB of u32 per CTA
    gen_idx<<<sms * 8, 512>>>(idx, n, nb_local * CL, 11u + CL); CK(cudaDeviceSynchronize());
    size_t sb = (size_t)nb_local * 4;
    CK(cudaFuncSetAttribute(k_dsmem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
    if (CL > 8) CK(cudaFuncSetAttribute(k_dsmem, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = sb;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(CL);
    int maxcl = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&maxcl, k_dsmem, &cfg);
    if (e != cudaSuccess) { printf("[dsmem CL=%d] occupancy query failed: %s\n", CL, cudaGetErrorString(e)); cudaGetLastError(); continue; }
    cfg.gridDim = dim3(maxcl * CL);
    CK(cudaMemset(sink, 0, 4));
    float t = timeit([&] { CK(cudaLaunchKernelEx(&cfg, k_dsmem, (const uint4*)idx, n4, nb_local, sink)); });
    uint32_t tot; CK(cudaMemcpy(&tot, sink, 4, cudaMemcpyDeviceToHost));
    printf("[dsmem CL=%2d] max clusters=%d (ctas=%d): %8.3f ms %7.1f Gupd/s  (check %u vs %u)\n", CL, maxcl, maxcl * CL, t,
           n / t * 1e-6, tot, (uint32_t)((n / 4 * 4) * 6));
  }

  // ---- shared f64 / u64 atomics, match.any
  {
    uint32_t nb = 25000;  // 200 KB of 8-byte bins
    gen_idx<<<sms * 8, 512>>>(idx, n, nb, 3u); CK(cudaDeviceSynchronize());
    size_t sb = (size_t)nb * 8;
    CK(cudaFuncSetAttribute(k_smem2<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
    CK(cudaFuncSetAttribute(k_smem2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
    CK(cudaFuncSetAttribute(k_smem2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
    CK(cudaFuncSetAttribute(k_smem2<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
    float t;
    t = timeit([&] { k_smem2<0><<<sms, 1024, sb>>>((uint4*)idx, n4, nb, sink); });
    printf("[smem f64 atomicAdd] %8.3f ms %7.1f Gupd/s\n", t, n / t * 1e-6);
    t = timeit([&] { k_smem2<1><<<sms, 1024, sb>>>((uint4*)idx, n4, nb, sink); });
    printf("[smem u64 atomicAdd] %8.3f ms %7.1f Gupd/s\n", t, n / t * 1e-6);
    t = timeit([&] { k_smem2<3><<<sms, 1024, sb>>>((uint4*)idx, n4, nb, sink); });
    printf("[smem u32 atomicMax] %8.3f ms %7.1f Gupd/s\n", t, n / t * 1e-6);
    t = timeit([&] { k_smem2<2><<<sms, 1024, sb>>>((uint4*)idx, n4, nb, sink); });
    printf("[match.any only    ] %8.3f ms %7.1f Gupd/s\n", t, n / t * 1e-6);
  }

  // ---- hybrid band in SMEM + global REDs, canvas 900x525
  {
    uint32_t nbins = 900u * 525u;
    gen_idx<<<sms * 8, 512>>>(idx, n, nbins, 21u); CK(cudaDeviceSynchronize());
    size_t sb = 56000 * 4;  // 224 KB
    CK(cudaFuncSetAttribute(k_hybrid<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
    CK(cudaFuncSetAttribute(k_hybrid<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
    CK(cudaFuncSetAttribute(k_hybrid<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
    float t;
    t = timeit([&] { k_hybrid<32><<<sms, 1024, sb>>>((uint4*)idx, n4, nbins, 56000u, canvas); });
    printf("[hybrid u32 band=56000 ] %8.3f ms %7.1f Gupd/s\n", t, n / t * 1e-6);
    t = timeit([&] { k_hybrid<16><<<sms, 1024, sb>>>((uint4*)idx, n4, nbins, 112000u, canvas); });
    printf("[hybrid u16 band=112000] %8.3f ms %7.1f Gupd/s\n", t, n / t * 1e-6);
    t = timeit([&] { k_hybrid<8><<<sms, 1024, sb>>>((uint4*)idx, n4, nbins, 224000u, canvas); });
    printf("[hybrid u8  band=224000] %8.3f ms %7.1f Gupd/s\n", t, n / t * 1e-6);
    // same with a canvas small enough to be fully resident as u8: 448x500
    gen_idx<<<sms * 8, 512>>>(idx, n, 224000u, 22u); CK(cudaDeviceSynchronize());
    t = timeit([&] { k_hybrid<8><<<sms, 1024, sb>>>((uint4*)idx, n4, 224000u, 224000u, canvas); });
    printf("[full-resident u8 224000 bins] %8.3f ms %7.1f Gupd/s\n", t, n / t * 1e-6);
    gen_idx<<<sms * 8, 512>>>(idx, n, 56000u, 23u); CK(cudaDeviceSynchronize());
    t = timeit([&] { k_hybrid<32><<<sms, 1024, sb>>>((uint4*)idx, n4, 56000u, 56000u, canvas); });
    printf("[full-resident u32 56000 bins] %8.3f ms %7.1f Gupd/s\n", t, n / t * 1e-6);
  }

  // ---- TMA bulk reduce flush
  {
    uint32_t bytes = 128 * 1024; int reps = 64;
    CK(cudaFuncSetAttribute(k_bulkred, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    float t = timeit([&] { k_bulkred<<<sms, 256, bytes>>>(canvas, bytes, reps); });
    double total = (double)bytes * reps * sms;
    printf("[bulk reduce add.u32] %8.3f ms  %7.1f GB/s aggregate, %6.2f us per 128KB flush per SM\n", t, total / t * 1e-6,
           t * 1e3 / reps);
  }
  printf("done\n");
  return 0;
}
