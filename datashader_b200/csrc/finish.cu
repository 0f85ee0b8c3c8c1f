// Canvas-sized passes: accumulator init, key decoding, row gathers, mean/sum finalisation, column bounds.
// These replace make_create (compiler.py:294-301) and each reduction's _finalize (reductions.py).
#include "common.cuh"
#include <limits.h>

template <typename T>
__global__ void k_fill(T* p, T v, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

static int grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  long long cap = (long long)dsb_num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

extern "C" int dsb_init_canvas(int32_t op, void* agg, int64_t ncell, void* stream) {
  if (!agg || ncell < 0) { dsb_set_error("dsb_init_canvas: bad arguments"); return DSB_ERR_ARG; }
  if (ncell == 0) return DSB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int th = 256;
  int g = grid_for(ncell, th);
  switch (op) {
    case DSB_OP_COUNT: cudaMemsetAsync(agg, 0, (size_t)ncell * 4, s); break;
    case DSB_OP_ANY: cudaMemsetAsync(agg, 0, (size_t)ncell, s); break;
    case DSB_OP_SUM: cudaMemsetAsync(agg, 0, (size_t)ncell * 8, s); break;
    case DSB_OP_MAX32: k_fill<int><<<g, th, 0, s>>>((int*)agg, INT_MIN, ncell); break;
    case DSB_OP_MIN32: k_fill<int><<<g, th, 0, s>>>((int*)agg, INT_MAX, ncell); break;
    case DSB_OP_MAX64: case DSB_OP_ARGMAX32: k_fill<long long><<<g, th, 0, s>>>((long long*)agg, LLONG_MIN, ncell); break;
    case DSB_OP_MIN64: case DSB_OP_ARGMIN32: case DSB_OP_MINROW: case DSB_OP_MATCHROW64:
      k_fill<long long><<<g, th, 0, s>>>((long long*)agg, LLONG_MAX, ncell); break;
    case DSB_OP_MAXROW: k_fill<long long><<<g, th, 0, s>>>((long long*)agg, -1LL, ncell); break;
    default: dsb_set_error("dsb_init_canvas: unknown op %d", op); return DSB_ERR_ARG;
  }
  DSB_CUDA_CHECK_LAUNCH("dsb_init_canvas");
  return DSB_OK;
}

__global__ void k_decode_minmax(const void* keys, int op, int dt, double* out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    double r;
    if (op == DSB_OP_MAX32 || op == DSB_OP_MIN32) {
      int k = ((const int*)keys)[i];
      bool empty = (op == DSB_OP_MAX32) ? (k == INT_MIN) : (k == INT_MAX);
      r = empty ? (double)NAN : value_from_key32(k, dt);
    } else {
      long long k = ((const long long*)keys)[i];
      bool empty = (op == DSB_OP_MAX64) ? (k == LLONG_MIN) : (k == LLONG_MAX);
      r = empty ? (double)NAN : f64_from_key64(k);
    }
    out[i] = r;
  }
}

extern "C" int dsb_decode_minmax(const void* keys, int32_t op, int32_t val_dtype, double* out, int64_t ncell, void* stream) {
  if (!keys || !out || op < DSB_OP_MAX32 || op > DSB_OP_MIN64) { dsb_set_error("dsb_decode_minmax: bad arguments"); return DSB_ERR_ARG; }
  if (ncell == 0) return DSB_OK;
  k_decode_minmax<<<grid_for(ncell, 256), 256, 0, (cudaStream_t)stream>>>(keys, op, val_dtype, out, ncell);
  DSB_CUDA_CHECK_LAUNCH("dsb_decode_minmax");
  return DSB_OK;
}

__global__ void k_decode_arg(const long long* packed, int op, int dt, long long row_offset, double* out_sel,
                             long long* out_row, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    long long p = packed[i];
    bool empty = (op == DSB_OP_ARGMAX32) ? (p == LLONG_MIN) : (p == LLONG_MAX);
    int k = (int)(p >> 32);
    uint32_t lo = (uint32_t)(p & 0xffffffffLL);
    uint32_t local = ((op == DSB_OP_ARGMAX32) ? ~lo : lo) - (uint32_t)row_offset;   // mod 2^32
    if (out_sel) out_sel[i] = empty ? (double)NAN : value_from_key32(k, dt);
    if (out_row) out_row[i] = empty ? -1LL : row_offset + (long long)local;
  }
}

extern "C" int dsb_decode_arg(const void* packed, int32_t op, int32_t val_dtype, int64_t row_offset, double* out_sel,
                              int64_t* out_row, int64_t ncell, void* stream) {
  if (!packed || (op != DSB_OP_ARGMAX32 && op != DSB_OP_ARGMIN32)) { dsb_set_error("dsb_decode_arg: bad arguments"); return DSB_ERR_ARG; }
  if (ncell == 0) return DSB_OK;
  k_decode_arg<<<grid_for(ncell, 256), 256, 0, (cudaStream_t)stream>>>((const long long*)packed, op, val_dtype, row_offset,
                                                                         out_sel, (long long*)out_row, ncell);
  DSB_CUDA_CHECK_LAUNCH("dsb_decode_arg");
  return DSB_OK;
}

__global__ void k_gather_rows(const long long* rows, long long row_offset, long long nrows, const void* lookup, int dt,
                              double* out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    long long r = rows[i];
    if (r < 0 || r == LLONG_MAX) { out[i] = (double)NAN; continue; }
    long long local = r - row_offset;
    if (local < 0 || local >= nrows) continue;     // another shard owns this row
    out[i] = load_f64(lookup, dt, local);
  }
}

extern "C" int dsb_gather_rows(const int64_t* rows, int64_t row_offset, int64_t n, const void* lookup, int32_t lookup_dtype,
                               double* out, int64_t ncell, void* stream) {
  if (!rows || !out || (!lookup && n > 0)) { dsb_set_error("dsb_gather_rows: bad arguments"); return DSB_ERR_ARG; }
  if (ncell == 0) return DSB_OK;
  k_gather_rows<<<grid_for(ncell, 256), 256, 0, (cudaStream_t)stream>>>((const long long*)rows, row_offset, n, lookup,
                                                                          lookup_dtype, out, ncell);
  DSB_CUDA_CHECK_LAUNCH("dsb_gather_rows");
  return DSB_OK;
}

__global__ void k_finish_minrow(long long* rows, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) if (rows[i] == LLONG_MAX) rows[i] = -1;
}

extern "C" int dsb_finish_minrow(int64_t* rows, int64_t ncell, void* stream) {
  if (!rows) { dsb_set_error("dsb_finish_minrow: null canvas"); return DSB_ERR_ARG; }
  if (ncell == 0) return DSB_OK;
  k_finish_minrow<<<grid_for(ncell, 256), 256, 0, (cudaStream_t)stream>>>((long long*)rows, ncell);
  DSB_CUDA_CHECK_LAUNCH("dsb_finish_minrow");
  return DSB_OK;
}

__global__ void k_finalize_mean(const double* sum, const unsigned int* cnt, double* out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    unsigned int c = cnt[i];
    out[i] = c > 0 ? __ddiv_rn(sum[i], (double)c) : (double)NAN;   // reductions.py:1292-1297
  }
}

extern "C" int dsb_finalize_mean(const double* sum, const void* count_u32, double* out, int64_t ncell, void* stream) {
  if (!sum || !count_u32 || !out) { dsb_set_error("dsb_finalize_mean: null pointer"); return DSB_ERR_ARG; }
  if (ncell == 0) return DSB_OK;
  k_finalize_mean<<<grid_for(ncell, 256), 256, 0, (cudaStream_t)stream>>>(sum, (const unsigned int*)count_u32, out, ncell);
  DSB_CUDA_CHECK_LAUNCH("dsb_finalize_mean");
  return DSB_OK;
}

__global__ void k_finalize_sum(const double* sum, const uint8_t* mask, double* out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = mask[i] ? sum[i] : (double)NAN;   // reductions.py:1091-1096
}

extern "C" int dsb_finalize_sum(const double* sum, const uint8_t* mask, double* out, int64_t ncell, void* stream) {
  if (!sum || !mask || !out) { dsb_set_error("dsb_finalize_sum: null pointer"); return DSB_ERR_ARG; }
  if (ncell == 0) return DSB_OK;
  k_finalize_sum<<<grid_for(ncell, 256), 256, 0, (cudaStream_t)stream>>>(sum, mask, out, ncell);
  DSB_CUDA_CHECK_LAUNCH("dsb_finalize_sum");
  return DSB_OK;
}

__global__ void k_finalize_sum_counted(const double* sum, const uint32_t* count, double* out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = count[i] ? sum[i] : (double)NAN;
}

extern "C" int dsb_finalize_sum_counted(const double* sum, const void* count_u32, double* out, int64_t ncell, void* stream) {
  if (!sum || !count_u32 || !out) { dsb_set_error("dsb_finalize_sum_counted: null pointer"); return DSB_ERR_ARG; }
  if (ncell == 0) return DSB_OK;
  k_finalize_sum_counted<<<grid_for(ncell, 256), 256, 0, (cudaStream_t)stream>>>(sum, (const uint32_t*)count_u32, out, ncell);
  DSB_CUDA_CHECK_LAUNCH("dsb_finalize_sum_counted");
  return DSB_OK;
}

// ---- column bounds: Glyph._compute_bounds_numba (glyphs/glyph.py:66-78) ----------------------------
__device__ __forceinline__ void atomic_min_f64(double* addr, double v) {
  // monotone CAS; NaNs never reach here
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a;
  while (v < __longlong_as_double(old)) {
    unsigned long long assumed = old;
    old = atomicCAS(a, assumed, __double_as_longlong(v));
    if (old == assumed) break;
  }
}
__device__ __forceinline__ void atomic_max_f64(double* addr, double v) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a;
  while (v > __longlong_as_double(old)) {
    unsigned long long assumed = old;
    old = atomicCAS(a, assumed, __double_as_longlong(v));
    if (old == assumed) break;
  }
}

__global__ void k_bounds_init(double* out) { out[0] = INFINITY; out[1] = -INFINITY; }

__global__ void __launch_bounds__(256) k_bounds(const void* col, int dt, long long n, double* out) {
  double mn = INFINITY, mx = -INFINITY;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    double v = load_f64(col, dt, i);
    if (v == v) { mn = fmin(mn, v); mx = fmax(mx, v); }
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ double smn[8], smx[8];
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { smn[w] = mn; smx[w] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; k++) { mn = fmin(mn, smn[k]); mx = fmax(mx, smx[k]); }
    if (mn <= mx) { atomic_min_f64(out, mn); atomic_max_f64(out + 1, mx); }
  }
}

extern "C" int dsb_bounds(const void* col, int32_t dtype, int64_t n, double* out_minmax, void* stream) {
  if (!out_minmax || (!col && n > 0) || dtype == DSB_NONE) { dsb_set_error("dsb_bounds: bad arguments"); return DSB_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  k_bounds_init<<<1, 1, 0, s>>>(out_minmax);
  if (n > 0) k_bounds<<<grid_for(n, 256 * 8), 256, 0, s>>>(col, dtype, n, out_minmax);
  DSB_CUDA_CHECK_LAUNCH("dsb_bounds");
  return DSB_OK;
}
