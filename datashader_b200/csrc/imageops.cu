// Post-shade image operations: the composite operators (datashader/composite.py), tf.spread's three stencil kernels
// and dynspread's density heuristic (transfer_functions/__init__.py:748-1051), tf.stack / set_background (:115-145,
// :748-768).
//
// The reference's spread is a serial SCATTER: every source pixel, in raster order, composites itself onto the
// (2 px + 1)^2 window around it, so for the non-commutative operators ('over', 'saturate', float 'add' in its
// rounding) the result at a pixel depends on the order in which the sources reached it.  Here each OUTPUT pixel is
// one thread that GATHERS its sources in exactly that order (y ascending, then x ascending) and folds them in a
// register - bit-identical results, no atomics, one coalesced pass.  All f64 arithmetic is unfused (_rn intrinsics)
// because numba/LLVM does not contract the reference's expressions.
#include "common.cuh"

__device__ __forceinline__ double m64(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double a64(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double d64(double a, double b) { return __ddiv_rn(a, b); }

// composite.py:43-49: min(255, uint32(c * 255)) per channel
__device__ __forceinline__ uint32_t pack_channel(double c) {
  const double v = m64(c, 255.0);
  uint32_t u = (v != v) ? 0u : (v >= 4294967295.0 ? 0xffffffffu : (v <= 0.0 ? 0u : (uint32_t)__double2ull_rz(v)));
  return u < 255u ? u : 255u;
}

// composite.py:72-125: op(src, dst)
__device__ __forceinline__ uint32_t composite_px(int how, uint32_t src, uint32_t dst) {
  if (how == DSB_COMP_SOURCE) return (src & 0xff000000u) ? src : dst;
  const double sr = d64((double)(src & 255u), 255.0), sg = d64((double)((src >> 8) & 255u), 255.0);
  const double sb = d64((double)((src >> 16) & 255u), 255.0), sa = d64((double)((src >> 24) & 255u), 255.0);
  const double dr = d64((double)(dst & 255u), 255.0), dg = d64((double)((dst >> 8) & 255u), 255.0);
  const double db = d64((double)((dst >> 16) & 255u), 255.0), da = d64((double)((dst >> 24) & 255u), 255.0);
  double a, r, g, b;
  if (how == DSB_COMP_OVER) {
    const double factor = a64(1.0, -sa);
    a = a64(sa, m64(da, factor));
    if (a == 0.0) return 0u;
    r = d64(a64(m64(sr, sa), m64(m64(dr, da), factor)), a);
    g = d64(a64(m64(sg, sa), m64(m64(dg, da), factor)), a);
    b = d64(a64(m64(sb, sa), m64(m64(db, da), factor)), a);
  } else if (how == DSB_COMP_ADD) {
    a = fmin(1.0, a64(sa, da));
    if (a == 0.0) return 0u;
    r = d64(a64(m64(sr, sa), m64(dr, da)), a);
    g = d64(a64(m64(sg, sa), m64(dg, da)), a);
    b = d64(a64(m64(sb, sa), m64(db, da)), a);
  } else {   // saturate
    a = fmin(1.0, a64(sa, da));
    if (a == 0.0) return 0u;
    const double factor = fmin(sa, a64(1.0, -da));
    r = d64(a64(m64(factor, sr), m64(dr, da)), a);
    g = d64(a64(m64(factor, sg), m64(dg, da)), a);
    b = d64(a64(m64(factor, sb), m64(db, da)), a);
  }
  return (pack_channel(a) << 24) | (pack_channel(b) << 16) | (pack_channel(g) << 8) | pack_channel(r);
}

__global__ void k_composite(const uint32_t* __restrict__ src, const uint32_t* __restrict__ dst, uint32_t dst_scalar, long long n,
                            int how, uint32_t* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = composite_px(how, src[i], dst ? dst[i] : dst_scalar);
}

extern "C" int dsb_composite(const uint32_t* src, const uint32_t* dst, uint32_t dst_scalar, int64_t n, int32_t how,
                             uint32_t* out, void* stream) {
  if (!src || !out) { dsb_set_error("dsb_composite: null pointer"); return DSB_ERR_ARG; }
  if (how < DSB_COMP_OVER || how > DSB_COMP_SOURCE) { dsb_set_error("dsb_composite: unknown operator %d", how); return DSB_ERR_ARG; }
  if (n == 0) return DSB_OK;
  long long g = (n + 255) / 256, cap = (long long)dsb_num_sms() * 8;
  k_composite<<<(int)(g < cap ? g : cap), 256, 0, (cudaStream_t)stream>>>(src, dst, dst_scalar, n, how, out);
  DSB_CUDA_CHECK_LAUNCH("dsb_composite");
  return DSB_OK;
}

// _build_spread_kernel, transfer_functions/__init__.py:880-915 (is_image=True)
__global__ void k_spread_image(const uint32_t* __restrict__ img, int H, int W, const uint8_t* __restrict__ mask, int w, int how,
                               uint32_t* __restrict__ out) {
  const int extra = w / 2;
  const long long n = (long long)H * W;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    const int py = (int)(p / W), px = (int)(p - (long long)py * W);
    uint32_t acc = 0;
    const int y0 = max(0, py - extra), y1 = min(H - 1, py + extra), x0 = max(0, px - extra), x1 = min(W - 1, px + extra);
    for (int y = y0; y <= y1; y++) {
      const int i = py + extra - y;
      for (int x = x0; x <= x1; x++) {
        if (!mask[i * w + (px + extra - x)]) continue;
        const uint32_t el = img[(long long)y * W + x];
        if (!((el >> 24) & 255u)) continue;          // transparent sources are skipped
        acc = (acc == 0u) ? el : composite_px(how, el, acc);
      }
    }
    out[p] = acc;
  }
}

extern "C" int dsb_spread_image(const uint32_t* img, int32_t H, int32_t W, const uint8_t* mask, int32_t w, int32_t how,
                                uint32_t* out, void* stream) {
  if (!img || !mask || !out || H <= 0 || W <= 0) { dsb_set_error("dsb_spread_image: bad arguments"); return DSB_ERR_ARG; }
  if (w < 1 || !(w & 1)) { dsb_set_error("dsb_spread_image: mask side must be odd"); return DSB_ERR_ARG; }
  if (how < DSB_COMP_OVER || how > DSB_COMP_SOURCE) { dsb_set_error("dsb_spread_image: unknown operator %d", how); return DSB_ERR_ARG; }
  long long n = (long long)H * W, g = (n + 127) / 128, cap = (long long)dsb_num_sms() * 16;
  k_spread_image<<<(int)(g < cap ? g : cap), 128, 0, (cudaStream_t)stream>>>(img, H, W, mask, w, how, out);
  DSB_CUDA_CHECK_LAUNCH("dsb_spread_image");
  return DSB_OK;
}

// composite.py:150-168 on scalars: op(el, out)
template <typename T>
__device__ __forceinline__ T arr_op(int how, T el, T o) {
  switch (how) {
    case DSB_ARR_ADD: return el + o;
    case DSB_ARR_MAX: return (o > el) ? o : el;      // Python max([src, dst]): dst only if strictly greater
    case DSB_ARR_MIN: return (o < el) ? o : el;
    default: return el ? el : o;                     // source_arr
  }
}

// MODE 0: float kernel (_build_float_kernel :852-877, NaN = empty); 1: int kernel (:825-849); 2: uint32 kernel with
// ignore_zeros.  Layer c of an [H, W, C] array is spread on its own (spread(): np.dstack over the categories).
template <typename T, int MODE>
__global__ void k_spread_array(const T* __restrict__ arr, int H, int W, int C, const uint8_t* __restrict__ mask, int w, int how,
                               T* __restrict__ out) {
  const int extra = w / 2;
  const long long n = (long long)H * W * C;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(p % C);
    const long long pix = p / C;
    const int py = (int)(pix / W), px = (int)(pix - (long long)py * W);
    T acc = MODE == 0 ? (T)NAN : (T)0;
    const int y0 = max(0, py - extra), y1 = min(H - 1, py + extra), x0 = max(0, px - extra), x1 = min(W - 1, px + extra);
    for (int y = y0; y <= y1; y++) {
      const int i = py + extra - y;
      for (int x = x0; x <= x1; x++) {
        if (!mask[i * w + (px + extra - x)]) continue;
        const T el = arr[((long long)y * W + x) * C + c];
        if (MODE == 0) {
          if (el != el) continue;
          acc = (acc != acc) ? el : arr_op<T>(how, el, acc);
        } else if (MODE == 2) {
          if (el == (T)0) continue;
          acc = (acc == (T)0) ? el : arr_op<T>(how, el, acc);
        } else {
          acc = arr_op<T>(how, el, acc);
        }
      }
    }
    out[p] = acc;
  }
}

template <typename T, int MODE>
static void launch_spread_array(const void* arr, int H, int W, int C, const uint8_t* mask, int w, int how, void* out, cudaStream_t s) {
  long long n = (long long)H * W * C, g = (n + 127) / 128, cap = (long long)dsb_num_sms() * 16;
  k_spread_array<T, MODE><<<(int)(g < cap ? g : cap), 128, 0, s>>>((const T*)arr, H, W, C, mask, w, how, (T*)out);
}

extern "C" int dsb_spread_array(const void* arr, int32_t dtype, int32_t H, int32_t W, int32_t C, const uint8_t* mask, int32_t w,
                                int32_t how, void* out, void* stream) {
  if (!arr || !mask || !out || H <= 0 || W <= 0 || C <= 0) { dsb_set_error("dsb_spread_array: bad arguments"); return DSB_ERR_ARG; }
  if (w < 1 || !(w & 1)) { dsb_set_error("dsb_spread_array: mask side must be odd"); return DSB_ERR_ARG; }
  if (how < DSB_ARR_ADD || how > DSB_ARR_SOURCE) { dsb_set_error("dsb_spread_array: unknown operator %d", how); return DSB_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  switch (dtype) {
    case DSB_F32: launch_spread_array<float, 0>(arr, H, W, C, mask, w, how, out, s); break;
    case DSB_F64: launch_spread_array<double, 0>(arr, H, W, C, mask, w, how, out, s); break;
    case DSB_I32: launch_spread_array<int32_t, 1>(arr, H, W, C, mask, w, how, out, s); break;
    case DSB_I64: launch_spread_array<long long, 1>(arr, H, W, C, mask, w, how, out, s); break;
    case DSB_U32: launch_spread_array<uint32_t, 2>(arr, H, W, C, mask, w, how, out, s); break;
    default: dsb_set_error("dsb_spread_array: dtype must be f32, f64, i32, i64 or u32"); return DSB_ERR_ARG;
  }
  DSB_CUDA_CHECK_LAUNCH("dsb_spread_array");
  return DSB_OK;
}

// _rgb_density / _array_density, transfer_functions/__init__.py:1004-1051.  KIND 0: image (alpha != 0), 1: float (not NaN),
// 2: integer (!= 0).  out2[0] = non-empty pixels, out2[1] = those with another non-empty pixel within px.
template <typename T, int KIND>
__device__ __forceinline__ bool occupied(T v) {
  if (KIND == 0) return (((uint32_t)v >> 24) & 255u) != 0u;
  if (KIND == 1) return v == v;
  return v != (T)0;
}

template <typename T, int KIND>
__global__ void k_density(const T* __restrict__ arr, int H, int W, int px, unsigned long long* __restrict__ out2) {
  const long long n = (long long)H * W;
  unsigned long long cnt = 0, has = 0;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    if (!occupied<T, KIND>(arr[p])) continue;
    cnt++;
    const int y = (int)(p / W), x = (int)(p - (long long)y * W);
    int neighbors = 0;
    for (int i = max(0, y - px); i < min(y + px + 1, H) && neighbors < 2; i++)
      for (int j = max(0, x - px); j < min(x + px + 1, W) && neighbors < 2; j++)
        neighbors += occupied<T, KIND>(arr[(long long)i * W + j]) ? 1 : 0;
    if (neighbors > 1) has++;
  }
  for (int o = 16; o > 0; o >>= 1) { cnt += __shfl_down_sync(0xffffffffu, cnt, o); has += __shfl_down_sync(0xffffffffu, has, o); }
  if ((threadIdx.x & 31) == 0) { if (cnt) atomicAdd(out2, cnt); if (has) atomicAdd(out2 + 1, has); }
}

extern "C" int dsb_density(const void* arr, int32_t dtype, int32_t is_image, int32_t H, int32_t W, int32_t px, uint64_t* out2,
                           void* stream) {
  if (!arr || !out2 || H <= 0 || W <= 0 || px < 0) { dsb_set_error("dsb_density: bad arguments"); return DSB_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(out2, 0, 16, s);
  long long n = (long long)H * W, g = (n + 127) / 128, cap = (long long)dsb_num_sms() * 16;
  const int grid = (int)(g < cap ? g : cap);
  unsigned long long* o = (unsigned long long*)out2;
  if (is_image) {
    if (dtype != DSB_U32) { dsb_set_error("dsb_density: images are uint32"); return DSB_ERR_ARG; }
    k_density<uint32_t, 0><<<grid, 128, 0, s>>>((const uint32_t*)arr, H, W, px, o);
  } else {
    switch (dtype) {
      case DSB_F32: k_density<float, 1><<<grid, 128, 0, s>>>((const float*)arr, H, W, px, o); break;
      case DSB_F64: k_density<double, 1><<<grid, 128, 0, s>>>((const double*)arr, H, W, px, o); break;
      case DSB_I32: k_density<int32_t, 2><<<grid, 128, 0, s>>>((const int32_t*)arr, H, W, px, o); break;
      case DSB_I64: k_density<long long, 2><<<grid, 128, 0, s>>>((const long long*)arr, H, W, px, o); break;
      case DSB_U32: k_density<uint32_t, 2><<<grid, 128, 0, s>>>((const uint32_t*)arr, H, W, px, o); break;
      default: dsb_set_error("dsb_density: dtype must be f32, f64, i32, i64 or u32"); return DSB_ERR_ARG;
    }
  }
  DSB_CUDA_CHECK_LAUNCH("dsb_density");
  return DSB_OK;
}
