// Shared device helpers for libdsb200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/dsb200.h"

#define DSB_SM_COUNT_FALLBACK 148

extern "C" void dsb_set_error(const char* fmt, ...);
int dsb_num_sms();
void dsb_count_launch();
void dsb_note_kernel(const char* fmt, ...);

#define DSB_CUDA_CHECK_LAUNCH(what)                                              \
  do {                                                                           \
    cudaError_t e_ = cudaGetLastError();                                         \
    if (e_ != cudaSuccess) {                                                     \
      dsb_set_error("%s: %s", what, cudaGetErrorString(e_));                    \
      return DSB_ERR_CUDA;                                                       \
    }                                                                            \
    dsb_count_launch();                                                          \
  } while (0)

// ---- order-preserving keys ------------------------------------------------------------------
// Signed keys so that the canvases are plain int32/int64 tensors that NCCL max/min understand.
// float bits -> sortable signed int: positive floats keep their bits, negative floats flip the
// magnitude bits.  key(NaN) is never produced (NaNs are skipped), which frees INT_MIN / INT_MAX as
// the "empty" sentinels of max / min canvases.
__device__ __forceinline__ int32_t key32_from_f32(float f) {
  // -0.0 is folded onto +0.0 first: the reference compares with < / >, for which the two zeros tie
  int32_t b = __float_as_int(f + 0.0f);
  return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float f32_from_key32(int32_t k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }
__device__ __forceinline__ int64_t key64_from_f64(double d) {
  int64_t b = __double_as_longlong(d + 0.0);
  return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double f64_from_key64(int64_t k) {
  return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL));
}

// ---- column loads ---------------------------------------------------------------------------
__device__ __forceinline__ double load_f64(const void* p, int dt, int64_t i) {
  switch (dt) {
    case DSB_F32: return (double)__ldg((const float*)p + i);
    case DSB_F64: return __ldg((const double*)p + i);
    case DSB_I8: return (double)__ldg((const int8_t*)p + i);
    case DSB_U8: return (double)__ldg((const uint8_t*)p + i);
    case DSB_I16: return (double)__ldg((const int16_t*)p + i);
    case DSB_U16: return (double)__ldg((const uint16_t*)p + i);
    case DSB_I32: return (double)__ldg((const int32_t*)p + i);
    case DSB_U32: return (double)__ldg((const uint32_t*)p + i);
    case DSB_I64: return (double)__ldg((const long long*)p + i);
    case DSB_U64: return (double)__ldg((const unsigned long long*)p + i);
  }
  return 0.0;
}
__device__ __forceinline__ bool is_float_dtype(int dt) { return dt == DSB_F32 || dt == DSB_F64; }

// NaN test on the column's own type (isnull, utils.py:598-603); integer columns are never null.
__device__ __forceinline__ bool col_isnan(const void* p, int dt, int64_t i) {
  if (dt == DSB_F32) { float f = __ldg((const float*)p + i); return f != f; }
  if (dt == DSB_F64) { double d = __ldg((const double*)p + i); return d != d; }
  return false;
}

__device__ __forceinline__ bool is_negzero(float f) { return __float_as_uint(f) == 0x80000000u; }
__device__ __forceinline__ bool is_negzero(double d) { return (unsigned long long)__double_as_longlong(d) == 0x8000000000000000ULL; }

// key32 of a <=32-bit value column; *isnan set for float NaN, *negzero for a float -0.0 (see DSB_NOTE_NEGZERO)
__device__ __forceinline__ int32_t load_key32(const void* p, int dt, int64_t i, bool* nan, bool* negzero) {
  *nan = false;
  *negzero = false;
  switch (dt) {
    case DSB_F32: { float f = __ldg((const float*)p + i); *nan = (f != f); *negzero = is_negzero(f); return key32_from_f32(f); }
    case DSB_I8: return (int32_t)__ldg((const int8_t*)p + i);
    case DSB_U8: return (int32_t)__ldg((const uint8_t*)p + i);
    case DSB_I16: return (int32_t)__ldg((const int16_t*)p + i);
    case DSB_U16: return (int32_t)__ldg((const uint16_t*)p + i);
    case DSB_I32: return __ldg((const int32_t*)p + i);
    case DSB_U32: return (int32_t)(__ldg((const uint32_t*)p + i) ^ 0x80000000u);
  }
  return 0;
}
__device__ __forceinline__ double value_from_key32(int32_t k, int dt) {
  switch (dt) {
    case DSB_F32: return (double)f32_from_key32(k);
    case DSB_U32: return (double)((uint32_t)k ^ 0x80000000u);
    default: return (double)k;
  }
}
__device__ __forceinline__ bool dtype_is_key32(int dt) {
  return dt == DSB_F32 || (dt >= DSB_I8 && dt <= DSB_U32);
}

// ---- the reference's pixel mapping ----------------------------------------------------------
// glyphs/points.py:193-203.  Bounds test in f64, UNFUSED multiply then add (numba/LLVM emits no
// FMA for x*sx+tx; nvcc would contract it, hence the explicit _rn intrinsics), truncating cast,
// and the upper-edge fold.  LogAxis.mapper = log10(float(val)) (core.py:129-132): under numba a
// float32 coordinate stays float32 through log10 and is widened afterwards.
// float32 coordinates: the C library's log10f of the reference's host, instruction for instruction (glibc 2.39, the FMA
// build of its logf) - csrc/log10f_glibc.h, proven equal to the library for every positive finite float by
// oracle/log10f_check.c.  CUDA's log10f is a different algorithm: last-bit differences moved points that sit on a pixel
// boundary by one pixel (round 1).  float64 coordinates still go through CUDA's log10 (not bit-identical to glibc's).
#define DSB_LG_FN __device__ __forceinline__
#define DSB_LG_FMA(a, b, c) fma((a), (b), (c))
#define DSB_LG_FMULF(a, b) __fmul_rn((a), (b))
#define DSB_LG_FADDF(a, b) __fadd_rn((a), (b))
#define DSB_LG_ASUINT(f) __float_as_uint(f)
#define DSB_LG_ASFLOAT(u) __uint_as_float(u)
#define DSB_LG_TABLE static __device__ const double
#include "log10f_glibc.h"

template <typename XY> __device__ __forceinline__ double axis_log(XY v);
template <> __device__ __forceinline__ double axis_log<float>(float v) {
  return (v > 0.0f && v < INFINITY) ? (double)lg_log10f(v) : (double)log10f(v);
}
template <> __device__ __forceinline__ double axis_log<double>(double v) { return log10(v); }

template <typename XY>
__device__ __forceinline__ int64_t map_to_cell(const dsb_view& v, XY xr, XY yr) {
  double x = (double)xr, y = (double)yr;
  if (!(v.xmin <= x && x <= v.xmax && v.ymin <= y && y <= v.ymax)) return -1;   // NaN fails the test
  double xm = v.x_log ? axis_log<XY>(xr) : x;
  double ym = v.y_log ? axis_log<XY>(yr) : y;
  int xx = __double2int_rz(__dadd_rn(__dmul_rn(xm, v.sx), v.tx));
  int yy = __double2int_rz(__dadd_rn(__dmul_rn(ym, v.sy), v.ty));
  if (xx >= v.width) xx = v.width - 1;
  if (yy >= v.height) yy = v.height - 1;
  // NOTE: a NaN/negative product cannot occur for an in-bounds linear coordinate (x >= xmin implies
  // x*sx+tx >= -ulp), and int() of a value in (-1, 0) truncates to 0 exactly as the reference does.
  if (xx < 0 || yy < 0) return -1;   // log axes with non-positive in-bounds values: undefined in the reference
  return (int64_t)yy * v.width + xx;
}
