// The accumulator ops shared by the point and line kernels: one call of the reference's generated append()
// (compiler.py:321-475) for one (pixel, row) pair, as commutative atomics (see dsb_op in include/dsb200.h).
#pragma once
#include "common.cuh"

__device__ __forceinline__ int load_cat(const void* p, int dt, long long i) {
  switch (dt) {
    case DSB_I8: return (int)__ldg((const int8_t*)p + i);
    case DSB_I16: return (int)__ldg((const int16_t*)p + i);
    case DSB_I32: return __ldg((const int32_t*)p + i);
    case DSB_I64: return (int)__ldg((const long long*)p + i);
    case DSB_U8: return (int)__ldg((const uint8_t*)p + i);
  }
  return -1;
}

// max / min accumulators are monotone, so a hit that cannot change the pixel needs no atomic: a plain (L1-cacheable)
// load of the current value filters it out.  A stale cached value is at worst less extreme than the real one, which
// only costs a redundant RED - never a lost update.  After the first few hits per pixel almost every hit is filtered
// (the running maximum of n values changes ~ln n times), which takes these reductions off the 180 G/s RED wall.
// Measured (900x525, 1e9 points): max / min 152 -> 191, first 154 -> 184, where(max) 153 -> 168 Gpts/s.  On canvases that
// are not L2-resident (8192^2, banded passes) the extra load costs more than the REDs it saves for first / where
// (32.7 -> 38.8 ms), so the point kernel turns the filter off there (FILTER = false).
template <bool FILTER, typename T> __device__ __forceinline__ void red_max(T* p, T k) {
  if (!FILTER || k > __ldcg((const T*)p)) atomicMax(p, k);
}
template <bool FILTER, typename T> __device__ __forceinline__ void red_min(T* p, T k) {
  if (!FILTER || k < __ldcg((const T*)p)) atomicMin(p, k);
}

template <bool FILTER = true>
__device__ __forceinline__ void apply_base(const dsb_base& b, long long cell, long long i, long long row, unsigned int* notes = nullptr) {
  // nan_check_column: skip the whole base when that column is null (compiler.py:439-446, 461-466)
  if (b.chk_dtype != DSB_NONE && col_isnan(b.chk, b.chk_dtype, i)) return;
  switch (b.op) {
    case DSB_OP_COUNT:
      if (b.val_dtype != DSB_NONE && col_isnan(b.val, b.val_dtype, i)) return;
      atomicAdd((unsigned int*)b.agg + cell, 1u);
      return;
    case DSB_OP_ANY:
      if (b.val_dtype != DSB_NONE && col_isnan(b.val, b.val_dtype, i)) return;
      ((uint8_t*)b.agg)[cell] = 1;   // idempotent store, same as the reference (reductions.py:872-873)
      return;
    case DSB_OP_SUM: {
      double f = load_f64(b.val, b.val_dtype, i);
      if (f != f) return;
      atomicAdd((double*)b.agg + cell, f);
      return;
    }
    case DSB_OP_MAX32:
    case DSB_OP_MIN32: {
      bool nan, nz;
      int32_t k = load_key32(b.val, b.val_dtype, i, &nan, &nz);
      if (nan) return;
      if (nz && notes) *notes = DSB_NOTE_NEGZERO;      // idempotent store; rare (the data holds a -0.0)
      if (b.op == DSB_OP_MAX32) red_max<FILTER>((int*)b.agg + cell, k);
      else red_min<FILTER>((int*)b.agg + cell, k);
      return;
    }
    case DSB_OP_MAX64:
    case DSB_OP_MIN64: {
      double f = load_f64(b.val, b.val_dtype, i);
      if (f != f) return;
      if (notes && is_negzero(f)) *notes = DSB_NOTE_NEGZERO;
      long long k = key64_from_f64(f);
      if (b.op == DSB_OP_MAX64) red_max<FILTER>((long long*)b.agg + cell, k);
      else red_min<FILTER>((long long*)b.agg + cell, k);
      return;
    }
    case DSB_OP_MAXROW:
      atomicMax((long long*)b.agg + cell, row);     // rows arrive in increasing order: nearly every hit is a new maximum
      return;
    case DSB_OP_MINROW:
      red_min<FILTER>((long long*)b.agg + cell, row);
      return;
    case DSB_OP_ARGMAX32:
    case DSB_OP_ARGMIN32: {
      bool nan, nz;
      int32_t k = load_key32(b.val, b.val_dtype, i, &nan, &nz);
      if (nan) return;
      // ties go to the earliest row: for max the row field is complemented so that a smaller row is larger
      // the row field is the low 32 bits of the GLOBAL row id, so chunks of one frame (< 2^32 rows) can
      // share a canvas; dsb_decode_arg rebuilds the full id from the frame's first row
      if (b.op == DSB_OP_ARGMAX32) {
        long long p = ((long long)k << 32) | (long long)(uint32_t)(~(uint32_t)row);
        red_max<FILTER>((long long*)b.agg + cell, p);
      } else {
        long long p = ((long long)k << 32) | (long long)(uint32_t)row;
        red_min<FILTER>((long long*)b.agg + cell, p);
      }
      return;
    }
    case DSB_OP_MATCHROW64: {
      double f = load_f64(b.val, b.val_dtype, i);
      if (f != f) return;
      if (key64_from_f64(f) == __ldg((const long long*)b.aux + cell)) red_min<FILTER>((long long*)b.agg + cell, row);
      return;
    }
  }
}

