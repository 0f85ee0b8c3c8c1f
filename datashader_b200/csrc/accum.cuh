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

__device__ __forceinline__ void apply_base(const dsb_base& b, long long cell, long long i, long long row) {
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
      bool nan;
      int32_t k = load_key32(b.val, b.val_dtype, i, &nan);
      if (nan) return;
      if (b.op == DSB_OP_MAX32) atomicMax((int*)b.agg + cell, k);
      else atomicMin((int*)b.agg + cell, k);
      return;
    }
    case DSB_OP_MAX64:
    case DSB_OP_MIN64: {
      double f = load_f64(b.val, b.val_dtype, i);
      if (f != f) return;
      long long k = key64_from_f64(f);
      if (b.op == DSB_OP_MAX64) atomicMax((long long*)b.agg + cell, k);
      else atomicMin((long long*)b.agg + cell, k);
      return;
    }
    case DSB_OP_MAXROW:
      atomicMax((long long*)b.agg + cell, row);
      return;
    case DSB_OP_MINROW:
      atomicMin((long long*)b.agg + cell, row);
      return;
    case DSB_OP_ARGMAX32:
    case DSB_OP_ARGMIN32: {
      bool nan;
      int32_t k = load_key32(b.val, b.val_dtype, i, &nan);
      if (nan) return;
      // ties go to the earliest row: for max the row field is complemented so that a smaller row is larger
      // the row field is the low 32 bits of the GLOBAL row id, so chunks of one frame (< 2^32 rows) can
      // share a canvas; dsb_decode_arg rebuilds the full id from the frame's first row
      if (b.op == DSB_OP_ARGMAX32) {
        long long p = ((long long)k << 32) | (long long)(uint32_t)(~(uint32_t)row);
        atomicMax((long long*)b.agg + cell, p);
      } else {
        long long p = ((long long)k << 32) | (long long)(uint32_t)row;
        atomicMin((long long*)b.agg + cell, p);
      }
      return;
    }
    case DSB_OP_MATCHROW64: {
      double f = load_f64(b.val, b.val_dtype, i);
      if (f != f) return;
      if (key64_from_f64(f) == __ldg((const long long*)b.aux + cell)) atomicMin((long long*)b.agg + cell, row);
      return;
    }
  }
}

