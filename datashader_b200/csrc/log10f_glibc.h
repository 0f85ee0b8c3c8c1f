/* glibc 2.39's log10f restated: compiled for the device by csrc/common.cuh (LogAxis on float32 coordinates) and for the host by
 * oracle/log10f_check.c, which proves it against the C library.
 *
 * The reference's LogAxis.mapper is log10(float(val)) (core.py:129-132); under numba a float32 coordinate stays float32,
 * LLVM lowers it to a call of the C library's log10f, and glibc 2.39 (this image) computes it as
 *     z = y * log10_2lo + ivln10 * logf(m);  return z + y * log10_2hi          (sysdeps/ieee754/flt-32/e_log10f.c)
 * in float arithmetic around its table-driven logf (e_logf.c, from ARM's optimized routines: 16-entry table, a cubic in
 * double).  On x86-64 CPUs with FMA the dynamic loader selects logf's FMA build (sysdeps/x86_64/fpu/multiarch), where
 * the compiler contracts the double multiply-adds; log10f itself has no FMA build.  The result is not correctly rounded
 * (it differs from round(log10(x)) for ~10 % of the inputs), so bit-exact pixel parity on log axes needs this very
 * sequence.  oracle/log10f_check.c proves the restatement against the C library for EVERY positive finite float. */
#ifndef DSB_LOG10F_GLIBC_H
#define DSB_LOG10F_GLIBC_H
#include <stdint.h>

#ifndef DSB_LG_FN
#define DSB_LG_FN static inline
#define DSB_LG_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define DSB_LG_FMULF(a, b) ((a) * (b))
#define DSB_LG_FADDF(a, b) ((a) + (b))
#define DSB_LG_ASUINT(f) lg_asuint(f)
#define DSB_LG_ASFLOAT(u) lg_asfloat(u)
static inline uint32_t lg_asuint(float f) { union { float f; uint32_t u; } c; c.f = f; return c.u; }
static inline float lg_asfloat(uint32_t u) { union { float f; uint32_t u; } c; c.u = u; return c.f; }
#define DSB_LG_TABLE static const double
#endif

DSB_LG_TABLE lg_invc[16] = {
  0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010bp+0, 0x1.3c995b0b80385p+0, 0x1.30d190c8864a5p+0,
  0x1.25e227b0b8eap+0, 0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0, 0x1.0953f419900a7p+0, 0x1p+0,
  0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aap-1, 0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1,
  0x1.767dcf5534862p-1};
DSB_LG_TABLE lg_logc[16] = {
  -0x1.57bf7808caadep-2, -0x1.2bef0a7c06ddbp-2, -0x1.01eae7f513a67p-2, -0x1.b31d8a68224e9p-3, -0x1.6574f0ac07758p-3,
  -0x1.1aa2bc79c81p-3, -0x1.a4e76ce8c0e5ep-4, -0x1.1973c5a611cccp-4, -0x1.252f438e10c1ep-5, 0x0p+0,
  0x1.aa5aa5df25984p-5, 0x1.c5e53aa362eb4p-4, 0x1.526e57720db08p-3, 0x1.bc2860d22477p-3, 0x1.1058bc8a07ee1p-2,
  0x1.4043057b6ee09p-2};

/* logf of a positive, finite, normal float (e_logf.c, FMA build) */
DSB_LG_FN float lg_logf(float x) {
  const double ln2 = 0x1.62e42fefa39efp-1, a0 = -0x1.00ea348b88334p-2, a1 = 0x1.5575b0be00b6ap-2, a2 = -0x1.ffffef20a4123p-2;
  uint32_t ix = DSB_LG_ASUINT(x);
  if (ix == 0x3f800000u) return 0.0f;
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (int)((tmp >> 19) & 15u);
  const int k = (int32_t)tmp >> 23;
  const uint32_t iz = ix - (tmp & 0xff800000u);
  const double z = (double)DSB_LG_ASFLOAT(iz);
  const double r = DSB_LG_FMA(z, lg_invc[i], -1.0);
  const double y0 = DSB_LG_FMA((double)k, ln2, lg_logc[i]);
  const double r2 = r * r;
  double y = DSB_LG_FMA(a1, r, a2);
  y = DSB_LG_FMA(a0, r2, y);
  y = DSB_LG_FMA(y, r2, y0 + r);
  return (float)y;
}

/* log10f of a positive, finite float (e_log10f.c; subnormals scaled by 2^25 first) */
DSB_LG_FN float lg_log10f(float x) {
  const float ivln10 = 4.3429449201e-01f, log10_2hi = 3.0102920532e-01f, log10_2lo = 7.9034151668e-07f;
  int32_t hx = (int32_t)DSB_LG_ASUINT(x), k = 0;
  if (hx < 0x00800000) { k -= 25; x = DSB_LG_FMULF(x, 3.3554432000e+07f); hx = (int32_t)DSB_LG_ASUINT(x); }
  k += (hx >> 23) - 127;
  const int32_t i = (int32_t)(((uint32_t)k & 0x80000000u) >> 31);
  hx = (hx & 0x007fffff) | ((0x7f - i) << 23);
  const float y = (float)(k + i);
  const float m = DSB_LG_ASFLOAT((uint32_t)hx);
  const float z = DSB_LG_FADDF(DSB_LG_FMULF(y, log10_2lo), DSB_LG_FMULF(ivln10, lg_logf(m)));
  return DSB_LG_FADDF(z, DSB_LG_FMULF(y, log10_2hi));
}
#endif
