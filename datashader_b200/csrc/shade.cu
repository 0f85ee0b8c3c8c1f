// tf.shade on the GPU: categorical colour mixing + alpha by eq_hist / log / cbrt / linear, and the
// histogram-equalisation machinery (histogram + block scan + CDF lookup).
//
// Replaces, for the covered paths, transfer_functions/__init__.py: eq_hist (:148-215), _colorize
// (:359-463), _interpolate_alpha (:466-532), _interpolate (:251-357), and the cupy/CUB calls plus
// interp2d_kernel (_cuda_utils.py:114-174) the reference's GPU path uses.  The float arithmetic follows
// numpy's histogram / linspace / interp formulas step by step (unfused f64) so that eq_hist and linear
// results are bit-identical to the reference; log / cbrt go through CUDA's log1p / pow (<= 2 ulp).
#include "common.cuh"

__device__ __forceinline__ double m64(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double a64(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double s64(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ double d64(double a, double b) { return __ddiv_rn(a, b); }

// ---- step 1: per-pixel totals + global statistics ----------------------------------------------
// stats[0] min entry (baseline, _colorize :398), [1] min total, [2] min non-zero total, [3] max total
__global__ void k_cat_totals(const uint32_t* __restrict__ counts, long long npix, int ncat,
                             unsigned long long* __restrict__ total, unsigned long long* stats) {
  unsigned long long mn_e = ~0ull, mn_t = ~0ull, mn_nz = ~0ull, mx_t = 0;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < npix; i += stride) {
    const uint32_t* p = counts + i * ncat;
    unsigned long long t = 0;
    for (int c = 0; c < ncat; c++) {
      uint32_t v = __ldcs(p + c);
      t += v;
      mn_e = min(mn_e, (unsigned long long)v);
    }
    total[i] = t;
    mn_t = min(mn_t, t);
    if (t) mn_nz = min(mn_nz, t);
    mx_t = max(mx_t, t);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn_e = min(mn_e, __shfl_xor_sync(0xffffffffu, mn_e, o));
    mn_t = min(mn_t, __shfl_xor_sync(0xffffffffu, mn_t, o));
    mn_nz = min(mn_nz, __shfl_xor_sync(0xffffffffu, mn_nz, o));
    mx_t = max(mx_t, __shfl_xor_sync(0xffffffffu, mx_t, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(stats + 0, mn_e); atomicMin(stats + 1, mn_t); atomicMin(stats + 2, mn_nz); atomicMax(stats + 3, mx_t);
  }
}

__global__ void k_stats_init(unsigned long long* stats) {
  stats[0] = ~0ull; stats[1] = ~0ull; stats[2] = ~0ull; stats[3] = 0;
}

// the same statistics for a 2-D u32 canvas viewed as its own total
__global__ void k_u32_totals(const uint32_t* __restrict__ v, long long npix, unsigned long long* __restrict__ total,
                             unsigned long long* stats) {
  unsigned long long mn_t = ~0ull, mn_nz = ~0ull, mx_t = 0;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < npix; i += stride) {
    unsigned long long t = v[i];
    total[i] = t;
    mn_t = min(mn_t, t);
    if (t) mn_nz = min(mn_nz, t);
    mx_t = max(mx_t, t);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn_t = min(mn_t, __shfl_xor_sync(0xffffffffu, mn_t, o));
    mn_nz = min(mn_nz, __shfl_xor_sync(0xffffffffu, mn_nz, o));
    mx_t = max(mx_t, __shfl_xor_sync(0xffffffffu, mx_t, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(stats + 0, mn_t); atomicMin(stats + 1, mn_t); atomicMin(stats + 2, mn_nz); atomicMax(stats + 3, mx_t);
  }
}

static int sgrid(long long n, int threads) {
  long long g = (n + threads - 1) / threads, cap = (long long)dsb_num_sms() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

extern "C" int dsb_shade_cat_totals(const uint32_t* counts, int64_t npix, int32_t ncat, uint64_t* total, uint64_t* stats,
                                    void* stream) {
  if (!counts || !total || !stats || npix < 0 || ncat < 1) { dsb_set_error("dsb_shade_cat_totals: bad arguments"); return DSB_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  k_stats_init<<<1, 1, 0, s>>>((unsigned long long*)stats);
  if (npix > 0) {
    if (ncat == 1) k_u32_totals<<<sgrid(npix, 256), 256, 0, s>>>(counts, npix, (unsigned long long*)total, (unsigned long long*)stats);
    else k_cat_totals<<<sgrid(npix, 256), 256, 0, s>>>(counts, npix, ncat, (unsigned long long*)total, (unsigned long long*)stats);
  }
  DSB_CUDA_CHECK_LAUNCH("dsb_shade_cat_totals");
  return DSB_OK;
}

// ---- step 2: histogram with numpy.histogram's uniform-bin rule ------------------------------------
struct EqParams {
  double first, last;      // outer edges (np.histogram _get_outer_edges; widened by 0.5 when equal)
  double step;             // (last - first) / nbins  (np.linspace)
  int nbins;
  int integer_mode;        // exact unique-value path (eq_hist :194-202): one bin per integer in [first, last]
};

__device__ __forceinline__ double lin_edge(const EqParams& e, int i) {
  // np.linspace: y = arange(num) * step + start, last element forced to stop
  return i == e.nbins ? e.last : a64(m64((double)i, e.step), e.first);
}

__device__ __forceinline__ int hist_index(const EqParams& e, double x) {
  if (e.integer_mode) return (int)(x - e.first);
  // numpy/lib/_histograms_impl.py fast path for equal-width bins
  double f = m64(d64(s64(x, e.first), s64(e.last, e.first)), (double)e.nbins);
  int idx = __double2int_rz(f);
  if (idx == e.nbins) idx -= 1;
  if (x < lin_edge(e, idx)) idx -= 1;
  else if (x >= lin_edge(e, idx + 1) && idx != e.nbins - 1) idx += 1;
  return idx;
}

__global__ void k_eqhist_hist(const unsigned long long* __restrict__ total, long long npix, unsigned long long offset,
                              int mask_zero, EqParams e, uint32_t* __restrict__ hist) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < npix; i += stride) {
    unsigned long long t = total[i];
    if (mask_zero && t == 0) continue;
    double x = s64((double)t, (double)offset);
    int idx = hist_index(e, x);
    if (idx >= 0 && idx < e.nbins) atomicAdd(hist + idx, 1u);
  }
}

__global__ void k_eqhist_hist_f64(const double* __restrict__ v, long long npix, double offset, EqParams e,
                                  uint32_t* __restrict__ hist) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < npix; i += stride) {
    double t = v[i];
    if (t != t) continue;
    int idx = hist_index(e, s64(t, offset));
    if (idx >= 0 && idx < e.nbins) atomicAdd(hist + idx, 1u);
  }
}

static EqParams make_params(double first, double last, int nbins, int integer_mode) {
  EqParams e;
  if (!integer_mode && first == last) { first -= 0.5; last += 0.5; }   // _get_outer_edges
  e.first = first; e.last = last; e.nbins = nbins; e.integer_mode = integer_mode;
  e.step = (last - first) / (double)nbins;
  return e;
}

extern "C" int dsb_eqhist_hist_u64(const uint64_t* total, int64_t npix, uint64_t offset, int32_t mask_zero, double first,
                                   double last, int32_t nbins, int32_t integer_mode, uint32_t* hist, void* stream) {
  if (!total || !hist || nbins < 1) { dsb_set_error("dsb_eqhist_hist_u64: bad arguments"); return DSB_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(hist, 0, (size_t)nbins * 4, s);
  if (npix > 0)
    k_eqhist_hist<<<sgrid(npix, 256), 256, 0, s>>>((const unsigned long long*)total, npix, offset, mask_zero,
                                                    make_params(first, last, nbins, integer_mode), hist);
  DSB_CUDA_CHECK_LAUNCH("dsb_eqhist_hist_u64");
  return DSB_OK;
}

extern "C" int dsb_eqhist_hist_f64(const double* vals, int64_t npix, double offset, double first, double last,
                                   int32_t nbins, int32_t integer_mode, uint32_t* hist, void* stream) {
  if (!vals || !hist || nbins < 1) { dsb_set_error("dsb_eqhist_hist_f64: bad arguments"); return DSB_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  cudaMemsetAsync(hist, 0, (size_t)nbins * 4, s);
  if (npix > 0) k_eqhist_hist_f64<<<sgrid(npix, 256), 256, 0, s>>>(vals, npix, offset, make_params(first, last, nbins, integer_mode), hist);
  DSB_CUDA_CHECK_LAUNCH("dsb_eqhist_hist_f64");
  return DSB_OK;
}

// ---- step 3: scan: drop empty bins, cumulative sum, CDF -------------------------------------------
// One CTA of 1024 threads; each thread owns a contiguous run of bins.  Two block scans: positions of the
// kept bins and the running count.  xp[j] = centre of kept bin j, cdf[j] = cumsum_j / cumsum_last.
// meta[0] = L (entries of xp/cdf), meta[1] = discrete_levels (non-empty bins).
__global__ void __launch_bounds__(1024) k_eqhist_scan(const uint32_t* __restrict__ hist, EqParams e, double* __restrict__ xp,
                                                      double* __restrict__ cdf, int* __restrict__ meta) {
  __shared__ unsigned long long s_cnt[1024];
  __shared__ int s_keep[1024];
  const int tid = threadIdx.x;
  const int per = (e.nbins + 1023) / 1024;
  const int lo = tid * per, hi = min(e.nbins, lo + per);
  unsigned long long cnt = 0;
  int keep = 0, nonempty = 0;
  for (int b = lo; b < hi; b++) {
    uint32_t h = hist[b];
    cnt += h;
    nonempty += (h > 0);
    keep += (e.integer_mode || h > 0);
  }
  s_cnt[tid] = cnt; s_keep[tid] = keep;
  __syncthreads();
  // Hillis-Steele inclusive scans (1024 elements, 10 steps)
  for (int o = 1; o < 1024; o <<= 1) {
    unsigned long long c = (tid >= o) ? s_cnt[tid - o] : 0;
    int k = (tid >= o) ? s_keep[tid - o] : 0;
    __syncthreads();
    s_cnt[tid] += c; s_keep[tid] += k;
    __syncthreads();
  }
  const unsigned long long total = s_cnt[1023];
  unsigned long long run = s_cnt[tid] - cnt;
  int pos = s_keep[tid] - keep;
  const double totald = (double)total;
  for (int b = lo; b < hi; b++) {
    uint32_t h = hist[b];
    run += h;
    if (e.integer_mode || h > 0) {
      // bin_centers = (edges[:-1] + edges[1:]) / 2 (eq_hist :205); integer mode: arange(vmin, vmax + 1)
      xp[pos] = e.integer_mode ? a64(e.first, (double)b) : d64(a64(lin_edge(e, b), lin_edge(e, b + 1)), 2.0);
      cdf[pos] = d64((double)run, totald);          // cdf = hist.cumsum() / float(cdf[-1])
      pos++;
    }
  }
  // discrete levels: reuse s_keep for a plain reduction
  __syncthreads();
  s_keep[tid] = nonempty;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (tid < o) s_keep[tid] += s_keep[tid + o];
    __syncthreads();
  }
  if (tid == 1023) meta[0] = pos;
  if (tid == 0) meta[1] = s_keep[0];
}

extern "C" int dsb_eqhist_scan(const uint32_t* hist, int32_t nbins, int32_t integer_mode, double first, double last,
                               double* xp, double* cdf, int32_t* meta, void* stream) {
  if (!hist || !xp || !cdf || !meta || nbins < 1) { dsb_set_error("dsb_eqhist_scan: bad arguments"); return DSB_ERR_ARG; }
  k_eqhist_scan<<<1, 1024, 0, (cudaStream_t)stream>>>(hist, make_params(first, last, nbins, integer_mode), xp, cdf, meta);
  DSB_CUDA_CHECK_LAUNCH("dsb_eqhist_scan");
  return DSB_OK;
}

// ---- np.interp ------------------------------------------------------------------------------------
// numpy/_core/src/multiarray/compiled_base.c:arr_interp - left/right fills, single-point case, and
// slope * (x - xp[j]) + fp[j] evaluated unfused.
__device__ __forceinline__ double np_interp(double x, const double* __restrict__ xp, const double* __restrict__ fp, int n,
                                            double left, double right) {
  if (x != x) return x;
  if (n == 1) return (x < xp[0]) ? left : ((x > xp[0]) ? right : fp[0]);
  if (x > xp[n - 1]) return right;
  if (x < xp[0]) return left;
  int lo = 0, hi = n - 1;            // find j with xp[j] <= x < xp[j+1]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (x >= xp[mid]) lo = mid; else hi = mid;
  }
  int j = (x >= xp[hi]) ? hi : lo;
  if (j == n - 1) return fp[j];
  if (x == xp[j]) return fp[j];
  double slope = d64(s64(fp[j + 1], fp[j]), s64(xp[j + 1], xp[j]));
  double r = a64(m64(slope, s64(x, xp[j])), fp[j]);
  if (r != r) {                      // numpy's fallback when the first form is NaN (inf slopes)
    r = a64(m64(slope, s64(x, xp[j + 1])), fp[j + 1]);
    if (r != r && fp[j] == fp[j + 1]) r = fp[j];
  }
  return r;
}

// a_scaled for one pixel: eq_hist -> CDF lookup, otherwise the analytic transfer functions
// DSB_HOW_F32: the canvas is float32, so numpy evaluates log1p / ** (1/3.) in float32 (the python-float exponent is
// cast to float32); the correctly rounded float32 result is obtained by rounding the float64 function value.
__device__ __forceinline__ double transfer(int how, double d, const double* xp, const double* cdf, int L) {
  const bool f32 = (how & DSB_HOW_F32) != 0;
  switch (how & 0xff) {
    case DSB_HOW_EQ_HIST: return np_interp(d, xp, cdf, L, cdf[0], cdf[L - 1]);
    case DSB_HOW_LOG: return f32 ? (double)(float)log1p(d) : log1p(d);
    case DSB_HOW_CBRT: return f32 ? (double)(float)pow(d, (double)(float)(1.0 / 3.0)) : pow(d, 1.0 / 3.0);
    default: return d;
  }
}

// span[0..1] = norm_span on the device: transfer(min d), transfer(max d) [+ _rescale_discrete_levels]
__global__ void k_norm_span(int how, double dmin, double dmax, const double* xp, const double* cdf, const int* meta,
                            int rescale, double* span) {
  int L = meta ? meta[0] : 0;
  double lo = transfer(how, dmin, xp, cdf, L), hi = transfer(how, dmax, xp, cdf, L);
  if (rescale && how == DSB_HOW_EQ_HIST) {     // :232-248
    double m = -0.5 / 98.0, c = 1.5 - 2 * m;
    double multiple = a64(m64(m, (double)meta[1]), c);
    if (multiple > 1) {
      double lower = s64(hi, m64(multiple, s64(hi, lo)));
      lo = lower > 0 ? lower : 0;
      hi = 1;
    }
  }
  span[0] = lo; span[1] = hi;
}

extern "C" int dsb_shade_norm_span(int32_t how, double dmin, double dmax, const double* xp, const double* cdf,
                                   const int32_t* meta, int32_t rescale, double* span, void* stream) {
  if (!span || (how == DSB_HOW_EQ_HIST && (!xp || !cdf || !meta))) { dsb_set_error("dsb_shade_norm_span: bad arguments"); return DSB_ERR_ARG; }
  k_norm_span<<<1, 1, 0, (cudaStream_t)stream>>>(how, dmin, dmax, xp, cdf, meta, rescale, span);
  DSB_CUDA_CHECK_LAUNCH("dsb_shade_norm_span");
  return DSB_OK;
}

// alpha = interp(a_scaled, norm_span, [min_alpha, alpha], left=0, right=255) -> uint8 (:530-531)
__device__ __forceinline__ uint32_t alpha_u8(double a_scaled, double lo, double hi, double min_alpha, double alpha) {
  if (a_scaled != a_scaled) return 0;
  double xp2[2] = {lo, hi}, fp2[2] = {min_alpha, alpha};
  double a = np_interp(a_scaled, xp2, fp2, 2, 0.0, 255.0);
  if (a != a) return 0;
  return (uint32_t)(unsigned char)__double2int_rz(a);
}

// ---- step 4: categorical colour mix + alpha (:359-463) --------------------------------------------
struct CatColorArgs {
  const uint32_t* counts;
  const unsigned long long* total;
  long long npix;
  int ncat;
  const float* rgb;          // [ncat, 3]
  uint32_t fallback_rgb;     // r | g << 8 | b << 16 of the "average of present categories" colour (:432-442)
  uint32_t baseline;         // nanmin(color_data) or color_baseline
  unsigned long long offset;
  int mask_zero;
  int how;
  const double* xp; const double* cdf; const int* meta; const double* span;
  double min_alpha, alpha;
  int clip_mode;             // explicit span (:507-514): 0 none, 1 the totals became float64 (zeros masked), 2 stay uint64
  double clip_lo, clip_hi;
  uint32_t* out;
};

__global__ void __launch_bounds__(256) k_cat_colorize(const CatColorArgs a) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int L = a.meta ? a.meta[0] : 0;
  const double lo = a.span[0], hi = a.span[1];
  for (; i < a.npix; i += stride) {
    const uint32_t* p = a.counts + i * a.ncat;
    float tot = 0.f, r = 0.f, g = 0.f, b = 0.f;
    for (int c = 0; c < a.ncat; c++) {
      float v = (float)(__ldcs(p + c) - a.baseline);          // u32 subtraction, then astype(float32)
      tot = __fadd_rn(tot, v);
      r = __fadd_rn(r, __fmul_rn(v, a.rgb[c * 3 + 0]));
      g = __fadd_rn(g, __fmul_rn(v, a.rgb[c * 3 + 1]));
      b = __fadd_rn(b, __fmul_rn(v, a.rgb[c * 3 + 2]));
    }
    uint32_t rgb;
    if (tot == 0.f) rgb = a.fallback_rgb;
    else {
      uint32_t ru = (uint32_t)(unsigned char)__float2int_rz(__fdiv_rn(r, tot));
      uint32_t gu = (uint32_t)(unsigned char)__float2int_rz(__fdiv_rn(g, tot));
      uint32_t bu = (uint32_t)(unsigned char)__float2int_rz(__fdiv_rn(b, tot));
      rgb = ru | (gu << 8) | (bu << 16);
    }
    unsigned long long t = a.total[i];
    uint32_t al = 0;
    if (!(a.mask_zero && t == 0)) {
      double d;
      if (a.clip_mode == 1) {            // masked_clip_2d on float64 totals, then total - offset
        double tv = (double)t;
        if (tv < a.clip_lo) tv = a.clip_lo; else if (tv > a.clip_hi) tv = a.clip_hi;
        d = s64(tv, (double)a.offset);
      } else if (a.clip_mode == 2) {     // on uint64 totals: the bound is truncated to an integer when it is stored
        if ((double)t < a.clip_lo) t = (unsigned long long)a.clip_lo; else if ((double)t > a.clip_hi) t = (unsigned long long)a.clip_hi;
        d = (double)(t - a.offset);
      } else {
        d = s64((double)t, (double)a.offset);
      }
      al = alpha_u8(transfer(a.how, d, a.xp, a.cdf, L), lo, hi, a.min_alpha, a.alpha);
    }
    a.out[i] = rgb | (al << 24);
  }
}

extern "C" int dsb_shade_cat_colorize(const uint32_t* counts, const uint64_t* total, int64_t npix, int32_t ncat,
                                      const float* rgb, uint32_t fallback_rgb, uint32_t baseline, uint64_t offset,
                                      int32_t mask_zero, int32_t how, const double* xp, const double* cdf,
                                      const int32_t* meta, const double* span, double min_alpha, double alpha,
                                      int32_t clip_mode, double clip_lo, double clip_hi, uint32_t* out, void* stream) {
  if (!counts || !total || !rgb || !span || !out || ncat < 1) { dsb_set_error("dsb_shade_cat_colorize: bad arguments"); return DSB_ERR_ARG; }
  if (how == DSB_HOW_EQ_HIST && (!xp || !cdf || !meta)) { dsb_set_error("dsb_shade_cat_colorize: eq_hist needs xp/cdf/meta"); return DSB_ERR_ARG; }
  if (npix == 0) return DSB_OK;
  CatColorArgs a;
  a.counts = counts; a.total = (const unsigned long long*)total; a.npix = npix; a.ncat = ncat; a.rgb = rgb;
  a.fallback_rgb = fallback_rgb; a.baseline = baseline; a.offset = offset; a.mask_zero = mask_zero; a.how = how;
  a.xp = xp; a.cdf = cdf; a.meta = meta; a.span = span; a.min_alpha = min_alpha; a.alpha = alpha; a.out = out;
  a.clip_mode = clip_mode; a.clip_lo = clip_lo; a.clip_hi = clip_hi;
  k_cat_colorize<<<sgrid(npix, 256), 256, 0, (cudaStream_t)stream>>>(a);
  DSB_CUDA_CHECK_LAUNCH("dsb_shade_cat_colorize");
  return DSB_OK;
}

// ---- 2-D colormapping (_interpolate :251-357) ------------------------------------------------------
// data is presented as f64 (NaN = masked) AFTER the offset subtraction; the cmap is a list of colours
// (n >= 2: r/g/b interpolated over linspace(span), alpha fixed) or a single colour (alpha ramp).
struct MapArgs {
  const double* data;        // d = value - offset, NaN where masked
  long long npix;
  int how;
  const double* xp; const double* cdf; const int* meta; const double* span;
  int ncolors;               // >= 2: list cmap; 1: single colour
  const double* cspan;       // [ncolors] linspace(span[0], span[1], ncolors)   (list cmap)
  const double* rs; const double* gs; const double* bs;
  double min_alpha, alpha;
  uint32_t* out;
};

__global__ void __launch_bounds__(256) k_map2d(const MapArgs a) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int L = a.meta ? a.meta[0] : 0;
  for (; i < a.npix; i += stride) {
    double d = a.data[i];
    if (d != d) {
      // masked: list cmaps give nan_to_num(NaN) = 0 in every channel; a single colour keeps its rgb with alpha 0
      a.out[i] = (a.ncolors >= 2) ? 0u : ((uint32_t)a.rs[0] | ((uint32_t)a.gs[0] << 8) | ((uint32_t)a.bs[0] << 16));
      continue;
    }
    double v = transfer(a.how, d, a.xp, a.cdf, L);
    uint32_t px;
    if (a.ncolors >= 2) {
      // interp(data, span, rspan, left=255) -> nan_to_num -> uint8 ; alpha where not NaN (:317-321)
      double r = np_interp(v, a.cspan, a.rs, a.ncolors, 255.0, a.rs[a.ncolors - 1]);
      double g = np_interp(v, a.cspan, a.gs, a.ncolors, 255.0, a.gs[a.ncolors - 1]);
      double b = np_interp(v, a.cspan, a.bs, a.ncolors, 255.0, a.bs[a.ncolors - 1]);
      uint32_t al = (v != v) ? 0u : (uint32_t)(unsigned char)(int)a.alpha;
      px = (uint32_t)(unsigned char)__double2int_rz(r) | ((uint32_t)(unsigned char)__double2int_rz(g) << 8) |
           ((uint32_t)(unsigned char)__double2int_rz(b) << 16) | (al << 24);
    } else {
      // single colour: alpha = interp(data, linspace(span, len(aspan)), aspan, left=0, right=255) (:323-330).
      // aspan is linear in the index, but the reference interpolates over the discretised ramp: reproduce it.
      int na = (int)a.alpha - (int)a.min_alpha + 1;
      double lo = a.span[0], hi = a.span[1];
      double al;
      if (v != v) al = 0.0;
      else if (na == 1) al = (v < lo) ? 0.0 : ((v > lo) ? 255.0 : a.min_alpha);
      else if (v > hi) al = 255.0;
      else if (v < lo) al = 0.0;
      else {
        // linspace(lo, hi, na): step = (hi - lo) / (na - 1); x_k = k * step + lo, x_{na-1} = hi
        double step = d64(s64(hi, lo), (double)(na - 1));
        int k = (step > 0) ? __double2int_rz(d64(s64(v, lo), step)) : 0;
        if (k > na - 1) k = na - 1;
        if (k < 0) k = 0;
        auto xk = [&](int q) { return q == na - 1 ? hi : a64(m64((double)q, step), lo); };
        while (k > 0 && v < xk(k)) k--;
        while (k < na - 1 && v >= xk(k + 1)) k++;
        if (k == na - 1 || v == xk(k)) al = a.min_alpha + k;
        else {
          double slope = d64(1.0, s64(xk(k + 1), xk(k)));
          al = a64(m64(slope, s64(v, xk(k))), a.min_alpha + k);
        }
      }
      px = (uint32_t)a.rs[0] | ((uint32_t)a.gs[0] << 8) | ((uint32_t)a.bs[0] << 16) |
           ((uint32_t)(unsigned char)__double2int_rz(al) << 24);
    }
    a.out[i] = px;
  }
}

extern "C" int dsb_shade_map2d(const double* data, int64_t npix, int32_t how, const double* xp, const double* cdf,
                               const int32_t* meta, const double* span, int32_t ncolors, const double* cspan,
                               const double* rs, const double* gs, const double* bs, double min_alpha, double alpha,
                               uint32_t* out, void* stream) {
  if (!data || !span || !rs || !gs || !bs || !out || ncolors < 1) { dsb_set_error("dsb_shade_map2d: bad arguments"); return DSB_ERR_ARG; }
  if (ncolors >= 2 && !cspan) { dsb_set_error("dsb_shade_map2d: list cmap needs cspan"); return DSB_ERR_ARG; }
  if ((how & 0xff) == DSB_HOW_EQ_HIST && (!xp || !cdf || !meta)) { dsb_set_error("dsb_shade_map2d: eq_hist needs xp/cdf/meta"); return DSB_ERR_ARG; }
  if (npix == 0) return DSB_OK;
  MapArgs a;
  a.data = data; a.npix = npix; a.how = how; a.xp = xp; a.cdf = cdf; a.meta = meta; a.span = span; a.ncolors = ncolors;
  a.cspan = cspan; a.rs = rs; a.gs = gs; a.bs = bs; a.min_alpha = min_alpha; a.alpha = alpha; a.out = out;
  k_map2d<<<sgrid(npix, 256), 256, 0, (cudaStream_t)stream>>>(a);
  DSB_CUDA_CHECK_LAUNCH("dsb_shade_map2d");
  return DSB_OK;
}
