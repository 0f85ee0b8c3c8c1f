// Error reporting and device queries shared by every entry point of libdsb200.
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>

static thread_local char g_err[512] = "";

extern "C" void dsb_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* dsb_last_error(void) { return g_err; }
extern "C" int dsb_abi_version(void) { return DSB_ABI_VERSION; }

int dsb_num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return DSB_SM_COUNT_FALLBACK;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = DSB_SM_COUNT_FALLBACK;
    cached = n; cached_dev = dev;
  }
  return cached;
}

// number of API calls that launched kernels so far (bench.py reports it as gpu_launches)
static unsigned long long g_launches = 0;
void dsb_count_launch() { __atomic_add_fetch(&g_launches, 1ULL, __ATOMIC_RELAXED); }
extern "C" int64_t dsb_launch_count(void) { return (int64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

// name of the aggregation kernel the most recent entry point chose (bench.py's roofline.kernel: read back from the
// library, so a silent change of route cannot be mislabelled)
static thread_local char g_kernel[160] = "";
void dsb_note_kernel(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_kernel, sizeof(g_kernel), fmt, ap);
  va_end(ap);
}
extern "C" const char* dsb_last_kernel(void) { return g_kernel; }
