/* Exhaustive check of datashader_b200/csrc/log10f_glibc.h against this box's C library: every positive finite float.
 *   gcc -O2 -ffp-contract=off -fopenmp -o _build/log10f_check log10f_check.c -lm && _build/log10f_check
 * Prints the number of inputs whose bit patterns differ (0 = the restatement IS the library's log10f here). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "../datashader_b200/csrc/log10f_glibc.h"

int main(int argc, char** argv) {
  const uint32_t lo = 1u, hi = 0x7f800000u;          /* (0, +inf) */
  const uint32_t step = argc > 1 ? (uint32_t)atoi(argv[1]) : 1u;
  unsigned long long bad = 0, total = 0;
  uint32_t first_bad = 0;
#pragma omp parallel for reduction(+ : bad, total) schedule(static)
  for (long long u = lo; u < (long long)hi; u += step) {
    const float x = lg_asfloat((uint32_t)u);
    volatile float xv = x;
    const float want = log10f(xv), got = lg_log10f(x);
    total++;
    if (lg_asuint(want) != lg_asuint(got)) {
      bad++;
#pragma omp critical
      if (!first_bad) first_bad = (uint32_t)u;
    }
  }
  printf("checked %llu floats, %llu differ", total, bad);
  if (bad) printf(" (first: 0x%08x = %.9g: libm %.9g, restated %.9g)", first_bad, lg_asfloat(first_bad), log10f(lg_asfloat(first_bad)), lg_log10f(lg_asfloat(first_bad)));
  printf("\n");
  return bad != 0;
}
