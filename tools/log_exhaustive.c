// Exhaustive comparison of fast_log_unit() (csrc/fast_log.h) with glibc's log over every possible free-flight draw:
// x = r / RAND_MAX, r = 1 .. 2^31 - 1.  Prints the number of arguments whose results differ by 0, 1, >1 ulp.
//   g++ -O2 -std=c++17 -ffp-contract=off -mfma -fopenmp -x c++ tools/log_exhaustive.c -o /tmp/log_exh && /tmp/log_exh
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../cnt_film_monte_carlo_b200/csrc/hop_core.h"
int main(int argc, char** argv) {
  const long long lo = argc > 1 ? atoll(argv[1]) : 1, hi = argc > 2 ? atoll(argv[2]) : 2147483647LL;
  long long       same = 0, one = 0, more = 0, divdiff = 0;
#pragma omp parallel for reduction(+ : same, one, more, divdiff) schedule(static)
  for (long long r = lo; r <= hi; ++r) {
    const double x = (double)r / 2147483647.0;
    const double xd = cntmc::div_by((double)r, cntmc::kRandMax, cntmc::kInvRandMax);
    if (x != xd) ++divdiff;
    const double a = log(x), b = cntmc::fast_log_unit(x);
    int64_t      ia, ib;
    memcpy(&ia, &a, 8);
    memcpy(&ib, &b, 8);
    const int64_t d = ia > ib ? ia - ib : ib - ia;
    if (d == 0) ++same; else if (d == 1) ++one; else ++more;
  }
  printf("{\"args\": %lld, \"identical\": %lld, \"one_ulp\": %lld, \"more\": %lld, \"div_by_mismatch\": %lld}\n", hi - lo + 1, same, one, more, divdiff);
  return more != 0 || divdiff != 0;
}
