// Host build of vloam-cmu-16833_b200/csrc/fdlibm_atan2f.h (the atan2f the CUDA scan registration uses) against the C
// library's atan2f, bit for bit.  Built and driven by tests/test_oracle_units.py (g++ -O2 -ffp-contract=off).
#include <math.h>

#include "../../vloam-cmu-16833_b200/csrc/fdlibm_atan2f.h"

extern "C" long fd_mismatches(long n, unsigned seed, float scale) {
  long diff = 0;
  unsigned s = seed;
  for (long i = 0; i < n; i++) {
    s = s * 1664525u + 1013904223u;
    float y = ((int)(s >> 8) - 8388608) / 8388608.0f * scale;
    s = s * 1664525u + 1013904223u;
    float x = ((int)(s >> 8) - 8388608) / 8388608.0f * scale;
    if (i % 7 == 0) x *= 1e-3f;
    if (i % 11 == 0) y *= 1e-4f;
    if (i % 13 == 0) x = 1.0f;
    if (i % 17 == 0) y = x;
    const float a = atan2f(y, x), b = vb_fdlibm::atan2f_fd(y, x);
    if (vb_fdlibm::fbits(a) != vb_fdlibm::fbits(b)) diff++;
  }
  return diff;
}
// 1 if both agree on (y, x) (NaN results compare equal)
extern "C" int fd_same(float y, float x) {
  const float a = atan2f(y, x), b = vb_fdlibm::atan2f_fd(y, x);
  return (a != a && b != b) || vb_fdlibm::fbits(a) == vb_fdlibm::fbits(b);
}
extern "C" float fd_atan2f(float y, float x) { return vb_fdlibm::atan2f_fd(y, x); }
