// vloam_b200 — atan2f with the bits of the reference's platform.
//
// scanRegistration turns azimuths into the relative time inside the sweep with std::atan2(float, float)
// (reference scan_registration.cpp:166-167, 234), i.e. the C library's atan2f.  Its unwrapping thresholds
// (:236-262) turn a last-place difference into a 2*pi jump for points that sit on them, which moves
// `intensity` by 0.1 and can change int(intensity), the scan id laserOdometry's ring-window tests read
// (laser_odometry.cpp:275-324).  CUDA's atan2f differs from glibc's in the last place for ~16 % of the
// arguments, so the device computes the function the way the library does: this is the fdlibm algorithm
// (e_atan2f.c / s_atanf.c: argument reduction to four intervals + an 11-term odd/even polynomial, all in
// float) that glibc shipped up to 2.40 — the library of every Ubuntu release ROS 1 ran on.  Every
// operation is a single IEEE float operation in the same order (no FMA contraction on the device:
// explicit round-to-nearest intrinsics; on the host compile with -ffp-contract=off).
// tests/test_oracle_units.py compiles this header for the host and checks it against the C library's
// atan2f bit for bit on tens of millions of arguments.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define VB_FD_FN __host__ __device__ __forceinline__
#else
#include <math.h>
#define VB_FD_FN static inline
#endif

namespace vb_fdlibm {

#if defined(__CUDA_ARCH__)
VB_FD_FN float fmul(float a, float b) { return __fmul_rn(a, b); }
VB_FD_FN float fadd(float a, float b) { return __fadd_rn(a, b); }
VB_FD_FN float fsub(float a, float b) { return __fsub_rn(a, b); }
VB_FD_FN float fdiv(float a, float b) { return __fdiv_rn(a, b); }
VB_FD_FN int32_t fbits(float x) { return __float_as_int(x); }
VB_FD_FN float fromBits(int32_t i) { return __int_as_float(i); }
#else
VB_FD_FN float fmul(float a, float b) { return a * b; }
VB_FD_FN float fadd(float a, float b) { return a + b; }
VB_FD_FN float fsub(float a, float b) { return a - b; }
VB_FD_FN float fdiv(float a, float b) { return a / b; }
VB_FD_FN int32_t fbits(float x) { int32_t i; memcpy(&i, &x, 4); return i; }
VB_FD_FN float fromBits(int32_t i) { float x; memcpy(&x, &i, 4); return x; }
#endif

// s_atanf.c
VB_FD_FN float atanf_fd(float x) {
  // atan(0.5), atan(1), atan(1.5), atan(inf): high and low parts
  const float hi0 = 4.6364760399e-01f, hi1 = 7.8539812565e-01f, hi2 = 9.8279368877e-01f, hi3 = 1.5707962513e+00f;
  const float lo0 = 5.0121582440e-09f, lo1 = 3.7748947079e-08f, lo2 = 3.4473217170e-08f, lo3 = 7.5497894159e-08f;
  const float a0 = 3.3333334327e-01f, a1 = -2.0000000298e-01f, a2 = 1.4285714924e-01f, a3 = -1.1111110449e-01f,
              a4 = 9.0908870101e-02f, a5 = -7.6918758452e-02f, a6 = 6.6610731184e-02f, a7 = -5.8335702866e-02f,
              a8 = 4.9768779427e-02f, a9 = -3.6531571299e-02f, a10 = 1.6285819933e-02f;
  const int32_t hx = fbits(x);
  const int32_t ix = hx & 0x7fffffff;
  int id;
  float hi = 0.f, lo = 0.f;
  if (ix >= 0x4c000000) {              // |x| >= 2^25
    if (ix > 0x7f800000) return fadd(x, x);   // NaN
    return hx > 0 ? fadd(hi3, lo3) : fsub(-hi3, lo3);
  }
  if (ix < 0x3ee00000) {               // |x| < 0.4375
    if (ix < 0x31000000) return x;     // |x| < 2^-29
    id = -1;
  } else {
    x = fromBits(ix);                  // fabsf
    if (ix < 0x3f980000) {             // |x| < 1.1875
      if (ix < 0x3f300000) { id = 0; hi = hi0; lo = lo0; x = fdiv(fsub(fmul(2.0f, x), 1.0f), fadd(2.0f, x)); }   // 7/16 <= |x| < 11/16
      else { id = 1; hi = hi1; lo = lo1; x = fdiv(fsub(x, 1.0f), fadd(x, 1.0f)); }                               // 11/16 <= |x| < 19/16
    } else {
      if (ix < 0x401c0000) { id = 2; hi = hi2; lo = lo2; x = fdiv(fsub(x, 1.5f), fadd(1.0f, fmul(1.5f, x))); }   // |x| < 2.4375
      else { id = 3; hi = hi3; lo = lo3; x = fdiv(-1.0f, x); }
    }
  }
  const float z = fmul(x, x);
  const float w = fmul(z, z);
  // the sum a_i z^(i+1), i = 0..10, split into its odd and even parts
  const float s1 = fmul(z, fadd(a0, fmul(w, fadd(a2, fmul(w, fadd(a4, fmul(w, fadd(a6, fmul(w, fadd(a8, fmul(w, a10)))))))))));
  const float s2 = fmul(w, fadd(a1, fmul(w, fadd(a3, fmul(w, fadd(a5, fmul(w, fadd(a7, fmul(w, a9)))))))));
  if (id < 0) return fsub(x, fmul(x, fadd(s1, s2)));
  const float r = fsub(hi, fsub(fsub(fmul(x, fadd(s1, s2)), lo), x));
  return hx < 0 ? -r : r;
}

// e_atan2f.c
VB_FD_FN float atan2f_fd(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  const int32_t hx = fbits(x), hy = fbits(y);
  const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return fadd(x, y);   // NaN
  if (hx == 0x3f800000) return atanf_fd(y);                      // x == 1
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);             // 2 * sign(x) + sign(y)
  if (iy == 0) {                                                 // y == 0
    if (m < 2) return y;
    return m == 2 ? fadd(pi, tiny) : fsub(-pi, tiny);
  }
  if (ix == 0) return hy < 0 ? fsub(-pi_o_2, tiny) : fadd(pi_o_2, tiny);
  if (ix == 0x7f800000) {                                        // x infinite
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return fadd(pi_o_4, tiny);
        case 1: return fsub(-pi_o_4, tiny);
        case 2: return fadd(fmul(3.0f, pi_o_4), tiny);
        default: return fsub(fmul(-3.0f, pi_o_4), tiny);
      }
    }
    switch (m) {
      case 0: return 0.0f;
      case 1: return -0.0f;
      case 2: return fadd(pi, tiny);
      default: return fsub(-pi, tiny);
    }
  }
  if (iy == 0x7f800000) return hy < 0 ? fsub(-pi_o_2, tiny) : fadd(pi_o_2, tiny);
  const int k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = fadd(pi_o_2, fmul(0.5f, pi_lo));               // |y / x| > 2^60
  else if (hx < 0 && k < -60) z = 0.0f;                          // |y| / x < -2^60
  else z = atanf_fd(fromBits(fbits(fdiv(y, x)) & 0x7fffffff));
  switch (m) {
    case 0: return z;
    case 1: return fromBits(fbits(z) ^ (int32_t)0x80000000);
    case 2: return fsub(pi, fsub(z, pi_lo));
    default: return fsub(fsub(z, pi_lo), pi);
  }
}

}  // namespace vb_fdlibm
