// exact_math.cuh -- IEEE-exact float helpers for the point->cell mapping.
//
// The reference bins every point with `atan2(y, x)` on floats (ground_removal.cpp:70), i.e. the host
// libm's atan2f.  glibc 2.39's atan2f is the classic fdlibm single-precision algorithm (e_atan2f.c /
// s_atanf.c): it is NOT correctly rounded, so neither CUDA's atan2f nor (float)atan2(double) reproduces
// its bits, and a 1-ulp difference flips the 80-way channel index of a few points per thousand frames.
// This header restates that algorithm with explicit round-to-nearest, FMA-free operations so the device
// result is bit-identical to the host libm the reference links against.
// The same code compiles for the host so the CPU test-suite can compare it with libm directly
// (lmot_selftest_atan2f in api.cu).
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>

#if defined(__CUDACC__)
#define LMOT_HD __host__ __device__ __forceinline__
#else
#define LMOT_HD inline
#endif

namespace lmot {

// --- explicit non-contracted IEEE binary32 ops -------------------------------------------------------
LMOT_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  volatile float r = a * b; return r;
#endif
}
LMOT_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  volatile float r = a + b; return r;
#endif
}
LMOT_HD float fsub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  volatile float r = a - b; return r;
#endif
}
LMOT_HD float fdiv(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  volatile float r = a / b; return r;
#endif
}
LMOT_HD float fsqrt(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return __builtin_sqrtf(a);
#endif
}
LMOT_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
LMOT_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}

// --- fdlibm s_atanf.c ---------------------------------------------------------------------------------
LMOT_HD float atanf_fdlibm(float x) {
  const float atanhi[4] = {u2f(0x3eed6338u), u2f(0x3f490fdau), u2f(0x3f7b985eu), u2f(0x3fc90fdau)};
  const float atanlo[4] = {u2f(0x31ac3769u), u2f(0x33222168u), u2f(0x33140fb4u), u2f(0x33a22168u)};
  const float aT0 = u2f(0x3eaaaaabu), aT1 = u2f(0xbe4ccccdu), aT2 = u2f(0x3e124925u), aT3 = u2f(0xbde38e38u),
              aT4 = u2f(0x3dba2e6eu), aT5 = u2f(0xbd9d8795u), aT6 = u2f(0x3d886b35u), aT7 = u2f(0xbd6ef16bu),
              aT8 = u2f(0x3d4bda59u), aT9 = u2f(0xbd15a221u), aT10 = u2f(0x3c8569d7u);
  const int32_t hx = (int32_t)f2u(x);
  const int32_t ix = hx & 0x7fffffff;
  int id;
  if (ix >= 0x4c000000) {                       // |x| >= 2^25
    if (ix > 0x7f800000) return fadd(x, x);     // NaN
    const float r = fadd(atanhi[3], atanlo[3]);
    return hx > 0 ? r : -r;
  }
  if (ix < 0x3ee00000) {                        // |x| < 0.4375
    if (ix < 0x31000000) {                      // |x| < 2^-29
      if (fadd(1.0e30f, x) > 1.0f) return x;
    }
    id = -1;
  } else {
    x = u2f((uint32_t)ix);                      // fabsf
    if (ix < 0x3f980000) {                      // |x| < 1.1875
      if (ix < 0x3f300000) { id = 0; x = fdiv(fsub(fmul(2.0f, x), 1.0f), fadd(2.0f, x)); }
      else                 { id = 1; x = fdiv(fsub(x, 1.0f), fadd(x, 1.0f)); }
    } else {
      if (ix < 0x401c0000) { id = 2; x = fdiv(fsub(x, 1.5f), fadd(1.0f, fmul(1.5f, x))); }
      else                 { id = 3; x = fdiv(-1.0f, x); }
    }
  }
  const float z = fmul(x, x);
  const float w = fmul(z, z);
  const float s1 = fmul(z, fadd(aT0, fmul(w, fadd(aT2, fmul(w, fadd(aT4, fmul(w, fadd(aT6, fmul(w, fadd(aT8, fmul(w, aT10)))))))))));
  const float s2 = fmul(w, fadd(aT1, fmul(w, fadd(aT3, fmul(w, fadd(aT5, fmul(w, fadd(aT7, fmul(w, aT9)))))))));
  if (id < 0) return fsub(x, fmul(x, fadd(s1, s2)));
  const float zz = fsub(atanhi[id], fsub(fsub(fmul(x, fadd(s1, s2)), atanlo[id]), x));
  return (hx < 0) ? -zz : zz;
}

// --- fdlibm e_atan2f.c --------------------------------------------------------------------------------
LMOT_HD float atan2f_fdlibm(float y, float x) {
  const float tiny = 1.0e-30f;
  const float pi_o_4 = u2f(0x3f490fdbu), pi_o_2 = u2f(0x3fc90fdbu), pi = u2f(0x40490fdbu), pi_lo = u2f(0xb3bbbd2eu);
  const int32_t hx = (int32_t)f2u(x), hy = (int32_t)f2u(y);
  const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return fadd(x, y);   // NaN
  if (hx == 0x3f800000) return atanf_fdlibm(y);                 // x == 1.0
  const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);            // 2*sign(x) + sign(y)
  if (iy == 0) {
    switch (m) {
      case 0: case 1: return y;
      case 2: return fadd(pi, tiny);
      default: return fsub(-pi, tiny);
    }
  }
  if (ix == 0) return (hy < 0) ? fsub(-pi_o_2, tiny) : fadd(pi_o_2, tiny);
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return fadd(pi_o_4, tiny);
        case 1: return fsub(-pi_o_4, tiny);
        case 2: return fadd(fmul(3.0f, pi_o_4), tiny);
        default: return fsub(fmul(-3.0f, pi_o_4), tiny);
      }
    } else {
      switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return fadd(pi, tiny);
        default: return fsub(-pi, tiny);
      }
    }
  }
  if (iy == 0x7f800000) return (hy < 0) ? fsub(-pi_o_2, tiny) : fadd(pi_o_2, tiny);
  const int32_t k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = fadd(pi_o_2, fmul(0.5f, pi_lo));
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = atanf_fdlibm(u2f(f2u(fdiv(y, x)) & 0x7fffffffu));
  switch (m) {
    case 0: return z;
    case 1: return u2f(f2u(z) ^ 0x80000000u);
    case 2: return fsub(pi, fsub(z, pi_lo));
    default: return fsub(fsub(z, pi_lo), pi);
  }
}

// component_clustering.cpp:40-48: cartesian cell (x*250+y) of a point, 0xFFFF outside the ROI; fp32, no contraction
LMOT_HD unsigned cart_cell_of(float x, float y, float roi, int num_grid) {
  const float half = fdiv(roi, 2.f);
  const float xC = fadd(x, half), yC = fadd(y, half);
  if (!(xC >= 0.f && xC < roi && yC >= 0.f && yC < roi)) return 0xFFFFu;
  const int xI = (int)floorf(fdiv(fmul((float)num_grid, xC), roi));
  const int yI = (int)floorf(fdiv(fmul((float)num_grid, yC), roi));
  return (unsigned)(xI * num_grid + yI);
}

}  // namespace lmot
