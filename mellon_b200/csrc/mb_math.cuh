// Lean float64 transcendental sequences for the covariance epilogue (K1 / K7).
//
// The FP64 pipe of sm_100a issues 64 DFMA / clk / SM and is what bounds the fused covariance
// build once the x.y contraction is out of the way, so the epilogue counts instructions:
//   * exp(-r): 64-entry table of 2^(j/64) + degree-5 polynomial on |t| <= ln2/128 (11 FP64 ops,
//     relative error < 2.5e-16) instead of libdevice's table-free degree-11 sequence (~25 ops);
//   * sqrt(s): MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, ~2^-22) + two coupled Newton steps
//     (6 FP64 ops, < 1 ulp) instead of libdevice's IEEE-rounded sequence with its slow-path branch.
// Both are compiled for the host as well (the seed becomes 1/sqrtf) so tests/ can check their
// accuracy on the CPU against libm over the whole argument range the kernels can see.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MB_HD __host__ __device__ __forceinline__
#else
#define MB_HD static inline
#endif

// 2^(j/64), j = 0..63, correctly rounded
#define MB_EXP2_TABLE_INIT                                                                          \
  {0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,          \
   0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,          \
   0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,          \
   0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,          \
   0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,          \
   0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,          \
   0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,          \
   0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,          \
   0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,          \
   0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,          \
   0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,          \
   0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,          \
   0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,          \
   0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,          \
   0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,          \
   0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0}

namespace mbmath {

MB_HD double hi_lo(int32_t hi, int32_t lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d;
  memcpy(&d, &u, 8);
  return d;
#endif
}
MB_HD int32_t hi_word(double d) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(d);
#else
  uint64_t u;
  memcpy(&u, &d, 8);
  return (int32_t)(u >> 32);
#endif
}
MB_HD int32_t lo_word(double d) {
#if defined(__CUDA_ARCH__)
  return __double2loint(d);
#else
  uint64_t u;
  memcpy(&u, &d, 8);
  return (int32_t)(uint32_t)u;
#endif
}

// exp(-r) for r >= 0 (r > 708 returns 0: the result would be subnormal).  `tab` = MB_EXP2_TABLE_INIT
// (shared memory in the kernels).
MB_HD double exp_neg(double r, const double* tab) {
  const double MAGIC = 6755399441055744.0;          // 1.5 * 2^52: low word of (v + MAGIC) is rint(v)
  const double K64 = 92.332482616893656877;         // 64 / ln 2
  const double C_HI = 0x1.62e42fefa3800p-7;         // ln2/64 with the low 11 mantissa bits cleared
  const double C_LO = 8.59050471673183e-16;         // ln2/64 - C_HI
  double t = fma(r, -K64, MAGIC);
  const int32_t n = lo_word(t);                     // n = rint(-r * 64/ln2) <= 0
  const double nn = t - MAGIC;
  double rem = fma(nn, -C_HI, -r);                  // -r - n ln2/64, |rem| <= ln2/128
  rem = fma(nn, -C_LO, rem);
  double p = fma(rem, 1.0 / 120.0, 1.0 / 24.0);
  p = fma(p, rem, 1.0 / 6.0);
  p = fma(p, rem, 0.5);
  p = fma(p, rem, 1.0);
  p = fma(p, rem, 1.0);
  const double tj = tab[n & 63];
  const int32_t m = n >> 6;                         // floor(n / 64)
  const double sc = hi_lo(hi_word(tj) + (m << 20), lo_word(tj));  // tj * 2^m (normal: m >= -1022)
  const double v = sc * p;
  return (hi_word(r) > 0x40862000) ? 0.0 : v;     // r > 708 (integer compare: keeps the FP64 pipe free)
}

// max(s, 1e-300) by an integer compare of the high word (negative doubles compare below as signed ints)
MB_HD double clamp_tiny(double s) { return (hi_word(s) < 0x01a56e1f) ? 1e-300 : s; }

// sqrt(s) for s in [1e-300, 1e300]
MB_HD double sqrt_pos(double s) {
  double y;
#if defined(__CUDA_ARCH__)
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
#else
  {  // host stand-in for the hardware seed: ~2^-22 relative
    int e;
    double mant = frexp(s, &e);
    if (e & 1) { mant *= 2.0; e -= 1; }
    y = ldexp((double)(1.0f / sqrtf((float)mant)), -e / 2);
  }
#endif
  double g = s * y;
  const double h0 = hi_lo(hi_word(y) - (1 << 20), lo_word(y));  // y / 2
  double e1 = fma(-h0, g, 0.5);
  g = fma(g, e1, g);
  const double h = fma(h0, e1, h0);
  const double d = fma(-g, g, s);
  return fma(d, h, g);
}

}  // namespace mbmath
