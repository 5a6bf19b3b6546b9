// c2ray_fastmath.cuh -- FP64 elementary functions without IEEE special-case scaffolding, for the argument ranges
// that occur on the hot path (positive normal inputs to log, |x| < 700 for exp, normal divisors).  The general-purpose
// CUDA log10/pow/exp and `/` cost 50-250 instructions each, most of it range checks and slow-path branches; these
// cost 5-35 and are branch-free.  Measured against libdevice (tools/check_fastmath.cu, 3e8 samples): fdiv equals a/b in
// every sample, fast_log10 is within 1.8e-15 absolute on [1e-20, 1e4] (1.5e-13 of a table row), fast_log within 2.2e-16
// relative, fast_exp within 4.4e-16 relative on |x| <= 700 -- the order of the libm/libdevice differences the parity
// tolerance (1e-8) already absorbs.
#pragma once

namespace c2 {

// 1/b : MUFU.RCP64H seed (measured relative error <= 2^-19.9) + one cubic step (e + e^2: error e^3 ~ 2^-60, then the
// rounding of two FMAs).  Measured over 3e8 log-uniform b in [1e-30, 1e30]: max |1 - b r| = 1.00 x 2^-53, the same as with
// a further Newton step (tools/check_fastmath.cu), which therefore is not taken.
__device__ __forceinline__ double fast_rcp(double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  e = fma(e, e, e);
  return fma(r, e, r);
}

// a/b to ~1 ulp (not correctly rounded), no slow path
__device__ __forceinline__ double fdiv(double a, double b) {
  const double r = fast_rcp(b);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}

// a/b given r ~ 1/b: one residual correction, result within ~0.5 ulp of a/b
__device__ __forceinline__ double fdiv_r(double a, double b, double r) {
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}

// Full-width FP64 literals of the band loop, kept in the constant bank: written as literals the compiler materialises
// each one with two UMOVs in front of every use (5.5 % of the sweep's instructions, profiles/), as constant-bank operands
// they cost nothing.  [0] log10(2)  [1] 2 log10(e)  [2] ln 2  [3] 1e-20  [4] (double)1.0e-7f  [5] (double)1.0e-4f
// [6] 1/dlogtau = 2000/24
__constant__ double d_lit[7] = {0.30102999566398119521, 0.86858896380650365530, 0.69314718055994530942, 1.0e-20,
                                (double)1.0e-7f, (double)1.0e-4f, 2000.0 / 24.0};

// 1/(2k+1), k = 9..1 : atanh series
__constant__ double d_logc[9] = {1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0,
                                 1.0 / 9.0,  1.0 / 7.0,  1.0 / 5.0,  1.0 / 3.0};

// x = 2^e * m, m in [sqrt(1/2), sqrt(2)) ; returns atanh(s), s = (m-1)/(m+1), i.e. log(m)/2 ; positive normal x only
__device__ __forceinline__ double log_core(double x, int& e) {
  int hi = __double2hiint(x);
  const int lo = __double2loint(x);
  e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  const int big = hi >= 0x3ff6a09f;  // m >= ~sqrt(2)
  hi -= big << 20;                   // m *= 0.5
  e += big;
  const double m = __hiloint2double(hi, lo);
  const double s = (m - 1.0) * fast_rcp(m + 1.0);
  const double z = s * s;
  // sum_{k=1..9} z^k/(2k+1), Estrin scheme (dependency depth 4 instead of 9)
  const double z2 = z * z, z4 = z2 * z2;
  const double q0 = fma(d_logc[7], z, d_logc[8]);
  const double q1 = fma(d_logc[5], z, d_logc[6]);
  const double q2 = fma(d_logc[3], z, d_logc[4]);
  const double q3 = fma(d_logc[1], z, d_logc[2]);
  const double r0 = fma(q1, z2, q0), r1 = fma(q3, z2, q2);
  const double p = fma(fma(d_logc[0], z4, r1), z4, r0) * z;  // atanh(s)/s - 1
  return fma(s, p, s);
}
__device__ __forceinline__ double fast_log10(double x) {
  int e;
  const double l = log_core(x, e);
  return fma((double)e, d_lit[0], l * d_lit[1]);  // e*log10(2) + 2 atanh(s) log10(e)
}
__device__ __forceinline__ double fast_log(double x) {
  int e;
  const double l = log_core(x, e);
  return fma((double)e, d_lit[2], l + l);
}

// exp(x) for |x| <= ~700 (clamped below at -700: exp(-700) ~ 1e-304 stands in for 0): k = rint(x/ln2),
// r = x - k ln2 (two-word ln2), degree-13 Taylor polynomial on |r| <= 0.347 (truncation 5e-18), scale by 2^k
// [0] log2(e)  [1],[2] -ln2 in two words  [3..14] 1/2! .. 1/13! (constant-bank operands, see d_lit)
__constant__ double d_expc[15] = {1.4426950408889634074, -6.93147180369123816490e-01, -1.90821492927058770002e-10,
                                  0.5, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0,
                                  1.0 / 362880.0, 1.0 / 3628800.0, 1.0 / 39916800.0, 1.0 / 479001600.0, 1.0 / 6227020800.0};
__device__ __forceinline__ double fast_exp(double x) {
  x = fmax(x, -700.0);
  const double kf = rint(x * d_expc[0]);
  double r = fma(kf, d_expc[1], x);
  r = fma(kf, d_expc[2], r);
  // Estrin on 1 + r + r^2/2! + ... + r^13/13!
  const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  const double p0 = fma(r, 1.0, 1.0);
  const double p1 = fma(r, d_expc[4], d_expc[3]);
  const double p2 = fma(r, d_expc[6], d_expc[5]);
  const double p3 = fma(r, d_expc[8], d_expc[7]);
  const double p4 = fma(r, d_expc[10], d_expc[9]);
  const double p5 = fma(r, d_expc[12], d_expc[11]);
  const double q0 = fma(p1, r2, p0), q1 = fma(p3, r2, p2), q2 = fma(p5, r2, p4);
  const double p6 = fma(r, d_expc[14], d_expc[13]);
  const double s = fma(fma(p6, r4, q2), r8, fma(q1, r4, q0));
  const int k = (int)kf;
  // 2^k * s by exponent arithmetic (s in [0.7, 1.5], k in [-1010, 1010])
  return __hiloint2double(__double2hiint(s) + (k << 20), __double2loint(s));
}

// x^y for positive normal x
__device__ __forceinline__ double fast_pow(double x, double y) { return fast_exp(y * fast_log(x)); }

}  // namespace c2
