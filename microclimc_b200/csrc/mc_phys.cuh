// Device-side physics primitives for the microclimc hot path (FP64, sm_100a).
//
// These are the worker functions the reference obtains from `microctools` (absent from the reference
// tree; SURVEY.md Appendix A gives every call site) plus the inline Tetens / latent-heat expressions of
// R/code.R.  They are written for the GPU: functions are split into a column-invariant part (evaluated
// once per column, `*_base`) and a per-step part, pow(x, 0.25) / pow(x, 1.75) are built from square
// roots, and nothing allocates.  Semantics are those documented in DESIGN.md §"compat layer".
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define MC_HD __host__ __device__ __forceinline__
#define MC_NI __host__ __device__ __noinline__      // large bodies: one copy in the instruction stream (I-cache)
#else
#define MC_HD inline
#define MC_NI inline
#endif
// CTA-wide rendezvous that keeps the warps of a block on the same stretch of code, so that one instruction-cache
// fill serves all of them (the hot loop is larger than the 32 KB L1.5 I-cache). No-op in the host (test) build.
#if defined(__CUDA_ARCH__)
#define MC_CTA_SYNC() __syncthreads()
#else
#define MC_CTA_SYNC() ((void)0)
#endif

namespace mc {

constexpr double kPi = 3.14159265358979323846;
constexpr double kSb = 5.67 * 1e-8;

// ---- fast FP64 primitives ------------------------------------------------------------------------
// Division and exp dominate the instruction stream of this path (profiles/README.md r01a: 1 445 IEEE divisions
// per column-step at ~20 instructions each, plus ~350 exp calls whose polynomial constants are rebuilt with
// move pairs).  On the device:
//  * frcp/fdiv: MUFU.RCP64H seed (rcp.approx.ftz.f64) + 2 Newton steps + one residual correction, branch-free,
//    <= 1 ulp from the correctly rounded quotient for normal, finite, non-zero operands. Sites whose operands can
//    legitimately be 0 / Inf (stomatal conductance at night) keep the IEEE `/`.
//  * fexp: Cody-Waite reduction to |r| <= ln2/128 with a 64-entry table of 2^(j/64) in shared memory + degree-5
//    polynomial, exponent insertion by integer add; relative error < 1.5 ulp; below -708 it returns e^-708 = 3e-308.
// The host build (tests) maps them to the libm / IEEE operations.
#if defined(__CUDACC__)
// 2^(j/64), j = 0..63, correctly rounded.  Every kernel that evaluates fexp copies the table to shared memory first
// (exp_tab_init): a per-lane index into constant memory would serialise, shared memory serves it in one or two wavefronts.
__constant__ double c_exp_tab[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};
__shared__ double s_exp_tab[64];
__device__ __forceinline__ void exp_tab_init() {
  if (threadIdx.x < 64) s_exp_tab[threadIdx.x] = c_exp_tab[threadIdx.x];
  __syncthreads();
}
#endif
#ifndef MC_RCP_CUBIC
#define MC_RCP_CUBIC 1
#endif
MC_HD double frcp(double b) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
#if MC_RCP_CUBIC
  // one third-order step instead of two Newton steps: the seed is good to 2^-23, so r (1 + e + e^2) is good to 2^-69 and the last
  // fma rounds once (0.502 ulp worst case in exact-arithmetic simulation, 0.5 for the quotient after fdiv's residual step);
  // 3 FP64 operations on a chain of 3 instead of 4 on a chain of 4 — at ~640 reciprocals per column-step, 4 % of the FP64 work
  const double e = fma(-b, r, 1.0);
  const double p = fma(e, e, e);
  return fma(r, p, r);
#else
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  return r;
#endif
#else
  return 1.0 / b;
#endif
}
#ifndef MC_FASTDIV
#define MC_FASTDIV 1
#endif
MC_HD double fdiv(double a, double b) {
#if defined(__CUDA_ARCH__)
  double r = frcp(b);
  double q = a * r;
#if MC_FASTDIV
  // q = a * (1/b) with the 0.5-ulp reciprocal: <= 1.5 ulp.  The exact-residual step fma(fma(-q, b, a), r, q) that made it 0.5 ulp is
  // two more dependent FP64 operations on a chain whose latency is what bounds the kernels (+3.3 % on the bench without it, r02n).
  return q;
#else
  return fma(fma(-q, b, a), r, q);
#endif
#else
  return a / b;
#endif
}
MC_HD double fexp(double x) {
#if defined(__CUDA_ARCH__)
  // x = (64 k + j) ln2/64 + r, |r| <= ln2/128:  exp(x) = 2^k * 2^(j/64) * exp(r), exp(r) - 1 by a degree-5 polynomial
  // (remainder r^6/720 < 4e-17).  10 FP64 operations on a dependent chain of 7 instead of 20 on a chain of 9.
  // Arguments below -708 are clamped first (the result is then e^-708 = 3e-308 instead of 0: the exponent insertion below must not
  // underflow); a NaN passes the clamp, gives n = 0 and comes out as NaN.  One compare + select here instead of three at the end.
  const double xc = (x < -708.0) ? -708.0 : x;
  double t = fma(xc, 9.233248261689366e+01, 6755399441055744.0);       // 64/ln2; round-to-nearest integer in the low word
  const int n = __double2loint(t);
  const double kd = t - 6755399441055744.0;
  double r = fma(kd, -0x1.62e42feep-7, xc);                            // ln2/64, high part (trailing zeros: kd * hi is exact)
  r = fma(kd, -0x1.a39ef35793c76p-39, r);                              // ln2/64, low part
  const double T = s_exp_tab[n & 63];
  const double r2 = r * r;
  const double a = fma(8.3333333333333332e-03, r, 4.1666666666666664e-02), b = fma(1.6666666666666666e-01, r, 0.5);
  const double d = fma(fma(a, r2, b), r2, r);                          // exp(r) - 1
  const double p = fma(T, d, T);
  double res = __hiloint2double(__double2hiint(p) + ((n >> 6) << 20), __double2loint(p));
  if (x > 709.0) res = __longlong_as_double(0x7ff0000000000000LL);
  return res;
#else
  return exp(x);
#endif
}

// fexp for arguments that are bounded by construction (|x| < 700: the Tetens / Clausius-Clapeyron exponents of physical temperatures,
// attenuation exponents of O(10)): the two range guards of fexp (a compare and two selects at the head of the dependent chain, another
// pair at its end) are dropped.  NaN still comes out as NaN (n = 0).  MC_EXPB 0 maps it back to fexp.
#ifndef MC_EXPB
#define MC_EXPB 1
#endif
MC_HD double fexp_b(double x) {
#if defined(__CUDA_ARCH__) && MC_EXPB
  double t = fma(x, 9.233248261689366e+01, 6755399441055744.0);
  const int n = __double2loint(t);
  const double kd = t - 6755399441055744.0;
  double r = fma(kd, -0x1.62e42feep-7, x);
  r = fma(kd, -0x1.a39ef35793c76p-39, r);
  const double T = s_exp_tab[n & 63];
  const double r2 = r * r;
  const double a = fma(8.3333333333333332e-03, r, 4.1666666666666664e-02), b = fma(1.6666666666666666e-01, r, 0.5);
  const double d = fma(fma(a, r2, b), r2, r);
  const double p = fma(T, d, T);
  return __hiloint2double(__double2hiint(p) + ((n >> 6) << 20), __double2loint(p));
#else
  return fexp(x);
#endif
}

// fexp for arguments that are never positive (attenuation exponents -K * PAI, which can be very negative when the sun is at the
// horizon): only the lower clamp is kept.
MC_HD double fexp_neg(double x) {
#if defined(__CUDA_ARCH__) && MC_EXPB
  return fexp_b((x < -708.0) ? -708.0 : x);
#else
  return fexp(x);
#endif
}

// Natural logarithm for finite positive normal x (the vapour pressures and ratios of this path): x = 2^e * m with
// m in [sqrt(1/2), sqrt(2)), s = (m-1)/(m+1), log m = 2s + s^3 * P(s^2) (atanh series, degree 8 in s^2, Estrin),
// result = e*ln2_hi + (2s + ... + e*ln2_lo); < 2 ulp. x <= 0, Inf, NaN, subnormals fall back to log().
#if defined(__CUDACC__)
__constant__ double c_log_poly[9] = {2.0 / 3, 2.0 / 5, 2.0 / 7, 2.0 / 9, 2.0 / 11, 2.0 / 13, 2.0 / 15, 2.0 / 17, 2.0 / 19};
#endif
MC_HD double flog(double x) {
#if defined(__CUDA_ARCH__)
  int hi = __double2hiint(x);
  if (hi < 0x00100000 || hi >= 0x7ff00000) return log(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  double m = __hiloint2double(hi, __double2loint(x));          // [1, 2)
  if (m > 1.4142135623730951) { m *= 0.5; e += 1; }
  const double sn = m - 1.0, sd = m + 1.0;
  const double s1 = fdiv(sn, sd);
  const double z = s1 * s1, z2 = z * z, z4 = z2 * z2;
  const double p01 = fma(c_log_poly[1], z, c_log_poly[0]), p23 = fma(c_log_poly[3], z, c_log_poly[2]);
  const double p45 = fma(c_log_poly[5], z, c_log_poly[4]), p67 = fma(c_log_poly[7], z, c_log_poly[6]);
  const double q03 = fma(p23, z2, p01), q47 = fma(p67, z2, p45);
  const double P = fma(fma(c_log_poly[8], z4, q47), z4, q03);
  const double ed = (double)e;
  // log m = 2 s1 + s1 * z * P, with the rounding error of s1 = sn/sd folded in: r = (sn - s1*sd)/sd
  const double r = fma(-s1, sd, sn) * frcp(sd);
  const double lo = fma(s1 * z, P, fma(ed, 1.90821492927058770002e-10, 2.0 * r));
  return fma(ed, 6.93147180369123816490e-01, fma(2.0, s1, lo));
#else
  return log(x);
#endif
}

// flog without its out-of-line branch (a straight-line body lets the scheduler interleave several calls): same polynomial; the special
// operands are patched in by selects — NaN and negative -> NaN, 0 -> -Inf, +Inf -> +Inf.  Subnormal operands (never a vapour pressure)
// are treated as 0.
MC_HD double flog_nb(double x) {
#if defined(__CUDA_ARCH__)
  int hi = __double2hiint(x);
  const bool bad = hi < 0x00100000 || hi >= 0x7ff00000;
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  double m = __hiloint2double(hi, __double2loint(x));          // [1, 2)
  const bool big = m > 1.4142135623730951;
  m = big ? m * 0.5 : m; e = big ? e + 1 : e;
  const double sn = m - 1.0, sd = m + 1.0;
  const double s1 = fdiv(sn, sd);
  const double z = s1 * s1, z2 = z * z, z4 = z2 * z2;
  const double p01 = fma(c_log_poly[1], z, c_log_poly[0]), p23 = fma(c_log_poly[3], z, c_log_poly[2]);
  const double p45 = fma(c_log_poly[5], z, c_log_poly[4]), p67 = fma(c_log_poly[7], z, c_log_poly[6]);
  const double q03 = fma(p23, z2, p01), q47 = fma(p67, z2, p45);
  const double P = fma(fma(c_log_poly[8], z4, q47), z4, q03);
  const double ed = (double)e;
  const double r = fma(-s1, sd, sn) * frcp(sd);
  const double lo = fma(s1 * z, P, fma(ed, 1.90821492927058770002e-10, 2.0 * r));
  const double res = fma(ed, 6.93147180369123816490e-01, fma(2.0, s1, lo));
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  const double special = (x != x || x < 0.0) ? __longlong_as_double(0x7ff8000000000000LL) : ((x == inf) ? inf : -inf);
  return bad ? special : res;
#else
  return log(x);
#endif
}

MC_HD double dmin(double a, double b) { return a < b ? a : b; }
MC_HD double dmax(double a, double b) { return a > b ? a : b; }
MC_HD double sq(double x) { return x * x; }
MC_HD double pow4(double x) { double x2 = x * x; return x2 * x2; }
// Square root from the MUFU.RSQ64H seed: two Newton steps on y ~ x^-1/2, s = x*y and one Heron correction; <= 1 ulp for
// finite positive normal x, exact 0 for x == 0, NaN for x < 0 (as sqrt). No branches, no out-of-line slow path.
#ifndef MC_SQRT_CUBIC
#define MC_SQRT_CUBIC 1
#endif
MC_HD double fsqrt(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if MC_SQRT_CUBIC
  // t = x y ~ sqrt(x) to 2^-22; e = 1 - x y^2; sqrt(x) = t (1 + e/2 + 3 e^2/8 + O(e^3)); then one Heron correction with the exact
  // residual, for which the seed's accuracy is enough (the correction is of relative size 2^-53).  8 FP64 operations on a chain of
  // 7 instead of 11 on a chain of 9; 0.4999 ulp worst case in exact-arithmetic simulation, as before.
  const double t = x * y;
  const double e = fma(-t, y, 1.0);
  const double q = e * fma(0.375, e, 0.5);
  const double s0 = fma(t, q, t);
#ifndef MC_SQRT_HERON
#define MC_SQRT_HERON 0
#endif
#if MC_SQRT_HERON
  const double s = fma(0.5 * y, fma(-s0, s0, x), s0);
#else
  const double s = s0;                   // <= 1.5 ulp without the Heron correction (3 operations, 2 chain levels less; r02n)
#endif
  return (x == 0.0) ? 0.0 : s;
#else
  const double hx = 0.5 * x;
  y = y * fma(-hx * y, y, 1.5);
  y = y * fma(-hx * y, y, 1.5);
  double s = x * y;
  s = fma(0.5 * y, fma(-s, s, x), s);
  return (x == 0.0) ? 0.0 : s;
#endif
#else
  return sqrt(x);
#endif
}
// fsqrt for operands that are never zero (a product of strictly positive factors): without the x == 0 patch (a compare and a select pair).
// NaN and negative operands still give NaN; +0 would give NaN instead of 0.
MC_HD double fsqrt_nz(double x) {
#if defined(__CUDA_ARCH__) && MC_SQRT_CUBIC && !MC_SQRT_HERON
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double t = x * y;
  const double e = fma(-t, y, 1.0);
  const double q = e * fma(0.375, e, 0.5);
  return fma(t, q, t);
#else
  return fsqrt(x);
#endif
}
MC_HD double qroot(double x) { return fsqrt(fsqrt(x)); }           // x^0.25
MC_HD double pow175(double x) { double r = fsqrt(x); return x * r * fsqrt(r); }   // x^1.75

// ---- thermodynamics ---------------------------------------------------------------------------
MC_HD double phair(double tc, double pk) { return 44.6 * fdiv(pk, 101.3) * fdiv(273.15, tc + 273.15); }
MC_HD double cpair(double tc) { return 2e-05 * tc * tc + 0.0002 * tc + 29.119; }
MC_HD double tetens(double tc) { return 0.6108 * fexp_b(fdiv(17.27 * tc, tc + 237.3)); }   // code.R:156,382,410,...
MC_HD double phair_c(double c, double tc) { return c * fdiv(273.15, tc + 273.15); }     // c = 44.6 * (pk / 101.3)
MC_HD double satvap(double tc) {                                   // ice = TRUE
  double e0, L;
  if (tc < 0) { e0 = 610.78 / 1000; L = 2.834e6; } else { e0 = 611.2 / 1000; L = 2.501e6 - 2340 * tc; }
  return e0 * fexp_b((L * (1 / 461.5)) * (1 / 273.15 - frcp(tc + 273.15)));
}
MC_HD double dewpoint(double ea, double tc) {
  double e0, L;
  if (tc < 0) { e0 = 610.78 / 1000; L = 2.834e6; } else { e0 = 611.2 / 1000; L = 2.501e6 - 2340 * tc; }
  double it = 1 / 273.15 - fdiv(461.5, L) * flog(fdiv(ea, e0));
  return frcp(it) - 273.15;
}

// ---- aerodynamics -----------------------------------------------------------------------------
MC_HD double zeroplanedis(double hgt, double PAI) {                // NMR.R:103-107
  if (PAI < 0.001) PAI = 0.001;
  double s = sqrt(7.5 * PAI);
  return (1 - (1 - exp(-s)) / s) * hgt;
}
MC_HD double roughlength_d(double hgt, double PAI, double d, double zm0) {   // NMR.R:109-113 (+ floor)
  double Be = sqrt(0.003 + (0.2 * PAI) / 2);
  double zm = (hgt - d) * exp(-0.4 / Be);
  return zm < zm0 ? zm0 : zm;
}
MC_HD double roughlength(double hgt, double PAI, double zm0) { return roughlength_d(hgt, PAI, zeroplanedis(hgt, PAI), zm0); }
MC_HD double mixinglength(double hgt, double PAI) {
  double d = zeroplanedis(hgt, PAI);
  double zm = roughlength_d(hgt, PAI, d, 0.003);
  return (0.32 * (hgt - d)) / log((hgt - d) / zm);
}
// attencoef = attencoef_base * sqrt(phi_m); the base is column- and layer-invariant in time
MC_HD double attencoef_base(double hgt, double PAI, double cd, double iw, double l_m) {
  return sqrt((cd * PAI * hgt) / (2 * l_m * iw));
}
MC_HD void diabatic_cor(double tc, double pk, double H, double uf, double zi, double d, double& psi_m, double& psi_h) {
  double Tk = tc + 273.15;
  double st = -fdiv(0.4 * 9.81 * (zi - d) * H, phair(tc, pk) * cpair(tc) * Tk * (uf * uf * uf));
  if (st < 0) { psi_h = -2 * flog((1 + fsqrt(1 - 16 * st)) / 2); psi_m = 0.6 * psi_h; }
  else { psi_h = 6 * flog(1 + st); psi_m = psi_h; }
  if (psi_m > 4) psi_m = 4;
  if (psi_h > 4) psi_h = 4;
  if (psi_m < -1.5) psi_m = -1.5;
  if (psi_h < -2.5) psi_h = -2.5;
}
// within-canopy conductance given precomputed mixing length and attenuation coefficient `a`
MC_HD double gcanopy_a(double uh, double z1, double z0, double tc1, double tc0, double hgt, double l_m, double a,
                       double iw, double phi_m, double pk) {
  double ph = phair((tc1 + tc0) / 2, pk);
  double e0 = fexp(-a * (fdiv(z0, hgt) - 1));
  double e1 = fexp(-a * (fdiv(z1, hgt) - 1));
  double g = fdiv(l_m * iw * ph * uh * a, hgt * (e0 - e1) * phi_m);
  double gfree = 0.0012 * ph * qroot(fdiv(fabs(tc1 - tc0), fabs(z1 - z0)));
  return (g > gfree) ? g : gfree;
}
MC_HD double gcanopy(double uh, double z1, double z0, double tc1, double tc0, double hgt, double PAI, double cd,
                     double iw, double phi_m, double pk) {
  double l_m = mixinglength(hgt, PAI);
  double a = attencoef_base(hgt, PAI, cd, iw, l_m) * sqrt(phi_m);
  return gcanopy_a(uh, z1, z0, tc1, tc0, hgt, l_m, a, iw, phi_m, pk);
}
MC_HD double gturb(double u, double zu, double z1, double z0, double hgt, double PAI, double tc, double psi_m,
                   double psi_h, double zm0, double pk) {
  double d = zeroplanedis(hgt, PAI);
  double zm = roughlength_d(hgt, PAI, d, zm0);
  double zh = 0.2 * zm;
  double ln = log((zu - d) / zm) + psi_m;
  if (!(ln >= 0.55)) ln = 0.55;
  double uf = (0.4 * u) / ln;
  double lo = z0 - d;
  if (!(lo > zh)) lo = zh;
  double lnh = log((z1 - d) / lo);
  double den = lnh + psi_h;
  if (den < 0.25 * lnh) den = 0.25 * lnh;
  return (0.4 * phair(tc, pk) * uf) / den;
}
// Prandtl number of air in gforcedfree: nu / Dh = 13.3 / 18.9 exactly (both carry the same temperature / pressure factor)
constexpr double kPrAir = 0.7037037037037037, kCbrtPrAir = 0.8894672162406483;
// Cf^2 and Cr^4 of the product form of gforcedfree (gha_fast2 in mc_fast.cuh, gha_hour in mc_steady.cuh):
// Cf = 0.664 * 18.9e-6 * Pr^(1/3) / sqrt(13.3e-6), Cr = 0.54 * 18.9e-6 * Pr^(1/4) / sqrt(13.3e-6)
constexpr double kGhaCf2 = 9.3684559114155109e-06, kGhaCr4 = 4.3162720585010526e-11;
MC_HD double gforcedfree(double d, double u, double tc, double dtc, double pk, double dtmin) {
  double Tk = tc + 273.15;
  double ph = phair(tc, pk);
  double sc = pow175(Tk * (1 / 273.15)) * fdiv(101.3, pk);
  double Dh = 18.9e-6 * sc;
  double nu = 13.3e-6 * sc;
  double adt = fabs(dtc);
  if (adt < dtmin) adt = dtmin;
  const double id = frcp(d);
  double Re = fdiv(u * d, nu);
  double Pr = fdiv(nu, Dh);
  double Gr = fdiv(9.81 * (d * d * d) * adt, Tk * (nu * nu));
  double gforced = (0.664 * Dh * ph * fsqrt(Re) * cbrt(Pr)) * id;
  double gfree = (0.54 * Dh * ph * qroot(Gr * Pr)) * id;
  return (gforced > gfree) ? gforced : gfree;
}
MC_HD double layercond(double Rsw, double gsmax, double q50) { double Qa = 4.6 * Rsw; return fdiv(gsmax * Qa, Qa + q50); }
MC_HD double soilrh(double theta, double b, double Psie, double Smax, double tc) {
#if defined(__CUDA_ARCH__)
  double matric = -Psie * fexp(-b * flog(fdiv(theta, Smax)));          // (theta / Smax)^-b
#else
  double matric = -Psie * pow(theta / Smax, -b);
#endif
  return fexp(fdiv(0.018 * matric, 8.31 * (tc + 273.15)));
}

// ---- sun / radiation --------------------------------------------------------------------------
// Solar altitude (microctools solalt; degrees).  Its nine trigonometric calls split three ways: five depend on the date only
// (equation of time, declination), two on the latitude only, and only cos(hour angle) and the atan depend on both — the steady
// kernel takes the date part from a per-hour table and the latitude part from the column's preparation (34 % of its stall samples
// sat in these calls, profiles r01q), the streaming kernel likewise.  Same operations in the same order as the one-piece form.
struct SolDay { double eot, sdec, cdec; };
MC_HD SolDay sol_day(double jd) {
  const double mm = 6.24004077 + 0.01720197 * (jd - 2451545);
  SolDay d;
  d.eot = -7.659 * sin(mm) + 9.863 * sin(2 * mm + 3.5932);
  const double declin = (kPi * 23.5 / 180) * cos(2 * kPi * fdiv(jd - 159.5, 365.25));
  d.sdec = sin(declin); d.cdec = cos(declin);
  return d;
}
MC_HD void sol_site(double lat, double& slat, double& clat) {
  const double latr = fdiv(lat * kPi, 180);
  slat = sin(latr); clat = cos(latr);
}
MC_HD double solalt_pre(double lt, double lon, double merid, double dst, const SolDay& d, double slat, double clat) {
  double st = lt + fdiv(4 * (lon - merid) + d.eot, 60) - dst;
  double tt = 0.261799 * (st - 12);
  double sh = d.sdec * slat + d.cdec * clat * cos(tt);
  return fdiv(180 * atan(fdiv(sh, fsqrt(1 - sh * sh))), kPi);
}
// The streaming and steady kernels never need the ANGLE: everything downstream of solalt is either a sign test, the 5-degree
// test, sin(sa) (difprop) or tan(90 - sa) (cank) — all functions of sh = sin(sa), which is what the formula produces before its
// atan.  Working on sh removes the atan -> degrees -> radians -> tan / sin round trips (two libm calls of ~70 serial instructions
// each per use; algebraically identical, numerically within a few ulp of the round trip, far inside the parity tolerance).
// The general column (mc_step.cuh) and the oracle keep the literal form.
constexpr double kSin5 = 0.08715574274765817;          // sin(5 degrees)
MC_HD double solsin_pre(double lt, double lon, double merid, double dst, const SolDay& d, double slat, double clat) {
  double st = lt + fdiv(4 * (lon - merid) + d.eot, 60) - dst;
  double tt = 0.261799 * (st - 12);
  return d.sdec * slat + d.cdec * clat * cos(tt);
}
// cank2 on sh: tan(90 deg - sa) = cos(sa) / sin(sa)
MC_HD void cank2_sh(double x, double sh, double den, double K5, double& K, double& Kc) {
  // sqrt(x^2 + cot^2) with cot^2 = (1 - sh^2) / sh^2 is sqrt((x^2 - 1) sh^2 + 1) / sh for sh > 0 (the only case Kf is selected in):
  // one square root and one division instead of two and two
  const double Kf = fdiv(fsqrt(fma(x * x - 1, sh * sh, 1.0)), sh * den);
  K = (sh > 0) ? Kf : 0.0;
  Kc = (sh < kSin5) ? K5 : Kf;
}
MC_HD double cank_sh(double x, double sh, double den) {
  const double cot = fdiv(fsqrt(1 - sh * sh), sh);
  return fdiv(fsqrt(x * x + cot * cot), den);
}
MC_HD double solalt(double lt, double lat, double lon, double jd, double merid, double dst) {
  double sl, cl;
  sol_site(lat, sl, cl);
  return solalt_pre(lt, lon, merid, dst, sol_day(jd), sl, cl);
}
MC_HD double difprop(double Rsw, double sa) {
  if (!(sa > 0) || !(Rsw > 0)) return 1.0;
  double kt = fdiv(Rsw, 1352.0 * sin(fdiv(sa * kPi, 180)));
  if (kt > 1) kt = 1;
  if (kt <= 0.22) return 1 - 0.09 * kt;
  if (kt <= 0.8) return 0.9511 - 0.1604 * kt + 4.388 * kt * kt - 16.638 * kt * kt * kt + 12.336 * kt * kt * kt * kt;
  return 0.165;
}
MC_HD double difprop_sh(double Rsw, double sh) {          // difprop with sin(sa) given
  if (!(sh > 0) || !(Rsw > 0)) return 1.0;
  double kt = fdiv(Rsw, 1352.0 * sh);
  if (kt > 1) kt = 1;
  if (kt <= 0.22) return 1 - 0.09 * kt;
  if (kt <= 0.8) return 0.9511 - 0.1604 * kt + 4.388 * kt * kt - 16.638 * kt * kt * kt + 12.336 * kt * kt * kt * kt;
  return 0.165;
}
MC_HD double cank_den(double x) { return x + 1.774 * pow(x + 1.182, -0.733); }   // column-invariant
MC_HD double cank(double x, double sa, double den) {
  double t = tan(fdiv((90 - sa) * kPi, 180));
  return fdiv(fsqrt(x * x + t * t), den);
}
MC_HD double clumptr(double tr, double clump) { return clump + (1 - clump) * tr; }
// shortwave under plant area l: `s` = sqrt(1 - ref), `trd` = clumped diffuse transmission exp(-s*l),
// `K` = beam extinction (unused when sa <= 0)
// cank for sa (0 when the sun is down) and for max(sa, 5 degrees) with ONE tangent: the second equals the first whenever
// sa >= 5 and is the column constant K5 = cank(x, 5, den) otherwise.  Values identical to two separate calls.
MC_HD void cank2(double x, double sa, double den, double K5, double& K, double& Kc) {
  const double Kf = cank(x, sa, den);
  K = (sa > 0) ? Kf : 0.0;
  Kc = (sa < 5) ? K5 : Kf;
}
MC_HD double cansw_k(double globrad, double dp, double sa, double l, double s, double trd, double K, double clump) {
  double trb = (sa > 0) ? clumptr(fexp(-s * K * l), clump) : clumptr(0.0, clump);
  return globrad * (dp * trd + (1 - dp) * trb);
}
MC_HD double psunlit_k(double l, double sa, double K, double clump) {
  if (!(sa > 0)) return clumptr(0.0, clump);
  return clumptr(fexp(-K * l), clump);
}

}  // namespace mc
