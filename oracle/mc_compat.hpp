// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement of the `microctools` worker functions that microclimc's hot path calls.
// The real `microctools` package is an un-vendored, un-pinned dependency of the reference
// (/root/reference/DESCRIPTION:23-24,29; NAMESPACE:20-21) and its source is ABSENT from the
// reference tree, so every function below is a *candidate definition* tagged with provenance:
//   [T] restates in-tree code (file:line given)         [V] follows an equation/prose in the vignette
//   [I] implied by inline duplicates in the reference    [R] recollection of the public package /
//   [D] designed here because nothing pins it                Campbell & Norman (1998)
// PARITY UNPINNED: the reference ships no tests or golden vectors and R is not installed here, so
// "oracle vs real R + microctools" cannot be verified in this container.  What IS pinned: the in-tree
// known answers of SURVEY.md Appendix F (Thomas, paraminit, soilinit) — see tests/test_oracle_golden.py.
//
// Plain scalar C++ (no vector temporaries needed at this level); compiled with
// -O2 -ffp-contract=off so the arithmetic is IEEE double in source order.
#pragma once
#include <cmath>
#include <algorithm>

namespace orc {

constexpr double PI = 3.14159265358979323846;   // R's `pi`

// ---- thermodynamics -------------------------------------------------------------------------
// [R] molar density of air (mol m^-3). Call sites: code.R:155,325,438,705; NMR.R:571,610,771,810.
inline double phair(double tc, double pk = 101.3) {
  double tk = tc + 273.15;
  return 44.6 * (pk / 101.3) * (273.15 / tk);
}
// [R] molar specific heat of air (J mol^-1 K^-1). Call sites: code.R:326,439,706,816,857; NMR.R:435.
inline double cpair(double tc) { return 2e-05 * tc * tc + 0.0002 * tc + 29.119; }

// [R]/[I] saturation vapour pressure (kPa), over ice below 0 degC (ice=TRUE default).
// Clausius-Clapeyron form; agrees with the inline Tetens expressions (code.R:156,382,410) to ~0.1 %.
inline void satvap_consts(double tc, bool ice, double& e0, double& L, double& T0) {
  if (ice && tc < 0) { e0 = 610.78 / 1000; L = 2.834e6; }
  else               { e0 = 611.2 / 1000;  L = 2.501e6 - 2340 * tc; }
  T0 = 273.15;
}
inline double satvap(double tc, bool ice = true) {
  double e0, L, T0; satvap_consts(tc, ice, e0, L, T0);
  return e0 * std::exp((L / 461.5) * (1 / T0 - 1 / (tc + 273.15)));
}
// [R]/[I] dew/frost point (degC): inverse of satvap with the constants selected by `tc`.
// Call sites: code.R:737,893; NMR.R:473.
inline double dewpoint(double ea, double tc, bool ice = true) {
  double e0, L, T0; satvap_consts(tc, ice, e0, L, T0);
  double it = 1 / T0 - (461.5 / L) * std::log(ea / e0);
  return 1 / it - 273.15;
}

// ---- aerodynamics ---------------------------------------------------------------------------
// [T] NMR.R:103-107
inline double zeroplanedis(double hgt, double PAI) {
  if (PAI < 0.001) PAI = 0.001;
  double s = std::sqrt(7.5 * PAI);
  return (1 - (1 - std::exp(-s)) / s) * hgt;
}
// [T] NMR.R:109-113 + [R] floor at zm0 (default 0.003)
inline double roughlength(double hgt, double PAI, double zm0 = 0.003) {
  double d = zeroplanedis(hgt, PAI);
  double Be = std::sqrt(0.003 + (0.2 * PAI) / 2);
  double zm = (hgt - d) * std::exp(-0.4 / Be);
  return zm < zm0 ? zm0 : zm;
}
// [R] mixing length (m). Call sites: code.R:392 (layer thickness as "hgt"), 795.
inline double mixinglength(double hgt, double PAI, double zm0 = 0.003) {
  double d = zeroplanedis(hgt, PAI);
  double zm = roughlength(hgt, PAI, zm0);
  return (0.32 * (hgt - d)) / std::log((hgt - d) / zm);
}
// [R] wind attenuation coefficient. Call sites: code.R:281,285,747; NMR.R:576,577,787.
inline double attencoef(double hgt, double PAI, double cd = 0.2, double iw = 0.5, double phi_m = 1) {
  double l_m = mixinglength(hgt, PAI);
  double a = std::sqrt((cd * PAI * hgt) / (2 * l_m * iw));
  return a * std::sqrt(phi_m);
}
// [R]+[V vig.md:101] Monin-Obukhov diabatic corrections. psi_m limited to [-1.5, 4] and psi_h to
// [-2.5, 4] (the vignette's stated range; -2.5 = -1.5/0.6). Call sites: code.R:730; NMR.R:781.
inline void diabatic_cor(double tc, double pk, double H, double uf, double zi, double d,
                         double& psi_m, double& psi_h) {
  double Tk = tc + 273.15;
  double ph = phair(tc, pk);
  double cp = cpair(tc);
  double st = -(0.4 * 9.81 * (zi - d) * H) / (ph * cp * Tk * (uf * uf * uf));
  if (st < 0) {                       // unstable
    psi_h = -2 * std::log((1 + std::sqrt(1 - 16 * st)) / 2);
    psi_m = 0.6 * psi_h;
  } else {                            // stable / neutral
    psi_h = 6 * std::log(1 + st);
    psi_m = psi_h;
  }
  if (psi_m > 4) psi_m = 4;
  if (psi_h > 4) psi_h = 4;
  if (psi_m < -1.5) psi_m = -1.5;
  if (psi_h < -2.5) psi_h = -2.5;
}
// [D] within-canopy stability from a bulk Richardson number of the previous profile; scalar result
// (the reference assigns the result to scalar slots). Neutral = 1.
// Call site: code.R:731  diabatic_cor_can(tc, previn$uz, z, vegp$PAI).
inline void diabatic_cor_can(const double* tc, const double* uz, const double* z, int m,
                             double& phi_m, double& phi_h) {
  double dz = z[m - 1] - z[0];
  double dtdz = (tc[m - 1] - tc[0]) / dz;
  double dudz = (uz[m - 1] - uz[0]) / dz;
  if (std::fabs(dudz) < 0.01) dudz = 0.01;
  double tm = 0; for (int i = 0; i < m; ++i) tm += tc[i];
  double Tk = tm / m + 273.15;
  double Ri = (9.81 / Tk) * dtdz / (dudz * dudz);
  if (Ri > 0.15) Ri = 0.15;
  if (Ri < -1) Ri = -1;
  if (Ri < 0) {
    phi_h = 1 / std::sqrt(1 - 16 * Ri);
    phi_m = std::sqrt(phi_h);
  } else {
    double st = Ri / (1 - 5 * Ri);
    phi_m = 1 + (6 * st) / (1 + st);
    phi_h = phi_m;
  }
}
// [R]+[D] above-canopy turbulent molar conductance between z0 and z1 (> z0).
// Call sites: NMR.R:602,624,802,824. Denominator floored at a quarter of its neutral value so that the
// conductance stays positive and strictly decreasing in z1 (needed by NMR.R:625,825).
inline double gturb(double u, double zu, double z1, double z0, double hgt, double PAI, double tc,
                    double psi_m = 0, double psi_h = 0, double zm0 = 0.004, double pk = 101.3) {
  double d = zeroplanedis(hgt, PAI);
  double zm = roughlength(hgt, PAI, zm0);
  double zh = 0.2 * zm;
  double ph = phair(tc, pk);
  double ln = std::log((zu - d) / zm) + psi_m;
  if (!(ln >= 0.55)) ln = 0.55;
  double uf = (0.4 * u) / ln;
  double lo = z0 - d;
  if (!(lo > zh)) lo = zh;
  double lnh = std::log((z1 - d) / lo);
  double den = lnh + psi_h;
  if (den < 0.25 * lnh) den = 0.25 * lnh;
  return (0.4 * ph * uf) / den;
}
// [R]+[V vig.md:575-581] within-canopy molar conductance between z0 < z1, floored by free convection.
// Call sites: code.R:752,765,769,885; NMR.R:606,638,640,806,836,838.
inline double gcanopy(double uh, double z1, double z0, double tc1, double tc0, double hgt, double PAI,
                      double cd = 0.2, double iw = 0.5, double phi_m = 1, double pk = 101.3) {
  double l_m = mixinglength(hgt, PAI);
  double a = attencoef(hgt, PAI, cd, iw, phi_m);
  double ph = phair((tc1 + tc0) / 2, pk);
  double e0 = std::exp(-a * (z0 / hgt - 1));
  double e1 = std::exp(-a * (z1 / hgt - 1));
  double g = (l_m * iw * ph * uh * a) / (hgt * (e0 - e1) * phi_m);
  double gfree = 0.0012 * ph * std::pow(std::fabs(tc1 - tc0) / std::fabs(z1 - z0), 0.25);
  return (g > gfree) ? g : gfree;   // pmax; a NaN forced term falls through to gfree
}
// [R]+[V vig.md:491-517] leaf boundary-layer conductance, forced vs free (pmax).
// Call sites: code.R:784 (5 args); NMR.R:607,641,807,839 (6 args, dtmin = 5).
inline double gforcedfree(double d, double u, double tc, double dtc, double pk = 101.3, double dtmin = 1) {
  double Tk = tc + 273.15;
  double ph = phair(tc, pk);
  double sc = std::pow(Tk / 273.15, 1.75) * (101.3 / pk);
  double Dh = 18.9e-6 * sc;        // thermal diffusivity
  double nu = 13.3e-6 * sc;        // kinematic viscosity
  double adt = std::fabs(dtc);
  if (adt < dtmin) adt = dtmin;
  double Re = (u * d) / nu;
  double Pr = nu / Dh;
  double Gr = (9.81 * (d * d * d) * adt) / (Tk * (nu * nu));
  double gforced = (0.664 * Dh * ph * std::sqrt(Re) * std::cbrt(Pr)) / d;
  double gfree = (0.54 * Dh * ph * std::pow(Gr * Pr, 0.25)) / d;
  return (gforced > gfree) ? gforced : gfree;
}
// [V vig.md:664] stomatal hyperbola; [R] Qa = 4.6 * Rsw. Call sites: code.R:781; NMR.R:608,643,808,841.
inline double layercond(double Rsw, double gsmax, double q50 = 100) {
  double Qa = 4.6 * Rsw;
  return (gsmax * Qa) / (Qa + q50);
}

// ---- soil -----------------------------------------------------------------------------------
// [R] C&N eq 4.13-style soil surface relative humidity. Call sites: code.R:159,160,416,837; NMR.R:442.
inline double soilrh(double theta, double b, double Psie, double Smax, double tc) {
  double matric = -Psie * std::pow(theta / Smax, -b);
  return std::exp((0.018 * matric) / (8.31 * (tc + 273.15)));
}
// [V vig.md:852-854]+[R Campbell 1985] soil conductances k (W m^-2 K^-1; node i <-> i+1, last <-> deep
// boundary) and heat storage cd (W m^-2 K^-1). Call site: code.R:790.  theta is a scalar (thz null) or, when runmodel was given
// soil moisture by depth (code.R:1146-1154: theta[, i] is then a vector over the soil nodes), the vector thz[0..sm): [D] node i takes
// its conductivity and heat capacity from its own moisture.
inline void soilk(double timestep, double theta, double Vq, double Vm, double Vo, double Mc, double rho,
                  const double* z, int sm, double* k, double* cd, const double* thz = nullptr) {
  double A = 0.65 - 0.78 * rho + 0.6 * rho * rho;
  double B = 1.06 * rho;
  double C = 1 + 2.6 / std::sqrt(Mc);
  double D = 0.03 + 0.1 * rho * rho;
  double ct = C * theta;
  double lam = A + B * theta - (A - D) * std::exp(-((ct * ct) * (ct * ct)));
  double Cv = (2.128 * Vq + 2.385 * Vm + 2.496 * Vo + 4.18 * theta) * 1e6;
  for (int i = 0; i < sm; ++i) {
    if (thz) {
      ct = C * thz[i];
      lam = A + B * thz[i] - (A - D) * std::exp(-((ct * ct) * (ct * ct)));
      Cv = (2.128 * Vq + 2.385 * Vm + 2.496 * Vo + 4.18 * thz[i]) * 1e6;
    }
    double dzk = (i < sm - 1) ? (z[i + 1] - z[i]) : (z[sm - 1] - z[sm - 2]);
    k[i] = lam / dzk;
    double zlo = (i > 0) ? z[i - 1] : 0.0;
    double zhi = (i < sm - 1) ? z[i + 1] : (z[sm - 1] + (z[sm - 1] - z[sm - 2]));
    cd[i] = (Cv * (zhi - zlo)) / (2 * timestep);
  }
}

// ---- time / sun -----------------------------------------------------------------------------
// [R] astronomical Julian day number at 12:00 of a calendar date (microclima::julday form).
inline double jday(int year, int month, int day) {
  int a = (14 - month) / 12;
  int y = year + 4800 - a;
  int mm = month + 12 * a - 3;
  long jdn = day + (153 * mm + 2) / 5 + 365L * y + y / 4 - y / 100 + y / 400 - 32045;
  return (double)jdn;
}
// [R] microclima::solartime + solalt (degrees). Call sites: code.R:38; NMR.R:219,365.
inline double solalt(double localtime, double lat, double lon, double jd, double merid = 0, double dst = 0) {
  double mm = 6.24004077 + 0.01720197 * (jd - 2451545);
  double eot = -7.659 * std::sin(mm) + 9.863 * std::sin(2 * mm + 3.5932);
  double st = localtime + (4 * (lon - merid) + eot) / 60 - dst;
  double tt = 0.261799 * (st - 12);
  double declin = (PI * 23.5 / 180) * std::cos(2 * PI * ((jd - 159.5) / 365.25));
  double latr = lat * PI / 180;
  double sh = std::sin(declin) * std::sin(latr) + std::cos(declin) * std::cos(latr) * std::cos(tt);
  return (180 * std::atan(sh / std::sqrt(1 - sh * sh))) / PI;
}
// [D] diffuse fraction from a clearness index (Erbs-type); only reachable when dp is NA
// (code.R:37), never on the runmodel path (code.R:1166-1167).
inline double difprop(double Rsw, double jd, double lt, double lat, double lon, double merid = 0, double dst = 0) {
  double sa = solalt(lt, lat, lon, jd, merid, dst);
  if (!(sa > 0) || !(Rsw > 0)) return 1.0;
  double kt = Rsw / (1352.0 * std::sin(sa * PI / 180));
  if (kt > 1) kt = 1;
  double kd;
  if (kt <= 0.22) kd = 1 - 0.09 * kt;
  else if (kt <= 0.8) kd = 0.9511 - 0.1604 * kt + 4.388 * kt * kt - 16.638 * kt * kt * kt + 12.336 * kt * kt * kt * kt;
  else kd = 0.165;
  return kd;
}

// ---- canopy radiation -----------------------------------------------------------------------
// [R] ellipsoidal extinction coefficient for solar altitude sa (deg): Campbell's (1986) PUBLISHED form
// K = sqrt(x^2 + tan^2(zenith)) / (x + 1.774 (x + 1.182)^-0.733) — exponent -0.733, the square root over the numerator only.  The
// vignette prints a different-looking formula at vig.md:415 (exponent 0.773 under one root); that typesetting is taken to be a slip and
// is NOT what is implemented here: the tag is [R] (recollection of the published equation), not [V].  mt_cansw / mt_psunlit traces of
// tools/r_parity/dump_reference.R are what would settle it.
inline double cank(double x, double sa) {
  double zen = (90 - sa) * PI / 180;
  double t = std::tan(zen);
  return std::sqrt(x * x + t * t) / (x + 1.774 * std::pow(x + 1.182, -0.733));
}
inline double clumptr(double tr, double clump) { return clump + (1 - clump) * tr; }
// [R] sunlit fraction under plant area l. Call sites: code.R:41; NMR.R:368.
inline double psunlit(double l, double x, double sa, double clump = 0) {
  if (!(sa > 0)) return clumptr(0.0, clump);
  return clumptr(std::exp(-cank(x, sa) * l), clump);
}
// [R] beam multiplier for sunlit leaves relative to the horizontal flux. Call sites: code.R:42; NMR.R:369.
inline double radmult(double x, double sa) { return cank(x, sa); }
// [V vig.md:392-393] diffuse transmission, K = 1 with reflectance correction. Call sites: NMR.R:658,856.
inline double cantransdif(double l, double ref, double clump = 0) {
  return clumptr(std::exp(-std::sqrt(1 - ref) * l), clump);
}
// [V]+[R] shortwave flux density under plant area l given a precomputed solar altitude.
// Call sites: code.R:43-45; NMR.R:370-372,642,656,840,854. The reference re-derives the solar
// altitude from (jd, lt, lat, long) inside every call; the value is identical, so it is passed in.
inline double cansw_sa(double globrad, double dp, double sa, double l, double x, double ref, double clump = 0) {
  double s = std::sqrt(1 - ref);
  double trd = clumptr(std::exp(-s * l), clump);
  double trb = (sa > 0) ? clumptr(std::exp(-s * cank(x, sa) * l), clump) : clumptr(0.0, clump);
  return globrad * (dp * trd + (1 - dp) * trb);
}
// [V vig.md:440]+[D] longwave under plant area l: absorbed by a leaf and incoming (downward).
// `ref` is the LW reflectance (1 - emissivity). Call sites: code.R:47,897; NMR.R:374,664,862.
inline void canlw(double tc, double l, double ref, double skyem, double clump, double& lwabs, double& lwin) {
  double sb = 5.67e-8;
  double tk = tc + 273.15;
  double lwout = sb * ((tk * tk) * (tk * tk));
  double tr = clumptr(std::exp(-std::sqrt(1 - ref) * l), clump);
  lwin = tr * skyem * lwout + (1 - tr) * (1 - ref) * lwout;
  lwabs = (1 - ref) * lwin;
}

}  // namespace orc
