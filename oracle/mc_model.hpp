// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// Literal CPU restatement (C++17, scalar, R evaluation order, 1-based vectors) of the reference's
// transient hot path:  /root/reference/R/code.R  (Thomas 94-125, ThomasV 153-169, windprofile 191-221,
// windcanopy 279-294, abovecanopytemp 319-331, leafabs 31-49, leaftemp 378-512, paraminit 566-585,
// soilinit 619-627, runonestep 687-902, .vegpsort/.snowsort 904-933, spinup 982-1013, runmodel 1125-1328).
// Every quirk listed in SURVEY.md Appendix B (Q1..Q23) is reproduced on purpose.
// The microctools layer-algebra functions (layermerge / layeraverage / layerinterp) have NO available
// source; they are defined here ([D], see DESIGN.md "compat layer").
// PARITY UNPINNED vs real R (see mc_compat.hpp header).  Pinned: SURVEY Appendix F known answers.
#pragma once
#include "mc_compat.hpp"
#include <cstring>
#include <limits>

namespace orc {

constexpr int MAXM = 64;    // max canopy nodes
constexpr int MAXS = 32;    // max soil nodes
constexpr int MAXN = MAXM + MAXS + 4;

// 1-based fixed-capacity vector, so the code below reads like the R it restates.
struct Vec {
  double a[MAXN + 2];
  int n = 0;
  Vec() {}
  explicit Vec(int n_, double v = 0.0) : n(n_) { for (int i = 0; i <= n_ + 1; ++i) a[i] = v; }
  double& operator()(int i) { return a[i]; }
  double operator()(int i) const { return a[i]; }
};
inline double vsum(const Vec& v, int lo, int hi) { double s = 0; for (int i = lo; i <= hi; ++i) s += v(i); return s; }
inline double vsum(const Vec& v) { return vsum(v, 1, v.n); }
inline double vmean(const Vec& v) { return vsum(v) / v.n; }
inline bool is_na(double x) { return std::isnan(x); }
constexpr double NA = std::numeric_limits<double>::quiet_NaN();

struct Vegp {   // SURVEY Appendix C (21 fields, vig.md:367-370)
  int m = 0;
  double hgt, x, lw, cd, hgtg, zm0, refls, refg, refw, reflp, vegem, gsmax, q50, cpw, phw, kwood, clump;
  Vec PAI, pLAI, iw, thickw;
};
struct Soilp {  // code.R:619-627
  int sm = 0;
  double Smax, Smin, alpha, n, Ksat, Vq, Vm, Vo, Mc, rho, b, psi_e;
  Vec z;
};
struct State {  // previn / runonestep result, 23 fields (code.R:581-584, 898-901)
  int m = 0, sm = 0;
  Vec tc, soiltc, tleaf, uz, z, sz, rh, Rabs, gt, gv, gha, L, Rswin, Rlwin;
  double tabove, zabove, relhum, tair, tsoil, pk, H, G, psi_m;
};
struct Climvars { double tair, relhum, pk, u, tsoil, skyem, Rsw, dp; };  // code.R:1194-1196
struct Tme { double jd, lt; };   // jday(tme) and tme$hour+tme$min/60+tme$sec/3600 (code.R:35-36)

// ------------------------------------------------------------------------------------------------
// Thomas  (code.R:94-125).  tc: N+2 values, k: N+1, cd: N (or recycled scalar), X: N (or scalar).
inline Vec Thomas(Vec tc, double tmsoil, double tair, const Vec& k, const Vec& cd, double f, const Vec& X) {
  int m = tc.n - 2;
  Vec tn(m + 2, 0.0);
  tn(m + 2) = tmsoil;
  tn(1) = tair;
  double g = 1 - f;
  Vec a(m + 3, 0.0), b(m + 3, 0.0), cc(m + 3, 0.0), d(m + 3, 0.0);
  auto cdv = [&](int xx) { return cd.n == 1 ? cd(1) : cd(xx - 1); };
  auto Xv = [&](int xx) { return X.n == 1 ? X(1) : X(xx - 1); };
  for (int xx = 2; xx <= m + 1; ++xx) tc(xx) = tc(xx) + g * Xv(xx);
  for (int xx = 2; xx <= m + 1; ++xx) cc(xx) = -k(xx) * f;
  for (int xx = 2; xx <= m + 1; ++xx) a(xx + 1) = cc(xx);
  for (int xx = 2; xx <= m + 1; ++xx) b(xx) = f * (k(xx) + k(xx - 1)) + cdv(xx);
  for (int xx = 2; xx <= m + 1; ++xx)
    d(xx) = g * k(xx - 1) * tc(xx - 1) + (cdv(xx) - g * (k(xx) + k(xx - 1))) * tc(xx) + g * k(xx) * tc(xx + 1);
  d(2) = d(2) + k(1) * tn(1) * f;
  d(m + 1) = d(m + 1) + k(m + 1) * f * tn(m + 2);
  if (m >= 2) {   // Q20: `2:m` counts down when m == 1; callers here always have m >= 2
    for (int i = 2; i <= m; ++i) {
      cc(i) = cc(i) / b(i);
      d(i) = d(i) / b(i);
      b(i + 1) = b(i + 1) - a(i + 1) * cc(i);
      d(i + 1) = d(i + 1) - a(i + 1) * d(i);
    }
  }
  tn(m + 1) = d(m + 1) / b(m + 1);
  for (int i = m; i >= 2; --i) tn(i) = d(i) - cc(i) * tn(i + 1);
  for (int xx = 2; xx <= m + 1; ++xx) {            // Q11: clamp to the OLD 3-point stencil
    double xmn = std::min(std::min(tc(xx), tc(xx - 1)), tc(xx + 1));
    double xmx = std::max(std::max(tc(xx), tc(xx - 1)), tc(xx + 1));
    if (tn(xx) < xmn) tn(xx) = xmn;
    if (tn(xx) > xmx) tn(xx) = xmx;
  }
  for (int xx = 2; xx <= m + 1; ++xx) tn(xx) = tn(xx) + f * Xv(xx);
  return tn;
}
inline Vec rev(const Vec& v) { Vec r(v.n); for (int i = 1; i <= v.n; ++i) r(i) = v(v.n + 1 - i); return r; }

// ThomasV  (code.R:153-169)
inline Vec ThomasV(const Vec& Vo_in, const Vec& tn, double pk, double theta, double thetap, double relhum,
                   double tair, double tsoil, const Vec& zth, const Vec& gt, const Vec& Vflux, double f,
                   const State& previn, const Soilp& soilp) {
  int m = zth.n;
  Vec ph(m);
  for (int i = 1; i <= m; ++i) ph(i) = phair(tn(i), pk);
  double ea = 0.6108 * std::exp(17.27 * tair / (tair + 237.3)) * (relhum / 100);
  double eap = 0.6108 * std::exp(17.27 * previn.tair / (previn.tair + 237.3)) * (previn.relhum / 100);
  double Vair = ea / pk;
  double rhsoil = soilrh(theta, soilp.b, soilp.psi_e, soilp.Smax, tsoil);
  double rhsoilp = soilrh(thetap, soilp.b, soilp.psi_e, soilp.Smax, previn.soiltc(1));
  if (rhsoil > 1) rhsoil = 1;
  if (rhsoilp > 1) rhsoilp = 1;
  double eas = 0.6108 * std::exp(17.27 * tsoil / (tsoil + 237.3)) * rhsoil;
  double easp = 0.6108 * std::exp(17.27 * previn.tsoil / (previn.tsoil + 237.3)) * rhsoilp;   // Q5
  double Vsoil = eas / pk;
  Vec Vo(m + 2);
  Vo(1) = easp / previn.pk;
  for (int i = 1; i <= m; ++i) Vo(i + 1) = Vo_in(i);
  Vo(m + 2) = eap / previn.pk;
  Vec Vn = Thomas(rev(Vo), Vsoil, Vair, rev(gt), rev(ph), f, Vflux);   // Q6: Vflux not reversed
  return rev(Vn);
}

// windprofile for ONE output height (code.R:191-221; the R is vectorised over zo)
inline double windprofile(double ui, double zi, double zo, double a, double PAI, double hgt, double psi_m = 0,
                          double hgtg = NA, double zm0 = 0.004) {
  if (is_na(hgtg)) hgtg = 0.05 * hgt;
  double d = zeroplanedis(hgt, PAI);
  double zm = roughlength(hgt, PAI, zm0);
  double zmg = 0.1 * hgtg;
  double ln1 = std::log((zi - d) / zm) + psi_m;
  double uf = 0.4 * (ui / ln1);
  double ln2 = std::log((zo - d) / zm) + psi_m;          // NaN when zo <= d (suppressWarnings)
  if (ln2 < 0.55) ln2 = 0.55;                            // NaN stays NaN, as in R
  double uo = (uf / 0.4) * ln2;
  if (zo < hgt) {
    ln2 = std::log((hgt - d) / zm) + psi_m;
    uo = (uf / 0.4) * ln2;
  }
  if (zo > (0.1 * hgt) && zo < hgt) uo = uo * std::exp(a * ((zo / hgt) - 1));
  if (zo <= (0.1 * hgt)) {
    uo = uo * std::exp(a * (((0.1 * hgt) / hgt) - 1));
    double ln1b = std::log((0.1 * hgt) / zmg) + psi_m;
    double ufb = 0.4 * (uo / ln1b);
    double ln2b = std::log(zo / zmg) + psi_m;
    uo = (ufb / 0.4) * ln2b;
  }
  return uo;
}

// windcanopy  (code.R:279-294); edgedist NaN == NA
inline double windcanopy(double uh, double z, double hgt, double PAI, double cd, double iw, double phi_m,
                         double edgedist, double uref, double zref) {
  double a = attencoef(hgt, PAI, cd, iw, phi_m);
  double uz = uh * std::exp(a * ((z / hgt) - 1));
  if (!is_na(edgedist)) {
    double a2 = attencoef(hgt, PAI, cd, iw, phi_m);
    double ur = windprofile(uref, zref, z, 0.7989537, 1.146, 0.12);
    double uwr = 2.5 * std::log((z - 0.08) / 0.01476);
    if (uwr < 1) uwr = 1;
    if (is_na(uwr)) uwr = 1;
    double edr = (hgt - edgedist / uwr) / hgt;
    double uhr = ur * std::exp(a2 * (edr - 1));
    // pmax(uz, uhr, na.rm = T)
    if (is_na(uz)) uz = uhr;
    else if (!is_na(uhr) && uhr > uz) uz = uhr;
  }
  return uz;
}

// abovecanopytemp  (code.R:319-331)
inline double abovecanopytemp(double tz, double uz, double zu, double zo, double H, double hgt, double PAI,
                              double zm0, double pk, double psi_h) {
  double d = zeroplanedis(hgt, PAI);
  double zm = roughlength(hgt, PAI, zm0);
  double zh = 0.2 * zm;
  double uf = (0.4 * uz) / (std::log((zu - d) / zm) + psi_h);
  double ph = phair(tz, pk);
  double cp = cpair(tz);
  double m = H / (0.4 * ph * cp * uf);
  double tref = tz + m * (std::log((zu - d) / zh) + psi_h);
  double tc = tref - m * (std::log((zo - d) / zh) + psi_h);
  return tc;
}

// leafabs  (code.R:31-49)
struct LeafAbs { Vec aRsw, aRlw, ref; };
inline LeafAbs leafabs(double Rsw, Tme tme, double tair, double /*tground*/, double lat, double lon, const Vec& PAI,
                       const Vec& pLAI, double x, double refls, double refw, double refg, double vegem,
                       double skyem, double dp, double merid, double dst, double clump) {
  int m = PAI.n;
  Vec cs(m);                      // cumsum(PAI)
  { double s = 0; for (int i = 1; i <= m; ++i) { s += PAI(i); cs(i) = s; } }
  Vec PAIc = rev(cs);             // rev(cumsum(PAI))  (sic)
  Vec PAIg = cs;
  double jd = tme.jd, lt = tme.lt;
  if (is_na(dp)) dp = difprop(Rsw, jd, lt, lat, lon, merid, dst);
  double sa = solalt(lt, lat, lon, jd, merid, dst);
  double sa2 = sa < 5 ? 5 : sa;    // Q23
  LeafAbs out; out.aRsw = Vec(m); out.aRlw = Vec(m); out.ref = Vec(m);
  double mul = radmult(x, sa2);
  double sumPAI = vsum(PAI);
  for (int i = 1; i <= m; ++i) out.ref(i) = pLAI(i) * refls + (1 - pLAI(i)) * refw;
  double mref = vmean(out.ref);
  double aG0 = refg * cansw_sa(Rsw, dp, sa, sumPAI, x, mref, clump);
  for (int i = 1; i <= m; ++i) {
    double ref = out.ref(i);
    double sunl = psunlit(PAIc(i), x, sa, clump);
    double aRsw = (1 - ref) * cansw_sa(Rsw, dp, sa, PAIc(i), x, ref, clump);
    double aGround = (1 - ref) * cansw_sa(aG0, dp, sa, PAIg(i), x, ref, clump);
    out.aRsw(i) = sunl * mul * aRsw + (1 - sunl) * aRsw + aGround;
    double lwabs, lwin;
    canlw(tair, PAIc(i), 1 - vegem, skyem, clump, lwabs, lwin);
    out.aRlw(i) = lwabs;
  }
  return out;
}

// leaftemp  (code.R:378-512)
// esoilv: with soil moisture by depth `theta` is a vector over the soil nodes, so esoil is one too and R recycles it over the canopy
// layers in every element-wise expression it meets (code.R:416-417, 424-437, 865): layer i sees esoilv[(i-1) %% length] (quirk Q24);
// `esoil` is its first element (tln$esoil[1], code.R:799-805).
struct LeafTemp { Vec tn, tleaf, ea, gtt, Rem, H, L; double esoil; Vec esoilv; };
inline double edf(double ea1, double ea2, double tc) {   // code.R:380-386
  double es = 0.6108 * std::exp(17.27 * tc / (tc + 237.3));
  double y = ea1 > ea2 ? ea1 - ea2 : 0;
  y = ea2 > es ? 0 : y;
  return y;
}
inline LeafTemp leaftemp(double tair, double relhum, double pk, double timestep, Vec gt, Vec gha, Vec gv, Vec Rabs,
                         const State& previn, const Vegp& vegp, const Soilp& soilp, double theta,
                         double zlafact = 1, double surfwet = 1, const double* thz = nullptr, int nthz = 0) {
  const Vec& z = previn.z;
  double lambda = (-42.575 * tair + 44994) * surfwet;
  int m = gt.n - 1;
  Vec zth(m), zla(m);
  for (int i = 1; i <= m - 1; ++i) zth(i) = z(i + 1) - z(i);
  zth(m) = vegp.hgt - (z(m) + z(m - 1)) / 2;
  for (int i = 1; i <= m; ++i) zla(i) = mixinglength(zth(i), vegp.PAI(i)) * 0.5 * zlafact;
  for (int i = 1; i <= m + 1; ++i) gt(i) = 0.5 * gt(i) + 0.5 * previn.gt(i);
  Vec gv2(m);
  for (int i = 1; i <= m; ++i) {
    double gv1 = 1 / ((1 / previn.gv(i)) + (1 / previn.gha(i)));
    double gv2i = 1 / ((1 / gv(i)) + (1 / gha(i)));
    gv(i) = 0.5 * gv1 + 0.5 * gv2i;
  }
  for (int i = 1; i <= m; ++i) gha(i) = 0.5 * gha(i) + 0.5 * previn.gha(i);
  double mtref = 0.5 * tair + 0.5 * previn.tair;
  double mrh = 0.5 * relhum + 0.5 * previn.relhum;
  Vec igtt(m + 1, 1 / gt(m + 1)), igtt2(m);
  igtt2(1) = 1 / gt(1);
  for (int i = m; i >= 1; --i) igtt(i) = igtt(i + 1) + 1 / gt(i);
  for (int i = 2; i <= m; ++i) igtt2(i) = igtt2(i - 1) + 1 / gt(i);
  Vec gtt(m), gtt2(m), zref(m);
  for (int i = 1; i <= m; ++i) { gtt(i) = 1 / igtt(i + 1); gtt2(i) = 1 / igtt2(i); zref(i) = (vegp.hgt + 2) - z(i); }
  double mpk = 0.5 * previn.pk + 0.5 * pk;
  double esref = satvap(mtref);
  double eref = (mrh / 100) * esref;
  double rhsoil = soilrh(theta, soilp.b, -soilp.psi_e, soilp.Smax, previn.soiltc(1));   // Q7
  double esoil = rhsoil * satvap(previn.soiltc(1));
  Vec esoilv(thz ? nthz : 1, esoil);
  if (thz) {
    for (int j = 1; j <= nthz; ++j) esoilv(j) = soilrh(thz[j - 1], soilp.b, -soilp.psi_e, soilp.Smax, previn.soiltc(1)) * satvap(previn.soiltc(1));
    esoil = esoilv(1);
  }
  const double esoil1 = esoil;
  LeafTemp o; o.esoilv = esoilv; o.tn = Vec(m); o.tleaf = Vec(m); o.ea = Vec(m); o.gtt = gtt; o.Rem = Vec(m); o.H = Vec(m); o.L = Vec(m);
  o.esoil = esoil1;
  double sb = 5.67 * 1e-8;
  for (int i = 1; i <= m; ++i) {
    esoil = esoilv((i - 1) % esoilv.n + 1);                                   // R recycling (Q24); the scalar case has length 1
    double ptc = previn.tc(i), ptl = previn.tleaf(i);
    double esj = 0.6108 * std::exp(17.27 * ptc / (ptc + 237.3));
    double eaj = (previn.rh(i) / 100) * esj;
    double estl = satvap(ptl);
    double delta = 4098 * (0.6108 * std::exp(17.27 * ptl / (ptl + 237.3))) / ((ptl + 237.3) * (ptl + 237.3));
    // steady-state test for vapour
    double PAIm = 2 * vegp.PAI(i) / zth(i);
    double gv2b = (PAIm / zth(i)) * gv(i);                                  // Q21
    double test = std::max(std::max(timestep * gtt(i) / zref(i), timestep * gv(i) / zla(i)), timestep * gtt2(i) / z(i));
    double btm = (1 / timestep) + 0.5 * (gtt(i) / zref(i) + gtt2(i) / z(i) + gv(i) / zla(i));
    double ae = eaj + 0.5 * ((gtt(i) / zref(i)) * edf(eref, eaj, ptc) + (gtt2(i) / z(i)) * edf(esoil, eaj, ptc) +
                             (gv(i) / zla(i)) * edf(estl, eaj, ptc)) / btm;
    double be = (0.25 * gv(i) * delta) / btm;
    if (test > 1) {
      ae = (gtt(i) * eref + gtt2(i) * esoil + gv2b * estl) / (gtt(i) + gtt2(i) + gv2b);
      be = (0.5 * delta) / (gtt(i) + gtt2(i) + gv2b);
    }
    // air temperature
    test = std::max(std::max(timestep * gtt(i) / zref(i), timestep * gha(i) / zla(i)), timestep * gtt2(i) / z(i));
    double ph = phair(ptc, previn.pk);
    double cp = cpair(ptc);
    double vden = vegp.thickw(i) * vegp.PAI(i);
    double ma = (timestep * PAIm * 2) / (cp * ph * (1 - vden));
    double K1 = gtt(i) * cp / zref(i), K2 = gtt2(i) * cp / z(i), K3 = gha(i) * cp * PAIm / zla(i);
    btm = 1 + 0.5 * ma * (K1 + K2 + K3);
    double aL = (ptc + 0.5 * ma * (K1 * mtref + K2 * previn.soiltc(1) + K3 * ptl)) / btm;
    double bL = (0.25 * ma * K3) / btm;
    if (test > 1) {
      K1 = gtt(i) * cp; K2 = gtt2(i) * cp; K3 = gha(i) * cp;
      aL = (K1 * mtref + K2 * previn.soiltc(1) + K3 * ptl) / (K1 + K2 + K3);
      bL = (0.5 * K3) / (K1 + K2 + K3);
    }
    if (bL < 0) bL = 0;
    double aH = cp * gha(i) * (ptl - aL);
    double bH = cp * gha(i) * (0.5 - 0.5 * bL);
    if (bH < 0) bH = 0;
    double aX = (lambda * gv(i) / mpk) * edf(estl, ae, ptc);
    double bX = (lambda * gv(i) / mpk) * (delta - be);
    if (bX < 0) bX = 0;
    double Ra = 0.5 * Rabs(i) + 0.5 * previn.Rabs(i);
    double tk = ptl + 273.15;
    double aR = vegp.vegem * sb * ((tk * tk) * (tk * tk));
    double bR = vegp.vegem * sb * 2 * (tk * tk * tk);
    double Ch = vden * vegp.cpw * vegp.phw;
    double ml = (timestep * PAIm) / Ch;
    double dTL = (ml * (Ra - aR - aX - aH)) / (zla(i) + ml * (bR + bX + bH));
    double dTL2 = (Ra - aR - aX - aH) / (bR + bX + bH);
    if (std::fabs(dTL2) < std::fabs(dTL)) dTL = dTL2;
    double tn = aL + bL * dTL;
    double tleaf = ptl + dTL;
    if (tleaf > tn) { double tmx = std::max(tair + 15, tleaf); if (tn > tmx) tn = tmx; }
    if (tleaf < tn) { double tmn = std::min(tair - 5, tleaf); if (tn < tmn) tn = tmn; }
    double eam = ae + be * dTL;
    double ean = eaj + 2 * (eam - eaj);
    double es = 0.6108 * std::exp(17.27 * tn / (tn + 237.3));
    if (ean > es) ean = es;
    double eamn = (relhum / 100) * 0.6108 * std::exp(17.27 * tair / (tair + 237.3));
    if (ean < eamn) ean = eamn;
    double a = std::log(ean / 0.6108);
    double Tdew = -(237.3 * a) / (a - 17.27);
    double dT1 = -(ptl + dTL - Tdew);
    if (dT1 < 0) dT1 = 0;
    double estl2 = 0.6108 * std::exp(17.27 * tleaf / (tleaf + 237.3));
    double Lc = lambda * gv(i) * (estl2 - ean) / mpk;
    double dT2 = -(ml / zla(i)) * Lc;
    if (dT2 < 0) dT2 = 0;
    double dT = std::min(dT1, dT2);
    tleaf = tleaf + dT;
    dTL = tleaf - ptl;
    double dTmx = ptc + 20.2 - ptl;
    double dTmn = ptc - 4.5 - ptl;
    if (dTL > dTmx) dTL = dTmx;
    if (dTL < dTmn) dTL = dTmn;
    o.tn(i) = aL + bL * dTL;
    o.tleaf(i) = ptl + dTL;
    o.ea(i) = ean;
    o.Rem(i) = aR + bR * dTL;
    o.H(i) = aH + bH * dTL;
    o.L(i) = aX + bX * dTL;
  }
  return o;
}

// paraminit  (code.R:566-585).  stats::spline with two knots == linear interpolation at n points.
inline Vec lin2(double y1, double y2, int n) {
  Vec v(n);
  for (int i = 1; i <= n; ++i) {
    double xx = 1 + (double)(i - 1) / (double)(n - 1);      // seq(1, 2, length.out = n)
    v(i) = y1 + (y2 - y1) * (xx - 1);
  }
  return v;
}
inline State paraminit(int m, int sm, double hgt, double tair, double u, double relhum, double tsoil, double Rsw) {
  State s; s.m = m; s.sm = sm;
  Vec tcb = lin2(tsoil, tair, m + sm);
  s.tc = Vec(m); for (int i = 1; i <= m; ++i) s.tc(i) = tcb(sm + i);
  s.soiltc = Vec(sm); for (int i = 1; i <= sm; ++i) s.soiltc(i) = tcb(sm + 1 - i);
  s.rh = Vec(m, relhum);
  s.Rabs = lin2(0.2 * Rsw + 20, Rsw + 20, m);
  s.tleaf = Vec(m); for (int i = 1; i <= m; ++i) s.tleaf(i) = s.tc(i) + 0.01 * s.Rabs(i);
  s.gt = Vec(m + 1, 50.0 * ((double)m / hgt) * (2.0 / m));
  s.gt(1) = 2 * s.gt(1);
  s.gt(m + 1) = s.gt(m + 1) * (hgt / m) * 0.5;
  s.gv = lin2(0.25, 0.32, m);
  s.gha = lin2(0.13, 0.19, m);
  s.z = Vec(m); for (int i = 1; i <= m; ++i) s.z(i) = (i - 0.5) / m * hgt;
  s.sz = Vec(sm); for (int i = 1; i <= sm; ++i) s.sz(i) = 2 / std::pow((double)sm, 2.42) * std::pow((double)i, 2.42);
  s.uz = Vec(m); for (int i = 1; i <= m; ++i) s.uz(i) = ((double)i / m) * u;
  s.tabove = tair; s.zabove = hgt; s.relhum = relhum; s.tair = tair; s.tsoil = tsoil; s.pk = 101.3;
  s.H = 0; s.L = Vec(m, 0.0); s.G = 0; s.Rswin = s.Rabs;
  s.Rlwin = Vec(m); for (int i = 1; i <= m; ++i) s.Rlwin(i) = 0.2 * s.Rabs(i);
  s.psi_m = 0;
  return s;
}
// soilinit node depths (code.R:622-623)
inline Vec soilinit_z(int m, double sdepth, double reqdepth) {
  Vec z(m);
  for (int i = 1; i <= m; ++i) z(i) = (sdepth / std::pow((double)m, 1.2)) * std::pow((double)i, 1.2);
  if (!is_na(reqdepth)) {
    double mn = 1e300; int sel = 1;
    for (int i = 1; i <= m; ++i) { double dd = std::fabs(z(i) - reqdepth); if (dd < mn) { mn = dd; sel = i; } }
    z(sel) = reqdepth;
  }
  return z;
}

// ---- layer algebra (microctools; source absent -> [D] defined here) -------------
struct LayerMerge { int u[MAXM + 2]; int ngroups; Vec gtx, zth; };
// layermerge(z, gt, hgt, timestep, ph)  — call site code.R:772.
// Greedy bottom-up grouping: layers are added to a group until the group's series conductance times the
// time step no longer exceeds its molar heat storage ( timestep / sum(1/gt) <= sum(ph*zth) ); an
// unsatisfied trailing group is absorbed into the group below it.
inline LayerMerge layermerge(const Vec& z, const Vec& gt, double /*hgt*/, double timestep, const Vec& ph) {
  int m = z.n;
  LayerMerge o; o.zth = Vec(m); o.gtx = Vec(m);
  for (int i = 1; i <= m; ++i) o.zth(i) = z(i) - (i > 1 ? z(i - 1) : 0.0);
  int g = 1; double R = 0, C = 0; bool open = false;
  for (int i = 1; i <= m; ++i) {
    R += 1 / gt(i); C += ph(i) * o.zth(i); o.u[i] = g; open = true;
    if (timestep / R <= C) { o.gtx(g) = 1 / R; ++g; R = 0; C = 0; open = false; }
  }
  if (open) {
    if (g > 1) { int lastg = g - 1; for (int i = 1; i <= m; ++i) if (o.u[i] == g) o.u[i] = lastg; o.ngroups = lastg; }
    else o.ngroups = 1;
  } else o.ngroups = g - 1;
  for (int gg = 1; gg <= o.ngroups; ++gg) {
    double Rs = 0; for (int i = 1; i <= m; ++i) if (o.u[i] == gg) Rs += 1 / gt(i);
    o.gtx(gg) = 1 / Rs;
  }
  o.gtx.n = o.ngroups;
  return o;
}
struct LayerAvg { int m; Vec tc, gt, z, Vo, ea, X, vden, cp, ph, TT; };
// layeraverage(lmm, tc, hgt, gha, gt, zla, z, Vo, ea, X, vden, pk, PAI, TT) — call site code.R:814.
// Thickness-weighted means per group; the group's node is its upper-median member c_g; conductances
// between merged nodes are the series sums of the member conductances between consecutive c_g.
inline LayerAvg layeraverage(const LayerMerge& lmm, const Vec& tc, double hgt, const Vec& gt, const Vec& z,
                             const Vec& Vo, const Vec& ea, const Vec& X, const Vec& vden, double pk, const Vec& TT) {
  int m = z.n, mg = lmm.ngroups;
  LayerAvg o; o.m = mg;
  o.tc = Vec(mg); o.gt = Vec(mg + 1); o.z = Vec(mg + 2); o.Vo = Vec(mg); o.ea = Vec(mg); o.X = Vec(mg);
  o.vden = Vec(mg); o.cp = Vec(mg); o.ph = Vec(mg); o.TT = Vec(mg);
  int cprev = 0;
  o.z(1) = 0;
  for (int g = 1; g <= mg; ++g) {
    int lo = 0, hi = 0;
    for (int i = 1; i <= m; ++i) if (lmm.u[i] == g) { if (!lo) lo = i; hi = i; }
    double w = 0, stc = 0, sVo = 0, sea = 0, sX = 0, svd = 0;
    for (int i = lo; i <= hi; ++i) {
      double wi = lmm.zth(i);
      w += wi; stc += wi * tc(i); sVo += wi * Vo(i); sea += wi * ea(i); sX += wi * X(i); svd += wi * vden(i);
    }
    o.tc(g) = stc / w; o.Vo(g) = sVo / w; o.ea(g) = sea / w; o.X(g) = sX / w; o.vden(g) = svd / w;
    o.cp(g) = cpair(o.tc(g)); o.ph(g) = phair(o.tc(g), pk);
    int c = (lo + hi + 1) / 2;
    o.z(g + 1) = z(c); o.TT(g) = TT(c);
    if (c - cprev == 1) o.gt(g) = gt(c);
    else { double R = 0; for (int j = cprev + 1; j <= c; ++j) R += 1 / gt(j); o.gt(g) = 1 / R; }
    cprev = c;
  }
  if (m + 1 - cprev == 1) o.gt(mg + 1) = gt(m + 1);
  else { double R = 0; for (int j = cprev + 1; j <= m + 1; ++j) R += 1 / gt(j); o.gt(mg + 1) = 1 / R; }
  o.z(mg + 2) = hgt;
  return o;
}
// layerinterp(T2, TT, y) — call sites code.R:846-847. Linear interpolation (R `approx`, rule = 1).
inline Vec layerinterp(const Vec& T2, const Vec& TT, const Vec& y) {
  Vec o(TT.n);
  for (int q = 1; q <= TT.n; ++q) {
    double v = TT(q);
    if (!(v >= T2(1)) || !(v <= T2(T2.n))) { o(q) = NA; continue; }
    int i = 1, j = T2.n;
    while (i < j - 1) { int ij = (i + j) / 2; if (v < T2(ij)) j = ij; else i = ij; }   // R approx1 bisection
    if (v == T2(j)) o(q) = y(j);
    else if (v == T2(i)) o(q) = y(i);
    else o(q) = y(i) + (y(j) - y(i)) * ((v - T2(i)) / (T2(j) - T2(i)));
  }
  return o;
}

// ---- .vegpsort / .snowsort (code.R:904-933) are parameter selection; .snowsort restated here ----
inline Vegp snowsort(const Vegp& vegp, bool snow) {
  Vegp v = vegp;
  if (snow) { v.lw = v.lw * 1.5; v.zm0 = 0.024; v.refls = 0.85; v.refg = 0.8; v.refw = 0.7; v.reflp = 0.8; v.gsmax = 10; v.q50 = 1; }
  return v;
}

struct StepCfg {   // runonestep's scalar arguments (code.R:687-689)
  double timestep = 3600, lat = 50, lon = -5, edgedist = 1000, sdepth = 2, reqhgt = NA, zu = 2;
  double merid = 0, dst = 0, n = 0.6, windhgt = 2, zlafact = 1, surfwet = 1;
  int metopen = 1;
};

// runonestep  (code.R:687-902).  Returns 0, or -1 when m < 3 (stop at code.R:699).
inline int runonestep(const Climvars& cv, const State& previn, Vegp vegp, const Soilp& soilp, const StepCfg& cfg,
                      Tme tme, double theta, double thetap, State& out, int* nmerged = nullptr,
                      const double* thz = nullptr /* theta by soil node (then theta == thz[0]) */) {
  const double timestep = cfg.timestep, edgedist = cfg.edgedist, reqhgt = cfg.reqhgt, zu = cfg.zu;
  for (int i = 1; i <= vegp.PAI.n; ++i) if (vegp.PAI(i) == 0) vegp.PAI(i) = 0.0001;     // code.R:691-696
  int m = previn.m;
  if (m < 3) return -1;
  double tair = cv.tair, relhum = cv.relhum, pk = cv.pk, u = cv.u, tsoil = cv.tsoil, skyem = cv.skyem, Rsw = cv.Rsw, dp = cv.dp;
  double H = previn.H;
  Vec tc = previn.tc; double ppk = previn.pk, hgt = vegp.hgt;
  Vec ph(m), cp(m), lambda(m);
  for (int i = 1; i <= m; ++i) { ph(i) = phair(tc(i), ppk); cp(i) = cpair(tc(i)); lambda(i) = -42.575 * tc(i) + 44994; }
  double pha = phair(tair, pk);
  double sumPAI = vsum(vegp.PAI);
  double u2;
  if (cfg.metopen) {
    if (cfg.windhgt != 2) u = u * 4.87 / std::log(67.8 * cfg.windhgt - 5.42);
    u2 = u * std::log(67.8 * (hgt + 2) - 5.42) / std::log(67.8 * 2 - 5.42);
  } else {
    u2 = u;
    if (cfg.windhgt != (2 + hgt)) u2 = windprofile(u, cfg.windhgt, hgt + 2, 2, sumPAI, hgt, previn.psi_m);
  }
  if (u2 < 0.5) u2 = 0.5;
  Vec z(m);
  for (int i = 1; i <= m; ++i) z(i) = (i - 0.5) / m * hgt;
  if (!is_na(reqhgt)) {                                                               // Q18
    double mn = 1e300; int sel = 1;
    for (int i = 1; i <= m; ++i) { double dd = std::fabs(z(i) - reqhgt); if (dd < mn) { mn = dd; sel = i; } }
    z(sel) = reqhgt;
  }
  Vec zt(m - 1);
  for (int i = 1; i <= m - 1; ++i) zt(i) = z(i + 1) - z(i);
  double zabove = vegp.hgt;
  if (!is_na(reqhgt)) if (reqhgt > hgt) zabove = reqhgt;
  double d = zeroplanedis(hgt, sumPAI);
  double zm = roughlength(hgt, sumPAI, vegp.zm0);
  double uf = (0.4 * u2) / std::log(((hgt + 2) - d) / zm);
  double psi_m, psi_h, phi_m, phi_h;
  diabatic_cor(tair, pk, H, uf, (hgt + 2), d, psi_m, psi_h);
  diabatic_cor_can(&tc.a[1], &previn.uz.a[1], &z.a[1], m, phi_m, phi_h);
  (void)phi_h;
  double tcan = abovecanopytemp(tair, u2, zu + hgt, zabove, H, hgt, sumPAI, vegp.zm0, pk, psi_h);
  double es0 = satvap(tair);
  double ea0 = (relhum / 100) * es0;
  double tfrost = dewpoint(ea0, tair);
  if (tcan < tfrost) tcan = tfrost;
  if (tcan > (tair + 15)) tcan = tair + 15;
  if (tcan < (tair - 4.5)) tcan = tair - 4.5;
  {
    double ea = 0.6108 * std::exp(17.27 * tair / (tair + 237.3)) * (relhum / 100);
    double eas = 0.6108 * std::exp(17.27 * tcan / (tcan + 237.3));
    relhum = (ea / eas) * 100; if (relhum > 100) relhum = 100;
  }
  // wind speed and turbulent conductances (code.R:747-770)
  double acm = attencoef(hgt, sumPAI, vegp.cd, vegp.iw(m), phi_m);
  Vec uz(m, 0.0), gt(m + 1, 0.0);
  double uh = windprofile(u2, hgt + 2, hgt, acm, sumPAI, hgt, psi_m, vegp.hgtg, vegp.zm0);
  uz(m) = windcanopy(uh, z(m), hgt, sumPAI, vegp.cd, vegp.iw(m), phi_m, edgedist, u, zu);
  gt(m) = gcanopy(uh, z(m), z(m - 1), tc(m), tc(m - 1), hgt, sumPAI, vegp.cd, vegp.iw(m), phi_m, pk);
  Vec tcx(m + 1);                                  // tc <- c(previn$soiltc[1], tc)
  tcx(1) = previn.soiltc(1);
  for (int i = 1; i <= m; ++i) tcx(i + 1) = tc(i);
  for (int i = m - 1; i >= 1; --i) {
    double zi = z(i + 1) + 0.5 * zt(i), zo = z(i + 1) - 0.5 * zt(i);
    double PAIi = vsum(vegp.PAI, 1, i);
    uh = windcanopy(uh, zo, zi, PAIi, vegp.cd, vegp.iw(i), phi_m, edgedist, u, zu);
    double z0 = (i > 1) ? z(i - 1) : 0;
    uz(i) = windcanopy(uh, zi, zo, PAIi, vegp.cd, vegp.iw(i), phi_m, edgedist, u, zu);    // Q9
    gt(i) = gcanopy(uh, z(i), z0, tcx(i + 1), tcx(i), zi, PAIi, vegp.cd, vegp.iw(i), phi_m, pk);
  }
  gt(m + 1) = gcanopy(uh, hgt, z(m), tcan, tcx(2), hgt, vegp.PAI(m) * 0.5, vegp.cd, vegp.iw(m), phi_m, pk);   // Q8
  LayerMerge lmm = layermerge(z, gt, hgt, timestep, ph);
  // absorbed radiation
  LeafAbs Rabss = leafabs(Rsw, tme, tair, previn.soiltc(1), cfg.lat, cfg.lon, vegp.PAI, vegp.pLAI, vegp.x, vegp.refls,
                          vegp.refw, vegp.refg, vegp.vegem, skyem, dp, cfg.merid, cfg.dst, vegp.clump);
  Vec Rabs(m), gv(m), gha(m);
  for (int i = 1; i <= m; ++i) {
    Rabs(i) = Rabss.aRsw(i) + Rabss.aRlw(i);
    gv(i) = layercond(Rabss.aRsw(i) / (1 - Rabss.ref(i)), vegp.gsmax, vegp.q50);
    double dtc = previn.tleaf(i) - previn.tc(i);
    gha(i) = 1.41 * gforcedfree(vegp.lw * 0.71, uz(i), tc(i), dtc, pk);
  }
  LeafTemp tln = leaftemp(tcan, relhum, pk, timestep, gt, gha, gv, Rabs, previn, vegp, soilp, theta, cfg.zlafact, cfg.surfwet, thz, previn.sm);
  Vec Vo(m);
  for (int i = 1; i <= m; ++i) Vo(i) = (0.6108 * std::exp(17.27 * tc(i) / (tc(i) + 237.3)) * (previn.rh(i) / 100)) / previn.pk;
  int sm = previn.sm;
  Vec ksoil(sm), cdsoil(sm);
  soilk(timestep, theta, soilp.Vq, soilp.Vm, soilp.Vo, soilp.Mc, soilp.rho, &soilp.z.a[1], sm, &ksoil.a[1], &cdsoil.a[1], thz);
  Vec sz = soilp.z;
  Vec vden(m), X(m), TT(m);
  for (int i = 1; i <= m; ++i) { vden(i) = vegp.PAI(i) * vegp.thickw(i); X(i) = tln.tn(i) - tc(i); }
  { double s = 0; for (int i = 1; i <= m; ++i) { s += (ph(i) / gt(i)) * (z(i) - (i > 1 ? z(i - 1) : 0.0)); TT(i) = s; } }
  // soil heat (code.R:799-811)
  double Xs;
  double vdef = tln.esoil - tln.ea(1);
  double ps1 = previn.soiltc(1);
  if (vdef > 0) {
    double mm = (lambda(1) * ph(1) * z(1)) / pk;
    double delta = 4098 * (0.6108 * std::exp(17.27 * ps1 / (ps1 + 237.3))) / ((ps1 + 237.3) * (ps1 + 237.3));
    double btm = 1 + 0.5 * (mm / cdsoil(1)) * delta;
    Xs = ((1 - vegp.refg) * (Rabss.aRsw(1) / cdsoil(1)) - (mm / cdsoil(1)) * (tln.esoil - tln.ea(1))) / btm;
  } else {
    Xs = (1 - vegp.refg) * (Rabss.aRsw(1) / cdsoil(1));
  }
  double mmsgn = Xs < 0 ? -1 : 1;
  double Xsmx = ((1 - vegp.refg) * Rabss.aRsw(1)) / (gha(1) * cp(1)) + (tln.tn(1) - ps1);
  if (std::fabs(Xs) > std::fabs(Xsmx)) Xs = Xsmx;
  Xs = std::fabs(Xs) * mmsgn;
  Vec tn(m), ea(m), tnsoil(sm);
  if (nmerged) *nmerged = lmm.ngroups;
  if (lmm.ngroups > 1) {                                                               // code.R:813-852
    LayerAvg lav = layeraverage(lmm, tc, vegp.hgt, gt, z, Vo, tln.ea, X, vden, ppk, TT);
    int mg = lav.m;
    Vec cda(mg), ka(mg + 1);
    for (int g = 1; g <= mg; ++g) cda(g) = lav.cp(g) * lav.ph(g) * (1 - lav.vden(g)) * (lav.z(g + 2) - lav.z(g)) / 2 * timestep;   // Q10
    for (int g = 1; g <= mg; ++g) ka(g) = lav.gt(g) * lav.cp(g);
    ka(mg + 1) = lav.gt(mg + 1) * cpair(previn.tabove);
    for (int g = 1; g <= mg; ++g) if (ka(g) > cda(g)) ka(g) = cda(g);
    int N = mg + sm;
    Vec k(N + 1), cd(N), XX(N), tc2(N + 2);
    for (int g = 1; g <= mg + 1; ++g) k(g) = ka(mg + 2 - g);
    for (int j = 1; j <= sm; ++j) k(mg + 1 + j) = ksoil(j);
    for (int g = 1; g <= mg; ++g) cd(g) = cda(mg + 1 - g);
    for (int j = 1; j <= sm; ++j) cd(mg + j) = cdsoil(j);
    for (int g = 1; g <= mg; ++g) XX(g) = lav.X(mg + 1 - g);
    XX(mg + 1) = Xs;
    for (int j = 2; j <= sm; ++j) XX(mg + j) = 0;
    tc2(1) = previn.tabove;
    for (int g = 1; g <= mg; ++g) tc2(1 + g) = lav.tc(mg + 1 - g);
    for (int j = 1; j <= sm; ++j) tc2(1 + mg + j) = previn.soiltc(j);
    tc2(N + 2) = previn.tsoil;
    Vec tn2 = Thomas(tc2, tsoil, tcan, k, cd, cfg.n, XX);
    Vec tnair(mg + 2);
    for (int g = 1; g <= mg + 2; ++g) tnair(g) = tn2(mg + 3 - g);
    for (int j = 1; j <= sm; ++j) tnsoil(j) = tn2(mg + 1 + j);
    Vec zth2(mg), Vmflux(mg), gt2(mg + 1), tn2b(mg);
    for (int g = 1; g <= mg; ++g) zth2(g) = lav.z(g + 1) - lav.z(g);
    for (int g = 1; g <= mg; ++g) { Vmflux(g) = (lav.ea(g) / pk) - lav.Vo(g); if (Vmflux(g) < 0) Vmflux(g) = 0; }
    for (int g = 1; g <= mg; ++g) { double gtmx = lav.ph(g) / zth2(g) * (2 * timestep); gt2(g) = lav.gt(g) > gtmx ? gtmx : lav.gt(g); }
    gt2(mg + 1) = lav.gt(mg + 1);
    for (int g = 1; g <= mg; ++g) tn2b(g) = tnair(g + 1);
    Vec Vn = ThomasV(lav.Vo, tn2b, pk, theta, thetap, relhum, tcan, tnsoil(1), zth2, gt2, Vmflux, cfg.n, previn, soilp);
    Vec vn(m);
    if (mg < m) {
      double TX = TT(m) + (pha / gt(m + 1)) * 2;
      Vec T2(mg + 2);
      T2(1) = 0; for (int g = 1; g <= mg; ++g) T2(g + 1) = lav.TT(g); T2(mg + 2) = TX;
      tn = layerinterp(T2, TT, tnair);
      vn = layerinterp(T2, TT, Vn);
    } else {
      for (int i = 1; i <= m; ++i) { tn(i) = tnair(i + 1); vn(i) = Vn(i + 1); }
    }
    for (int i = 1; i <= m; ++i) ea(i) = vn(i) * pk;
  } else {                                                                              // code.R:854-867
    Vec XX(sm, 0.0); XX(1) = Xs;
    int half = (int)std::nearbyint(m / 2.0);                // round(length(gha)/2, 0): half-to-even
    double cs = 0; for (int i = 1; i <= half; ++i) cs += 1 / gt(i);
    double gt2 = 1 / cs;
    double ka = gt2 * cpair(previn.tabove);
    Vec k(sm + 1), tc2(sm + 2);
    k(1) = ka; for (int j = 1; j <= sm; ++j) k(1 + j) = ksoil(j);
    tc2(1) = previn.tabove; for (int j = 1; j <= sm; ++j) tc2(1 + j) = previn.soiltc(j); tc2(sm + 2) = previn.tsoil;
    Vec tn2 = Thomas(tc2, tsoil, tcan, k, cdsoil, cfg.n, XX);
    for (int j = 1; j <= sm; ++j) tnsoil(j) = tn2(1 + j);
    double TX = TT(m) + (pha / gt(m + 1)) * 2;
    double eair = (relhum / 100) * 0.6108 * std::exp(17.27 * tcan / (tcan + 237.3));
    for (int i = 1; i <= m; ++i) {
      double wgt = TT(i) / TX;
      tn(i) = (1 - wgt) * tnsoil(1) + wgt * tcan;
      ea(i) = (1 - wgt) * tln.esoilv((i - 1) % tln.esoilv.n + 1) + wgt * eair;      // tln$esoil recycled (Q24)
    }
  }
  if (tnsoil(1) < tfrost) tnsoil(1) = tfrost;
  if (tnsoil(1) > tair + 30) tnsoil(1) = tair + 30;
  Vec rh(m);
  for (int i = 1; i <= m; ++i) {
    double es = 0.6108 * std::exp(17.27 * tn(i) / (tn(i) + 237.3));
    rh(i) = (ea(i) / es) * 100;
    if (rh(i) > 100) rh(i) = 100;
    if (rh(i) < 0) rh(i) = 0;
  }
  double shade = (1 - std::exp(-Rabss.ref(m) * sumPAI));
  double alb = shade * Rabss.ref(m) + (1 - shade) * vegp.refg;
  double tkg = 0.5 * tnsoil(1) + 0.5 * ps1 + 273.15;
  double Rlw = 5.67 * 1e-8 * 0.97 * ((tkg * tkg) * (tkg * tkg));
  Vec L(m);
  for (int i = 1; i <= m; ++i) L(i) = tln.L(i) * vegp.PAI(i);        // Q1: the "clamp" on the next R token is a no-op
  double Lt = vsum(L);
  int sel = 1; { double mn = 1e300; for (int i = 1; i <= m; ++i) { double zd = std::fabs(z(i) - d + 0.2 * zm); if (zd < mn) { mn = zd; sel = i; } } }
  double gt2g = gcanopy(uh, d + 0.2 * zm, 0, tn(sel), tnsoil(1), hgt, sumPAI, vegp.cd, vmean(vegp.iw), phi_m, pk);
  double G = gt2g * cp(1) * (tn(sel) - tnsoil(1));
  double Gmx = std::fabs((1 - alb) * Rsw - Rlw - Lt);
  if (G > Gmx) G = Gmx;
  if (G < -Gmx) G = -Gmx;
  H = (1 - alb) * Rsw - Rlw - G - Lt;
  for (int i = 1; i <= m; ++i) { double tdws = dewpoint(tln.ea(i), tn(i)); if (tn(i) < tdws) tn(i) = tdws; }   // Q12
  out.m = m; out.sm = sm;
  out.Rswin = Vec(m); out.Rlwin = Vec(m);
  for (int i = 1; i <= m; ++i) {
    out.Rswin(i) = Rabss.aRsw(i) / (1 - Rabss.ref(i));
    double lwabs, lwin; canlw(tn(i), sumPAI, 1 - vegp.vegem, cv.skyem, 0, lwabs, lwin);
    out.Rlwin(i) = lwin;
  }
  out.tc = tn; out.soiltc = tnsoil; out.tleaf = tln.tleaf; out.tabove = tcan; out.uz = uz; out.z = z; out.sz = sz;
  out.zabove = zabove; out.rh = rh; out.relhum = relhum; out.tair = tair; out.tsoil = tsoil; out.pk = pk;
  out.Rabs = Rabs; out.gt = gt; out.gv = gv; out.gha = gha; out.H = H; out.L = L; out.G = G; out.psi_m = psi_m;
  return 0;
}

}  // namespace orc
