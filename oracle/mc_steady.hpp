// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// Literal CPU restatement of the reference's steady-state path, evaluated per (column, hour) point
// (the R code is vectorised over the T hours and has no time recurrence):
//   /root/reference/R/NicheMapRcode.R  .leafabs2 361-376, .windprofile 378-404, tleafS 432-493,
//   .snowtempf 495-512, .runmodelsnow 514-677, runmodelS 717-885.
// Quirks Q14-Q17 of SURVEY Appendix B are reproduced (Q16's double subset is a whole-series
// re-indexing and is applied by the host wrapper, see microclimc_b200/api.py `_q16_remap`).
// PARITY UNPINNED vs real R (see mc_compat.hpp header).
#pragma once
#include "mc_model.hpp"

namespace orc {

// .leafabs2 (NMR.R:361-376); returns aRsw + aRlw
inline double leafabs2(double Rsw, Tme tme, double tair, double /*tground*/, double lat, double lon, double PAIt,
                       double PAIu, double pLAI, double x, double refls, double refw, double refg, double vegem,
                       double skyem, double dp, double merid, double dst = 0, double clump = 0) {
  double sa = solalt(tme.lt, lat, lon, tme.jd, merid, dst);
  double sa2 = sa < 5 ? 5 : sa;
  double ref = pLAI * refls + (1 - pLAI) * refw;
  double sunl = psunlit(PAIu, x, sa, clump);
  double mul = radmult(x, sa2);
  double aRsw = (1 - ref) * cansw_sa(Rsw, dp, sa, PAIu, x, ref, clump);
  double aGround = refg * cansw_sa(Rsw, dp, sa, PAIt, x, ref, clump);
  aGround = (1 - ref) * cansw_sa(aGround, dp, sa, PAIt, x, ref, clump);
  aRsw = sunl * mul * aRsw + (1 - sunl) * aRsw + aGround;
  double lwabs, lwin;
  canlw(tair, PAIu, 1 - vegem, skyem, clump, lwabs, lwin);
  return aRsw + lwabs;
}

// .windprofile (NMR.R:378-404), scalar zo; Q14 exponent grouping and the `zi` test kept verbatim
inline double windprofile2(double ui, double zi, double zo, double a, double hgt, double PAI, double psi_m = 0,
                           double hgtg = NA, double zm0 = 0.004) {
  if (is_na(hgtg)) hgtg = 0.05 * hgt;
  double d = zeroplanedis(hgt, PAI);
  double zm = roughlength(hgt, PAI, zm0);
  double ln2 = std::log((zi - d) / zm) + psi_m;
  if (ln2 < 0.55) ln2 = 0.55;
  double uf = (0.4 * ui) / ln2;
  double uo;
  if (zo >= hgt) {
    ln2 = std::log((zo - d) / zm) + psi_m;
    if (ln2 < 0.55) ln2 = 0.55;
    uo = (uf / 0.4) * ln2;
  } else {
    ln2 = std::log((hgt - d) / zm) + psi_m;
    if (ln2 < 0.55) ln2 = 0.55;
    double uh = (uf / 0.4) * ln2;
    if (zi >= (0.1 * hgt)) {
      uo = uh * std::exp(a * (zo / hgt) - 1);
    } else {
      uo = uh * std::exp(a * ((0.1 * hgt) / hgt) - 1);
      double zmg = 0.1 * hgtg;
      double ln1 = std::log((0.1 * hgt) / zmg) + psi_m;
      double ufb = 0.4 * (uo / ln1);
      double ln2b = std::log(zo / zmg) + psi_m;
      uo = (ufb / 0.4) * ln2b;
    }
  }
  return uo;
}

// tleafS (NMR.R:432-493)
struct TleafS { double tleaf, tn, rh; };
inline TleafS tleafS(double tair, double tground, double relhum, double pk, double theta, double gtt, double gt0,
                     double gha, double gv, double gL, double Rabs, double vegem, double soilb, double Psie,
                     double Smax, double surfwet, double leafdens) {
  double cp = cpair(tair);
  double aL = (gtt * (tair + 273.15) + gt0 * (tground + 273.15)) / (gtt + gt0);
  double bL = (leafdens * gL) / (gtt + gt0);
  double es = satvap(tair);
  double eref = (relhum / 100) * es;
  double rhsoil = soilrh(theta, soilb, Psie, Smax, tground);
  double esoil = rhsoil * satvap(tground);
  double sb = 5.67 * 1e-8;
  double tk = tair + 273.15;
  double Rnet = Rabs - 0.97 * sb * ((tk * tk) * (tk * tk));
  double tle = tair + (0.5 * Rnet) / (cp * gha);
  double wgt1 = leafdens < 1 ? leafdens / 2 : 1 - 0.5 / leafdens;
  double wgt2 = 1 - wgt1;
  double tapprox = wgt1 * tle + wgt2 * tair;
  double delta = 4098 * (0.6108 * std::exp(17.27 * tapprox / (tapprox + 237.3))) / ((tapprox + 237.3) * (tapprox + 237.3));
  double ae = (gtt * eref + gt0 * esoil + gv * es) / (gtt + gt0 + gv);
  double be = (gv * delta) / (gtt + gt0 + gv);
  double bH = gha * cp;
  double lambda = (-42.575 * tair + 44994);
  double aX = ((lambda * gv) / pk) * (surfwet * es - ae);
  double bX = ((lambda * gv) / pk) * (surfwet * delta - be);
  if (aX < 0) aX = 0;
  if (bX < 0) bX = 0;
  double aR = sb * vegem * ((aL * aL) * (aL * aL));
  double bR = 4 * vegem * sb * ((aL * aL * aL) * bL + (tk * tk * tk));
  double dTL = (Rabs - aR - aX) / (1 + bR + bX + bH);
  double tn = aL - 273.15 + bL * dTL;
  double tleaf = tn + dTL;
  double eanew = ae + be * dTL;
  if (eanew < 0.01) eanew = 0.01;
  double tmin = dewpoint(eanew, tn, true);
  double esnew = satvap(tn);
  if (eanew > esnew) eanew = esnew;
  double rh = (eanew / esnew) * 100;
  if (tleaf < tmin) tleaf = tmin;
  if (tn < tmin) tn = tmin;
  double tmax = (tn + 20 < 80) ? tn + 20 : 80;
  if (tleaf > tmax) tleaf = tmax;
  double tmx = std::max(tground + 5, tair + 5);
  if (tn > tmx) tn = tmx;
  tmx = std::max(tground + 20, tair + 20);
  if (tleaf > tmx) tleaf = tmx;
  double tmn = tair - 7;
  if (tn < tmn) tn = tmn;
  tmn = tair - 20;
  if (tleaf < tmn) tleaf = tmn;
  return {tleaf, tn, rh};
}

// .snowtempf (NMR.R:495-512) for one hour. `snowdepth` in metres (already /100 by the caller; the
// second /100 at NMR.R:503 is the reference's). SN[0..7] = SN1..SN8. Returns false when the layer
// lookup would index past SN8 (an R "undefined columns" error: reqhgt < 0.025 m) or before SN1.
inline bool snowtempf(double snowdepth, const double* SN, double reqhgt, double& stemp) {
  static const double nd[9] = {3.0, 2.0, 1.0, 0.5, 0.2, 0.1, 0.05, 0.025, 0.0};   // rev(c(0,2.5,...,300))/100
  int selb = 0;
  for (int i = 1; i <= 9; ++i) if (nd[i - 1] - reqhgt <= 0) { selb = i; break; }
  int sela = selb - 1;
  if (sela < 1 || selb > 8) return false;
  double sm = std::fabs(reqhgt - nd[sela - 1]) + std::fabs(reqhgt - nd[selb - 1]);
  double wgt1 = 1 - std::fabs(reqhgt - nd[sela - 1]) / sm;
  double wgt2 = 1 - std::fabs(reqhgt - nd[selb - 1]) / sm;
  if (nd[sela - 1] > (snowdepth / 100)) { wgt1 = 0; wgt2 = 1; }
  stemp = wgt1 * SN[sela - 1] + wgt2 * SN[selb - 1];     // snow[, sela+2] == SN<sela>
  if (snowdepth < reqhgt) stemp = 0;
  return true;
}

struct SteadyVeg {   // what runmodelS reads from vegp, after the per-hour reductions at NMR.R:730-760
  double hgt, PAIt, LAI, PAIu, x, lw, cd, iwmean, refls, refw, refg, vegem, gsmax, q50, clump;   // LAI = mean(pLAI*PAI)
};
struct SteadyClim { double tair, relhum, pk, u, swrad, difrad, skyem; Tme tme; };
struct SteadyNmr { double tground, theta, snowdep_cm; const double* SN; };   // D0cm, WC0cm, SNOWDEP, SN1..8
struct SteadyOut { double Tloc, tleaf, RHloc, RSWloc, RLWloc, windspeed, dp; };
struct SteadyCfg { double reqhgt, lat, lon, windhgt = 2, surfwet = 1, groundem = 0.95; int metopen = 1; };

// Per-hour vegetation reductions (NMR.R:730-760).  PAI, pLAI: m layer values of this hour.
// `snowfloor`: .runmodelsnow floors PAIt at 0.001 and maps NaN pLAI to 0.8 (NMR.R:531,535).
inline void steady_veg_reduce(const double* PAIin, const double* pLAIin, int m, double hgt, double reqhgt,
                              double& PAIt, double& LAIout, double& PAIu) {
  double PAI[MAXM + 1];
  for (int i = 1; i <= m; ++i) { PAI[i] = PAIin[i - 1]; if (PAI[i] < 0.0001) PAI[i] = 0.0001; }   // NMR.R:730
  PAIt = 0; double LAI = 0;
  for (int i = 1; i <= m; ++i) { PAIt += PAI[i]; LAI += pLAIin[i - 1] * PAI[i]; }
  LAI = LAI / m;                                                                                // apply(LAI,2,mean)
  LAIout = LAI;                                                                                 // pLAI <- LAI/PAIt by the caller
  double z[MAXM + 1];
  for (int i = 1; i <= m; ++i) z[i] = (i - 0.5) / m * hgt;
  if (reqhgt < hgt) {
    int first = 0, nsel = 0;
    for (int i = 1; i <= m; ++i) if (z[i] > reqhgt) { if (!first) first = i; ++nsel; }
    if (nsel > 1) {
      // NMR.R:741-751: `if (length(wgt2) > 1)` is never true (wgt2 is a scalar, or numeric(0) when
      // sel[1]-1 == 0), so the interpolation towards the layer below is dead code: PAIu <- PAIu1.
      double PAIu1 = 0; for (int i = first; i <= m; ++i) PAIu1 += PAI[i];
      PAIu = PAIu1;
    } else {
      double zu = z[m];
      double wgt = (hgt - reqhgt) / (hgt - zu);
      PAIu = wgt * PAI[m];
    }
  } else PAIu = 0;
}

// .runmodelsnow (NMR.R:514-677) for one hour. Returns false if the snow-layer lookup is out of range.
inline bool runmodelsnow_point(const SteadyClim& c, SteadyVeg v, const Soilp& soilp, const SteadyNmr& nm,
                               const SteadyCfg& cfg, SteadyOut& o) {
  double reqhgt = cfg.reqhgt, lat = cfg.lat, lon = cfg.lon;
  double snowtemp = nm.SN[0];
  double snowdep = nm.snowdep_cm / 100;
  double tair = c.tair, relhum = c.relhum, pk = c.pk, u = c.u, hgt = v.hgt;
  double PAIt = v.PAIt; if (PAIt < 0.001) PAIt = 0.001;
  double pLAI = v.LAI / PAIt; if (is_na(pLAI)) pLAI = 0.8;
  double PAIu = v.PAIu;
  double u2;
  if (cfg.metopen) {
    if (cfg.windhgt != 2) u = u * 4.87 / std::log(67.8 * cfg.windhgt - 5.42);
    u2 = u * std::log(67.8 * (hgt + 2) - 5.42) / std::log(67.8 * 2 - 5.42);
  } else {
    u2 = u;
    if (cfg.windhgt != (2 + hgt)) u2 = windprofile(u, cfg.windhgt, hgt + 2, 2, PAIt, hgt);
  }
  if (u2 < 0.5) u2 = 0.5;
  double zm = hgt > snowdep ? roughlength(hgt, PAIt) : 0.003;
  double d = hgt > snowdep ? zeroplanedis(hgt, PAIt) : snowdep - 0.003;
  double hgt2 = hgt > snowdep ? hgt : snowdep;
  double a1 = attencoef(hgt, PAIt, v.cd, v.iwmean);
  double a2 = attencoef(hgt, 0, v.cd, v.iwmean);
  double uz1 = windprofile2(u2, hgt + 2, reqhgt, a1, hgt, PAIt);
  double uz2 = windprofile2(u2, snowdep + 2, reqhgt, a2, 0, 0);
  double uz = hgt > snowdep ? uz1 : uz2;
  double uf = (0.4 * u2) / std::log((hgt2 + 2 - d) / zm);
  if (uf < 0.065) uf = 0.065;
  hgt2 = hgt < snowdep ? snowdep : hgt;
  double uh = (uf / 0.4) * std::log((hgt2 - d) / zm);
  if (uh < uf) uh = uf;
  bool below = reqhgt < snowdep;
  double tz2 = 0;
  if (below) { if (!snowtempf(snowdep, nm.SN, reqhgt, tz2)) return false; }
  double leafdens = (PAIt / hgt) * 1.2;
  double dp = c.difrad / c.swrad;
  if (is_na(dp)) dp = 0.5;
  if (std::isinf(dp)) dp = 0.5;
  if (dp > 1) dp = 1;
  double gtt = gturb(u2, hgt + 2, hgt + 2, hgt, hgt, PAIt, tair, 0, 0, 0.004, pk);
  double tz, tleaf, rh, Rsw, Rlw;
  double sb = 5.67 * 1e-8;
  if (reqhgt >= hgt) {
    double gt0 = gcanopy(uh, hgt, 0, tair, tair, hgt, PAIt, v.cd, v.iwmean, 1, pk);
    double gha = 1.41 * gforcedfree(v.lw * 2 * 0.71, uh, tair, 5, pk, 5);
    double gC = layercond(c.swrad, v.gsmax, v.q50);
    double gv = 1 / (1 / gC + 1 / gha);
    double gtz = phair(tair) * uh;
    double gL = 1 / (1 / gha + 1 / gtz);
    double Rabs = leafabs2(c.swrad, c.tme, tair, snowtemp, lat, lon, PAIt, 0, pLAI, v.x, 0.95, 0.95, 0.95, 0.85,
                           c.skyem, dp, /*merid <- clump, Q15*/ v.clump);
    TleafS Th = tleafS(tair, snowtemp, relhum, pk, 0.5, gtt, gt0, gha, gv, gL, Rabs, 0.85, soilp.b, soilp.psi_e,
                       soilp.Smax, 1, leafdens);
    double tn = Th.tn; tleaf = Th.tleaf;
    if (reqhgt == hgt) { tz = tn; rh = Th.rh; }
    else {
      double gtc = gturb(u2, hgt + 2, reqhgt, hgt, hgt, PAIt, tair, 0, 0, 0.004, pk);
      double gt2 = 1 / (1 / gtt - 1 / gtc);
      double eref = (relhum / 100) * satvap(tair);
      double ecan = (Th.rh / 100) * satvap(tn);
      double ea = (gt2 * eref + gtc * ecan) / (gt2 + gtc);
      tz = (gt2 * tair + gtc * tn) / (gt2 + gtc);
      rh = (ea / satvap(tz)) * 100;
    }
    if (rh > 100) rh = 100;
    Rsw = c.swrad;
    double tk = tz + 273.15;
    Rlw = (1 - c.skyem) * 0.85 * sb * ((tk * tk) * (tk * tk));
  } else {
    double gtc = gcanopy(uh, hgt, reqhgt, tair, tair, hgt, PAIt, v.cd, v.iwmean, 1, pk);
    gtt = 1 / (1 / gtt + 1 / gtc);
    double gt0 = gcanopy(uh, reqhgt, 0, tair, tair, hgt, PAIt, v.cd, v.iwmean, 1, pk);
    double gha = 1.41 * gforcedfree(v.lw * 0.71 * 2, uz, tair, 5, pk, 5);
    double merid_def = std::nearbyint(lon / 15) * 15;         // cansw default merid = round(long/15, 0) * 15
    double sa = solalt(c.tme.lt, lat, lon, c.tme.jd, merid_def, 0);
    double PAR = cansw_sa(c.swrad, dp, sa, PAIu, v.x, 0.6);
    double gC = layercond(PAR, v.gsmax, v.q50);
    double gv = 1 / (1 / gC + 1 / gha);
    double gtz = phair(tair) * uz;
    double gL = 1 / (1 / gha + 1 / gtz);
    double Rabs = leafabs2(c.swrad, c.tme, tair, snowtemp, lat, lon, PAIt, PAIu, pLAI, v.x, 0.95, 0.95, 0.95, 0.85,
                           c.skyem, dp, /*Q15*/ v.clump);
    TleafS Th = tleafS(tair, snowtemp, relhum, pk, 0.5, gtt, gt0, gha, gv, gL, Rabs, v.vegem, soilp.b, soilp.psi_e,
                       soilp.Smax, 1, leafdens);
    tz = Th.tn; tleaf = Th.tleaf; rh = Th.rh;
    Rsw = cansw_sa(c.swrad, dp, sa, PAIu, v.x, 0.95);
    double Rdif = cantransdif(PAIu, 0.95) * c.swrad * dp;
    dp = Rdif / Rsw;
    if (is_na(dp)) dp = 1;
    if (std::isinf(dp)) dp = 1;
    if (dp < 0) dp = 0;
    if (dp > 1) dp = 1;
    double lwabs; canlw(tair, PAIu, 0.85, c.skyem, v.clump, lwabs, Rlw);
  }
  if (below) {
    tz = tz2; tleaf = tz2; rh = 100; uz = 0; Rsw = 0;
    double tk = tz2 + 273.15;
    Rlw = sb * 0.85 * ((tk * tk) * (tk * tk));
  }
  o = {tz, tleaf, rh, Rsw, Rlw, uz, dp};
  return true;
}

// runmodelS (NMR.R:717-885) for one hour, including the snow overlay (NMR.R:866-883).
// Returns 0 ok, -2 snow-layer lookup out of range.
inline int runmodelS_point(const SteadyClim& c, const SteadyVeg& v, const Soilp& soilp, const SteadyNmr& nm,
                           const SteadyCfg& cfg, SteadyOut& o) {
  double reqhgt = cfg.reqhgt, lat = cfg.lat, lon = cfg.lon;
  double tair = c.tair, relhum = c.relhum, pk = c.pk, u = c.u, hgt = v.hgt;
  double PAIt = v.PAIt, pLAI = v.LAI / v.PAIt, PAIu = v.PAIu;
  double u2;
  if (cfg.metopen) {
    if (cfg.windhgt != 2) u = u * 4.87 / std::log(67.8 * cfg.windhgt - 5.42);
    u2 = u * std::log(67.8 * (hgt + 2) - 5.42) / std::log(67.8 * 2 - 5.42);
  } else {
    u2 = u;
    if (cfg.windhgt != (2 + hgt)) u2 = windprofile(u, cfg.windhgt, hgt + 2, 2, PAIt, hgt);
  }
  double cp = cpair(tair); (void)cp;
  double ph = phair(tair, pk); (void)ph;
  if (u2 < 0.5) u2 = 0.5;
  double zm = roughlength(hgt, PAIt);
  double d = zeroplanedis(hgt, PAIt);
  double uf = (0.4 * u2) / std::log((2 + hgt - d) / zm);
  double sb = 5.67 * 1e-8;
  double tk = tair + 273.15;
  double Rlw = (1 - c.skyem) * sb * v.vegem * ((tk * tk) * (tk * tk));
  double Rnet = (1 - 0.23) * c.swrad - Rlw;
  double H = 0.2 * Rnet;
  double psi_m, psi_h;
  diabatic_cor(tair, pk, H, uf, hgt + 2, d, psi_m, psi_h);
  if (psi_m > 2.5) psi_m = 2.5;
  if (psi_h > 2.5) psi_h = 2.5;
  if (psi_m < -2.5) psi_m = -2.5;
  if (psi_h < -2.5) psi_h = -2.5;
  double a = attencoef(hgt, PAIt, v.cd, v.iwmean);
  double uz = windprofile2(u2, hgt + 2, reqhgt, a, hgt, PAIt);
  double ln2 = std::log((hgt - d) / zm);
  if (ln2 < 0.55) ln2 = 0.55;
  double uh = (uf / 0.4) * ln2;
  double tground = nm.tground, theta = nm.theta;
  double leafdens = PAIt / hgt;
  double dp = c.difrad / c.swrad;
  if (is_na(dp)) dp = 0.5;
  if (std::isinf(dp)) dp = 0.5;
  if (dp > 1) dp = 1;
  double gtt = gturb(u2, hgt + 2, hgt + 2, hgt, hgt, PAIt, tair, psi_m, psi_h, 0.004, pk);
  double tz, tleaf, rh, Rsw;
  if (reqhgt >= hgt) {
    double gt0 = gcanopy(uh, hgt, 0, tair, tair, hgt, PAIt, v.cd, v.iwmean, 1, pk);
    double gha = 1.41 * gforcedfree(v.lw * 0.71, uh, tair, 5, pk, 5);
    double gC = layercond(c.swrad, v.gsmax, v.q50);
    double gv = 1 / (1 / gC + 1 / gha);
    double gtz = phair(tair) * uh;
    double gL = 1 / (1 / gha + 1 / gtz);
    double Rabs = leafabs2(c.swrad, c.tme, tair, tground, lat, lon, PAIt, 0, pLAI, v.x, v.refls, v.refw, v.refg,
                           v.vegem, c.skyem, dp, /*Q15*/ v.clump);
    TleafS Th = tleafS(tair, tground, relhum, pk, theta, gtt, gt0, gha, gv, gL, Rabs, v.vegem, soilp.b, soilp.psi_e,
                       soilp.Smax, cfg.surfwet, leafdens);
    double tn = Th.tn; tleaf = Th.tleaf;
    if (reqhgt == hgt) { tz = tn; rh = Th.rh; }
    else {
      double gtc = gturb(u2, hgt + 2, reqhgt, hgt, hgt, PAIt, tair, psi_m, psi_h, 0.004, pk);
      double gt2 = 1 / (1 / gtt - 1 / gtc);
      double eref = (relhum / 100) * satvap(tair);
      double ecan = (Th.rh / 100) * satvap(tn);
      double ea = (gt2 * eref + gtc * ecan) / (gt2 + gtc);
      tz = (gt2 * tair + gtc * tn) / (gt2 + gtc);
      rh = (ea / satvap(tz)) * 100;
    }
    if (rh > 100) rh = 100;
    Rsw = c.swrad;
  } else {
    double gtc = gcanopy(uh, hgt, reqhgt, tair, tair, hgt, PAIt, v.cd, v.iwmean, 1, pk);
    gtt = 1 / (1 / gtt + 1 / gtc);
    double gt0 = gcanopy(uh, reqhgt, 0, tair, tair, hgt, PAIt, v.cd, v.iwmean, 1, pk);
    double gha = 1.41 * gforcedfree(v.lw * 0.71, uz, tair, 5, pk, 5);
    double merid_def = std::nearbyint(lon / 15) * 15;
    double sa = solalt(c.tme.lt, lat, lon, c.tme.jd, merid_def, 0);
    double PAR = cansw_sa(c.swrad, dp, sa, PAIu, v.x, 0.2);
    double gC = layercond(PAR, v.gsmax, v.q50);
    double gv = 1 / (1 / gC + 1 / gha);
    double gtz = phair(tair) * uz;
    double gL = 1 / (1 / gha + 1 / gtz);
    double Rabs = leafabs2(c.swrad, c.tme, tair, tground, lat, lon, PAIt, PAIu, pLAI, v.x, v.refls, v.refw, v.refg,
                           v.vegem, c.skyem, dp, /*Q15*/ v.clump);
    TleafS Th = tleafS(tair, tground, relhum, pk, theta, gtt, gt0, gha, gv, gL, Rabs, v.vegem, soilp.b, soilp.psi_e,
                       soilp.Smax, cfg.surfwet, leafdens);
    tz = Th.tn; tleaf = Th.tleaf; rh = Th.rh;
    Rsw = cansw_sa(c.swrad, dp, sa, PAIu, v.x, v.refls);
    double Rdif = cantransdif(PAIu, v.refls) * c.swrad * dp;
    dp = Rdif / Rsw;
    if (is_na(dp)) dp = 1;
    if (std::isinf(dp)) dp = 1;
    if (dp < 0) dp = 0;
    if (dp > 1) dp = 1;
    double lwabs; canlw(tair, PAIu, 1 - v.vegem, c.skyem, v.clump, lwabs, Rlw);
  }
  o = {tz, tleaf, rh, Rsw, Rlw, uz, dp};
  if (nm.snowdep_cm > 0) {                                  // NMR.R:873-883 (per-hour equivalent)
    SteadyOut s;
    if (!runmodelsnow_point(c, v, soilp, nm, cfg, s)) return -2;
    o = s;
  }
  return 0;
}

}  // namespace orc
