// Steady-state path on the device: runmodelS + .runmodelsnow + tleafS + .leafabs2 + .windprofile + .snowtempf
// (R/NicheMapRcode.R:717-885, 514-677, 432-493, 361-376, 378-404, 495-512), one thread per column looping over a
// slice of the hours.  There is no time recurrence on this path, so the (column, hour) points are
// independent; the time axis is split over blockIdx.y when there are too few columns to fill the GPU.
// Everything that depends only on the column (and the seasonal PAI factor) — displacement/roughness lengths, the
// log-law ratios, attenuation coefficient, canopy-conductance integrals, diffuse transmissions — is hoisted into
// `prep()`; the hourly part keeps only the forcing-dependent transcendentals.
#pragma once
#include "mc_step.cuh"
#ifndef MC_STEADY_SYNC
#define MC_STEADY_SYNC 1
#endif

namespace mc {

struct SteadyArgs {
  long long ncol; int m, sm; long long T;
  const double *vegp, *soilp, *lat, *lon;
  const double* clim;        // [T][9]  temp, relhum, pres, windspeed, swrad, difrad, skyem, jd, lt
  const double* nmr;         // [T][11] or [T][11][ncol]: D0cm, WC0cm, SNOWDEP(cm), SN1..SN8
  int nmr_percol;
  const double* season;      // [T] or null
  const double* vegt;        // [T][2m+1][nvt] or null: PAI[m], pLAI[m] (zm0 unused here) per hour (vegp$PAI as m x T matrix, NMR.R:730-735)
  long long nvt;             // 1: shared by all columns, ncol: per column
  const double* sol;         // [T][13] eot, sin / cos of the declination, dp, ph, ph0, es, eref, cp, (tair + 273.15)^4, gS, gT, lpk per hour (k_steady_hour), or null: computed in place
  double reqhgt; int metopen; double windhgt, surfwet;
  double* out;               // [T][7][ncol] or null
  double* stats;             // [6][ncol] (written by the reduce kernel)
  double* part;              // [nchunk][6][ncol] partial sums/min/max, or null
  int* flags;                // [ncol]; bit 2: snow-layer lookup out of range
};

// What an hour's forcing row determines for every column: the date-only solar terms, the diffuse fraction and the molar
// densities of air at the forcing pressure and at 101.3 kPa (k_steady_hour writes them once per hour; the same expressions
// run in place when there is no table).
struct HourPre { SolDay sd; double dp, ph, ph0, es, eref, cp, p4, gS, gT, lpk; };   // lpk = latent heat of vaporisation (J / mol) / pk   // es = satvap(tair), eref = relhum/100 * es, cp = cpair(tair), p4 = (tair + 273.15)^4
MC_HD HourPre hour_pre(const double* cr) {
  HourPre h;
  h.sd = sol_day(cr[7]);
  double dp = cr[5] / cr[4];                                   // difrad / swrad (IEEE division: 0/0 and x/0 are meaningful)
  if (dp != dp) dp = 0.5;
  if (isinf(dp)) dp = 0.5;
  if (dp > 1) dp = 1;
  h.dp = dp; h.ph = phair(cr[0], cr[2]); h.ph0 = phair(cr[0], 101.3);
  h.es = satvap(cr[0]); h.eref = fdiv(cr[1], 100) * h.es; h.cp = cpair(cr[0]); h.p4 = pow4(cr[0] + 273.15);
  {                                                            // hour-only pieces of gforcedfree (leaf boundary-layer conductance), see gha_hour
    const double Tk = cr[0] + 273.15;
    const double sc = pow175(Tk * (1 / 273.15)) * fdiv(101.3, cr[2]);
    h.gS = h.ph * fsqrt(sc);
    h.gT = fsqrt(frcp(Tk));
  }
  h.lpk = fdiv(-42.575 * cr[0] + 44994, cr[2]);
  return h;
}

template <int MM>
struct SteadyCol {
  int m;
  double hgt, x, lw, cd, refls, refw, refg, vegem, gsmax, q50, clump, iwmean, soilb, psi_e, Smax, lat, lon, slat, clat;
  double PAIb[MM + 2], pLAIb[MM + 2];
  // prepared (column + season)
  double PAIt, LAI, PAIu, d, zm3, zm4, lnuf, ilnuf, lnh3, a, l_m, La, leafdens;
  double wA, wB, wE, iwA, iwA1; // .windprofile factors: ln(zi), ln(zo or hgt), exp term; 1 / ln(zi) (and of the snow-surface variant)
  double gt_lnu, gt_lnh, gtc_lnh;   // gturb pieces for z1 = hgt+2 and z1 = reqhgt
  double gc_top_den, gc_c_den, gc_0_den;   // hgt*(e0-e1) of the three gcanopy calls
  double gd, gid, gd3a, gd_s, gid_s, gd3a_s;   // column-only pieces of gha_hour for the leaf dimension d = 0.71 lw (twice that under snow): Cf^2 d, 1.41 / d, sqrt(Cr^4 9.81 d^3 * 5)
  double ref, sref, kden, k5, emv, slw, trd_u, trd_t, trlw_u;
  double s02, trd02, sls, trdls, cdif;     // PAR (ref 0.2) and Rsw (ref refls) transmissions, cantransdif
  double merid_def;
  bool same_merid;
  // snow variants
  double a1, a2, PAIt_s, pLAI_s, zm3_s, d_s, wA1, wB1, wE1, wB2, leafdens_s, gts_lnu, gts_lnh, gtcs_lnh, gcs_top_den, gcs_c_den,
      gcs_0_den, l_m_s, trds_u, trds_t, s06, trd06, s95, trd95, cdif95, trlw85_u, sref_s, ref_s;

  MC_HD void load(const SteadyArgs& a, long long c) {
    m = a.m;
    auto V = [&](int r) { return a.vegp[(long long)r * a.ncol + c]; };
    hgt = V(V_HGT); x = V(V_X); lw = V(V_LW); cd = V(V_CD); refls = V(V_REFLS); refg = V(V_REFG); refw = V(V_REFW);
    vegem = V(V_VEGEM); gsmax = V(V_GSMAX); q50 = V(V_Q50); clump = V(V_CLUMP);
    double s = 0;
    for (int i = 1; i <= m; ++i) { PAIb[i] = V(V_NSCAL + i - 1); pLAIb[i] = V(V_NSCAL + m + i - 1); s += V(V_NSCAL + 2 * m + i - 1); }
    iwmean = s / m;
    auto S = [&](int r) { return a.soilp[(long long)r * a.ncol + c]; };
    Smax = S(S_SMAX); soilb = S(S_B); psi_e = S(S_PSIE);
    lat = a.lat[c]; lon = a.lon[c];
    sol_site(lat, slat, clat);
  }
  // hgt*(e0-e1) of gcanopy between z0 < z1 for attenuation a (phi_m = 1)
  MC_HD double gc_den(double z1, double z0, double av) const {
    return hgt * (exp(-av * (z0 / hgt - 1)) - exp(-av * (z1 / hgt - 1)));
  }
  MC_HD void prep(const SteadyArgs& A, double season) {
    const double reqhgt = A.reqhgt;
    // vegetation reductions (NMR.R:730-760)
    double cs = 0, ls = 0, pu = 0;
    int first = 0, nsel = 0;
    for (int i = 1; i <= m; ++i) if ((i - 0.5) / m * hgt > reqhgt) { if (!first) first = i; ++nsel; }
    double ptop = 0;
    for (int i = 1; i <= m; ++i) {
      double p = PAIb[i] * season; if (p < 0.0001) p = 0.0001;
      cs += p; ls += pLAIb[i] * p;
      if (nsel > 1 && i >= first) pu += p;
      if (i == m) ptop = p;
    }
    PAIt = cs; LAI = ls / m;
    if (reqhgt < hgt) {
      if (nsel > 1) PAIu = pu;
      else { double zu = (m - 0.5) / m * hgt; PAIu = ((hgt - reqhgt) / (hgt - zu)) * ptop; }
    } else PAIu = 0;
    La = log(67.8 * (hgt + 2) - 5.42);
    d = zeroplanedis(hgt, PAIt);
    zm3 = roughlength_d(hgt, PAIt, d, 0.003);
    zm4 = roughlength_d(hgt, PAIt, d, 0.004);
    lnuf = log((2 + hgt - d) / zm3); ilnuf = 1 / lnuf;
    lnh3 = log((hgt - d) / zm3); if (lnh3 < 0.55) lnh3 = 0.55;
    l_m = mixinglength(hgt, PAIt);
    a = attencoef_base(hgt, PAIt, cd, iwmean, l_m) * sqrt(1.0);
    leafdens = PAIt / hgt;
    wind_factors(hgt + 2, reqhgt, a, hgt, d, zm4, wA, wB, wE); iwA = 1 / wA;
    gturb_factors(hgt + 2, hgt + 2, hgt, d, zm4, gt_lnu, gt_lnh);
    { double dum; gturb_factors(hgt + 2, reqhgt, hgt, d, zm4, dum, gtc_lnh); }
    gc_top_den = gc_den(hgt, 0, a);
    gc_c_den = gc_den(hgt, reqhgt, a);
    gc_0_den = gc_den(reqhgt, 0, a);
    // radiation
    double pl = LAI / PAIt;
    ref = pl * refls + (1 - pl) * refw;
    sref = sqrt(1 - ref);
    kden = cank_den(x);
    k5 = cank(x, 5.0, kden);
    { const double d = lw * 0.71; gd = kGhaCf2 * d; gid = 1.41 / d; gd3a = sqrt(kGhaCr4 * ((9.81 * (d * d * d)) * 5)); }
    { const double d = lw * 2 * 0.71; gd_s = kGhaCf2 * d; gid_s = 1.41 / d; gd3a_s = sqrt(kGhaCr4 * ((9.81 * (d * d * d)) * 5)); }
    emv = 1 - (1 - vegem);
    slw = sqrt(emv);
    trd_u = clumptr(exp(-sref * PAIu), 0.0);            // Q15: clump lands in `merid`; cansw/canlw see clump = 0
    trd_t = clumptr(exp(-sref * PAIt), 0.0);
    trlw_u = clumptr(exp(-slw * PAIu), 0.0);
    s02 = sqrt(1 - 0.2); trd02 = exp(-s02 * PAIu);
    sls = sqrt(1 - refls); trdls = exp(-sls * PAIu);
    cdif = exp(-sqrt(1 - refls) * PAIu);
    merid_def = nearbyint(lon / 15) * 15;
    same_merid = (lon - merid_def) == (lon - clump);       // what solsin_pre sees of the meridian
    // ---- snow-surface variants (.runmodelsnow, NMR.R:514-677) ----
    PAIt_s = PAIt < 0.001 ? 0.001 : PAIt;
    pLAI_s = LAI / PAIt_s; if (pLAI_s != pLAI_s) pLAI_s = 0.8;
    zm3_s = roughlength(hgt, PAIt_s, 0.003);
    d_s = zeroplanedis(hgt, PAIt_s);
    l_m_s = mixinglength(hgt, PAIt_s);
    a1 = attencoef_base(hgt, PAIt_s, cd, iwmean, l_m_s) * sqrt(1.0);
    { double lm0 = mixinglength(hgt, 0.0); a2 = attencoef_base(hgt, 0.0, cd, iwmean, lm0) * sqrt(1.0); }
    { double zm4s = roughlength_d(hgt, PAIt_s, d_s, 0.004);
      wind_factors(hgt + 2, reqhgt, a1, hgt, d_s, zm4s, wA1, wB1, wE1); iwA1 = 1 / wA1;
      gturb_factors(hgt + 2, hgt + 2, hgt, d_s, zm4s, gts_lnu, gts_lnh);
      double dum; gturb_factors(hgt + 2, reqhgt, hgt, d_s, zm4s, dum, gtcs_lnh); }
    { double l2 = log((reqhgt - 0.0) / 0.004); if (l2 < 0.55) l2 = 0.55; wB2 = l2; }
    leafdens_s = (PAIt_s / hgt) * 1.2;
    gcs_top_den = gc_den(hgt, 0, a1); gcs_c_den = gc_den(hgt, reqhgt, a1); gcs_0_den = gc_den(reqhgt, 0, a1);
    ref_s = pLAI_s * 0.95 + (1 - pLAI_s) * 0.95;
    sref_s = sqrt(1 - ref_s);
    trds_u = exp(-sref_s * PAIu); trds_t = exp(-sref_s * PAIt_s);
    s06 = sqrt(1 - 0.6); trd06 = exp(-s06 * PAIu);
    s95 = sqrt(1 - 0.95); trd95 = exp(-s95 * PAIu);
    cdif95 = exp(-sqrt(1 - 0.95) * PAIu);
    trlw85_u = clumptr(exp(-sqrt(1 - 0.85) * PAIu), clump);
  }
  // .windprofile (NMR.R:378-404) split into u-independent factors: uo = ((0.4*ui)/wA / 0.4) * wB [* wE]
  MC_HD static void wind_factors(double zi, double zo, double av, double h, double dd, double zmm, double& fA, double& fB, double& fE) {
    double l = log((zi - dd) / zmm); if (l < 0.55) l = 0.55;
    fA = l;
    if (zo >= h) { double l2 = log((zo - dd) / zmm); if (l2 < 0.55) l2 = 0.55; fB = l2; fE = -1.0; }
    else { double l2 = log((h - dd) / zmm); if (l2 < 0.55) l2 = 0.55; fB = l2; fE = exp(av * (zo / h) - 1); }   // Q14
  }
  // the same with the column-only reciprocal 1 / fA given: ((0.4 ui) / fA / 0.4) fB = ui (1 / fA) fB up to rounding
  MC_HD static double wind_apply_r(double ui, double ifA, double fB, double fE) {
    const double uo = (ui * ifA) * fB;
    return fE < 0 ? uo : uo * fE;
  }
  MC_HD static double wind_apply(double ui, double fA, double fB, double fE) {
    double uf = fdiv(0.4 * ui, fA);
    double uo = fdiv(uf, 0.4) * fB;
    return fE < 0 ? uo : uo * fE;
  }
  MC_HD static void gturb_factors(double zu, double z1, double z0, double dd, double zmm, double& lnu, double& lnh) {
    double zh = 0.2 * zmm;
    lnu = log((zu - dd) / zmm);
    double lo = z0 - dd; if (!(lo > zh)) lo = zh;
    lnh = log((z1 - dd) / lo);
  }
  MC_HD static double gturb_apply(double u, double lnu, double lnh, double ph, double psi_m, double psi_h) {
    double ln = lnu + psi_m; if (!(ln >= 0.55)) ln = 0.55;
    double den = lnh + psi_h; if (den < 0.25 * lnh) den = 0.25 * lnh;
    return fdiv((0.4 * 0.4) * ph * u, ln * den);                  // 0.4 ph uf / den with uf = 0.4 u / ln: one division
  }
  // gcanopy with tc1 == tc0 (free-convection floor is exactly 0)
  MC_HD static double gcan_apply(double lmix, double iw, double ph, double uh, double av, double den) {
    double g = fdiv(lmix * iw * ph * uh * av, den * 1.0);
    return (g > 0.0) ? g : 0.0;
  }
  // .leafabs2 (NMR.R:361-376) with merid <- clump (Q15), clump = 0
  MC_HD double leafabs2(double Rsw, double sa, double K, double mul, double tair, double PAIt_, double PAIu_, double refv, double srefv,
                        double trdu, double trdt, double refgv, double emvv, double trlwu, double skyem, double dp, double p4) const {
    double sunl = psunlit_k(PAIu_, sa, K, 0.0);
    double aRsw = (1 - refv) * cansw_k(Rsw, dp, sa, PAIu_, srefv, trdu, K, 0.0);
    double aGround = refgv * cansw_k(Rsw, dp, sa, PAIt_, srefv, trdt, K, 0.0);
    aGround = (1 - refv) * cansw_k(aGround, dp, sa, PAIt_, srefv, trdt, K, 0.0);
    aRsw = sunl * mul * aRsw + (1 - sunl) * aRsw + aGround;
    double lwout = kSb * p4;
    double lwin = trlwu * skyem * lwout + (1 - trlwu) * emvv * lwout;
    return aRsw + emvv * lwin;
  }
  // 1.41 * gforcedfree(d, u, tair, 5, pk, 5) (mc_phys.cuh gforcedfree) with what the forced and the free branch share taken out of the
  // maximum, as gha_fast2 of the streaming kernel: with Dh = 18.9e-6 sc, nu = 13.3e-6 sc
  //   1.41 max(gforced, gfree) = (1.41 / d) ph sqrt(sc) sqrt(max(Cf^2 d u, sqrt(Cr^4 9.81 d^3 5 / Tk)))
  // and here |dT| = 5 is fixed, so the free-convection term is a column-only factor cB = sqrt(Cr^4 9.81 d^3 5) times an hour-only one
  // gT = sqrt(1 / Tk): one square root per call instead of three.  cA = Cf^2 d, cI = 1.41 / d (prep()); gS = ph sqrt(sc) (HourPre).
  MC_HD static double gha_hour(double cA, double cI, double cB, double u, const HourPre& hp) {
    const double A = cA * u, B = cB * hp.gT;
    return (cI * hp.gS) * fsqrt((A > B) ? A : B);
  }
  // 1 / (1/a + 1/b) as a*b / (a+b): one division; a == 0 (closed stomata at night) still gives 0 as the IEEE form does
  MC_HD static double hpar(double a, double b) { return fdiv(a * b, a + b); }
  // tleafS (NMR.R:432-493)
  MC_HD static void tleafS(double tair, double tground, double relhum, double pk, double theta, double gtt, double gt0, double gha,
                           double gv, double gL, double Rabs, double vegemv, double soilbv, double Psie, double Smaxv, double surfwet,
                           double leafd, const HourPre& hp, double& tleaf, double& tn, double& rh) {
    double cp = hp.cp;
    double gs = gtt + gt0;
    const double igs = frcp(gs);
    double aL = (gtt * (tair + 273.15) + gt0 * (tground + 273.15)) * igs;
    double bL = (leafd * gL) * igs;
    double es = hp.es;
    double eref = hp.eref;
    double esoil = soilrh(theta, soilbv, Psie, Smaxv, tground) * satvap(tground);
    double tk = tair + 273.15;
    double Rnet = Rabs - 0.97 * kSb * hp.p4;
    double tle = tair + fdiv(0.5 * Rnet, cp * gha);
    double wgt1 = leafd < 1 ? leafd / 2 : 1 - fdiv(0.5, leafd);
    double tapprox = wgt1 * tle + (1 - wgt1) * tair;
    const double r237 = frcp(tapprox + 237.3);                    // Tetens and its slope share the reciprocal
    double delta = (4098 * (0.6108 * fexp_b((17.27 * tapprox) * r237))) * (r237 * r237);
    double g3 = gs + gv;
    const double ig3 = frcp(g3);
    double ae = (gtt * eref + gt0 * esoil + gv * es) * ig3;
    double be = (gv * delta) * ig3;
    double bH = gha * cp;
    double lg = hp.lpk * gv;
    double aX = lg * (surfwet * es - ae);
    double bX = lg * (surfwet * delta - be);
    if (aX < 0) aX = 0;
    if (bX < 0) bX = 0;
    double aR = kSb * vegemv * pow4(aL);
    double bR = 4 * vegemv * kSb * ((aL * aL * aL) * bL + (tk * tk * tk));
    double dTL = fdiv(Rabs - aR - aX, 1 + bR + bX + bH);
    tn = aL - 273.15 + bL * dTL;
    tleaf = tn + dTL;
    double eanew = ae + be * dTL;
    if (eanew < 0.01) eanew = 0.01;
    double tmin = dewpoint(eanew, tn);
    double esnew = satvap(tn);
    if (eanew > esnew) eanew = esnew;
    rh = fdiv(eanew, esnew) * 100;
    if (tleaf < tmin) tleaf = tmin;
    if (tn < tmin) tn = tmin;
    double tmax = (tn + 20 < 80) ? tn + 20 : 80;
    if (tleaf > tmax) tleaf = tmax;
    double tmx = dmax(tground + 5, tair + 5);
    if (tn > tmx) tn = tmx;
    tmx = dmax(tground + 20, tair + 20);
    if (tleaf > tmx) tleaf = tmx;
    if (tn < tair - 7) tn = tair - 7;
    if (tleaf < tair - 20) tleaf = tair - 20;
  }
  MC_HD double wind2m(const SteadyArgs& A, double u, double PAIv) const {
    double u2;
    if (A.metopen) {
      if (A.windhgt != 2) u = fdiv(u * 4.87, flog(67.8 * A.windhgt - 5.42));
      u2 = fdiv(u * La, log(67.8 * 2 - 5.42));
    } else {
      u2 = u;
      if (A.windhgt != (2 + hgt)) u2 = Column<3, 2>::windprofile(u, A.windhgt, hgt + 2, 2, PAIv, hgt, 0, 0.05 * hgt, 0.004);
    }
    if (u2 < 0.5) u2 = 0.5;
    return u2;
  }

  // one hour without snow (runmodelS body, NMR.R:761-865); o[7] = Tloc, tleaf, RHloc, RSWloc, RLWloc, windspeed, dp
  MC_HD void point(const SteadyArgs& A, const double* cr, const HourPre& hp, double tground, double theta, double* o) const {
    const double reqhgt = A.reqhgt;
    const double tair = cr[0], relhum = cr[1], pk = cr[2], swrad = cr[4], skyem = cr[6];
    const double u2 = wind2m(A, cr[3], PAIt);
    const double uf = (0.4 * u2) * ilnuf;
    double Rlw = (1 - skyem) * kSb * vegem * hp.p4;
    const double Hh = 0.2 * ((1 - 0.23) * swrad - Rlw);
    double psi_m, psi_h;
    diabatic_cor(tair, pk, Hh, uf, hgt + 2, d, psi_m, psi_h);
    if (psi_m > 2.5) psi_m = 2.5;
    if (psi_h > 2.5) psi_h = 2.5;
    if (psi_m < -2.5) psi_m = -2.5;
    if (psi_h < -2.5) psi_h = -2.5;
    const double uz = wind_apply_r(u2, iwA, wB, wE);
    const double uh = (uf * 2.5) * lnh3;
    double dp = hp.dp;
    const SolDay& sd = hp.sd;
    const double ph = hp.ph;
    double gtt = gturb_apply(u2, gt_lnu, gt_lnh, ph, psi_m, psi_h);
    const double sa_l = solsin_pre(cr[8], lon, /*merid <- clump (Q15)*/ clump, 0, sd, slat, clat);
    double K_l, mul;
    cank2_sh(x, sa_l, kden, k5, K_l, mul);       // sa_l, sa: sin(solar altitude), see mc_phys.cuh
    const double ph0 = hp.ph0;
    double tz, tleaf, rh, Rsw;
    if (reqhgt >= hgt) {
      double gt0 = gcan_apply(l_m, iwmean, ph, uh, a, gc_top_den);
      double gha = gha_hour(gd, gid, gd3a, uh, hp);
      double gC = layercond(swrad, gsmax, q50);
      double gv = hpar(gC, gha);
      double gL = hpar(gha, ph0 * uh);
      double Rabs = leafabs2(swrad, sa_l, K_l, mul, tair, PAIt, 0.0, ref, sref, 1.0, trd_t, refg, emv, 1.0, skyem, dp, hp.p4);
      double tn, rhn;
      tleafS(tair, tground, relhum, pk, theta, gtt, gt0, gha, gv, gL, Rabs, vegem, soilb, psi_e, Smax, A.surfwet, leafdens, hp, tleaf, tn, rhn);
      if (reqhgt == hgt) { tz = tn; rh = rhn; }
      else {
        double gtc = gturb_apply(u2, gt_lnu, gtc_lnh, ph, psi_m, psi_h);
        double gt2 = 1 / (1 / gtt - 1 / gtc);
        double eref = hp.eref;
        double ecan = fdiv(rhn, 100) * satvap(tn);
        double ea = (gt2 * eref + gtc * ecan) / (gt2 + gtc);          // (gt2 may be +-Inf when gtt == gtc: IEEE division kept)
        tz = (gt2 * tair + gtc * tn) / (gt2 + gtc);
        rh = fdiv(ea, satvap(tz)) * 100;
      }
      if (rh > 100) rh = 100;
      Rsw = swrad;
    } else {
      double gtc = gcan_apply(l_m, iwmean, ph, uh, a, gc_c_den);
      gtt = hpar(gtt, gtc);
      double gt0 = gcan_apply(l_m, iwmean, ph, uh, a, gc_0_den);
      double gha = gha_hour(gd, gid, gd3a, uz, hp);
      // second solar position (default meridian; the first one carries quirk Q15): the same numbers whenever the two meridians
      // coincide (clump = 0 and |long| < 7.5 degrees, e.g.), then reused instead of recomputed
      double sa = sa_l, K = K_l;
      if (!same_merid) {
        sa = solsin_pre(cr[8], lon, merid_def, 0, sd, slat, clat);
        K = (sa > 0) ? cank_sh(x, sa, kden) : 0.0;
      }
      double PAR = cansw_k(swrad, dp, sa, PAIu, s02, trd02, K, 0.0);
      double gC = layercond(PAR, gsmax, q50);
      double gv = hpar(gC, gha);
      double gL = hpar(gha, ph0 * uz);
      double Rabs = leafabs2(swrad, sa_l, K_l, mul, tair, PAIt, PAIu, ref, sref, trd_u, trd_t, refg, emv, trlw_u, skyem, dp, hp.p4);
      tleafS(tair, tground, relhum, pk, theta, gtt, gt0, gha, gv, gL, Rabs, vegem, soilb, psi_e, Smax, A.surfwet, leafdens, hp, tleaf, tz, rh);
      Rsw = cansw_k(swrad, dp, sa, PAIu, sls, trdls, K, 0.0);
      double Rdif = cdif * swrad * dp;
      dp = Rdif / Rsw;
      if (dp != dp) dp = 1;
      if (isinf(dp)) dp = 1;
      if (dp < 0) dp = 0;
      if (dp > 1) dp = 1;
      double tr = clumptr(fexp(-slw * PAIu), clump);
      double lwout = kSb * hp.p4;
      Rlw = tr * skyem * lwout + (1 - tr) * emv * lwout;
    }
    o[0] = tz; o[1] = tleaf; o[2] = rh; o[3] = Rsw; o[4] = Rlw; o[5] = uz; o[6] = dp;
  }

  // .snowtempf (NMR.R:495-512); false when the SN column lookup is out of range
  MC_HD static bool snowtempf(double snowdepth, const double* SN, double reqhgt, double& stemp) {
    const double nd[9] = {3.0, 2.0, 1.0, 0.5, 0.2, 0.1, 0.05, 0.025, 0.0};
    int selb = 0;
    for (int i = 1; i <= 9; ++i) if (nd[i - 1] - reqhgt <= 0) { selb = i; break; }
    int sela = selb - 1;
    if (sela < 1 || selb > 8) return false;
    double smm = fabs(reqhgt - nd[sela - 1]) + fabs(reqhgt - nd[selb - 1]);
    double wgt1 = 1 - fabs(reqhgt - nd[sela - 1]) / smm;
    double wgt2 = 1 - fabs(reqhgt - nd[selb - 1]) / smm;
    if (nd[sela - 1] > (snowdepth / 100)) { wgt1 = 0; wgt2 = 1; }
    stemp = wgt1 * SN[sela - 1] + wgt2 * SN[selb - 1];
    if (snowdepth < reqhgt) stemp = 0;
    return true;
  }
  // one hour with snow on the ground (.runmodelsnow body, NMR.R:514-677)
  MC_HD bool point_snow(const SteadyArgs& A, const double* cr, const HourPre& hp, double snowdep_cm, const double* SN, double* o) const {
    const double reqhgt = A.reqhgt;
    const double snowtemp = SN[0];
    const double snowdep = snowdep_cm / 100;
    const double tair = cr[0], relhum = cr[1], pk = cr[2], swrad = cr[4], skyem = cr[6];
    const double u2 = wind2m(A, cr[3], PAIt_s);
    const bool above = hgt > snowdep;
    const double zm = above ? zm3_s : 0.003;
    const double dd = above ? d_s : snowdep - 0.003;
    double hgt2 = above ? hgt : snowdep;
    double uz;
    if (above) uz = wind_apply_r(u2, iwA1, wB1, wE1);
    else {
      double l = flog(fdiv(snowdep + 2 - 0.0, 0.004)); if (l < 0.55) l = 0.55;
      uz = wind_apply(u2, l, wB2, -1.0);
    }
    double uf = fdiv(0.4 * u2, flog(fdiv(hgt2 + 2 - dd, zm)));
    if (uf < 0.065) uf = 0.065;
    hgt2 = hgt < snowdep ? snowdep : hgt;
    double uh = fdiv(uf, 0.4) * flog(fdiv(hgt2 - dd, zm));
    if (uh < uf) uh = uf;
    const bool below = reqhgt < snowdep;
    double tz2 = 0;
    if (below && !snowtempf(snowdep, SN, reqhgt, tz2)) return false;
    double dp = hp.dp;
    const SolDay& sd = hp.sd;
    const double ph = hp.ph, ph0 = hp.ph0;
    double gtt = gturb_apply(u2, gts_lnu, gts_lnh, ph, 0, 0);
    const double sa_l = solsin_pre(cr[8], lon, clump, 0, sd, slat, clat);
    double K_l, mul;
    cank2_sh(x, sa_l, kden, k5, K_l, mul);       // sa_l, sa: sin(solar altitude), see mc_phys.cuh
    const double emv85 = 1 - (1 - 0.85);
    double tz, tleaf, rh, Rsw, Rlw;
    if (reqhgt >= hgt) {
      double gt0 = gcan_apply(l_m_s, iwmean, ph, uh, a1, gcs_top_den);
      double gha = gha_hour(gd_s, gid_s, gd3a_s, uh, hp);
      double gC = layercond(swrad, gsmax, q50);
      double gv = hpar(gC, gha);
      double gL = hpar(gha, ph0 * uh);
      double Rabs = leafabs2(swrad, sa_l, K_l, mul, tair, PAIt_s, 0.0, ref_s, sref_s, 1.0, trds_t, 0.95, emv85, 1.0, skyem, dp, hp.p4);
      double tn, rhn;
      tleafS(tair, snowtemp, relhum, pk, 0.5, gtt, gt0, gha, gv, gL, Rabs, 0.85, soilb, psi_e, Smax, 1, leafdens_s, hp, tleaf, tn, rhn);
      if (reqhgt == hgt) { tz = tn; rh = rhn; }
      else {
        double gtc = gturb_apply(u2, gts_lnu, gtcs_lnh, ph, 0, 0);
        double gt2 = 1 / (1 / gtt - 1 / gtc);
        double eref = hp.eref;
        double ecan = fdiv(rhn, 100) * satvap(tn);
        double ea = (gt2 * eref + gtc * ecan) / (gt2 + gtc);          // (gt2 may be +-Inf when gtt == gtc: IEEE division kept)
        tz = (gt2 * tair + gtc * tn) / (gt2 + gtc);
        rh = fdiv(ea, satvap(tz)) * 100;
      }
      if (rh > 100) rh = 100;
      Rsw = swrad;
      Rlw = (1 - skyem) * 0.85 * kSb * pow4(tz + 273.15);
    } else {
      double gtc = gcan_apply(l_m_s, iwmean, ph, uh, a1, gcs_c_den);
      gtt = hpar(gtt, gtc);
      double gt0 = gcan_apply(l_m_s, iwmean, ph, uh, a1, gcs_0_den);
      double gha = gha_hour(gd_s, gid_s, gd3a_s, uz, hp);
      // second solar position (default meridian; the first one carries quirk Q15): the same numbers whenever the two meridians
      // coincide (clump = 0 and |long| < 7.5 degrees, e.g.), then reused instead of recomputed
      double sa = sa_l, K = K_l;
      if (!same_merid) {
        sa = solsin_pre(cr[8], lon, merid_def, 0, sd, slat, clat);
        K = (sa > 0) ? cank_sh(x, sa, kden) : 0.0;
      }
      double PAR = cansw_k(swrad, dp, sa, PAIu, s06, trd06, K, 0.0);
      double gC = layercond(PAR, gsmax, q50);
      double gv = hpar(gC, gha);
      double gL = hpar(gha, ph0 * uz);
      double trlw = clumptr(fexp(-fsqrt(emv85) * PAIu), 0.0);
      double Rabs = leafabs2(swrad, sa_l, K_l, mul, tair, PAIt_s, PAIu, ref_s, sref_s, trds_u, trds_t, 0.95, emv85, trlw, skyem, dp, hp.p4);
      tleafS(tair, snowtemp, relhum, pk, 0.5, gtt, gt0, gha, gv, gL, Rabs, vegem, soilb, psi_e, Smax, 1, leafdens_s, hp, tleaf, tz, rh);
      Rsw = cansw_k(swrad, dp, sa, PAIu, s95, trd95, K, 0.0);
      double Rdif = cdif95 * swrad * dp;
      dp = Rdif / Rsw;
      if (dp != dp) dp = 1;
      if (isinf(dp)) dp = 1;
      if (dp < 0) dp = 0;
      if (dp > 1) dp = 1;
      double lwout = kSb * hp.p4;
      Rlw = trlw85_u * skyem * lwout + (1 - trlw85_u) * (1 - 0.85) * lwout;     // canlw(tair, PAIu, 0.85, ...): 0.85 passed as `ref`
    }
    if (below) { tz = tz2; tleaf = tz2; rh = 100; uz = 0; Rsw = 0; Rlw = kSb * 0.85 * pow4(tz2 + 273.15); }
    o[0] = tz; o[1] = tleaf; o[2] = rh; o[3] = Rsw; o[4] = Rlw; o[5] = uz; o[6] = dp;
    return true;
  }
};

template <int MM>
MC_HD void steady_column(const SteadyArgs& A, long long c, int chunk, int nchunk, bool live = true) {
  SteadyCol<MM> col;
  col.load(A, c);
  if (!A.season && !A.vegt) col.prep(A, 1.0);
  const long long per = (A.T + nchunk - 1) / nchunk;
  const long long tb = (long long)chunk * per;
  const long long te = (tb + per < A.T) ? tb + per : A.T;
  double st[6] = {0, 1e300, -1e300, 0, 1e300, -1e300};
  int flag = 0;
  for (long long t = tb; t < te; ++t) {
#if defined(__CUDA_ARCH__) && MC_STEADY_SYNC
    // The hour's code (~25 KB of SASS) is larger than what 20 resident warps at 20 different places leave each other in the
    // instruction cache: the block's warps start every hour together, so one fill serves four warps (padding lanes loop too).
    // +4.5 % on config 2, +3 % on config 4; a second rendezvous in the middle of the hour costs 2.5 % instead.
    __syncthreads();
#endif
    if (A.vegt) {
      const double* row = A.vegt + t * (2 * A.m + 1) * A.nvt + (A.nvt == 1 ? 0 : c);
      for (int i = 1; i <= A.m; ++i) { col.PAIb[i] = row[(long long)(i - 1) * A.nvt]; col.pLAIb[i] = row[(long long)(A.m + i - 1) * A.nvt]; }
      col.prep(A, 1.0);
    } else if (A.season) col.prep(A, A.season[t]);
    const double* cr = A.clim + t * 9;
    double nm[11];
    if (A.nmr_percol) { for (int q = 0; q < 11; ++q) nm[q] = A.nmr[(t * 11 + q) * A.ncol + c]; }
    else { for (int q = 0; q < 11; ++q) nm[q] = A.nmr[t * 11 + q]; }
    HourPre hp;
    if (A.sol) {
      const double* h = A.sol + t * 13;
      hp.sd.eot = h[0]; hp.sd.sdec = h[1]; hp.sd.cdec = h[2]; hp.dp = h[3]; hp.ph = h[4]; hp.ph0 = h[5];
      hp.es = h[6]; hp.eref = h[7]; hp.cp = h[8]; hp.p4 = h[9]; hp.gS = h[10]; hp.gT = h[11]; hp.lpk = h[12];
    } else hp = hour_pre(cr);
    double o[7];
    if (nm[2] > 0) {                                             // NMR.R:873-883: snow rows come from .runmodelsnow
      if (!col.point_snow(A, cr, hp, nm[2], nm + 3, o)) { flag |= 4; for (int q = 0; q < 7; ++q) o[q] = NAN; }
    } else col.point(A, cr, hp, nm[0], nm[1], o);
    if (A.out && live) for (int q = 0; q < 7; ++q) A.out[(t * 7 + q) * A.ncol + c] = o[q];
    st[0] += o[0]; if (o[0] < st[1]) st[1] = o[0]; if (o[0] > st[2]) st[2] = o[0];
    st[3] += o[1]; if (o[1] < st[4]) st[4] = o[1]; if (o[1] > st[5]) st[5] = o[1];
  }
  if (!live) return;
  if (A.part) {
    double* p = A.part + (long long)chunk * 6 * A.ncol;
    for (int q = 0; q < 6; ++q) p[q * A.ncol + c] = st[q];
  }
#if defined(__CUDA_ARCH__)
  if (flag && A.flags) atomicOr(&A.flags[c], flag);
#else
  if (flag && A.flags) A.flags[c] |= flag;
#endif
}

}  // namespace mc
