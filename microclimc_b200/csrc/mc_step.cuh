// One column of the transient model, owned by one GPU thread.
//
// `Column<MM, MS>` holds the column's parameters, recurrent state and hoisted (time-invariant) geometry in
// thread-private arrays and advances it one time step with `step()` — the whole of the reference's
// `runonestep` (R/code.R:687-902: wind/conductance recurrence 747-770, layermerge 772, leafabs 776-778,
// leaftemp 785 [code.R:378-512], soil forcing 789-811, Thomas/ThomasV branches 813-867 [code.R:94-169],
// clamps and fluxes 868-897) fused into a single pass.  Differences from a literal translation:
//   * every quantity that depends only on (column parameters, node heights) — zero-plane displacement,
//     roughness, mixing lengths, attenuation-coefficient bases, the reference-grass edge profile, diffuse
//     and long-wave transmissions — is computed once per launch in `precompute()`;
//   * the per-layer vector expressions of leaftemp/leafabs are fused into one loop per layer with the
//     intermediates in registers;
//   * the Thomas solve keeps only the two sweep arrays.
// All arrays are 1-based (index 0 unused) so the code can be checked against the R line by line.
// The code is plain C++ (MC_HD) so tests can also build it for the host to debug parity without a GPU;
// the product library only ever instantiates it inside CUDA kernels.
#pragma once
#include "mc_phys.cuh"

namespace mc {

struct RunCfg {          // runonestep / runmodel scalar arguments (code.R:687-689, 1125-1129)
  double timestep, edgedist, reqhgt, zu, merid, dst, n, windhgt, zlafact, surfwet;
  int metopen;
  int dbg;                 // timing ablations of the streaming kernel (MC_DBG_SKIP; results are meaningless when non-zero)
};
// thz: soil moisture by soil node of this step (theta[, i] of runmodel's theta matrix, code.R:1146-1154, 1202) or null; theta == thz[0] then
struct Forcing { double tair, relhum, pk, u, tsoil, skyem, Rsw, dp, jd, lt, theta, thetap; const double* thz = nullptr; };

// state / vegp / soilp row indices of the column-interleaved arrays (include/microclimc_b200.h)
enum { V_HGT, V_X, V_LW, V_CD, V_HGTG, V_ZM0, V_REFLS, V_REFG, V_REFW, V_REFLP, V_VEGEM, V_GSMAX, V_Q50, V_CPW, V_PHW,
       V_KWOOD, V_CLUMP, V_NSCAL };
enum { S_SMAX, S_SMIN, S_ALPHA, S_N, S_KSAT, S_VQ, S_VM, S_VO, S_MC, S_RHO, S_B, S_PSIE, S_NSCAL };
enum { P_TABOVE, P_ZABOVE, P_RELHUM, P_TAIR, P_TSOIL, P_PK, P_H, P_G, P_PSIM, P_NSCAL };
MC_HD int nstate_rows(int m, int sm) { return P_NSCAL + 11 * m + (m + 1) + 2 * sm; }

template <int MM, int MS>
struct Column {
  static constexpr int MN = MM + MS + 3;
  int m, sm;
  // ---- parameters (vegp, soilp) ----
  double hgt, x, lw, cd, hgtg, zm0, refls, refg, refw, vegem, gsmax, q50, cpw, phw, clump;
  double PAI[MM + 2], PAIbase[MM + 2], pLAI[MM + 2], iw[MM + 2], thickw[MM + 2];
  double Smax, Vq, Vm, Vo_s, Mc, rho, soilb, psi_e, zs[MS + 2];
  double lat, lon;
  // ---- recurrent state (previn) ----
  double tc[MM + 2], tleaf[MM + 2], uz[MM + 2], zprev[MM + 2], rh[MM + 2], Rabs[MM + 2], gv[MM + 2], gha[MM + 2],
      gt[MM + 3], soiltc[MS + 2];
  double tabove, relhum, tair, tsoil, pk, H, psi_m;
  // ---- output-only fields of the state ----
  double L[MM + 2], Rswin[MM + 2], Rlwin[MM + 2], G, zabove;
  // ---- hoisted geometry ----
  double z[MM + 2];
  double zabove_g, sumPAI, d, zm, lnref, lzu, lzuh, lzoh, lnhd, La, lm_top, a0_top, ex_top, gl_top, ed_top, gx0_top, gx1_top;
  double lm_cap, a0_cap, gxc0, lm_g, a0_g, zg, gxg0, gxg1, iwmean;
  int selG;
  double a0i[MM + 2], a0o[MM + 2], lmi[MM + 2], gl_o[MM + 2], gl_i[MM + 2], ed_o[MM + 2], ed_i[MM + 2], ex1[MM + 2],
      ex2[MM + 2], gx0[MM + 2], gx1[MM + 2], zi_[MM + 2];
  double PAIg[MM + 2], sref[MM + 2], trdc[MM + 2], trdg[MM + 2], trlw[MM + 2];
  double mref, s_m, trd_m, kden, emv, trlw_sum;
  double zla[MM + 2], zthp[MM + 2];
  bool zprev_is_z;
  // true only when every thread of the block runs step() together (general kernel): block-wide rendezvous per layer keeps the
  // warps on the same instructions (I-cache). Must stay false when single threads call step() (streaming kernel's fallback).
  bool cta_sync = false;
  // soil conductance cache (soilk depends on theta only)
  double th_cache, ksoil[MS + 2], cdsoil[MS + 2];
  // grass-reference constants of windcanopy's edge profile (code.R:286)
  double g_d, g_zm, g_lnh;

  // ------------------------------------------------------------------------------------------
  MC_HD void load_params(const double* __restrict__ vegp, const double* __restrict__ soilp, const double* __restrict__ latp,
                         const double* __restrict__ lonp, long long ncol, long long c, int m_, int sm_) {
    m = m_; sm = sm_;
    auto V = [&](int r) { return vegp[(long long)r * ncol + c]; };
    hgt = V(V_HGT); x = V(V_X); lw = V(V_LW); cd = V(V_CD); hgtg = V(V_HGTG); zm0 = V(V_ZM0); refls = V(V_REFLS);
    refg = V(V_REFG); refw = V(V_REFW); vegem = V(V_VEGEM); gsmax = V(V_GSMAX); q50 = V(V_Q50); cpw = V(V_CPW);
    phw = V(V_PHW); clump = V(V_CLUMP);
    for (int i = 1; i <= m; ++i) {
      PAIbase[i] = V(V_NSCAL + i - 1); pLAI[i] = V(V_NSCAL + m + i - 1); iw[i] = V(V_NSCAL + 2 * m + i - 1);
      thickw[i] = V(V_NSCAL + 3 * m + i - 1);
    }
    auto S = [&](int r) { return soilp[(long long)r * ncol + c]; };
    Smax = S(S_SMAX); Vq = S(S_VQ); Vm = S(S_VM); Vo_s = S(S_VO); Mc = S(S_MC); rho = S(S_RHO); soilb = S(S_B); psi_e = S(S_PSIE);
    for (int i = 1; i <= sm; ++i) zs[i] = S(S_NSCAL + i - 1);
    lat = latp[c]; lon = lonp[c];
    th_cache = -1e300;
  }
  MC_NI void load_state(const double* __restrict__ st, long long ncol, long long c) {
    auto P = [&](int r) { return st[(long long)r * ncol + c]; };
    tabove = P(P_TABOVE); relhum = P(P_RELHUM); tair = P(P_TAIR); tsoil = P(P_TSOIL); pk = P(P_PK); H = P(P_H); psi_m = P(P_PSIM);
    zabove = P(P_ZABOVE); G = P(P_G);
    int r = P_NSCAL;
    for (int i = 1; i <= m; ++i) tc[i] = P(r + i - 1);
    r += m; for (int i = 1; i <= m; ++i) tleaf[i] = P(r + i - 1);
    r += m; for (int i = 1; i <= m; ++i) uz[i] = P(r + i - 1);
    r += m; for (int i = 1; i <= m; ++i) zprev[i] = P(r + i - 1);
    r += m; for (int i = 1; i <= m; ++i) rh[i] = P(r + i - 1);
    r += m; for (int i = 1; i <= m; ++i) Rabs[i] = P(r + i - 1);
    r += m; for (int i = 1; i <= m; ++i) gv[i] = P(r + i - 1);
    r += m; for (int i = 1; i <= m; ++i) gha[i] = P(r + i - 1);
    r += m; for (int i = 1; i <= m; ++i) L[i] = P(r + i - 1);
    r += m; for (int i = 1; i <= m; ++i) Rswin[i] = P(r + i - 1);
    r += m; for (int i = 1; i <= m; ++i) Rlwin[i] = P(r + i - 1);
    r += m; for (int i = 1; i <= m + 1; ++i) gt[i] = P(r + i - 1);
    r += m + 1; for (int i = 1; i <= sm; ++i) soiltc[i] = P(r + i - 1);
  }
  MC_NI void store_state(double* __restrict__ st, long long ncol, long long c) const {
    auto W = [&](int r, double v) { st[(long long)r * ncol + c] = v; };
    W(P_TABOVE, tabove); W(P_ZABOVE, zabove); W(P_RELHUM, relhum); W(P_TAIR, tair); W(P_TSOIL, tsoil); W(P_PK, pk);
    W(P_H, H); W(P_G, G); W(P_PSIM, psi_m);
    int r = P_NSCAL;
    for (int i = 1; i <= m; ++i) W(r + i - 1, tc[i]);
    r += m; for (int i = 1; i <= m; ++i) W(r + i - 1, tleaf[i]);
    r += m; for (int i = 1; i <= m; ++i) W(r + i - 1, uz[i]);
    r += m; for (int i = 1; i <= m; ++i) W(r + i - 1, zprev[i]);
    r += m; for (int i = 1; i <= m; ++i) W(r + i - 1, rh[i]);
    r += m; for (int i = 1; i <= m; ++i) W(r + i - 1, Rabs[i]);
    r += m; for (int i = 1; i <= m; ++i) W(r + i - 1, gv[i]);
    r += m; for (int i = 1; i <= m; ++i) W(r + i - 1, gha[i]);
    r += m; for (int i = 1; i <= m; ++i) W(r + i - 1, L[i]);
    r += m; for (int i = 1; i <= m; ++i) W(r + i - 1, Rswin[i]);
    r += m; for (int i = 1; i <= m; ++i) W(r + i - 1, Rlwin[i]);
    r += m; for (int i = 1; i <= m + 1; ++i) W(r + i - 1, gt[i]);
    r += m + 1; for (int i = 1; i <= sm; ++i) W(r + i - 1, soiltc[i]);
    r += sm; for (int i = 1; i <= sm; ++i) W(r + i - 1, zs[i]);         // sz <- soilp$z (code.R:791)
  }

  // paraminit (code.R:566-585): two-knot spline == linear interpolation at n equally spaced points
  MC_HD static double lin2(double y1, double y2, int i, int n) {
    double xx = 1 + (double)(i - 1) / (double)(n - 1);
    return y1 + (y2 - y1) * (xx - 1);
  }
  MC_HD void paraminit(double hgt_, double tair_, double u_, double relhum_, double tsoil_, double Rsw_) {
    for (int i = 1; i <= m; ++i) tc[i] = lin2(tsoil_, tair_, sm + i, m + sm);
    for (int i = 1; i <= sm; ++i) soiltc[i] = lin2(tsoil_, tair_, sm + 1 - i, m + sm);
    for (int i = 1; i <= m; ++i) {
      rh[i] = relhum_;
      Rabs[i] = lin2(0.2 * Rsw_ + 20, Rsw_ + 20, i, m);
      tleaf[i] = tc[i] + 0.01 * Rabs[i];
      gv[i] = lin2(0.25, 0.32, i, m);
      gha[i] = lin2(0.13, 0.19, i, m);
      zprev[i] = (i - 0.5) / m * hgt_;
      uz[i] = ((double)i / m) * u_;
      L[i] = 0; Rswin[i] = Rabs[i]; Rlwin[i] = 0.2 * Rabs[i];
    }
    double g0 = 50.0 * ((double)m / hgt_) * (2.0 / m);
    for (int i = 1; i <= m + 1; ++i) gt[i] = g0;
    gt[1] = 2 * gt[1];
    gt[m + 1] = gt[m + 1] * (hgt_ / m) * 0.5;
    tabove = tair_; zabove = hgt_; relhum = relhum_; tair = tair_; tsoil = tsoil_; pk = 101.3; H = 0; G = 0; psi_m = 0;
  }
  // sz of paraminit (code.R:579) is only reported, never read back; written by the caller when needed.

  // reference-grass edge profile factor of windcanopy (code.R:286-289) for node height zq below a layer
  // top `htop`:  ur = (uf/0.4) * gl ;  edr - 1 = ed.   Returns false when zq < 0.12 (generic path needed).
  MC_HD bool grass_consts(double zq, double htop, double edgedist, double& gl, double& ed) const {
    double uwr = 2.5 * log((zq - 0.08) / 0.01476);
    if (uwr < 1) uwr = 1;
    if (uwr != uwr) uwr = 1;
    double edr = (htop - edgedist / uwr) / htop;
    ed = edr - 1;
    double ln2 = log((zq - g_d) / g_zm);
    if (ln2 < 0.55) ln2 = 0.55;
    gl = ln2;
    return !(zq < 0.12);
  }
  // generic windprofile (code.R:191-221) for one height
  MC_NI static double windprofile(double ui, double zi, double zo, double a, double PAI_, double hgt_, double psi, double hgtg_,
                                  double zm0_) {
    double dd = zeroplanedis(hgt_, PAI_);
    double zmm = roughlength_d(hgt_, PAI_, dd, zm0_);
    double zmg = 0.1 * hgtg_;
    double ln1 = log((zi - dd) / zmm) + psi;
    double uf = 0.4 * (ui / ln1);
    double ln2 = log((zo - dd) / zmm) + psi;
    if (ln2 < 0.55) ln2 = 0.55;
    double uo = (uf / 0.4) * ln2;
    if (zo < hgt_) { ln2 = log((hgt_ - dd) / zmm) + psi; uo = (uf / 0.4) * ln2; }
    if (zo > (0.1 * hgt_) && zo < hgt_) uo = uo * exp(a * ((zo / hgt_) - 1));
    if (zo <= (0.1 * hgt_)) {
      uo = uo * exp(a * (((0.1 * hgt_) / hgt_) - 1));
      double ln1b = log((0.1 * hgt_) / zmg) + psi;
      double ufb = 0.4 * (uo / ln1b);
      double ln2b = log(zo / zmg) + psi;
      uo = (ufb / 0.4) * ln2b;
    }
    return uo;
  }
  // windcanopy (code.R:279-294) with the hoisted pieces: a = attenuation coefficient, ex = z/hgt - 1,
  // (gl, ed, fast) from grass_consts, ugr = uf_grass/0.4
  MC_HD double windcanopy_h(double uh, double a, double ex, double gl, double ed, bool fast, double zq, double uref,
                            double ufg, const RunCfg& cfg) const {
    double uzv = uh * exp(a * ex);
    if (cfg.edgedist == cfg.edgedist) {            // !is.na(edgedist)
      double ur = fast ? (ufg / 0.4) * gl : windprofile(uref, cfg.zu, zq, 0.7989537, 1.146, 0.12, 0, 0.05 * 0.12, 0.004);
      double uhr = ur * exp(a * ed);
      if (uzv != uzv) uzv = uhr;
      else if (uhr == uhr && uhr > uzv) uzv = uhr;
    }
    return uzv;
  }

  // ---- time-invariant geometry (recomputed each step only when PAI varies in time) -------------
  MC_NI void precompute(const RunCfg& cfg, double season) {
    for (int i = 1; i <= m; ++i) { double p = PAIbase[i] * season; PAI[i] = (p == 0) ? 0.0001 : p; }   // code.R:691-696
    const double reqhgt = cfg.reqhgt;
    for (int i = 1; i <= m; ++i) z[i] = (i - 0.5) / m * hgt;
    if (reqhgt == reqhgt) {                                                                            // code.R:719
      double mn = 1e300; int sel = 1;
      for (int i = 1; i <= m; ++i) { double dd = fabs(z[i] - reqhgt); if (dd < mn) { mn = dd; sel = i; } }
      z[sel] = reqhgt;
    }
    zabove_g = hgt;
    if (reqhgt == reqhgt && reqhgt > hgt) zabove_g = reqhgt;
    sumPAI = 0; for (int i = 1; i <= m; ++i) sumPAI += PAI[i];
    { double s = 0; for (int i = 1; i <= m; ++i) s += iw[i]; iwmean = s / m; }
    d = zeroplanedis(hgt, sumPAI);
    zm = roughlength_d(hgt, sumPAI, d, zm0);
    double zh = 0.2 * zm;
    lnref = log(((hgt + 2) - d) / zm);
    lzu = log((cfg.zu + hgt - d) / zm);
    lzuh = log((cfg.zu + hgt - d) / zh);
    lzoh = log((zabove_g - d) / zh);
    lnhd = log((hgt - d) / zm);
    La = log(67.8 * (hgt + 2) - 5.42);
    // grass reference (hgt 0.12, PAI 1.146, zm0 0.004)
    g_d = zeroplanedis(0.12, 1.146);
    g_zm = roughlength_d(0.12, 1.146, g_d, 0.004);
    g_lnh = log((cfg.zu - g_d) / g_zm);
    // top layer
    lm_top = mixinglength(hgt, sumPAI);
    a0_top = attencoef_base(hgt, sumPAI, cd, iw[m], lm_top);
    ex_top = (z[m] / hgt) - 1;
    zi_[m] = grass_consts(z[m], hgt, cfg.edgedist, gl_top, ed_top) ? 1.0 : 0.0;
    gx0_top = z[m - 1] / hgt - 1;
    gx1_top = z[m] / hgt - 1;
    // layers m-1 .. 1
    double cs = 0;
    for (int i = 1; i <= m; ++i) { cs += PAI[i]; PAIg[i] = cs; }
    for (int i = m - 1; i >= 1; --i) {
      double zt = z[i + 1] - z[i];
      double zi = z[i + 1] + 0.5 * zt, zo = z[i + 1] - 0.5 * zt;
      double PAIi = PAIg[i];
      lmi[i] = mixinglength(zi, PAIi);
      a0i[i] = attencoef_base(zi, PAIi, cd, iw[i], lmi[i]);
      double lmo = mixinglength(zo, PAIi);
      a0o[i] = attencoef_base(zo, PAIi, cd, iw[i], lmo);
      ex1[i] = (zo / zi) - 1;
      ex2[i] = (zi / zo) - 1;
      bool f1 = grass_consts(zo, zi, cfg.edgedist, gl_o[i], ed_o[i]);
      bool f2 = grass_consts(zi, zo, cfg.edgedist, gl_i[i], ed_i[i]);
      zi_[i] = (f1 && f2) ? 1.0 : 0.0;
      double z0 = (i > 1) ? z[i - 1] : 0;
      gx0[i] = z0 / zi - 1;
      gx1[i] = z[i] / zi - 1;
    }
    // conductance canopy top <-> reference (code.R:769) and ground flux conductance (code.R:885)
    lm_cap = mixinglength(hgt, PAI[m] * 0.5);
    a0_cap = attencoef_base(hgt, PAI[m] * 0.5, cd, iw[m], lm_cap);
    gxc0 = z[m] / hgt - 1;
    lm_g = lm_top;
    a0_g = attencoef_base(hgt, sumPAI, cd, iwmean, lm_g);
    zg = d + 0.2 * zm;
    gxg0 = 0 / hgt - 1;
    gxg1 = zg / hgt - 1;
    { double mn = 1e300; selG = 1; for (int i = 1; i <= m; ++i) { double zd = fabs(z[i] - d + 0.2 * zm); if (zd < mn) { mn = zd; selG = i; } } }
    // radiation (leafabs, code.R:31-49)
    kden = cank_den(x);
    emv = 1 - (1 - vegem);
    double slw = sqrt(emv);
    double rs = 0;
    for (int i = 1; i <= m; ++i) {
      double ref = pLAI[i] * refls + (1 - pLAI[i]) * refw;
      rs += ref;
      sref[i] = sqrt(1 - ref);
    }
    mref = rs / m;
    s_m = sqrt(1 - mref);
    trd_m = clumptr(exp(-s_m * sumPAI), clump);
    for (int i = 1; i <= m; ++i) {
      double pc = PAIg[m + 1 - i];                    // rev(cumsum(PAI))
      trdc[i] = clumptr(exp(-sref[i] * pc), clump);
      trdg[i] = clumptr(exp(-sref[i] * PAIg[i]), clump);
      trlw[i] = clumptr(exp(-slw * pc), clump);
    }
    trlw_sum = clumptr(exp(-slw * sumPAI), 0.0);
  }
  // leaftemp's layer geometry depends on previn$z (code.R:387-392)
  MC_NI void precompute_leafgeom(const RunCfg& cfg) {
    for (int i = 1; i <= m - 1; ++i) zthp[i] = zprev[i + 1] - zprev[i];
    zthp[m] = hgt - (zprev[m] + zprev[m - 1]) / 2;
    for (int i = 1; i <= m; ++i) zla[i] = mixinglength(zthp[i], PAI[i]) * 0.5 * cfg.zlafact;
    zprev_is_z = true;
    for (int i = 1; i <= m; ++i) if (zprev[i] != z[i]) zprev_is_z = false;
  }

  MC_NI void soilk(double timestep, double theta, const double* thz = nullptr) {
    if (!thz && theta == th_cache) return;
    th_cache = thz ? -1e300 : theta;
    double A = 0.65 - 0.78 * rho + 0.6 * rho * rho;
    double B = 1.06 * rho;
    double C = 1 + 2.6 / sqrt(Mc);
    double D = 0.03 + 0.1 * rho * rho;
    double ct = C * theta;
    double lam = A + B * theta - (A - D) * exp(-((ct * ct) * (ct * ct)));
    double Cv = (2.128 * Vq + 2.385 * Vm + 2.496 * Vo_s + 4.18 * theta) * 1e6;
    for (int i = 1; i <= sm; ++i) {
      if (thz) {                                   // moisture by depth: node i from its own theta
        ct = C * thz[i - 1];
        lam = A + B * thz[i - 1] - (A - D) * exp(-((ct * ct) * (ct * ct)));
        Cv = (2.128 * Vq + 2.385 * Vm + 2.496 * Vo_s + 4.18 * thz[i - 1]) * 1e6;
      }
      double dzk = (i < sm) ? (zs[i + 1] - zs[i]) : (zs[sm] - zs[sm - 1]);
      ksoil[i] = lam / dzk;
      double zlo = (i > 1) ? zs[i - 1] : 0.0;
      double zhi = (i < sm) ? zs[i + 1] : (zs[sm] + (zs[sm] - zs[sm - 1]));
      cdsoil[i] = (Cv * (zhi - zlo)) / (2 * timestep);
    }
  }

  // Thomas (code.R:94-125) on 1-based local arrays: tcv[N+2] (modified), k[N+1], cdv[N], Xv[N] -> tn[N+2].
  MC_NI static void thomas(int N, double* tcv, double tmsoil, double tairv, const double* k, const double* cdv, double f,
                           const double* Xv, double* tn, double* cc, double* dd) {
    const double g = 1 - f;
    tn[N + 2] = tmsoil; tn[1] = tairv;
    for (int xx = 2; xx <= N + 1; ++xx) tcv[xx] = tcv[xx] + g * Xv[xx - 1];
    for (int xx = 2; xx <= N + 1; ++xx)
      dd[xx] = g * k[xx - 1] * tcv[xx - 1] + (cdv[xx - 1] - g * (k[xx] + k[xx - 1])) * tcv[xx] + g * k[xx] * tcv[xx + 1];
    dd[2] = dd[2] + k[1] * tn[1] * f;
    dd[N + 1] = dd[N + 1] + k[N + 1] * f * tn[N + 2];
    double b = f * (k[2] + k[1]) + cdv[1];
    for (int i = 2; i <= N; ++i) {
      double ci = (-k[i] * f) / b;
      cc[i] = ci;
      dd[i] = dd[i] / b;
      double a1 = -k[i] * f;                               // a[i+1] <- cc[i] (before normalisation)
      b = (f * (k[i + 1] + k[i]) + cdv[i]) - a1 * ci;
      dd[i + 1] = dd[i + 1] - a1 * dd[i];
    }
    tn[N + 1] = dd[N + 1] / b;
    for (int i = N; i >= 2; --i) tn[i] = dd[i] - cc[i] * tn[i + 1];
    for (int xx = 2; xx <= N + 1; ++xx) {
      double xmn = dmin(dmin(tcv[xx], tcv[xx - 1]), tcv[xx + 1]);
      double xmx = dmax(dmax(tcv[xx], tcv[xx - 1]), tcv[xx + 1]);
      double v = tn[xx];
      if (v < xmn) v = xmn;
      if (v > xmx) v = xmx;
      tn[xx] = v + f * Xv[xx - 1];
    }
  }

  // ---- one time step (runonestep, code.R:687-902) ---------------------------------------------
  // returns the number of merged air layers used by the diffusion solve
  MC_NI int step(const Forcing& fc, const RunCfg& cfg) {
    const double timestep = cfg.timestep;
    const double tairn = fc.tair, pkn = fc.pk, skyem = fc.skyem, Rsw = fc.Rsw;
    double relhumn = fc.relhum, u = fc.u;
    const double ppk = pk, Hprev = H, ps1 = soiltc[1];
    const double pha = phair(tairn, pkn);
    // wind 2 m above canopy (code.R:709-716)
    double u2;
    if (cfg.metopen) {
      if (cfg.windhgt != 2) u = u * 4.87 / log(67.8 * cfg.windhgt - 5.42);
      u2 = u * La / log(67.8 * 2 - 5.42);
    } else {
      u2 = u;
      if (cfg.windhgt != (2 + hgt)) u2 = windprofile(u, cfg.windhgt, hgt + 2, 2, sumPAI, hgt, psi_m, 0.05 * hgt, 0.004);
    }
    if (u2 < 0.5) u2 = 0.5;
    // diabatic corrections (code.R:727-732)
    double uf = (0.4 * u2) / lnref;
    double psi_mn, psi_h;
    diabatic_cor(tairn, pkn, Hprev, uf, hgt + 2, d, psi_mn, psi_h);
    double phi_m;
    {
      double dz = z[m] - z[1];
      double dtdz = (tc[m] - tc[1]) / dz;
      double dudz = (uz[m] - uz[1]) / dz;
      if (fabs(dudz) < 0.01) dudz = 0.01;
      double tm = 0; for (int i = 1; i <= m; ++i) tm += tc[i];
      double Tk = tm / m + 273.15;
      double Ri = (9.81 / Tk) * dtdz / (dudz * dudz);
      if (Ri > 0.15) Ri = 0.15;
      if (Ri < -1) Ri = -1;
      if (Ri < 0) { double phh = 1 / sqrt(1 - 16 * Ri); phi_m = sqrt(phh); }
      else { double st = Ri / (1 - 5 * Ri); phi_m = 1 + (6 * st) / (1 + st); }
    }
    // canopy-top temperature and humidity (code.R:734-745; abovecanopytemp code.R:319-331)
    double tcan;
    {
      double ufh = (0.4 * u2) / (lzu + psi_h);
      double mm = Hprev / (0.4 * pha * cpair(tairn) * ufh);
      double tref = tairn + mm * (lzuh + psi_h);
      tcan = tref - mm * (lzoh + psi_h);
    }
    const double tfrost = dewpoint((relhumn / 100) * satvap(tairn), tairn);
    if (tcan < tfrost) tcan = tfrost;
    if (tcan > (tairn + 15)) tcan = tairn + 15;
    if (tcan < (tairn - 4.5)) tcan = tairn - 4.5;
    const double tet_tair = tetens(tairn);
    const double tet_tcan = tetens(tcan);
    {
      double ea = tet_tair * (relhumn / 100);
      relhumn = (ea / tet_tcan) * 100;
      if (relhumn > 100) relhumn = 100;
    }
    // ---- wind speed and turbulent conductances, top -> bottom (code.R:747-770) ----
    const double sqphi = sqrt(phi_m);
    const double ufg = 0.4 * (u / g_lnh);                     // grass-profile friction velocity (uref = u, zref = zu)
    double gtn[MM + 3], uzn[MM + 2];
    double uh;
    {
      double ln1 = lnref + psi_mn;
      double ufw = 0.4 * (u2 / ln1);
      double ln2 = lnhd + psi_mn;
      if (ln2 < 0.55) ln2 = 0.55;
      uh = (ufw / 0.4) * ln2;                                  // windprofile(u2, hgt+2, hgt, ...) (zo == hgt)
    }
    {
      double a = a0_top * sqphi;
      uzn[m] = windcanopy_h(uh, a, ex_top, gl_top, ed_top, zi_[m] != 0.0, z[m], u, ufg, cfg);
      double ph = phair((tc[m] + tc[m - 1]) / 2, pkn);
      double e0 = exp(-a * gx0_top), e1 = exp(-a * gx1_top);
      double g = (lm_top * iw[m] * ph * uh * a) / (hgt * (e0 - e1) * phi_m);
      double gfree = 0.0012 * ph * qroot(fabs(tc[m] - tc[m - 1]) / fabs(z[m] - z[m - 1]));
      gtn[m] = (g > gfree) ? g : gfree;
    }
    for (int i = m - 1; i >= 1; --i) {
      if (cta_sync) MC_CTA_SYNC();
      bool fast = zi_[i] != 0.0;
      double zt = z[i + 1] - z[i];
      double zi = z[i + 1] + 0.5 * zt, zo = z[i + 1] - 0.5 * zt;
      double ai = a0i[i] * sqphi, ao = a0o[i] * sqphi;
      uh = windcanopy_h(uh, ai, ex1[i], gl_o[i], ed_o[i], fast, zo, u, ufg, cfg);
      uzn[i] = windcanopy_h(uh, ao, ex2[i], gl_i[i], ed_i[i], fast, zi, u, ufg, cfg);          // Q9
      double tlo = (i > 1) ? tc[i - 1] : ps1;
      double z0 = (i > 1) ? z[i - 1] : 0;
      double ph = phair((tc[i] + tlo) / 2, pkn);
      double e0 = exp(-ai * gx0[i]), e1 = exp(-ai * gx1[i]);
      double g = (lmi[i] * iw[i] * ph * uh * ai) / (zi * (e0 - e1) * phi_m);
      double gfree = 0.0012 * ph * qroot(fabs(tc[i] - tlo) / fabs(z[i] - z0));
      gtn[i] = (g > gfree) ? g : gfree;
    }
    {                                                          // Q8: leftover uh, tc[1]
      double a = a0_cap * sqphi;
      double ph = phair((tcan + tc[1]) / 2, pkn);
      double e0 = exp(-a * gxc0), e1 = exp(-a * (hgt / hgt - 1));
      double g = (lm_cap * iw[m] * ph * uh * a) / (hgt * (e0 - e1) * phi_m);
      double gfree = 0.0012 * ph * qroot(fabs(tcan - tc[1]) / fabs(hgt - z[m]));
      gtn[m + 1] = (g > gfree) ? g : gfree;
    }
    // ---- layermerge (code.R:772; definition in DESIGN.md) ----
    int gend[MM + 2];          // last member of each merged group
    int ng = 0;
    {
      double R = 0, C = 0; bool open = false;
      for (int i = 1; i <= m; ++i) {
        double zth = z[i] - (i > 1 ? z[i - 1] : 0.0);
        R += 1 / gtn[i]; C += phair(tc[i], ppk) * zth; open = true;
        if (timestep / R <= C) { gend[++ng] = i; R = 0; C = 0; open = false; }
      }
      if (open) { if (ng > 0) gend[ng] = m; else { ng = 1; gend[1] = m; } }
    }
    // ---- absorbed radiation, stomatal and boundary-layer conductances (code.R:776-784) ----
    double dp = fc.dp;
    const double sa = solalt(fc.lt, lat, lon, fc.jd, cfg.merid, cfg.dst);
    if (dp != dp) dp = difprop(Rsw, sa);
    const double K = (sa > 0) ? cank(x, sa, kden) : 0.0;
    const double mul = cank(x, sa < 5 ? 5.0 : sa, kden);
    const double aG0 = refg * cansw_k(Rsw, dp, sa, sumPAI, s_m, trd_m, K, clump);
    const double lwout = kSb * pow4(tairn + 273.15);
    double aRsw1 = 0, ref_m = 0;
    double Rabsn[MM + 2], gvn[MM + 2], ghan[MM + 2];
    for (int i = 1; i <= m; ++i) {
      if (cta_sync) MC_CTA_SYNC();
      double ref = pLAI[i] * refls + (1 - pLAI[i]) * refw;
      double pc = PAIg[m + 1 - i];
      double sunl = psunlit_k(pc, sa, K, clump);
      double aR = (1 - ref) * cansw_k(Rsw, dp, sa, pc, sref[i], trdc[i], K, clump);
      double aGr = (1 - ref) * cansw_k(aG0, dp, sa, PAIg[i], sref[i], trdg[i], K, clump);
      double aRsw = sunl * mul * aR + (1 - sunl) * aR + aGr;
      double lwin = trlw[i] * skyem * lwout + (1 - trlw[i]) * emv * lwout;
      double aRlw = emv * lwin;
      if (i == 1) aRsw1 = aRsw;
      if (i == m) ref_m = ref;
      Rabsn[i] = aRsw + aRlw;
      double rsw = aRsw / (1 - ref);
      Rswin[i] = rsw;
      gvn[i] = layercond(rsw, gsmax, q50);
      ghan[i] = 1.41 * gforcedfree(lw * 0.71, uzn[i], tc[i], tleaf[i] - tc[i], pkn, 1.0);
    }
    // ---- leaftemp (code.R:378-512), fused per layer ----
    double gtt[MM + 2], gtt2[MM + 2];
    {
      double acc = 1 / (0.5 * gtn[m + 1] + 0.5 * gt[m + 1]);
      for (int i = m; i >= 1; --i) { gtt[i] = 1 / acc; acc = acc + 1 / (0.5 * gtn[i] + 0.5 * gt[i]); }
      double acc2 = 1 / (0.5 * gtn[1] + 0.5 * gt[1]);
      gtt2[1] = 1 / acc2;
      for (int i = 2; i <= m; ++i) { acc2 = acc2 + 1 / (0.5 * gtn[i] + 0.5 * gt[i]); gtt2[i] = 1 / acc2; }
    }
    const double lambda_l = (-42.575 * tcan + 44994) * cfg.surfwet;
    const double mtref = 0.5 * tcan + 0.5 * tair;
    const double mrh = 0.5 * relhumn + 0.5 * relhum;
    const double mpk = 0.5 * ppk + 0.5 * pkn;
    const double eref = (mrh / 100) * satvap(mtref);
    const double esoil1 = soilrh(fc.theta, soilb, -psi_e, Smax, ps1) * satvap(ps1);           // Q7
    // with moisture by depth esoil is a vector over the soil nodes that R recycles over the canopy layers (quirk Q24)
    double esoilv[MS + 2];
    if (fc.thz) { const double sv = satvap(ps1); for (int j = 1; j <= sm; ++j) esoilv[j] = soilrh(fc.thz[j - 1], soilb, -psi_e, Smax, ps1) * sv; }
    const double eamn = (relhumn / 100) * tet_tcan;
    double tnl[MM + 2], tleafn[MM + 2], eal[MM + 2];
    for (int i = 1; i <= m; ++i) {
      if (cta_sync) MC_CTA_SYNC();
      const double esoil = fc.thz ? esoilv[(i - 1) % sm + 1] : esoil1;
      const double ptc = tc[i], ptl = tleaf[i];
      const double zp = zprev[i], zth = zthp[i], zl = zla[i];
      const double zrf = (hgt + 2) - zp;
      const double esj = tetens(ptc);
      const double eaj = (rh[i] / 100) * esj;
      const double estl = satvap(ptl);
      const double delta = 4098 * tetens(ptl) / sq(ptl + 237.3);
      const double gvp = 1 / ((1 / gv[i]) + (1 / gha[i]));
      const double gvc = 1 / ((1 / gvn[i]) + (1 / ghan[i]));
      const double gvm = 0.5 * gvp + 0.5 * gvc;
      const double gham = 0.5 * ghan[i] + 0.5 * gha[i];
      const double PAIm = 2 * PAI[i] / zth;
      const double gv2b = (PAIm / zth) * gvm;                                                  // Q21
      const double g1 = gtt[i] / zrf, g2 = gtt2[i] / zp, g3 = gvm / zl;
      double test = dmax(dmax(timestep * gtt[i] / zrf, timestep * gvm / zl), timestep * gtt2[i] / zp);
      double ae, be;
      if (test > 1) {
        double den = gtt[i] + gtt2[i] + gv2b;
        ae = (gtt[i] * eref + gtt2[i] * esoil + gv2b * estl) / den;
        be = (0.5 * delta) / den;
      } else {
        double btm = (1 / timestep) + 0.5 * (g1 + g2 + g3);
        // edf(ea1, ea2, tc): max(ea1 - ea2, 0) unless ea2 > es(tc)  (code.R:380-386)
        double sat = eaj > esj ? 0.0 : 1.0;
        double e1 = (eref > eaj ? eref - eaj : 0.0) * sat;
        double e2 = (esoil > eaj ? esoil - eaj : 0.0) * sat;
        double e3 = (estl > eaj ? estl - eaj : 0.0) * sat;
        ae = eaj + 0.5 * (g1 * e1 + g2 * e2 + g3 * e3) / btm;
        be = (0.25 * gvm * delta) / btm;
      }
      test = dmax(dmax(timestep * gtt[i] / zrf, timestep * gham / zl), timestep * gtt2[i] / zp);
      const double ph = phair(ptc, ppk);
      const double cp = cpair(ptc);
      const double vden = thickw[i] * PAI[i];
      double aL, bL;
      if (test > 1) {
        double K1 = gtt[i] * cp, K2 = gtt2[i] * cp, K3 = gham * cp;
        aL = (K1 * mtref + K2 * ps1 + K3 * ptl) / (K1 + K2 + K3);
        bL = (0.5 * K3) / (K1 + K2 + K3);
      } else {
        double ma = (timestep * PAIm * 2) / (cp * ph * (1 - vden));
        double K1 = gtt[i] * cp / zrf, K2 = gtt2[i] * cp / zp, K3 = gham * cp * PAIm / zl;
        double btm = 1 + 0.5 * ma * (K1 + K2 + K3);
        aL = (ptc + 0.5 * ma * (K1 * mtref + K2 * ps1 + K3 * ptl)) / btm;
        bL = (0.25 * ma * K3) / btm;
      }
      if (bL < 0) bL = 0;
      const double aH = cp * gham * (ptl - aL);
      double bH = cp * gham * (0.5 - 0.5 * bL);
      if (bH < 0) bH = 0;
      const double lg = lambda_l * gvm / mpk;
      double edf_l = (estl > ae ? estl - ae : 0.0);
      if (ae > esj) edf_l = 0;
      const double aX = lg * edf_l;
      double bX = lg * (delta - be);
      if (bX < 0) bX = 0;
      const double Ra = 0.5 * Rabsn[i] + 0.5 * Rabs[i];
      const double tk = ptl + 273.15;
      const double aR = vegem * kSb * pow4(tk);
      const double bR = vegem * kSb * 2 * (tk * tk * tk);
      const double Ch = vden * cpw * phw;
      const double ml = (timestep * PAIm) / Ch;
      const double num = Ra - aR - aX - aH, bsum = bR + bX + bH;
      double dTL = (ml * num) / (zl + ml * bsum);
      const double dTL2 = num / bsum;
      if (fabs(dTL2) < fabs(dTL)) dTL = dTL2;
      double tn = aL + bL * dTL;
      double tlf = ptl + dTL;
      if (tlf > tn) { double tmx = dmax(tcan + 15, tlf); if (tn > tmx) tn = tmx; }
      if (tlf < tn) { double tmn = dmin(tcan - 5, tlf); if (tn < tmn) tn = tmn; }
      const double eam = ae + be * dTL;
      double ean = eaj + 2 * (eam - eaj);
      const double es = tetens(tn);
      if (ean > es) ean = es;
      if (ean < eamn) ean = eamn;
      const double al = log(ean / 0.6108);
      const double Tdew = -(237.3 * al) / (al - 17.27);
      double dT1 = -(ptl + dTL - Tdew);
      if (dT1 < 0) dT1 = 0;
      const double Lc = lambda_l * gvm * (tetens(tlf) - ean) / mpk;
      double dT2 = -(ml / zl) * Lc;
      if (dT2 < 0) dT2 = 0;
      tlf = tlf + dmin(dT1, dT2);
      dTL = tlf - ptl;
      const double dTmx = ptc + 20.2 - ptl, dTmn = ptc - 4.5 - ptl;
      if (dTL > dTmx) dTL = dTmx;
      if (dTL < dTmn) dTL = dTmn;
      tnl[i] = aL + bL * dTL;
      tleafn[i] = ptl + dTL;
      eal[i] = ean;
      L[i] = (aX + bX * dTL) * PAI[i];                                                         // code.R:880 (Q1)
    }
    // ---- soil conductances and surface forcing (code.R:789-811) ----
    if (cta_sync) MC_CTA_SYNC();
    soilk(timestep, fc.theta, fc.thz);
    const double esoil = esoil1;                                 // tln$esoil[1] (code.R:799-805)
    double TT[MM + 2];
    { double s = 0; for (int i = 1; i <= m; ++i) { s += (phair(tc[i], ppk) / gtn[i]) * (z[i] - (i > 1 ? z[i - 1] : 0.0)); TT[i] = s; } }
    double Xs;
    {
      double a1 = (1 - refg) * (aRsw1 / cdsoil[1]);
      if (esoil - eal[1] > 0) {
        double mm = ((-42.575 * tc[1] + 44994) * phair(tc[1], ppk) * z[1]) / pkn;
        double delta = 4098 * tetens(ps1) / sq(ps1 + 237.3);
        double btm = 1 + 0.5 * (mm / cdsoil[1]) * delta;
        Xs = (a1 - (mm / cdsoil[1]) * (esoil - eal[1])) / btm;
      } else Xs = a1;
      double sgn = Xs < 0 ? -1.0 : 1.0;
      double Xsmx = ((1 - refg) * aRsw1) / (ghan[1] * cpair(tc[1])) + (tnl[1] - ps1);
      if (fabs(Xs) > fabs(Xsmx)) Xs = Xsmx;
      Xs = fabs(Xs) * sgn;
    }
    double tnn[MM + 2], ean2[MM + 2], tnsoil[MS + 2];
    const double TX = TT[m] + (pha / gtn[m + 1]) * 2;
    if (ng > 1) {
      // ---- canopy air not in equilibrium: layeraverage -> Thomas -> ThomasV -> layerinterp (code.R:813-852) ----
      const int mg = ng, N = mg + sm;
      double ltc[MM + 2], lgt[MM + 3], lz[MM + 4], lVo[MM + 2], lea[MM + 2], lX[MM + 2], lvd[MM + 2], lcp[MM + 2], lph[MM + 2],
          lTT[MM + 2];
      int cprev = 0, lo = 1;
      lz[1] = 0;
      for (int g = 1; g <= mg; ++g) {
        int hi = gend[g];
        double w = 0, stc = 0, sVo = 0, sea = 0, sX = 0, svd = 0;
        for (int i = lo; i <= hi; ++i) {
          double wi = z[i] - (i > 1 ? z[i - 1] : 0.0);
          double Vo = (tetens(tc[i]) * (rh[i] / 100)) / ppk;
          w += wi; stc += wi * tc[i]; sVo += wi * Vo; sea += wi * eal[i]; sX += wi * (tnl[i] - tc[i]);
          svd += wi * (PAI[i] * thickw[i]);
        }
        ltc[g] = stc / w; lVo[g] = sVo / w; lea[g] = sea / w; lX[g] = sX / w; lvd[g] = svd / w;
        lcp[g] = cpair(ltc[g]); lph[g] = phair(ltc[g], ppk);
        int c = (lo + hi + 1) / 2;
        lz[g + 1] = z[c]; lTT[g] = TT[c];
        if (c - cprev == 1) lgt[g] = gtn[c];
        else { double R = 0; for (int j = cprev + 1; j <= c; ++j) R += 1 / gtn[j]; lgt[g] = 1 / R; }
        cprev = c; lo = hi + 1;
      }
      if (m + 1 - cprev == 1) lgt[mg + 1] = gtn[m + 1];
      else { double R = 0; for (int j = cprev + 1; j <= m + 1; ++j) R += 1 / gtn[j]; lgt[mg + 1] = 1 / R; }
      lz[mg + 2] = hgt;
      double k[MN + 1], cdv[MN], XX[MN], tc2[MN + 2], tn2[MN + 2], cc[MN + 2], dd[MN + 2];
      for (int g = 1; g <= mg; ++g) {
        double cda = lcp[g] * lph[g] * (1 - lvd[g]) * (lz[g + 2] - lz[g]) / 2 * timestep;      // Q10
        double ka = lgt[g] * lcp[g];
        if (ka > cda) ka = cda;
        k[mg + 2 - g] = ka; cdv[mg + 1 - g] = cda; XX[mg + 1 - g] = lX[g]; tc2[1 + mg + 1 - g] = ltc[g];
      }
      k[1] = lgt[mg + 1] * cpair(tabove);
      for (int j = 1; j <= sm; ++j) { k[mg + 1 + j] = ksoil[j]; cdv[mg + j] = cdsoil[j]; XX[mg + j] = 0; tc2[1 + mg + j] = soiltc[j]; }
      XX[mg + 1] = Xs;
      tc2[1] = tabove; tc2[N + 2] = tsoil;
      thomas(N, tc2, fc.tsoil, tcan, k, cdv, cfg.n, XX, tn2, cc, dd);
      // tnair = rev(tn2[1:(mg+2)]): (soil1, air bottom..top, tcan); tnsoil = tn2[(mg+2):(N+1)]
      for (int j = 1; j <= sm; ++j) tnsoil[j] = tn2[mg + 1 + j];
      double tnair[MM + 4];
      for (int g = 1; g <= mg + 2; ++g) tnair[g] = tn2[mg + 3 - g];
      // vapour exchange: ThomasV (code.R:153-169) on reversed arrays
      double Vn[MM + 4];
      {
        double rhs = soilrh(fc.theta, soilb, psi_e, Smax, tnsoil[1]); if (rhs > 1) rhs = 1;
        double rhsp = soilrh(fc.thetap, soilb, psi_e, Smax, ps1); if (rhsp > 1) rhsp = 1;
        double Vair = (tet_tcan * (relhumn / 100)) / pkn;
        double eap = tetens(tair) * (relhum / 100);
        double Vsoil = (tetens(tnsoil[1]) * rhs) / pkn;
        double easp = tetens(tsoil) * rhsp;                                                     // Q5
        // rev(Vo) with Vo = c(easp/ppk, lav$Vo, eap/ppk)
        double vt[MM + 4], vk[MM + 3], vcd[MM + 2], vX[MM + 2], vn[MM + 4];
        vt[1] = eap / ppk; vt[mg + 2] = easp / ppk;
        for (int g = 1; g <= mg; ++g) vt[1 + g] = lVo[mg + 1 - g];
        for (int g = 1; g <= mg + 1; ++g) {
          double gt2 = lgt[g];
          if (g <= mg) { double gtmx = lph[g] / (lz[g + 1] - lz[g]) * (2 * timestep); if (gt2 > gtmx) gt2 = gtmx; }
          vk[mg + 2 - g] = gt2;
        }
        for (int g = 1; g <= mg; ++g) {
          vcd[mg + 1 - g] = phair(tnair[g + 1], pkn);
          double vf = (lea[g] / pkn) - lVo[g]; if (vf < 0) vf = 0;
          vX[g] = vf;                                                                           // Q6: not reversed
        }
        thomas(mg, vt, Vsoil, Vair, vk, vcd, cfg.n, vX, vn, cc, dd);
        for (int g = 1; g <= mg + 2; ++g) Vn[g] = vn[mg + 3 - g];
      }
      if (mg < m) {
        // layerinterp(T2 = c(0, lav$TT, TX), TT, y)  (R approx, bisection + linear)
        for (int q = 1; q <= m; ++q) {
          double v = TT[q];
          auto T2 = [&](int j) { return j == 1 ? 0.0 : (j == mg + 2 ? TX : lTT[j - 1]); };
          if (!(v >= 0.0) || !(v <= TX)) { tnn[q] = NAN; ean2[q] = NAN; continue; }
          int i = 1, j = mg + 2;
          while (i < j - 1) { int ij = (i + j) / 2; if (v < T2(ij)) j = ij; else i = ij; }
          double ta, va;
          if (v == T2(j)) { ta = tnair[j]; va = Vn[j]; }
          else if (v == T2(i)) { ta = tnair[i]; va = Vn[i]; }
          else {
            double w = (v - T2(i)) / (T2(j) - T2(i));
            ta = tnair[i] + (tnair[j] - tnair[i]) * w;
            va = Vn[i] + (Vn[j] - Vn[i]) * w;
          }
          tnn[q] = ta; ean2[q] = va * pkn;
        }
      } else {
        for (int i = 1; i <= m; ++i) { tnn[i] = tnair[i + 1]; ean2[i] = Vn[i + 1] * pkn; }
      }
    } else {
      // ---- canopy air in equilibrium with above canopy (code.R:854-867) ----
      double k[MS + 3], XX[MS + 2], tc2[MS + 4], tn2[MS + 4], cc[MS + 4], dd[MS + 4];
      int half = (int)nearbyint(m / 2.0);
      double cs = 0; for (int i = 1; i <= half; ++i) cs += 1 / gtn[i];
      k[1] = (1 / cs) * cpair(tabove);
      for (int j = 1; j <= sm; ++j) { k[1 + j] = ksoil[j]; XX[j] = 0; tc2[1 + j] = soiltc[j]; }
      XX[1] = Xs;
      tc2[1] = tabove; tc2[sm + 2] = tsoil;
      thomas(sm, tc2, fc.tsoil, tcan, k, &cdsoil[0], cfg.n, XX, tn2, cc, dd);
      for (int j = 1; j <= sm; ++j) tnsoil[j] = tn2[1 + j];
      double eair = (relhumn / 100) * tet_tcan;
      for (int i = 1; i <= m; ++i) {
        double wgt = TT[i] / TX;
        tnn[i] = (1 - wgt) * tnsoil[1] + wgt * tcan;
        ean2[i] = (1 - wgt) * (fc.thz ? esoilv[(i - 1) % sm + 1] : esoil) + wgt * eair;
      }
    }
    // ---- limits, humidity, fluxes (code.R:868-897) ----
    if (cta_sync) MC_CTA_SYNC();
    if (tnsoil[1] < tfrost) tnsoil[1] = tfrost;
    if (tnsoil[1] > tairn + 30) tnsoil[1] = tairn + 30;
    double Lt = 0;
    for (int i = 1; i <= m; ++i) Lt += L[i];
    const double shade = (1 - exp(-ref_m * sumPAI));
    const double alb = shade * ref_m + (1 - shade) * refg;
    const double Rlw = kSb * 0.97 * pow4(0.5 * tnsoil[1] + 0.5 * ps1 + 273.15);
    double Gn;
    {
      double a = a0_g * sqphi;
      double t1 = tnn[selG], t0 = tnsoil[1];
      double ph = phair((t1 + t0) / 2, pkn);
      double e0 = exp(-a * gxg0), e1 = exp(-a * gxg1);
      double g = (lm_g * iwmean * ph * uh * a) / (hgt * (e0 - e1) * phi_m);
      double gfree = 0.0012 * ph * qroot(fabs(t1 - t0) / fabs(zg - 0));
      double gt2 = (g > gfree) ? g : gfree;
      Gn = gt2 * cpair(tc[1]) * (t1 - t0);
    }
    const double net = (1 - alb) * Rsw - Rlw;
    const double Gmx = fabs(net - Lt);
    if (Gn > Gmx) Gn = Gmx;
    if (Gn < -Gmx) Gn = -Gmx;
    const double Hn = net - Gn - Lt;
    // ---- commit the new state ----
    for (int i = 1; i <= m; ++i) {
      if (cta_sync) MC_CTA_SYNC();
      double tn = tnn[i];
      double es = tetens(tn);
      double r = (ean2[i] / es) * 100;
      if (r > 100) r = 100;
      if (r < 0) r = 0;
      rh[i] = r;
      double tdws = dewpoint(eal[i], tn);                                                       // Q12
      if (tn < tdws) tn = tdws;
      tc[i] = tn;
      Rlwin[i] = trlw_sum * skyem * (kSb * pow4(tn + 273.15)) + (1 - trlw_sum) * emv * (kSb * pow4(tn + 273.15));
      tleaf[i] = tleafn[i]; uz[i] = uzn[i]; Rabs[i] = Rabsn[i]; gv[i] = gvn[i]; gha[i] = ghan[i]; gt[i] = gtn[i];
    }
    gt[m + 1] = gtn[m + 1];
    for (int j = 1; j <= sm; ++j) soiltc[j] = tnsoil[j];
    tabove = tcan; zabove = zabove_g; relhum = relhumn; tair = tairn; tsoil = fc.tsoil; pk = pkn; H = Hn; G = Gn; psi_m = psi_mn;
    if (!zprev_is_z) { for (int i = 1; i <= m; ++i) zprev[i] = z[i]; precompute_leafgeom(cfg); }
    return ng;
  }

  // runmodel's per-step output harvest (code.R:1222-1265): tout, tleaf, rh, Rswin, Rlwin, L, H, G
  MC_NI void harvest(double reqhgt, int charmode, double temp_i, double relhum0, double rsw0, double skyem0, double* o) const {
    double stc = 0, stl = 0, srh = 0, ssw = 0, slw = 0, sL = 0;
    for (int i = 1; i <= m; ++i) { stc += tc[i]; stl += tleaf[i]; srh += rh[i]; ssw += Rswin[i]; slw += Rlwin[i]; sL += L[i]; }
    o[6] = H; o[7] = G;
    if (charmode || reqhgt != reqhgt) { o[0] = stc / m; o[1] = stl / m; o[2] = srh / m; o[3] = ssw / m; o[4] = slw / m; o[5] = sL; return; }
    o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0;
    if (reqhgt >= hgt) {                                                                        // Q13
      o[0] = tabove; o[1] = stl / m;
      double es1 = tetens(temp_i), es2 = tetens(tabove);
      double r = (((relhum0 / 100) * es1) / es2) * 100; if (r > 100) r = 100;
      o[2] = r; o[3] = rsw0; o[4] = skyem0 * 5.67 * 1e-8 * pow4(tabove + 273.15); o[5] = sL;
    }
    if (reqhgt >= 0 && reqhgt < hgt) {
      int sel = 0; for (int i = 1; i <= m; ++i) if (zprev[i] == reqhgt) { sel = i; break; }
      if (sel) { o[0] = tc[sel]; o[1] = tleaf[sel]; o[2] = rh[sel]; o[3] = Rswin[sel]; o[4] = Rlwin[sel]; o[5] = L[sel]; }
      else { for (int q = 0; q < 6; ++q) o[q] = NAN; }
    }
    if (reqhgt <= 0) {
      int sel = 1; double mn = 1e300;
      for (int i = 1; i <= sm; ++i) { double dd = fabs(zs[i] + reqhgt); if (dd < mn) { mn = dd; sel = i; } }
      o[0] = soiltc[sel]; o[1] = stl / m; o[2] = srh / m; o[3] = ssw / m; o[4] = slw / m; o[5] = sL;
    }
  }
};

}  // namespace mc
