// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
// PARITY UNPINNED vs real R: the reference ships no tests or golden vectors and neither R nor microctools can run here;
// pinned only by SURVEY Appendix F known answers, the independent numpy transliteration goldens and the decoded datasets.
//
// extern "C" entry points (ctypes-callable) around the CPU restatement in mc_model.hpp / mc_steady.hpp,
// plus the drivers spinup (code.R:982-1013) and the runmodel hot loop + output harvest (code.R:1193-1266),
// batched over independent columns with the SAME column-interleaved array layouts the product C-ABI uses
// (include/microclimc_b200.h), so tests can feed both sides the same buffers.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
#include "mc_steady.hpp"
#include <cstdint>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace {
// ---- flat layouts (must agree with include/microclimc_b200.h; tests/test_abi_layout.py checks) ----
enum { V_HGT, V_X, V_LW, V_CD, V_HGTG, V_ZM0, V_REFLS, V_REFG, V_REFW, V_REFLP, V_VEGEM, V_GSMAX, V_Q50, V_CPW, V_PHW,
       V_KWOOD, V_CLUMP, V_NSCAL };
enum { S_SMAX, S_SMIN, S_ALPHA, S_N, S_KSAT, S_VQ, S_VM, S_VO, S_MC, S_RHO, S_B, S_PSIE, S_NSCAL };
enum { P_TABOVE, P_ZABOVE, P_RELHUM, P_TAIR, P_TSOIL, P_PK, P_H, P_G, P_PSIM, P_NSCAL };
enum { F_TAIR, F_RELHUM, F_PK, F_U, F_SKYEM, F_RSW, F_DP, F_THETA, F_JD, F_LT, F_N };
enum { C_TIMESTEP, C_EDGEDIST, C_REQHGT, C_ZU, C_MERID, C_DST, C_N, C_METOPEN, C_WINDHGT, C_ZLAFACT, C_SURFWET,
       C_SPINSTEPS, C_HARVEST /*0 numeric/NA reqhgt, 1 "all"/"allclim"*/, C_N_ };

inline int nveg(int m) { return V_NSCAL + 4 * m; }
inline int nsoil(int sm) { return S_NSCAL + sm; }
inline int nstate(int m, int sm) { return P_NSCAL + 11 * m + (m + 1) + 2 * sm; }

Vegp load_vegp(const double* v, int64_t ncol, int64_t c, int m) {
  Vegp p; p.m = m;
  auto g = [&](int r) { return v[(int64_t)r * ncol + c]; };
  p.hgt = g(V_HGT); p.x = g(V_X); p.lw = g(V_LW); p.cd = g(V_CD); p.hgtg = g(V_HGTG); p.zm0 = g(V_ZM0);
  p.refls = g(V_REFLS); p.refg = g(V_REFG); p.refw = g(V_REFW); p.reflp = g(V_REFLP); p.vegem = g(V_VEGEM);
  p.gsmax = g(V_GSMAX); p.q50 = g(V_Q50); p.cpw = g(V_CPW); p.phw = g(V_PHW); p.kwood = g(V_KWOOD); p.clump = g(V_CLUMP);
  p.PAI = Vec(m); p.pLAI = Vec(m); p.iw = Vec(m); p.thickw = Vec(m);
  for (int i = 1; i <= m; ++i) {
    p.PAI(i) = g(V_NSCAL + i - 1); p.pLAI(i) = g(V_NSCAL + m + i - 1);
    p.iw(i) = g(V_NSCAL + 2 * m + i - 1); p.thickw(i) = g(V_NSCAL + 3 * m + i - 1);
  }
  return p;
}
Soilp load_soilp(const double* s, int64_t ncol, int64_t c, int sm) {
  Soilp p; p.sm = sm;
  auto g = [&](int r) { return s[(int64_t)r * ncol + c]; };
  p.Smax = g(S_SMAX); p.Smin = g(S_SMIN); p.alpha = g(S_ALPHA); p.n = g(S_N); p.Ksat = g(S_KSAT); p.Vq = g(S_VQ);
  p.Vm = g(S_VM); p.Vo = g(S_VO); p.Mc = g(S_MC); p.rho = g(S_RHO); p.b = g(S_B); p.psi_e = g(S_PSIE);
  p.z = Vec(sm);
  for (int i = 1; i <= sm; ++i) p.z(i) = g(S_NSCAL + i - 1);
  return p;
}
State load_state(const double* s, int64_t ncol, int64_t c, int m, int sm) {
  State p; p.m = m; p.sm = sm;
  auto g = [&](int r) { return s[(int64_t)r * ncol + c]; };
  p.tabove = g(P_TABOVE); p.zabove = g(P_ZABOVE); p.relhum = g(P_RELHUM); p.tair = g(P_TAIR); p.tsoil = g(P_TSOIL);
  p.pk = g(P_PK); p.H = g(P_H); p.G = g(P_G); p.psi_m = g(P_PSIM);
  Vec* arrs[11] = {&p.tc, &p.tleaf, &p.uz, &p.z, &p.rh, &p.Rabs, &p.gv, &p.gha, &p.L, &p.Rswin, &p.Rlwin};
  int r = P_NSCAL;
  for (int a = 0; a < 11; ++a) { *arrs[a] = Vec(m); for (int i = 1; i <= m; ++i) (*arrs[a])(i) = g(r++); }
  p.gt = Vec(m + 1); for (int i = 1; i <= m + 1; ++i) p.gt(i) = g(r++);
  p.soiltc = Vec(sm); for (int i = 1; i <= sm; ++i) p.soiltc(i) = g(r++);
  p.sz = Vec(sm); for (int i = 1; i <= sm; ++i) p.sz(i) = g(r++);
  return p;
}
void store_state(const State& p, double* s, int64_t ncol, int64_t c) {
  int m = p.m, sm = p.sm;
  auto w = [&](int r, double v) { s[(int64_t)r * ncol + c] = v; };
  w(P_TABOVE, p.tabove); w(P_ZABOVE, p.zabove); w(P_RELHUM, p.relhum); w(P_TAIR, p.tair); w(P_TSOIL, p.tsoil);
  w(P_PK, p.pk); w(P_H, p.H); w(P_G, p.G); w(P_PSIM, p.psi_m);
  const Vec* arrs[11] = {&p.tc, &p.tleaf, &p.uz, &p.z, &p.rh, &p.Rabs, &p.gv, &p.gha, &p.L, &p.Rswin, &p.Rlwin};
  int r = P_NSCAL;
  for (int a = 0; a < 11; ++a) for (int i = 1; i <= m; ++i) w(r++, (*arrs[a])(i));
  for (int i = 1; i <= m + 1; ++i) w(r++, p.gt(i));
  for (int i = 1; i <= sm; ++i) w(r++, p.soiltc(i));
  for (int i = 1; i <= sm; ++i) w(r++, p.sz(i));
}
StepCfg load_cfg(const double* cfg, double lat, double lon) {
  StepCfg c;
  c.timestep = cfg[C_TIMESTEP]; c.lat = lat; c.lon = lon; c.edgedist = cfg[C_EDGEDIST]; c.reqhgt = cfg[C_REQHGT];
  c.zu = cfg[C_ZU]; c.merid = cfg[C_MERID]; c.dst = cfg[C_DST]; c.n = cfg[C_N]; c.metopen = cfg[C_METOPEN] != 0;
  c.windhgt = cfg[C_WINDHGT]; c.zlafact = cfg[C_ZLAFACT]; c.surfwet = cfg[C_SURFWET];
  return c;
}
// forcing of column c at step t: shared base row, optional per-column affine perturbation with the clamps
// documented in include/microclimc_b200.h (mc_forcing).
void forcing_at(const double* forcing, int64_t t, const double* fscale, const double* foffset, int64_t ncol, int64_t c,
                double* f /*F_N*/) {
  const double* row = forcing + t * F_N;
  for (int v = 0; v < F_N; ++v) f[v] = row[v];
  if (fscale && foffset) {
    for (int v = 0; v <= F_THETA; ++v) f[v] = row[v] * fscale[(int64_t)v * ncol + c] + foffset[(int64_t)v * ncol + c];
    if (f[F_RELHUM] > 100) f[F_RELHUM] = 100;
    if (f[F_RELHUM] < 1) f[F_RELHUM] = 1;
    if (f[F_U] < 0) f[F_U] = 0;
    if (f[F_SKYEM] > 1) f[F_SKYEM] = 1;
    if (f[F_SKYEM] < 0) f[F_SKYEM] = 0;
    if (f[F_RSW] < 0) f[F_RSW] = 0;
    if (f[F_THETA] < 0.001) f[F_THETA] = 0.001;
  }
}
void apply_season(Vegp& v, const Vegp& base, double s) {
  for (int i = 1; i <= base.m; ++i) v.PAI(i) = base.PAI(i) * s;
}
// runmodel output harvest for one step (code.R:1222-1265): tout, tleaf, rh, Rswin, Rlwin, L, H, G
void harvest(const State& p, const Vegp& vegp, const Soilp& soilp, double reqhgt, int charmode, double temp_i,
             const double* f0 /*forcing row 0 of this column*/, double* o /*8*/) {
  int m = p.m;
  double sb = 5.67 * 1e-8;
  auto meanv = [&](const Vec& v) { return vsum(v) / v.n; };
  o[6] = p.H; o[7] = p.G;
  if (charmode || is_na(reqhgt)) {     // NA reqhgt (code.R:1222-1228); the character modes keep full tables instead
    o[0] = meanv(p.tc); o[1] = meanv(p.tleaf); o[2] = meanv(p.rh); o[3] = meanv(p.Rswin); o[4] = meanv(p.Rlwin); o[5] = vsum(p.L);
    return;
  }
  o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0;
  if (reqhgt >= vegp.hgt) {            // Q13: whole climdata vectors assigned into scalars -> element 1
    o[0] = p.tabove; o[1] = meanv(p.tleaf);
    double es1 = 0.6108 * std::exp(17.27 * temp_i / (temp_i + 237.3));
    double es2 = 0.6108 * std::exp(17.27 * p.tabove / (p.tabove + 237.3));
    double ea = (f0[F_RELHUM] / 100) * es1;
    double relhum = (ea / es2) * 100; if (relhum > 100) relhum = 100;
    o[2] = relhum; o[3] = f0[F_RSW];
    double tk = p.tabove + 273.15;
    o[4] = f0[F_SKYEM] * 5.67 * 1e-8 * ((tk * tk) * (tk * tk));
    o[5] = vsum(p.L);
  }
  if (reqhgt >= 0 && reqhgt < vegp.hgt) {
    int sel = 0; for (int i = 1; i <= m; ++i) if (p.z(i) == reqhgt) { sel = i; break; }
    if (sel) { o[0] = p.tc(sel); o[1] = p.tleaf(sel); o[2] = p.rh(sel); o[3] = p.Rswin(sel); o[4] = p.Rlwin(sel); o[5] = p.L(sel); }
    else { for (int q = 0; q < 6; ++q) o[q] = NA; }
  }
  if (reqhgt <= 0) {
    int sel = 1; double mn = 1e300;
    for (int i = 1; i <= soilp.sm; ++i) { double dd = std::fabs(soilp.z(i) + reqhgt); if (dd < mn) { mn = dd; sel = i; } }
    o[0] = p.soiltc(sel); o[1] = meanv(p.tleaf); o[2] = meanv(p.rh); o[3] = meanv(p.Rswin); o[4] = meanv(p.Rlwin); o[5] = vsum(p.L);
  }
  (void)sb;
}
}  // namespace

extern "C" {

int orc_nveg(int m) { return nveg(m); }
int orc_nsoil(int sm) { return nsoil(sm); }
int orc_nstate(int m, int sm) { return nstate(m, sm); }
int orc_nforc() { return F_N; }
int orc_ncfg() { return C_N_; }

// ---- scalar compat probes (unit tests) -------------------------------------------------------
double orc_satvap(double tc) { return satvap(tc); }
double orc_dewpoint(double ea, double tc) { return dewpoint(ea, tc); }
double orc_phair(double tc, double pk) { return phair(tc, pk); }
double orc_cpair(double tc) { return cpair(tc); }
double orc_zeroplanedis(double h, double p) { return zeroplanedis(h, p); }
double orc_roughlength(double h, double p, double zm0) { return roughlength(h, p, zm0); }
double orc_mixinglength(double h, double p) { return mixinglength(h, p); }
double orc_attencoef(double h, double p, double cd, double iw, double phi) { return attencoef(h, p, cd, iw, phi); }
double orc_solalt(double lt, double lat, double lon, double jd, double merid, double dst) { return solalt(lt, lat, lon, jd, merid, dst); }
double orc_jday(int y, int mo, int d) { return jday(y, mo, d); }
double orc_gcanopy(double uh, double z1, double z0, double tc1, double tc0, double hgt, double PAI, double cd, double iw, double phi_m, double pk) {
  return gcanopy(uh, z1, z0, tc1, tc0, hgt, PAI, cd, iw, phi_m, pk);
}
double orc_gturb(double u, double zu, double z1, double z0, double hgt, double PAI, double tc, double psi_m, double psi_h, double zm0, double pk) {
  return gturb(u, zu, z1, z0, hgt, PAI, tc, psi_m, psi_h, zm0, pk);
}
double orc_gforcedfree(double d, double u, double tc, double dtc, double pk, double dtmin) { return gforcedfree(d, u, tc, dtc, pk, dtmin); }
double orc_windprofile(double ui, double zi, double zo, double a, double PAI, double hgt, double psi_m, double hgtg, double zm0) {
  return windprofile(ui, zi, zo, a, PAI, hgt, psi_m, hgtg, zm0);
}
double orc_windprofile2(double ui, double zi, double zo, double a, double hgt, double PAI, double psi_m) {   // .windprofile, NMR.R:378-404
  return windprofile2(ui, zi, zo, a, hgt, PAI, psi_m);
}
double orc_windcanopy(double uh, double z, double hgt, double PAI, double cd, double iw, double phi_m, double edgedist, double uref, double zref) {
  return windcanopy(uh, z, hgt, PAI, cd, iw, phi_m, edgedist, uref, zref);
}
double orc_abovecanopytemp(double tz, double uz, double zu, double zo, double H, double hgt, double PAI, double zm0, double pk, double psi_h) {
  return abovecanopytemp(tz, uz, zu, zo, H, hgt, PAI, zm0, pk, psi_h);
}
double orc_leafem(double tc, double vegem) { double tk = tc + 273.15; return vegem * 5.67 * 1e-8 * ((tk * tk) * (tk * tk)); }   // code.R:59-62

// Thomas (code.R:94-125): tc[N+2], k[N+1], cd[ncd] (ncd = N or 1), X[nX] (nX = N or 1) -> tn[N+2]
void orc_thomas(int N, const double* tc, double tmsoil, double tair, const double* k, const double* cd, int ncd, double f,
                const double* X, int nX, double* tn) {
  Vec vtc(N + 2), vk(N + 1), vcd(ncd), vX(nX);
  for (int i = 1; i <= N + 2; ++i) vtc(i) = tc[i - 1];
  for (int i = 1; i <= N + 1; ++i) vk(i) = k[i - 1];
  for (int i = 1; i <= ncd; ++i) vcd(i) = cd[i - 1];
  for (int i = 1; i <= nX; ++i) vX(i) = X[i - 1];
  Vec o = Thomas(vtc, tmsoil, tair, vk, vcd, f, vX);
  for (int i = 1; i <= N + 2; ++i) tn[i - 1] = o(i);
}

// paraminit (code.R:566-585) -> state[NSTATE] (single column, ncol = 1 layout)
void orc_paraminit(int m, int sm, double hgt, double tair, double u, double relhum, double tsoil, double Rsw, double* state) {
  State s = paraminit(m, sm, hgt, tair, u, relhum, tsoil, Rsw);
  store_state(s, state, 1, 0);
}
void orc_soilinit_z(int m, double sdepth, double reqdepth, double* z) {
  Vec v = soilinit_z(m, sdepth, reqdepth);
  for (int i = 1; i <= m; ++i) z[i - 1] = v(i);
}

// runonestep (code.R:687-902) for ncol columns; state in/out [NSTATE][ncol].
// climvars [8][ncol]: tair, relhum, pk, u, tsoil, skyem, Rsw, dp.  tme = (jd, lt).  Returns 0 or -1 (m < 3).
int orc_runonestep(int64_t ncol, int m, int sm, const double* vegp, const double* soilp, const double* climvars,
                   const double* lat, const double* lon, const double* cfg, double jd, double lt, const double* theta,
                   const double* thetap, double* state, int* nmerged) {
  if (m < 3) return -1;
  for (int64_t c = 0; c < ncol; ++c) {
    Vegp v = load_vegp(vegp, ncol, c, m);
    Soilp s = load_soilp(soilp, ncol, c, sm);
    State p = load_state(state, ncol, c, m, sm), q;
    Climvars cv = {climvars[0 * ncol + c], climvars[1 * ncol + c], climvars[2 * ncol + c], climvars[3 * ncol + c],
                   climvars[4 * ncol + c], climvars[5 * ncol + c], climvars[6 * ncol + c], climvars[7 * ncol + c]};
    StepCfg sc = load_cfg(cfg, lat[c], lon[c]);
    int nm = 0;
    int rc = runonestep(cv, p, v, s, sc, Tme{jd, lt}, theta[c], thetap[c], q, &nm);
    if (rc) return rc;
    if (nmerged) nmerged[c] = nm;
    store_state(q, state, ncol, c);
  }
  return 0;
}

// Transient driver: optional paraminit + spinup (code.R:982-1013 as called from code.R:1155-1159), then the
// hourly loop with output harvest (code.R:1193-1266).
//  forcing [T][F_N] shared; fscale/foffset [8][ncol] or NULL; tsoil [ncol] (NaN -> mean of the column's temp);
//  pai_season [T] or NULL (PAI(t) = PAI * season[t]); state [NSTATE][ncol] in (when spin_steps == 0) / out;
//  series [T][8][nsamp] for columns samp_idx[]; stats [6][ncol] = mean,min,max of tout then of tleaf;
//  fullseries [T][NSTATE][nsamp]; flags [ncol] (bit0: non-finite tout/tleaf seen).
//  vegt [T][2m+1][nvt] or NULL: PAI[m], pLAI[m], zm0 of each forcing row (.vegpsort, code.R:904-916), one series shared by all columns
//  (nvt == 1) or one per column (nvt == ncol); thetaz [T][sm] or NULL: soil moisture by depth (theta matrix, code.R:1146-1154), the
//  per-column perturbation of the theta row applies to every depth.
int orc_run_transient2(int64_t ncol, int m, int sm, int64_t T, const double* vegp, const double* soilp, const double* forcing,
                       const double* fscale, const double* foffset, const double* tsoil_in, const double* lat,
                       const double* lon, const double* cfg, const double* pai_season, double* state, double* series,
                       const int64_t* samp_idx, int64_t nsamp, double* stats, double* fullseries, int32_t* flags, int nthreads,
                       const double* vegt, int64_t nvt, const double* thetaz);
int orc_run_transient(int64_t ncol, int m, int sm, int64_t T, const double* vegp, const double* soilp, const double* forcing,
                      const double* fscale, const double* foffset, const double* tsoil_in, const double* lat,
                      const double* lon, const double* cfg, const double* pai_season, double* state, double* series,
                      const int64_t* samp_idx, int64_t nsamp, double* stats, double* fullseries, int32_t* flags, int nthreads) {
  return orc_run_transient2(ncol, m, sm, T, vegp, soilp, forcing, fscale, foffset, tsoil_in, lat, lon, cfg, pai_season, state, series,
                            samp_idx, nsamp, stats, fullseries, flags, nthreads, nullptr, 0, nullptr);
}
int orc_run_transient2(int64_t ncol, int m, int sm, int64_t T, const double* vegp, const double* soilp, const double* forcing,
                       const double* fscale, const double* foffset, const double* tsoil_in, const double* lat,
                       const double* lon, const double* cfg, const double* pai_season, double* state, double* series,
                       const int64_t* samp_idx, int64_t nsamp, double* stats, double* fullseries, int32_t* flags, int nthreads,
                       const double* vegt, int64_t nvt, const double* thetaz) {
  if (m < 3) return -1;
  if (thetaz && sm > m) return -3;          // R would recycle the canopy vectors over the longer soil vector: not a defined run
  const int NS = nstate(m, sm);
  const int spin = (int)cfg[C_SPINSTEPS];
  const int charmode = (int)cfg[C_HARVEST];
  std::vector<int64_t> samp_of(ncol, -1);
  for (int64_t j = 0; j < nsamp; ++j) samp_of[samp_idx[j]] = j;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t c = 0; c < ncol; ++c) {
    Vegp vbase = load_vegp(vegp, ncol, c, m);
    Vegp v = vbase;
    Soilp s = load_soilp(soilp, ncol, c, sm);
    StepCfg sc = load_cfg(cfg, lat[c], lon[c]);
    double f0[F_N], f[F_N];
    forcing_at(forcing, 0, fscale, foffset, ncol, c, f0);
    double tmean = 0;
    for (int64_t t = 0; t < T; ++t) { forcing_at(forcing, t, fscale, foffset, ncol, c, f); tmean += f[F_TAIR]; }
    tmean = tmean / (double)T;                                   // mean(climdata$temp)
    double reqhgt = sc.reqhgt;
    double spinhgt = charmode ? vbase.hgt / 2 : ((!is_na(reqhgt) && reqhgt > 0) ? reqhgt : NA);   // code.R:1136-1140
    State p, q;
    auto vegpsort = [&](int64_t t) {                             // .vegpsort(vegp, t + 1), code.R:904-916
      const int64_t cc = (nvt == 1) ? 0 : c;
      for (int i = 1; i <= m; ++i) {
        v.PAI(i) = vegt[(t * (2 * m + 1) + (i - 1)) * nvt + cc];
        v.pLAI(i) = vegt[(t * (2 * m + 1) + m + (i - 1)) * nvt + cc];
      }
      v.zm0 = vegt[(t * (2 * m + 1) + 2 * m) * nvt + cc];
    };
    double thz[MAXN + 2], thz0[MAXN + 2];
    auto theta_at = [&](int64_t t, double* o) {                  // theta[, t + 1] of this column
      for (int j = 0; j < sm; ++j) {
        double th = thetaz[t * sm + j];
        if (fscale && foffset) { th = th * fscale[(int64_t)F_THETA * ncol + c] + foffset[(int64_t)F_THETA * ncol + c]; if (th < 0.001) th = 0.001; }
        o[j] = th;
      }
    };
    if (thetaz) theta_at(0, thz0);
    if (spin > 0) {                                              // spinup (code.R:982-1013)
      if (pai_season) apply_season(v, vbase, pai_season[0]);
      if (vegt) vegpsort(0);
      p = paraminit(m, sm, v.hgt, f0[F_TAIR], f0[F_U], f0[F_RELHUM], tmean, f0[F_RSW]);
      Climvars cv = {f0[F_TAIR], f0[F_RELHUM], f0[F_PK], f0[F_U], tmean, f0[F_SKYEM], f0[F_RSW], f0[F_DP]};
      StepCfg ss = sc; ss.reqhgt = spinhgt; ss.zlafact = 1; ss.surfwet = 1;    // Q22
      double Hsum = 0;
      for (int i = 1; i <= spin; ++i) {
        if (thetaz) runonestep(cv, p, v, s, ss, Tme{f0[F_JD], f0[F_LT]}, thz0[0], thz0[0], q, nullptr, thz0);   // theta[,1], theta[,1] (code.R:1158)
        else runonestep(cv, p, v, s, ss, Tme{f0[F_JD], f0[F_LT]}, f0[F_THETA], f0[F_THETA], q);
        Hsum += q.H;
        q.H = Hsum / i;                                          // previn$H <- mean(H)
        p = q;
      }
    } else {
      p = load_state(state, ncol, c, m, sm);
    }
    double tsoil = (tsoil_in && !is_na(tsoil_in[c])) ? tsoil_in[c] : tmean;
    StepCfg sl = sc; sl.reqhgt = charmode ? spinhgt : reqhgt;    // reqhgt2 (code.R:1199)
    double st[6] = {0, 1e300, -1e300, 0, 1e300, -1e300};
    int32_t flag = 0;
    for (int64_t t = 0; t < T; ++t) {
      forcing_at(forcing, t, fscale, foffset, ncol, c, f);
      if (pai_season) apply_season(v, vbase, pai_season[t]);
      if (vegt) vegpsort(t);
      Climvars cv = {f[F_TAIR], f[F_RELHUM], f[F_PK], f[F_U], tsoil, f[F_SKYEM], f[F_RSW], f[F_DP]};
      if (thetaz) { theta_at(t, thz); runonestep(cv, p, v, s, sl, Tme{f[F_JD], f[F_LT]}, thz[0], /*Q2: theta[1,1]*/ thz0[0], q, nullptr, thz); }
      else runonestep(cv, p, v, s, sl, Tme{f[F_JD], f[F_LT]}, f[F_THETA], /*Q2*/ f0[F_THETA], q);
      p = q;
      double o[8];
      harvest(p, v, s, reqhgt, charmode, f[F_TAIR], f0, o);
      st[0] += o[0]; if (o[0] < st[1]) st[1] = o[0]; if (o[0] > st[2]) st[2] = o[0];
      st[3] += o[1]; if (o[1] < st[4]) st[4] = o[1]; if (o[1] > st[5]) st[5] = o[1];
      if (!std::isfinite(o[0]) || !std::isfinite(o[1])) flag |= 1;
      int64_t j = samp_of[c];
      if (j >= 0) {
        if (series) for (int q8 = 0; q8 < 8; ++q8) series[(t * 8 + q8) * nsamp + j] = o[q8];
        if (fullseries) store_state(p, fullseries + t * NS * nsamp, nsamp, j);
      }
    }
    if (stats) {
      stats[0 * ncol + c] = st[0] / (double)T; stats[1 * ncol + c] = st[1]; stats[2 * ncol + c] = st[2];
      stats[3 * ncol + c] = st[3] / (double)T; stats[4 * ncol + c] = st[4]; stats[5 * ncol + c] = st[5];
    }
    if (flags) flags[c] = flag;
    store_state(p, state, ncol, c);
  }
  return 0;
}

// Steady-state driver (runmodelS + .runmodelsnow, NMR.R:717-885, 514-677) over ncol columns x T hours.
//  vegp [NVEG][ncol] (static per-layer PAI/pLAI; PAI(t) = PAI * pai_season[t] when given);
//  clim [T][9]: temp, relhum, pres, windspeed, swrad, difrad, skyem, jd, lt;
//  nmr  [T][3 + 8]: D0cm, WC0cm, SNOWDEP(cm), SN1..SN8   (shared across columns) -- or per column when
//  nmr_percol != 0: [T][11][ncol];  scfg: reqhgt, metopen, windhgt, surfwet, groundem
//  out [T][7][ncol]: Tloc, tleaf, RHloc, RSWloc, RLWloc, windspeed, dp;   stats [6][ncol] (Tloc, tleaf)
//  vegt [T][2m+1][nvt] or NULL: PAI[m], pLAI[m] (and zm0, unused on this path) per hour — runmodelS takes vegp$PAI as an m x T matrix
//  (NMR.R:730-735: apply(vegp$PAI, 2, sum)); nvt == 1 shared by all columns, nvt == ncol one per column.
int orc_run_steady2(int64_t ncol, int m, int sm, int64_t T, const double* vegp, const double* soilp, const double* clim,
                    const double* nmr, int nmr_percol, const double* lat, const double* lon, const double* scfg,
                    const double* pai_season, double* out, double* stats, int nthreads, const double* vegt, int64_t nvt);
int orc_run_steady(int64_t ncol, int m, int sm, int64_t T, const double* vegp, const double* soilp, const double* clim,
                   const double* nmr, int nmr_percol, const double* lat, const double* lon, const double* scfg,
                   const double* pai_season, double* out, double* stats, int nthreads) {
  return orc_run_steady2(ncol, m, sm, T, vegp, soilp, clim, nmr, nmr_percol, lat, lon, scfg, pai_season, out, stats, nthreads, nullptr, 0);
}
int orc_run_steady2(int64_t ncol, int m, int sm, int64_t T, const double* vegp, const double* soilp, const double* clim,
                    const double* nmr, int nmr_percol, const double* lat, const double* lon, const double* scfg,
                    const double* pai_season, double* out, double* stats, int nthreads, const double* vegt, int64_t nvt) {
  int rc_all = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t c = 0; c < ncol; ++c) {
    Vegp v = load_vegp(vegp, ncol, c, m);
    Soilp s = load_soilp(soilp, ncol, c, sm);
    SteadyCfg cfg; cfg.reqhgt = scfg[0]; cfg.metopen = scfg[1] != 0; cfg.windhgt = scfg[2]; cfg.surfwet = scfg[3];
    cfg.groundem = scfg[4]; cfg.lat = lat[c]; cfg.lon = lon[c];
    double st[6] = {0, 1e300, -1e300, 0, 1e300, -1e300};
    double pai[MAXM], plai[MAXM];
    for (int64_t t = 0; t < T; ++t) {
      const double* cr = clim + t * 9;
      double season = pai_season ? pai_season[t] : 1.0;
      for (int i = 0; i < m; ++i) { pai[i] = v.PAI(i + 1) * season; plai[i] = v.pLAI(i + 1); }
      if (vegt) {
        const int64_t cc = (nvt == 1) ? 0 : c;
        for (int i = 0; i < m; ++i) { pai[i] = vegt[(t * (2 * m + 1) + i) * nvt + cc]; plai[i] = vegt[(t * (2 * m + 1) + m + i) * nvt + cc]; }
      }
      SteadyVeg sv;
      sv.hgt = v.hgt; sv.x = v.x; sv.lw = v.lw; sv.cd = v.cd; sv.iwmean = vmean(v.iw); sv.refls = v.refls; sv.refw = v.refw;
      sv.refg = v.refg; sv.vegem = v.vegem; sv.gsmax = v.gsmax; sv.q50 = v.q50; sv.clump = v.clump;
      steady_veg_reduce(pai, plai, m, v.hgt, cfg.reqhgt, sv.PAIt, sv.LAI, sv.PAIu);
      SteadyClim sc = {cr[0], cr[1], cr[2], cr[3], cr[4], cr[5], cr[6], Tme{cr[7], cr[8]}};
      double SN[8]; SteadyNmr nm;
      if (nmr_percol) {
        const double* nr = nmr + t * 11 * ncol;
        nm.tground = nr[0 * ncol + c]; nm.theta = nr[1 * ncol + c]; nm.snowdep_cm = nr[2 * ncol + c];
        for (int q = 0; q < 8; ++q) SN[q] = nr[(3 + q) * ncol + c];
      } else {
        const double* nr = nmr + t * 11;
        nm.tground = nr[0]; nm.theta = nr[1]; nm.snowdep_cm = nr[2];
        for (int q = 0; q < 8; ++q) SN[q] = nr[3 + q];
      }
      nm.SN = SN;
      SteadyOut o;
      int rc = runmodelS_point(sc, sv, s, nm, cfg, o);
      if (rc) {
#pragma omp atomic write
        rc_all = rc;
        o = {NA, NA, NA, NA, NA, NA, NA};
      }
      if (out) {
        double* op = out + t * 7 * ncol;
        op[0 * ncol + c] = o.Tloc; op[1 * ncol + c] = o.tleaf; op[2 * ncol + c] = o.RHloc; op[3 * ncol + c] = o.RSWloc;
        op[4 * ncol + c] = o.RLWloc; op[5 * ncol + c] = o.windspeed; op[6 * ncol + c] = o.dp;
      }
      st[0] += o.Tloc; if (o.Tloc < st[1]) st[1] = o.Tloc; if (o.Tloc > st[2]) st[2] = o.Tloc;
      st[3] += o.tleaf; if (o.tleaf < st[4]) st[4] = o.tleaf; if (o.tleaf > st[5]) st[5] = o.tleaf;
    }
    if (stats) {
      stats[0 * ncol + c] = st[0] / (double)T; stats[1 * ncol + c] = st[1]; stats[2 * ncol + c] = st[2];
      stats[3 * ncol + c] = st[3] / (double)T; stats[4 * ncol + c] = st[4]; stats[5 * ncol + c] = st[5];
    }
  }
  return rc_all;
}

// tleafS probe (NMR.R:432-493), 17 scalars in -> (tleaf, tn, rh)
void orc_tleafS(const double* a, double* o) {
  TleafS r = tleafS(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[15], a[16]);
  o[0] = r.tleaf; o[1] = r.tn; o[2] = r.rh;
}
int orc_snowtempf(double snowdepth_m, const double* SN, double reqhgt, double* o) { return snowtempf(snowdepth_m, SN, reqhgt, *o) ? 0 : -2; }

}  // extern "C"
