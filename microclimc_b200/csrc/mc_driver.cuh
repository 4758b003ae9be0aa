// Per-column time drivers: spinup (R/code.R:982-1013 as invoked by runmodel, code.R:1155-1159) and the
// runmodel hot loop with its output harvest (code.R:1193-1266).  The time loop lives inside the kernel: a thread
// loads its column once, advances it through all requested steps with the state in thread-private storage, and
// writes the state back once.  Forcing is a shared base series [T][NFORC] (broadcast loads: every lane of a warp
// reads the same address) optionally modified by a per-column affine perturbation.
#pragma once
#include "mc_step.cuh"

namespace mc {

enum { F_TAIR, F_RELHUM, F_PK, F_U, F_SKYEM, F_RSW, F_DP, F_THETA, F_JD, F_LT, F_N };
constexpr int kNPert = 8;

struct TransientArgs {
  long long ncol;            // columns in this shard
  int m, sm;
  long long T;               // rows of `forcing`
  long long t0, t1;          // steps to run: [t0, t1)
  const double* vegp;        // [NVEG][ncol]
  const double* soilp;       // [NSOIL][ncol]
  const double* lat;         // [ncol]
  const double* lon;         // [ncol]
  const double* forcing;     // [T][F_N]
  const double* fsol;        // [T][3] date-only solar terms of each forcing row (sol_day: eot, sin / cos of the declination) or null
  const double* fscale;      // [8][ncol] or null
  const double* foffset;     // [8][ncol] or null
  const double* tsoil;       // [ncol] or null (NaN entries -> mean of the column's air temperature)
  const double* season;      // [T] or null: PAI(t) = PAI * season[t]
  double* state;             // [NSTATE][ncol]
  double* series;            // [t1-t0][8][nsamp] or null
  double* full;              // [t1-t0][NSTATE][nsamp] or null
  const long long* samp_of;  // [ncol] -> sample slot or -1 (null when nsamp == 0)
  long long nsamp;
  double* stats;             // [6][ncol] or null: mean/min/max of tout, then of tleaf, over [t0,t1)
  int* flags;                // [ncol] or null
  RunCfg cfg;
  int spin_steps;            // > 0: paraminit + spinup before the loop (state input ignored)
  int charmode;              // reqhgt "all"/"allclim"
  int spin_standalone;       // spin-up called as spinup() itself: keeps reqhgt, zlafact, surfwet (no Q22 / spinhgt mapping)
  int flag_or;               // OR-ed into every flags[] entry this launch writes (8: the column left the streaming form)
  // general time-varying vegetation (.vegpsort, code.R:904-916): per step PAI[m], pLAI[m], zm0, shared by all columns (nvt == 1)
  // or per column (nvt == ncol); null: time-invariant (or `season`)
  const double* vegt;        // [T][2m+1][nvt]
  long long nvt;
  const double* thetaz;      // [T][sm] soil moisture by depth (theta matrix of runmodel, code.R:1146-1154) or null
  // statistics carried across calls (mc_set_stats_accumulate): raw sums / minima / maxima and the flag bits live in stats_raw / flags_raw
  // between calls; `stats_prev_n` > 0: continue from them (steps already in them), else start afresh.  stats (when given) always receives
  // the statistics over all accumulated steps.
  double* stats_raw;         // [6][ncol] or null
  int* flags_raw;            // [ncol] or null
  long long stats_prev_n;
};
// statistics of a column at the start / end of a launch (shared by both kernels, so that a column redone by the general kernel continues
// from the same partial sums the streaming kernel would have used)
MC_HD void stats_begin(const TransientArgs& a, long long c, double* st, int& flag) {
  st[0] = 0; st[1] = 1e300; st[2] = -1e300; st[3] = 0; st[4] = 1e300; st[5] = -1e300; flag = 0;
  if (a.stats_raw && a.stats_prev_n > 0) {
    for (int q = 0; q < 6; ++q) st[q] = a.stats_raw[(long long)q * a.ncol + c];
    if (a.flags_raw) flag = a.flags_raw[c];
  }
}
MC_HD void stats_end(const TransientArgs& a, long long c, const double* st, int flag) {
  if (a.stats_raw) for (int q = 0; q < 6; ++q) a.stats_raw[(long long)q * a.ncol + c] = st[q];
  if (a.flags_raw) a.flags_raw[c] = flag;
  if (a.stats) {
    const double n = (double)((a.stats_raw ? a.stats_prev_n : 0) + (a.t1 - a.t0));
    a.stats[0 * a.ncol + c] = st[0] / n; a.stats[1 * a.ncol + c] = st[1]; a.stats[2 * a.ncol + c] = st[2];
    a.stats[3 * a.ncol + c] = st[3] / n; a.stats[4 * a.ncol + c] = st[4]; a.stats[5 * a.ncol + c] = st[5];
  }
  if (a.flags) a.flags[c] = flag;
}

MC_HD void forcing_at(const TransientArgs& a, long long t, long long c, double* f) {
  const double* row = a.forcing + t * F_N;
#pragma unroll
  for (int v = 0; v < F_N; ++v) f[v] = row[v];
  if (a.fscale && a.foffset) {
#pragma unroll
    for (int v = 0; v < kNPert; ++v) f[v] = row[v] * a.fscale[(long long)v * a.ncol + c] + a.foffset[(long long)v * a.ncol + c];
    if (f[F_RELHUM] > 100) f[F_RELHUM] = 100;
    if (f[F_RELHUM] < 1) f[F_RELHUM] = 1;
    if (f[F_U] < 0) f[F_U] = 0;
    if (f[F_SKYEM] > 1) f[F_SKYEM] = 1;
    if (f[F_SKYEM] < 0) f[F_SKYEM] = 0;
    if (f[F_RSW] < 0) f[F_RSW] = 0;
    if (f[F_THETA] < 0.001) f[F_THETA] = 0.001;
  }
}

// `live` is false for the padding threads of the last block: they shadow the last column through every barrier and
// write nothing.
template <int MM, int MS>
MC_HD void run_column(const TransientArgs& a, long long c, bool live = true) {
  Column<MM, MS> col;
  col.cta_sync = true;
  col.load_params(a.vegp, a.soilp, a.lat, a.lon, a.ncol, c, a.m, a.sm);
  const int m = a.m;
  double f0[F_N], f[F_N];
  forcing_at(a, 0, c, f0);
  // mean(climdata$temp) of this column (code.R:992, 1165), summed in time order
  double tmean;
  if (a.fscale && a.foffset) {
    const double sc = a.fscale[(long long)F_TAIR * a.ncol + c], of = a.foffset[(long long)F_TAIR * a.ncol + c];
    double s = 0;
    for (long long t = 0; t < a.T; ++t) s += a.forcing[t * F_N + F_TAIR] * sc + of;
    tmean = s / (double)a.T;
  } else {
    double s = 0;
    for (long long t = 0; t < a.T; ++t) s += a.forcing[t * F_N + F_TAIR];
    tmean = s / (double)a.T;
  }
  // .vegpsort(vegp, t + 1) (code.R:904-916): PAI, pLAI and zm0 of forcing row t
  auto vegpsort = [&](long long t) {
    const long long cc = (a.nvt == 1) ? 0 : c;
    const double* row = a.vegt + t * (2 * m + 1) * a.nvt + cc;
    for (int i = 1; i <= m; ++i) { col.PAIbase[i] = row[(long long)(i - 1) * a.nvt]; col.pLAI[i] = row[(long long)(m + i - 1) * a.nvt]; }
    col.zm0 = row[(long long)(2 * m) * a.nvt];
  };
  // theta[, t + 1] of this column (code.R:1202): the per-column perturbation of the theta row applies to every depth
  double thz[MS + 1], thz0[MS + 1];
  auto theta_at = [&](long long t, double* o) {
    for (int j = 0; j < a.sm; ++j) {
      double th = a.thetaz[t * a.sm + j];
      if (a.fscale && a.foffset) { th = th * a.fscale[(long long)F_THETA * a.ncol + c] + a.foffset[(long long)F_THETA * a.ncol + c]; if (th < 0.001) th = 0.001; }
      o[j] = th;
    }
  };
  if (a.thetaz) theta_at(0, thz0);
  const double reqhgt = a.cfg.reqhgt;
  const double spinhgt = a.charmode ? col.hgt / 2 : ((reqhgt == reqhgt && reqhgt > 0) ? reqhgt : NAN);   // code.R:1136-1140
  RunCfg cl = a.cfg;
  cl.reqhgt = a.charmode ? spinhgt : reqhgt;                                                              // reqhgt2, code.R:1199
  if (a.spin_steps > 0) {
    RunCfg cs = a.cfg;
    if (!a.spin_standalone) { cs.reqhgt = spinhgt; cs.zlafact = 1; cs.surfwet = 1; }                     // Q22
    if (a.vegt) vegpsort(0);
    col.precompute(cs, a.season ? a.season[0] : 1.0);
    col.paraminit(col.hgt, f0[F_TAIR], f0[F_U], f0[F_RELHUM], tmean, f0[F_RSW]);
    col.precompute_leafgeom(cs);
    Forcing fc = {f0[F_TAIR], f0[F_RELHUM], f0[F_PK], f0[F_U], tmean, f0[F_SKYEM], f0[F_RSW], f0[F_DP], f0[F_JD], f0[F_LT],
                  f0[F_THETA], f0[F_THETA]};
    if (a.thetaz) { fc.theta = thz0[0]; fc.thetap = thz0[0]; fc.thz = thz0; }                             // theta[,1], theta[,1] (code.R:1158)
    double Hsum = 0;
    for (int i = 1; i <= a.spin_steps; ++i) {
      MC_CTA_SYNC();
      col.step(fc, cs);
      Hsum += col.H;
      col.H = Hsum / i;                                                                                   // previn$H <- mean(H)
    }
    bool same = (cs.reqhgt == cl.reqhgt) || (cs.reqhgt != cs.reqhgt && cl.reqhgt != cl.reqhgt);
    if (a.vegt) vegpsort(a.t0);
    if (!same || a.season || a.vegt || cs.zlafact != cl.zlafact) { col.precompute(cl, a.season ? a.season[a.t0] : 1.0); col.precompute_leafgeom(cl); }
  } else {
    col.load_state(a.state, a.ncol, c);
    if (a.vegt) vegpsort(a.t0);
    col.precompute(cl, a.season ? a.season[a.t0] : 1.0);
    col.precompute_leafgeom(cl);
  }
  const double tsoil = (a.tsoil && a.tsoil[c] == a.tsoil[c]) ? a.tsoil[c] : tmean;
  double st[6];
  int flag;
  stats_begin(a, c, st, flag);
  flag |= a.flag_or;
  const long long slot = (a.nsamp > 0 && a.samp_of) ? a.samp_of[c] : -1;
  const int NS = nstate_rows(m, a.sm);
  for (long long t = a.t0; t < a.t1; ++t) {
    MC_CTA_SYNC();
    forcing_at(a, t, c, f);
    if (a.vegt && t > a.t0) vegpsort(t);
    if ((a.season || a.vegt) && t > a.t0) { col.precompute(cl, a.season ? a.season[t] : 1.0); col.precompute_leafgeom(cl); }
    Forcing fc = {f[F_TAIR], f[F_RELHUM], f[F_PK], f[F_U], tsoil, f[F_SKYEM], f[F_RSW], f[F_DP], f[F_JD], f[F_LT], f[F_THETA],
                  /*Q2*/ f0[F_THETA]};
    if (a.thetaz) { theta_at(t, thz); fc.theta = thz[0]; fc.thetap = thz0[0]; fc.thz = thz; }             // thetap: theta[1, 1] (Q2)
    int ng = col.step(fc, cl);
    if (ng > 1) flag |= 2;
    double o[8];
    col.harvest(reqhgt, a.charmode, f[F_TAIR], f0[F_RELHUM], f0[F_RSW], f0[F_SKYEM], o);
    st[0] += o[0]; if (o[0] < st[1]) st[1] = o[0]; if (o[0] > st[2]) st[2] = o[0];
    st[3] += o[1]; if (o[1] < st[4]) st[4] = o[1]; if (o[1] > st[5]) st[5] = o[1];
    if (!isfinite(o[0]) || !isfinite(o[1])) flag |= 1;
    if (slot >= 0 && live) {
      if (a.series) for (int q = 0; q < 8; ++q) a.series[((t - a.t0) * 8 + q) * a.nsamp + slot] = o[q];
      if (a.full) col.store_state(a.full + (t - a.t0) * NS * a.nsamp, a.nsamp, slot);
    }
  }
  if (!live) return;
  stats_end(a, c, st, flag);
  col.store_state(a.state, a.ncol, c);
}

// one step with explicit climvars per column (runonestep as an API call, code.R:687-689)
struct OneStepArgs {
  long long ncol; int m, sm;
  const double *vegp, *soilp, *lat, *lon;
  const double* climvars;    // [8][ncol]: tair, relhum, pk, u, tsoil, skyem, Rsw, dp
  const double *theta, *thetap;   // [ncol]
  double jd, lt;
  double* state; int* nmerged;
  RunCfg cfg;
};
template <int MM, int MS>
MC_HD void onestep_column(const OneStepArgs& a, long long c) {
  Column<MM, MS> col;
  col.load_params(a.vegp, a.soilp, a.lat, a.lon, a.ncol, c, a.m, a.sm);
  col.load_state(a.state, a.ncol, c);
  col.precompute(a.cfg, 1.0);
  col.precompute_leafgeom(a.cfg);
  auto C = [&](int r) { return a.climvars[(long long)r * a.ncol + c]; };
  Forcing fc = {C(0), C(1), C(2), C(3), C(4), C(5), C(6), C(7), a.jd, a.lt, a.theta[c], a.thetap[c]};
  int ng = col.step(fc, a.cfg);
  if (a.nmerged) a.nmerged[c] = ng;
  col.store_state(a.state, a.ncol, c);
}

// batched paraminit (code.R:566-585); args [6][ncol]: hgt, tair, u, relhum, tsoil, Rsw
template <int MM, int MS>
MC_HD void paraminit_column(long long ncol, int m, int sm, const double* args, double* state, long long c) {
  Column<MM, MS> col;
  col.m = m; col.sm = sm;
  auto A = [&](int r) { return args[(long long)r * ncol + c]; };
  col.paraminit(A(0), A(1), A(2), A(3), A(4), A(5));
  for (int i = 1; i <= sm; ++i) col.zs[i] = 2 / pow((double)sm, 2.42) * pow((double)i, 2.42);    // sz (code.R:579)
  col.store_state(state, ncol, c);
}

}  // namespace mc
