// TEST-ONLY host build of the device column code (microclimc_b200/csrc/mc_step.cuh, mc_driver.cuh).
// Lets the CPU-only test suite check the product's fused step logic against the oracle without a GPU.
// It is NOT part of the product library, is never loaded by microclimc_b200/, and is not a CPU fallback.
#include <cmath>
#include <cstdlib>
#include <cstdint>
#include "../../microclimc_b200/csrc/mc_driver.cuh"
#include "../../microclimc_b200/csrc/mc_fast.cuh"
#include "../../microclimc_b200/csrc/mc_steady.cuh"
#include <vector>

using namespace mc;
static RunCfg mkcfg(const double* c) {
  RunCfg r; r.timestep = c[0]; r.edgedist = c[1]; r.reqhgt = c[2]; r.zu = c[3]; r.merid = c[4]; r.dst = c[5]; r.n = c[6];
  r.metopen = c[7] != 0; r.windhgt = c[8]; r.zlafact = c[9]; r.surfwet = c[10]; r.dbg = 0;
  return r;
}
extern "C" int dbg_run_transient2(int64_t ncol, int m, int sm, int64_t T, const double* vegp, const double* soilp, const double* forcing,
                      const double* fscale, const double* foffset, const double* tsoil_in, const double* lat,
                      const double* lon, const double* cfg, const double* pai_season, double* state, double* series,
                      const int64_t* samp_idx, int64_t nsamp, double* stats, double* fullseries, int32_t* flags, int,
                      const double* vegt, int64_t nvt, const double* thetaz);
extern "C" int dbg_run_transient(int64_t ncol, int m, int sm, int64_t T, const double* vegp, const double* soilp, const double* forcing,
                      const double* fscale, const double* foffset, const double* tsoil_in, const double* lat,
                      const double* lon, const double* cfg, const double* pai_season, double* state, double* series,
                      const int64_t* samp_idx, int64_t nsamp, double* stats, double* fullseries, int32_t* flags, int) {
  return dbg_run_transient2(ncol, m, sm, T, vegp, soilp, forcing, fscale, foffset, tsoil_in, lat, lon, cfg, pai_season, state, series, samp_idx,
                            nsamp, stats, fullseries, flags, 0, nullptr, 0, nullptr);
}
extern "C" int dbg_run_transient2(int64_t ncol, int m, int sm, int64_t T, const double* vegp, const double* soilp, const double* forcing,
                      const double* fscale, const double* foffset, const double* tsoil_in, const double* lat,
                      const double* lon, const double* cfg, const double* pai_season, double* state, double* series,
                      const int64_t* samp_idx, int64_t nsamp, double* stats, double* fullseries, int32_t* flags, int,
                      const double* vegt, int64_t nvt, const double* thetaz) {
  TransientArgs a{};
  a.vegt = vegt; a.nvt = nvt; a.thetaz = thetaz;
  a.ncol = ncol; a.m = m; a.sm = sm; a.T = T; a.t0 = 0; a.t1 = T; a.vegp = vegp; a.soilp = soilp; a.lat = lat; a.lon = lon;
  a.forcing = forcing; a.fscale = fscale; a.foffset = foffset; a.tsoil = tsoil_in; a.season = pai_season; a.state = state;
  a.series = series; a.full = fullseries; a.nsamp = nsamp; a.stats = stats; a.flags = flags; a.cfg = mkcfg(cfg);
  a.spin_steps = (int)cfg[11]; a.charmode = (int)cfg[12];
  long long* samp_of = new long long[ncol];
  for (int64_t c = 0; c < ncol; ++c) samp_of[c] = -1;
  for (int64_t j = 0; j < nsamp; ++j) samp_of[samp_idx[j]] = j;
  a.samp_of = samp_of;
  for (int64_t c = 0; c < ncol; ++c) {
    if (m <= 20 && sm <= 10) run_column<20, 10>(a, c); else run_column<64, 32>(a, c);
  }
  delete[] samp_of;
  return 0;
}
// Streaming form (mc_fast.cuh) with the direct-memory accessor: geometry pass, then every column through
// fast_run_column — spin-up launch first, then the hourly loop, exactly as mc_lib.cu sequences the two kernels.
extern "C" int dbg_run_transient_fast(int64_t ncol, int m, int sm, int64_t T, const double* vegp, const double* soilp, const double* forcing,
                      const double* fscale, const double* foffset, const double* tsoil_in, const double* lat,
                      const double* lon, const double* cfg, const double* pai_season, double* state, double* series,
                      const int64_t* samp_idx, int64_t nsamp, double* stats, double* fullseries, int32_t* flags, int) {
  if (pai_season || m > 20 || m <= kNG || sm > 10) return -9;
  TransientArgs t{};
  t.ncol = ncol; t.m = m; t.sm = sm; t.T = T; t.t0 = 0; t.t1 = T; t.vegp = vegp; t.soilp = soilp; t.lat = lat; t.lon = lon;
  t.forcing = forcing; t.fscale = fscale; t.foffset = foffset; t.tsoil = tsoil_in; t.season = nullptr; t.state = state;
  t.series = series; t.full = fullseries; t.nsamp = nsamp; t.stats = stats; t.flags = flags; t.cfg = mkcfg(cfg);
  t.spin_steps = (int)cfg[11]; t.charmode = (int)cfg[12];
  std::vector<long long> samp_of(ncol, -1);
  for (int64_t j = 0; j < nsamp; ++j) samp_of[samp_idx[j]] = j;
  t.samp_of = samp_of.data();
  FastLayout L{m, sm};
  const long long ntile = (ncol + kTile - 1) / kTile;
  std::vector<double> P(ntile * L.nP() * kTile), S(ntile * L.nS() * kTile), W(ntile * L.nW() * kTile);
  std::vector<int> fail(ncol, 0); std::vector<long long> flist(2 * ncol, 0); unsigned fcount[2] = {0, 0};
  FastArgs a{t, P.data(), S.data(), W.data(), nullptr, 0, nullptr, fail.data(), flist.data(), fcount};
  for (int spin = 1; spin >= 0; --spin) {
    if (spin ? t.spin_steps <= 0 : T <= 0) continue;
    a.spin_mode = spin;
    for (long long c = 0; c < ntile * kTile; ++c) {
      const long long cl = c < ncol ? c : ncol - 1;
      geom_column<20, 10>(a, effective_cfg(t, vegp[(long long)mc::V_HGT * ncol + cl], spin != 0), c, cl);
    }
    for (long long c = 0; c < ncol; ++c) {
      DirectMem mem;
      mem.L = L;
      mem.P = P.data() + (c / kTile) * L.nP() * kTile + (c % kTile);
      mem.S = S.data() + (c / kTile) * L.nS() * kTile + (c % kTile);
      mem.W = W.data() + (c / kTile) * L.nW() * kTile + (c % kTile);
      mem.ia = mem.ib = 1;
      // the instantiation mc_lib.cu would launch: REG (regular node grid, reqhgt NA) only with m = 20, sm = 10
      // (MC_HOST_FORCE_REG=1: the regular-grid instantiation at any time step, to exercise its fallback for leaftemp's transient balances)
      const char* fr = getenv("MC_HOST_FORCE_REG");
      TransientArgs tq = t; if (fr && fr[0] == '1') tq.cfg.timestep = 3600.0;
      if (m == 20 && sm == 10 && fast_regular_grid(tq, spin != 0)) fast_run_column<DirectMem, 20, 10, true>(a, mem, c, true);
      else fast_run_column<DirectMem, 20, 10, false>(a, mem, c, true);
    }
  }
  // columns that left the streaming form: redone by the general column from the call's initial state, as mc_lib.cu sequences it
  for (int q = 0; q < 2; ++q) {
    if (q == 0 ? t.spin_steps <= 0 : T <= 0) continue;
    TransientArgs g = t;
    if (q == 1) g.spin_steps = 0;
    g.flag_or = 8;
    for (unsigned k = 0; k < fcount[q]; ++k) run_column<20, 10>(g, flist[(q ? ncol : 0) + k]);
  }
  return 0;
}
extern "C" long long dbg_general_steps(int reset) { long long v = g_mc_general_steps; if (reset) g_mc_general_steps = 0; return v; }
extern "C" int dbg_runonestep(int64_t ncol, int m, int sm, const double* vegp, const double* soilp, const double* climvars,
                   const double* lat, const double* lon, const double* cfg, double jd, double lt, const double* theta,
                   const double* thetap, double* state, int* nmerged) {
  OneStepArgs a{}; a.ncol = ncol; a.m = m; a.sm = sm; a.vegp = vegp; a.soilp = soilp; a.lat = lat; a.lon = lon; a.climvars = climvars;
  a.theta = theta; a.thetap = thetap; a.jd = jd; a.lt = lt; a.state = state; a.nmerged = nmerged; a.cfg = mkcfg(cfg);
  for (int64_t c = 0; c < ncol; ++c) { if (m <= 20 && sm <= 10) onestep_column<20, 10>(a, c); else onestep_column<64, 32>(a, c); }
  return 0;
}

extern "C" int dbg_run_steady2(int64_t ncol, int m, int sm, int64_t T, const double* vegp, const double* soilp, const double* clim,
                   const double* nmr, int nmr_percol, const double* lat, const double* lon, const double* scfg,
                   const double* pai_season, double* out, double* stats, int, const double* vegt, int64_t nvt);
extern "C" int dbg_run_steady(int64_t ncol, int m, int sm, int64_t T, const double* vegp, const double* soilp, const double* clim,
                   const double* nmr, int nmr_percol, const double* lat, const double* lon, const double* scfg,
                   const double* pai_season, double* out, double* stats, int) {
  return dbg_run_steady2(ncol, m, sm, T, vegp, soilp, clim, nmr, nmr_percol, lat, lon, scfg, pai_season, out, stats, 0, nullptr, 0);
}
extern "C" int dbg_run_steady2(int64_t ncol, int m, int sm, int64_t T, const double* vegp, const double* soilp, const double* clim,
                   const double* nmr, int nmr_percol, const double* lat, const double* lon, const double* scfg,
                   const double* pai_season, double* out, double* stats, int, const double* vegt, int64_t nvt) {
  SteadyArgs a{};
  a.vegt = vegt; a.nvt = nvt; a.ncol = ncol; a.m = m; a.sm = sm; a.T = T; a.vegp = vegp; a.soilp = soilp; a.lat = lat; a.lon = lon; a.clim = clim;
  a.nmr = nmr; a.nmr_percol = nmr_percol; a.season = pai_season; a.reqhgt = scfg[0]; a.metopen = scfg[1] != 0; a.windhgt = scfg[2];
  a.surfwet = scfg[3]; a.out = out;
  int* flags = new int[ncol](); a.flags = flags;
  double* part = new double[6 * ncol]; a.part = part;
  int rc = 0;
  for (int64_t c = 0; c < ncol; ++c) {
    if (m <= 20) steady_column<20>(a, c, 0, 1); else steady_column<64>(a, c, 0, 1);
    if (flags[c] & 4) rc = -2;
    if (stats) { for (int q = 0; q < 6; ++q) stats[q * ncol + c] = part[q * ncol + c]; stats[0 * ncol + c] /= (double)T; stats[3 * ncol + c] /= (double)T; }
  }
  delete[] flags; delete[] part;
  return rc;
}
