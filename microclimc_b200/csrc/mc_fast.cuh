// Streaming form of the transient step (runonestep, R/code.R:687-902): time-invariant PAI, depth-uniform soil moisture, any time step.
// Steps with up to kNG merged canopy air layers (hourly steps give 1-5) are solved in the idle stage buffers, steps with more in the
// lane's own arrays (the wide solve).  A column that leaves the envelope (lowest nodes under 0.12 m with the edge profile; a state whose
// previn$z is not the launch's node grid in a regular-grid or parameter-class launch; a transient leaf balance on the regular grid) is
// dropped from the streaming launch: the kernel records it in a fail list, stops storing for it, and the library re-runs the listed
// columns from the launch's initial state through the general kernel (k_transient, mc_step.cuh) right after — the general column is NOT
// part of this kernel (it used to be: 14 KB stack frame, and its call boundary pinned the register allocation of the hot loop).
// Otherwise a state whose node heights differ from the current geometry (first step after paraminit when reqhgt snaps a node,
// code.R:573 vs 719) stays in the streaming form: the leaf-geometry rows are built from previn$z for the first step and rebuilt in
// place from the current heights after it (fast_regeom).
//
// Why a second form: the general column keeps ~16 KB of thread-private state + hoisted geometry + work arrays and is
// bound by local-memory latency (profiles/README.md r01b: long_scoreboard 64 % of stalls, FP64 pipe 13 % busy).
// Here nothing per-layer lives in thread-private memory:
//   * a geometry kernel writes every time-invariant per-layer constant once per launch into a tile-major workspace
//     P[tile][row][128 columns];  state S and the few inter-pass work rows W use the same tiling;
//   * the step is three sweeps over the layers — A: top -> bottom (wind, turbulent conductances, code.R:747-770),
//     B: bottom -> top (leafabs + leaftemp, code.R:776-785, with the layermerge / layeraverage sums accumulated on
//     the way), then the Thomas / ThomasV solves on the few merged nodes (code.R:813-860), C: bottom -> top
//     (layerinterp or the equilibrium blend, humidity, limits, code.R:846-897) — and each (sweep, layer) is one
//     pipeline stage whose rows (2-27 KB per 128-column tile) are
//     fetched by TMA 1-D bulk copies (cp.async.bulk ... mbarrier::complete_tx) into a double-buffered shared-memory
//     stage while the previous layer is being computed; results go back with coalesced 8-byte stores;
//   * the block's four warps advance stage by stage together, so one instruction-cache fill serves all of them.
// The same code builds for the host (tests/hostdbg) with a direct-memory accessor, which is how its arithmetic is
// checked against the CPU oracle without a GPU.
#pragma once
#include <type_traits>
#include "mc_driver.cuh"

namespace mc {

constexpr int kTile = 128;   // columns per CTA == columns per workspace tile
// Algebraic forms that keep transcendental work off the per-layer chains (each changes results at rounding level only):
//   * sweep A carries log(uh): windcanopy's two products uh * exp(a ex), ur * exp(a ed) and their maximum become two sums and a maximum,
//     the gcanopy quotient uh / (e0 - e1) becomes 1 / (exp(-a gx0 - lu) - exp(-a gx1 - lu)): 3 exponentials per layer instead of 6 (the gl
//     rows hold log(gl));
//   * leaf boundary-layer conductance with the common factor (ph / d) sqrt(sc) of the forced and free branches taken out of the maximum
//     (gha_fast2): 4 square roots and no reciprocal instead of 5 and one;
//   * the working state keeps the series conductance gv gha / (gv + gha) of the previous step (all that leaftemp reads of previn$gv)
//     instead of gv; gv itself is written with the harvest rows when the state is emitted.
// Same-box A/B of the three together: 589-594 -> 613-614 M column-steps/s.
#if !defined(__CUDACC__)
static long long g_mc_general_steps = 0;     // host (test) build: columns that left the streaming form (launch-level count)
#endif

// ---- workspace rows -----------------------------------------------------------------------------------------------
// Scalar rows of a column.  The first 47 are grouped by the phase that reads them, each group contiguous, so that ONE TMA bulk copy
// brings a group into stage-buffer rows that are idle at that time, a phase ahead of its use (the reads used to be ~20 exposed L2 / DRAM
// round trips per step):
//   G0  (26 rows) forcing perturbation scale / offset [8 + 8] and the phase-0 constants: fetched during the end of the previous step;
//   G1a ( 9 rows) canopy-top conductance and soil-conductivity constants: fetched during sweep A, read right after it;
//   G1b (12 rows) radiation geometry and soil-humidity constants of the sweep-B scalars: fetched during sweep A.
enum { P0_FS0, P0_FO0 = P0_FS0 + 8,
       P0_HGT = P0_FO0 + 8, P0_LA, P0_LNREF, P0_D, P0_DZTB, P0_LZU, P0_LZUH, P0_LZOH, P0_GLNH, P0_LNHD, P0_G0END,
       P0_A0CAP = P0_G0END, P0_GXC0, P0_CGCAP, P0_RDZCAP, P0_SKC, P0_SKA, P0_SKB, P0_SKAD, P0_CVDRY, P0_G1AEND,
       P0_LON = P0_G1AEND, P0_SLAT, P0_CLAT, P0_X, P0_KDEN, P0_K5, P0_SUMPAI, P0_SM, P0_TRDM, P0_SOILB, P0_PSIE, P0_SMAX, P0_G1BEND,
       P0_WL1 = P0_G1BEND, P0_WL2, P0_A0G, P0_GXG1, P0_CGG, P0_RZG,
       P0_CLUMP, P0_REFG, P0_EMV, P0_VEGEM, P0_TRLWSUM, P0_REFM, P0_LAT,
       P0_GSMAX, P0_Q50, P0_LWD, P0_Z1,
       P0_SELG, P0_SELH, P0_SELS, P0_ZABOVE, P0_FAST, P0_CPW, P0_PHW, P0_ZDIFF, P0_NFIX };          // then 1/DZK[sm], DZC[sm], Z[m]
constexpr int kG0Rows = P0_G0END, kG1aRows = P0_G1AEND - P0_G0END, kG1bRows = P0_G1BEND - P0_G1AEND;
constexpr int kGRow0 = 15;                        // G1a / G1b land in rows 15.. of a stage buffer (sweep-A stages use rows 0..13)
// Per-layer rows.  The first PA_NREG / PB_NREG of a block are staged in every launch.  The next ones are pure functions of the node grid
// z: on the REGULAR grid z[i] = (i - 0.5) hgt / m that paraminit builds (code.R:573) and runonestep keeps whenever reqhgt is NA (code.R:719
// replaces one node otherwise) they depend on the column through hgt alone, and a launch whose columns all have that grid (template
// parameter REG of the kernel) forms them from m / hgt and small-integer reciprocals instead of streaming them: 5 + 4 rows of 8 bytes per
// layer and step less, 1.44 of the 8.4 KB the kernel moves per column-step.  PB_CVD (the cumulative sum of dz * vden, read at the merged-layer
// group boundaries by the solve), PB_MAC (leaftemp's transient air balance, which only the non-regular instantiation evaluates, and hourly steps
// take it only for a node moved above the canopy) and PB_VDEN (read by fast_regeom only) are never staged: direct reads where needed.
// (On the regular grid the inner-node rows GLI, EDI of layer i are those of the outer node of layer i + 1 — the same height, zi(i) = zo(i + 1)
// — up to the factor zi(i + 1) / zo(i) = (i + 2) / i in ed: sweep A, going down, carries them in registers and stages them only for
// layer m - 1, which has no such neighbour.  SREF = sqrt(1 - ref) stands for three former rows: 1 - ref = SREF^2, sref * pc = SREF * PC.)
enum { PA_A0I, PA_A0O, PA_GLO, PA_EDO, PA_CG, PA_NREG5, PA_GLI = PA_NREG5, PA_EDI, PA_NREG, PA_EX1 = PA_NREG, PA_EX2, PA_GX0, PA_GX1, PA_RDZ, PA_N };
enum { PB_SREF, PB_PC, PB_SGC, PB_TRDC, PB_TRDG, PB_TRLW, PB_PAI, PB_IZL, PB_MZ, PB_NREG, PB_IZTH = PB_NREG, PB_IZP, PB_IZRF, PB_DZ,
       PB_NSTG, PB_CVD = PB_NSTG, PB_MAC, PB_VDEN, PB_N };
// 1 / k for the small integers the regular grid needs (k <= 2 m - 1, m <= 20 in the streaming form)
#if defined(__CUDACC__)
__constant__ double c_rinv[40] = {0.0,     1.0 / 1,  1.0 / 2,  1.0 / 3,  1.0 / 4,  1.0 / 5,  1.0 / 6,  1.0 / 7,  1.0 / 8,  1.0 / 9,
                                  1.0 / 10, 1.0 / 11, 1.0 / 12, 1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19,
                                  1.0 / 20, 1.0 / 21, 1.0 / 22, 1.0 / 23, 1.0 / 24, 1.0 / 25, 1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29,
                                  1.0 / 30, 1.0 / 31, 1.0 / 32, 1.0 / 33, 1.0 / 34, 1.0 / 35, 1.0 / 36, 1.0 / 37, 1.0 / 38, 1.0 / 39};
#endif
MC_HD double rinv_small(int k) {
#if defined(__CUDA_ARCH__)
  return c_rinv[k];
#else
  return 1.0 / (double)k;
#endif
}
enum { S0_TABOVE, S0_RELHUM, S0_TAIR, S0_TSOIL, S0_PK, S0_H, S0_PSIM, S0_G, S0_TCSUM, S0_GTCAP, S0_N };   // then soiltc[sm]
enum { SL_TC, SL_TLEAF, SL_RH, SL_RABS, SL_GV, SL_GHA, SL_N };
enum { WL_GTN, WL_UZN, WL_TT, WL_EAL, WL_L, WL_RSWIN, WL_RLWIN, WL_GVN, WL_N };   // GVN: previn$gv, written with the harvest rows
constexpr int kWLiveB = 2;                        // W rows sweep B reads back from sweep A (GTN, UZN)
constexpr int kWLiveC = 2;                        // W rows sweep C reads back from sweep B (TT, EAL)
constexpr int kCL = 5;                            // layers per sweep-C stage (their TT/EAL rows are contiguous: 10 rows)
static_assert(WL_GTN == 0 && WL_UZN == 1, "sweep B stages the first kWLiveB rows of a layer's block");
constexpr int kStageRowsB = PB_NSTG + SL_N + kWLiveB + 1;   // sweep-B stage: geometry, state, work rows and the old gt row = 23 rows (19 on the regular grid)
constexpr int kStageRows = 27;                    // rows of 128 doubles per stage buffer (the solver scratch needs 2 x 27)
static_assert(kStageRowsB <= kStageRows, "sweep-B stage exceeds the stage buffer");
static_assert(kG0Rows <= kStageRows && kGRow0 + kG1aRows <= kStageRows && kGRow0 + kG1bRows <= kStageRows && PA_N + 2 <= kGRow0, "scalar groups do not fit the idle stage rows");
constexpr int kG2Spill = kStageRowsB;              // first stage row no sweep-B stage uses
static_assert(3 * 10 <= kStageRows + (kStageRows - kG2Spill), "the solve's soil rows do not fit the idle stage rows");
constexpr int kNG = 6;                            // merged air layers the solve holds in the idle stage buffers (the common case: hourly steps give 1-5)
constexpr int kNGW = 20;                          // ... and in the lane's own (thread-private) scratch when a step has more: the wide solve

struct FastLayout {
  int m, sm;
  MC_HD int np0() const { return P0_NFIX + 2 * sm + m; }
  MC_HD int pz(int i) const { return P0_NFIX + 2 * sm + (i - 1); }
  MC_HD int pa(int i, int r) const { return np0() + (i - 1) * PA_N + r; }
  MC_HD int pb(int i, int r) const { return np0() + m * PA_N + (i - 1) * PB_N + r; }
  MC_HD int nP() const { return np0() + m * (PA_N + PB_N); }
  MC_HD int soiltc(int j) const { return S0_N + (j - 1); }
  MC_HD int sl(int i, int r) const { return S0_N + sm + (i - 1) * SL_N + r; }
  MC_HD int uz(int i) const { return S0_N + sm + m * SL_N + (i - 1); }
  MC_HD int gt(int i) const { return S0_N + sm + m * SL_N + m + (i - 1); }          // i = 1..m+1
  MC_HD int nS() const { return S0_N + sm + m * SL_N + m + (m + 1); }
  // per-layer blocks of 6 rows (GTN, UZN, L, RSWIN, RLWIN, GVN), then the sweep-C rows (TT, EAL) of all layers contiguously
  MC_HD int wl(int i, int r) const {
    return (r == WL_TT || r == WL_EAL) ? m * (WL_N - 2) + (i - 1) * 2 + (r - WL_TT) : (i - 1) * (WL_N - 2) + (r < WL_TT ? r : r - 2);
  }
  MC_HD int nW() const { return m * WL_N; }
};

struct FastArgs {
  TransientArgs t;          // the ABI-layout arrays and run description (mc_driver.cuh)
  double* P;                // [ntile][nP][128]  geometry, written by geom_column
  double* S;                // [ntile][nS][128]  working state
  double* W;                // [ntile][nW][128]  inter-pass work rows
  unsigned long long* dbgbuf;   // [8] phase cycle counters (MC_DBG_SKIP bit 16) or null
  int spin_mode;            // 1: paraminit + t.spin_steps steps on forcing row 0 with the running mean of H (spinup, code.R:982-1013)
  // columns that left the streaming form: fail[c] = 1 (during the spin-up launch) or 2 (during the run launch), appended to
  // flist[0][..] / flist[1][..] with counts fcount[0] / fcount[1]; re-run by the general kernel after the streaming launches
  // Parameter classes (mc_set_classes): tile t takes its per-LAYER geometry rows from tile tile_src[t] — the first tile of the launch whose
  // 128 columns all belong to the same class as t's (itself when its columns are mixed) — so a grid with K << ncol distinct vegetation /
  // soil parameter sets streams those 4.6 KB per column-step from L2 instead of DRAM.  Null: every tile reads its own rows.
  const int* tile_src;
  int* fail;                // [ncol], zeroed before the first launch of a call
  long long* flist;         // [2][ncol]
  unsigned* fcount;         // [2]
};

// The six leaftemp rows of a layer that depend on previn$z (code.R:387-392): 1/zth, 1/zla, 1/zref, 1/z, the two time-step factors.
struct LeafGeom { double izth, izl, izrf, izp, mac, mz; };
MC_HD LeafGeom leafgeom_rows(double hgt, double zp, double zth, double zl, double PAIi, double vden, double cpw, double phw, double timestep) {
  const double PAIm = 2 * PAIi / zth;
  LeafGeom g;
  g.izth = 1 / zth; g.izl = 1 / zl; g.izrf = 1 / ((hgt + 2) - zp); g.izp = 1 / zp;
  g.mac = (timestep * PAIm * 2) / (1 - vden);
  g.mz = ((timestep * PAIm) / (vden * cpw * phw)) / zl;
  return g;
}

// ---- geometry: one thread per column, once per launch (uses the general column's precompute) -------------------------
template <int MM, int MS>
MC_HD void geom_column(const FastArgs& a, const RunCfg& cfg, long long c, long long cl /* clamped source column */) {
  Column<MM, MS> col;
  const TransientArgs& t = a.t;
  col.load_params(t.vegp, t.soilp, t.lat, t.lon, t.ncol, cl, t.m, t.sm);
  col.precompute(cfg, 1.0);
  const int m = t.m, sm = t.sm;
  // previn$z of the launch's first step: paraminit's regular grid (code.R:573) for the spin-up launch, the z rows of the incoming
  // state otherwise.  When it differs from the current heights, the kernel rebuilds the leaf-geometry rows after that step.
  bool zdiff = false;
  for (int i = 1; i <= m; ++i) {
    col.zprev[i] = a.spin_mode ? (i - 0.5) / m * col.hgt : t.state[(long long)(P_NSCAL + 3 * m + i - 1) * t.ncol + cl];
    if (col.zprev[i] != col.z[i]) zdiff = true;
  }
  col.precompute_leafgeom(cfg);
  FastLayout L{m, sm};
  double* P = a.P + (c / kTile) * (long long)L.nP() * kTile + (c % kTile);
  auto put = [&](int row, double v) { P[(long long)row * kTile] = v; };
  const double hgt = col.hgt;
  for (int v = 0; v < kNPert; ++v) {              // forcing perturbation of the column (identity when the run has none)
    put(P0_FS0 + v, (t.fscale && t.foffset) ? t.fscale[(long long)v * t.ncol + cl] : 1.0);
    put(P0_FO0 + v, (t.fscale && t.foffset) ? t.foffset[(long long)v * t.ncol + cl] : 0.0);
  }
  put(P0_HGT, hgt); put(P0_LA, col.La); put(P0_LNREF, col.lnref); put(P0_LZU, col.lzu); put(P0_LZUH, col.lzuh); put(P0_LZOH, col.lzoh);
  put(P0_LNHD, col.lnhd); put(P0_D, col.d); put(P0_GLNH, col.g_lnh); put(P0_DZTB, col.z[m] - col.z[1]); put(P0_SUMPAI, col.sumPAI);
  {
    const double zm4 = roughlength_d(hgt, col.sumPAI, col.d, 0.004);
    put(P0_WL1, log((cfg.windhgt - col.d) / zm4)); put(P0_WL2, log(((hgt + 2) - col.d) / zm4));
  }
  put(P0_A0CAP, col.a0_cap); put(P0_GXC0, col.gxc0); put(P0_CGCAP, (col.lm_cap * col.iw[m]) / hgt); put(P0_RDZCAP, 1 / fabs(hgt - col.z[m]));
  put(P0_A0G, col.a0_g); put(P0_GXG1, col.gxg1); put(P0_CGG, (col.lm_g * col.iwmean) / hgt); put(P0_RZG, 1 / fabs(col.zg - 0));
  put(P0_K5, cank(col.x, 5.0, col.kden));
  { double sl, cl; sol_site(col.lat, sl, cl); put(P0_SLAT, sl); put(P0_CLAT, cl); }
  put(P0_X, col.x); put(P0_KDEN, col.kden); put(P0_CLUMP, col.clump); put(P0_REFG, col.refg); put(P0_SM, col.s_m); put(P0_TRDM, col.trd_m);
  put(P0_EMV, col.emv); put(P0_VEGEM, col.vegem); put(P0_TRLWSUM, col.trlw_sum);
  put(P0_REFM, col.pLAI[m] * col.refls + (1 - col.pLAI[m]) * col.refw); put(P0_LAT, col.lat); put(P0_LON, col.lon);
  put(P0_GSMAX, col.gsmax); put(P0_Q50, col.q50); put(P0_LWD, col.lw * 0.71); put(P0_SMAX, col.Smax); put(P0_SOILB, col.soilb);
  put(P0_PSIE, col.psi_e);
  {                                                                  // soilk (see Column::soilk): theta-independent pieces
    const double rho = col.rho;
    const double A = 0.65 - 0.78 * rho + 0.6 * rho * rho, B = 1.06 * rho, C = 1 + 2.6 / sqrt(col.Mc), D = 0.03 + 0.1 * rho * rho;
    put(P0_SKA, A); put(P0_SKB, B); put(P0_SKC, C); put(P0_SKAD, A - D);
    put(P0_CVDRY, 2.128 * col.Vq + 2.385 * col.Vm + 2.496 * col.Vo_s);
    for (int j = 1; j <= sm; ++j) {
      const double dzk = (j < sm) ? (col.zs[j + 1] - col.zs[j]) : (col.zs[sm] - col.zs[sm - 1]);
      const double zlo = (j > 1) ? col.zs[j - 1] : 0.0;
      const double zhi = (j < sm) ? col.zs[j + 1] : (col.zs[sm] + (col.zs[sm] - col.zs[sm - 1]));
      put(P0_NFIX + (j - 1), 1 / dzk); put(P0_NFIX + sm + (j - 1), zhi - zlo);
    }
  }
  put(P0_Z1, col.z[1]); put(P0_SELG, (double)col.selG);
  for (int i = 1; i <= m; ++i) put(L.pz(i), col.z[i]);
  {                                                                  // output node selection of runmodel (code.R:1241-1263)
    int selh = 0;
    if (cfg.reqhgt == cfg.reqhgt) for (int i = 1; i <= m; ++i) if (col.z[i] == cfg.reqhgt) { selh = i; break; }
    int sels = 1; double mn = 1e300;
    if (cfg.reqhgt == cfg.reqhgt) for (int j = 1; j <= sm; ++j) { double dd = fabs(col.zs[j] + cfg.reqhgt); if (dd < mn) { mn = dd; sels = j; } }
    put(P0_SELH, (double)selh); put(P0_SELS, (double)sels);
  }
  put(P0_ZABOVE, col.zabove_g);
  put(P0_CPW, col.cpw); put(P0_PHW, col.phw); put(P0_ZDIFF, zdiff ? 1.0 : 0.0);
  bool fast = true;
  if (cfg.edgedist == cfg.edgedist) for (int i = 1; i <= m; ++i) if (col.zi_[i] == 0.0) fast = false;
  put(P0_FAST, fast ? 1.0 : 0.0);
  // sweep A rows: layer m (canopy top, code.R:749-753) then the recurrence layers (code.R:756-767)
  put(L.pa(m, PA_A0I), col.a0_top); put(L.pa(m, PA_A0O), col.a0_top); put(L.pa(m, PA_EX1), col.ex_top); put(L.pa(m, PA_EX2), col.ex_top);
  put(L.pa(m, PA_GLO), log(col.gl_top)); put(L.pa(m, PA_GLI), log(col.gl_top)); put(L.pa(m, PA_EDO), col.ed_top); put(L.pa(m, PA_EDI), col.ed_top);
  put(L.pa(m, PA_GX0), col.gx0_top); put(L.pa(m, PA_GX1), col.gx1_top); put(L.pa(m, PA_CG), (col.lm_top * col.iw[m]) / hgt);
  put(L.pa(m, PA_RDZ), 1 / fabs(col.z[m] - col.z[m - 1]));
  for (int i = m - 1; i >= 1; --i) {
    const double zt = col.z[i + 1] - col.z[i], zi = col.z[i + 1] + 0.5 * zt;
    const double z0 = (i > 1) ? col.z[i - 1] : 0;
    put(L.pa(i, PA_A0I), col.a0i[i]); put(L.pa(i, PA_A0O), col.a0o[i]); put(L.pa(i, PA_EX1), col.ex1[i]); put(L.pa(i, PA_EX2), col.ex2[i]);
    put(L.pa(i, PA_GLO), log(col.gl_o[i])); put(L.pa(i, PA_GLI), log(col.gl_i[i])); put(L.pa(i, PA_EDO), col.ed_o[i]); put(L.pa(i, PA_EDI), col.ed_i[i]);
    put(L.pa(i, PA_GX0), col.gx0[i]); put(L.pa(i, PA_GX1), col.gx1[i]); put(L.pa(i, PA_CG), (col.lmi[i] * col.iw[i]) / zi);
    put(L.pa(i, PA_RDZ), 1 / fabs(col.z[i] - z0));
  }
  // sweep B rows: leafabs (code.R:31-49) and leaftemp (code.R:378-512) constants of the layer
  double cvd = 0;
  for (int i = 1; i <= m; ++i) {
    const double ref = col.pLAI[i] * col.refls + (1 - col.pLAI[i]) * col.refw;
    const double pc = col.PAIg[m + 1 - i];
    const double vden = col.thickw[i] * col.PAI[i];
    const LeafGeom lg = leafgeom_rows(hgt, col.zprev[i], col.zthp[i], col.zla[i], col.PAI[i], vden, col.cpw, col.phw, cfg.timestep);
    put(L.pb(i, PB_SREF), col.sref[i]); put(L.pb(i, PB_PC), pc); put(L.pb(i, PB_SGC), col.sref[i] * col.PAIg[i]);      // sref = sqrt(1 - ref)
    put(L.pb(i, PB_TRDC), col.trdc[i]); put(L.pb(i, PB_TRDG), col.trdg[i]); put(L.pb(i, PB_TRLW), col.trlw[i]); put(L.pb(i, PB_PAI), col.PAI[i]);
    put(L.pb(i, PB_IZTH), lg.izth); put(L.pb(i, PB_IZL), lg.izl); put(L.pb(i, PB_IZRF), lg.izrf);
    put(L.pb(i, PB_IZP), lg.izp); put(L.pb(i, PB_MAC), lg.mac);
    put(L.pb(i, PB_MZ), lg.mz);
    const double dzi = col.z[i] - (i > 1 ? col.z[i - 1] : 0.0);
    cvd += dzi * vden;                                                 // sum over the layers up to i of dz * vden (layeraverage's vden sum of a group = a difference of two)
    put(L.pb(i, PB_DZ), dzi); put(L.pb(i, PB_CVD), cvd); put(L.pb(i, PB_VDEN), vden);
  }
}

// ---- memory accessors ------------------------------------------------------------------------------------------------
// DirectMem: host build and the device's non-pipelined reads. Every accessor addresses this thread's lane of a tile.
struct DirectMem {
  const double* P; double* S; double* W;      // lane-adjusted tile bases
  FastLayout L;
  int ia, ib;
  double parked[64];                            // stands in for the device's tensor-memory scratch (PipeMem::park)
  MC_HD void park(int slot, double v) { parked[slot] = v; }
  MC_HD void park_done() {}
  MC_HD void unpark8(int slot0, double* v) { for (int q = 0; q < 8; ++q) v[q] = parked[slot0 + q]; }
  MC_HD double unpark(int slot) { return parked[slot]; }
  double scratch[2 * kStageRows];               // stands in for the idle stage buffers (PipeMem::scr_*)
  MC_HD double scr_get(int idx) const { return scratch[idx]; }
  MC_HD void scr_put(int idx, double v) { scratch[idx] = v; }
  MC_HD int warp_max(int v) const { return v; }
  MC_HD void st_stream(double* p, double v) const { *p = v; }
  MC_HD void st_keep(double* p, double v) const { *p = v; }
  MC_HD double ld_keep(const double* p) const { return *p; }
  MC_HD void discard_AB() {}
  MC_HD void discard_C() {}
  MC_HD void g0_issue() {}
  MC_HD void g0_wait() {}
  MC_HD void g1_wait() {}
  MC_HD void g2_wait() {}
  MC_HD double g2(int r) const { return r < 2 * L.sm ? P[(long long)(P0_NFIX + r) * kTile] : S[(long long)L.soiltc(r - 2 * L.sm + 1) * kTile]; }
  MC_HD void end_begin(bool) {}
  MC_HD double g0(int r) const { return P[(long long)r * kTile]; }
  MC_HD double g1a(int r) const { return P[(long long)r * kTile]; }
  MC_HD double g1b(int r) const { return P[(long long)r * kTile]; }
  MC_HD void step_begin() {}
  MC_HD void sweepB_begin() {}
  MC_HD void sweepC_begin() {}
  MC_HD void stageA(int i) { ia = i; }
  MC_HD void stageB(int i) { ib = i; }
  MC_HD void stageC(int i) { ib = i; }
  MC_HD double c_w(int r) const { return W[(long long)L.wl(ib, WL_TT + r) * kTile]; }
  MC_HD double c_wq(int q, int r) const { return W[(long long)L.wl(ib + q, WL_TT + r) * kTile]; }   // layer ib + q of the current sweep-C stage
  MC_HD double a_p(int r) const { return P[(long long)L.pa(ia, r) * kTile]; }
  MC_HD double a_tcl() const { return S[(long long)L.sl(ia - 1, SL_TC) * kTile]; }          // only for ia > 1
  MC_HD double a_gto() const { return S[(long long)L.gt(ia) * kTile]; }
  MC_HD double b_p(int r) const { return P[(long long)L.pb(ib, r) * kTile]; }
  MC_HD double b_s(int r) const { return S[(long long)L.sl(ib, r) * kTile]; }
  MC_HD double b_w(int r) const { return W[(long long)L.wl(ib, r) * kTile]; }
  MC_HD double b_gto() const { return S[(long long)L.gt(ib) * kTile]; }                     // read before the layer stores its new gt
  MC_HD double p_cvd(int i) const { return P[(long long)L.pb(i, PB_CVD) * kTile]; }
  MC_HD double p_mac(int i) const { return P[(long long)L.pb(i, PB_MAC) * kTile]; }
};

// make BOUNDS=1: every computed index into the stage buffers, the tensor-memory scratch and the tile workspaces is range-checked on the
// device (printf + trap).  compute-sanitizer is not available on the GPU pool (profiles/r02a_compute_sanitizer_refused.txt); this build is
// what the memory-safety record of profiles/ was produced with.  Never the shipped build (the checks cost ~15 %).
#if defined(MC_BOUNDS) && defined(__CUDA_ARCH__)
#define MC_CHK(cond, what, v) do { if (!(cond)) { printf("[mc bounds] %s = %lld out of range (block %d thread %d)\n", what, (long long)(v), blockIdx.x, threadIdx.x); __trap(); } } while (0)
#else
#define MC_CHK(cond, what, v) ((void)0)
#endif
#if defined(__CUDACC__)
#define MC_D __device__ __forceinline__
// PipeMem: per-stage rows arrive in shared memory through TMA bulk copies; two buffers, one mbarrier each.
MC_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
MC_D void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
MC_D void mbar_expect_tx(unsigned bar32, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar32), "r"(bytes) : "memory");
}
MC_D void mbar_wait(unsigned bar32, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar32), "r"(parity) : "memory");
}
MC_D void bulk_g2s(unsigned dst32, const void* src, unsigned bytes, unsigned bar32) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst32), "l"(src), "r"(bytes),
               "r"(bar32)
               : "memory");
}
// streamed rows (geometry, state) are read once per step with a reuse distance far beyond L2: mark them evict-first so
// that the short-lived work rows and the threads' local scratch stay resident
// (The descriptors createpolicy.fractional.L2::evict_first / evict_last return for fraction 1.0 are fixed encodings — the same constants
// CUTLASS passes as TMA cache hints, cute/arch/copy_sm90_desc.hpp — so they are immediates here and occupy no registers between uses.)
constexpr unsigned long long kL2EvictFirst = 0x12F0000000000000ull, kL2EvictLast = 0x14F0000000000000ull;
MC_D unsigned long long l2_evict_first_policy() { return kL2EvictFirst; }
MC_D void bulk_g2s_stream(unsigned dst32, const void* src, unsigned bytes, unsigned bar32, unsigned long long pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst32), "l"(src),
               "r"(bytes), "r"(bar32), "l"(pol)
               : "memory");
}
MC_D void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Work rows W are written in one sweep and read back once, half a step later, by the same CTA: keep them in L2 (evict-last) and drop
// the dead lines after the read (discard.global.L2: no write-back of a dirty line nobody will read again).  MC_L2HINT: 0 none,
// 1 hints only, 2 hints + discard.
#ifndef MC_L2HINT
#define MC_L2HINT 2
#endif
MC_D unsigned long long l2_evict_last_policy() { return kL2EvictLast; }
MC_D void st_hint(double* p, double v, unsigned long long pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
MC_D double ld_hint(const double* p, unsigned long long pol) {
  double v;
  asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
MC_D void discard_line(const void* p) { asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory"); }

template <bool REG>
struct PipeMemT {
  static constexpr int PBS = REG ? PB_NREG : PB_NSTG;   // geometry rows of a sweep-B stage
  // geometry rows of the sweep-A stage of layer i: on the regular grid GLI / EDI are staged for layer m - 1 only (see the PA enum)
  MC_D int pas(int i) const { return REG ? (i == L.m - 1 ? (int)PA_NREG : (int)PA_NREG5) : (int)PA_N; }
  int an;                                      // pas() of the current sweep-A stage
  const double* P; double* S; double* W;      // lane-adjusted tile bases (for stores and the few direct reads)
  const double* Pt; const double* St; const double* Wt;   // tile bases (lane 0), for the bulk copies
  const double* Pl;                            // tile whose per-layer geometry rows this tile reads (Pt, or its class's first tile)
  const double* Pll;                           // Pl + lane, for the rows that are read directly
  MC_D double p_cvd(int i) const { return Pll[(long long)L.pb(i, PB_CVD) * kTile]; }
  MC_D double p_mac(int i) const { return Pll[(long long)L.pb(i, PB_MAC) * kTile]; }
  FastLayout L;
  double* buf;                                 // [2][kStageRows][kTile] in shared memory
  unsigned long long* bar;                     // [2]
  unsigned bar32, buf32;                       // the same two as 32-bit shared-window addresses (no generic->shared conversion in the loops)
  int lane;                                    // threadIdx.x
  unsigned k;                                  // stages issued so far (uniform over the block)
  const double* cur;                           // current stage buffer + lane
  const double* cbase;                         // sweep C: start of the current group's rows
  static constexpr unsigned long long pol = kL2EvictFirst;        // L2 evict-first policy for the streamed rows
  static constexpr unsigned long long pol_keep = kL2EvictLast;    // L2 evict-last policy for the work rows
#if MC_L2HINT
  MC_D void st_stream(double* p, double v) const { st_hint(p, v, pol); }
  MC_D void st_keep(double* p, double v) const { st_hint(p, v, pol_keep); }
#else
  MC_D void st_stream(double* p, double v) const { *p = v; }
  MC_D void st_keep(double* p, double v) const { *p = v; }
#endif
#ifndef MC_KEEP
#define MC_KEEP 0          // bit 0: per-column forcing perturbation rows, bit 1: the state's scalar rows — loaded / stored with the L2 evict-last policy
#endif
  MC_D double ld_keep(const double* p) const { return ld_hint(p, pol_keep); }
  // all lines of the A->B rows (GTN, UZN of every layer) / of the B->C rows (TT, EAL, contiguous) of this tile
  MC_D void discard_AB() {
#if MC_L2HINT >= 2
    const int nl = L.m * 2 * (kTile * 8 / 128);                        // 2 rows per layer, 8 lines per row
    for (int q = lane; q < nl; q += kTile) {
      const int i = q / 16, r = q % 16;                                // layer, line within its two (adjacent) rows
      discard_line(reinterpret_cast<const char*>(Wt + (long long)L.wl(i + 1, WL_GTN) * kTile) + r * 128);
    }
#endif
  }
  MC_D void discard_C() {
#if MC_L2HINT >= 2
    const int nl = L.m * kWLiveC * (kTile * 8 / 128);
    const char* base = reinterpret_cast<const char*>(Wt + (long long)L.wl(1, WL_TT) * kTile);
    for (int q = lane; q < nl; q += kTile) discard_line(base + q * 128);
#endif
  }
  // Tensor memory as a per-thread scratchpad: the CTA owns 128 columns x 128 lanes x 32 bit, i.e. 64 doubles per thread
  // (thread t of warp w <-> lane 32*(w%4) + t%32).  Step scalars that are only needed once per layer are parked there instead
  // of being spilled to local memory (12-cycle LDTM vs an L2 round trip).  All 32 lanes must execute these together.
  unsigned tbase;                              // TMEM address of this warp's lane quarter, column 0 of the allocation
  long long tcyc[8], tlast;                    // phase timing (MC_DBG_SKIP bit 16)
  MC_D void park(int slot, double v) {
    MC_CHK(slot >= 0 && slot < 64, "tensor-memory slot (park)", slot);
    __syncwarp();
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(tbase + 2 * slot), "r"(__double2loint(v)), "r"(__double2hiint(v)));
  }
  MC_D void park_done() { asm volatile("tcgen05.wait::st.sync.aligned;"); }
  // per-thread scratch in the stage buffers while no stage is in flight (54 doubles: row idx of the two buffers, own lane)
  MC_D double scr_get(int idx) const { MC_CHK(idx >= 0 && idx < 2 * kStageRows, "solver scratch row (get)", idx); return buf[idx * kTile + lane]; }
  MC_D void scr_put(int idx, double v) { MC_CHK(idx >= 0 && idx < 2 * kStageRows, "solver scratch row (put)", idx); buf[idx * kTile + lane] = v; }
  MC_D int warp_max(int v) const { __syncwarp(); return __reduce_max_sync(0xffffffffu, v); }
  MC_D double unpark(int slot) {
    MC_CHK(slot >= 0 && slot < 64, "tensor-memory slot (unpark)", slot);
    unsigned lo, hi;
    __syncwarp();
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(tbase + 2 * slot));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    return __hiloint2double(hi, lo);
  }
  MC_D void unpark8(int slot0, double* v) {
    MC_CHK(slot0 >= 0 && slot0 + 8 <= 64, "tensor-memory slot (unpark8)", slot0);
    unsigned r[16];
    __syncwarp();
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tbase + 2 * slot0));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __hiloint2double(r[2 * q + 1], r[2 * q]);
  }
  // ---- scalar groups (see the P0 enum): one more mbarrier (bar[2]), used strictly in turn: G0 of a step, then its G1, then the next G0
  unsigned gk;                                 // completed group waits (parity of bar[2])
  const double* g0p; const double* g1ap; const double* g1bp;   // lane-adjusted smem rows of the current groups
#ifndef MC_GKEEP
#define MC_GKEEP 0
#endif
  MC_D void g_issue(unsigned dst32, int row0, int nrows) { bulk_g2s_stream(dst32, Pt + (long long)row0 * kTile, nrows * kTile * 8, bar32 + 16, MC_GKEEP ? pol_keep : pol); }
  // G0 for the coming step goes to the buffer its first sweep-A stage does NOT use (stage k lands in buffer k & 1)
  MC_D void g0_issue() {
    if (lane == 0) {
      mbar_expect_tx(bar32 + 16, kG0Rows * kTile * 8);
      g_issue(buf32 + ((k + 1) & 1) * (kStageRows * kTile * 8), 0, kG0Rows);
    }
    g0p = buf + ((k + 1) & 1) * (kStageRows * kTile) + lane;
  }
  MC_D void g_wait() { mbar_wait(bar32 + 16, gk & 1); ++gk; }
  MC_D void g0_wait() { g_wait(); }
  MC_D void g1_wait() { g_wait(); }
  // G2: the solve's soil rows — 1 / dzk [sm], dzc [sm] of the geometry and the sm soil temperatures of the state — fetched behind the LAST
  // sweep-B stage, which has no successor to prefetch: the other stage buffer is idle from that stage's barrier until sweep C's first
  // copy.  Rows 0.. of that buffer, and what does not fit its kStageRows rows goes to the spare rows (kG2Spill..) of the current buffer.
  // (They used to be 3 sm direct reads at the head of the solve: 42 % of its stall samples, r02t.)
  const double* g2a; const double* g2b;
  MC_D void g2_wait() { g_wait(); }
  MC_D double g2(int r) const { return r < kStageRows ? g2a[r * kTile] : g2b[(kG2Spill + r - kStageRows) * kTile]; }
  // after sweep C: every lane is done with the stage rows and the node arrays; fetch the next step's G0 behind the end of this one
  MC_D void end_begin(bool more) {
    // (no proxy fence: sweep C only READ the stage rows and the node arrays; the generic-proxy writes to them — the solve's scratch — were
    // fenced by sweepC_begin.  The barrier orders every lane's last read before the copy overwrites the rows.)
    __syncthreads();
    if (more) g0_issue();
  }
  MC_D double g0(int r) const { MC_CHK(r >= 0 && r < kG0Rows, "G0 row", r); return g0p[r * kTile]; }
  MC_D double g1a(int r) const { MC_CHK(r >= P0_A0CAP && r < P0_G1AEND, "G1a row", r); return g1ap[(kGRow0 + r - P0_A0CAP) * kTile]; }
  MC_D double g1b(int r) const { MC_CHK(r >= P0_LON && r < P0_G1BEND, "G1b row", r); return g1bp[(kGRow0 + r - P0_LON) * kTile]; }
  MC_D void issueA(int i, unsigned kk) {
    const unsigned b = buf32 + (kk & 1) * (kStageRows * kTile * 8);
    const unsigned mb = bar32 + (kk & 1) * 8;
    const int PAS = pas(i);
    const unsigned bytes = (PAS + (i > 1 ? 1 : 0) + 1) * kTile * 8;
    mbar_expect_tx(mb, bytes);
    bulk_g2s_stream(b, Pl + (long long)L.pa(i, 0) * kTile, PAS * kTile * 8, mb, Pl == Pt ? pol : pol_keep);
#ifndef MC_AKEEP
#define MC_AKEEP 1      // 1: the tc / gt rows sweep A reads are kept in L2 (evict-last) for sweep B's read of the same rows half a step later
#endif
    if (i > 1) bulk_g2s_stream(b + PAS * kTile * 8, St + (long long)L.sl(i - 1, SL_TC) * kTile, kTile * 8, mb, MC_AKEEP ? pol_keep : pol);
    bulk_g2s_stream(b + (PAS + 1) * kTile * 8, St + (long long)L.gt(i) * kTile, kTile * 8, mb, MC_AKEEP ? pol_keep : pol);
  }
  MC_D void issueB(int i, unsigned kk) {
    const unsigned b = buf32 + (kk & 1) * (kStageRows * kTile * 8);
    const unsigned mb = bar32 + (kk & 1) * 8;
    mbar_expect_tx(mb, (PBS + SL_N + kWLiveB + 1) * kTile * 8);
    bulk_g2s_stream(b, Pl + (long long)L.pb(i, 0) * kTile, PBS * kTile * 8, mb, Pl == Pt ? pol : pol_keep);
    bulk_g2s_stream(b + PBS * kTile * 8, St + (long long)L.sl(i, 0) * kTile, SL_N * kTile * 8, mb, pol);
#if MC_L2HINT
    bulk_g2s_stream(b + (PBS + SL_N) * kTile * 8, Wt + (long long)L.wl(i, 0) * kTile, kWLiveB * kTile * 8, mb, pol_keep);
#else
    bulk_g2s(b + (PBS + SL_N) * kTile * 8, Wt + (long long)L.wl(i, 0) * kTile, kWLiveB * kTile * 8, mb);
#endif
    bulk_g2s_stream(b + (PBS + SL_N + kWLiveB) * kTile * 8, St + (long long)L.gt(i) * kTile, kTile * 8, mb, pol);
  }
  MC_D void issueC(int i, unsigned kk) {                  // rows of layers i .. i+kCL-1 (or up to m)
    const unsigned b = buf32 + (kk & 1) * (kStageRows * kTile * 8);
    const unsigned mb = bar32 + (kk & 1) * 8;
    const int nl = (L.m - i + 1 < kCL) ? (L.m - i + 1) : kCL;
    mbar_expect_tx(mb, nl * kWLiveC * kTile * 8);
#if MC_L2HINT
    bulk_g2s_stream(b, Wt + (long long)L.wl(i, WL_TT) * kTile, nl * kWLiveC * kTile * 8, mb, pol_keep);
#else
    bulk_g2s(b, Wt + (long long)L.wl(i, WL_TT) * kTile, nl * kWLiveC * kTile * 8, mb);
#endif
  }
  // generic-proxy stores of the previous sweep -> visible to the async proxy, block-wide; then start the next sweep
  MC_D void step_begin() {
    fence_proxy_async();
    __syncthreads();
    if (lane == 0) issueA(L.m, k);
  }
  MC_D void sweepB_begin() {
    fence_proxy_async();
    __syncthreads();
    if (lane == 0) issueB(1, k);
  }
  MC_D void sweepC_begin() {
    fence_proxy_async();
    __syncthreads();
    if (lane == 0) issueC(1, k);
  }
  MC_D void stageC(int i) {
    const int q = (i - 1) % kCL;
    if (q == 0) {
      __syncthreads();
      if (lane == 0 && i + kCL <= L.m) issueC(i + kCL, k + 1);
      mbar_wait(bar32 + (k & 1) * 8, (k >> 1) & 1);
      cbase = buf + (k & 1) * (kStageRows * kTile) + lane;
      ++k;
    }
    cur = cbase + q * kWLiveC * kTile;
  }
  MC_D void stageA(int i) {
    __syncthreads();                                   // everybody has left the buffer the next stage will fill
    if (i == L.m) {
      // sweep B's first stage will land in buffer (k + m) & 1: G1b (read after that copy has been issued) goes to the other one,
      // G1a (read before it) to that one; both in rows 15.., which the sweep-A stages leave alone
      const unsigned ba = (k + L.m) & 1, bb = ba ^ 1;
      if (lane == 0) {
        mbar_expect_tx(bar32 + 16, (kG1aRows + kG1bRows) * kTile * 8);
        g_issue(buf32 + (ba * kStageRows + kGRow0) * (kTile * 8), P0_A0CAP, kG1aRows);
        g_issue(buf32 + (bb * kStageRows + kGRow0) * (kTile * 8), P0_LON, kG1bRows);
      }
      g1ap = buf + ba * (kStageRows * kTile) + lane; g1bp = buf + bb * (kStageRows * kTile) + lane;
    }
    if (lane == 0 && i > 1) issueA(i - 1, k + 1);
    mbar_wait(bar32 + (k & 1) * 8, (k >> 1) & 1);
    cur = buf + (k & 1) * (kStageRows * kTile) + lane;
    an = pas(i);
    ++k;
  }
  MC_D void stageB(int i) {
    __syncthreads();
    if (lane == 0 && i < L.m) issueB(i + 1, k + 1);
    if (i == L.m) {                                  // G2 behind the last stage (see g2)
      const int ns = 3 * L.sm, na = ns < kStageRows ? ns : kStageRows;
      const unsigned fb = (k + 1) & 1, cb = k & 1;
      if (lane == 0) {
        mbar_expect_tx(bar32 + 16, ns * kTile * 8);
        const unsigned dst = buf32 + fb * (kStageRows * kTile * 8);
        bulk_g2s_stream(dst, Pt + (long long)P0_NFIX * kTile, 2 * L.sm * kTile * 8, bar32 + 16, pol);
        bulk_g2s_stream(dst + 2 * L.sm * kTile * 8, St + (long long)L.soiltc(1) * kTile, (na - 2 * L.sm) * kTile * 8, bar32 + 16, pol);
        if (ns > na) bulk_g2s_stream(buf32 + (cb * kStageRows + kG2Spill) * (kTile * 8), St + (long long)L.soiltc(1 + na - 2 * L.sm) * kTile, (ns - na) * kTile * 8, bar32 + 16, pol);
      }
      g2a = buf + fb * (kStageRows * kTile) + lane; g2b = buf + cb * (kStageRows * kTile) + lane;
    }
    mbar_wait(bar32 + (k & 1) * 8, (k >> 1) & 1);
    cur = buf + (k & 1) * (kStageRows * kTile) + lane;
    ++k;
  }
  MC_D double a_p(int r) const { MC_CHK(r >= 0 && r < an, "sweep-A stage row", r); return cur[r * kTile]; }
  MC_D double a_tcl() const { return cur[an * kTile]; }
  MC_D double a_gto() const { return cur[(an + 1) * kTile]; }
  MC_D double b_p(int r) const { MC_CHK(r >= 0 && r < PBS, "sweep-B geometry row", r); return cur[r * kTile]; }
  MC_D double b_s(int r) const { MC_CHK(r >= 0 && r < SL_N, "sweep-B state row", r); return cur[(PBS + r) * kTile]; }
  MC_D double b_w(int r) const { MC_CHK(r >= 0 && r < kWLiveB, "sweep-B work row", r); return cur[(PBS + SL_N + r) * kTile]; }
  MC_D double b_gto() const { return cur[(PBS + SL_N + kWLiveB) * kTile]; }
  MC_D double c_w(int r) const { return cur[r * kTile]; }
  MC_D double c_wq(int q, int r) const { return cbase[(q * kWLiveC + r) * kTile]; }
};
#endif

// 1.41 * gforcedfree(d, u, tc, dtc, pk, dtmin = 1) (microctools; mc_phys.cuh gforcedfree), Pr = nu / Dh = 13.3 / 18.9, with what the forced and
// the free branch share taken out of the maximum.  With Dh = 18.9e-6 sc, nu = 13.3e-6 sc:
//   gforced = Cf (ph / d) sqrt(sc) sqrt(u d),  gfree = Cr (ph / d) sqrt(sc) (9.81 d^3 |dT| / Tk)^(1/4)
//   => 1.41 max(gforced, gfree) = (1.41 / d) ph sqrt(sc max(Cf^2 d u, sqrt(Cr^4 9.81 d^3 |dT| / Tk)))
// with Cf = 0.664 * 18.9e-6 * Pr^(1/3) / sqrt(13.3e-6), Cr = 0.54 * 18.9e-6 * Pr^(1/4) / sqrt(13.3e-6).  kA = Cf^2 d, kB = Cr^4 9.81 d^3 and
// kI = 1.41 / d are column constants (fast_park_constants).  Four square roots on two independent chains and no reciprocal, instead of
// five and one in a row.  A negative or NaN wind speed loses the comparison as the NaN of sqrt(Re) did.
MC_HD double gha_fast2(double kA, double kB, double kI, double u, double Tk, double iTk, double ph, double dtc, double c101) {
  const double x = Tk * (1 / 273.15);
  const double r = fsqrt_nz(x);
  const double sc = (x * r) * (fsqrt_nz(r) * c101);               // (Tk / 273.15)^1.75 * 101.3 / pk
  double adt = fabs(dtc);
  if (adt < 1.0) adt = 1.0;
  const double A = u * kA, B = fsqrt_nz((kB * adt) * iTk);
  const double M = (A > B) ? A : B;
  return (kI * ph) * fsqrt_nz(sc * M);
}

// Scratch map of the solves (indices into Mem::scr_*; the C stages later overwrite rows 0, 1, 27, 28 only):
//   heat   : K(i) at SK0+i (i = 1..17), CD(i) at SCD0+i (1..16), TC(i) at STC0+i (1..18; holds the solution afterwards)
//   vapour : K, CD, TC in the (then dead) heat K/CD rows;  node arrays for sweep C: T2, Yt, Ye (1..8 each)
// sweep C stages use rows 0..9 of both buffers: the node arrays sit in rows 10..25 of buffer 0 and 10..17 of buffer 1; the new
// soil temperatures are written to the state rows right after the heat solve.
enum { SK0 = 0, SCD0 = 17, STC0 = 33, VK0 = 0, VCD0 = 8, VTC0 = 15, NT20 = 9, NYT0 = 17, NYE0 = 36 };
static_assert(2 * kCL <= NT20 + 1 && NYE0 + 1 >= kStageRows + 2 * kCL, "sweep-C stage rows overlap the node arrays");
static_assert(STC0 + kNG + 10 + 2 < 2 * kStageRows, "solver scratch exceeds the two stage buffers");
struct ScrMapStage { static constexpr int K0 = SK0, CD0 = SCD0, TC0 = STC0, vK0 = VK0, vCD0 = VCD0, vTC0 = VTC0, T20 = NT20, YT0 = NYT0, YE0 = NYE0; };
// The wide solve: a step in which a lane closes more than kNG merged layers (calm, stably stratified hours) used to drop the column to the
// general kernel, whose redo launch costs 38 ms per 168-step launch whatever the number of such columns (one thread per column, serial).
// Such a lane now runs the same solve code on arrays sized for up to m merged layers in thread-private memory (1.9 KB of stack, touched
// only then); the warp's other lanes idle through it.  Same arithmetic, same order.
struct ScrMapLane { static constexpr int K0 = 0, CD0 = 32, TC0 = 63, vK0 = 96, vCD0 = 118, vTC0 = 139, T20 = 162, YT0 = 185, YE0 = 208, N = 232; };
static_assert(ScrMapLane::CD0 > kNGW + 10 + 1 && ScrMapLane::TC0 - ScrMapLane::CD0 > kNGW + 10 && ScrMapLane::vK0 - ScrMapLane::TC0 > kNGW + 10 + 2 &&
              ScrMapLane::vCD0 - ScrMapLane::vK0 > kNGW + 1 && ScrMapLane::vTC0 - ScrMapLane::vCD0 > kNGW && ScrMapLane::T20 - ScrMapLane::vTC0 > kNGW + 2 &&
              ScrMapLane::YT0 - ScrMapLane::T20 > kNGW + 2 && ScrMapLane::YE0 - ScrMapLane::YT0 > kNGW + 2 && ScrMapLane::N - ScrMapLane::YE0 > kNGW + 2, "lane scratch map");
struct LaneScr {
  double a[ScrMapLane::N];
  MC_HD double scr_get(int idx) const { return a[idx]; }
  MC_HD void scr_put(int idx, double v) { a[idx] = v; }
};

// Thomas (code.R:94-125) in fused form: one forward pass forms the right-hand side on the fly and eliminates, one backward
// pass substitutes and applies the old-stencil limits (quirk Q11).  Coefficients come from the scratch rows and are replaced in
// place by the elimination factors (cc, dd) as the forward pass consumes them.  Same operations in the same order as
// Column::thomas.  N may differ per lane (N = 0: nothing to solve).
template <class Mem>
MC_HD void thomas_s(Mem& mem, int k0, int cd0, int tc0, int N, double tmsoil, double tairv, double f, const double* xv, int nx) {
  if (N < 1) return;
  const double g = 1 - f;
  double t_lo = mem.scr_get(tc0 + 1);
  double t_mid = mem.scr_get(tc0 + 2) + g * (nx >= 1 ? xv[1] : 0.0);
  double k_lo = mem.scr_get(k0 + 1);
  double a1p = 0, ccp = 0, dnp = 0, tnN1 = 0;
#pragma unroll 1
  for (int x = 2; x <= N + 1; ++x) {
    const double k_mid = mem.scr_get(k0 + x), cd = mem.scr_get(cd0 + x - 1);
    double t_hi = mem.scr_get(tc0 + x + 1);
    if (x + 1 <= N + 1) t_hi = t_hi + g * (x <= nx ? xv[x] : 0.0);
    double d = g * k_lo * t_lo + (cd - g * (k_mid + k_lo)) * t_mid + g * k_mid * t_hi;
    if (x == 2) d = d + k_lo * tairv * f;
    if (x == N + 1) d = d + k_mid * f * tmsoil;
    double b = f * (k_mid + k_lo) + cd;
    if (x > 2) { d = d - a1p * dnp; b = b - a1p * ccp; }
    mem.scr_put(tc0 + x, t_mid);                       // tcv' = tcv + (1-f) X, needed again for the limits
    if (x <= N) {
      const double ib = frcp(b);
      const double a1 = -k_mid * f;
      const double c = a1 * ib, dn = d * ib;
      mem.scr_put(k0 + x, c); mem.scr_put(cd0 + x - 1, dn);      // elimination factors replace the (now dead) coefficients
      a1p = a1; ccp = c; dnp = dn;
    } else tnN1 = fdiv(d, b);
    t_lo = t_mid; t_mid = t_hi; k_lo = k_mid;
  }
  double tu = tnN1;
  double t_up = mem.scr_get(tc0 + N + 2), t_c = mem.scr_get(tc0 + N + 1);
#pragma unroll 1
  for (int x = N + 1; x >= 2; --x) {
    if (x <= N) tu = mem.scr_get(cd0 + x - 1) - mem.scr_get(k0 + x) * tu;
    const double t_dn = mem.scr_get(tc0 + x - 1);
    const double xmn = dmin(dmin(t_c, t_dn), t_up), xmx = dmax(dmax(t_c, t_dn), t_up);
    double v = tu;
    if (v < xmn) v = xmn;
    if (v > xmx) v = xmx;
    mem.scr_put(tc0 + x, v + f * ((x - 1) <= nx ? xv[x - 1] : 0.0));
    t_up = t_c; t_c = t_dn;
  }
}

struct StepOut { double o[8]; int ng; };

// wall-clock per phase of the step (MC_DBG_SKIP bit 16): one lane per block accumulates clock64() deltas
// Both are compiled in only with -DMC_ABLATE (make ABLATE=1): the production kernel carries no debug tests in its loops.
#if defined(__CUDA_ARCH__) && defined(MC_ABLATE)
#define MC_TICK(slot) do { if (cfg.dbg & 16) { long long t_ = clock64(); mem.tcyc[slot] += t_ - mem.tlast; mem.tlast = t_; } } while (0)
#define MC_SKIP(bit) (cfg.dbg & (bit))
#elif defined(__CUDA_ARCH__)
// production build: one performance-monitor trigger (PMTRIG) per phase boundary and step, as a marker that splits a profile's
// SASS listing by phase (tools/ncu_phase_split.py)
#define MC_TICK(slot) asm volatile("pmevent %0;" ::"n"(slot))
#define MC_SKIP(bit) false
#else
#define MC_TICK(slot) ((void)0)
#define MC_SKIP(bit) false
#endif

// Step context: per-thread scalars handed from one phase of the step to the next.  On the device they live in tensor
// memory (PipeMem::park / unpark), so that no phase has to keep another phase's values in registers (or spill them to
// local memory, which is an L2 round trip here); the host build keeps them in a plain array.
enum { X_TAIRN, X_PKN, X_SKYEM, X_RSW, X_RELHUMN, X_TCAN, X_TAIRP, X_RELHUMP,
       X_PPK, X_PS1, X_TET_TAIR, X_TET_TCAN, X_TFROST, X_PSIM, X_PHIM, X_TABOVE,
       X_TSOILP, X_PHA, X_TC1, X_UH, X_SQPHI, X_CS, X_GTNCAP, X_RGCAP,
       X_XS, X_CP1, X_TT, X_LT, X_STL, X_SSW, X_LAM, X_CV,
       // sweep-B invariants, 8 per group, one LDTM each
       KR_K, KR_MUL, KR_DP, KR_RSW, KR_AG0, KR_CLUMP, KR_SKLW, KR_EMLW,
       KG_IDZ, KG_GSMAX, KG_Q50, KG_LWD, KG_ILWD, KG_LWD3, KG_C101, KG_CPK,
       KT_EREF, KT_ESOIL, KT_MTREF, KT_PS1, KT_RGCAP, KT_TNS1, KT_TSEL, KT_CPPK,
       KU_LAMBDA, KU_IMPK, KU_VSB, KU_TCAN, KU_EAMN, KU_IPPK, KU_PKN, KU_REFG, K_NPARK };
static_assert(K_NPARK <= 64, "64 doubles of tensor-memory scratch per thread");

// Column constants of sweep B that live in tensor memory for the whole launch (nothing else writes these slots).
template <class Mem>
MC_HD void fast_park_constants(Mem& mem) {
  const double* const P = mem.P;
  auto p0 = [=](int r) { return P[(long long)r * kTile]; };
  const double lwd = p0(P0_LWD);
  mem.park(KR_CLUMP, p0(P0_CLUMP));
  mem.park(KG_IDZ, (double)mem.L.m / p0(P0_HGT));                  // m / hgt: 1 / dz of the regular grid (used by the REG instantiation only)
  mem.park(KG_GSMAX, p0(P0_GSMAX)); mem.park(KG_Q50, p0(P0_Q50));
  mem.park(KG_LWD, kGhaCf2 * lwd); mem.park(KG_ILWD, 1.41 / lwd); mem.park(KG_LWD3, kGhaCr4 * (9.81 * (lwd * lwd * lwd)));   // kA, kI, kB of gha_fast2
  mem.park(KU_VSB, p0(P0_VEGEM) * kSb); mem.park(KU_REFG, p0(P0_REFG));
  mem.park_done();
}

// One time step of one column in streaming form.  Every lane computes (tensor-memory accesses are warp-collective);
// `run` false (a column that takes this step through the general path) only suppresses the stores into the state rows.
// Returns false when the column left the fast envelope (nothing of the state has been modified then).
template <class Mem, int MS, bool REG>
MC_HD bool fast_step(Mem& mem, const Forcing& fc, const SolDay& sday, const RunCfg& cfg /* launch-uniform fields only */, double surfwet, bool run, bool emit, bool more /* another step follows in this launch */,
                     double reqhgt, int charmode,
                     double relhum0, double rsw0, double skyem0, StepOut& out) {
  const FastLayout L = mem.L;
  const int m = L.m, sm = L.sm;
  const double* const P = mem.P; double* const S = mem.S; double* const W = mem.W;
  auto p0 = [=](int r) { MC_CHK(r >= 0 && r < L.nP(), "geometry row", r); return P[(long long)r * kTile]; };
#if defined(__CUDA_ARCH__) && (MC_KEEP & 2)
  auto s0 = [&](int r) { return mem.ld_keep(S + (long long)r * kTile); };
  auto sput = [&](int r, double v) { mem.st_keep(S + (long long)r * kTile, v); };
#else
  auto s0 = [=](int r) { MC_CHK(r >= 0 && r < L.nS(), "state row (read)", r); return S[(long long)r * kTile]; };
  auto sput = [=](int r, double v) { MC_CHK(r >= 0 && r < L.nS(), "state row (write)", r); S[(long long)r * kTile] = v; };
#endif
  auto wput = [=](int i, int r, double v) { MC_CHK(i >= 1 && i <= L.m && r >= 0 && r < WL_N, "work row (write), layer", i); W[(long long)L.wl(i, r) * kTile] = v; };
  auto wget = [=](int i, int r) { MC_CHK(i >= 1 && i <= L.m && r >= 0 && r < WL_N, "work row (read), layer", i); return W[(long long)L.wl(i, r) * kTile]; };
  MC_TICK(7);
  mem.step_begin();
  bool fastok = run;
  // ======================= phase 0: scalars of the step (code.R:700-745) =======================
  double uh, sqphi, iphi, ugr, cpk, ps1a, tcu;
  double lu = 0, lugr = 0;                            // log(uh), log(ugr): sweep A works on logarithms
  double idz = 0;                                     // REG: m / hgt
  const bool edge = cfg.edgedist == cfg.edgedist;
  {
    const double tairn = fc.tair, pkn = fc.pk;
    double relhumn = fc.relhum, u = fc.u;
    const double ppk = s0(S0_PK), Hprev = s0(S0_H), ps1 = s0(L.soiltc(1));
    const double hgt = mem.g0(P0_HGT);
    cpk = 44.6 * fdiv(pkn, 101.3);                                                  // phair(t, pk) = c * 273.15 / (t + 273.15)
    const double pha = phair(tairn, pkn);
    double u2;
    if (cfg.metopen) {
      if (cfg.windhgt != 2) u = fdiv(u * 4.87, flog(67.8 * cfg.windhgt - 5.42));
      u2 = fdiv(u * mem.g0(P0_LA), log(67.8 * 2 - 5.42));
    } else {
      u2 = u;
      if (cfg.windhgt != (2 + hgt)) {                                        // windprofile(u, windhgt, hgt+2, ...) with zo above the canopy
        const double psi = s0(S0_PSIM);
        double ln2 = p0(P0_WL2) + psi;
        if (ln2 < 0.55) ln2 = 0.55;
        u2 = fdiv(0.4 * fdiv(u, p0(P0_WL1) + psi), 0.4) * ln2;
      }
    }
    if (u2 < 0.5) u2 = 0.5;
    const double lnref = mem.g0(P0_LNREF);
    const double uf = fdiv(0.4 * u2, lnref);
    double psi_mn, psi_h;
    diabatic_cor(tairn, pkn, Hprev, uf, hgt + 2, mem.g0(P0_D), psi_mn, psi_h);
    const double tcm = s0(L.sl(m, SL_TC)), tc1 = s0(L.sl(1, SL_TC));
    double phi_m;
    {
      const double dz = mem.g0(P0_DZTB);
      const double idz = frcp(dz);
      const double dtdz = (tcm - tc1) * idz;
      double dudz = (s0(L.uz(m)) - s0(L.uz(1))) * idz;
      if (fabs(dudz) < 0.01) dudz = 0.01;
      const double Tk = fdiv(s0(S0_TCSUM), (double)m) + 273.15;
      double Ri = fdiv(fdiv(9.81, Tk) * dtdz, dudz * dudz);
      if (Ri > 0.15) Ri = 0.15;
      if (Ri < -1) Ri = -1;
      if (Ri < 0) { double phh = frcp(fsqrt(1 - 16 * Ri)); phi_m = fsqrt(phh); }
      else { double st = fdiv(Ri, 1 - 5 * Ri); phi_m = 1 + fdiv(6 * st, 1 + st); }
    }
    double tcan;
    {
      const double ufh = fdiv(0.4 * u2, mem.g0(P0_LZU) + psi_h);
      const double mm = fdiv(Hprev, 0.4 * pha * cpair(tairn) * ufh);
      const double tref = tairn + mm * (mem.g0(P0_LZUH) + psi_h);
      tcan = tref - mm * (mem.g0(P0_LZOH) + psi_h);
    }
    const double tfrost = dewpoint(fdiv(relhumn, 100) * satvap(tairn), tairn);
    if (tcan < tfrost) tcan = tfrost;
    if (tcan > (tairn + 15)) tcan = tairn + 15;
    if (tcan < (tairn - 4.5)) tcan = tairn - 4.5;
    const double tet_tair = tetens(tairn), tet_tcan = tetens(tcan);
    {
      const double ea = tet_tair * fdiv(relhumn, 100);
      relhumn = fdiv(ea, tet_tcan) * 100;
      if (relhumn > 100) relhumn = 100;
    }
    sqphi = fsqrt(phi_m); iphi = frcp(phi_m);
    ugr = fdiv(0.4 * fdiv(u, mem.g0(P0_GLNH)), 0.4);                                    // grass-profile ur = ugr * gl (uref = u, zref = zu)
    {
      const double ufw = 0.4 * fdiv(u2, lnref + psi_mn);
      double ln2 = mem.g0(P0_LNHD) + psi_mn;
      if (ln2 < 0.55) ln2 = 0.55;
      uh = fdiv(ufw, 0.4) * ln2;
    }
    ps1a = ps1; tcu = tcm;
    if (REG) idz = (double)m * frcp(hgt);              // 1 / dz of the regular grid
    lu = flog_nb(uh); lugr = flog_nb(ugr);                 // (uh >= 0.5 * ... > 0; ugr = 0 in a calm hour gives -Inf, which loses every maximum)
    mem.park(X_TAIRN, tairn); mem.park(X_PKN, pkn); mem.park(X_SKYEM, fc.skyem); mem.park(X_RSW, fc.Rsw); mem.park(X_RELHUMN, relhumn);
    mem.park(X_TCAN, tcan); mem.park(X_TAIRP, s0(S0_TAIR)); mem.park(X_RELHUMP, s0(S0_RELHUM));
    mem.park(X_PPK, ppk); mem.park(X_PS1, ps1); mem.park(X_TET_TAIR, tet_tair); mem.park(X_TET_TCAN, tet_tcan); mem.park(X_TFROST, tfrost);
    mem.park(X_PSIM, psi_mn); mem.park(X_PHIM, phi_m); mem.park(X_TABOVE, s0(S0_TABOVE));
    mem.park(X_TSOILP, s0(S0_TSOIL)); mem.park(X_PHA, pha); mem.park(X_TC1, tc1);
    mem.park_done();
  }
  MC_TICK(0);
  // ======================= sweep A: wind and turbulent conductances, top -> bottom (code.R:747-770) =======================
  {
    double sfx = 0, cs = 0;
    double lglo_up = 0, edo_up = 0;                     // REG: log(gl), ed of the outer node of the layer above
    const int half = (int)nearbyint(m / 2.0);
    for (int i = m; i >= 1; --i) {
      mem.stageA(i);
      if (MC_SKIP(1)) continue;
      const double tcl = (i > 1) ? mem.a_tcl() : ps1a;
      const double ai = mem.a_p(PA_A0I) * sqphi;
      double uzn;
      const double ph = phair_c(cpk, (tcu + tcl) / 2);
      // windcanopy (code.R:279-294) on logarithms: log(uh exp(a ex)) = lu + a ex, log(ur exp(a ed)) = log(ugr) + log(gl) + a ed; the same
      // NaN-aware maximum; only the wind speed that is stored is exponentiated.  gcanopy's uh / (e0 - e1) takes lu into the exponents.
      auto wcl = [&](double l, double a, double ex, double lgl, double ed) {
        double t = fma(a, ex, l);
        if (edge) {
          const double r = fma(a, ed, lugr + lgl);
          if (t != t) t = r;
          else if (r == r && r > t) t = r;
        }
        return t;
      };
      // the rows that depend on the node grid alone: on the regular grid zi = (i + 1) dz, zo = i dz, z[i] = (i - 0.5) dz, z0 = (i - 1.5) dz
      // (0 under the lowest node), so ex1 = zo / zi - 1 = -1 / (i + 1), ex2 = zi / zo - 1 = 1 / i, gx0 = z0 / zi - 1, gx1 = z[i] / zi - 1;
      // the top layer refers to hgt = m dz (Column::precompute)
      double ex1, ex2, gx0, gx1, rdz;
      if (REG) {
        const double r1 = rinv_small(i == m ? m : i + 1);
        ex1 = (i == m) ? -0.5 * r1 : -r1; ex2 = (i == m) ? ex1 : rinv_small(i);
        gx0 = (i == m) ? -1.5 * r1 : (i > 1 ? -2.5 * r1 : -1.0); gx1 = (i == m) ? -0.5 * r1 : -1.5 * r1;
        rdz = (i == 1) ? 2 * idz : idz;
      } else { ex1 = mem.a_p(PA_EX1); ex2 = mem.a_p(PA_EX2); gx0 = mem.a_p(PA_GX0); gx1 = mem.a_p(PA_GX1); rdz = mem.a_p(PA_RDZ); }
      double lz;
      const double lglo = mem.a_p(PA_GLO), edo = mem.a_p(PA_EDO);
      if (i == m) lz = wcl(lu, ai, ex1, lglo, edo);
      else {
        // inner node of the layer: its rows, or on the regular grid (below layer m - 1) the outer node of the layer above — the same
        // height — with ed rescaled from that layer's top to this layer's outer node: zi(i + 1) / zo(i) = (i + 2) / i
        const double lgli = (!REG || i == m - 1) ? mem.a_p(PA_GLI) : lglo_up;
        const double edi = (!REG || i == m - 1) ? mem.a_p(PA_EDI) : edo_up * ((double)(i + 2) * rinv_small(i));
        lu = wcl(lu, ai, ex1, lglo, edo);
        lz = wcl(lu, mem.a_p(PA_A0O) * sqphi, ex2, lgli, edi);                           // Q9
      }
      lglo_up = lglo; edo_up = edo;
      uzn = fexp_neg(lz);                               // (clamped below only: lz <= log of the wind above the canopy)
      const double nlu = (lu < -600.0) ? 600.0 : -lu;   // (without the edge term the product can underflow; the exponents stay finite)
      const double e0 = fexp_b(fma(-ai, gx0, nlu)), e1 = fexp_b(fma(-ai, gx1, nlu));      // (gx in [-1, 0])
      const double g = fdiv(mem.a_p(PA_CG) * ph * ai, e0 - e1) * iphi;
      const double gfree = 0.0012 * ph * qroot(fabs(tcu - tcl) * rdz);
      const double gtn = (g > gfree) ? g : gfree;
      const double rn = frcp(gtn);
      const double rg = frcp(0.5 * gtn + 0.5 * mem.a_gto());
      double* Wl = W + (long long)L.wl(i, 0) * kTile;
      mem.st_keep(Wl + WL_GTN * kTile, gtn); mem.st_keep(Wl + WL_UZN * kTile, uzn);
      // phz is rebuilt in sweep B from rows it stages anyway (old tc of the layer, dz); rg and its sum above each layer are not handed to sweep B through memory: B recomputes rg from gtn and the old gt row
      // and rebuilds the sum above layer i as (total - sum up to i)
      sfx += rg;
      if (i <= half) cs += rn;
      tcu = tcl;
    }
    // canopy top <-> reference (code.R:769), Q8: leftover uh, lowest air node
    uh = fexp_neg(lu);
    const double tcan = mem.unpark(X_TCAN), tc1 = mem.unpark(X_TC1);
    mem.g1_wait();                                     // the G1 rows were fetched behind the sweep
    const double a = mem.g1a(P0_A0CAP) * sqphi;
    const double ph = phair_c(cpk, (tcan + tc1) / 2);
    const double e0 = fexp_b(-a * mem.g1a(P0_GXC0));
    const double g = fdiv(mem.g1a(P0_CGCAP) * ph * uh * a, e0 - 1.0) * iphi;
    const double gfree = 0.0012 * ph * qroot(fabs(tcan - tc1) * mem.g1a(P0_RDZCAP));
    {                                                  // soilk (microctools; Column::soilk) for this step's theta
      const double ct = mem.g1a(P0_SKC) * fc.theta;
      mem.park(X_LAM, mem.g1a(P0_SKA) + mem.g1a(P0_SKB) * fc.theta - mem.g1a(P0_SKAD) * fexp(-((ct * ct) * (ct * ct))));
      mem.park(X_CV, (mem.g1a(P0_CVDRY) + 4.18 * fc.theta) * 1e6);
    }
    const double gtncap = (g > gfree) ? g : gfree;
    const double rgcap = frcp(0.5 * gtncap + 0.5 * s0(S0_GTCAP));
    mem.park(X_UH, uh); mem.park(X_SQPHI, sqphi); mem.park(X_CS, cs); mem.park(X_GTNCAP, gtncap); mem.park(X_RGCAP, rgcap);
    mem.park(KT_TNS1, sfx);                            // sum of rg over all layers (the slot is free until the solve)
  }
  MC_TICK(1);
  mem.sweepB_begin();
  // ======================= scalars of sweep B: radiation geometry, reference vapour (code.R:776-785) =======================
  bool day;
  {
    double x0[8], x1[8];
    mem.park_done();
    mem.unpark8(X_TAIRN, x0); mem.unpark8(X_PPK, x1);
    const double tairn = x0[X_TAIRN], pkn = x0[X_PKN], skyem = x0[X_SKYEM], Rsw = x0[X_RSW], relhumn = x0[X_RELHUMN], tcan = x0[X_TCAN];
    const double ppk = x1[X_PPK - X_PPK], ps1 = x1[X_PS1 - X_PPK], tet_tcan = x1[X_TET_TCAN - X_PPK];
    const double clump = mem.unpark(KR_CLUMP), refg = mem.unpark(KU_REFG), emv = p0(P0_EMV);   // launch-resident (fast_park_constants)
    double dp = fc.dp;
    // `sa` is sin(solar altitude) from here on: only its sign, the 5-degree test, sin and cot of the angle are used (mc_phys.cuh)
    const double sa = solsin_pre(fc.lt, mem.g1b(P0_LON), cfg.merid, cfg.dst, sday, mem.g1b(P0_SLAT), mem.g1b(P0_CLAT));
    if (dp != dp) dp = difprop_sh(Rsw, sa);
    const double xx = mem.g1b(P0_X), kden = mem.g1b(P0_KDEN);
    double K, mul;
    cank2_sh(xx, sa, kden, mem.g1b(P0_K5), K, mul);
    const double aG0 = refg * cansw_k(Rsw, dp, sa, mem.g1b(P0_SUMPAI), mem.g1b(P0_SM), mem.g1b(P0_TRDM), K, clump);
    const double lwout = kSb * pow4(tairn + 273.15);
    day = sa > 0;
    mem.park(KR_K, K); mem.park(KR_MUL, mul); mem.park(KR_DP, dp); mem.park(KR_RSW, Rsw); mem.park(KR_AG0, aG0);
    mem.park(KR_SKLW, emv * (skyem * lwout - emv * lwout)); mem.park(KR_EMLW, emv * (emv * lwout));   // slope and intercept of lwabs in trlw
    mem.park(KG_C101, fdiv(101.3, pkn)); mem.park(KG_CPK, (44.6 * fdiv(pkn, 101.3)) * 273.15);
    const double mtref = 0.5 * tcan + 0.5 * x0[X_TAIRP];
    const double mrh = 0.5 * relhumn + 0.5 * x0[X_RELHUMP];
    const double mpk = 0.5 * ppk + 0.5 * pkn;
    const double eref = fdiv(mrh, 100) * satvap(mtref);
    const double esoil = soilrh(fc.theta, mem.g1b(P0_SOILB), -mem.g1b(P0_PSIE), mem.g1b(P0_SMAX), ps1) * satvap(ps1);     // Q7
    const double eamn = fdiv(relhumn, 100) * tet_tcan;
    mem.park(KT_EREF, eref); mem.park(KT_ESOIL, esoil); mem.park(KT_MTREF, mtref); mem.park(KT_PS1, ps1); mem.park(KT_RGCAP, mem.unpark(X_RGCAP));
    mem.park(KT_CPPK, (44.6 * fdiv(ppk, 101.3)) * 273.15);
    if (REG) mem.park(KT_TSEL, frcp(mem.unpark(KG_IDZ)));             // dz = hgt / m of the regular grid (the slot is free until the solve parks tsel)
    mem.park(KU_LAMBDA, fdiv((-42.575 * tcan + 44994) * surfwet, mpk));     // lambda / mpk: the only form sweep B uses
    mem.park(KU_TCAN, tcan); mem.park(KU_EAMN, eamn); mem.park(KU_IPPK, frcp(ppk)); mem.park(KU_PKN, pkn);
    mem.park_done();
  }
  MC_TICK(2);
  // ======================= sweep B: leafabs + leaftemp, bottom -> top, with the merged-layer sums =======================
  // layermerge / layeraverage bookkeeping (code.R:772, 814): closed groups 1..ng and the open one
  int ng = 0, glo[kNGW + 2], ghi[kNGW + 2];
  double gw[kNGW + 2], gtc[kNGW + 2], gVo[kNGW + 2], gea[kNGW + 2], gX[kNGW + 2];
  const bool runB = run;
  {
    const int selHb = (int)p0(P0_SELH);
    const double its = 1 / cfg.timestep;              // launch-uniform
    bool balances_ok = true;                            // every layer so far takes leaftemp's equilibrium balances
    double acc2 = 0, TT = 0, Lt = 0, stl = 0, ssw = 0;
    double oR = 0, oC = 0, ow = 0, otc = 0, oVo = 0, oea = 0, oX = 0;
    int olo = 1;
    bool open = false;
    for (int i = 1; i <= m; ++i) {
      mem.stageB(i);
      if (MC_SKIP(2)) continue;
      const double ptc = mem.b_s(SL_TC), ptl = mem.b_s(SL_TLEAF);
      const double uzn = mem.b_w(WL_UZN), gtn = mem.b_w(WL_GTN);
      // leafabs (code.R:31-49) of this layer
      const double sref = mem.b_p(PB_SREF), omr = sref * sref;                   // 1 - ref
      double rsw, Rabsn, gvn, ghan, idzb = 0;
      const double Tkc = ptc + 273.15, iTkc = frcp(Tkc);
      {
        double kr[8];
        mem.unpark8(KR_K, kr);
        const double K = kr[KR_K - KR_K], clump = kr[KR_CLUMP - KR_K], dp = kr[KR_DP - KR_K];
        const double trbc = day ? clumptr(fexp_neg(-((sref * mem.b_p(PB_PC)) * K)), clump) : clumptr(0.0, clump);
        const double trbg = day ? clumptr(fexp_neg(-(mem.b_p(PB_SGC) * K)), clump) : clumptr(0.0, clump);
        const double sunl = day ? clumptr(fexp_neg(-K * mem.b_p(PB_PC)), clump) : clumptr(0.0, clump);
        // code.R:43-46 per unit (1 - ref): Rswin = aRsw / (1 - ref) (code.R:895) is formed first, aRsw = (1 - ref) * Rswin
        const double xR = kr[KR_RSW - KR_K] * (dp * mem.b_p(PB_TRDC) + (1 - dp) * trbc);
        const double xG = kr[KR_AG0 - KR_K] * (dp * mem.b_p(PB_TRDG) + (1 - dp) * trbg);
        rsw = xR * (sunl * kr[KR_MUL - KR_K] + (1 - sunl)) + xG;
        const double aRsw = omr * rsw;
        const double trl = mem.b_p(PB_TRLW);
        double kg[8];
        mem.unpark8(KG_IDZ, kg);
        const double aRlw = kr[KR_EMLW - KR_K] + trl * kr[KR_SKLW - KR_K];       // canlw()$lwabs, code.R:47: emv * (trl * sky + (1 - trl) * canopy)
        Rabsn = aRsw + aRlw;
        const double Qa = 4.6 * rsw;
        idzb = kg[KG_IDZ - KG_IDZ];
        gvn = fdiv(kg[KG_GSMAX - KG_IDZ] * Qa, Qa + kg[KG_Q50 - KG_IDZ]);       // layercond
        ghan = gha_fast2(kg[KG_LWD - KG_IDZ], kg[KG_LWD3 - KG_IDZ], kg[KG_ILWD - KG_IDZ], uzn, Tkc, iTkc, kg[KG_CPK - KG_IDZ] * iTkc, ptl - ptc,
                         kg[KG_C101 - KG_IDZ]);
      }
      double kt[8];
      mem.unpark8(KT_EREF, kt);
      const double eref = kt[KT_EREF - KT_EREF], esoil = kt[KT_ESOIL - KT_EREF], mtref = kt[KT_MTREF - KT_EREF], ps1 = kt[KT_PS1 - KT_EREF];
      const double timestep = cfg.timestep;
      // leaftemp (code.R:378-512)
      const double rg = frcp(0.5 * gtn + 0.5 * mem.b_gto());                     // as in sweep A
      acc2 += rg;
      const double sfxi = (i == m) ? 0.0 : kt[KT_TNS1 - KT_EREF] - acc2;            // sum of rg above this layer
      const double gtt = frcp(kt[KT_RGCAP - KT_EREF] + sfxi);
      const double gtt2 = frcp(acc2);
      // previn$z geometry of the layer (code.R:387-392): rows, or on the regular grid zp = (i - 0.5) dz, zth = dz, zref = hgt + 2 - zp
      const double izl = mem.b_p(PB_IZL), PAIi = mem.b_p(PB_PAI);
      const double izth = REG ? idzb : mem.b_p(PB_IZTH);
      const double PAIm = 2 * PAIi * izth;
      const double esj = 0.6108 * fexp_b((17.27 * ptc) * frcp(ptc + 237.3));    // Tetens, code.R:410
      const double eaj = (mem.b_s(SL_RH) * 0.01) * esj;
      const double estl = satvap(ptl);
      const double r237l = frcp(ptl + 237.3);
      const double tetl = 0.6108 * fexp_b((17.27 * ptl) * r237l);
      const double delta = (4098 * tetl) * (r237l * r237l);                      // code.R:415
      const double gvo = mem.b_s(SL_GV), ghao = mem.b_s(SL_GHA);
      // 1/(1/a + 1/b) = a*b/(a+b): one division; a = 0 (closed stomata at night) still gives 0
      const double gvp = gvo;                                                    // the working row holds the previous step's gvc (fast_run_column converts previn$gv)
      const double gvc = fdiv(gvn * ghan, gvn + ghan);
      const double gvm = 0.5 * gvp + 0.5 * gvc;
      const double gham = 0.5 * ghan + 0.5 * ghao;
      // leaftemp switches each balance between a transient form and an equilibrium form: equilibrium when timestep * max(g1, g3, g2) > 1,
      // g1 = gtt / zref, g2 = gtt2 / z, g3 = gv / zla (gha / zla for heat) (code.R:419-449).  On the regular node grid every layer of every
      // test and bench input is in the equilibrium form at the hourly or longer steps the streaming form runs with (the leaf term alone is
      // 10^2-10^4): the REG instantiation does not carry the transient forms; a lane that would need one leaves the streaming form and is
      // redone by the general kernel.  Its test avoids the reciprocals (gtt / zref > 1 / timestep <=> timestep gtt > zref; zref, z > 0 there).
      // A node moved above the canopy (reqhgt > hgt) does take the transient forms at hourly steps: the general instantiation has both.
      bool eqv = true, eqh = true;
      if (REG) {
        const double zp = kt[KT_TSEL - KT_EREF] * ((double)i - 0.5);                // (i - 0.5) dz;  zref = hgt + 2 - z
        const bool big12 = (timestep * gtt > fma(kt[KT_TSEL - KT_EREF], (double)m - ((double)i - 0.5), 2.0)) || (timestep * gtt2 > zp);
        const double tzl = timestep * izl;
        if (!(big12 || tzl * gvm > 1) || !(big12 || tzl * gham > 1)) balances_ok = false;
      }
      const double g1 = REG ? 0.0 : gtt * mem.b_p(PB_IZRF), g2 = REG ? 0.0 : gtt2 * mem.b_p(PB_IZP), g3 = gvm * izl, g3h = gham * izl;
      if (!REG) { eqv = timestep * dmax(dmax(g1, g3), g2) > 1; eqh = timestep * dmax(dmax(g1, g3h), g2) > 1; }
      double ae, be;
      if (REG || eqv) {
        const double gv2b = (PAIm * izth) * gvm;                                // Q21
        const double iden = frcp(gtt + gtt2 + gv2b);
        ae = (gtt * eref + gtt2 * esoil + gv2b * estl) * iden;
        be = (0.5 * delta) * iden;
      } else {
        const double ibtm = frcp(its + 0.5 * (g1 + g2 + g3));
        const double sat = eaj > esj ? 0.0 : 1.0;                               // edf(), code.R:380-386
        const double e1 = (eref > eaj ? eref - eaj : 0.0) * sat;
        const double e2 = (esoil > eaj ? esoil - eaj : 0.0) * sat;
        const double e3 = (estl > eaj ? estl - eaj : 0.0) * sat;
        ae = eaj + 0.5 * (g1 * e1 + g2 * e2 + g3 * e3) * ibtm;
        be = (0.25 * gvm * delta) * ibtm;
      }
      const double cp = cpair(ptc);
      const double phc = kt[KT_CPPK - KT_EREF] * iTkc;
      double aL, bL;
      if (REG || eqh) {
        // code.R:447-449 with the common factor cp of K1, K2, K3 cancelled between numerator and denominator
        const double iden = frcp(gtt + gtt2 + gham);
        aL = (gtt * mtref + gtt2 * ps1 + gham * ptl) * iden;
        bL = (0.5 * gham) * iden;
      } else {
        // ma K_i = (mac / (cp ph)) (g_i cp) = mac g_i / ph and 1 / ph = Tk / (cpk 273.15): no division, cp cancels
        const double ma = mem.p_mac(i) * (Tkc * frcp(kt[KT_CPPK - KT_EREF]));   // (PB_MAC is not staged)
        const double K1 = g1, K2 = g2, K3 = g3h * PAIm;
        const double ibtm = frcp(1 + 0.5 * ma * (K1 + K2 + K3));
        aL = (ptc + 0.5 * ma * (K1 * mtref + K2 * ps1 + K3 * ptl)) * ibtm;
        bL = (0.25 * ma * K3) * ibtm;
      }
      if (bL < 0) bL = 0;
      const double aH = cp * gham * (ptl - aL);
      double bH = cp * gham * (0.5 - 0.5 * bL);
      if (bH < 0) bH = 0;
      // (the second group of step scalars is fetched here, after the vapour / heat balances, so that it does not occupy registers over them)
      double ku[8];
      mem.unpark8(KU_LAMBDA, ku);
      const double eamn = ku[KU_EAMN - KU_LAMBDA];
      const double lg = ku[KU_LAMBDA - KU_LAMBDA] * gvm;                           // (lambda / mpk) * gv, code.R:456
      double edf_l = (estl > ae ? estl - ae : 0.0);
      if (ae > esj) edf_l = 0;
      const double aX = lg * edf_l;
      double bX = lg * (delta - be);
      if (bX < 0) bX = 0;
      const double Ra = 0.5 * Rabsn + 0.5 * mem.b_s(SL_RABS);
      const double tk = ptl + 273.15;
      const double aRe = ku[KU_VSB - KU_LAMBDA] * pow4(tk);
      const double bRe = ku[KU_VSB - KU_LAMBDA] * 2 * (tk * tk * tk);
      const double mz = mem.b_p(PB_MZ);                                          // ml / zla
      const double num = Ra - aRe - aX - aH, bsum = bRe + bX + bH;
      // (ml*num) / (zla + ml*bsum).  Two pieces of the reference are dead arithmetic and are not evaluated: dTL2 = num / bsum replaces dTL only
      // when it is SMALLER in magnitude (code.R:467-469), and |num| / bsum >= |num| / (bsum + zla/ml) for every bsum > 0, ml/zla > 0 (bsum holds
      // 2 eps sigma T^3 > 0; both are NaN together); the limits of tn against max(tair + 15, tleaf) / min(tair - 5, tleaf) (code.R:472-477) can
      // never bind, since they are applied only where tleaf is already beyond tn on that side.  The general column keeps both literally.
      double dTL = fdiv(mz * num, 1 + mz * bsum);
      double tnl = aL + bL * dTL;
      double tlf = ptl + dTL;
      const double eam = ae + be * dTL;
      double ean = eaj + 2 * (eam - eaj);
      const double es = tetens(tnl);
      if (ean > es) ean = es;
      if (ean < eamn) ean = eamn;
      const double al = flog_nb(ean * (1 / 0.6108));          // (branch-free form: ean >= eamn > 0, and a straight-line body lets the scheduler interleave it)
      const double Tdew = -fdiv(237.3 * al, al - 17.27);
      double dT1 = -(ptl + dTL - Tdew);
      if (dT1 < 0) dT1 = 0;
      // dew on the leaf (code.R:497-503): min(dT1, dT2) with dT2 >= 0 is 0 whenever the leaf is not below the dew point (dT1 == 0), so the
      // second Tetens evaluation is only made for lanes that need it (a NaN dT1 still takes the branch and propagates as before)
      if (!(dT1 <= 0)) {
        const double Lc = lg * (tetens(tlf) - ean);
        double dT2 = -mz * Lc;
        if (dT2 < 0) dT2 = 0;
        tlf = tlf + dmin(dT1, dT2);
      }
      dTL = tlf - ptl;
      const double dTmx = ptc + 20.2 - ptl, dTmn = ptc - 4.5 - ptl;
      if (dTL > dTmx) dTL = dTmx;
      if (dTL < dTmn) dTL = dTmn;
      tnl = aL + bL * dTL;
      const double tleafn = ptl + dTL;
      const double Lp = (aX + bX * dTL) * mem.b_p(PB_PAI);                                   // code.R:880 (Q1)
      if (i == 1) {                                                               // soil-surface forcing (code.R:799-811)
        const double refg = ku[KU_REFG - KU_LAMBDA], pkn = ku[KU_PKN - KU_LAMBDA];
        const double esoil = mem.unpark(KT_ESOIL), ps1 = mem.unpark(KT_PS1);   // (fetched again: not carried over the leaf balance of every layer)
        const double cd1 = (mem.unpark(X_CV) * p0(P0_NFIX + sm)) * (0.5 * its);
        const double icd1 = frcp(cd1);
        const double aRsw = (mem.b_p(PB_SREF) * mem.b_p(PB_SREF)) * rsw;        // (1 - ref) * Rswin, as formed for Rabs above
        const double a1 = (1 - refg) * (aRsw * icd1);
        double Xs;
        if (esoil - ean > 0) {
          const double mm = fdiv((-42.575 * ptc + 44994) * phc * p0(P0_Z1), pkn);
          const double dl = fdiv(4098 * tetens(ps1), sq(ps1 + 237.3));
          const double btm = 1 + 0.5 * (mm * icd1) * dl;
          Xs = fdiv(a1 - (mm * icd1) * (esoil - ean), btm);
        } else Xs = a1;
        const double sgn = Xs < 0 ? -1.0 : 1.0;
        const double Xsmx = fdiv((1 - refg) * aRsw, ghan * cp) + (tnl - ps1);
        if (fabs(Xs) > fabs(Xsmx)) Xs = Xsmx;
        Xs = fabs(Xs) * sgn;
        mem.park(X_XS, Xs); mem.park(X_CP1, cp);        // parked here: nothing of it stays live over the other layers
      }
      // layermerge (code.R:772) greedy grouping with the layeraverage sums (code.R:814) of the open group
      {
        const double wi = REG ? ((i == 1) ? 0.5 * kt[KT_TSEL - KT_EREF] : kt[KT_TSEL - KT_EREF]) : mem.b_p(PB_DZ);   // z[i] - z[i-1]
        const double phz = phc * wi, rn = frcp(mem.b_w(WL_GTN));       // phair(previn$tc, previn$pk) * dz as in sweep A
        TT += phz * rn;                                                           // cumsum((ph/gt) * dz), code.R:797
        oR += rn; oC += phz; open = true;
        ow += wi; otc += wi * ptc; oVo += wi * (eaj * ku[KU_IPPK - KU_LAMBDA]); oea += wi * ean; oX += wi * (tnl - ptc);
        if (timestep <= oC * oR) {                                               // timestep / sum(1/gt) <= sum(ph*zth)
          if (ng < kNGW) {             // (at most m <= kNGW groups can close)
            ++ng;
            glo[ng] = olo; ghi[ng] = i; gw[ng] = ow; gtc[ng] = otc; gVo[ng] = oVo; gea[ng] = oea; gX[ng] = oX;
          }
          olo = i + 1; oR = 0; oC = 0; ow = 0; otc = 0; oVo = 0; oea = 0; oX = 0; open = false;
        }
      }
      if (runB) {
        double* Sl = S + (long long)L.sl(i, 0) * kTile;
        mem.st_stream(Sl + SL_TLEAF * kTile, tleafn); mem.st_stream(Sl + SL_RABS * kTile, Rabsn); mem.st_stream(Sl + SL_GV * kTile, gvc);
        mem.st_stream(Sl + SL_GHA * kTile, ghan);
        // previn$uz is only read at the lowest and highest node (code.R:731), so the other rows are written when the state is
        // harvested (emit) and not on every step
        // (stage rows are read again here rather than held in registers since the top of the layer)
        if (emit || i == 1 || i == m) sput(L.uz(i), mem.b_w(WL_UZN));
        mem.st_stream(S + (long long)L.gt(i) * kTile, mem.b_w(WL_GTN));
      }
      // (the EAL row carries log(ean / 0.6108), which this sweep has anyway: sweep C needs ean only inside dewpoint()'s logarithm)
      mem.st_keep(W + (long long)L.wl(i, WL_TT) * kTile, TT); mem.st_keep(W + (long long)L.wl(i, WL_EAL) * kTile, al);
      if (emit || i == selHb) { wput(i, WL_L, Lp); wput(i, WL_RSWIN, rsw); }
      if (emit) wput(i, WL_GVN, gvn);
      Lt += Lp; stl += tleafn; ssw += rsw;
    }
    if (open) {                                     // trailing group joins the last closed one
      if (ng > 0) { ghi[ng] = m; gw[ng] += ow; gtc[ng] += otc; gVo[ng] += oVo; gea[ng] += oea; gX[ng] += oX; }
      else { ng = 1; glo[1] = 1; ghi[1] = m; }
    }
    if (REG && !balances_ok) fastok = false;          // a transient balance somewhere in the column: general path
    mem.park(X_TT, TT); mem.park(X_LT, Lt); mem.park(X_STL, stl); mem.park(X_SSW, ssw);
    mem.park_done();
  }
  MC_TICK(3);
  // ======================= heat and vapour exchange on the merged nodes (code.R:813-867) =======================
  // Solver storage: coefficients in the block's idle stage buffers, replaced in place by the elimination factors; ONE Thomas call site
  // per solve for both branches (equilibrium = no merged air node), so warps with mixed lanes do not run the solves twice.
  int nn = 2, mgx = 0;
  double TX, y1t = 0, y2t = 0, y1e = 0, y2e = 0, tns1_raw = 0, tsel_raw = 0;
  const bool wide = ng > kNG;                         // this lane's step has more merged layers than the stage-buffer solve holds
  LaneScr lscr;                                       // (thread-private; touched only by wide lanes; its node arrays are read by sweep C)
  {
    double x0[8], x1[8], x2[8];
    mem.unpark8(X_TAIRN, x0); mem.unpark8(X_PPK, x1); mem.unpark8(X_TSOILP, x2);
    const double pkn = x0[X_PKN], relhumn = x0[X_RELHUMN], tcan = x0[X_TCAN], tairp = x0[X_TAIRP], relhump = x0[X_RELHUMP];
    const double ppk = x1[X_PPK - X_PPK], ps1 = x1[X_PS1 - X_PPK], tet_tcan = x1[X_TET_TCAN - X_PPK], tabove = x1[X_TABOVE - X_PPK];
    const double tsoilp = x2[X_TSOILP - X_TSOILP], pha = x2[X_PHA - X_TSOILP], cs = x2[X_CS - X_TSOILP], gtncap = x2[X_GTNCAP - X_TSOILP];
    const double Xs = mem.unpark(X_XS), TT = mem.unpark(X_TT), lam = mem.unpark(X_LAM), Cv = mem.unpark(X_CV);
    const double esoil = mem.unpark(KT_ESOIL), eamn = mem.unpark(KU_EAMN);          // (all tensor-memory reads before lanes diverge)
    const double timestep = cfg.timestep, i2ts = 1 / (2 * timestep);
    TX = TT + fdiv(pha, gtncap) * 2;
    mem.g2_wait();                                     // the soil rows were fetched behind the last sweep-B stage
    if (!MC_SKIP(4)) {
      // the solve on storage `scr` (stage-buffer rows, or the lane's own arrays) with row map MAP and capacity CAP merged layers
      auto solve = [&](auto& scr, auto map_tag, auto cap_tag) {
      using MAP = decltype(map_tag);
      constexpr int CAP = decltype(cap_tag)::value;
      mgx = (ng > 1) ? ng : 0;
      const int mg = mgx, N = mg + sm;
      double ltc[CAP + 2], lgt[CAP + 3], lz[CAP + 4], lVo[CAP + 2], lea[CAP + 2], lph[CAP + 2], lTT[CAP + 2], xv[CAP + 3], vX[CAP + 2], Yt[CAP + 3];
      {                                                  // soil nodes FIRST: their rows (G2) sit in stage rows the scratch stores below reuse;
                                                         // all row loads, then the scratch stores (own lane only: no cross-lane hazard)
        double rk[MS], rc[MS], rt[MS];
#pragma unroll
        for (int j = 1; j <= MS; ++j) if (j <= sm) { rk[j - 1] = mem.g2(j - 1); rc[j - 1] = mem.g2(sm + (j - 1)); rt[j - 1] = mem.g2(2 * sm + (j - 1)); }
#pragma unroll
        for (int j = 1; j <= MS; ++j) if (j <= sm) {
          scr.scr_put(MAP::K0 + mg + 1 + j, lam * rk[j - 1]); scr.scr_put(MAP::CD0 + mg + j, (Cv * rc[j - 1]) * i2ts); scr.scr_put(MAP::TC0 + 1 + mg + j, rt[j - 1]);
        }
      }
      if (mg > 0) {                                   // layeraverage (code.R:814) and the air part of the system (code.R:815-826)
        int cprev = 0;
        lz[1] = 0;
#pragma unroll 1
        for (int g = 1; g <= mg; ++g) {
          const double iw = frcp(gw[g]);
          ltc[g] = gtc[g] * iw; lVo[g] = gVo[g] * iw; lea[g] = gea[g] * iw;
          lph[g] = phair(ltc[g], ppk);
          const int c = (glo[g] + ghi[g] + 1) / 2;
          lz[g + 1] = P[(long long)L.pz(c) * kTile]; lTT[g] = wget(c, WL_TT);
          if (c - cprev == 1) lgt[g] = wget(c, WL_GTN);
          else {
            double R = 0;
#pragma unroll 1
            for (int j = cprev + 1; j <= c; ++j) R += frcp(wget(j, WL_GTN));
            lgt[g] = frcp(R);
          }
          cprev = c;
          xv[mg + 1 - g] = gX[g] * iw;
        }
        if (m + 1 - cprev == 1) lgt[mg + 1] = gtncap;
        else {
          double R = 0;
#pragma unroll 1
          for (int j = cprev + 1; j <= m; ++j) R += frcp(wget(j, WL_GTN));
          R += frcp(gtncap); lgt[mg + 1] = frcp(R);
        }
        lz[mg + 2] = p0(P0_HGT);
#pragma unroll 1
        for (int g = 1; g <= mg; ++g) {
          // sum of dz * vden over the group from the cumulative row (two direct reads per group, only on steps with merged layers)
          const double gvd = mem.p_cvd(ghi[g]) - (glo[g] > 1 ? mem.p_cvd(glo[g] - 1) : 0.0);
          const double lcp = cpair(ltc[g]), lvd = gvd * frcp(gw[g]);
          const double cda = lcp * lph[g] * (1 - lvd) * (lz[g + 2] - lz[g]) / 2 * timestep;       // Q10
          double ka = lgt[g] * lcp;
          if (ka > cda) ka = cda;
          scr.scr_put(MAP::K0 + mg + 2 - g, ka); scr.scr_put(MAP::CD0 + mg + 1 - g, cda); scr.scr_put(MAP::TC0 + 1 + mg + 1 - g, ltc[g]);
        }
        scr.scr_put(MAP::K0 + 1, lgt[mg + 1] * cpair(tabove));
      } else {
        scr.scr_put(MAP::K0 + 1, frcp(cs) * cpair(tabove));   // one lumped conductance for the lower half of the canopy (code.R:856-857)
      }
      xv[mg + 1] = Xs;
      scr.scr_put(MAP::TC0 + 1, tabove); scr.scr_put(MAP::TC0 + N + 2, tsoilp);
      thomas_s(scr, MAP::K0, MAP::CD0, MAP::TC0, N, fc.tsoil, tcan, cfg.n, xv, mg + 1);
      // STC(2..N+1) now holds tn2 (code.R:829); tnsoil[j] = tn2[mg+1+j]
      const double tns1 = scr.scr_get(MAP::TC0 + mg + 2);
      double Vsoil = 0, Vair = 0;
      if (mg > 0) {
        // tnair = rev(tn2[1:(mg+2)]) = (soil 1, air bottom..top, tcan); ThomasV (code.R:153-169) on reversed arrays
        nn = mg + 2;
        Yt[mg + 2] = tcan;
        for (int g = 1; g <= mg + 1; ++g) Yt[g] = scr.scr_get(MAP::TC0 + mg + 3 - g);
        double rhs = soilrh(fc.theta, p0(P0_SOILB), p0(P0_PSIE), p0(P0_SMAX), tns1); if (rhs > 1) rhs = 1;
        double rhsp = soilrh(fc.thetap, p0(P0_SOILB), p0(P0_PSIE), p0(P0_SMAX), ps1); if (rhsp > 1) rhsp = 1;
        const double ipkn = frcp(pkn), ippk = frcp(ppk);
        Vair = (tet_tcan * fdiv(relhumn, 100)) * ipkn;
        const double eap = tetens(tairp) * fdiv(relhump, 100);
        Vsoil = (tetens(tns1) * rhs) * ipkn;
        const double easp = tetens(tsoilp) * rhsp;                               // Q5
        scr.scr_put(MAP::vTC0 + 1, eap * ippk); scr.scr_put(MAP::vTC0 + mg + 2, easp * ippk);
        for (int g = 1; g <= mg; ++g) scr.scr_put(MAP::vTC0 + 1 + g, lVo[mg + 1 - g]);
        for (int g = 1; g <= mg + 1; ++g) {
          double gt2 = lgt[g];
          if (g <= mg) { const double gtmx = fdiv(lph[g], lz[g + 1] - lz[g]) * (2 * timestep); if (gt2 > gtmx) gt2 = gtmx; }
          scr.scr_put(MAP::vK0 + mg + 2 - g, gt2);
        }
        for (int g = 1; g <= mg; ++g) {
          scr.scr_put(MAP::vCD0 + mg + 1 - g, phair(Yt[g + 1], pkn));
          double vf = (lea[g] * ipkn) - lVo[g]; if (vf < 0) vf = 0;
          vX[g] = vf;                                                            // Q6: not reversed
        }
      }
      thomas_s(scr, MAP::vK0, MAP::vCD0, MAP::vTC0, mg, Vsoil, Vair, cfg.n, vX, mg);           // (lanes without merged layers idle through it)
      // tnsoil (code.R:830, 859): nodes 2..sm go straight to the state rows (before the node arrays below reuse their scratch
      // rows); node 1 passes the limits at the end of the step first.  The node nearest to -reqhgt is kept for the harvest.
      {
        const int ss = (int)p0(P0_SELS);
        double tsel = tns1;
        for (int j = 2; j <= sm; ++j) {
          const double v = scr.scr_get(MAP::TC0 + mg + 1 + j);
          if (runB) sput(L.soiltc(j), v);
          if (j == ss) tsel = v;
        }
        tns1_raw = tns1; tsel_raw = tsel;
      }
      if (mg > 0) {
        // Vn = rev(vn): Ye[g] = vn[mg+3-g] with vn[1] = Vair, vn[mg+2] = Vsoil; mole fractions, scaled by pk after the interpolation
        double Ye[CAP + 3];
        Ye[1] = Vsoil; Ye[mg + 2] = Vair;
        for (int g = 2; g <= mg + 1; ++g) Ye[g] = scr.scr_get(MAP::vTC0 + mg + 3 - g);
        // node arrays for sweep C, in scratch rows the C stages do not touch
        for (int g = 1; g <= mg + 2; ++g) { scr.scr_put(MAP::YT0 + g, Yt[g]); scr.scr_put(MAP::YE0 + g, Ye[g]); }
        scr.scr_put(MAP::T20 + 1, 0.0);
        for (int g = 1; g <= mg; ++g) scr.scr_put(MAP::T20 + g + 1, lTT[g]);
        scr.scr_put(MAP::T20 + mg + 2, TX);
      } else {
        y1t = tns1; y2t = tcan; y1e = esoil; y2e = eamn;                           // eair = (relhum/100)*tetens(tcan), code.R:863
      }
      };
      if (!wide) solve(mem, ScrMapStage{}, std::integral_constant<int, kNG>{});
      else solve(lscr, ScrMapLane{}, std::integral_constant<int, kNGW>{});
    }
  }
  mem.park(KT_TNS1, tns1_raw); mem.park(KT_TSEL, tsel_raw); mem.park_done();
  MC_TICK(4);
  mem.sweepC_begin();
  mem.discard_AB();                                   // (after the barrier in sweepC_begin: every lane's solve has read its GTN rows)
  // ======================= sweep C: back to the m nodes, humidity, limits (code.R:846-874, 893-897) =======================
  double stc = 0, srh = 0, slw = 0, tnselG = 0;
  {
    const double trlw_sum = p0(P0_TRLWSUM), emv = p0(P0_EMV);
    const double skyem = mem.unpark(X_SKYEM), pkn = mem.unpark(X_PKN);
    const int selG = (int)p0(P0_SELG), selH = (int)p0(P0_SELH);
    const double iTX = frcp(TX);
    const double clw = trlw_sum * skyem + (1 - trlw_sum) * emv;
    // (Unrolling the kCL layers of a stage so that their branch-free tails interleave was tried and LOST 6-15 %: the sweep grows from 0.5 k to
    // 1.4 k instructions and the step no longer fits the instruction cache as well — tools/abv.sh cil1..cil5, profiles/README.md r02e.)
    for (int i = 1; i <= m; ++i) {
      mem.stageC(i);
      if (MC_SKIP(8)) continue;
      const double v = mem.c_w(0), lal = mem.c_w(1);             // lal = log(ean / 0.6108) from sweep B
      double tnn, ea2;
      if (nn == 2) {
        const double wgt = v * iTX;
        tnn = (1 - wgt) * y1t + wgt * y2t;
        ea2 = (1 - wgt) * y1e + wgt * y2e;
      } else if (!(v >= 0.0) || !(v <= TX)) { tnn = NAN; ea2 = NAN; }
      else {                                            // layerinterp: R approx(), bisection + linear (code.R:846-847)
        auto node = [&](int fast0, int wide0, int idx) { return wide ? lscr.a[wide0 + idx] : mem.scr_get(fast0 + idx); };
        int a = 1, b = nn;
        while (a < b - 1) { const int ab = (a + b) / 2; if (v < node(NT20, ScrMapLane::T20, ab)) b = ab; else a = ab; }
        const double T2a = node(NT20, ScrMapLane::T20, a), T2b = node(NT20, ScrMapLane::T20, b);
        const double Yta = node(NYT0, ScrMapLane::YT0, a), Ytb = node(NYT0, ScrMapLane::YT0, b);
        const double Yea = node(NYE0, ScrMapLane::YE0, a), Yeb = node(NYE0, ScrMapLane::YE0, b);
        if (v == T2b) { tnn = Ytb; ea2 = Yeb * pkn; }
        else if (v == T2a) { tnn = Yta; ea2 = Yea * pkn; }
        else {
          const double w = fdiv(v - T2a, T2b - T2a);
          tnn = Yta + (Ytb - Yta) * w;
          ea2 = (Yea + (Yeb - Yea) * w) * pkn;                                     // ea <- vn * pk (code.R:852)
        }
      }
      double r = fdiv(ea2, 0.6108 * fexp_b((17.27 * tnn) * frcp(tnn + 237.3))) * 100;
      if (r > 100) r = 100;
      if (r < 0) r = 0;
      double tdws;                                                               // dewpoint(tln$ea, tn), Q12
      {
        const bool ice = tnn < 0;
        // log(ean / e0) = log(ean / 0.6108) + log(0.6108 / e0), e0 = 0.61078 over ice, 0.6112 over water
        const double le0 = ice ? 3.27444784653359042e-05 : -6.54664507833277280e-04;
        // 1 / (1/273.15 - (461.5 / Lv) * L) = Lv / (Lv/273.15 - 461.5 L): one division instead of two
        const double Lv = ice ? 2.834e6 : 2.501e6 - 2340 * tnn;
        tdws = fdiv(Lv, Lv * (1 / 273.15) - 461.5 * (lal + le0)) - 273.15;
      }
      const double tn = (tnn < tdws) ? tdws : tnn;
      const double lwt = kSb * pow4(tn + 273.15);
      const double Rlwin = clw * lwt;                                             // (trlw_sum skyem + (1 - trlw_sum) emv) lwt
      if (i == selG) tnselG = tnn;
      if (runB) { mem.st_stream(S + (long long)L.sl(i, SL_TC) * kTile, tn); mem.st_stream(S + (long long)L.sl(i, SL_RH) * kTile, r); }
      if (emit || i == selH) wput(i, WL_RLWIN, Rlwin);
      stc += tn; srh += r; slw += Rlwin;
    }
  }
  MC_TICK(5);
  mem.discard_C();
  mem.end_begin(more);
  // ======================= soil limits and fluxes (code.R:868-891), output harvest (code.R:1222-1265) =======================
  double x0[8], x1[8];
  mem.unpark8(X_TAIRN, x0); mem.unpark8(X_PPK, x1);
  const double Lt = mem.unpark(X_LT), stl = mem.unpark(X_STL), ssw = mem.unpark(X_SSW);
  const double sqphi_e = mem.unpark(X_SQPHI), uh_e = mem.unpark(X_UH), cp1_e = mem.unpark(X_CP1), gtncap_e = mem.unpark(X_GTNCAP);
  const double tns1_in = mem.unpark(KT_TNS1), tsel_in = mem.unpark(KT_TSEL);
  if (!runB) return !run;                         // (after the last warp-collective tensor-memory access of the step)
  const double tairn = x0[X_TAIRN], pkn = x0[X_PKN], Rsw = x0[X_RSW], tcan = x0[X_TCAN];
  const double ps1 = x1[X_PS1 - X_PPK], tet_tair = x1[X_TET_TAIR - X_PPK], tet_tcan = x1[X_TET_TCAN - X_PPK], tfrost = x1[X_TFROST - X_PPK];
  const double phi_m = x1[X_PHIM - X_PPK];
  const double hgt = p0(P0_HGT), refg = p0(P0_REFG);
  double tns1 = tns1_in;
  if (tns1 < tfrost) tns1 = tfrost;
  if (tns1 > tairn + 30) tns1 = tairn + 30;
  const double refm = p0(P0_REFM);
  const double shade = (1 - fexp(-refm * p0(P0_SUMPAI)));
  const double alb = shade * refm + (1 - shade) * refg;
  const double Rlw = kSb * 0.97 * pow4(0.5 * tns1 + 0.5 * ps1 + 273.15);
  double Gn;
  {
    const double a = p0(P0_A0G) * sqphi_e;
    const double t1 = tnselG, t0 = tns1;
    const double ph = phair((t1 + t0) / 2, pkn);
    const double e0 = fexp(-a * (-1.0)), e1 = fexp(-a * p0(P0_GXG1));
    const double g = fdiv(p0(P0_CGG) * ph * uh_e * a, (e0 - e1) * phi_m);
    const double gfree = 0.0012 * ph * qroot(fabs(t1 - t0) * p0(P0_RZG));
    const double gt2 = (g > gfree) ? g : gfree;
    Gn = gt2 * cp1_e * (t1 - t0);
  }
  const double net = (1 - alb) * Rsw - Rlw;
  const double Gmx = fabs(net - Lt);
  if (Gn > Gmx) Gn = Gmx;
  if (Gn < -Gmx) Gn = -Gmx;
  const double Hn = net - Gn - Lt;
  sput(L.soiltc(1), tns1);
  sput(S0_TABOVE, tcan); sput(S0_RELHUM, x0[X_RELHUMN]); sput(S0_TAIR, tairn); sput(S0_TSOIL, fc.tsoil); sput(S0_PK, pkn); sput(S0_H, Hn);
  sput(S0_PSIM, x1[X_PSIM - X_PPK]); sput(S0_G, Gn); sput(S0_TCSUM, stc); sput(S0_GTCAP, gtncap_e);
  MC_TICK(6);
  double* o = out.o;
  out.ng = ng;
  o[6] = Hn; o[7] = Gn;
  if (charmode || reqhgt != reqhgt) { o[0] = stc / m; o[1] = stl / m; o[2] = srh / m; o[3] = ssw / m; o[4] = slw / m; o[5] = Lt; return fastok; }
  o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0;
  if (reqhgt >= hgt) {                                                         // Q13
    o[0] = tcan; o[1] = stl / m;
    double r = (((relhum0 / 100) * tet_tair) / tet_tcan) * 100; if (r > 100) r = 100;
    o[2] = r; o[3] = rsw0; o[4] = skyem0 * 5.67 * 1e-8 * pow4(tcan + 273.15); o[5] = Lt;
  }
  if (reqhgt >= 0 && reqhgt < hgt) {
    const int selH = (int)p0(P0_SELH);
    if (selH) {
      o[0] = s0(L.sl(selH, SL_TC)); o[1] = s0(L.sl(selH, SL_TLEAF)); o[2] = s0(L.sl(selH, SL_RH)); o[3] = wget(selH, WL_RSWIN);
      o[4] = wget(selH, WL_RLWIN); o[5] = wget(selH, WL_L);
    } else { for (int q = 0; q < 6; ++q) o[q] = NAN; }
  }
  if (reqhgt <= 0) { const int ss = (int)p0(P0_SELS); o[0] = (ss == 1) ? tns1 : tsel_in; o[1] = stl / m; o[2] = srh / m; o[3] = ssw / m; o[4] = slw / m; o[5] = Lt; }
  return fastok;
}

// ---- leaf-geometry rows after a first step whose previn$z differed from the current heights ---------------------------------
// leaftemp takes its layer geometry from previn$z (code.R:387-392); runonestep returns z = the current heights (code.R:719, 898),
// so from the second step on the rows follow the current heights.  One thread rewrites the six rows of its own column lane.
MC_NI void fast_regeom(double* P /* lane-adjusted */, int m, int sm, double timestep, double zlafact) {
  const FastLayout L{m, sm};
  auto g = [&](int r) { return P[(long long)r * kTile]; };
  const double hgt = g(P0_HGT), cpw = g(P0_CPW), phw = g(P0_PHW);
#pragma unroll 1
  for (int i = 1; i <= m; ++i) {
    const double zi = g(L.pz(i));
    const double zth = (i < m) ? g(L.pz(i + 1)) - zi : hgt - (zi + g(L.pz(m - 1))) / 2;
    const double PAIi = g(L.pb(i, PB_PAI));
    const double zl = mixinglength(zth, PAIi) * 0.5 * zlafact;
    const LeafGeom lg = leafgeom_rows(hgt, zi, zth, zl, PAIi, g(L.pb(i, PB_VDEN)), cpw, phw, timestep);
    P[(long long)L.pb(i, PB_IZTH) * kTile] = lg.izth; P[(long long)L.pb(i, PB_IZL) * kTile] = lg.izl;
    P[(long long)L.pb(i, PB_IZRF) * kTile] = lg.izrf; P[(long long)L.pb(i, PB_IZP) * kTile] = lg.izp;
    P[(long long)L.pb(i, PB_MAC) * kTile] = lg.mac; P[(long long)L.pb(i, PB_MZ) * kTile] = lg.mz;
  }
}

// ABI state rows of one column from the working rows (state layout of include/microclimc_b200.h)
MC_HD void abi_state_from_ws(const FastLayout& L, const double* __restrict__ P, const double* __restrict__ S, const double* __restrict__ W,
                             const double* __restrict__ soilz /* soilp z rows */, long long soil_ld, double* __restrict__ st, long long ld) {
  const int m = L.m, sm = L.sm;
  auto g = [&](int r) { return S[(long long)r * kTile]; };
  auto w = [&](int r, double v) { st[(long long)r * ld] = v; };
  w(P_TABOVE, g(S0_TABOVE)); w(P_ZABOVE, P[(long long)P0_ZABOVE * kTile]); w(P_RELHUM, g(S0_RELHUM)); w(P_TAIR, g(S0_TAIR)); w(P_TSOIL, g(S0_TSOIL));
  w(P_PK, g(S0_PK)); w(P_H, g(S0_H)); w(P_G, g(S0_G)); w(P_PSIM, g(S0_PSIM));
  int r = P_NSCAL;
  for (int i = 1; i <= m; ++i) w(r + i - 1, g(L.sl(i, SL_TC)));
  r += m; for (int i = 1; i <= m; ++i) w(r + i - 1, g(L.sl(i, SL_TLEAF)));
  r += m; for (int i = 1; i <= m; ++i) w(r + i - 1, g(L.uz(i)));
  r += m; for (int i = 1; i <= m; ++i) w(r + i - 1, P[(long long)L.pz(i) * kTile]);
  r += m; for (int i = 1; i <= m; ++i) w(r + i - 1, g(L.sl(i, SL_RH)));
  r += m; for (int i = 1; i <= m; ++i) w(r + i - 1, g(L.sl(i, SL_RABS)));
  r += m; for (int i = 1; i <= m; ++i) w(r + i - 1, W[(long long)L.wl(i, WL_GVN) * kTile]);
  r += m; for (int i = 1; i <= m; ++i) w(r + i - 1, g(L.sl(i, SL_GHA)));
  r += m; for (int i = 1; i <= m; ++i) w(r + i - 1, W[(long long)L.wl(i, WL_L) * kTile]);
  r += m; for (int i = 1; i <= m; ++i) w(r + i - 1, W[(long long)L.wl(i, WL_RSWIN) * kTile]);
  r += m; for (int i = 1; i <= m; ++i) w(r + i - 1, W[(long long)L.wl(i, WL_RLWIN) * kTile]);
  r += m; for (int i = 1; i <= m; ++i) w(r + i - 1, g(L.gt(i)));
  w(r + m, g(S0_GTCAP));
  r += m + 1; for (int j = 1; j <= sm; ++j) w(r + j - 1, g(L.soiltc(j)));
  r += sm; for (int j = 1; j <= sm; ++j) w(r + j - 1, soilz[(long long)(j - 1) * soil_ld]);
}

// ---- per-column configuration of a launch ---------------------------------------------------------------------------------
// runmodel maps reqhgt to the spin-up height and to reqhgt2 (code.R:1136-1140, 1199); spinup from runmodel drops zlafact /
// surfwet (Q22).  `spin` selects the spin-up variant.
MC_HD RunCfg effective_cfg(const TransientArgs& t, double hgt, bool spin) {
  const double reqhgt = t.cfg.reqhgt;
  const double spinhgt = t.charmode ? hgt / 2 : ((reqhgt == reqhgt && reqhgt > 0) ? reqhgt : NAN);
  RunCfg c = t.cfg;
  if (spin) { if (!t.spin_standalone) { c.reqhgt = spinhgt; c.zlafact = 1; c.surfwet = 1; } }
  else c.reqhgt = t.charmode ? spinhgt : reqhgt;
  return c;
}

// The launch runs on the regular node grid (template parameter REG of the streaming kernel) when its effective reqhgt is NA for every
// column: runonestep then keeps paraminit's z (code.R:573, 719).  `spin` selects the spin-up launch.  (With charmode the spin-up height
// is hgt / 2 per column, never NA.)
MC_HD bool fast_regular_grid(const TransientArgs& t, bool spin) {
  if (t.charmode) return false;
  if (!(t.cfg.timestep >= 3600.0)) return false;      // (shorter steps reach leaftemp's transient balances, which only the general instantiation has)
  const RunCfg c = effective_cfg(t, 1.0, spin);
  return !(c.reqhgt == c.reqhgt);
}

// mean(climdata$temp) of one column (code.R:992, 1165), summed in time order
MC_HD double column_tmean(const TransientArgs& a, long long c) {
  double s = 0;
  if (a.fscale && a.foffset) {
    const double sc = a.fscale[(long long)F_TAIR * a.ncol + c], of = a.foffset[(long long)F_TAIR * a.ncol + c];
    for (long long t = 0; t < a.T; ++t) s += a.forcing[t * F_N + F_TAIR] * sc + of;
  } else {
    for (long long t = 0; t < a.T; ++t) s += a.forcing[t * F_N + F_TAIR];
  }
  return s / (double)a.T;
}

// forcing_at (mc_driver.cuh) with the column's perturbation taken from the G0 rows (scale / offset were copied into the geometry rows by
// k_geom and fetched behind the end of the previous step: the 16 loads used to be a DRAM round trip at the start of every step)
template <class Mem>
MC_HD void fast_forcing_at(const Mem& mem, const TransientArgs& a, long long t, long long c, double* f) {
  const double* row = a.forcing + t * F_N;
#pragma unroll
  for (int v = 0; v < F_N; ++v) f[v] = row[v];
  if (a.fscale && a.foffset) {
#pragma unroll
    for (int v = 0; v < kNPert; ++v) f[v] = row[v] * mem.g0(P0_FS0 + v) + mem.g0(P0_FO0 + v);
    if (f[F_RELHUM] > 100) f[F_RELHUM] = 100;
    if (f[F_RELHUM] < 1) f[F_RELHUM] = 1;
    if (f[F_U] < 0) f[F_U] = 0;
    if (f[F_SKYEM] > 1) f[F_SKYEM] = 1;
    if (f[F_SKYEM] < 0) f[F_SKYEM] = 0;
    if (f[F_RSW] < 0) f[F_RSW] = 0;
    if (f[F_THETA] < 0.001) f[F_THETA] = 0.001;
  }
}

// The whole launch for one column: spin_mode 1 = paraminit + spin-up steps on forcing row 0 (code.R:982-1013);
// spin_mode 0 = steps [t0, t1) of runmodel's loop with the output harvest (code.R:1193-1266).
// `c` is the source column (clamped for padding lanes), `live` false for padding lanes.
template <class Mem, int MM, int MS, bool REG>
MC_HD void fast_run_column(const FastArgs& a, Mem& mem, long long c, bool live) {
  const TransientArgs& t = a.t;
  const FastLayout L = mem.L;
  const int m = L.m, sm = L.sm;
  double* S = mem.S; double* W = mem.W; const double* P = mem.P;
  auto sput = [=](int r, double v) { S[(long long)r * kTile] = v; };
  const double hgt = P[(long long)P0_HGT * kTile];
  const RunCfg cfg = effective_cfg(t, hgt, a.spin_mode != 0);
  const bool fastcol = P[(long long)P0_FAST * kTile] != 0.0;
  const bool zdiff = P[(long long)P0_ZDIFF * kTile] != 0.0;
  double f0[F_N], f[F_N];
  forcing_at(t, 0, c, f0);
  const bool have_tsoil = t.tsoil && t.tsoil[c] == t.tsoil[c];
  const double tmean = (a.spin_mode || !have_tsoil) ? column_tmean(t, c) : 0.0;
  const int NS = nstate_rows(m, sm);
  if (a.spin_mode) {
    // paraminit (code.R:566-585) straight into the working rows
    const double tair_ = f0[F_TAIR], u_ = f0[F_U], relhum_ = f0[F_RELHUM], tsoil_ = tmean, Rsw_ = f0[F_RSW];
    auto lin2 = [](double y1, double y2, int i, int n) { double xx = 1 + (double)(i - 1) / (double)(n - 1); return y1 + (y2 - y1) * (xx - 1); };
    double tcs = 0;
    const double g0 = 50.0 * ((double)m / hgt) * (2.0 / m);
    for (int i = 1; i <= m; ++i) {
      const double tci = lin2(tsoil_, tair_, sm + i, m + sm), Ra = lin2(0.2 * Rsw_ + 20, Rsw_ + 20, i, m);
      tcs += tci;
      sput(L.sl(i, SL_TC), tci); sput(L.sl(i, SL_TLEAF), tci + 0.01 * Ra); sput(L.sl(i, SL_RH), relhum_); sput(L.sl(i, SL_RABS), Ra);
      { const double gv = lin2(0.25, 0.32, i, m), gha = lin2(0.13, 0.19, i, m);
        sput(L.sl(i, SL_GV), fdiv(gv * gha, gv + gha)); sput(L.sl(i, SL_GHA), gha); }
      sput(L.uz(i), ((double)i / m) * u_);
      sput(L.gt(i), i == 1 ? 2 * g0 : g0);
    }
    const double gcap = g0 * (hgt / m) * 0.5;
    sput(L.gt(m + 1), gcap); sput(S0_GTCAP, gcap); sput(S0_TCSUM, tcs);
    for (int j = 1; j <= sm; ++j) sput(L.soiltc(j), lin2(tsoil_, tair_, sm + 1 - j, m + sm));
    sput(S0_TABOVE, tair_); sput(S0_RELHUM, relhum_); sput(S0_TAIR, tair_); sput(S0_TSOIL, tsoil_); sput(S0_PK, 101.3); sput(S0_H, 0.0);
    sput(S0_PSIM, 0.0); sput(S0_G, 0.0);
  } else {
    // (loads of the ABI state and stores of the working rows never alias: tell the compiler, so that it batches the loads)
    const double* __restrict__ const stp = t.state + c;
    auto st = [&](int r) { return stp[(long long)r * t.ncol]; };
    sput(S0_TABOVE, st(P_TABOVE)); sput(S0_RELHUM, st(P_RELHUM)); sput(S0_TAIR, st(P_TAIR)); sput(S0_TSOIL, st(P_TSOIL)); sput(S0_PK, st(P_PK));
    sput(S0_H, st(P_H)); sput(S0_PSIM, st(P_PSIM)); sput(S0_G, st(P_G));
    double tcs = 0;
    for (int i = 1; i <= m; ++i) {
      const double tci = st(P_NSCAL + i - 1);
      tcs += tci;
      sput(L.sl(i, SL_TC), tci); sput(L.sl(i, SL_TLEAF), st(P_NSCAL + m + i - 1)); sput(L.uz(i), st(P_NSCAL + 2 * m + i - 1));
      sput(L.sl(i, SL_RH), st(P_NSCAL + 4 * m + i - 1)); sput(L.sl(i, SL_RABS), st(P_NSCAL + 5 * m + i - 1));
      { const double gv = st(P_NSCAL + 6 * m + i - 1), gha = st(P_NSCAL + 7 * m + i - 1);      // (the same quotient the step forms: restarts stay bit-identical)
        sput(L.sl(i, SL_GV), fdiv(gv * gha, gv + gha)); sput(L.sl(i, SL_GHA), gha); }
      sput(L.gt(i), st(P_NSCAL + 11 * m + i - 1));
      W[(long long)L.wl(i, WL_L) * kTile] = st(P_NSCAL + 8 * m + i - 1); W[(long long)L.wl(i, WL_RSWIN) * kTile] = st(P_NSCAL + 9 * m + i - 1);
      W[(long long)L.wl(i, WL_RLWIN) * kTile] = st(P_NSCAL + 10 * m + i - 1);
    }
    sput(S0_TCSUM, tcs); sput(S0_GTCAP, st(P_NSCAL + 12 * m)); sput(L.gt(m + 1), st(P_NSCAL + 12 * m));
    for (int j = 1; j <= sm; ++j) sput(L.soiltc(j), st(P_NSCAL + 12 * m + 1 + j - 1));
  }
  const double reqhgt = t.cfg.reqhgt;
  StepOut so;
  // ONE loop (one inlined copy of the step) serves both modes; `spin` is launch-uniform.
  //   spin-up (code.R:982-1013): t.spin_steps steps on forcing row 0, tsoil = mean air temperature, previn$H <- running mean of H;
  //   runmodel (code.R:1193-1266): steps [t0, t1) with the output harvest.
  const bool spin = a.spin_mode != 0;
  const long long nsteps = spin ? (long long)t.spin_steps : t.t1 - t.t0;
  const double tsoil = spin ? tmean : (have_tsoil ? t.tsoil[c] : tmean);
  const double surfwet = cfg.surfwet;
  double stt[6];
  double Hsum = 0;
  int flag;
  stats_begin(t, c, stt, flag);
  const long long slot = (!spin && t.nsamp > 0 && t.samp_of) ? t.samp_of[c] : -1;
  // `dead`: padding lane, or a column that left the streaming form (now or in the spin-up launch of the same call): it keeps
  // computing with its stores suppressed (tensor-memory accesses and barriers are collective) and the general kernel redoes it.
  bool dead = !live || a.fail[c] != 0;
  auto drop = [&]() {
    dead = true;
    a.fail[c] = spin ? 1 : 2;
#if defined(__CUDA_ARCH__)
    const unsigned k = atomicAdd(a.fcount + (spin ? 0 : 1), 1u);
#elif !defined(__CUDACC__)
    const unsigned k = a.fcount[spin ? 0 : 1]++;
    ++g_mc_general_steps;
#else
    const unsigned k = 0;                               // (nvcc's host pass of a device-only path)
#endif
    a.flist[(spin ? 0 : t.ncol) + k] = c;
  };
  if (!dead && !fastcol) drop();
  // with shared per-layer rows (parameter classes) a tile cannot rebuild its leaf-geometry rows after the first step: such a column
  // (incoming node heights differ from the current geometry) goes to the general kernel
  if (!dead && zdiff && (a.tile_src || REG)) drop();       // (REG: the launch assumes the regular node grid in previn$z as well)
  fast_park_constants(mem);
  if (nsteps > 0) mem.g0_issue();
  for (long long it = 0; it < nsteps; ++it) {
    const long long tt = spin ? 0 : t.t0 + it;
    mem.g0_wait();
    fast_forcing_at(mem, t, tt, c, f);
    const Forcing fc = {f[F_TAIR], f[F_RELHUM], f[F_PK], f[F_U], tsoil, f[F_SKYEM], f[F_RSW], f[F_DP], f[F_JD], f[F_LT], f[F_THETA],
                        /*Q2*/ f0[F_THETA]};
    const bool emit = !dead && (it == nsteps - 1 || (slot >= 0 && t.full));
    SolDay sday;                                          // date-only solar terms: from the per-row table when the library made one
    if (t.fsol) { sday.eot = t.fsol[tt * 3]; sday.sdec = t.fsol[tt * 3 + 1]; sday.cdec = t.fsol[tt * 3 + 2]; }
    else sday = sol_day(f[F_JD]);
    const bool ok = fast_step<Mem, MS, REG>(mem, fc, sday, t.cfg, surfwet, !dead, emit, it + 1 < nsteps, reqhgt, t.charmode, f0[F_RELHUM], f0[F_RSW], f0[F_SKYEM], so);
    if (!dead && !ok) drop();
    if (!REG && it == 0 && zdiff && !a.tile_src) fast_regeom(const_cast<double*>(P), m, sm, cfg.timestep, cfg.zlafact);   // (ordered before the next step's bulk copies by step_begin's fence + barrier)
    if (dead) continue;
    if (spin) { Hsum += S[(long long)S0_H * kTile]; sput(S0_H, Hsum / (double)(it + 1)); continue; }   // previn$H <- mean(H)
    if (so.ng > 1) flag |= 2;
    const double* o = so.o;
    stt[0] += o[0]; if (o[0] < stt[1]) stt[1] = o[0]; if (o[0] > stt[2]) stt[2] = o[0];
    stt[3] += o[1]; if (o[1] < stt[4]) stt[4] = o[1]; if (o[1] > stt[5]) stt[5] = o[1];
    if (!isfinite(o[0]) || !isfinite(o[1])) flag |= 1;
    if (slot >= 0) {
      if (t.series) for (int q = 0; q < 8; ++q) t.series[(it * 8 + q) * t.nsamp + slot] = o[q];
      if (t.full) abi_state_from_ws(L, P, S, W, t.soilp + (long long)S_NSCAL * t.ncol + c, t.ncol, t.full + it * NS * t.nsamp + slot, t.nsamp);
    }
  }
  if (!dead && !spin) stats_end(t, c, stt, flag);
  if (!dead) abi_state_from_ws(L, P, S, W, t.soilp + (long long)S_NSCAL * t.ncol + c, t.ncol, t.state + c, t.ncol);
}

}  // namespace mc
