// C-ABI implementation (include/microclimc_b200.h): context, static column split over GPUs, kernel launches.
// One host thread + one stream per GPU, no collective, no CPU fallback.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <vector>
#include <functional>
#include <algorithm>
#include "../../include/microclimc_b200.h"
#include "mc_driver.cuh"
#include "mc_fast.cuh"
#include "mc_steady.cuh"

using namespace mc;

// ---------------------------------------------------------------------------------------------------
// kernels: one column per thread; the column index is the fastest-varying index of every array, so each
// warp-wide load/store touches 32 consecutive doubles (two full 128-B lines).
// ---------------------------------------------------------------------------------------------------
constexpr int kBlock = 128;

// `list` null: every column of the shard.  Otherwise the columns list[0 .. *count) — the ones that left the streaming form
// (mc_fast.cuh); blocks beyond the count exit at once, the last block's spare threads shadow its last column.
template <int MM, int MS>
__global__ void __launch_bounds__(kBlock) k_transient(TransientArgs a, const long long* __restrict__ list, const unsigned* __restrict__ count) {
  long long n = a.ncol;
  if (list) { n = *count; if ((long long)blockIdx.x * kBlock >= n) return; }
  exp_tab_init();
  const long long j = (long long)blockIdx.x * kBlock + threadIdx.x;
  const bool live = j < n;
  const long long jj = live ? j : n - 1;
  run_column<MM, MS>(a, list ? list[jj] : jj, live);
}
// Streaming form of the transient run (mc_fast.cuh): geometry once per launch, then the time loop with TMA-staged layers.
template <int MM, int MS>
__global__ void __launch_bounds__(kTile) k_geom(const __grid_constant__ FastArgs a) {
  exp_tab_init();
  const long long c = (long long)blockIdx.x * kTile + threadIdx.x;
  const long long cl = c < a.t.ncol ? c : a.t.ncol - 1;
  const double hgt = a.t.vegp[(long long)V_HGT * a.t.ncol + cl];
  geom_column<MM, MS>(a, effective_cfg(a.t, hgt, a.spin_mode != 0), c, cl);
}
// Only the forcing-perturbation rows of the geometry workspace (16 of its ~670 rows): what a new mc_set_forcing changes.
__global__ void __launch_bounds__(kTile) k_geom_pert(const __grid_constant__ FastArgs a) {
  const long long c = (long long)blockIdx.x * kTile + threadIdx.x;
  const long long cl = c < a.t.ncol ? c : a.t.ncol - 1;
  const FastLayout L{a.t.m, a.t.sm};
  double* P = a.P + (long long)blockIdx.x * L.nP() * kTile + threadIdx.x;
  const bool pert = a.t.fscale && a.t.foffset;
  for (int v = 0; v < kNPert; ++v) {
    P[(long long)(P0_FS0 + v) * kTile] = pert ? a.t.fscale[(long long)v * a.t.ncol + cl] : 1.0;
    P[(long long)(P0_FO0 + v) * kTile] = pert ? a.t.foffset[(long long)v * a.t.ncol + cl] : 0.0;
  }
}
constexpr int kFastSmem = 2 * kStageRows * kTile * 8 + 64;
// EXACT: the layer counts equal the template bounds (m = MM, sm = MS, all BASELINE configs), so every row offset is a
// compile-time constant; otherwise they are read from the arguments (6 < m <= MM, sm <= MS).
template <int MM, int MS, bool EXACT, bool REG>
// CTAs per SM the register allocation is bounded for.  4 (128 registers) is what shared memory (55 KB per CTA) and tensor memory (128 columns
// per CTA) allow as well.  Measured around it on the same box (tools/abv.sh, profiles/README.md r02s): 5 (96 registers, still 4 resident) runs
// 12 % slower, 3 (168 registers, no spills, 3 resident) 4 % slower — each warp is 28 % faster with the registers, which does not make up for
// a quarter fewer warps.
#ifndef MC_FAST_CTAS
#define MC_FAST_CTAS 4
#endif
__global__ void __launch_bounds__(kTile, MC_FAST_CTAS) k_transient_fast(const __grid_constant__ FastArgs a) {
  exp_tab_init();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* buf = reinterpret_cast<double*>(smem_raw);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw + 2 * kStageRows * kTile * 8);
  unsigned* tmem_base_s = reinterpret_cast<unsigned*>(bar + 3);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_init(bar + 2, 1);       // two stage barriers, one for the scalar groups
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {         // one warp allocates the block's 128 tensor-memory columns (64 doubles per thread of scratch)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_s)), "n"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const long long c = (long long)blockIdx.x * kTile + threadIdx.x;
  const bool live = c < a.t.ncol;
  FastLayout L{EXACT ? MM : a.t.m, EXACT ? MS : a.t.sm};
  PipeMemT<REG> mem;
  mem.L = L;
  mem.Pt = a.P + (long long)blockIdx.x * L.nP() * kTile; mem.St = a.S + (long long)blockIdx.x * L.nS() * kTile;
  mem.Wt = a.W + (long long)blockIdx.x * L.nW() * kTile;
  mem.Pl = a.tile_src ? a.P + (long long)a.tile_src[blockIdx.x] * L.nP() * kTile : mem.Pt;
  mem.Pll = mem.Pl + threadIdx.x;
  mem.P = mem.Pt + threadIdx.x; mem.S = const_cast<double*>(mem.St) + threadIdx.x; mem.W = const_cast<double*>(mem.Wt) + threadIdx.x;
  mem.buf = buf; mem.bar = bar; mem.bar32 = smem_u32(bar); mem.buf32 = smem_u32(buf); mem.lane = threadIdx.x; mem.k = 0; mem.gk = 0; mem.g0p = mem.g1ap = mem.g1bp = buf; mem.cur = buf;
  mem.tbase = *tmem_base_s + ((unsigned)((threadIdx.x >> 5) * 32) << 16);
  for (int q = 0; q < 8; ++q) mem.tcyc[q] = 0;
  mem.tlast = clock64();
  fast_run_column<PipeMemT<REG>, MM, MS, REG>(a, mem, live ? c : a.t.ncol - 1, live);
  if (a.dbgbuf && threadIdx.x == 0) for (int q = 0; q < 8; ++q) atomicAdd(a.dbgbuf + q, (unsigned long long)mem.tcyc[q]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_base_s), "n"(128) : "memory");
}
template <int MM, int MS>
__global__ void __launch_bounds__(kBlock) k_onestep(OneStepArgs a) {
  exp_tab_init();
  long long c = (long long)blockIdx.x * kBlock + threadIdx.x;
  if (c < a.ncol) onestep_column<MM, MS>(a, c);
}
template <int MM, int MS>
__global__ void __launch_bounds__(kBlock) k_paraminit(long long ncol, int m, int sm, const double* args, double* state) {
  exp_tab_init();
  long long c = (long long)blockIdx.x * kBlock + threadIdx.x;
  if (c < ncol) paraminit_column<MM, MS>(ncol, m, sm, args, state, c);
}
#ifndef MC_STEADY_CTAS
#define MC_STEADY_CTAS 5
#endif
template <int MM>
__global__ void __launch_bounds__(kBlock, MC_STEADY_CTAS) k_steady(SteadyArgs a) {
  exp_tab_init();
  long long c = (long long)blockIdx.x * kBlock + threadIdx.x;
  const bool live = c < a.ncol;
  // (Keeping the hour's 36 prepared column values in shared memory instead of thread-private storage was tried: 7.0 -> 8.0 ms on config 2 —
  // the L1-cached local loads the 96-register build makes are as cheap as LDS and the compiler keeps part of them in registers; r02j.)
  steady_column<MM>(a, live ? c : a.ncol - 1, blockIdx.y, gridDim.y, live);
}
__global__ void k_samp_scatter(long long ns, const long long* __restrict__ cols, long long* __restrict__ samp_of) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < ns) samp_of[cols[k]] = k;
}
// date-only part of the solar geometry, once per row of the series instead of once per (column, row); `stride` doubles per row,
// Julian day at `jdcol`
__global__ void k_sol_rows(long long T, const double* rows, int stride, int jdcol, double* sol) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const SolDay d = sol_day(rows[t * stride + jdcol]);
  sol[t * 3] = d.eot; sol[t * 3 + 1] = d.sdec; sol[t * 3 + 2] = d.cdec;
}
// hour-only quantities of the steady path, once per hour instead of once per (column, hour)
__global__ void k_steady_hour(long long T, const double* clim, double* out) {
  exp_tab_init();                                            // satvap uses fexp's shared-memory table
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const HourPre h = hour_pre(clim + t * 9);
  double* o = out + t * 13;
  o[0] = h.sd.eot; o[1] = h.sd.sdec; o[2] = h.sd.cdec; o[3] = h.dp; o[4] = h.ph; o[5] = h.ph0;
  o[6] = h.es; o[7] = h.eref; o[8] = h.cp; o[9] = h.p4; o[10] = h.gS; o[11] = h.gT; o[12] = h.lpk;
}
__global__ void k_steady_reduce(long long ncol, int nchunk, long long T, const double* part, double* stats) {
  long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncol) return;
  double s0 = 0, mn0 = 1e300, mx0 = -1e300, s1 = 0, mn1 = 1e300, mx1 = -1e300;
  for (int k = 0; k < nchunk; ++k) {
    const double* p = part + (long long)k * 6 * ncol;
    s0 += p[0 * ncol + c]; mn0 = dmin(mn0, p[1 * ncol + c]); mx0 = dmax(mx0, p[2 * ncol + c]);
    s1 += p[3 * ncol + c]; mn1 = dmin(mn1, p[4 * ncol + c]); mx1 = dmax(mx1, p[5 * ncol + c]);
  }
  stats[0 * ncol + c] = s0 / (double)T; stats[1 * ncol + c] = mn0; stats[2 * ncol + c] = mx0;
  stats[3 * ncol + c] = s1 / (double)T; stats[4 * ncol + c] = mn1; stats[5 * ncol + c] = mx1;
}

// Diagnostic: the fast FP64 primitives of mc_phys.cuh on plain arrays (tests bound their error in ulp).
__global__ void k_eval_primitive(int op, long long n, const double* x, const double* y, double* out) {
  exp_tab_init();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double a = x[i];
  double r;
  switch (op) {
    case 0: r = mc::fexp(a); break;
    case 1: r = mc::flog(a); break;
    case 2: r = mc::frcp(a); break;
    case 3: r = mc::fsqrt(a); break;
    default: r = mc::fdiv(a, y[i]); break;
  }
  out[i] = r;
}

// FP64 pipe peak probe: 8 independent DFMA chains per thread (the roofline denominator for this FP64-bound path;
// MEASURED_PEAKS.json has no FP64 entry).
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 16
  for (int i = 0; i < iters; ++i) {                  // unrolled: the loop bookkeeping would otherwise take 3 issue slots per 8 DFMA
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// ---------------------------------------------------------------------------------------------------
namespace {

std::string g_create_err;

struct DBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t need(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return (T*)p; }
};

// One persistent host thread per GPU of a multi-GPU context (created in mc_create, joined in mc_destroy): the calling thread posts
// the same job to every worker and waits for all of them, so no thread is spawned per API call and each worker keeps its device
// current.  The library's entry points are not re-entrant per context (one caller at a time), which is all the hand-off assumes.
// pinned host staging (sampled series travel through it before they are scattered into the caller's arrays)
struct HBuf {
  void* p = nullptr; size_t cap = 0;
  cudaError_t need(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocPortable);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return (T*)p; }
};

struct Worker {
  std::thread th; std::mutex mu; std::condition_variable cv;
  const std::function<int(struct Shard&)>* job = nullptr;
  int rc = 0; bool busy = false, quit = false;
};

struct Shard {
  int dev = 0; long long c0 = 0, nc = 0;
  Worker* worker = nullptr;
  std::vector<int64_t> samp_cached;            // sample columns whose slots the device-side samp_of currently holds
  bool samp_valid = false;
  HBuf hser, hfull;
  cudaStream_t st = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
  DBuf vegp, soilp, lat, lon, state, state_bak, fscale, foffset, forcing, season, tsoil, stats, series, full, samp_of, flags,
      args6, clim8, theta, thetap, nmerged, sclim, snmr, sseason, sout, spart, ssol, fsol, fP, fS, fW, dbg, ffail, flist, fcount, samp_cols, vegt, thetaz, stats_raw, flags_raw, tile_src;
  bool fast_attr = false;
  // Geometry cache of the streaming path.  k_geom's output (the P workspace: 3.9 ms per 10^6 columns and launch) depends on the column inputs
  // (vegp, soilp, lat, lon: `inputs_epoch` counts their uploads; the forcing perturbation: `pert_epoch`, and a launch after a new
  // mc_set_forcing alone rewrites just those 16 rows, k_geom_pert), the run description, the launch kind (spin-up or run) and
  // previn$z of the incoming state.  A run launch re-uses the rows of the previous run launch when all of that is unchanged AND both launches
  // are "clean": the incoming node heights equal the launch's own heights, which holds when the state was last written by a launch with the same
  // effective reqhgt (runonestep returns z = its current heights, code.R:898) — then no column has differing previn$z and the kernel does not
  // rebuild rows in place.  A year run week by week computes the geometry once instead of 52 times.
  struct GeomTag { double c[10]; int metopen, charmode, standalone, m, sm; long long ncol; unsigned long long epoch; };
  GeomTag geom_tag{}; bool geom_valid = false; unsigned long long geom_pert_epoch = 0;
  unsigned long long inputs_epoch = 1, pert_epoch = 1;
  bool z_valid = false; double z_reqhgt = 0; int z_charmode = 0; unsigned long long z_epoch = 0;   // whose heights the state's z rows hold
  bool bak_z_valid = false; double bak_z_reqhgt = 0; int bak_z_charmode = 0; unsigned long long bak_z_epoch = 0;   // the same for state_bak
  bool has_pert = false, has_season = false, has_state = false, bak_valid = false;
  long long nvt = 0, vegt_T = 0, thetaz_T = 0; bool has_thetaz = false;
  bool has_classes = false; long long class_tiles_shared = 0;   // mc_set_classes: tile_src loaded; tiles that read another tile's rows       // time-varying vegetation / soil moisture by depth loaded (mc_set_vegt, mc_set_thetaz)
  float ms = 0; long long launches = 0, pending_launches = 0;   // pending: kernels of mc_set_forcing, reported with the next run
  // replay info for mc_bench_resident
  int last_kind = 0;   // 0 none, 1 transient, 2 steady
  TransientArgs last_t; SteadyArgs last_s; int last_nchunk = 1;
  std::string err;
};

}  // namespace

struct mc_ctx {
  int m = 0, sm = 0;
  long long ncol = 0, T = 0;
  std::vector<Shard> shards;
  double cfg[MC_C_N_];
  bool has_cfg = false;
  std::string err;
  double last_ms = 0; long long last_launches = 0;
  bool stats_accum = false; long long stats_n = 0;   // mc_set_stats_accumulate: statistics carried across mc_run_transient calls, steps in them
  bool keep_replay_state = false;   // set by the first mc_bench_resident call: transient runs then keep a copy of their incoming state
};

namespace {

int fail(mc_ctx* ctx, int code, const std::string& msg) { if (ctx) ctx->err = msg; return code; }

#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { s.err = std::string(#call) + ": " + cudaGetErrorString(e_); return MC_ERR_CUDA; } } while (0)

void worker_main(Shard* s) {
  cudaSetDevice(s->dev);
  Worker& w = *s->worker;
  std::unique_lock<std::mutex> lk(w.mu);
  for (;;) {
    w.cv.wait(lk, [&] { return w.busy || w.quit; });
    if (w.quit) return;
    lk.unlock();
    const int rc = (*w.job)(*s);
    lk.lock();
    w.rc = rc; w.busy = false;
    w.cv.notify_all();
  }
}
// run fn on every shard — inline for one GPU, on the per-GPU worker threads otherwise — and return when all are done
int for_shards(mc_ctx* ctx, const std::function<int(Shard&)>& fn) {
  std::vector<int> rc(ctx->shards.size(), 0);
  if (ctx->shards.size() == 1) {
    cudaSetDevice(ctx->shards[0].dev);
    rc[0] = fn(ctx->shards[0]);
  } else {
    for (auto& s : ctx->shards) {
      std::lock_guard<std::mutex> lk(s.worker->mu);
      s.worker->job = &fn; s.worker->busy = true;
      s.worker->cv.notify_all();
    }
    for (size_t i = 0; i < ctx->shards.size(); ++i) {
      Worker& w = *ctx->shards[i].worker;
      std::unique_lock<std::mutex> lk(w.mu);
      w.cv.wait(lk, [&] { return !w.busy; });
      rc[i] = w.rc;
    }
  }
  for (size_t i = 0; i < rc.size(); ++i)
    if (rc[i]) { ctx->err = "gpu " + std::to_string(ctx->shards[i].dev) + ": " + ctx->shards[i].err; return rc[i]; }
  return 0;
}

// copy rows x [c0, c0+nc) of a host [rows][ncol] array to a dense device [rows][nc] array (and back)
cudaError_t h2d_rows(double* dst, const double* src, long long rows, long long ncol, long long c0, long long nc, cudaStream_t st) {
  return cudaMemcpy2DAsync(dst, nc * sizeof(double), src + c0, ncol * sizeof(double), nc * sizeof(double), rows,
                           cudaMemcpyHostToDevice, st);
}
cudaError_t d2h_rows(double* dst, const double* src, long long rows, long long ncol, long long c0, long long nc, cudaStream_t st) {
  return cudaMemcpy2DAsync(dst + c0, ncol * sizeof(double), src, nc * sizeof(double), nc * sizeof(double), rows,
                           cudaMemcpyDeviceToHost, st);
}

RunCfg mk_cfg(const double* c) {
  RunCfg r;
  r.timestep = c[MC_C_TIMESTEP]; r.edgedist = c[MC_C_EDGEDIST]; r.reqhgt = c[MC_C_REQHGT]; r.zu = c[MC_C_ZU];
  r.merid = c[MC_C_MERID]; r.dst = c[MC_C_DST]; r.n = c[MC_C_N]; r.metopen = c[MC_C_METOPEN] != 0; r.windhgt = c[MC_C_WINDHGT];
  r.zlafact = c[MC_C_ZLAFACT]; r.surfwet = c[MC_C_SURFWET];
  r.dbg = 0;
#ifdef MC_ABLATE                                   // measurement builds only (make ABLATE=1): section-skip / phase-timer knobs
  const char* e = getenv("MC_DBG_SKIP");
  r.dbg = e ? atoi(e) : 0;
#endif
  return r;
}

template <class F> int dispatch_ms(int m, int sm, F&& f) {
  if (m <= 20 && sm <= 10) return f(std::integral_constant<int, 20>(), std::integral_constant<int, 10>());
  return f(std::integral_constant<int, 64>(), std::integral_constant<int, 32>());
}
inline unsigned nblocks(long long n) { return (unsigned)((n + kBlock - 1) / kBlock); }

// The streaming kernels cover time-invariant PAI, depth-uniform soil moisture and the small template (6 < m <= 20, sm <= 10), at any time
// step: hourly or longer steps without an output height run the regular-grid instantiation, everything else the general instantiation
// (all per-layer rows staged, leaftemp's transient balances included); steps with more than six merged air layers (the rule at short
// steps) are solved by the wide solve.  The rest runs the general one-column-per-thread kernel.  Measurement builds (make ABLATE=1) can
// force the general kernel with MC_FORCE_GENERAL=1.
bool fast_eligible(const TransientArgs& a) {
#ifdef MC_ABLATE
  const char* e = getenv("MC_FORCE_GENERAL");
  if (e && e[0] == '1') return false;
#endif
  return !a.season && !a.vegt && !a.thetaz && a.m <= 20 && a.m > kNG && a.sm <= 10;
}
// whose node heights the state holds after this call: those of its last launch (the run launch, else the spin-up launch)
static void note_state_heights(Shard& s, const TransientArgs& a) {
  const bool spin_last = !(a.t1 > a.t0);
  s.z_valid = true; s.z_reqhgt = effective_cfg(a, 1.0, spin_last).reqhgt; s.z_charmode = a.charmode; s.z_epoch = s.inputs_epoch;
}
static bool same_bits(double x, double y) { return std::memcmp(&x, &y, sizeof(double)) == 0; }
int launch_transient(Shard& s, const TransientArgs& a) {
  if (!fast_eligible(a)) {
    int rc = dispatch_ms(a.m, a.sm, [&](auto MM, auto MS) {
      k_transient<decltype(MM)::value, decltype(MS)::value><<<nblocks(a.ncol), kBlock, 0, s.st>>>(a, nullptr, nullptr);
      s.launches++;
      return 0;
    });
    note_state_heights(s, a);
    return rc;
  }
  const FastLayout L{a.m, a.sm};
  const long long ntile = (a.ncol + kTile - 1) / kTile;
  CU(s.fP.need(sizeof(double) * ntile * L.nP() * kTile)); CU(s.fS.need(sizeof(double) * ntile * L.nS() * kTile));
  CU(s.fW.need(sizeof(double) * ntile * L.nW() * kTile));
  CU(s.ffail.need(sizeof(int) * a.ncol)); CU(s.flist.need(sizeof(long long) * 2 * a.ncol)); CU(s.fcount.need(sizeof(unsigned) * 2));
  CU(cudaMemsetAsync(s.ffail.p, 0, sizeof(int) * a.ncol, s.st)); CU(cudaMemsetAsync(s.fcount.p, 0, sizeof(unsigned) * 2, s.st));
  if (!s.fast_attr) {
    CU(cudaFuncSetAttribute(k_transient_fast<20, 10, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFastSmem));
    CU(cudaFuncSetAttribute(k_transient_fast<20, 10, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFastSmem));
    CU(cudaFuncSetAttribute(k_transient_fast<20, 10, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFastSmem));
    s.fast_attr = true;
  }
  FastArgs fa{a, s.fP.as<double>(), s.fS.as<double>(), s.fW.as<double>(), nullptr, 0, nullptr, s.ffail.as<int>(), s.flist.as<long long>(),
              s.fcount.as<unsigned>()};
  // parameter classes: shared per-layer rows only in launches whose incoming node heights equal the geometry's (the spin-up launch
  // starts from paraminit's regular grid, which differs whenever reqhgt snaps a node: code.R:573 vs 719)
  const bool classes_ok_spin = !(a.cfg.reqhgt == a.cfg.reqhgt) && !a.charmode;
#ifdef MC_ABLATE
  if (a.cfg.dbg & 16) {
    CU(s.dbg.need(sizeof(unsigned long long) * 8));
    CU(cudaMemsetAsync(s.dbg.p, 0, sizeof(unsigned long long) * 8, s.st));
    fa.dbgbuf = s.dbg.as<unsigned long long>();
  }
#endif
  for (int spin = 1; spin >= 0; --spin) {
    if (spin ? a.spin_steps <= 0 : a.t1 <= a.t0) continue;
    fa.spin_mode = spin;
    fa.tile_src = (s.has_classes && (spin == 0 || classes_ok_spin)) ? s.tile_src.as<int>() : nullptr;
    // geometry: recomputed unless this is a run launch with the inputs, run description and (clean) node heights of the run launch that
    // filled the workspace last (see Shard::GeomTag)
    Shard::GeomTag tag{};
    {
      const RunCfg& c = a.cfg;
      const double cv[10] = {c.timestep, c.edgedist, c.reqhgt, c.zu, c.merid, c.dst, c.n, c.windhgt, c.zlafact, c.surfwet};
      std::memcpy(tag.c, cv, sizeof(cv));
      tag.metopen = c.metopen; tag.charmode = a.charmode; tag.standalone = a.spin_standalone; tag.m = a.m; tag.sm = a.sm;
      tag.ncol = a.ncol; tag.epoch = s.inputs_epoch;
    }
    const double zreq = effective_cfg(a, 1.0, spin != 0).reqhgt;
    // the state's z rows are this launch's heights: written by a launch of the same effective reqhgt (or, after this call's own spin-up
    // launch, by that one)
    const bool spun = spin == 0 && a.spin_steps > 0;
    const double zprev_req = spun ? effective_cfg(a, 1.0, true).reqhgt : s.z_reqhgt;
    const bool clean = spin == 0 && (spun || (s.z_valid && s.z_epoch == s.inputs_epoch && s.z_charmode == a.charmode)) && same_bits(zprev_req, zreq) && !a.charmode;
    const bool reuse = clean && s.geom_valid && std::memcmp(&tag, &s.geom_tag, sizeof(tag)) == 0;
    if (!reuse) {
      k_geom<20, 10><<<(unsigned)ntile, kTile, 0, s.st>>>(fa);
      s.launches++;
      s.geom_valid = clean; s.geom_tag = tag; s.geom_pert_epoch = s.pert_epoch;   // (a spin-up launch or one with differing previn$z leaves rows nobody may re-use)
    } else if (s.geom_pert_epoch != s.pert_epoch) {
      k_geom_pert<<<(unsigned)ntile, kTile, 0, s.st>>>(fa);
      s.launches++;
      s.geom_pert_epoch = s.pert_epoch;
    }
    // REG: every column of the launch is on paraminit's regular node grid (reqhgt NA): the rows that depend on the grid alone are not streamed
    if (a.m == 20 && a.sm == 10 && fast_regular_grid(a, spin != 0)) k_transient_fast<20, 10, true, true><<<(unsigned)ntile, kTile, kFastSmem, s.st>>>(fa);
    else if (a.m == 20 && a.sm == 10) k_transient_fast<20, 10, true, false><<<(unsigned)ntile, kTile, kFastSmem, s.st>>>(fa);
    else k_transient_fast<20, 10, false, false><<<(unsigned)ntile, kTile, kFastSmem, s.st>>>(fa);
    s.launches++;
  }
  note_state_heights(s, a);
  // Columns that left the streaming form are redone by the general kernel from the call's initial state: list 0 (left during the
  // spin-up launch) with the spin-up, list 1 (left during the run launch) from the state the spin-up launch stored.  Both launches
  // cover the worst case and their blocks beyond the list length exit at once (no host round trip for the count).
  for (int q = 0; q < 2; ++q) {
    if (q == 0 ? a.spin_steps <= 0 : a.t1 <= a.t0) continue;
    TransientArgs g = a;
    if (q == 1) g.spin_steps = 0;
    g.flag_or = 8;
    k_transient<20, 10><<<nblocks(a.ncol), kBlock, 0, s.st>>>(g, s.flist.as<long long>() + (q ? a.ncol : 0), s.fcount.as<unsigned>() + q);
    s.launches++;
  }
#ifdef MC_ABLATE
  if (fa.dbgbuf) {
    unsigned long long h[8];
    CU(cudaMemcpyAsync(h, fa.dbgbuf, sizeof(h), cudaMemcpyDeviceToHost, s.st));
    CU(cudaStreamSynchronize(s.st));
    unsigned long long tot = 0; for (int q = 0; q < 8; ++q) tot += h[q];
    const char* nm[8] = {"phase0", "sweepA+cap", "Bprep", "sweepB", "solve", "sweepC", "end", "between-steps"};
    for (int q = 0; q < 8; ++q) fprintf(stderr, "[mc phase] %-14s %6.2f %%\n", nm[q], 100.0 * (double)h[q] / (double)(tot ? tot : 1));
  }
#endif
  return 0;
}
int launch_steady(Shard& s, const SteadyArgs& a, int nchunk) {
  if (a.sol) {
    k_steady_hour<<<nblocks(a.T), kBlock, 0, s.st>>>(a.T, a.clim, const_cast<double*>(a.sol));
    s.launches++;
  }
  dim3 grid(nblocks(a.ncol), nchunk);
  if (a.m <= 20) k_steady<20><<<grid, kBlock, 0, s.st>>>(a);
  else k_steady<64><<<grid, kBlock, 0, s.st>>>(a);
  s.launches++;
  if (a.part) {
    k_steady_reduce<<<nblocks(a.ncol), kBlock, 0, s.st>>>(a.ncol, nchunk, a.T, a.part, a.stats);
    s.launches++;
  }
  return 0;
}
void collect_timing(mc_ctx* ctx) {
  double ms = 0; long long l = 0;
  for (auto& s : ctx->shards) { if (s.ms > ms) ms = s.ms; l += s.launches; }
  ctx->last_ms = ms; ctx->last_launches = l;
}

}  // namespace

extern "C" {

int mc_version(void) { return MC_ABI_VERSION; }
int mc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
int mc_nveg(int m) { return MC_V_NSCAL + 4 * m; }
int mc_nsoil(int sm) { return MC_S_NSCAL + sm; }
int mc_nstate(int m, int sm) { return nstate_rows(m, sm); }
const char* mc_last_error(const mc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

mc_ctx* mc_create(int m, int sm, int n_gpus, const int* gpu_ids) {
  if (m < 3) { g_create_err = "At least three canopy layers must be provided"; return nullptr; }
  if (m > 64 || sm > 32 || sm < 2) { g_create_err = "supported sizes: 3 <= m <= 64, 2 <= sm <= 32"; return nullptr; }
  int ndev = mc_device_count();
  if (ndev <= 0) { g_create_err = "no CUDA device visible: microclimc_b200 has no CPU fallback"; return nullptr; }
  if (n_gpus <= 0) n_gpus = 1;
  if (n_gpus > ndev) { g_create_err = "requested " + std::to_string(n_gpus) + " GPUs but only " + std::to_string(ndev) + " visible"; return nullptr; }
  mc_ctx* ctx = new mc_ctx();
  ctx->m = m; ctx->sm = sm;
  ctx->shards.resize(n_gpus);
  for (int i = 0; i < n_gpus; ++i) {
    Shard& s = ctx->shards[i];
    s.dev = gpu_ids ? gpu_ids[i] : i;
    if (cudaSetDevice(s.dev) != cudaSuccess || cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&s.e0) != cudaSuccess || cudaEventCreate(&s.e1) != cudaSuccess) {
      g_create_err = std::string("CUDA init failed on device ") + std::to_string(s.dev) + ": " + cudaGetErrorString(cudaGetLastError());
      delete ctx; return nullptr;
    }
  }
  if (n_gpus > 1)
    for (auto& s : ctx->shards) { s.worker = new Worker(); s.worker->th = std::thread(worker_main, &s); }
  std::memset(ctx->cfg, 0, sizeof(ctx->cfg));
  return ctx;
}
void mc_destroy(mc_ctx* ctx) {
  if (!ctx) return;
  for (auto& s : ctx->shards) {
    if (s.worker) {
      { std::lock_guard<std::mutex> lk(s.worker->mu); s.worker->quit = true; s.worker->cv.notify_all(); }
      s.worker->th.join();
      delete s.worker; s.worker = nullptr;
    }
    cudaSetDevice(s.dev);
    DBuf* bufs[] = {&s.vegp, &s.soilp, &s.lat, &s.lon, &s.state, &s.state_bak, &s.fscale, &s.foffset, &s.forcing, &s.season, &s.tsoil,
                    &s.stats, &s.series, &s.full, &s.samp_of, &s.flags, &s.args6, &s.clim8, &s.theta, &s.thetap, &s.nmerged,
                    &s.sclim, &s.snmr, &s.sseason, &s.sout, &s.spart, &s.ssol, &s.fsol, &s.fP, &s.fS, &s.fW, &s.dbg, &s.ffail, &s.flist,
                    &s.fcount, &s.samp_cols, &s.vegt, &s.thetaz, &s.stats_raw, &s.flags_raw, &s.tile_src};
    for (DBuf* b : bufs) b->release();
    s.hser.release(); s.hfull.release();
    if (s.e0) cudaEventDestroy(s.e0);
    if (s.e1) cudaEventDestroy(s.e1);
    if (s.st) cudaStreamDestroy(s.st);
  }
  delete ctx;
}

int mc_set_columns(mc_ctx* ctx, int64_t ncol, const double* vegp, const double* soilp, const double* lat, const double* lon) {
  if (!ctx || ncol <= 0 || !vegp || !soilp || !lat || !lon) return fail(ctx, MC_ERR_ARG, "mc_set_columns: bad argument");
  ctx->ncol = ncol;
  const int G = (int)ctx->shards.size();
  const long long per = (ncol + G - 1) / G;
  for (int i = 0; i < G; ++i) {
    Shard& s = ctx->shards[i];
    s.c0 = std::min<long long>((long long)i * per, ncol);
    s.nc = std::min<long long>(per, ncol - s.c0);
    s.has_state = false; s.has_pert = false; s.last_kind = 0; s.samp_valid = false; s.nvt = 0; s.has_thetaz = false; s.has_classes = false;
    s.inputs_epoch++; s.geom_valid = false; s.z_valid = false;
  }
  const int NV = mc_nveg(ctx->m), NSo = mc_nsoil(ctx->sm), NSt = mc_nstate(ctx->m, ctx->sm);
  return for_shards(ctx, [&](Shard& s) -> int {
    if (s.nc == 0) return 0;
    CU(s.vegp.need(sizeof(double) * NV * s.nc)); CU(s.soilp.need(sizeof(double) * NSo * s.nc));
    CU(s.lat.need(sizeof(double) * s.nc)); CU(s.lon.need(sizeof(double) * s.nc));
    CU(s.state.need(sizeof(double) * NSt * s.nc));
    CU(h2d_rows(s.vegp.as<double>(), vegp, NV, ncol, s.c0, s.nc, s.st));
    CU(h2d_rows(s.soilp.as<double>(), soilp, NSo, ncol, s.c0, s.nc, s.st));
    CU(h2d_rows(s.lat.as<double>(), lat, 1, ncol, s.c0, s.nc, s.st));
    CU(h2d_rows(s.lon.as<double>(), lon, 1, ncol, s.c0, s.nc, s.st));
    CU(cudaStreamSynchronize(s.st));
    return 0;
  });
}

int mc_set_forcing(mc_ctx* ctx, int64_t T, const double* forcing, const double* fscale, const double* foffset, const double* pai_season) {
  if (!ctx || T <= 0 || !forcing || ctx->ncol <= 0) return fail(ctx, MC_ERR_ARG, "mc_set_forcing: bad argument (set columns first)");
  if ((fscale == nullptr) != (foffset == nullptr)) return fail(ctx, MC_ERR_ARG, "mc_set_forcing: fscale and foffset must both be given");
  ctx->T = T;
  const long long ncol = ctx->ncol;
  return for_shards(ctx, [&](Shard& s) -> int {
    if (s.nc == 0) return 0;
    CU(s.forcing.need(sizeof(double) * T * MC_F_N));
    CU(cudaMemcpyAsync(s.forcing.p, forcing, sizeof(double) * T * MC_F_N, cudaMemcpyHostToDevice, s.st));
    CU(s.fsol.need(sizeof(double) * T * 3));
    k_sol_rows<<<nblocks(T), kBlock, 0, s.st>>>(T, s.forcing.as<double>(), MC_F_N, F_JD, s.fsol.as<double>());
    CU(cudaGetLastError());
    s.pending_launches++;
    s.has_pert = fscale != nullptr;
    s.pert_epoch++;                                  // (the perturbation rows are copied into the geometry workspace)
    if (fscale) {
      CU(s.fscale.need(sizeof(double) * MC_NPERT * s.nc)); CU(s.foffset.need(sizeof(double) * MC_NPERT * s.nc));
      CU(h2d_rows(s.fscale.as<double>(), fscale, MC_NPERT, ncol, s.c0, s.nc, s.st));
      CU(h2d_rows(s.foffset.as<double>(), foffset, MC_NPERT, ncol, s.c0, s.nc, s.st));
    }
    s.has_season = pai_season != nullptr;
    if (pai_season) {
      CU(s.season.need(sizeof(double) * T));
      CU(cudaMemcpyAsync(s.season.p, pai_season, sizeof(double) * T, cudaMemcpyHostToDevice, s.st));
    }
    CU(cudaStreamSynchronize(s.st));
    return 0;
  });
}

int mc_set_vegt(mc_ctx* ctx, int64_t T, int64_t nvt, const double* vegt) {
  if (!ctx || ctx->ncol <= 0) return fail(ctx, MC_ERR_ARG, "mc_set_vegt: set columns first");
  if (!vegt) { for (auto& s : ctx->shards) s.nvt = 0; return 0; }
  if (T <= 0) return fail(ctx, MC_ERR_ARG, "mc_set_vegt: T must be positive");
  if (nvt != 1 && nvt != ctx->ncol) return fail(ctx, MC_ERR_ARG, "mc_set_vegt: nvt must be 1 (shared series) or ncol");
  const long long rows = T * (2 * ctx->m + 1), ncol = ctx->ncol;
  return for_shards(ctx, [&](Shard& s) -> int {
    if (s.nc == 0) return 0;
    const long long n = (nvt == 1) ? 1 : s.nc;
    CU(s.vegt.need(sizeof(double) * rows * n));
    if (nvt == 1) CU(cudaMemcpyAsync(s.vegt.p, vegt, sizeof(double) * rows, cudaMemcpyHostToDevice, s.st));
    else CU(h2d_rows(s.vegt.as<double>(), vegt, rows, ncol, s.c0, s.nc, s.st));
    CU(cudaStreamSynchronize(s.st));
    s.nvt = n; s.vegt_T = T;
    return 0;
  });
}
int mc_set_thetaz(mc_ctx* ctx, int64_t T, const double* thetaz) {
  if (!ctx || ctx->ncol <= 0) return fail(ctx, MC_ERR_ARG, "mc_set_thetaz: set columns first");
  if (!thetaz) { for (auto& s : ctx->shards) s.has_thetaz = false; return 0; }
  if (T <= 0) return fail(ctx, MC_ERR_ARG, "mc_set_thetaz: T must be positive");
  if (ctx->sm > ctx->m) return fail(ctx, MC_ERR_ARG, "mc_set_thetaz: soil moisture by depth needs sm <= m (R recycles the soil vector over the canopy layers)");
  const int sm = ctx->sm;
  return for_shards(ctx, [&](Shard& s) -> int {
    if (s.nc == 0) return 0;
    CU(s.thetaz.need(sizeof(double) * T * sm));
    CU(cudaMemcpyAsync(s.thetaz.p, thetaz, sizeof(double) * T * sm, cudaMemcpyHostToDevice, s.st));
    CU(cudaStreamSynchronize(s.st));
    s.has_thetaz = true; s.thetaz_T = T;
    return 0;
  });
}

int mc_set_classes(mc_ctx* ctx, const int32_t* class_of) {
  if (!ctx || ctx->ncol <= 0) return fail(ctx, MC_ERR_ARG, "mc_set_classes: set columns first");
  if (!class_of) { for (auto& s : ctx->shards) s.has_classes = false; return 0; }
  return for_shards(ctx, [&](Shard& s) -> int {
    if (s.nc == 0) return 0;
    const long long ntile = (s.nc + kTile - 1) / kTile;
    std::vector<int> src(ntile);
    std::vector<std::pair<int32_t, int>> first;              // class id -> first class-uniform tile of this shard
    long long shared = 0;
    for (long long t = 0; t < ntile; ++t) {
      const long long c0 = s.c0 + t * kTile, c1 = std::min<long long>(c0 + kTile, s.c0 + s.nc);
      const int32_t id = class_of[c0];
      bool uniform = id >= 0;
      for (long long c = c0 + 1; c < c1 && uniform; ++c) uniform = class_of[c] == id;
      src[t] = (int)t;
      if (!uniform) continue;
      auto it = std::find_if(first.begin(), first.end(), [&](const std::pair<int32_t, int>& p) { return p.first == id; });
      if (it == first.end()) first.emplace_back(id, (int)t);
      else { src[t] = it->second; ++shared; }
    }
    CU(s.tile_src.need(sizeof(int) * ntile));
    CU(cudaMemcpyAsync(s.tile_src.p, src.data(), sizeof(int) * ntile, cudaMemcpyHostToDevice, s.st));
    CU(cudaStreamSynchronize(s.st));
    s.has_classes = true; s.class_tiles_shared = shared;
    return 0;
  });
}
int mc_set_stats_accumulate(mc_ctx* ctx, int on) {
  if (!ctx) return MC_ERR_ARG;
  ctx->stats_accum = on != 0; ctx->stats_n = 0;
  return 0;
}
int mc_set_config(mc_ctx* ctx, const double* cfg) {
  if (!ctx || !cfg) return fail(ctx, MC_ERR_ARG, "mc_set_config: bad argument");
  std::memcpy(ctx->cfg, cfg, sizeof(ctx->cfg));
  ctx->has_cfg = true;
  return 0;
}

int mc_set_state(mc_ctx* ctx, const double* state) {
  if (!ctx || !state || ctx->ncol <= 0) return fail(ctx, MC_ERR_ARG, "mc_set_state: bad argument (set columns first)");
  const int NSt = mc_nstate(ctx->m, ctx->sm);
  const long long ncol = ctx->ncol;
  return for_shards(ctx, [&](Shard& s) -> int {
    if (s.nc == 0) return 0;
    CU(h2d_rows(s.state.as<double>(), state, NSt, ncol, s.c0, s.nc, s.st));
    CU(cudaStreamSynchronize(s.st));
    s.has_state = true; s.z_valid = false;
    return 0;
  });
}
int mc_get_state(mc_ctx* ctx, double* state) {
  if (!ctx || !state || ctx->ncol <= 0) return fail(ctx, MC_ERR_ARG, "mc_get_state: bad argument");
  for (auto& s : ctx->shards) if (s.nc && !s.has_state) return fail(ctx, MC_ERR_STATE, "mc_get_state: no state on the device yet");
  const int NSt = mc_nstate(ctx->m, ctx->sm);
  const long long ncol = ctx->ncol;
  return for_shards(ctx, [&](Shard& s) -> int {
    if (s.nc == 0) return 0;
    CU(d2h_rows(state, s.state.as<double>(), NSt, ncol, s.c0, s.nc, s.st));
    CU(cudaStreamSynchronize(s.st));
    return 0;
  });
}

int mc_paraminit(mc_ctx* ctx, const double* args) {
  if (!ctx || !args || ctx->ncol <= 0) return fail(ctx, MC_ERR_ARG, "mc_paraminit: bad argument (set columns first)");
  const long long ncol = ctx->ncol;
  const int m = ctx->m, sm = ctx->sm;
  int rc = for_shards(ctx, [&](Shard& s) -> int {
    s.ms = 0; s.launches = 0;
    if (s.nc == 0) return 0;
    CU(s.args6.need(sizeof(double) * 6 * s.nc));
    CU(h2d_rows(s.args6.as<double>(), args, 6, ncol, s.c0, s.nc, s.st));
    CU(cudaEventRecord(s.e0, s.st));
    dispatch_ms(m, sm, [&](auto MM, auto MS) {
      k_paraminit<decltype(MM)::value, decltype(MS)::value><<<nblocks(s.nc), kBlock, 0, s.st>>>(s.nc, m, sm, s.args6.as<double>(), s.state.as<double>());
      return 0;
    });
    s.launches++;
    CU(cudaEventRecord(s.e1, s.st));
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s.st));
    CU(cudaEventElapsedTime(&s.ms, s.e0, s.e1));
    s.has_state = true; s.z_valid = false;
    return 0;
  });
  collect_timing(ctx);
  return rc;
}

int mc_run_onestep(mc_ctx* ctx, const double* climvars, double jd, double lt, const double* theta, const double* thetap, int32_t* nmerged) {
  if (!ctx || !climvars || !theta || !thetap || ctx->ncol <= 0) return fail(ctx, MC_ERR_ARG, "mc_run_onestep: bad argument");
  if (!ctx->has_cfg) return fail(ctx, MC_ERR_ARG, "mc_run_onestep: mc_set_config first");
  for (auto& s : ctx->shards) if (s.nc && !s.has_state) return fail(ctx, MC_ERR_STATE, "mc_run_onestep: no state (mc_paraminit / mc_set_state first)");
  const long long ncol = ctx->ncol;
  const int m = ctx->m, sm = ctx->sm;
  RunCfg cfg = mk_cfg(ctx->cfg);
  int rc = for_shards(ctx, [&](Shard& s) -> int {
    s.ms = 0; s.launches = 0;
    if (s.nc == 0) return 0;
    CU(s.clim8.need(sizeof(double) * 8 * s.nc)); CU(s.theta.need(sizeof(double) * s.nc)); CU(s.thetap.need(sizeof(double) * s.nc));
    CU(s.nmerged.need(sizeof(int) * s.nc));
    CU(h2d_rows(s.clim8.as<double>(), climvars, 8, ncol, s.c0, s.nc, s.st));
    CU(h2d_rows(s.theta.as<double>(), theta, 1, ncol, s.c0, s.nc, s.st));
    CU(h2d_rows(s.thetap.as<double>(), thetap, 1, ncol, s.c0, s.nc, s.st));
    OneStepArgs a{};
    a.ncol = s.nc; a.m = m; a.sm = sm; a.vegp = s.vegp.as<double>(); a.soilp = s.soilp.as<double>(); a.lat = s.lat.as<double>();
    a.lon = s.lon.as<double>(); a.climvars = s.clim8.as<double>(); a.theta = s.theta.as<double>(); a.thetap = s.thetap.as<double>();
    a.jd = jd; a.lt = lt; a.state = s.state.as<double>(); a.nmerged = s.nmerged.as<int>(); a.cfg = cfg;
    CU(cudaEventRecord(s.e0, s.st));
    dispatch_ms(m, sm, [&](auto MM, auto MS) {
      k_onestep<decltype(MM)::value, decltype(MS)::value><<<nblocks(s.nc), kBlock, 0, s.st>>>(a);
      return 0;
    });
    s.launches++;
    CU(cudaEventRecord(s.e1, s.st));
    CU(cudaGetLastError());
    if (nmerged) CU(cudaMemcpyAsync(nmerged + s.c0, s.nmerged.p, sizeof(int) * s.nc, cudaMemcpyDeviceToHost, s.st));
    CU(cudaStreamSynchronize(s.st));
    CU(cudaEventElapsedTime(&s.ms, s.e0, s.e1));
    s.z_valid = false;
    return 0;
  });
  collect_timing(ctx);
  return rc;
}

static int run_transient_impl(mc_ctx* ctx, int64_t t0, int64_t t1, int spin_steps, int standalone, const double* tsoil, const int64_t* samp_idx,
                              int64_t nsamp, double* series, double* fullseries, double* stats, int32_t* flags) {
  if (!ctx || ctx->ncol <= 0 || ctx->T <= 0) return fail(ctx, MC_ERR_ARG, "mc_run_transient: set columns and forcing first");
  if (!ctx->has_cfg) return fail(ctx, MC_ERR_ARG, "mc_run_transient: mc_set_config first");
  if (t0 < 0 || t1 < t0 || t1 > ctx->T) return fail(ctx, MC_ERR_ARG, "mc_run_transient: bad step range");
  if ((series || fullseries) && (!samp_idx || nsamp <= 0)) return fail(ctx, MC_ERR_ARG, "mc_run_transient: series requested without samp_idx");
  if (!ctx->cfg[MC_C_METOPEN]) {   // code.R:1141-1143 — checked per column on the host mirror; here a cheap global guard is not possible
  }
  if (spin_steps <= 0) for (auto& s : ctx->shards) if (s.nc && !s.has_state)
    return fail(ctx, MC_ERR_STATE, "mc_run_transient: no state on the device (spin-up disabled and no mc_set_state / mc_paraminit)");
  for (int64_t j = 0; j < nsamp; ++j) if (samp_idx[j] < 0 || samp_idx[j] >= ctx->ncol) return fail(ctx, MC_ERR_ARG, "mc_run_transient: samp_idx out of range");
  for (auto& s : ctx->shards) {
    if (s.nc && s.nvt && s.vegt_T != ctx->T) return fail(ctx, MC_ERR_ARG, "mc_run_transient: mc_set_vegt rows differ from the forcing rows");
    if (s.nc && s.has_thetaz && s.thetaz_T != ctx->T) return fail(ctx, MC_ERR_ARG, "mc_run_transient: mc_set_thetaz rows differ from the forcing rows");
  }
  if (nsamp > 1 && (series || fullseries)) {          // a column may be sampled once: its device slot is written once
    std::vector<int64_t> srt(samp_idx, samp_idx + nsamp);
    std::sort(srt.begin(), srt.end());
    if (std::adjacent_find(srt.begin(), srt.end()) != srt.end()) return fail(ctx, MC_ERR_ARG, "mc_run_transient: duplicate entries in samp_idx");
  }
  const long long ncol = ctx->ncol;
  const int m = ctx->m, sm = ctx->sm, NSt = mc_nstate(m, sm);
  const long long nt = t1 - t0;
  RunCfg cfg = mk_cfg(ctx->cfg);
  const int charmode = ctx->cfg[MC_C_HARVEST_CHAR] != 0;
  int rc = for_shards(ctx, [&](Shard& s) -> int {
    s.ms = 0; s.launches = s.pending_launches; s.pending_launches = 0;
    if (s.nc == 0) return 0;
    // sample slots owned by this shard: slot k of the shard <-> entry slots[k] of samp_idx <-> local column cols[k]
    std::vector<int64_t> slots, cols;
    if (nsamp > 0 && (series || fullseries))
      for (int64_t j = 0; j < nsamp; ++j) {
        long long c = samp_idx[j] - s.c0;
        if (c >= 0 && c < s.nc) { slots.push_back(j); cols.push_back(c); }
      }
    const long long ns = (long long)slots.size();
    TransientArgs a{};
    a.ncol = s.nc; a.m = m; a.sm = sm; a.T = ctx->T; a.t0 = t0; a.t1 = t1;
    a.vegp = s.vegp.as<double>(); a.soilp = s.soilp.as<double>(); a.lat = s.lat.as<double>(); a.lon = s.lon.as<double>();
    a.forcing = s.forcing.as<double>(); a.fsol = s.fsol.as<double>();
    a.fscale = s.has_pert ? s.fscale.as<double>() : nullptr; a.foffset = s.has_pert ? s.foffset.as<double>() : nullptr;
    a.season = s.has_season ? s.season.as<double>() : nullptr;
    a.vegt = s.nvt ? s.vegt.as<double>() : nullptr; a.nvt = s.nvt;
    a.thetaz = s.has_thetaz ? s.thetaz.as<double>() : nullptr;
    a.state = s.state.as<double>(); a.cfg = cfg; a.spin_steps = spin_steps; a.charmode = charmode; a.spin_standalone = standalone;
    if (tsoil) {
      CU(s.tsoil.need(sizeof(double) * s.nc));
      CU(h2d_rows(s.tsoil.as<double>(), tsoil, 1, ncol, s.c0, s.nc, s.st));
      a.tsoil = s.tsoil.as<double>();
    }
    if (ns > 0) {
      // column -> slot map on the device: filled with -1 and scattered from the short column list by a small kernel, and kept as long
      // as the sample set does not change (a weekly-chunked year re-uses it 52 times) — no per-call upload proportional to ncol
      if (!s.samp_valid || s.samp_cached != cols) {
        CU(s.samp_of.need(sizeof(long long) * s.nc)); CU(s.samp_cols.need(sizeof(long long) * ns));
        CU(cudaMemsetAsync(s.samp_of.p, 0xFF, sizeof(long long) * s.nc, s.st));
        CU(cudaMemcpyAsync(s.samp_cols.p, cols.data(), sizeof(long long) * ns, cudaMemcpyHostToDevice, s.st));
        k_samp_scatter<<<nblocks(ns), kBlock, 0, s.st>>>(ns, s.samp_cols.as<long long>(), s.samp_of.as<long long>());
        CU(cudaGetLastError());
        s.launches++;
        CU(cudaStreamSynchronize(s.st));                 // (cols is a local: the copy must have read it before it goes away)
        s.samp_cached = cols; s.samp_valid = true;
      }
      a.samp_of = s.samp_of.as<long long>(); a.nsamp = ns;
      if (series) { CU(s.series.need(sizeof(double) * nt * MC_O_N * ns)); a.series = s.series.as<double>(); }
      if (fullseries) { CU(s.full.need(sizeof(double) * nt * NSt * ns)); a.full = s.full.as<double>(); }
    }
    if (stats) { CU(s.stats.need(sizeof(double) * 6 * s.nc)); a.stats = s.stats.as<double>(); }
    if (flags) { CU(s.flags.need(sizeof(int) * s.nc)); a.flags = s.flags.as<int>(); }
    if (ctx->stats_accum && nt > 0) {
      CU(s.stats_raw.need(sizeof(double) * 6 * s.nc)); CU(s.flags_raw.need(sizeof(int) * s.nc));
      a.stats_raw = s.stats_raw.as<double>(); a.flags_raw = s.flags_raw.as<int>(); a.stats_prev_n = ctx->stats_n;
    }
    if (spin_steps <= 0 && ctx->keep_replay_state) {   // opt-in (mc_bench_resident): a copy of the incoming state to replay from
      CU(s.state_bak.need(sizeof(double) * NSt * s.nc));
      CU(cudaMemcpyAsync(s.state_bak.p, s.state.p, sizeof(double) * NSt * s.nc, cudaMemcpyDeviceToDevice, s.st));
      s.bak_valid = true;
      s.bak_z_valid = s.z_valid; s.bak_z_reqhgt = s.z_reqhgt; s.bak_z_charmode = s.z_charmode; s.bak_z_epoch = s.z_epoch;
    } else s.bak_valid = false;
    CU(cudaEventRecord(s.e0, s.st));
    { int lrc = launch_transient(s, a); if (lrc) return lrc; }
    CU(cudaEventRecord(s.e1, s.st));
    CU(cudaGetLastError());
    s.last_kind = 1; s.last_t = a;
    if (stats) CU(d2h_rows(stats, s.stats.as<double>(), 6, ncol, s.c0, s.nc, s.st));
    if (flags) CU(cudaMemcpyAsync(flags + s.c0, s.flags.p, sizeof(int) * s.nc, cudaMemcpyDeviceToHost, s.st));
    const double *hser = nullptr, *hfull = nullptr;
    if (ns > 0 && series) {
      const size_t b = sizeof(double) * (size_t)nt * MC_O_N * ns;
      CU(s.hser.need(b)); CU(cudaMemcpyAsync(s.hser.p, s.series.p, b, cudaMemcpyDeviceToHost, s.st)); hser = s.hser.as<double>();
    }
    if (ns > 0 && fullseries) {
      const size_t b = sizeof(double) * (size_t)nt * NSt * ns;
      CU(s.hfull.need(b)); CU(cudaMemcpyAsync(s.hfull.p, s.full.p, b, cudaMemcpyDeviceToHost, s.st)); hfull = s.hfull.as<double>();
    }
    CU(cudaStreamSynchronize(s.st));
    CU(cudaEventElapsedTime(&s.ms, s.e0, s.e1));
    s.has_state = true;
    // scatter this shard's sample slots into the caller's arrays (disjoint slots per shard)
    // (row by row: both sides are walked contiguously — column by column this was 1.4 M cache-missing element copies, ~5 ms per call for
    // 1 024 sampled columns x 168 h; with one shard owning every slot in order the rows are identical and copied whole)
    bool whole = ns == (long long)nsamp;
    for (long long k = 0; whole && k < ns; ++k) whole = slots[k] == k;
    auto scatter = [&](double* dst, const double* src, long long rows) {
      if (whole) { memcpy(dst, src, sizeof(double) * (size_t)rows * (size_t)ns); return; }
      for (long long r = 0; r < rows; ++r) {
        double* d = dst + r * nsamp; const double* h = src + r * ns;
        for (long long k = 0; k < ns; ++k) d[slots[k]] = h[k];
      }
    };
    if (ns > 0 && series) scatter(series, hser, nt * MC_O_N);
    if (ns > 0 && fullseries) scatter(fullseries, hfull, nt * NSt);
    return 0;
  });
  collect_timing(ctx);
  if (rc == 0 && ctx->stats_accum) ctx->stats_n += nt;
  return rc;
}

int mc_run_transient(mc_ctx* ctx, int64_t t0, int64_t t1, const double* tsoil, const int64_t* samp_idx, int64_t nsamp,
                     double* series, double* fullseries, double* stats, int32_t* flags) {
  if (!ctx) return MC_ERR_ARG;
  int spin = (t0 == 0) ? (int)ctx->cfg[MC_C_SPINSTEPS] : 0;
  return run_transient_impl(ctx, t0, t1, spin, 0, tsoil, samp_idx, nsamp, series, fullseries, stats, flags);
}
int mc_spinup(mc_ctx* ctx, int steps) {
  if (!ctx || steps <= 0) return fail(ctx, MC_ERR_ARG, "mc_spinup: steps must be positive");
  return run_transient_impl(ctx, 0, 0, steps, 1, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
}

int mc_run_steady(mc_ctx* ctx, int64_t T, const double* clim, const double* nmr, int nmr_percol, const double* scfg,
                  const double* pai_season, double* out, double* stats) {
  if (!ctx || ctx->ncol <= 0 || T <= 0 || !clim || !nmr || !scfg) return fail(ctx, MC_ERR_ARG, "mc_run_steady: bad argument (set columns first)");
  const long long ncol = ctx->ncol;
  const int m = ctx->m, sm = ctx->sm;
  // snow-layer lookup guard (.snowtempf, NMR.R:495-508): the below-snow branch needs 0.025 <= reqhgt < 3
  const double reqhgt = scfg[MC_SG_REQHGT];
  int rc = for_shards(ctx, [&](Shard& s) -> int {
    s.ms = 0; s.launches = 0;
    if (s.nc == 0) return 0;
    CU(s.sclim.need(sizeof(double) * T * MC_SC_N));
    CU(cudaMemcpyAsync(s.sclim.p, clim, sizeof(double) * T * MC_SC_N, cudaMemcpyHostToDevice, s.st));
    if (nmr_percol) {
      CU(s.snmr.need(sizeof(double) * T * MC_SN_N * s.nc));
      CU(h2d_rows(s.snmr.as<double>(), nmr, T * MC_SN_N, ncol, s.c0, s.nc, s.st));
    } else {
      CU(s.snmr.need(sizeof(double) * T * MC_SN_N));
      CU(cudaMemcpyAsync(s.snmr.p, nmr, sizeof(double) * T * MC_SN_N, cudaMemcpyHostToDevice, s.st));
    }
    if (pai_season) {
      CU(s.sseason.need(sizeof(double) * T));
      CU(cudaMemcpyAsync(s.sseason.p, pai_season, sizeof(double) * T, cudaMemcpyHostToDevice, s.st));
    }
    // split time so that the grid has enough threads to fill the GPU when there are few columns
    int nchunk = 1;
    const long long want_threads = 148LL * 2048;
    if (s.nc < want_threads) nchunk = (int)std::min<long long>((want_threads + s.nc - 1) / s.nc, std::max<long long>(1, T / 64));
    if (nchunk > 65535) nchunk = 65535;
    if (nchunk < 1) nchunk = 1;
    SteadyArgs a{};
    a.ncol = s.nc; a.m = m; a.sm = sm; a.T = T; a.vegp = s.vegp.as<double>(); a.soilp = s.soilp.as<double>();
    a.lat = s.lat.as<double>(); a.lon = s.lon.as<double>(); a.clim = s.sclim.as<double>(); a.nmr = s.snmr.as<double>();
    a.nmr_percol = nmr_percol; a.season = pai_season ? s.sseason.as<double>() : nullptr;
    if (s.nvt) {
      if (s.vegt_T != T) { s.err = "mc_run_steady: mc_set_vegt rows differ from T"; return MC_ERR_ARG; }
      a.vegt = s.vegt.as<double>(); a.nvt = s.nvt;
    }
    a.reqhgt = reqhgt; a.metopen = scfg[MC_SG_METOPEN] != 0; a.windhgt = scfg[MC_SG_WINDHGT]; a.surfwet = scfg[MC_SG_SURFWET];
    if (out) { CU(s.sout.need(sizeof(double) * T * MC_SO_N * s.nc)); a.out = s.sout.as<double>(); }
    if (stats) {
      CU(s.stats.need(sizeof(double) * 6 * s.nc)); a.stats = s.stats.as<double>();
      CU(s.spart.need(sizeof(double) * 6 * s.nc * nchunk)); a.part = s.spart.as<double>();
    }
    CU(s.flags.need(sizeof(int) * s.nc)); a.flags = s.flags.as<int>();
    CU(cudaMemsetAsync(s.flags.p, 0, sizeof(int) * s.nc, s.st));
    CU(s.ssol.need(sizeof(double) * T * 13)); a.sol = s.ssol.as<double>();
    CU(cudaEventRecord(s.e0, s.st));
    { int lrc = launch_steady(s, a, nchunk); if (lrc) return lrc; }
    CU(cudaEventRecord(s.e1, s.st));
    CU(cudaGetLastError());
    s.last_kind = 2; s.last_s = a; s.last_nchunk = nchunk;
    if (out) CU(d2h_rows(out, s.sout.as<double>(), T * MC_SO_N, ncol, s.c0, s.nc, s.st));
    if (stats) CU(d2h_rows(stats, s.stats.as<double>(), 6, ncol, s.c0, s.nc, s.st));
    std::vector<int> hflags(s.nc);
    CU(cudaMemcpyAsync(hflags.data(), s.flags.p, sizeof(int) * s.nc, cudaMemcpyDeviceToHost, s.st));
    CU(cudaStreamSynchronize(s.st));
    CU(cudaEventElapsedTime(&s.ms, s.e0, s.e1));
    for (int f : hflags) if (f & 4) { s.err = "snow-layer lookup out of range (.snowtempf needs 0.025 <= reqhgt < 3 m under snow)"; return MC_ERR_SNOWLAYER; }
    return 0;
  });
  collect_timing(ctx);
  return rc;
}

double mc_fp64_peak_tflops(mc_ctx* ctx) {
  if (!ctx || ctx->shards.empty()) return -1.0;
  Shard& s = ctx->shards[0];
  cudaSetDevice(s.dev);
  const int blocks = 148 * 16, threads = 256, iters = 20000;
  double* buf = nullptr;
  if (cudaMalloc(&buf, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
  float best = 1e30f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(s.e0, s.st);
    k_dfma_peak<<<blocks, threads, 0, s.st>>>(buf, iters, 0.999999, 1e-9);
    cudaEventRecord(s.e1, s.st);
    cudaStreamSynchronize(s.st);
    float ms = 0; cudaEventElapsedTime(&ms, s.e0, s.e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaFree(buf);
  if (cudaGetLastError() != cudaSuccess) return -1.0;
  double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
  return flops / (best * 1e-3) / 1e12;
}

int mc_eval_primitive(mc_ctx* ctx, int op, int64_t n, const double* x, const double* y, double* out) {
  if (!ctx || ctx->shards.empty() || n < 0 || !x || !out || op < 0 || op > 4 || (op == 4 && !y))
    return fail(ctx, MC_ERR_ARG, "mc_eval_primitive: bad argument");
  if (n == 0) return 0;
  Shard& s = ctx->shards[0];
  cudaSetDevice(s.dev);
  double *dx = nullptr, *dy = nullptr, *dout = nullptr;
  const size_t bytes = sizeof(double) * (size_t)n;
  cudaError_t e = cudaMalloc(&dx, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&dout, bytes);
  if (e == cudaSuccess && y) e = cudaMalloc(&dy, bytes);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dx, x, bytes, cudaMemcpyHostToDevice, s.st);
  if (e == cudaSuccess && y) e = cudaMemcpyAsync(dy, y, bytes, cudaMemcpyHostToDevice, s.st);
  if (e == cudaSuccess) {
    k_eval_primitive<<<(unsigned)((n + 127) / 128), 128, 0, s.st>>>(op, n, dx, dy, dout);
    e = cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, s.st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s.st);
  cudaFree(dx); cudaFree(dy); cudaFree(dout);
  if (e != cudaSuccess) return fail(ctx, MC_ERR_CUDA, std::string("mc_eval_primitive: ") + cudaGetErrorString(e));
  return 0;
}

void* mc_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void mc_host_free(void* p) { if (p) cudaFreeHost(p); }

double mc_last_kernel_ms(const mc_ctx* ctx) { return ctx ? ctx->last_ms : -1.0; }
int64_t mc_last_kernel_launches(const mc_ctx* ctx) { return ctx ? ctx->last_launches : 0; }

double mc_bench_resident(mc_ctx* ctx, int reps) {
  if (!ctx || reps <= 0) return -1.0;
  const int NSt = mc_nstate(ctx->m, ctx->sm);
  if (!ctx->keep_replay_state) {
    // The state copy a replay needs is not kept by default (it doubles the resident state).  The first call arms it: a transient run
    // without spin-up issued after this call can be replayed; one issued before cannot (its incoming state is gone).
    ctx->keep_replay_state = true;
    for (auto& s : ctx->shards)
      if (s.nc && s.last_kind == 1 && s.last_t.spin_steps <= 0 && !s.bak_valid) {
        fail(ctx, MC_ERR_STATE, "mc_bench_resident: replay armed; repeat the mc_run_transient call, then call mc_bench_resident again");
        return -1.0;
      }
  }
  int rc = for_shards(ctx, [&](Shard& s) -> int {
    s.ms = 0; s.launches = 0;
    if (s.nc == 0 || s.last_kind == 0) return 0;
    float total = 0;
    for (int r = 0; r < reps; ++r) {
      if (s.last_kind == 1 && s.last_t.spin_steps <= 0) {
        CU(cudaMemcpyAsync(s.state.p, s.state_bak.p, sizeof(double) * NSt * s.nc, cudaMemcpyDeviceToDevice, s.st));
        s.z_valid = s.bak_z_valid; s.z_reqhgt = s.bak_z_reqhgt; s.z_charmode = s.bak_z_charmode; s.z_epoch = s.bak_z_epoch;   // (the restored state's heights)
      }
      CU(cudaEventRecord(s.e0, s.st));
      { int lrc = (s.last_kind == 1) ? launch_transient(s, s.last_t) : launch_steady(s, s.last_s, s.last_nchunk); if (lrc) return lrc; }
      CU(cudaEventRecord(s.e1, s.st));
      CU(cudaGetLastError());
      CU(cudaStreamSynchronize(s.st));
      float ms = 0; CU(cudaEventElapsedTime(&ms, s.e0, s.e1));
      total += ms;
    }
    s.ms = total / reps;
    return 0;
  });
  collect_timing(ctx);
  if (rc) return -1.0;
  return ctx->last_ms;
}

}  // extern "C"
