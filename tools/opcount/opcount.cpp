// TEST / MEASUREMENT INFRASTRUCTURE ONLY — never shipped, never loaded by microclimc_b200/.
//
// Operator census for the roofline denominator: both the literal CPU oracle (oracle/*.hpp) and the
// shipped fused column code (microclimc_b200/csrc/mc_*.cuh, host build) are compiled with `double` replaced by a
// counting scalar, so one run reports exactly how many FP64 additions, multiplications, divisions, square roots,
// exponentials, logarithms, powers, cube roots and trigonometric calls a column-timestep costs.
// tools/opcount/run_opcount.py drives it and writes profiles/cost_per_colstep.json.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <vector>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double real_t;
enum { K_ADD, K_MUL, K_DIV, K_SQRT, K_EXP, K_LOG, K_POW, K_CBRT, K_TRIG, K_CMP, K_N };
static thread_local unsigned long long g_cnt[K_N];
static thread_local bool g_on = true;
#define TICK(k) do { if (g_on) ++g_cnt[k]; } while (0)

struct Cnt {
  real_t v;
  constexpr Cnt() : v(0) {}
  constexpr Cnt(real_t x) : v(x) {}
  explicit operator int() const { return (int)v; }
  explicit operator long long() const { return (long long)v; }
  explicit operator long() const { return (long)v; }
  explicit operator bool() const { return v != 0; }
  Cnt& operator+=(Cnt o) { TICK(K_ADD); v += o.v; return *this; }
  Cnt& operator-=(Cnt o) { TICK(K_ADD); v -= o.v; return *this; }
  Cnt& operator*=(Cnt o) { TICK(K_MUL); v *= o.v; return *this; }
  Cnt& operator/=(Cnt o) { TICK(K_DIV); v /= o.v; return *this; }
};
inline Cnt operator+(Cnt a, Cnt b) { TICK(K_ADD); return Cnt(a.v + b.v); }
inline Cnt operator-(Cnt a, Cnt b) { TICK(K_ADD); return Cnt(a.v - b.v); }
inline Cnt operator*(Cnt a, Cnt b) { TICK(K_MUL); return Cnt(a.v * b.v); }
inline Cnt operator/(Cnt a, Cnt b) { TICK(K_DIV); return Cnt(a.v / b.v); }
inline Cnt operator-(Cnt a) { return Cnt(-a.v); }
inline Cnt operator+(Cnt a) { return a; }
inline bool operator<(Cnt a, Cnt b) { TICK(K_CMP); return a.v < b.v; }
inline bool operator>(Cnt a, Cnt b) { TICK(K_CMP); return a.v > b.v; }
inline bool operator<=(Cnt a, Cnt b) { TICK(K_CMP); return a.v <= b.v; }
inline bool operator>=(Cnt a, Cnt b) { TICK(K_CMP); return a.v >= b.v; }
inline bool operator==(Cnt a, Cnt b) { TICK(K_CMP); return a.v == b.v; }
inline bool operator!=(Cnt a, Cnt b) { TICK(K_CMP); return a.v != b.v; }
#define FN1(name, kind) inline Cnt name(Cnt x) { TICK(kind); return Cnt(::name(x.v)); }
FN1(sqrt, K_SQRT) FN1(exp, K_EXP) FN1(log, K_LOG) FN1(cbrt, K_CBRT) FN1(sin, K_TRIG) FN1(cos, K_TRIG) FN1(tan, K_TRIG) FN1(atan, K_TRIG)
inline Cnt pow(Cnt x, Cnt y) { TICK(K_POW); return Cnt(::pow(x.v, y.v)); }
inline Cnt fabs(Cnt x) { return Cnt(::fabs(x.v)); }
inline Cnt nearbyint(Cnt x) { return Cnt(::nearbyint(x.v)); }
inline bool isnan(Cnt x) { return std::isnan(x.v); }
inline bool isinf(Cnt x) { return std::isinf(x.v); }
inline bool isfinite(Cnt x) { return std::isfinite(x.v); }
namespace std {
using ::sqrt; using ::exp; using ::log; using ::cbrt; using ::sin; using ::cos; using ::tan; using ::atan; using ::pow; using ::fabs;
using ::nearbyint; using ::isnan; using ::isinf; using ::isfinite;
template <> struct numeric_limits<Cnt> {
  static constexpr bool is_specialized = true;
  static constexpr Cnt quiet_NaN() { return Cnt(numeric_limits<real_t>::quiet_NaN()); }
};
}  // namespace std

#define double Cnt
#include "../../oracle/mc_oracle.cpp"
#include "../../tests/hostdbg/dbg_host.cpp"
#undef double

extern "C" void cnt_reset() { std::memset(g_cnt, 0, sizeof(g_cnt)); }
extern "C" void cnt_get(unsigned long long* out) { for (int k = 0; k < K_N; ++k) out[k] = g_cnt[k]; }
