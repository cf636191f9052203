#pragma once
// stand-in for the Ceres API the reference's LiDAR path uses.  AutoDiffCostFunction evaluates the reference's own cost
// functors (lidarFactor.hpp) on dual numbers; Problem records the residual blocks; Solve() hands them to the oracle's
// restatement of the Ceres 1.14 trust-region / Levenberg-Marquardt / DENSE_QR loop (oracle/lm.c) through its block
// hook, so the residuals and Jacobians inside a solve come from the reference functors and only the minimiser itself
// is the oracle's.  Library stand-in, not reference source.
#include <cmath>
#include <cstdlib>
#include <vector>
#include "lmono_oracle.h"
namespace ceres {
template <class T, int N> struct Jet {
  T a; T v[N];
  Jet() : a(0) { for (auto& e : v) e = 0; }
  Jet(const T& s) : a(s) { for (auto& e : v) e = 0; }              // NOLINT: T(double) conversions are what the functors write
  Jet(int s) : a(s) { for (auto& e : v) e = 0; }                   // NOLINT
};
#define JET_BIN(op, A, DV) \
  template <class T, int N> inline Jet<T, N> operator op(const Jet<T, N>& f, const Jet<T, N>& g) { Jet<T, N> h; h.a = A; for (int i = 0; i < N; ++i) h.v[i] = DV; return h; }
JET_BIN(+, f.a + g.a, f.v[i] + g.v[i])
JET_BIN(-, f.a - g.a, f.v[i] - g.v[i])
JET_BIN(*, f.a * g.a, f.a * g.v[i] + f.v[i] * g.a)
#undef JET_BIN
template <class T, int N> inline Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {     // jet.h: h.v = (f.v - f.a / g.a * g.v) / g.a
  Jet<T, N> h; const T gi = T(1) / g.a; const T fg = f.a * gi; h.a = fg; for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - fg * g.v[i]) * gi; return h; }
template <class T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f) { Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <class T, int N> inline Jet<T, N> operator*(const Jet<T, N>& f, T s) { Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
template <class T, int N> inline Jet<T, N> operator*(T s, const Jet<T, N>& f) { return f * s; }
template <class T, int N> inline bool operator<(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a < g.a; }
template <class T, int N> inline bool operator>(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a > g.a; }
template <class T, int N> inline bool operator>=(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a >= g.a; }
template <class T, int N> inline bool operator<=(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a <= g.a; }
#define JET_FN(name, A, D) \
  template <class T, int N> inline Jet<T, N> name(const Jet<T, N>& f) { Jet<T, N> h; h.a = A; const T d = D; for (int i = 0; i < N; ++i) h.v[i] = d * f.v[i]; return h; }
JET_FN(sqrt, std::sqrt(f.a), T(1) / (T(2) * h.a))
JET_FN(sin, std::sin(f.a), std::cos(f.a))
JET_FN(cos, std::cos(f.a), -std::sin(f.a))
JET_FN(acos, std::acos(f.a), -T(1) / std::sqrt(T(1) - f.a * f.a))
JET_FN(abs, std::fabs(f.a), (f.a < T(0) ? T(-1) : T(1)))
#undef JET_FN

class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  int num_residuals() const { return nres_; }
 protected:
  int nres_ = 0;
};
template <class Functor, int kRes, int N0, int N1> class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(Functor* f) : f_(f) { nres_ = kRes; }
  ~AutoDiffCostFunction() override { delete f_; }
  bool Evaluate(double const* const* p, double* r, double** jac) const override {
    if (!jac) return (*f_)(p[0], p[1], r);
    typedef Jet<double, N0 + N1> J;
    J x0[N0], x1[N1], out[kRes];
    for (int i = 0; i < N0; ++i) { x0[i] = J(p[0][i]); x0[i].v[i] = 1.0; }
    for (int i = 0; i < N1; ++i) { x1[i] = J(p[1][i]); x1[i].v[N0 + i] = 1.0; }
    if (!(*f_)(x0, x1, out)) return false;
    for (int k = 0; k < kRes; ++k) {
      r[k] = out[k].a;
      if (jac[0]) for (int i = 0; i < N0; ++i) jac[0][k * N0 + i] = out[k].v[i];
      if (jac[1]) for (int i = 0; i < N1; ++i) jac[1][k * N1 + i] = out[k].v[N0 + i];
    }
    return true;
  }
 private:
  Functor* f_;
};
class LossFunction { public: virtual ~LossFunction() {} };
class HuberLoss : public LossFunction { public: explicit HuberLoss(double a) : a_(a) {} double a_; };
class LocalParameterization { public: virtual ~LocalParameterization() {} };
class EigenQuaternionParameterization : public LocalParameterization {};
enum LinearSolverType { DENSE_QR };
class Problem {
 public:
  struct Options {};
  Problem() {} explicit Problem(const Options&) {}
  ~Problem() { for (auto& b : blocks) delete b.cost; for (auto* l : losses) delete l; for (auto* l : locals) delete l; }
  void AddParameterBlock(double* p, int size, LocalParameterization* lp = nullptr) {
    if (size == 4) q = p; else if (size == 3) t = p; else std::abort();
    if (lp) { bool seen = false; for (auto* l : locals) seen |= l == lp; if (!seen) locals.push_back(lp); } }
  void AddResidualBlock(CostFunction* c, LossFunction* l, double* p0, double* p1) {
    if (p0 != q || p1 != t) std::abort();
    blocks.push_back({ c }); bool seen = false; for (auto* k : losses) seen |= k == l; if (l && !seen) losses.push_back(l); }
  struct Block { CostFunction* cost; };
  std::vector<Block> blocks; std::vector<LossFunction*> losses; std::vector<LocalParameterization*> locals;
  double* q = nullptr; double* t = nullptr;
};
class Solver { public:
  struct Options { LinearSolverType linear_solver_type = DENSE_QR; int max_num_iterations = 50; bool minimizer_progress_to_stdout = false;
                   bool check_gradients = false; double gradient_check_relative_precision = 1e-8; };
  struct Summary { o_solve_summary oracle; };
};
namespace refstub_detail {
inline int block_hook(void* user, int i, const double q[4], const double t[3], double r[3], double Jq[3][4], double Jt[3][3], int want_jac) {
  const Problem* p = (const Problem*)user;
  const CostFunction* c = p->blocks[(size_t)i].cost;
  const int n = c->num_residuals();
  double const* params[2] = { q, t };
  double jq[12], jt[9]; double* jac[2] = { jq, jt };
  if (!c->Evaluate(params, r, want_jac ? jac : nullptr)) std::abort();
  if (want_jac) for (int k = 0; k < n; ++k) { for (int j = 0; j < 4; ++j) Jq[k][j] = jq[k * 4 + j]; for (int j = 0; j < 3; ++j) Jt[k][j] = jt[k * 3 + j]; }
  return n;
}
}
inline void Solve(const Solver::Options& o, Problem* p, Solver::Summary* s) {
  for (auto* l : p->losses) if (((HuberLoss*)l)->a_ != 0.1) std::abort();       // the oracle's loss is HuberLoss(0.1), the only one the reference uses
  std::vector<o_factor> f(p->blocks.size());
  for (size_t i = 0; i < f.size(); ++i) { f[i] = o_factor(); f[i].type = p->blocks[i].cost->num_residuals() == 3 ? O_FACTOR_EDGE : O_FACTOR_PLANE_NORM; }
  o_pose x; for (int k = 0; k < 4; ++k) x.q[k] = p->q[k]; for (int k = 0; k < 3; ++k) x.t[k] = p->t[k];
  lmono_cpu_lm_set_block_hook(refstub_detail::block_hook, p);
  lmono_cpu_lm_solve(f.data(), (int)f.size(), &x, o.max_num_iterations, &s->oracle);
  lmono_cpu_lm_set_block_hook(nullptr, nullptr);
  for (int k = 0; k < 4; ++k) p->q[k] = x.q[k]; for (int k = 0; k < 3; ++k) p->t[k] = x.t[k];
}
}
