// mpc_types.cuh -- problem constants, per-problem solver state, status codes and math wrappers shared by the solver
// core (warp_core.cuh), the kernels (mpcb200.cu) and the host-side test build (tests/host_sim).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MPC_HD __host__ __device__ __forceinline__
#else
#define MPC_HD inline
#endif

namespace mpcb200 {

// ------------------------------------------------------------------ status codes (mirror FORCESNLPsolver.h:70-106)
enum : int {
  ST_OPTIMAL = 1,        // converged
  ST_MAXIT = 0,          // iteration limit
  ST_STALLED = 3,        // feasible, at mu_min, step stalled at the rounding-noise floor above tol_step (usable, less accurate)
  ST_NAN = -6,           // NaN/Inf met
  ST_NOPROGRESS = -7,    // line search failed repeatedly
  ST_INFEASIBLE_X0 = -8, // pinned stage violates a constraint (friction row infeasible / nonconvex, x0 inside obstacle)
};

enum : int { HESS_GN = 0, HESS_EXACT = 1 };

template <typename T>
struct ParamsT {
  int N;
  int max_iter;
  int hessian;       // HESS_GN | HESS_EXACT (exact Lagrangian Hessian with adjoint multipliers, GN fallback)
  int ls_max;        // max backtracking trials
  T dt, l_wb, l_fric;
  T Q[5], R[2];
  T dd_min, dd_max, a_max, de_min, de_max, v_min, v_max;
  T r_sum, ego_off;
  T mu0, mu_min, mu_factor, tol_step, tol_feas, tau_min, bound_push;
  T acc_factor; int acc_iters;   // acceptable-level exit (warp core)
  T screen_inv_curv;             // obstacle rows with mu/s^2 < 1/screen_inv_curv are screened out for the iteration (0: never)
  T trust_step;                  // Newton-trust acceptance (no merit test) for feasible iterates and steps below this
  int stall_iters;               // stall exit after this many iterations at mu_min without halving the step
  T mu_min_alpha;                // the barrier parameter is reduced only after a step of at least this length
  T mu_up_alpha, mu_up_factor, mu_max;   // barrier warm-up: raise mu while the first steps are blocked below mu_up_alpha
  T mu_factor_full;                      // barrier reduction factor after a full (alpha = 1) primal and dual step
  T kappa_sigma;                 // multipliers are kept within [mu/(kappa s), kappa mu/s] after every step
  int init_rollout;              // 1: initial states = Euler rollout of the initial controls from the pinned state
  T stiff_slack;                 // float32: an acceptable-level exit with a live obstacle slack below this is flagged ST_STALLED
  T mu_warm, warm_push, kappa_warm;   // dual warm start (init_warm): restart barrier parameter, interior push, multiplier band
};

// per-problem scalars that persist across launches (one launch per SQP iteration mode)
template <typename T>
struct ProbState {
  T mu, rho;
  T a0_lo, a0_hi;     // stage-0 friction box (constants of the pinned stage)
  T kkt;              // last step inf-norm (diagnostic)
  T d_al, d_ap, d_ad, d_c1, d_dphi; int d_blk;   // diagnostics of the last iteration
  T best;             // smallest accepted step length*norm seen at mu_min (stall detection)
  T pstep;            // full Newton step norm of the previous iteration if it was taken in full near mu_min, else 0 (final-phase rate estimate)
  int status, iters, done, nfail, nsoc, nacc, centered, nstall;
};

// ------------------------------------------------------------------ math wrappers
MPC_HD float m_sqrt(float x) { return sqrtf(x); }
MPC_HD double m_sqrt(double x) { return sqrt(x); }
MPC_HD float m_abs(float x) { return fabsf(x); }
MPC_HD double m_abs(double x) { return fabs(x); }
MPC_HD float m_max(float a, float b) { return fmaxf(a, b); }
MPC_HD double m_max(double a, double b) { return fmax(a, b); }
MPC_HD float m_min(float a, float b) { return fminf(a, b); }
MPC_HD double m_min(double a, double b) { return fmin(a, b); }
MPC_HD float m_log1p(float x) { return log1pf(x); }
MPC_HD double m_log1p(double x) { return log1p(x); }
// float sine / cosine without the library's large-argument slow path (Payne-Hanek: ~100 instructions inlined at every call
// site, never executed for headings and steering angles): three-constant Cody-Waite reduction by pi/2 -- exact for
// |x| < 400 rad, degrading gracefully beyond -- and the usual minimax polynomials on [-pi/4, pi/4]; max error 1.5 ulp of 1.
// Pure fma arithmetic: the host emulator and the device compute the same thing.
MPC_HD void m_sincos(float x, float* s, float* c) {
  const float j = rintf(x * 0.636619772f);
  float r = fmaf(-j, 1.57079601e+00f, x);
  r = fmaf(-j, 3.13916473e-07f, r);
  r = fmaf(-j, 5.39030253e-15f, r);
  const int q = (int)j;
  const float r2 = r * r;
  const float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f), r2 * r, r);
  const float cp = fmaf(fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f), r2 * r2,
                        fmaf(-0.5f, r2, 1.0f));
  const float ss = (q & 1) ? cp : sp, cc = (q & 1) ? sp : cp;
  *s = (q & 2) ? -ss : ss;
  *c = ((q + 1) & 2) ? -cc : cc;
}
MPC_HD double m_tan(double x) { return tan(x); }
MPC_HD void m_sincos(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  sincos(x, s, c);
#else
  *s = sin(x); *c = cos(x);
#endif
}
template <typename T> MPC_HD bool m_finite(T x) { return (x - x) == T(0); }
// reciprocal / reciprocal square root used where a few ulp do not matter (barrier weights, step-length limits, merit
// ratios): on the device one MUFU op (+ one Newton step for rsqrt) instead of the IEEE division / sqrt sequences with
// their slow-path calls; float64 and the host build use the exact operations.
MPC_HD float m_rcp(float x) {
#if defined(__CUDA_ARCH__)
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
  return 1.0f / x;
#endif
}
MPC_HD double m_rcp(double x) { return 1.0 / x; }
MPC_HD float m_tan(float x) { float sn, cs; m_sincos(x, &sn, &cs); return sn * m_rcp(cs); }      // steering angle, |x| <= 1.066: ~2 ulp
MPC_HD float m_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
  float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * (1.5f - 0.5f * x * r * r);
#else
  return 1.0f / sqrtf(x);
#endif
}
MPC_HD double m_rsqrt(double x) { return 1.0 / sqrt(x); }
// square root where a few ulp do not matter (thresholds, the barrier schedule): x * rsqrt(x), no IEEE slow path
MPC_HD float m_sqrt_fast(float x) { const float y = fmaxf(x, 1e-30f); return y * m_rsqrt(y); }
MPC_HD double m_sqrt_fast(double x) { return sqrt(x); }
// natural log by one MUFU op (lg2.approx: absolute error ~1e-7 for arguments near 1); callers add the first-order
// correction for the rounding of the argument themselves (trial_merit)
MPC_HD float m_fastlog(float u) {
#if defined(__CUDA_ARCH__)
  float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(u)); return r * 0.69314718056f;
#else
  return logf(u);
#endif
}
MPC_HD double m_fastlog(double u) { return log(u); }
MPC_HD float m_eps(float) { return 6e-8f; }
MPC_HD double m_eps(double) { return 1.2e-16; }
// |r| with a dead zone at the rounding level of the distance h it was computed from: the slack residual of a far,
// inactive obstacle row (lane following: dummy obstacle ~130 m away, quirk Q11) is pure rounding noise of h.
template <typename T> MPC_HD T m_resid(T r, T h) { return m_max(m_abs(r) - T(8) * m_eps(T(0)) * h, T(0)); }
// slack of a bound row computed from the primal value; floored at a few ulps of the bound so that an iterate that
// rounds onto its bound (fp32: mu/nu can be below one ulp of x) gives a stiff but finite barrier weight.
MPC_HD float m_slack(float x) { return fmaxf(x, 2.5e-7f); }
MPC_HD double m_slack(double x) { return fmax(x, 1e-15); }

// dual slots per stage k (u_k rows then x_{k+1} rows)
enum : int { V_DD_LO = 0, V_DD_HI, V_A_HI, V_A_LO, V_DE_LO, V_DE_HI, V_V_LO, V_V_HI, V_OB0, V_OB1, V_OB2, NV = 11 };


}  // namespace mpcb200
