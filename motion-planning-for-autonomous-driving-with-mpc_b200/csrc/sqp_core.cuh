// sqp_core.cuh -- per-problem nonlinear-MPC solver core (one ego instance = one CUDA lane).
//
// Solves the NLP that the reference hands to IPOPT each MPC step
// (/root/reference/MPC_Planner/optimizer.py:513-560, constraints :373-411, bounds :413-491, cost :493-511)
// with a Gauss-Newton / Newton SQP-type primal-dual interior-point iteration whose linear system -- the
// block-tridiagonal KKT matrix in stage order (u_0, x_1, u_1, x_2, ...) -- is factored and solved by a
// Riccati (block LDL^T) sweep that exploits the 6-non-zero structure of A_k = I + dt*df/dx and the constant
// B = dt*[e_delta e_v].  New code: the reference contains no solver of its own (it calls casadi/IPOPT).
//
// The pinned stage (X_0 = X_ref[:,0], optimizer.py:378) is eliminated: unknowns are u_k (k=0..N-1), x_k (k=1..N).
// The stage-0 friction row |a_0^2 + v_0^2 tan(delta_0)/2.578| <= a_max (optimizer.py:378, 424-425) is, with X_0
// pinned, exactly a box on a_0 and is handled as such.
//
// Everything here is `__host__ __device__` scalar code parameterised on the arithmetic type T (float | double)
// and on a strided workspace accessor, so the SAME source is what the CUDA kernels run (one lane per problem,
// workspace column `ws[idx*32 + lane]` in shared memory) and what tests/host_sim compiles with g++ to debug the
// algorithm in a container without a GPU.  The host build is test tooling only; the product never runs it.
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
};

// per-problem scalars that persist across launches (one launch per SQP iteration mode)
template <typename T>
struct ProbState {
  T mu, rho;
  T a0_lo, a0_hi;     // stage-0 friction box (constants of the pinned stage)
  T kkt;              // last step inf-norm (diagnostic)
  T d_al, d_ap, d_ad, d_c1, d_dphi; int d_blk;   // diagnostics of the last iteration
  int status, iters, done, nfail, nsoc;
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
MPC_HD float m_tan(float x) { return tanf(x); }
MPC_HD double m_tan(double x) { return tan(x); }
MPC_HD void m_sincos(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
  sincosf(x, s, c);
#else
  *s = sinf(x); *c = cosf(x);
#endif
}
MPC_HD void m_sincos(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  sincos(x, s, c);
#else
  *s = sin(x); *c = cos(x);
#endif
}
template <typename T> MPC_HD bool m_finite(T x) { return (x - x) == T(0); }
MPC_HD float m_eps(float) { return 6e-8f; }
MPC_HD double m_eps(double) { return 1.2e-16; }
// |r| with a dead zone at the rounding level of the distance h it was computed from: the slack residual of a far,
// inactive obstacle row (lane following: dummy obstacle ~130 m away, quirk Q11) is pure rounding noise of h.
template <typename T> MPC_HD T m_resid(T r, T h) { return m_max(m_abs(r) - T(8) * m_eps(T(0)) * h, T(0)); }
// slack of a bound row computed from the primal value; floored at a few ulps of the bound so that an iterate that
// rounds onto its bound (fp32: mu/nu can be below one ulp of x) gives a stiff but finite barrier weight.
MPC_HD float m_slack(float x) { return fmaxf(x, 2.5e-7f); }
MPC_HD double m_slack(double x) { return fmax(x, 1e-15); }

// ------------------------------------------------------------------ workspace layout (words of T per problem)
// DEVIATION COORDINATES.  Stage j's state is stored as xt_j = x_j - rho_j, where rho_j = X_ref[:, min(j+1, N)] is the
// reference row the cost pairs stage j with (optimizer.py:509, quirk Q2).  The loader forms xt_j and the position
// increments c_j = rho_j - rho_{j+1} in float64 before rounding to T, so in fp32 the dynamics defects
//     d_k = xt_k - xt_{k+1} + dt*f(x_k,u_k) + c_k
// are differences of O(1 m) numbers (ulp 1e-7) instead of O(100 m) absolute coordinates (ulp 8e-6).
// [rho 5(N+1)] [XT 5(N+1)] [U 2N] [KK 12N] [V 11N] [S 3N] [TR 3(N+1)] [CP 2N]
struct Layout {
  int N, o_xref, o_X, o_U, o_K, o_V, o_S, o_T, o_C, words;
  MPC_HD explicit Layout(int N_) : N(N_) {
    o_xref = 0;
    o_X = o_xref + 5 * (N + 1);
    o_U = o_X + 5 * (N + 1);
    o_K = o_U + 2 * N;
    o_V = o_K + 12 * N;
    o_S = o_V + 11 * N;
    o_T = o_S + 3 * N;
    o_C = o_T + 3 * (N + 1);
    words = o_C + 2 * N;
  }
};

// dual slots per stage k (u_k rows then x_{k+1} rows)
enum : int { V_DD_LO = 0, V_DD_HI, V_A_HI, V_A_LO, V_DE_LO, V_DE_HI, V_V_LO, V_V_HI, V_OB0, V_OB1, V_OB2, NV = 11 };

template <typename T, int STRIDE>
struct Ws {
  T* p;
  MPC_HD T& operator()(int i) const { return p[(size_t)i * STRIDE]; }
};

#define SI(i, j) ((i) <= (j) ? ((i) * 5 - (i) * ((i) - 1) / 2 + ((j) - (i))) : ((j) * 5 - (j) * ((j) - 1) / 2 + ((i) - (j))))

template <typename T, int STRIDE>
struct Solver {
  typedef Ws<T, STRIDE> W;
  const ParamsT<T>& P;
  const Layout L;
  W ws;
  const T* obs;   // 6 values: obstacle circle centres (centre, front, rear) in the problem's SHIFTED frame
  MPC_HD Solver(const ParamsT<T>& P_, W ws_, const T* obs_) : P(P_), L(P_.N), ws(ws_), obs(obs_) {}

  // accessors
  MPC_HD T& xr(int k, int j) const { return ws(L.o_xref + 5 * k + j); }   // rho_k (frame shifted to the pinned position)
  MPC_HD T& X(int k, int j) const { return ws(L.o_X + 5 * k + j); }       // xt_k = x_k - rho_k
  MPC_HD T& CP(int k, int j) const { return ws(L.o_C + 2 * k + j); }      // (rho_k - rho_{k+1}) positions, from float64
  MPC_HD T xa(int k, int j) const { return X(k, j) + xr(k, j); }          // absolute (shifted-frame) state
  // reference increment of component j between stages k and k+1
  MPC_HD T cinc(int k, int j) const { return (j < 2) ? CP(k, j) : (xr(k, j) - xr(k + 1, j)); }
  MPC_HD T& U(int k, int j) const { return ws(L.o_U + 2 * k + j); }
  MPC_HD T& KK(int k, int j) const { return ws(L.o_K + 12 * k + j); }   // 0..9 K (row-major 2x5), 10..11 kff
  MPC_HD T& V(int k, int j) const { return ws(L.o_V + NV * k + j); }
  MPC_HD T& S(int k, int j) const { return ws(L.o_S + 3 * k + j); }
  MPC_HD T& TR(int k, int j) const { return ws(L.o_T + 3 * k + j); }     // sin psi_k, cos psi_k, tan delta_k
  // after the forward sweep the K block of stage k holds the step: 0..4 dx_{k+1}, 5..6 du_k
  MPC_HD T& DX(int k, int j) const { return ws(L.o_K + 12 * k + j); }
  MPC_HD T& DU(int k, int j) const { return ws(L.o_K + 12 * k + 5 + j); }

  // ---------------------------------------------------------------- obstacle row j at state (sx,sy,psi): value + gradient
  MPC_HD void obst(int j, T sx, T sy, T sn, T cs, T& h, T& gx, T& gy, T& gp) const {
    const T sg = (j == 0) ? T(0) : (j == 1 ? T(1) : T(-1));
    const T o = sg * P.ego_off;
    const T dx = sx + o * cs - obs[2 * j], dy = sy + o * sn - obs[2 * j + 1];
    h = m_sqrt(dx * dx + dy * dy);
    const T ih = T(1) / m_max(h, T(1e-12));
    gx = dx * ih; gy = dy * ih;
    gp = o * (gy * cs - gx * sn);
  }

  // ---------------------------------------------------------------- problem I/O (float64 row-major arrays of ONE problem)
  // xref [N+1][5] (row 0 = pinned state), Xin [N+1][5], Uin [N][2]; obstacle_abs[6] absolute circle centres.
  // Differences are formed in float64, then rounded to T (see Layout comment).  obs_out receives the obstacle circle
  // centres in the frame shifted to the pinned position.
  MPC_HD void load(const double* xref, const double* Xin, const double* Uin, const double* obstacle_abs, T* obs_out) const {
    const int N = P.N;
    const double ox = xref[0], oy = xref[1];
    for (int j = 0; j < 3; ++j) { obs_out[2 * j] = (T)(obstacle_abs[2 * j] - ox); obs_out[2 * j + 1] = (T)(obstacle_abs[2 * j + 1] - oy); }
    for (int k = 0; k <= N; ++k) {
      const double* rho = xref + 5 * ((k + 1 < N) ? (k + 1) : N);
      const double* xin = (k == 0) ? xref : (Xin + 5 * k);           // stage 0 is pinned to X_ref[:,0]
      xr(k, 0) = (T)(rho[0] - ox); xr(k, 1) = (T)(rho[1] - oy);
      xr(k, 2) = (T)rho[2]; xr(k, 3) = (T)rho[3]; xr(k, 4) = (T)rho[4];
      for (int j = 0; j < 5; ++j) X(k, j) = (T)(xin[j] - rho[j]);
      if (k < N) {
        const double* rho1 = xref + 5 * ((k + 2 < N) ? (k + 2) : N);
        CP(k, 0) = (T)(rho[0] - rho1[0]); CP(k, 1) = (T)(rho[1] - rho1[1]);
        U(k, 0) = (T)Uin[2 * k]; U(k, 1) = (T)Uin[2 * k + 1];
      }
    }
  }
  MPC_HD void store(const double* xref, double* Xout, double* Uout) const {
    const int N = P.N;
    for (int k = 0; k <= N; ++k) {
      const double* rho = xref + 5 * ((k + 1 < N) ? (k + 1) : N);
      for (int j = 0; j < 5; ++j) Xout[5 * k + j] = (k == 0) ? xref[j] : ((double)X(k, j) + rho[j]);
      if (k < N) { Uout[2 * k] = (double)U(k, 0); Uout[2 * k + 1] = (double)U(k, 1); }
    }
  }

  // ---------------------------------------------------------------- stage helpers
  struct Trig { T sn, cs, tn; };

  // defect of stage k: d = xt_k - xt_{k+1} + dt*f(x_k,u_k) + c_k   (x0d/x1d deviations, v abs speed of x_k)
  MPC_HD void defect(int k, const T* x0d, const T* x1d, T v, const Trig& t, T u0, T u1, T* d) const {
    const T dt = P.dt;
    d[0] = (x0d[0] - x1d[0]) + dt * v * t.cs + CP(k, 0);
    d[1] = (x0d[1] - x1d[1]) + dt * v * t.sn + CP(k, 1);
    d[2] = (x0d[2] - x1d[2]) + dt * u0 + (xr(k, 2) - xr(k + 1, 2));
    d[3] = (x0d[3] - x1d[3]) + dt * u1 + (xr(k, 3) - xr(k + 1, 3));
    d[4] = (x0d[4] - x1d[4]) + dt * v * t.tn / P.l_wb + (xr(k, 4) - xr(k + 1, 4));
  }

  // ---------------------------------------------------------------- initialisation
  // Requires rho, XT, U, CP already in the workspace.  Pushes the start point strictly inside the bounds
  // (IPOPT's bound_push idea) and initialises slacks/duals on the central path.
  MPC_HD void init(ProbState<T>& st) const {
    const int N = P.N;
    st.mu = P.mu0; st.rho = T(1); st.status = ST_MAXIT; st.iters = 0; st.done = 0; st.nfail = 0; st.nsoc = 0; st.kkt = T(0);
    const T de0 = xa(0, 2), v0 = xa(0, 3);
    const T s0 = v0 * v0 * m_tan(de0) / P.l_fric;
    bool bad = false;
    // friction row: |a0^2 + s0| <= a_max.  Feasible set is the box |a0| <= sqrt(a_max - s0) when |s0| < a_max.
    if (!(s0 < P.a_max) || !(s0 > -P.a_max)) bad = true;
    const T amax0 = m_sqrt(m_max(P.a_max - s0, T(1e-12)));
    st.a0_hi = m_min(amax0, P.a_max);
    st.a0_lo = -amax0;
    if (de0 < P.de_min || de0 > P.de_max || v0 < P.v_min || v0 > P.v_max) bad = true;
    {
      T sn, cs; m_sincos(xa(0, 4), &sn, &cs);
      for (int j = 0; j < 3; ++j) {
        T h, gx, gy, gp; obst(j, xa(0, 0), xa(0, 1), sn, cs, h, gx, gy, gp);
        if (h < P.r_sum) bad = true;
      }
    }
    if (bad) { st.status = ST_INFEASIBLE_X0; st.done = 1; }
    const T kp = P.bound_push;
    for (int k = 0; k < N; ++k) {
      const T pdd = m_min(kp, kp * (P.dd_max - P.dd_min));
      T dd = m_min(m_max(U(k, 0), P.dd_min + pdd), P.dd_max - pdd);
      T ahi = (k == 0) ? st.a0_hi : P.a_max;
      T a = U(k, 1);
      if (k == 0) {
        const T pa = m_min(kp * m_max(T(1), ahi), kp * (ahi - st.a0_lo));
        a = m_min(m_max(a, st.a0_lo + pa), ahi - pa);
      } else {
        a = m_min(a, ahi - kp * m_max(T(1), m_abs(ahi)));
      }
      U(k, 0) = dd; U(k, 1) = a;
      const T pde = m_min(kp * m_max(T(1), m_abs(P.de_max)), kp * (P.de_max - P.de_min));
      const T pv = m_min(kp * m_max(T(1), m_abs(P.v_max)), kp * (P.v_max - P.v_min));
      T de = m_min(m_max(xa(k + 1, 2), P.de_min + pde), P.de_max - pde);
      T vv = m_min(m_max(xa(k + 1, 3), P.v_min + pv), P.v_max - pv);
      X(k + 1, 2) = de - xr(k + 1, 2);
      X(k + 1, 3) = vv - xr(k + 1, 3);
      V(k, V_DD_LO) = st.mu / (dd - P.dd_min);
      V(k, V_DD_HI) = st.mu / (P.dd_max - dd);
      V(k, V_A_HI) = st.mu / (ahi - a);
      V(k, V_A_LO) = (k == 0) ? st.mu / (a - st.a0_lo) : T(0);
      V(k, V_DE_LO) = st.mu / (de - P.de_min);
      V(k, V_DE_HI) = st.mu / (P.de_max - de);
      V(k, V_V_LO) = st.mu / (vv - P.v_min);
      V(k, V_V_HI) = st.mu / (P.v_max - vv);
      T sn, cs; m_sincos(xa(k + 1, 4), &sn, &cs);
      for (int j = 0; j < 3; ++j) {
        T h, gx, gy, gp; obst(j, xa(k + 1, 0), xa(k + 1, 1), sn, cs, h, gx, gy, gp);
        const T c = h - P.r_sum;
        const T s = m_max(c, kp * m_max(T(1), P.r_sum));
        S(k, j) = s;
        V(k, V_OB0 + j) = st.mu / s;
      }
    }
  }

  // ---------------------------------------------------------------- backward (factor) sweep
  // Returns false if an exact-Hessian stage block was not positive definite (caller retries with GN).
  MPC_HD bool backward(const ProbState<T>& st, int hess) const {
    const int N = P.N;
    const T dt = P.dt, mu = st.mu;
    T Pm[15], p[5], lam[5];
#pragma unroll
    for (int i = 0; i < 15; ++i) Pm[i] = T(0);
#pragma unroll
    for (int i = 0; i < 5; ++i) { p[i] = T(0); lam[i] = T(0); }
    // x_N: deviation, absolute, trig
    T x1d[5], x1a[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) { x1d[j] = X(N, j); x1a[j] = x1d[j] + xr(N, j); }
    Trig t1;
    m_sincos(x1a[4], &t1.sn, &t1.cs);
    t1.tn = m_tan(x1a[2]);
    TR(N, 0) = t1.sn; TR(N, 1) = t1.cs; TR(N, 2) = t1.tn;

    for (int k = N - 1; k >= 0; --k) {
      // ---- terms of x_{k+1}: cost (stages 1..N-1 only, quirk Q1), box barriers, obstacle barriers
      if (k + 1 <= N - 1) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const T gq = T(2) * P.Q[j] * x1d[j];
          Pm[SI(j, j)] += T(2) * P.Q[j];
          p[j] += gq; lam[j] += gq;
        }
      }
      {
        const T slo = m_slack(x1a[2] - P.de_min), shi = m_slack(P.de_max - x1a[2]);
        const T vlo = V(k, V_DE_LO), vhi = V(k, V_DE_HI);
        Pm[SI(2, 2)] += vlo / slo + vhi / shi;
        p[2] += -mu / slo + mu / shi;
        lam[2] += -vlo + vhi;
      }
      {
        const T slo = m_slack(x1a[3] - P.v_min), shi = m_slack(P.v_max - x1a[3]);
        const T vlo = V(k, V_V_LO), vhi = V(k, V_V_HI);
        Pm[SI(3, 3)] += vlo / slo + vhi / shi;
        p[3] += -mu / slo + mu / shi;
        lam[3] += -vlo + vhi;
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        T h, gx, gy, gp; obst(j, x1a[0], x1a[1], t1.sn, t1.cs, h, gx, gy, gp);
        // slack reset s <- max(s, c(x)) (Nocedal & Wright 19.3): the distance is convex, so its linearisation
        // under-estimates it and the slack would otherwise creep behind the true clearance and jam the step length.
        T s = S(k, j);
        if (h - P.r_sum > s) { s = h - P.r_sum; S(k, j) = s; }
        const T nu = V(k, V_OB0 + j);
        const T w = nu / s, r = (h - P.r_sum) - s;
        const T cg = -(mu / s - w * r);
        Pm[SI(0, 0)] += w * gx * gx; Pm[SI(0, 1)] += w * gx * gy; Pm[SI(0, 4)] += w * gx * gp;
        Pm[SI(1, 1)] += w * gy * gy; Pm[SI(1, 4)] += w * gy * gp; Pm[SI(4, 4)] += w * gp * gp;
        p[0] += cg * gx; p[1] += cg * gy; p[4] += cg * gp;
        lam[0] -= nu * gx; lam[1] -= nu * gy; lam[4] -= nu * gp;
      }
      // ---- linearisation of stage k dynamics at (x_k, u_k)
      T x0d[5], x0a[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) { x0d[j] = X(k, j); x0a[j] = x0d[j] + xr(k, j); }
      const T u0 = U(k, 0), u1 = U(k, 1);
      Trig t0;
      m_sincos(x0a[4], &t0.sn, &t0.cs);
      t0.tn = m_tan(x0a[2]);
      TR(k, 0) = t0.sn; TR(k, 1) = t0.cs; TR(k, 2) = t0.tn;
      const T v = x0a[3];
      const T sec2 = T(1) + t0.tn * t0.tn;
      const T e03 = dt * t0.cs, e04 = -dt * v * t0.sn, e13 = dt * t0.sn, e14 = dt * v * t0.cs;
      const T e42 = dt * v * sec2 / P.l_wb, e43 = dt * t0.tn / P.l_wb;
      T d[5];
      defect(k, x0d, x1d, v, t0, u0, u1, d);
      // ---- control terms
      T Ru0, Ru1, ru0, ru1;
      {
        const T slo = m_slack(u0 - P.dd_min), shi = m_slack(P.dd_max - u0);
        Ru0 = T(2) * P.R[0] + V(k, V_DD_LO) / slo + V(k, V_DD_HI) / shi;
        ru0 = T(2) * P.R[0] * u0 - mu / slo + mu / shi;
        const T ahi = (k == 0) ? st.a0_hi : P.a_max;
        const T sh = m_slack(ahi - u1);
        Ru1 = T(2) * P.R[1] + V(k, V_A_HI) / sh;
        ru1 = T(2) * P.R[1] * u1 + mu / sh;
        if (k == 0) {
          const T sl = m_slack(u1 - st.a0_lo);
          Ru1 += V(k, V_A_LO) / sl;
          ru1 -= mu / sl;
        }
      }
      // ---- Riccati step
      T Pd[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        T a = p[i];
#pragma unroll
        for (int j = 0; j < 5; ++j) a += Pm[SI(i, j)] * d[j];
        Pd[i] = a;
      }
      const T dt2 = dt * dt;
      const T G00 = Ru0 + dt2 * Pm[SI(2, 2)], G01 = dt2 * Pm[SI(2, 3)], G11 = Ru1 + dt2 * Pm[SI(3, 3)];
      const T g0 = ru0 + dt * Pd[2], g1 = ru1 + dt * Pd[3];
      const T det = G00 * G11 - G01 * G01;
      if (hess == HESS_EXACT) {
        if (!(G00 > T(0)) || !(det > T(1e-8) * G00 * G11)) return false;
      }
      const T idet = T(1) / det;
      const T I00 = G11 * idet, I01 = -G01 * idet, I11 = G00 * idet;
      const T k0 = -(I00 * g0 + I01 * g1), k1 = -(I01 * g0 + I11 * g1);
      KK(k, 10) = k0; KK(k, 11) = k1;
      if (k > 0) {
        T M[5][5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          M[i][0] = Pm[SI(i, 0)];
          M[i][1] = Pm[SI(i, 1)];
          M[i][2] = Pm[SI(i, 2)] + Pm[SI(i, 4)] * e42;
          M[i][3] = Pm[SI(i, 3)] + Pm[SI(i, 0)] * e03 + Pm[SI(i, 1)] * e13 + Pm[SI(i, 4)] * e43;
          M[i][4] = Pm[SI(i, 4)] + Pm[SI(i, 0)] * e04 + Pm[SI(i, 1)] * e14;
        }
        T H0[5], H1[5], K0[5], K1[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          H0[j] = dt * M[2][j]; H1[j] = dt * M[3][j];
          K0[j] = -(I00 * H0[j] + I01 * H1[j]);
          K1[j] = -(I01 * H0[j] + I11 * H1[j]);
          KK(k, j) = K0[j]; KK(k, 5 + j) = K1[j];
        }
        T AtPd[5], Atl[5];
        AtPd[0] = Pd[0]; AtPd[1] = Pd[1];
        AtPd[2] = Pd[2] + e42 * Pd[4];
        AtPd[3] = Pd[3] + e03 * Pd[0] + e13 * Pd[1] + e43 * Pd[4];
        AtPd[4] = Pd[4] + e04 * Pd[0] + e14 * Pd[1];
        Atl[0] = lam[0]; Atl[1] = lam[1];
        Atl[2] = lam[2] + e42 * lam[4];
        Atl[3] = lam[3] + e03 * lam[0] + e13 * lam[1] + e43 * lam[4];
        Atl[4] = lam[4] + e04 * lam[0] + e14 * lam[1];
        T Pn[15];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
#pragma unroll
          for (int j = i; j < 5; ++j) {
            T a = M[i][j];
            if (i == 2) a += e42 * M[4][j];
            if (i == 3) a += e03 * M[0][j] + e13 * M[1][j] + e43 * M[4][j];
            if (i == 4) a += e04 * M[0][j] + e14 * M[1][j];
            Pn[SI(i, j)] = a + H0[i] * K0[j] + H1[i] * K1[j];
          }
        }
        if (hess == HESS_EXACT) {
          // + dt * sum_i lam_{k+1,i} * hess f_i(x_k) on the (delta, v, psi) block (adjoint multiplier estimate)
          Pn[SI(4, 4)] += dt * (-lam[0] * v * t0.cs - lam[1] * v * t0.sn);
          Pn[SI(3, 4)] += dt * (-lam[0] * t0.sn + lam[1] * t0.cs);
          Pn[SI(2, 2)] += dt * lam[4] * T(2) * v / P.l_wb * sec2 * t0.tn;
          Pn[SI(2, 3)] += dt * lam[4] * sec2 / P.l_wb;
        }
#pragma unroll
        for (int i = 0; i < 15; ++i) Pm[i] = Pn[i];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          p[i] = AtPd[i] + H0[i] * k0 + H1[i] * k1;
          lam[i] = Atl[i];
        }
      }
#pragma unroll
      for (int j = 0; j < 5; ++j) { x1d[j] = x0d[j]; x1a[j] = x0a[j]; }
      t1 = t0;
    }
    return true;
  }

  // ---------------------------------------------------------------- forward sweep: step, step-length limits, merit slope
  struct FwdOut { T a_p, a_d, dphi, c1, step_inf, mag; int blk, cur; };

  MPC_HD void row_limits(T s, T nu, T ds, T mu, T tau, FwdOut& o) const {
    const T dnu = (mu - nu * s - nu * ds) / s;
    if (ds < T(0) && -tau * s / ds < o.a_p) { o.a_p = -tau * s / ds; o.blk = o.cur; }
    o.cur++;
    if (dnu < T(0)) o.a_d = m_min(o.a_d, -tau * nu / dnu);
    o.dphi += -mu * ds / s;
  }

  MPC_HD FwdOut forward(const ProbState<T>& st) const {
    const int N = P.N;
    const T dt = P.dt, mu = st.mu;
    const T tau = m_max(P.tau_min, T(1) - mu);
    FwdOut o; o.a_p = T(1); o.a_d = T(1); o.dphi = T(0); o.c1 = T(0); o.step_inf = T(0); o.mag = T(0); o.blk = -1; o.cur = 0;
    T dx[5] = {T(0), T(0), T(0), T(0), T(0)};
    T x0d[5], x0a[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) { x0d[j] = X(0, j); x0a[j] = x0d[j] + xr(0, j); }
    for (int k = 0; k < N; ++k) {
      Trig t0; t0.sn = TR(k, 0); t0.cs = TR(k, 1); t0.tn = TR(k, 2);
      const T v = x0a[3];
      const T sec2 = T(1) + t0.tn * t0.tn;
      const T u0 = U(k, 0), u1 = U(k, 1);
      T du0 = KK(k, 10), du1 = KK(k, 11);
      if (k > 0) {
#pragma unroll
        for (int j = 0; j < 5; ++j) { du0 += KK(k, j) * dx[j]; du1 += KK(k, 5 + j) * dx[j]; }
      }
      T x1d[5], x1a[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) { x1d[j] = X(k + 1, j); x1a[j] = x1d[j] + xr(k + 1, j); }
      T d[5];
      defect(k, x0d, x1d, v, t0, u0, u1, d);
      T nx[5];
      nx[0] = dx[0] + dt * t0.cs * dx[3] - dt * v * t0.sn * dx[4] + d[0];
      nx[1] = dx[1] + dt * t0.sn * dx[3] + dt * v * t0.cs * dx[4] + d[1];
      nx[2] = dx[2] + dt * du0 + d[2];
      nx[3] = dx[3] + dt * du1 + d[3];
      nx[4] = dx[4] + dt * v * sec2 / P.l_wb * dx[2] + dt * t0.tn / P.l_wb * dx[3] + d[4];
#pragma unroll
      for (int j = 0; j < 5; ++j) { o.c1 += m_abs(d[j]); o.mag += m_abs(x0d[j]) + m_abs(x1d[j]); }
      o.mag += dt * (T(2) * m_abs(v) + m_abs(u0) + m_abs(u1)) + m_abs(CP(k, 0)) + m_abs(CP(k, 1));
      // cost slope
      o.dphi += T(2) * P.R[0] * u0 * du0 + T(2) * P.R[1] * u1 * du1;
      if (k + 1 <= N - 1) {
#pragma unroll
        for (int j = 0; j < 5; ++j) o.dphi += T(2) * P.Q[j] * x1d[j] * nx[j];
      }
      // inequality rows of u_k
      o.cur = 16 * k;
      row_limits(m_slack(u0 - P.dd_min), V(k, V_DD_LO), du0, mu, tau, o);
      row_limits(m_slack(P.dd_max - u0), V(k, V_DD_HI), -du0, mu, tau, o);
      const T ahi = (k == 0) ? st.a0_hi : P.a_max;
      row_limits(m_slack(ahi - u1), V(k, V_A_HI), -du1, mu, tau, o);
      if (k == 0) row_limits(m_slack(u1 - st.a0_lo), V(k, V_A_LO), du1, mu, tau, o);
      // rows of x_{k+1}
      row_limits(m_slack(x1a[2] - P.de_min), V(k, V_DE_LO), nx[2], mu, tau, o);
      row_limits(m_slack(P.de_max - x1a[2]), V(k, V_DE_HI), -nx[2], mu, tau, o);
      row_limits(m_slack(x1a[3] - P.v_min), V(k, V_V_LO), nx[3], mu, tau, o);
      row_limits(m_slack(P.v_max - x1a[3]), V(k, V_V_HI), -nx[3], mu, tau, o);
      const T sn1 = TR(k + 1, 0), cs1 = TR(k + 1, 1);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        T h, gx, gy, gp; obst(j, x1a[0], x1a[1], sn1, cs1, h, gx, gy, gp);
        const T s = S(k, j);
        const T r = (h - P.r_sum) - s;
        const T ds = gx * nx[0] + gy * nx[1] + gp * nx[4] + r;
        o.c1 += m_resid(r, h);
        row_limits(s, V(k, V_OB0 + j), ds, mu, tau, o);
      }
      // store the step in place of the gains of this stage (dead from here on)
#pragma unroll
      for (int j = 0; j < 5; ++j) { DX(k, j) = nx[j]; dx[j] = nx[j]; o.step_inf = m_max(o.step_inf, m_abs(nx[j])); }
      DU(k, 0) = du0; DU(k, 1) = du1;
      o.step_inf = m_max(o.step_inf, m_max(m_abs(du0), m_abs(du1)));
#pragma unroll
      for (int j = 0; j < 5; ++j) { x0d[j] = x1d[j]; x0a[j] = x1a[j]; }
    }
    return o;
  }

  // ---------------------------------------------------------------- merit difference phi(alpha) - phi(0)
  // dphi: cost + barrier difference (computed term by term, no cancellation); c1: l1 infeasibility at the trial point;
  // nz: magnitude sum of the terms (for the rounding-noise allowance); returns false if a slack would leave the interior.
  // reshoot = true: second-order correction.  Instead of the linear step x + al*dx the trial states are re-simulated
  // from the trial controls, x^_{k+1} = x^_k + dt f(x^_k, u_k + al*du_k) - (1 - al) d_k, so the dynamics defects shrink
  // EXACTLY by (1 - al) and the l1 merit cannot reject a good Newton step because of second-order defect growth (Maratos
  // effect).  The re-simulated step overwrites DX so that commit() applies it.
  MPC_HD bool trial(const ProbState<T>& st, T al, T& dphi, T& c1, T& nz, bool reshoot = false) const {
    const int N = P.N;
    const T mu = st.mu;
    dphi = T(0); c1 = T(0); nz = T(0);
    T xad[5], xaa[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) { xad[j] = X(0, j); xaa[j] = xad[j] + xr(0, j); }
    Trig ta; ta.sn = TR(0, 0); ta.cs = TR(0, 1); ta.tn = TR(0, 2);
    bool ok = true;
    T lg = T(0), lga = T(0);
    auto lrow = [&](T num, T den) {
      const T rt = num / den;
      ok = ok && (rt > T(-1));
      const T l = m_log1p(m_max(rt, T(-0.999999)));
      lg += l; lga += m_abs(l);
    };
    for (int k = 0; k < N; ++k) {
      const T u0 = U(k, 0), u1 = U(k, 1);
      const T du0 = al * DU(k, 0), du1 = al * DU(k, 1);
      const T nu0 = u0 + du0, nu1 = u1 + du1;
      T xbd[5], xba[5], dxb[5], x1a[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        dxb[j] = al * DX(k, j);
        const T xd = X(k + 1, j);
        x1a[j] = xd + xr(k + 1, j);
        xbd[j] = xd + dxb[j];
      }
      T d[5];
      if (reshoot) {
        // current defect of this stage (old point), then place x^_{k+1} so that the new defect is (1 - al) * d_old
        T x0d[5], x1d[5], dold[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) { x0d[j] = X(k, j); x1d[j] = X(k + 1, j); }
        Trig t0; t0.sn = TR(k, 0); t0.cs = TR(k, 1); t0.tn = TR(k, 2);
        defect(k, x0d, x1d, x0d[3] + xr(k, 3), t0, u0, u1, dold);
        T zero[5] = {T(0), T(0), T(0), T(0), T(0)};
        T roll[5];
        defect(k, xad, zero, xaa[3], ta, nu0, nu1, roll);      // = xt^_k + dt f(x^_k, u^_k) + c_k
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          xbd[j] = roll[j] - (T(1) - al) * dold[j];
          dxb[j] = xbd[j] - x1d[j];
          DX(k, j) = dxb[j] / al;
          d[j] = (T(1) - al) * dold[j];
        }
      } else {
        defect(k, xad, xbd, xaa[3], ta, nu0, nu1, d);
      }
#pragma unroll
      for (int j = 0; j < 5; ++j) { xba[j] = xbd[j] + xr(k + 1, j); c1 += m_abs(d[j]); }
      // cost difference (exact, no cancellation)
      {
        const T t0 = P.R[0] * du0 * (T(2) * u0 + du0), t1 = P.R[1] * du1 * (T(2) * u1 + du1);
        dphi += t0 + t1; nz += m_abs(t0) + m_abs(t1);
      }
      if (k + 1 <= N - 1) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const T t0 = P.Q[j] * dxb[j] * (T(2) * X(k + 1, j) + dxb[j]);
          dphi += t0; nz += m_abs(t0);
        }
      }
      // barrier differences: -mu * log(s_new / s_old)
      lrow(du0, m_slack(u0 - P.dd_min));
      lrow(-du0, m_slack(P.dd_max - u0));
      const T ahi = (k == 0) ? st.a0_hi : P.a_max;
      lrow(-du1, m_slack(ahi - u1));
      if (k == 0) lrow(du1, m_slack(u1 - st.a0_lo));
      lrow(dxb[2], m_slack(x1a[2] - P.de_min));
      lrow(-dxb[2], m_slack(P.de_max - x1a[2]));
      lrow(dxb[3], m_slack(x1a[3] - P.v_min));
      lrow(-dxb[3], m_slack(P.v_max - x1a[3]));
      // obstacle rows: the slack moves with its own Newton step ds (from the OLD linearisation)
      const T sn0 = TR(k + 1, 0), cs0 = TR(k + 1, 1);
      Trig tb; m_sincos(xba[4], &tb.sn, &tb.cs); tb.tn = m_tan(xba[2]);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        T h, gx, gy, gp; obst(j, x1a[0], x1a[1], sn0, cs0, h, gx, gy, gp);
        const T s = S(k, j);
        const T r = (h - P.r_sum) - s;
        const T ds = al * (gx * DX(k, 0) + gy * DX(k, 1) + gp * DX(k, 4) + r);
        lrow(ds, s);
        T hb, g1, g2, g3; obst(j, xba[0], xba[1], tb.sn, tb.cs, hb, g1, g2, g3);
        c1 += m_resid((hb - P.r_sum) - (s + ds), hb);
      }
#pragma unroll
      for (int j = 0; j < 5; ++j) { xad[j] = xbd[j]; xaa[j] = xba[j]; }
      ta = tb;
    }
    dphi -= mu * lg;
    nz += mu * lga;
    return ok;
  }

  // ---------------------------------------------------------------- commit the step
  MPC_HD void commit(ProbState<T>& st, T al, T ad) const {
    const int N = P.N;
    const T mu = st.mu;
    for (int k = 0; k < N; ++k) {
      const T u0 = U(k, 0), u1 = U(k, 1);
      const T du0 = DU(k, 0), du1 = DU(k, 1);
      T dx[5], x1a[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) { dx[j] = DX(k, j); x1a[j] = xa(k + 1, j); }
      auto upd = [&](int slot, T s, T ds) {
        const T nu = V(k, slot);
        const T dnu = (mu - nu * s - nu * ds) / s;
        V(k, slot) = m_max(nu + ad * dnu, T(1e-30));
      };
      upd(V_DD_LO, m_slack(u0 - P.dd_min), du0);
      upd(V_DD_HI, m_slack(P.dd_max - u0), -du0);
      const T ahi = (k == 0) ? st.a0_hi : P.a_max;
      upd(V_A_HI, m_slack(ahi - u1), -du1);
      if (k == 0) upd(V_A_LO, m_slack(u1 - st.a0_lo), du1);
      upd(V_DE_LO, m_slack(x1a[2] - P.de_min), dx[2]);
      upd(V_DE_HI, m_slack(P.de_max - x1a[2]), -dx[2]);
      upd(V_V_LO, m_slack(x1a[3] - P.v_min), dx[3]);
      upd(V_V_HI, m_slack(P.v_max - x1a[3]), -dx[3]);
      const T sn1 = TR(k + 1, 0), cs1 = TR(k + 1, 1);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        T h, gx, gy, gp; obst(j, x1a[0], x1a[1], sn1, cs1, h, gx, gy, gp);
        const T s = S(k, j);
        const T r = (h - P.r_sum) - s;
        const T ds = gx * dx[0] + gy * dx[1] + gp * dx[4] + r;
        upd(V_OB0 + j, s, ds);
        S(k, j) = m_slack(s + al * ds);
      }
      U(k, 0) = u0 + al * du0; U(k, 1) = u1 + al * du1;
#pragma unroll
      for (int j = 0; j < 5; ++j) X(k + 1, j) += al * dx[j];
    }
  }

  // ---------------------------------------------------------------- complementarity statistics after a step
  MPC_HD void compl_stats(const ProbState<T>& st, T& avg, T& cmax) const {
    const int N = P.N;
    T sum = T(0); cmax = T(0); int cnt = 0;
    auto acc = [&](T s, T nu) { const T c = s * nu; sum += c; cmax = m_max(cmax, c); ++cnt; };
    for (int k = 0; k < N; ++k) {
      const T u0 = U(k, 0), u1 = U(k, 1);
      acc(m_slack(u0 - P.dd_min), V(k, V_DD_LO)); acc(m_slack(P.dd_max - u0), V(k, V_DD_HI));
      const T ahi = (k == 0) ? st.a0_hi : P.a_max;
      acc(m_slack(ahi - u1), V(k, V_A_HI));
      if (k == 0) acc(m_slack(u1 - st.a0_lo), V(k, V_A_LO));
      const T de = xa(k + 1, 2), vv = xa(k + 1, 3);
      acc(m_slack(de - P.de_min), V(k, V_DE_LO)); acc(m_slack(P.de_max - de), V(k, V_DE_HI));
      acc(m_slack(vv - P.v_min), V(k, V_V_LO)); acc(m_slack(P.v_max - vv), V(k, V_V_HI));
#pragma unroll
      for (int j = 0; j < 3; ++j) acc(S(k, j), V(k, V_OB0 + j));
    }
    avg = sum / T(cnt);
  }

  // ---------------------------------------------------------------- one SQP / interior-point iteration
  MPC_HD void iterate(ProbState<T>& st) const {
    if (st.done) return;
    if (!backward(st, P.hessian)) backward(st, HESS_GN);
    FwdOut f = forward(st);
    if (!m_finite(f.step_inf) || !m_finite(f.dphi)) { st.status = ST_NAN; st.done = 1; return; }
    // penalty parameter of the l1 merit
    if (f.c1 > T(0)) {
      const T need = f.dphi / (T(0.5) * f.c1);
      if (need > st.rho) st.rho = need * T(1.5) + T(1);
    }
    const T slope = f.dphi - st.rho * f.c1;
    const T epsm = m_eps(T(0));
    T al = f.a_p;
    bool accepted = false;
    for (int t = 0; t < P.ls_max; ++t) {
      T dphi, c1, nz;
      bool ok = trial(st, al, dphi, c1, nz);
      T dm = dphi + st.rho * (c1 - f.c1);
      T noise = T(8) * epsm * (nz + st.rho * f.mag);
#ifdef MPC_DEBUG_LS
      printf("      ls t=%d al=%.3e ok=%d dphi=%.4e c1=%.4e (c1_0 %.4e) dm=%.4e thresh=%.4e noise=%.3e nz=%.3e mag=%.3e\n", t, (double)al, (int)ok,
             (double)dphi, (double)c1, (double)f.c1, (double)dm, (double)(T(1e-4) * al * slope + noise), (double)noise, (double)nz, (double)f.mag);
#endif
      if (ok && m_finite(dm) && dm <= T(1e-4) * al * slope + noise) { accepted = true; break; }
      if (t == 0 && ok && m_finite(dm)) {
        // second-order correction (see trial): re-simulate the states from the trial controls
        ok = trial(st, al, dphi, c1, nz, true);
        dm = dphi + st.rho * (c1 - f.c1);
        noise = T(8) * epsm * (nz + st.rho * f.mag);
        if (ok && m_finite(dm) && dm <= T(1e-4) * al * slope + noise) { accepted = true; st.nsoc++; break; }
      }
      al *= T(0.5);
    }
    if (!accepted) {
      st.nfail++;
      if (st.nfail >= 3) { st.status = ST_NOPROGRESS; st.done = 1; return; }
    } else {
      st.nfail = 0;
    }
    commit(st, al, f.a_d);
    st.d_al = al; st.d_ap = f.a_p; st.d_ad = f.a_d; st.d_c1 = f.c1; st.d_dphi = f.dphi; st.d_blk = f.blk;
    st.iters++;
    st.kkt = f.step_inf;
    // barrier update + convergence
    T avg, cmax; compl_stats(st, avg, cmax);
    if (st.mu <= P.mu_min * T(1.0001) && al >= T(0.5) && al * f.step_inf <= P.tol_step && f.c1 <= P.tol_feas) {
      st.status = ST_OPTIMAL; st.done = 1; return;
    }
    if (al >= T(0.5)) {
      const T mu_new = m_max(P.mu_min, m_min(st.mu, P.mu_factor * avg));
      if (mu_new < st.mu) st.rho = m_max(T(1), st.rho * T(0.5));
      st.mu = mu_new;
    }
  }
};

}  // namespace mpcb200
