// warp_core.cuh -- warp-cooperative nonlinear-MPC solver core: ONE WARP PER EGO INSTANCE (one NLP).
//
// Solves the NLP the reference hands to IPOPT each MPC step (/root/reference/MPC_Planner/optimizer.py:513-560;
// constraints :373-411, bounds :413-491, cost :493-511) with a Gauss-Newton / Newton SQP-type primal-dual interior-point
// iteration (l1-merit line search with second-order correction, monotone barrier) mapped onto 32 lanes:
//
//   * phases that are independent per stage -- linearisation of the Euler multiple-shooting defects
//     (optimizer.py:380-382) and the 10 inequality rows per stage into the stage KKT block, step-length limits,
//     line-search merit terms, the step commit -- run LANE = STAGE (k = lane, lane+32, ...), with warp reductions
//     (REDUX / shuffle butterflies) for the merit, the fraction-to-the-boundary limits and the convergence test;
//   * the backward Riccati sweep (block LDL^T of the block-tridiagonal KKT matrix in stage order
//     u_0, x_1, u_1, x_2, ...) runs LANE = MATRIX ENTRY: the cost-to-go is the 5x6 augmented block [P | p]
//     (lane 6i+j holds entry (i,j), only the upper triangle of P is ever a shuffle source, so P stays exactly
//     symmetric); one stage = 15 shuffles in two rounds, ~25 FMAs, one reciprocal of the 2x2 control block
//     determinant.  A_k = I + dt*df/dx has 6 off-identity entries and B = dt*[e_delta e_v] is constant, so the
//     products P*A and A^T*M touch at most 3-5 sources per entry;
//   * the forward sweep runs LANE = STATE COMPONENT (5 lanes + 5 broadcast shuffles per stage).
//
// The per-problem KKT slab (iterate, reference, multipliers, slacks, trig cache, stage KKT blocks, gains, step) is
// 90N+21 words (94N+21 with the exact-Hessian adjoint words) and lives in shared memory; stage records have an odd stride so LANE = STAGE accesses are
// bank-conflict free and LANE = ENTRY accesses of one record are contiguous.
//
// New code: the reference contains no solver of its own (it calls casadi/IPOPT).  The same source compiles for the
// device (nvcc) and, through the fiber emulator in warp_ctx.cuh, for the host-side algorithm tests (tests/host_sim).
#pragma once
#include "mpc_types.cuh"
#include "warp_ctx.cuh"

namespace mpcb200 {

// ------------------------------------------------------------------ slab layout (words of T per problem)
// stage record k = 0..N-1  (terms of u_k, of the dynamics x_k -> x_{k+1} and of x_{k+1})
enum : int {
  R_U = 0,      // 2  controls u_k
  R_V = 2,      // 11 multipliers of the inequality rows (slot order: V_DD_LO ... V_OB2, mpc_types.cuh)
  R_S = 13,     // 3  obstacle slacks
  R_CP = 16,    // 2  reference position increment rho_k - rho_{k+1} (from float64)
  R_E = 18,     // 6  off-identity entries of A_k: e03 e04 e13 e14 e42 e43
  R_D = 24,     // 5  dynamics defect d_k
  R_ZERO = 29,  // 1  constant 0 (target of structurally-zero coefficient indices)
  R_ONE = 30,   // 1  constant 1
  R_H = 31,     // 8  Hessian of the x_{k+1} terms: h00 h01 h04 h11 h14 h44 h22 h33
  R_GX = 39,    // 5  gradient of the x_{k+1} terms (barrier gradient at the current mu)
  R_RU = 44,    // 4  Ru0/dt^2 Ru1/dt^2 ru0/dt ru1/dt (control Hessian diagonal / gradient incl. barrier terms, pre-scaled for the sweep)
  R_KK = 48,    // 12 gains times dt: dt*(K0[0..4] k0 K1[0..4] k1)
  R_DX = 60,    // 5  step dx_{k+1}
  R_DU = 65,    // 2  step du_k
  R_FAR = 67,   // 1  obstacle rows of x_{k+1} screened out this iteration (1) or live (0)
  R_LC = 68,    // 5  multiplier-weighted gradient of the x_{k+1} terms (adjoint recursion: exact Hessian only, LAST so that
                //    Gauss-Newton kernels use a shorter record)
  REC_STRIDE = 73,      // record stride with the adjoint terms (exact Hessian, host emulator)
  REC_STRIDE_GN = 69    // Gauss-Newton kernels: 68 words + 1 pad (odd strides keep lane = stage accesses bank-conflict free);
                        // 480 B less per problem at N = 30 = the difference between 7 and 8 resident CTAs per SM
};
MPC_HD constexpr int rec_stride_for(int hessian_mode) { return hessian_mode == 0 /* HESS_GN */ ? REC_STRIDE_GN : REC_STRIDE; }
// state record k = 0..N
enum : int {
  S_XR = 0,     // 5 rho_k (reference row paired with stage k, frame shifted to the pinned position)
  S_XT = 5,     // 5 deviation state xt_k = x_k - rho_k
  S_TR = 10,    // 3 sin(psi_k) cos(psi_k) tan(delta_k)
  S_XTT = 13,   // 5 trial deviation state
  S_TRT = 18,   // 3 trig of the trial state
  ST_STRIDE = 21
};
enum : int { E03 = 0, E04, E13, E14, E42, E43 };
enum : int { H00 = 0, H01, H04, H11, H14, H44, H22, H33 };

struct WLayout {
  int N, o_state, o_rec, words;
  MPC_HD explicit WLayout(int N_, int rec_stride = REC_STRIDE) : N(N_) {
    o_state = 0;
    o_rec = ST_STRIDE * (N + 1);
    words = (o_rec + rec_stride * N + 3) & ~3;     // multiple of 4 words: 16-byte granularity for bulk copies
  }
};

// record offset of E[l][m] = (A_k - I)[l][m], or the ZERO slot
MPC_HD int e_off(int l, int m) {
  if (l == 0 && m == 3) return R_E + E03;
  if (l == 0 && m == 4) return R_E + E04;
  if (l == 1 && m == 3) return R_E + E13;
  if (l == 1 && m == 4) return R_E + E14;
  if (l == 4 && m == 2) return R_E + E42;
  if (l == 4 && m == 3) return R_E + E43;
  return R_ZERO;
}

// per-lane index tables of the serial sweeps (registers; computed once per kernel)
struct LaneTab {
  int i, j;        // entry of the 5x6 augmented block [P | p] this lane holds
  int own;         // 1 for the affine column j == 5
  int hidx;        // record offset of the x_{k+1} Hessian / gradient term this lane accumulates
  int exw;         // exact-Hessian addition: 0 none, 1 (4,4), 2 (3,4), 3 (2,2), 4 (2,3)
  int c[5];        // record offsets of Atilde[t][j], t = 0..4   (M = [P|p] * Atilde)
  int s1[5];       // source lanes of P[i][t] (upper-triangle storage)
  int e[3];        // record offsets of E[l_t][i]                  (Pn = A^T M - H^T G^-1 H)
  int s2[3];       // source lanes of M[l_t][j]
  int s_m2j, s_m3j, s_m2i, s_m3i;
  int fc[5], fc0;  // forward sweep, lanes 0..4: coefficients of row `lane` and of the constant term
  MPC_HD explicit LaneTab(int lane) {
    const int l = lane < 30 ? lane : 0;
    i = l / 6; j = l % 6;
    own = (j == 5) ? 1 : 0;
    hidx = R_ZERO; exw = 0;
    if (j == 5) hidx = R_GX + i;
    else if (i <= j) {
      if (i == 0 && j == 0) hidx = R_H + H00;
      if (i == 0 && j == 1) hidx = R_H + H01;
      if (i == 0 && j == 4) hidx = R_H + H04;
      if (i == 1 && j == 1) hidx = R_H + H11;
      if (i == 1 && j == 4) hidx = R_H + H14;
      if (i == 4 && j == 4) { hidx = R_H + H44; exw = 1; }
      if (i == 2 && j == 2) { hidx = R_H + H22; exw = 3; }
      if (i == 3 && j == 3) hidx = R_H + H33;
      if (i == 3 && j == 4) exw = 2;
      if (i == 2 && j == 3) exw = 4;
    }
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      c[t] = (j == 5) ? (R_D + t) : ((t == j) ? R_ONE : e_off(t, j));
      s1[t] = (i <= t) ? (6 * i + t) : (6 * t + i);
    }
    // rows l with E[l][i] != 0:  i = 2: {4},  i = 3: {0, 1, 4},  i = 4: {0, 1}   (static indexing only: registers)
    const int l0 = (i == 2) ? 4 : 0, l1 = 1, l2 = 4;
    e[0] = (i >= 2) ? e_off(l0, i) : R_ZERO;
    e[1] = (i >= 3) ? e_off(l1, i) : R_ZERO;
    e[2] = (i == 3) ? e_off(l2, i) : R_ZERO;
    s2[0] = 6 * l0 + j; s2[1] = 6 * l1 + j; s2[2] = 6 * l2 + j;
    s_m2j = 12 + j; s_m3j = 18 + j; s_m2i = 12 + i; s_m3i = 18 + i;
    const int r = lane < 5 ? lane : 0;
#pragma unroll
    for (int t = 0; t < 5; ++t) fc[t] = (r == 2) ? (R_KK + t) : (r == 3) ? (R_KK + 6 + t) : e_off(r, t);
    fc0 = (r == 2) ? (R_KK + 5) : (r == 3) ? (R_KK + 11) : R_ZERO;
  }
};

// The slab accessor.  On the device it indexes the CTA's dynamic shared memory by a 32-bit word offset, so every access
// is an LDS/STS with immediate field offsets (a generic `T*` member made nvcc fall back to 64-bit generic LD/ST).
#if defined(__CUDACC__)
extern __shared__ __align__(128) unsigned char mpc_dyn_smem[];
template <typename T>
struct SlabRef {
  int base;          // word offset of this problem's slab inside the dynamic shared memory
  MPC_HD T& operator[](int i) const { return reinterpret_cast<T*>(mpc_dyn_smem)[base + i]; }
};
#else
template <typename T>
struct SlabRef {
  T* p;
  int base;
  MPC_HD T& operator[](int i) const { return p[base + i]; }
};
#endif

template <bool B> struct FetchTag { static constexpr bool value = B; };      // compile-time "prefetch the next record" flag of the sweeps

// HM: Hessian mode fixed at compile time (HESS_GN / HESS_EXACT: the kernels, so that a Gauss-Newton kernel carries no
// exact-Hessian code between its hot phases -- instruction-cache footprint) or HESS_RUNTIME (P.hessian decides: the
// host emulator).
enum : int { HESS_RUNTIME = -1 };
template <typename T, int HM = HESS_RUNTIME>
struct WarpSolver {
  static constexpr int RS = (HM == 0 /* HESS_GN */) ? REC_STRIDE_GN : REC_STRIDE;      // stage-record stride of this instantiation
  const ParamsT<T>& P;
  const WLayout L;
  const SlabRef<T> sl;   // this problem's slab
  const T* obs;      // obstacle circle centres (centre, front, rear) in the problem's shifted frame
  const WarpCtx& w;
  const int lane;
  const LaneTab tb;
  const T il_wb;     // 1 / wheelbase
  const T idt_;      // 1 / dt
  const T irows;     // 1 / (number of inequality rows) = 1 / (10 N + 1)

  MPC_HD WarpSolver(const ParamsT<T>& P_, const SlabRef<T>& slab, const T* obs_, const WarpCtx& w_)
      : P(P_), L(P_.N, RS), sl(slab), obs(obs_), w(w_), lane(w_.lane()), tb(w_.lane()), il_wb(T(1) / P_.l_wb), idt_(T(1) / P_.dt),
        irows(T(1) / T(10 * P_.N + 1)) {}

  MPC_HD T& sx(int k, int f) const { return sl[L.o_state + ST_STRIDE * k + f]; }
  MPC_HD T& rc(int k, int f) const { return sl[L.o_rec + RS * k + f]; }
  MPC_HD T xa(int k, int j) const { return sx(k, S_XT + j) + sx(k, S_XR + j); }
  struct RecRef {        // one stage record (serial sweeps: every lane reads the same record)
    const SlabRef<T>& s; int o;
    MPC_HD T operator[](int f) const { return s[o + f]; }
  };

  struct Trig { T sn, cs, tn; };
  MPC_HD Trig trig_of(T psi, T delta) const { Trig t; m_sincos(psi, &t.sn, &t.cs); t.tn = m_tan(delta); return t; }

  // obstacle row j at (sx, sy, psi): distance + gradient (optimizer.py:384-403; distinct rows only, quirk Q6)
  MPC_HD void obst(int j, T px, T py, T sn, T cs, T& h, T& gx, T& gy, T& gp) const {
    const T sg = (j == 0) ? T(0) : (j == 1 ? T(1) : T(-1));
    const T o = sg * P.ego_off;
    const T dx = px + o * cs - obs[2 * j], dy = py + o * sn - obs[2 * j + 1];
    const T d2 = m_max(dx * dx + dy * dy, T(1e-24));
    const T ih = m_rsqrt(d2);
    h = d2 * ih;
    gx = dx * ih; gy = dy * ih;
    gp = o * (gy * cs - gx * sn);
  }

  // Row screening.  An obstacle row whose slack s = distance - r_sum is so large that its barrier curvature mu/s^2 is
  // below `screen_curv` contributes less than rounding error to the KKT system (lane following: dummy obstacle 130 m away,
  // quirk Q11).  Such rows are skipped for the iteration (decided once in linearize(), stored in R_FAR); their slack and
  // multiplier stay frozen and they count with s*nu = mu in the complementarity average.  Conservative test on the centre
  // distance: every circle pair is at least (centre distance - obs_spread - ego_off) apart.
  MPC_HD T far_threshold2(T mu) const {
    if (!(P.screen_inv_curv > T(0))) return T(3e38);       // screening off
    const T sp1 = m_sqrt_fast((obs[2] - obs[0]) * (obs[2] - obs[0]) + (obs[3] - obs[1]) * (obs[3] - obs[1]));
    const T sp2 = m_sqrt_fast((obs[4] - obs[0]) * (obs[4] - obs[0]) + (obs[5] - obs[1]) * (obs[5] - obs[1]));
    const T reach = P.r_sum + P.ego_off + m_max(sp1, sp2) + m_sqrt_fast(mu * P.screen_inv_curv);
    return reach * reach;
  }
  MPC_HD bool is_far(T px, T py, T thr2) const {
    const T dx = px - obs[0], dy = py - obs[1];
    return dx * dx + dy * dy > thr2;
  }

  // defect of stage k: d = xt_k - xt_{k+1} + dt*f(x_k,u_k) + (rho_k - rho_{k+1})      (optimizer.py:380-382, Euler)
  MPC_HD void defect(int k, const T* x0d, const T* x1d, T v, const Trig& t, T u0, T u1, T* d) const {
    const T dt = P.dt;
    d[0] = (x0d[0] - x1d[0]) + dt * v * t.cs + rc(k, R_CP);
    d[1] = (x0d[1] - x1d[1]) + dt * v * t.sn + rc(k, R_CP + 1);
    d[2] = (x0d[2] - x1d[2]) + dt * u0 + (sx(k, S_XR + 2) - sx(k + 1, S_XR + 2));
    d[3] = (x0d[3] - x1d[3]) + dt * u1 + (sx(k, S_XR + 3) - sx(k + 1, S_XR + 3));
    d[4] = (x0d[4] - x1d[4]) + dt * v * t.tn * il_wb + (sx(k, S_XR + 4) - sx(k + 1, S_XR + 4));
  }

  // ---------------------------------------------------------------- problem I/O (float64 row-major arrays of ONE problem)
  // xref [N+1][5] (row 0 = pinned state) is read lane = STAGE (on the device it sits in the warp's shared-memory staging,
  // filled by one TMA bulk copy); the warm start Xin [N+1][5], Uin [N][2] (either may be null: cold start = the reference's
  // step-0 guess, X_0 tiled and zero controls, optimizer.py:578-583) is read lane = ELEMENT straight from global / pinned host
  // memory: coalesced 256-byte requests, every byte touched once.  Differences are formed in float64, then rounded.
  MPC_HD void load(const double* xref, const double* Xin, const double* Uin, const double* obstacle_abs, T* obs_out) const {
    const int N = P.N;
    const double ox = xref[0], oy = xref[1];
    for (int j = 0; j < 3; ++j) { obs_out[2 * j] = (T)(obstacle_abs[2 * j] - ox); obs_out[2 * j + 1] = (T)(obstacle_abs[2 * j + 1] - oy); }
    for (int k = lane; k <= N; k += 32) {
      const double* rho = xref + 5 * ((k + 1 < N) ? (k + 1) : N);
      sx(k, S_XR + 0) = (T)(rho[0] - ox); sx(k, S_XR + 1) = (T)(rho[1] - oy);
      sx(k, S_XR + 2) = (T)rho[2]; sx(k, S_XR + 3) = (T)rho[3]; sx(k, S_XR + 4) = (T)rho[4];
      if (k == 0 || !Xin) {                                           // stage 0 is pinned to X_ref[:,0]; no warm start = X_0 tiled
        for (int j = 0; j < 5; ++j) sx(k, S_XT + j) = (T)(xref[j] - rho[j]);
      }
      if (k < N) {
        const double* rho1 = xref + 5 * ((k + 2 < N) ? (k + 2) : N);
        rc(k, R_CP) = (T)(rho[0] - rho1[0]); rc(k, R_CP + 1) = (T)(rho[1] - rho1[1]);
        if (!Uin) { rc(k, R_U) = T(0); rc(k, R_U + 1) = T(0); }
        rc(k, R_ZERO) = T(0); rc(k, R_ONE) = T(1);
      }
    }
    if (Xin) {
      for (int e = 5 + lane; e < 5 * (N + 1); e += 32) {
        const int k = e / 5, j = e - 5 * k;
        sx(k, S_XT + j) = (T)(Xin[e] - xref[5 * ((k + 1 < N) ? (k + 1) : N) + j]);
      }
    }
    if (Uin) {
      for (int e = lane; e < 2 * N; e += 32) rc(e >> 1, R_U + (e & 1)) = (T)Uin[e];
    }
    w.sync();
    if (P.init_rollout) {
      // single-shooting start: x_{k+1} = x_k + dt f(x_k, u_k) from the pinned state (float64, every lane redundantly;
      // lane k % 32 keeps stage k+1).  The caller's X rows are ignored.
      double x[5];
      for (int j = 0; j < 5; ++j) x[j] = xref[j];
      for (int k = 0; k < N; ++k) {
        const double u0 = (double)rc(k, R_U), u1 = (double)rc(k, R_U + 1);
        const double v = x[3], sn = sin(x[4]), cs = cos(x[4]), tn = tan(x[2]);
        x[0] += (double)P.dt * v * cs; x[1] += (double)P.dt * v * sn; x[2] += (double)P.dt * u0; x[3] += (double)P.dt * u1;
        x[4] += (double)P.dt * v * tn / (double)P.l_wb;
        if (lane == (k & 31)) {
          const double* rho = xref + 5 * ((k + 2 < N) ? (k + 2) : N);
          for (int j = 0; j < 5; ++j) sx(k + 1, S_XT + j) = (T)(x[j] - rho[j]);
        }
      }
      w.sync();
    }
  }
  // solution out, lane = ELEMENT: coalesced stores straight from the slab to global / pinned host memory (rho is added back
  // in float64 from the xref block)
  MPC_HD void store(const double* xref, double* Xout, double* Uout) const {
    const int N = P.N;
    w.sync();
    for (int e = lane; e < 5 * (N + 1); e += 32) {
      const int k = e / 5, j = e - 5 * k;
      Xout[e] = (k == 0) ? xref[j] : ((double)sx(k, S_XT + j) + xref[5 * ((k + 1 < N) ? (k + 1) : N) + j]);
    }
    for (int e = lane; e < 2 * N; e += 32) Uout[e] = (double)rc(e >> 1, R_U + (e & 1));
    w.sync();
  }

  // Feasibility tolerance of the PINNED stage: in a closed loop the pinned state is the plant's answer to the previous solution,
  // which sits ON its active bounds up to the rounding of the arithmetic (float32: delta_1 = delta_max + 1 ulp would otherwise
  // flag every later MPC step ST_INFEASIBLE_X0 -- measured: 19 % of the float32 Lanker closed-loop steps).
  MPC_HD static T x0_tol() { return sizeof(T) == 4 ? T(2e-5) : T(1e-9); }

  // ---------------------------------------------------------------- initialisation (lane = stage)
  // Pushes the start point strictly inside the bounds and puts slacks / multipliers on the central path; fills the
  // trig cache.  Every lane ends with the same (uniform) ProbState.
  MPC_HD void init(ProbState<T>& st) const {
    const int N = P.N;
    st.mu = P.mu0; st.rho = T(1); st.status = ST_MAXIT; st.iters = 0; st.done = 0; st.nfail = 0; st.nsoc = 0; st.nacc = 0; st.centered = 0; st.nstall = 0; st.best = T(1e30); st.kkt = T(0); st.pstep = T(0);
    st.d_al = st.d_ap = st.d_ad = st.d_c1 = st.d_dphi = T(0); st.d_blk = 0;
    const T de0 = xa(0, 2), v0 = xa(0, 3);
    const T s0 = v0 * v0 * m_tan(de0) / P.l_fric;
    bool bad = false;
    // friction row (optimizer.py:378, 424-425): |a0^2 + s0| <= a_max is the box |a0| <= sqrt(a_max - s0) when |s0| < a_max
    if (!(s0 < P.a_max) || !(s0 > -P.a_max)) bad = true;
    const T amax0 = m_sqrt(m_max(P.a_max - s0, T(1e-12)));
    st.a0_hi = m_min(amax0, P.a_max);
    st.a0_lo = -amax0;
    if (de0 < P.de_min - x0_tol() || de0 > P.de_max + x0_tol() || v0 < P.v_min - x0_tol() || v0 > P.v_max + x0_tol()) bad = true;
    {
      T sn, cs; m_sincos(xa(0, 4), &sn, &cs);
      for (int j = 0; j < 3; ++j) {
        T h, gx, gy, gp; obst(j, xa(0, 0), xa(0, 1), sn, cs, h, gx, gy, gp);
        if (h < P.r_sum - x0_tol()) bad = true;
      }
    }
    if (bad) { st.status = ST_INFEASIBLE_X0; st.done = 1; }
    const T kp = P.bound_push;
    if (lane == 0) {
      const Trig t = trig_of(xa(0, 4), xa(0, 2));
      sx(0, S_TR) = t.sn; sx(0, S_TR + 1) = t.cs; sx(0, S_TR + 2) = t.tn;
      sx(0, S_TRT) = t.sn; sx(0, S_TRT + 1) = t.cs; sx(0, S_TRT + 2) = t.tn;
      for (int j = 0; j < 5; ++j) sx(0, S_XTT + j) = sx(0, S_XT + j);
    }
    for (int k = lane; k < N; k += 32) {
      const T mu = st.mu;
      const T pdd = m_min(kp, kp * (P.dd_max - P.dd_min));
      T dd = m_min(m_max(rc(k, R_U), P.dd_min + pdd), P.dd_max - pdd);
      const T ahi = (k == 0) ? st.a0_hi : P.a_max;
      T a = rc(k, R_U + 1);
      if (k == 0) {
        const T pa = m_min(kp * m_max(T(1), ahi), kp * (ahi - st.a0_lo));
        a = m_min(m_max(a, st.a0_lo + pa), ahi - pa);
      } else {
        a = m_min(a, ahi - kp * m_max(T(1), m_abs(ahi)));
      }
      rc(k, R_U) = dd; rc(k, R_U + 1) = a;
      const T pde = m_min(kp * m_max(T(1), m_abs(P.de_max)), kp * (P.de_max - P.de_min));
      const T pv = m_min(kp * m_max(T(1), m_abs(P.v_max)), kp * (P.v_max - P.v_min));
      const T de = m_min(m_max(xa(k + 1, 2), P.de_min + pde), P.de_max - pde);
      const T vv = m_min(m_max(xa(k + 1, 3), P.v_min + pv), P.v_max - pv);
      sx(k + 1, S_XT + 2) = de - sx(k + 1, S_XR + 2);
      sx(k + 1, S_XT + 3) = vv - sx(k + 1, S_XR + 3);
      rc(k, R_V + V_DD_LO) = mu * m_rcp(dd - P.dd_min);
      rc(k, R_V + V_DD_HI) = mu * m_rcp(P.dd_max - dd);
      rc(k, R_V + V_A_HI) = mu * m_rcp(ahi - a);
      rc(k, R_V + V_A_LO) = (k == 0) ? mu * m_rcp(a - st.a0_lo) : T(0);
      rc(k, R_V + V_DE_LO) = mu * m_rcp(de - P.de_min);
      rc(k, R_V + V_DE_HI) = mu * m_rcp(P.de_max - de);
      rc(k, R_V + V_V_LO) = mu * m_rcp(vv - P.v_min);
      rc(k, R_V + V_V_HI) = mu * m_rcp(P.v_max - vv);
      const Trig t = trig_of(xa(k + 1, 4), xa(k + 1, 2));
      sx(k + 1, S_TR) = t.sn; sx(k + 1, S_TR + 1) = t.cs; sx(k + 1, S_TR + 2) = t.tn;
      for (int j = 0; j < 3; ++j) {
        T h, gx, gy, gp; obst(j, xa(k + 1, 0), xa(k + 1, 1), t.sn, t.cs, h, gx, gy, gp);
        const T c = h - P.r_sum;
        const T s = m_max(c, kp * m_max(T(1), P.r_sum));
        rc(k, R_S + j) = s;
        rc(k, R_V + V_OB0 + j) = mu * m_rcp(s);
      }
      for (int j = 0; j < 5; ++j) rc(k, R_DX + j) = T(0);
      rc(k, R_DU) = T(0); rc(k, R_DU + 1) = T(0);
    }
    w.sync();
  }

  // ---------------------------------------------------------------- dual warm start across MPC steps
  // The dual half of shift_movement (optimizer.py:652-653 shifts only the primal arrays; IPOPT restarts its multipliers
  // every step): inequality multipliers and obstacle slacks of stage k+1 become those of stage k, the last stage is repeated.
  MPC_HD void shift_duals() const {
    const int N = P.N;
    for (int k0 = 0; k0 < N; k0 += 32) {
      const int k = k0 + lane;
      T v[NV + 3];
      if (k < N) {
        const int ks = (k + 1 < N) ? k + 1 : N - 1;
#pragma unroll
        for (int j = 0; j < NV + 3; ++j) v[j] = rc(ks, R_V + j);          // R_V (11) and R_S (3) are adjacent
      }
      w.sync();
      if (k < N) {
#pragma unroll
        for (int j = 0; j < NV + 3; ++j) rc(k, R_V + j) = v[j];
      }
      w.sync();
    }
  }
  // Dual block of one problem in the caller's float64 layout (include/mpcb200.h, mpcb200_solve_dual): [N][14] = the 11
  // inequality multipliers of a stage (slot order V_DD_LO ... V_OB2) followed by its 3 obstacle slacks, then mu at exit and a
  // validity word.  lane = ELEMENT: coalesced, every byte touched once.  R_V (11) and R_S (3) are adjacent in the stage record.
  MPC_HD static int lam_words(int N) { return 14 * N + 2; }
  MPC_HD bool duals_valid(const double* lam) const { return lam[14 * P.N + 1] == 1.0; }
  MPC_HD void load_duals(const double* lam) const {
    const int N = P.N;
    for (int e = lane; e < 14 * N; e += 32) { const int k = e / 14, j = e - 14 * k; rc(k, R_V + j) = (T)lam[e]; }
    w.sync();
  }
  MPC_HD void store_duals(double* lam, T mu) const {
    const int N = P.N;
    w.sync();
    for (int e = lane; e < 14 * N; e += 32) { const int k = e / 14, j = e - 14 * k; lam[e] = (double)rc(k, R_V + j); }
    if (lane == 0) { lam[14 * N] = (double)mu; lam[14 * N + 1] = 1.0; }
    w.sync();
  }

  // New parameter block for a slab that keeps the previous MPC step's solution (controls, slacks, multipliers): only the
  // reference-derived words and the pinned stage are rewritten.
  MPC_HD void load_reference(const double* xref, const double* obstacle_abs, T* obs_out) const {
    const int N = P.N;
    const double ox = xref[0], oy = xref[1];
    for (int j = 0; j < 3; ++j) { obs_out[2 * j] = (T)(obstacle_abs[2 * j] - ox); obs_out[2 * j + 1] = (T)(obstacle_abs[2 * j + 1] - oy); }
    for (int k = lane; k <= N; k += 32) {
      const double* rho = xref + 5 * ((k + 1 < N) ? (k + 1) : N);
      sx(k, S_XR + 0) = (T)(rho[0] - ox); sx(k, S_XR + 1) = (T)(rho[1] - oy);
      sx(k, S_XR + 2) = (T)rho[2]; sx(k, S_XR + 3) = (T)rho[3]; sx(k, S_XR + 4) = (T)rho[4];
      if (k == 0) { for (int j = 0; j < 5; ++j) sx(0, S_XT + j) = (T)(xref[j] - rho[j]); }
      if (k < N) {
        const double* rho1 = xref + 5 * ((k + 2 < N) ? (k + 2) : N);
        rc(k, R_CP) = (T)(rho[0] - rho1[0]); rc(k, R_CP + 1) = (T)(rho[1] - rho1[1]);
      }
    }
    w.sync();
  }
  // Primal warm start of an MPC step from the previous step's controls (still in the slab).  Two candidate control
  // sequences -- the previous plan shifted one stage (shift_movement's guess, optimizer.py:652: right when the plan is
  // executed as predicted) and the previous plan as it is (right when the window is frozen, quirk Q8, or when the pinned
  // stage's friction box clamps u_0 every step, quirks Q3 / Q7) -- are rolled out from the new pinned state by two lanes
  // at once; the one with the lower cost + bound violation is kept, its single-shooting states become the start point (zero
  // dynamics defects), and controls / slacks / multipliers are shifted to match.  Returns the chosen shift (0 or 1).
  MPC_HD int warm_primal(const ProbState<T>& st0) const {
    const int N = P.N;
    const int c = lane & 1;                       // candidate of this lane: shift by c stages
    T a0_lo, a0_hi;
    {
      const T de0 = xa(0, 2), v0 = xa(0, 3);
      const T s0 = v0 * v0 * m_tan(de0) / P.l_fric;
      const T amax0 = m_sqrt(m_max(P.a_max - s0, T(1e-12)));
      a0_hi = m_min(amax0, P.a_max); a0_lo = -amax0;
    }
    (void)st0;
    T score = T(0);
    {
      T xd[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) xd[j] = sx(0, S_XT + j);
      const T zero[5] = {T(0), T(0), T(0), T(0), T(0)};
      for (int k = 0; k < N; ++k) {
        const int ks = (k + c < N) ? (k + c) : (N - 1);
        T u0 = m_min(m_max(rc(ks, R_U), P.dd_min), P.dd_max);
        T u1 = m_min(rc(ks, R_U + 1), P.a_max);
        if (k == 0) u1 = m_min(m_max(u1, a0_lo), a0_hi);
        const Trig t = trig_of(xd[4] + sx(k, S_XR + 4), xd[2] + sx(k, S_XR + 2));
        T roll[5];
        defect(k, xd, zero, xd[3] + sx(k, S_XR + 3), t, u0, u1, roll);
        score += P.R[0] * u0 * u0 + P.R[1] * u1 * u1;
        if (k + 1 <= N - 1) {
#pragma unroll
          for (int j = 0; j < 5; ++j) score += P.Q[j] * roll[j] * roll[j];
        }
        const T de = roll[2] + sx(k + 1, S_XR + 2), vv = roll[3] + sx(k + 1, S_XR + 3);
        const T viol = m_max(m_max(P.de_min - de, de - P.de_max), m_max(P.v_min - vv, vv - P.v_max));
        score += T(1e4) * m_max(viol, T(0));
#pragma unroll
        for (int j = 0; j < 5; ++j) xd[j] = roll[j];
      }
    }
    const T s_shift = w.shfl(score, 1), s_keep = w.shfl(score, 0);
    const int pick = (s_shift < s_keep) ? 1 : 0;          // uniform
    if (pick) {
      // shift controls, slacks and multipliers one stage (R_U, R_V, R_S are adjacent: 16 words)
      for (int k0 = 0; k0 < N; k0 += 32) {
        const int k = k0 + lane;
        T v[16];
        if (k < N) {
          const int ks = (k + 1 < N) ? k + 1 : N - 1;
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = rc(ks, R_U + j);
        }
        w.sync();
        if (k < N) {
#pragma unroll
          for (int j = 0; j < 16; ++j) rc(k, R_U + j) = v[j];
        }
        w.sync();
      }
    }
    // states of the chosen plan (every lane redundantly; lane k % 32 keeps stage k+1), controls clamped as rolled out
    {
      T xd[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) xd[j] = sx(0, S_XT + j);
      const T zero[5] = {T(0), T(0), T(0), T(0), T(0)};
      for (int k = 0; k < N; ++k) {
        T u0 = m_min(m_max(rc(k, R_U), P.dd_min), P.dd_max);
        T u1 = m_min(rc(k, R_U + 1), P.a_max);
        if (k == 0) u1 = m_min(m_max(u1, a0_lo), a0_hi);
        const Trig t = trig_of(xd[4] + sx(k, S_XR + 4), xd[2] + sx(k, S_XR + 2));
        T roll[5];
        defect(k, xd, zero, xd[3] + sx(k, S_XR + 3), t, u0, u1, roll);
        if (lane == (k & 31)) {
          rc(k, R_U) = u0; rc(k, R_U + 1) = u1;
#pragma unroll
          for (int j = 0; j < 5; ++j) sx(k + 1, S_XT + j) = roll[j];
        }
#pragma unroll
        for (int j = 0; j < 5; ++j) xd[j] = roll[j];
      }
    }
    w.sync();
    return pick;
  }
  // Start of a solve whose slab already holds slacks / multipliers (previous MPC step, shifted; or an imported dual block):
  // the iterate is pushed inside its bounds only by `warm_push`, every row keeps its multiplier -- brought to within
  // [mu / kappa_warm, kappa_warm * mu] / s of the central path of the restart barrier parameter mu_warm -- and the barrier
  // schedule restarts at mu_warm instead of mu0.  Every lane ends with the same (uniform) ProbState.
  MPC_HD void init_warm(ProbState<T>& st) const {
    const int N = P.N;
    st.mu = P.mu_warm; st.rho = T(1); st.status = ST_MAXIT; st.iters = 0; st.done = 0; st.nfail = 0; st.nsoc = 0; st.nacc = 0; st.centered = 0; st.nstall = 0; st.best = T(1e30); st.kkt = T(0); st.pstep = T(0);
    st.d_al = st.d_ap = st.d_ad = st.d_c1 = st.d_dphi = T(0); st.d_blk = 0;
    const T de0 = xa(0, 2), v0 = xa(0, 3);
    const T s0 = v0 * v0 * m_tan(de0) / P.l_fric;
    bool bad = false;
    if (!(s0 < P.a_max) || !(s0 > -P.a_max)) bad = true;
    const T amax0 = m_sqrt(m_max(P.a_max - s0, T(1e-12)));
    st.a0_hi = m_min(amax0, P.a_max);
    st.a0_lo = -amax0;
    if (de0 < P.de_min - x0_tol() || de0 > P.de_max + x0_tol() || v0 < P.v_min - x0_tol() || v0 > P.v_max + x0_tol()) bad = true;
    {
      T sn, cs; m_sincos(xa(0, 4), &sn, &cs);
      for (int j = 0; j < 3; ++j) {
        T h, gx, gy, gp; obst(j, xa(0, 0), xa(0, 1), sn, cs, h, gx, gy, gp);
        if (h < P.r_sum - x0_tol()) bad = true;
      }
    }
    if (bad) { st.status = ST_INFEASIBLE_X0; st.done = 1; }
    const T kp = P.warm_push, mu = st.mu, kap = P.kappa_warm, ikap = T(1) / P.kappa_warm;
    if (lane == 0) {
      const Trig t = trig_of(xa(0, 4), xa(0, 2));
      sx(0, S_TR) = t.sn; sx(0, S_TR + 1) = t.cs; sx(0, S_TR + 2) = t.tn;
      sx(0, S_TRT) = t.sn; sx(0, S_TRT + 1) = t.cs; sx(0, S_TRT + 2) = t.tn;
      for (int j = 0; j < 5; ++j) sx(0, S_XTT + j) = sx(0, S_XT + j);
    }
    auto recentre = [&](T nu, T s) { const T c = mu * m_rcp(s); return m_min(m_max(nu, c * ikap), c * kap); };
    for (int k = lane; k < N; k += 32) {
      const T pdd = m_min(kp, kp * (P.dd_max - P.dd_min));
      T dd = m_min(m_max(rc(k, R_U), P.dd_min + pdd), P.dd_max - pdd);
      const T ahi = (k == 0) ? st.a0_hi : P.a_max;
      T a = rc(k, R_U + 1);
      if (k == 0) {
        const T pa = m_min(kp * m_max(T(1), ahi), kp * (ahi - st.a0_lo));
        a = m_min(m_max(a, st.a0_lo + pa), ahi - pa);
      } else {
        a = m_min(a, ahi - kp * m_max(T(1), m_abs(ahi)));
      }
      rc(k, R_U) = dd; rc(k, R_U + 1) = a;
      const T pde = m_min(kp * m_max(T(1), m_abs(P.de_max)), kp * (P.de_max - P.de_min));
      const T pv = m_min(kp * m_max(T(1), m_abs(P.v_max)), kp * (P.v_max - P.v_min));
      const T de = m_min(m_max(xa(k + 1, 2), P.de_min + pde), P.de_max - pde);
      const T vv = m_min(m_max(xa(k + 1, 3), P.v_min + pv), P.v_max - pv);
      sx(k + 1, S_XT + 2) = de - sx(k + 1, S_XR + 2);
      sx(k + 1, S_XT + 3) = vv - sx(k + 1, S_XR + 3);
      rc(k, R_V + V_DD_LO) = recentre(rc(k, R_V + V_DD_LO), dd - P.dd_min);
      rc(k, R_V + V_DD_HI) = recentre(rc(k, R_V + V_DD_HI), P.dd_max - dd);
      rc(k, R_V + V_A_HI) = recentre(rc(k, R_V + V_A_HI), ahi - a);
      {                                                                               // exists at stage 0 only
        const T old = rc(k, R_V + V_A_LO);
        rc(k, R_V + V_A_LO) = (k == 0) ? ((old > T(0)) ? recentre(old, a - st.a0_lo) : mu * m_rcp(a - st.a0_lo)) : T(0);
      }
      rc(k, R_V + V_DE_LO) = recentre(rc(k, R_V + V_DE_LO), de - P.de_min);
      rc(k, R_V + V_DE_HI) = recentre(rc(k, R_V + V_DE_HI), P.de_max - de);
      rc(k, R_V + V_V_LO) = recentre(rc(k, R_V + V_V_LO), vv - P.v_min);
      rc(k, R_V + V_V_HI) = recentre(rc(k, R_V + V_V_HI), P.v_max - vv);
      const Trig t = trig_of(xa(k + 1, 4), xa(k + 1, 2));
      sx(k + 1, S_TR) = t.sn; sx(k + 1, S_TR + 1) = t.cs; sx(k + 1, S_TR + 2) = t.tn;
      for (int j = 0; j < 3; ++j) {
        T h, gx, gy, gp; obst(j, xa(k + 1, 0), xa(k + 1, 1), t.sn, t.cs, h, gx, gy, gp);
        const T c = h - P.r_sum;
        const T s = m_max(m_max(c, rc(k, R_S + j)), kp * m_max(T(1), P.r_sum));
        rc(k, R_S + j) = s;
        rc(k, R_V + V_OB0 + j) = recentre(rc(k, R_V + V_OB0 + j), s);
      }
      for (int j = 0; j < 5; ++j) rc(k, R_DX + j) = T(0);
      rc(k, R_DU) = T(0); rc(k, R_DU + 1) = T(0);
    }
    w.sync();
  }

  // ---------------------------------------------------------------- phase A: stage KKT blocks (lane = stage)
  MPC_HD void linearize(const ProbState<T>& st) const {
    const int N = P.N;
    const T dt = P.dt, mu = st.mu, idt = idt_;
    const T thr2 = far_threshold2(mu);
    for (int k = lane; k < N; k += 32) {
      T x0d[5], x1d[5], x1a[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) { x0d[j] = sx(k, S_XT + j); x1d[j] = sx(k + 1, S_XT + j); x1a[j] = x1d[j] + sx(k + 1, S_XR + j); }
      const T v = x0d[3] + sx(k, S_XR + 3);
      Trig t0, t1;
      t0.sn = sx(k, S_TR); t0.cs = sx(k, S_TR + 1); t0.tn = sx(k, S_TR + 2);
      t1.sn = sx(k + 1, S_TR); t1.cs = sx(k + 1, S_TR + 1);
      const T u0 = rc(k, R_U), u1 = rc(k, R_U + 1);
      // dynamics: A_k = I + dt*df/dx (configuration.py:364-368), defect
      const T sec2 = T(1) + t0.tn * t0.tn;
      rc(k, R_E + E03) = dt * t0.cs; rc(k, R_E + E04) = -dt * v * t0.sn;
      rc(k, R_E + E13) = dt * t0.sn; rc(k, R_E + E14) = dt * v * t0.cs;
      rc(k, R_E + E42) = dt * v * sec2 * il_wb; rc(k, R_E + E43) = dt * t0.tn * il_wb;
      T d[5];
      defect(k, x0d, x1d, v, t0, u0, u1, d);
#pragma unroll
      for (int j = 0; j < 5; ++j) rc(k, R_D + j) = d[j];
      // x_{k+1} terms: cost (stages 1..N-1 only, quirk Q1), box barriers, obstacle barriers
      T hd[5], g[5], lc[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) { hd[j] = T(0); g[j] = T(0); lc[j] = T(0); }
      if (k + 1 <= N - 1) {
#pragma unroll
        for (int j = 0; j < 5; ++j) { const T gq = T(2) * P.Q[j] * x1d[j]; hd[j] = T(2) * P.Q[j]; g[j] = gq; lc[j] = gq; }
      }
      {
        const T ilo = m_rcp(m_slack(x1a[2] - P.de_min)), ihi = m_rcp(m_slack(P.de_max - x1a[2]));
        const T vlo = rc(k, R_V + V_DE_LO), vhi = rc(k, R_V + V_DE_HI);
        hd[2] += vlo * ilo + vhi * ihi;
        g[2] += mu * (ihi - ilo);
        lc[2] += -vlo + vhi;
      }
      {
        const T ilo = m_rcp(m_slack(x1a[3] - P.v_min)), ihi = m_rcp(m_slack(P.v_max - x1a[3]));
        const T vlo = rc(k, R_V + V_V_LO), vhi = rc(k, R_V + V_V_HI);
        hd[3] += vlo * ilo + vhi * ihi;
        g[3] += mu * (ihi - ilo);
        lc[3] += -vlo + vhi;
      }
      T h01 = T(0), h04 = T(0), h14 = T(0);
      const bool far = is_far(x1a[0], x1a[1], thr2);
      rc(k, R_FAR) = far ? T(1) : T(0);
      if (!far) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        T h, gx, gy, gp; obst(j, x1a[0], x1a[1], t1.sn, t1.cs, h, gx, gy, gp);
        // slack reset s <- max(s, c(x)) (Nocedal & Wright 19.3)
        T s = rc(k, R_S + j);
        if (h - P.r_sum > s) { s = h - P.r_sum; rc(k, R_S + j) = s; }
        const T nu = rc(k, R_V + V_OB0 + j);
        const T is = m_rcp(s);
        const T wgt = nu * is, r = (h - P.r_sum) - s;
        const T cg = -(mu * is - wgt * r);
        hd[0] += wgt * gx * gx; h01 += wgt * gx * gy; h04 += wgt * gx * gp;
        hd[1] += wgt * gy * gy; h14 += wgt * gy * gp; hd[4] += wgt * gp * gp;
        g[0] += cg * gx; g[1] += cg * gy; g[4] += cg * gp;
        lc[0] -= nu * gx; lc[1] -= nu * gy; lc[4] -= nu * gp;
      }
      }
      rc(k, R_H + H00) = hd[0]; rc(k, R_H + H01) = h01; rc(k, R_H + H04) = h04; rc(k, R_H + H11) = hd[1];
      rc(k, R_H + H14) = h14; rc(k, R_H + H44) = hd[4]; rc(k, R_H + H22) = hd[2]; rc(k, R_H + H33) = hd[3];
#pragma unroll
      for (int j = 0; j < 5; ++j) { rc(k, R_GX + j) = g[j]; if (HM != HESS_GN) rc(k, R_LC + j) = lc[j]; }
      // control terms
      {
        const T ilo = m_rcp(m_slack(u0 - P.dd_min)), ihi = m_rcp(m_slack(P.dd_max - u0));
        T Ru0 = T(2) * P.R[0] + rc(k, R_V + V_DD_LO) * ilo + rc(k, R_V + V_DD_HI) * ihi;
        T ru0 = T(2) * P.R[0] * u0 + mu * (ihi - ilo);
        const T ahi = (k == 0) ? st.a0_hi : P.a_max;
        const T ish = m_rcp(m_slack(ahi - u1));
        T Ru1 = T(2) * P.R[1] + rc(k, R_V + V_A_HI) * ish;
        T ru1 = T(2) * P.R[1] * u1 + mu * ish;
        if (k == 0) {
          const T isl = m_rcp(m_slack(u1 - st.a0_lo));
          Ru1 += rc(k, R_V + V_A_LO) * isl;
          ru1 -= mu * isl;
        }
        rc(k, R_RU) = Ru0 * (idt * idt); rc(k, R_RU + 1) = Ru1 * (idt * idt); rc(k, R_RU + 2) = ru0 * idt; rc(k, R_RU + 3) = ru1 * idt;
      }
    }
    w.sync();
  }

  // ---------------------------------------------------------------- phase C: backward Riccati sweep (lane = entry of [P | p])
  // Returns false (uniformly) if an exact-Hessian control block was not positive definite (caller retries with GN).
  //
  // Per stage, with M = [P|p] * [A d; 0 1], G = R~ + dt^2 P[2:4,2:4], J = -dt^2 G^-1 and mm = M[2:4, j] (+ ru/dt on the
  // affine column):   T = J mm  ( = dt * [K | kff], stored as the gains ),   Pn = A^T M + M[2:4, i]^T T.
  // The stage loop is software-pipelined by hand: the record loads of stage k-1 are issued between the shuffles of
  // stage k and their consumers, so the LDS issue slots hide under the shuffle latency of the dependent chain.
  struct BwdCoef { T h, c0, c1, c2, c3, c4, e0, e1, e2, Ru0, Ru1, rp0, rp1; };

  template <int HESS>
  MPC_HD bool backward_t() const {
    const int N = P.N;
    const T dt = P.dt;
    T Pij = T(0);
    T lam[5] = {T(0), T(0), T(0), T(0), T(0)};
    const T ownf = (T)tb.own;
    bool ok = true;
    // per-lane record pointers, bumped one record per stage (one IADD each instead of index arithmetic per load)
    const int last = L.o_rec + RS * (N - 1);
    const T* ph = &sl[last + tb.hidx];
    const T* pc0 = &sl[last + tb.c[0]]; const T* pc1 = &sl[last + tb.c[1]]; const T* pc2 = &sl[last + tb.c[2]];
    const T* pc3 = &sl[last + tb.c[3]]; const T* pc4 = &sl[last + tb.c[4]];
    const T* pe0 = &sl[last + tb.e[0]]; const T* pe1 = &sl[last + tb.e[1]]; const T* pe2 = &sl[last + tb.e[2]];
    const T* pr = &sl[last + R_RU];
    T* pk = &sl[last + R_KK + tb.j];
    // loads the record the pointers stand on, then steps them one record down (constant bump: folded into the load offsets of
    // the two-stage loop body).  Nothing is read below record 0: the stage that would prefetch it is instantiated WITHOUT the
    // fetch (`more` is a compile-time tag), in the loop epilogue.
    auto fetch = [&](BwdCoef& q) {
      q.h = *ph; q.c0 = *pc0; q.c1 = *pc1; q.c2 = *pc2; q.c3 = *pc3; q.c4 = *pc4; q.e0 = *pe0; q.e1 = *pe1; q.e2 = *pe2;
      q.Ru0 = pr[0]; q.Ru1 = pr[1]; q.rp0 = pr[2]; q.rp1 = pr[3];
      ph -= RS; pc0 -= RS; pc1 -= RS; pc2 -= RS; pc3 -= RS; pc4 -= RS;
      pe0 -= RS; pe1 -= RS; pe2 -= RS; pr -= RS;
    };
    auto stage = [&](int k, const BwdCoef& cur, BwdCoef& nxt, auto more) -> bool {
      Pij += cur.h;                                   // x_{k+1} terms
      // round 1: control block + M = [P|p] * Atilde
      const T p22 = w.shfl(Pij, 14), p23 = w.shfl(Pij, 15), p33 = w.shfl(Pij, 21);
      const T q0 = w.shfl(Pij, tb.s1[0]), q1 = w.shfl(Pij, tb.s1[1]), q2 = w.shfl(Pij, tb.s1[2]);
      const T q3 = w.shfl(Pij, tb.s1[3]), q4 = w.shfl(Pij, tb.s1[4]);
      if (decltype(more)::value) fetch(nxt);   // record k-1, off the dependent chain
      const T M = ownf * Pij + ((cur.c0 * q0 + cur.c1 * q1) + (cur.c2 * q2 + cur.c3 * q3) + cur.c4 * q4);
      // G / dt^2 (the stored control diagonal is pre-divided): J = -dt^2 G^-1 = -(G / dt^2)^-1 needs no dt^2 on the chain
      const T G00 = cur.Ru0 + p22, G01 = p23, G11 = cur.Ru1 + p33;
      const T det = G00 * G11 - G01 * G01;
      if (HESS == HESS_EXACT) {
        if (!(G00 > T(0)) || !(det > T(1e-8) * G00 * G11)) return false;      // uniform across lanes
      }
      const T cdet = m_rcp(det);
      const T J00 = -cdet * G11, J01 = cdet * G01, J11 = -cdet * G00;
      // round 2
      const T m2j = w.shfl(M, tb.s_m2j), m3j = w.shfl(M, tb.s_m3j);
      const T m2i = w.shfl(M, tb.s_m2i), m3i = w.shfl(M, tb.s_m3i);
      const T r0 = w.shfl(M, tb.s2[0]), r1 = w.shfl(M, tb.s2[1]), r2 = w.shfl(M, tb.s2[2]);
      const T mm2 = m2j + ownf * cur.rp0, mm3 = m3j + ownf * cur.rp1;
      const T T0 = J00 * mm2 + J01 * mm3, T1 = J01 * mm2 + J11 * mm3;
      T Pn = (M + cur.e0 * r0) + (cur.e1 * r1 + cur.e2 * r2) + (m2i * T0 + m3i * T1);
      if (lane < 6) { pk[0] = T0; pk[6] = T1; }
      pk -= RS;
      if (HESS == HESS_EXACT) {
        // adjoint multipliers lam_{k+1} (every lane keeps the 5-vector), then
        // + dt * sum_i lam_{k+1,i} * hess f_i(x_k) on the (delta, v, psi) block of P_k
        const RecRef r{sl, L.o_rec + RS * k};
#pragma unroll
        for (int jj = 0; jj < 5; ++jj) lam[jj] += r[R_LC + jj];
        const T v = xa(k, 3);
        const T sn = sx(k, S_TR), cs = sx(k, S_TR + 1), tn = sx(k, S_TR + 2);
        const T sec2 = T(1) + tn * tn;
        T ex = T(0);
        if (tb.exw == 1) ex = dt * (-lam[0] * v * cs - lam[1] * v * sn);
        if (tb.exw == 2) ex = dt * (-lam[0] * sn + lam[1] * cs);
        if (tb.exw == 3) ex = dt * lam[4] * T(2) * v * il_wb * sec2 * tn;
        if (tb.exw == 4) ex = dt * lam[4] * sec2 * il_wb;
        Pn += ex;
        const T f03 = r[R_E + E03], f04 = r[R_E + E04], f13 = r[R_E + E13], f14 = r[R_E + E14];
        const T f42 = r[R_E + E42], f43 = r[R_E + E43];
        const T l2 = lam[2] + f42 * lam[4];
        const T l3 = lam[3] + f03 * lam[0] + f13 * lam[1] + f43 * lam[4];
        const T l4 = lam[4] + f04 * lam[0] + f14 * lam[1];
        lam[2] = l2; lam[3] = l3; lam[4] = l4;
      }
      Pij = Pn;
      return true;
    };
    // two stages per trip with ping-pong coefficient registers (no register moves between stages)
    BwdCoef ca, cb;
    const FetchTag<true> yes; const FetchTag<false> no;
    fetch(ca);                                            // record N-1
    int k = N - 1;
    for (; k >= 2; k -= 2) {                              // both stages of a trip have a record below them to prefetch
      if (!stage(k, ca, cb, yes)) { ok = false; break; }
      if (!stage(k - 1, cb, ca, yes)) { ok = false; break; }
    }
    if (ok) {
      if (k == 1) ok = stage(1, ca, cb, yes) && stage(0, cb, ca, no);
      else ok = stage(0, ca, cb, no);
    }
    w.sync();
    return ok;
  }
  // factor + solve with the configured Hessian; an indefinite exact-Hessian control block falls back to Gauss-Newton
  MPC_HD void backward() const {
    if (HM == HESS_GN) { backward_t<HESS_GN>(); return; }
    if (HM == HESS_EXACT || P.hessian == HESS_EXACT) { if (backward_t<HESS_EXACT>()) return; }
    backward_t<HESS_GN>();
  }

  // ---------------------------------------------------------------- phase D: forward sweep (lane = state component)
  // dx_{k+1} = A dx_k + B du_k + d_k with du_k = kff + K dx_k: lane r < 5 owns component r.  Rows 0, 1, 4 take their
  // coefficients from A_k, rows 2, 3 from the stored gains dt*[K | kff]; the 5 new components are re-broadcast by
  // shuffles.  Record loads are prefetched one stage ahead (off the dependent chain).
  struct FwdCoef { T f0, f1, f2, f3, f4, fc, d; };
  MPC_HD void forward_sweep() const {
    const int N = P.N;
    const T idt = idt_;
    const int row = lane < 5 ? lane : 0;
    T dx0 = T(0), dx1 = T(0), dx2 = T(0), dx3 = T(0), dx4 = T(0), mine = T(0);
    const T* pf0 = &sl[L.o_rec + tb.fc[0]]; const T* pf1 = &sl[L.o_rec + tb.fc[1]]; const T* pf2 = &sl[L.o_rec + tb.fc[2]];
    const T* pf3 = &sl[L.o_rec + tb.fc[3]]; const T* pf4 = &sl[L.o_rec + tb.fc[4]]; const T* pfc = &sl[L.o_rec + tb.fc0];
    T* pd = &sl[L.o_rec + row];
    const T* pl = &sl[L.o_rec + row];
    // loads record j (where the pointers stand), then steps them up; the last stage is instantiated without the fetch
    auto fetch = [&](FwdCoef& q) {
      q.f0 = *pf0; q.f1 = *pf1; q.f2 = *pf2; q.f3 = *pf3; q.f4 = *pf4; q.fc = *pfc; q.d = pl[R_D];
      pf0 += RS; pf1 += RS; pf2 += RS; pf3 += RS; pf4 += RS; pfc += RS; pl += RS;
    };
    auto stage = [&](const FwdCoef& cur, FwdCoef& nxt, auto more) {
      const T acc = (cur.fc + cur.f0 * dx0) + (cur.f1 * dx1 + cur.f2 * dx2) + (cur.f3 * dx3 + cur.f4 * dx4);
      const T nx = (mine + cur.d) + acc;
      dx0 = w.shfl(nx, 0); dx1 = w.shfl(nx, 1); dx2 = w.shfl(nx, 2); dx3 = w.shfl(nx, 3); dx4 = w.shfl(nx, 4);
      mine = nx;
      if (lane < 5) pd[R_DX] = nx;
      if (lane == 2 || lane == 3) pd[R_DU - 2] = acc * idt;
      pd += RS;
      if (decltype(more)::value) fetch(nxt);   // record k+1
    };
    FwdCoef ca, cb;
    const FetchTag<true> yes; const FetchTag<false> no;
    fetch(ca);                                            // record 0
    int k = 0;
    for (; k + 2 < N; k += 2) { stage(ca, cb, yes); stage(cb, ca, yes); }      // stages k, k+1 with a record k+2 <= N-1 after them
    if (k + 2 == N) { stage(ca, cb, yes); stage(cb, ca, no); }
    else stage(ca, cb, no);
    w.sync();
  }

  // ---------------------------------------------------------------- phase E: step-length limits, merit slope (lane = stage)
  // a_p / a_d hold, until the final reduction, the LARGEST relative decrease -ds/s and -dnu/nu over the rows; the
  // fraction-to-the-boundary step lengths are then min(1, tau / that) -- one division per iteration instead of per row.
  struct FwdOut { T a_p, a_d, lim, dphi, c1, step_inf, mag; int blk, cur; };

  MPC_HD void row_limits(T s, T nu, T ds, T mu, T tau, FwdOut& o) const {
    const T is = m_rcp(s), inu = m_rcp(nu);
    const T t = ds * is;                                  // ds / s
    const T q = (mu * is) * inu - T(1) - t;               // dnu / nu  with dnu = (mu - nu*s - nu*ds) / s
#ifdef MPC_DIAG
    if (-t > o.a_p) o.blk = o.cur;
    o.cur++;
#endif
    o.a_p = m_max(o.a_p, -t);
    o.a_d = m_max(o.a_d, -q);
    o.dphi -= mu * t;
  }

  MPC_HD FwdOut forward_stats(const ProbState<T>& st) const {
    const int N = P.N;
    const T dt = P.dt, mu = st.mu;
    const T tau = m_max(P.tau_min, T(1) - mu);
    FwdOut o; o.a_p = T(0); o.a_d = T(0); o.dphi = T(0); o.c1 = T(0); o.step_inf = T(0); o.mag = T(0); o.blk = -1; o.cur = 0;
    for (int k = lane; k < N; k += 32) {
      o.cur = 16 * k;
      T x0d[5], x1d[5], x1a[5], nx[5], d[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        x0d[j] = sx(k, S_XT + j); x1d[j] = sx(k + 1, S_XT + j); x1a[j] = x1d[j] + sx(k + 1, S_XR + j);
        nx[j] = rc(k, R_DX + j); d[j] = rc(k, R_D + j);
      }
      const T v = x0d[3] + sx(k, S_XR + 3);
      const T u0 = rc(k, R_U), u1 = rc(k, R_U + 1);
      const T du0 = rc(k, R_DU), du1 = rc(k, R_DU + 1);
#pragma unroll
      for (int j = 0; j < 5; ++j) { o.c1 += m_abs(d[j]); o.mag += m_abs(x0d[j]) + m_abs(x1d[j]); }
      o.mag += dt * (T(2) * m_abs(v) + m_abs(u0) + m_abs(u1)) + m_abs(rc(k, R_CP)) + m_abs(rc(k, R_CP + 1));
      o.dphi += T(2) * P.R[0] * u0 * du0 + T(2) * P.R[1] * u1 * du1;
      if (k + 1 <= N - 1) {
#pragma unroll
        for (int j = 0; j < 5; ++j) o.dphi += T(2) * P.Q[j] * x1d[j] * nx[j];
      }
      row_limits(m_slack(u0 - P.dd_min), rc(k, R_V + V_DD_LO), du0, mu, tau, o);
      row_limits(m_slack(P.dd_max - u0), rc(k, R_V + V_DD_HI), -du0, mu, tau, o);
      const T ahi = (k == 0) ? st.a0_hi : P.a_max;
      row_limits(m_slack(ahi - u1), rc(k, R_V + V_A_HI), -du1, mu, tau, o);
      if (k == 0) row_limits(m_slack(u1 - st.a0_lo), rc(k, R_V + V_A_LO), du1, mu, tau, o);
      row_limits(m_slack(x1a[2] - P.de_min), rc(k, R_V + V_DE_LO), nx[2], mu, tau, o);
      row_limits(m_slack(P.de_max - x1a[2]), rc(k, R_V + V_DE_HI), -nx[2], mu, tau, o);
      row_limits(m_slack(x1a[3] - P.v_min), rc(k, R_V + V_V_LO), nx[3], mu, tau, o);
      row_limits(m_slack(P.v_max - x1a[3]), rc(k, R_V + V_V_HI), -nx[3], mu, tau, o);
      const T sn1 = sx(k + 1, S_TR), cs1 = sx(k + 1, S_TR + 1);
      if (rc(k, R_FAR) == T(0)) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        T h, gx, gy, gp; obst(j, x1a[0], x1a[1], sn1, cs1, h, gx, gy, gp);
        const T s = rc(k, R_S + j);
        const T r = (h - P.r_sum) - s;
        const T ds = gx * nx[0] + gy * nx[1] + gp * nx[4] + r;
        o.c1 += m_resid(r, h);
        row_limits(s, rc(k, R_V + V_OB0 + j), ds, mu, tau, o);
      }
      }
#pragma unroll
      for (int j = 0; j < 5; ++j) o.step_inf = m_max(o.step_inf, m_abs(nx[j]));
      o.step_inf = m_max(o.step_inf, m_max(m_abs(du0), m_abs(du1)));
    }
    const bool fin = m_finite(o.step_inf) && m_finite(o.dphi) && m_finite(o.a_p) && m_finite(o.a_d);
    const bool allfin = w.all(fin);
#ifdef MPC_DIAG
    { const T mine = o.a_p; const T mx = w.max_nonneg(m_max(o.a_p, T(0))); int code = (mine == mx) ? o.blk : -1;
      for (int m = 16; m; m >>= 1) { const int other = w.shfl_xor(code, m); code = code > other ? code : other; }
      o.blk = code; }
#endif
    o.a_p = w.max_nonneg(m_max(o.a_p, T(0))); o.a_d = w.max_nonneg(m_max(o.a_d, T(0)));
    o.lim = tau * m_rcp(m_max(m_max(o.a_p, o.a_d), T(1e-3)));        // uncapped fraction-to-the-boundary length (primal and dual)
    o.a_p = (o.a_p > tau) ? tau * m_rcp(o.a_p) : T(1);
    o.a_d = (o.a_d > tau) ? tau * m_rcp(o.a_d) : T(1);
    o.step_inf = w.max_nonneg(fin ? o.step_inf : T(0));
    o.dphi = w.sum(o.dphi); o.c1 = w.sum(o.c1); o.mag = w.sum(o.mag);
    if (!allfin) o.step_inf = T(NAN);     // the caller turns this into ST_NAN
    return o;
  }

  // ---------------------------------------------------------------- phase F: merit difference phi(alpha) - phi(0) (lane = stage)
  // trial_points: x^_{k+1} = x_{k+1} + al*dx_{k+1} and its trig (reused by the next linearisation when accepted).
  MPC_HD void trial_points(T al) const {
    const int N = P.N;
    for (int k = lane; k < N; k += 32) {
      T xb[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) { xb[j] = sx(k + 1, S_XT + j) + al * rc(k, R_DX + j); sx(k + 1, S_XTT + j) = xb[j]; }
      const Trig t = trig_of(xb[4] + sx(k + 1, S_XR + 4), xb[2] + sx(k + 1, S_XR + 2));
      sx(k + 1, S_TRT) = t.sn; sx(k + 1, S_TRT + 1) = t.cs; sx(k + 1, S_TRT + 2) = t.tn;
    }
    w.sync();
  }
  // second-order correction: re-simulate the trial states from the trial controls,
  //   x^_{k+1} = x^_k + dt f(x^_k, u_k + al*du_k) - (1 - al) d_k,
  // so the dynamics defects shrink EXACTLY by (1 - al) (Maratos effect); the re-simulated step overwrites DX so that
  // the multiplier / slack updates of commit() see it.  Serial in k, computed redundantly by every lane (rare path).
  MPC_HD void trial_points_reshoot(T al) const {
    const int N = P.N;
    T xd[5], xaa[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) { xd[j] = sx(0, S_XT + j); xaa[j] = xd[j] + sx(0, S_XR + j); }
    Trig ta; ta.sn = sx(0, S_TR); ta.cs = sx(0, S_TR + 1); ta.tn = sx(0, S_TR + 2);
    const T zero[5] = {T(0), T(0), T(0), T(0), T(0)};
    for (int k = 0; k < N; ++k) {
      const T nu0 = rc(k, R_U) + al * rc(k, R_DU), nu1 = rc(k, R_U + 1) + al * rc(k, R_DU + 1);
      T roll[5];
      defect(k, xd, zero, xaa[3], ta, nu0, nu1, roll);      // = xt^_k + dt f(x^_k, u^_k) + c_k
      T xb[5], dxn[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        xb[j] = roll[j] - (T(1) - al) * rc(k, R_D + j);
        dxn[j] = (xb[j] - sx(k + 1, S_XT + j)) / al;
        xaa[j] = xb[j] + sx(k + 1, S_XR + j);
        xd[j] = xb[j];
      }
      ta = trig_of(xaa[4], xaa[2]);
      if (lane == (k & 31)) {
#pragma unroll
        for (int j = 0; j < 5; ++j) { sx(k + 1, S_XTT + j) = xb[j]; rc(k, R_DX + j) = dxn[j]; }
        sx(k + 1, S_TRT) = ta.sn; sx(k + 1, S_TRT + 1) = ta.cs; sx(k + 1, S_TRT + 2) = ta.tn;
      }
    }
    w.sync();
  }

  // dphi: cost + barrier difference (term by term, no cancellation); c1: l1 infeasibility at the trial point;
  // nz: magnitude sum of the terms (rounding-noise allowance); returns false if a slack would leave the interior.
  MPC_HD bool trial_merit(const ProbState<T>& st, T al, bool reshoot, T& dphi, T& c1, T& nz) const {
    const int N = P.N;
    const T mu = st.mu;
    dphi = T(0); c1 = T(0); nz = T(0);
    bool ok = true;
    T lg = T(0), lga = T(0);
    // The barrier terms -mu*log(s_new/s_old) = -mu*log1p(rt) of the 11-12 rows are gathered first; when EVERY ratio of
    // the warp's stages is small (the usual case once the iterate is close) the logs are a 4-term series instead of 12
    // log1pf calls -- a warp-uniform choice, so no divergence.
    for (int k0 = 0; k0 < N; k0 += 32) {
      const int k = k0 + lane;
      const bool act = k < N;
      T rts[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) rts[i] = T(0);
      if (act) {
      int nr = 0;
      auto lrow = [&](T num, T den) { rts[nr++] = num * m_rcp(den); };
      const T u0 = rc(k, R_U), u1 = rc(k, R_U + 1);
      const T du0 = al * rc(k, R_DU), du1 = al * rc(k, R_DU + 1);
      const T nu0 = u0 + du0, nu1 = u1 + du1;
      T xad[5], xbd[5], xba[5], dxb[5], x1d[5], x1a[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        xad[j] = sx(k, S_XTT + j);
        xbd[j] = sx(k + 1, S_XTT + j);
        xba[j] = xbd[j] + sx(k + 1, S_XR + j);
        dxb[j] = al * rc(k, R_DX + j);
        x1d[j] = sx(k + 1, S_XT + j);
        x1a[j] = x1d[j] + sx(k + 1, S_XR + j);
      }
      Trig ta; ta.sn = sx(k, S_TRT); ta.cs = sx(k, S_TRT + 1); ta.tn = sx(k, S_TRT + 2);
      T d[5];
      if (reshoot) {
#pragma unroll
        for (int j = 0; j < 5; ++j) d[j] = (T(1) - al) * rc(k, R_D + j);
      } else {
        defect(k, xad, xbd, xad[3] + sx(k, S_XR + 3), ta, nu0, nu1, d);
      }
#pragma unroll
      for (int j = 0; j < 5; ++j) c1 += m_abs(d[j]);
      {
        const T t0 = P.R[0] * du0 * (T(2) * u0 + du0), t1 = P.R[1] * du1 * (T(2) * u1 + du1);
        dphi += t0 + t1; nz += m_abs(t0) + m_abs(t1);
      }
      if (k + 1 <= N - 1) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const T t0 = P.Q[j] * dxb[j] * (T(2) * x1d[j] + dxb[j]);
          dphi += t0; nz += m_abs(t0);
        }
      }
      lrow(du0, m_slack(u0 - P.dd_min));
      lrow(-du0, m_slack(P.dd_max - u0));
      const T ahi = (k == 0) ? st.a0_hi : P.a_max;
      lrow(-du1, m_slack(ahi - u1));
      lrow(dxb[2], m_slack(x1a[2] - P.de_min));
      lrow(-dxb[2], m_slack(P.de_max - x1a[2]));
      lrow(dxb[3], m_slack(x1a[3] - P.v_min));
      lrow(-dxb[3], m_slack(P.v_max - x1a[3]));
      // obstacle rows: the slack moves with its own Newton step ds (from the OLD linearisation)
      const T sn0 = sx(k + 1, S_TR), cs0 = sx(k + 1, S_TR + 1);
      const T snb = sx(k + 1, S_TRT), csb = sx(k + 1, S_TRT + 1);
      if (rc(k, R_FAR) == T(0)) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        T h, gx, gy, gp; obst(j, x1a[0], x1a[1], sn0, cs0, h, gx, gy, gp);
        const T s = rc(k, R_S + j);
        const T r = (h - P.r_sum) - s;
        const T ds = al * (gx * rc(k, R_DX) + gy * rc(k, R_DX + 1) + gp * rc(k, R_DX + 4) + r);
        rts[7 + j] = ds * m_rcp(s);
        T hb, g1, g2, g3; obst(j, xba[0], xba[1], snb, csb, hb, g1, g2, g3);
        c1 += m_resid((hb - P.r_sum) - (s + ds), hb);
      }
      }
      if (k == 0) rts[10] = du1 * m_rcp(m_slack(u1 - st.a0_lo));          // stage-0 friction box, lower side
      }   // act
      if (sizeof(T) == 4) {
        // float: ONE logarithm per stage -- the sum of the twelve log(1 + x_i) is the log of the product of the (1 + x_i)
        // (one FFMA per row; 0.01^12 .. 100^12 stays inside the float range).  The product carries ~12 roundings (7e-7
        // absolute in the log); these terms enter the merit times mu, next to cost terms whose own rounding is orders of
        // magnitude larger, and the allowance below covers it.  log(prod): 4-term series in d = prod - 1 below 0.02
        // (relative error < d^4/5 = 3e-8), else one MUFU logarithm.  (12 log1pf calls were ~400 instructions of the
        // per-iteration instruction stream, 12 MUFU logarithms with rounding correction still ~180.)
        T prod = T(1), xs = T(0);
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          ok = ok && (rts[i] > T(-1));
          const T x = m_max(rts[i], T(-0.999999));
          prod += prod * x;
          xs += m_abs(x);
        }
        const T d = prod - T(1);
        const T ser = d * (T(1) + d * (T(-0.5) + d * (T(1) / T(3) - T(0.25) * d)));
        const T l = (m_abs(d) < T(0.02)) ? ser : m_fastlog(prod);
        lg += l; lga += xs + T(2);
      } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
          ok = ok && (rts[i] > T(-1));
          const T l = m_log1p(m_max(rts[i], T(-0.999999)));
          lg += l; lga += m_abs(l);
        }
      }
    }
    dphi -= mu * lg;
    nz += mu * lga;
    ok = w.all(ok);
    dphi = w.sum(dphi); c1 = w.sum(c1); nz = w.sum(nz);
    w.sync();          // the next trial / the restored forward sweep overwrites the trial states and the step these lanes just read
    return ok;
  }

  // ---------------------------------------------------------------- phase G: commit the step + complementarity statistics
  MPC_HD void commit(ProbState<T>& st, T al, T ad, T& avg, T& cmax, T& smin_ob) const {
    const int N = P.N;
    const T mu = st.mu, ikap = m_rcp(P.kappa_sigma);
    T sum = T(0); cmax = T(0);
    T smin = T(1e30);                              // smallest new slack of a live obstacle row (how stiff the barrier system is)
    for (int k = lane; k < N; k += 32) {
      const T u0 = rc(k, R_U), u1 = rc(k, R_U + 1);
      const T du0 = rc(k, R_DU), du1 = rc(k, R_DU + 1);
      T dx[5], x1a[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) { dx[j] = rc(k, R_DX + j); x1a[j] = xa(k + 1, j); }
      const T nu0 = u0 + al * du0, nu1 = u1 + al * du1;
      T xn[5];
#pragma unroll
      for (int j = 0; j < 5; ++j) xn[j] = sx(k + 1, S_XTT + j);
      const T nde = xn[2] + sx(k + 1, S_XR + 2), nvv = xn[3] + sx(k + 1, S_XR + 3);
      // multiplier update at the OLD point, complementarity product at the NEW point
      auto upd = [&](int slot, T s, T ds, T snew) {
        const T nu = rc(k, R_V + slot);
        const T dnu = (mu - nu * s - nu * ds) * m_rcp(s);
        T nn = m_max(nu + ad * dnu, T(1e-30));
        // lower safeguard on the multiplier, nu >= mu / (kappa_sigma s)  (IPOPT's kappa_sigma, Waechter & Biegler 2006 eq. 16)
        const T cen = mu * m_rcp(snew);
        nn = m_max(nn, cen * ikap);
        rc(k, R_V + slot) = nn;
        const T c = snew * nn; sum += c; cmax = m_max(cmax, c);
      };
      upd(V_DD_LO, m_slack(u0 - P.dd_min), du0, m_slack(nu0 - P.dd_min));
      upd(V_DD_HI, m_slack(P.dd_max - u0), -du0, m_slack(P.dd_max - nu0));
      const T ahi = (k == 0) ? st.a0_hi : P.a_max;
      upd(V_A_HI, m_slack(ahi - u1), -du1, m_slack(ahi - nu1));
      if (k == 0) upd(V_A_LO, m_slack(u1 - st.a0_lo), du1, m_slack(nu1 - st.a0_lo));
      upd(V_DE_LO, m_slack(x1a[2] - P.de_min), dx[2], m_slack(nde - P.de_min));
      upd(V_DE_HI, m_slack(P.de_max - x1a[2]), -dx[2], m_slack(P.de_max - nde));
      upd(V_V_LO, m_slack(x1a[3] - P.v_min), dx[3], m_slack(nvv - P.v_min));
      upd(V_V_HI, m_slack(P.v_max - x1a[3]), -dx[3], m_slack(P.v_max - nvv));
      const T sn1 = sx(k + 1, S_TR), cs1 = sx(k + 1, S_TR + 1);
      if (rc(k, R_FAR) == T(0)) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        T h, gx, gy, gp; obst(j, x1a[0], x1a[1], sn1, cs1, h, gx, gy, gp);
        const T s = rc(k, R_S + j);
        const T r = (h - P.r_sum) - s;
        const T ds = gx * dx[0] + gy * dx[1] + gp * dx[4] + r;
        const T snew = m_slack(s + al * ds);
        upd(V_OB0 + j, s, ds, snew);
        rc(k, R_S + j) = snew;
        smin = m_min(smin, snew);
      }
      } else {
        sum += T(3) * mu;                          // screened rows sit on the central path: s*nu = mu
      }
      rc(k, R_U) = nu0; rc(k, R_U + 1) = nu1;
#pragma unroll
      for (int j = 0; j < 5; ++j) sx(k + 1, S_XT + j) = xn[j];
      sx(k + 1, S_TR) = sx(k + 1, S_TRT); sx(k + 1, S_TR + 1) = sx(k + 1, S_TRT + 1); sx(k + 1, S_TR + 2) = sx(k + 1, S_TRT + 2);
    }
    sum = w.sum(sum);
    cmax = w.max_nonneg(m_max(cmax, T(0)));
    smin_ob = w.min_nonneg(m_max(smin, T(0)));
    avg = sum * irows;
    w.sync();
  }

  // ---------------------------------------------------------------- one SQP / interior-point iteration (uniform control flow)
  // remaining error after this step, from the contraction r = step / previous step of the last two full steps: the next
  // ratio is taken as r^1.5 (Newton-like iterations square it, Gauss-Newton's linear rate keeps it: 1.5 sits between), the tail as
  // a geometric series in that ratio
  MPC_HD bool rate_ok(T step, T prev) const {
    const T r = step * m_rcp(prev);
    const T q = rate_pow() ? r * m_sqrt_fast(r) : r;
    return step * q <= P.tol_step * (T(1) - q);
  }
  MPC_HD static bool rate_pow() {
#if !defined(__CUDACC__) && defined(MPC_DIAG)
    static const bool v = getenv("MPC_RATEPOW") ? atoi(getenv("MPC_RATEPOW")) != 0 : true;
    return v;
#else
    return true;
#endif
  }
  MPC_HD static bool tune_extrap() {
#if !defined(__CUDACC__) && defined(MPC_DIAG)
    static const bool v = getenv("MPC_EXTRAP") ? atoi(getenv("MPC_EXTRAP")) != 0 : true;
    return v;
#else
    return true;
#endif
  }
  MPC_HD static bool tune_pred() {
#if !defined(__CUDACC__) && defined(MPC_DIAG)
    static const bool v = getenv("MPC_PRED") ? atoi(getenv("MPC_PRED")) != 0 : true;
    return v;
#else
    return true;
#endif
  }
  MPC_HD void iterate(ProbState<T>& st) const {
    if (st.done) return;
    linearize(st);
    backward();
    forward_sweep();
    FwdOut f = forward_stats(st);
    if (!m_finite(f.step_inf) || !m_finite(f.dphi)) { st.status = ST_NAN; st.done = 1; return; }
    // penalty parameter of the l1 merit (left alone once the infeasibility is at its rounding-noise floor: dividing by
    // a noise-level c1 would blow rho up and hand the line search to the noise)
    const T epsm = m_eps(T(0));
    if (f.c1 > T(8) * epsm * f.mag) {
      const T need = f.dphi * m_rcp(T(0.5) * f.c1);
      if (need > st.rho) st.rho = need * T(1.5) + T(1);
    }
    const T slope = f.dphi - st.rho * f.c1;
    const T cfloor = T(8) * epsm * f.mag;                       // rounding-noise floor of the l1 infeasibility
    const bool trust = (f.c1 <= cfloor) && (f.step_inf <= P.trust_step);
    T al = f.a_p;
    // Final-phase extrapolation.  On a weakly active row (slack and multiplier both -> 0) Newton's step on s nu = mu HALVES per
    // iteration -- a double root; twice the Newton step restores fast convergence.  Detected from two scalars (this full step is
    // 0.4 .. 0.6 of the previous full step, both at the final barrier parameter); the doubled length is the first line-search
    // trial (Armijo as usual, the plain Newton step is the second), capped by the uncapped fraction-to-the-boundary length.
    T ad = f.a_d;
    bool extrap = false;
    if (tune_extrap() && st.pstep > T(0) && st.mu <= P.mu_min * T(1.0001) && f.a_p >= T(1) && f.a_d >= T(1) &&
        f.step_inf > T(0.4) * st.pstep && f.step_inf < T(0.6) * st.pstep && f.lim > T(1.5)) {
      al = m_min(T(2), f.lim); extrap = true;
    }
    bool accepted = false;
    for (int t = 0; t < P.ls_max; ++t) {
      T dphi, c1, nz;
      trial_points(al);
      bool ok = trial_merit(st, al, false, dphi, c1, nz);
      T dm = dphi + st.rho * (c1 - f.c1);
      T noise = T(8) * epsm * (nz + st.rho * f.mag);
#if defined(MPC_DEBUG_LS) && !defined(__CUDACC__)
      if (lane == 0) printf("      ls t=%d al=%.3e ok=%d dphi=%.4e c1=%.4e (c1_0 %.4e) dm=%.4e thresh=%.4e noise=%.3e nz=%.3e mag=%.3e slope=%.3e\n", t, (double)al, (int)ok,
             (double)dphi, (double)c1, (double)f.c1, (double)dm, (double)(T(1e-4) * al * slope + noise), (double)noise, (double)nz, (double)f.mag, (double)slope);
#endif
      if (ok && m_finite(dm) && dm <= T(1e-4) * al * slope + noise) { accepted = true; break; }
      // Newton-trust acceptance: feasible to working precision and a short step.  There the l1 merit cannot judge the
      // step any more -- defect rounding noise times the (large, unknown) dynamics multipliers exceeds the predicted
      // change -- while the undamped Newton iteration converges quadratically; stalls are caught by the stall exit.
      if (trust && ok && m_finite(dm) && c1 <= T(2) * cfloor) { accepted = true; break; }
      if (t == 0 && ok && m_finite(dm)) {
        trial_points_reshoot(al);
        ok = trial_merit(st, al, true, dphi, c1, nz);
        dm = dphi + st.rho * (c1 - f.c1);
        noise = T(8) * epsm * (nz + st.rho * f.mag);
        if (ok && m_finite(dm) && dm <= T(1e-4) * al * slope + noise) { accepted = true; st.nsoc++; break; }
        // the re-simulated step replaced DX: restore the Newton step for the shorter trials
        forward_sweep();
      }
      al *= T(0.5);
    }
    if (!accepted) {
      st.nfail++;
      if (st.nfail >= 3) { st.status = ST_NOPROGRESS; st.done = 1; return; }
      trial_points(al);        // the (short) step is taken anyway: trial point for commit
    } else {
      st.nfail = 0;
    }
    T avg, cmax, smin_ob;
    if (extrap && al > T(1)) ad = al;                       // the multipliers take the same extrapolated length
    commit(st, al, ad, avg, cmax, smin_ob);
    st.d_al = al; st.d_ap = f.a_p; st.d_ad = f.a_d; st.d_c1 = f.c1; st.d_dphi = f.dphi; st.d_blk = f.blk;
    st.iters++;
    st.kkt = f.step_inf;
    // rate estimate for the next iteration: this full Newton step's norm if it was taken as it is near the final barrier
    // parameter (an extrapolated or damped step says nothing about the rate)
    const T pstep_new = (al == T(1) && st.mu <= T(2) * P.mu_min) ? f.step_inf : T(0);
    if (st.mu <= P.mu_min * T(1.0001) && f.c1 <= P.tol_feas) {
      if (al >= T(0.5) && al * f.step_inf <= P.tol_step) { st.status = ST_OPTIMAL; st.done = 1; return; }
      // rate-based exit: two full steps in a row near the final barrier parameter shrinking at rate r = step / previous step <
      // 0.5 leave a remaining error of about step r / (1 - r); below tol_step the next iteration would only confirm it
      // (not in float32 with a stiff live obstacle row: there the step sizes are at the rounding-noise floor of the KKT solve and
      // say nothing about the distance to the optimum -- measured: 8 of the 4096 config-3 instances ended 1.1e-3 .. 2.7e-3 off)
      if (tune_pred() && al >= T(1) && st.pstep > T(0) && f.step_inf < T(0.5) * st.pstep && (sizeof(T) == 8 || smin_ob >= P.stiff_slack) &&
          rate_ok(f.step_inf, st.pstep)) { st.status = ST_OPTIMAL; st.done = 1; return; }
      // "acceptable" exit (IPOPT's acceptable_tol / acceptable_iter idea): with strongly active rows the barrier
      // weights reach 1/mu_min and the Newton step has a rounding-noise floor above tol_step; a run of steps at that
      // floor is convergence, not progress
      // In float32 a run of steps at the acceptable level WITH an active obstacle row (slack below stiff_slack: barrier weight
      // nu / s of 1e4 and more on a rank-1 term) is the rounding-noise floor of the KKT solve, measured up to 1.4e-3 from the
      // optimum -- outside the stated 1e-3: such an exit is flagged ST_STALLED so that the float64 refinement pass polishes it.
      if (al * f.step_inf <= P.acc_factor * P.tol_step) {
        if (++st.nacc >= P.acc_iters) { st.status = (sizeof(T) == 4 && smin_ob < P.stiff_slack) ? ST_STALLED : ST_OPTIMAL; st.done = 1; return; }
      }
      else st.nacc = 0;
      // stall exit: the step no longer halves -- the iterate sits at the rounding-noise floor of the arithmetic
      // (fp32 with barrier weights ~1/mu_min on an active obstacle row).  Usable, flagged ST_STALLED.
      const T sz = al * f.step_inf;
      if (sz < T(0.5) * st.best) { st.best = sz; st.nstall = 0; }
      else if (++st.nstall >= P.stall_iters) { st.status = ST_STALLED; st.done = 1; return; }
    }
    st.pstep = pstep_new;
    // Barrier warm-up: while no step of length >= 0.5 has been taken, a step blocked hard by the fraction-to-the-boundary
    // rule (alpha < mu_up_alpha) means the barrier is invisible next to the cost gradient -- the iteration would crawl
    // one bound per step.  Raise mu (and the multipliers with it, keeping s*nu on the central path) instead.
    if (!st.centered) {
      if (al >= T(0.5)) st.centered = 1;
      else if (al < P.mu_up_alpha && st.mu * P.mu_up_factor <= P.mu_max) {
        st.mu *= P.mu_up_factor;
        for (int k = lane; k < P.N; k += 32) {
#pragma unroll
          for (int j = 0; j < NV; ++j) rc(k, R_V + j) *= P.mu_up_factor;
        }
        w.sync();
        return;
      }
    }
    if (al >= P.mu_min_alpha) {
      // monotone barrier update, linear (mu_factor) far out and superlinear (avg^1.5, as IPOPT's theta_mu) close in
      // (a full primal and dual step means the linearisation was trusted all the way: reduce faster)
      const T fac = (al >= T(1) && f.a_d >= T(1)) ? P.mu_factor_full : P.mu_factor;
      const T mu_new = m_max(P.mu_min, m_min(st.mu, m_min(fac * avg, avg * m_sqrt_fast(avg))));
      if (mu_new < st.mu) st.rho = m_max(T(1), st.rho * T(0.5));
      st.mu = mu_new;
    }
  }
};

}  // namespace mpcb200
