// forces_core.cuh -- warp-cooperative solver core for the reference's FORCESPRO FORMULATION of the MPC problem
// (/root/reference/MPC_Planner/optimizer.py:86-246): ONE WARP PER EGO INSTANCE, like warp_core.cuh, but for
//
//   stage variable  z_k = [deltaDot, aLong | xPos, yPos, delta, v, psi], k = 0 .. N-1             optimizer.py:93, 204-205
//   dynamics        x_{k+1} = one RK4 step of the kinematic single-track model, k = 0 .. N-2      :90-98  (x_0 = xinit, :224)
//   inequalities    lb <= z_k <= ub (symmetric acceleration bounds, :108-109), the friction circle
//                   aLong^2 + (v psiDot)^2 <= a_max^2 at EVERY stage and the 3 x 3 SQUARED circle distances >= r_sum^2  :110-155
//   objective       stage least squares with per-stage reference (path point, ramped desired speed, path heading) and a
//                   terminal stage with its own weights and no input terms                          :163-195, 288-317
//
// The reference hands this NLP to the closed-source FORCESPRO SQP core (one QP per call, BFGS).  Here it is solved to
// convergence by the same primal-dual interior-point / Gauss-Newton SQP iteration as warp_core.cuh.  What changes against the
// CasADi-formulation core, and why it is a second core and not a flag:
//   * RK4 makes B_k = d x_{k+1} / d u_k state dependent and dense (8 non-zeros), A_k has 13;
//   * the friction row couples aLong_k with (delta_k, v_k): a cross term S_k = d2 / du dx in the stage Hessian;
//   so the backward sweep is the GENERAL Riccati recursion
//       G = R + B'PB,  H = S + B'PA,  P_k = Q + A'PA - H' G^-1 H,   K = -G^-1 H
//   on the 5x6 augmented block [P | p] with LANE = MATRIX ENTRY: round 1 gathers row i of P (5 shuffles) for M = [P|p][A d; 0 1]
//   and P B; round 2 gathers column j and column i of M and P B (20 shuffles); G (2x2) is formed redundantly by every lane.
//   The closed-loop matrices A + BK, d + B kff are stored by the same lanes, so the forward sweep is one 6-term dot product per
//   lane (lanes 0..4: dx_{k+1}; lanes 5, 6: du_k) and 5 broadcast shuffles per stage.
//   Stage-parallel phases (linearisation incl. the RK4 Jacobian by the chain rule of forces_model.cuh, step limits, merit,
//   commit) run LANE = STAGE as before.
// Stage record k holds ALL terms of z_k (the FORCESPRO stage), so P_k = value function of x_k including its own stage terms.
//
// Two statements differ from the literal model without changing its solution set (same as oracle/forces_nlp.py): the vacuous
// lower bound 0 <= aLong^2 + (v psiDot)^2 is not a row, and the last stage's inputs -- which appear in no cost term and no
// dynamics -- get the regular input weights so that the minimum-norm member u_{N-1} = 0 of the optimal set is returned.
//
// New code (the reference has no solver of its own).  Compiles for the device and for the host emulator (tests/host_sim).
#pragma once
#include "mpc_types.cuh"
#include "warp_ctx.cuh"
#include "forces_model.cuh"
#include "warp_core.cuh"      // SlabRef, FetchTag

namespace mpcb200 {

// multiplier slots of a stage
enum : int { FV_DD_LO = 0, FV_DD_HI, FV_A_LO, FV_A_HI, FV_DE_LO, FV_DE_HI, FV_V_LO, FV_V_HI, FV_FR, FV_OB0, FNV = 18 };
// stage record k = 0..N-1 (words of T)
enum : int {
  FR_U = 0,      // 2  inputs u_k
  FR_V = 2,      // 18 multipliers (FV_*: 8 bound rows, friction, 9 circle pairs)
  FR_S = 20,     // 10 slacks of the nonlinear rows: friction, 9 circle pairs (ego circle outer, obstacle circle inner)
  FR_A = 30,     // 25 A_k = d x_{k+1} / d x_k, row major
  FR_B = 55,     // 10 B_k = d x_{k+1} / d u_k, [t][c] at 2t + c
  FR_D = 65,     // 5  defect d_k = c(z_k) - x_{k+1}
  FR_ZERO = 70,  // 1  constant 0
  FR_H = 71,     // 9  Hessian of the x_k terms: h00 h01 h04 h11 h14 h44 h22 h23 h33
  FR_GX = 80,    // 5  gradient of the x_k terms (barrier gradient at the current mu)
  FR_RU = 85,    // 2  input Hessian diagonal
  FR_RG = 87,    // 2  input gradient
  FR_SX = 89,    // 2  cross terms d2 / d aLong d(delta, v) of the friction row
  FR_KK = 91,    // 12 gains K0[0..4] k0 K1[0..4] k1
  FR_ACL = 103,  // 30 closed loop [A + BK | d + B kff], entry (i, j) at 6i + j.  (Aliasing it onto A_k | B_k -- 11 instead of 9
                 //    problems per SM -- was measured: no gain at large batches, and the Gauss-Newton fallback after a rejected
                 //    curvature sweep then has to linearise again: USA_Lanker -22 %.  Kept separate.)
  FR_DX = 133,   // 5  step dx_k
  FR_DU = 138,   // 2  step du_k
  FR_FAR = 140,  // 1  circle rows of x_k screened out this iteration
  FR_LC = 141,   // 5  multiplier-weighted gradient of the x_k terms (adjoint recursion of the dynamics-curvature term)
  FREC_PLAIN = 147,  // record stride without the road-boundary rows (odd strides keep lane = stage accesses bank-conflict free)
  FR_BV = 146,   // 6  multipliers of the road-boundary rows: left boundary x (centre, front, rear circle), right boundary x (...)
  FR_BS = 152,   // 6  their slacks
  FREC_RB = 159  // record stride with them
};
MPC_HD constexpr int forces_rec_stride(bool rb) { return rb ? FREC_RB : FREC_PLAIN; }
// state record k = 0..N-1
enum : int {
  FS_XR = 0,     // 5 reference of the stage cost [path_x, path_y, 0, v_des, psi_ref] (positions relative to xinit)
  FS_XT = 5,     // 5 deviation state x_k - ref_k
  FS_XTT = 10,   // 5 trial deviation state
  FS_OB = 15,    // 6 obstacle circle centres of this stage (relative to xinit)
  FS_CP = 21,    // 2 reference position increment ref_k - ref_{k+1}, formed in float64
  FS_TR = 23,    // 3 sin(psi_k) cos(psi_k) tan(delta_k) at the current iterate
  FST = 27
};
enum : int { FH00 = 0, FH01, FH04, FH11, FH14, FH44, FH22, FH23, FH33 };

struct FLayout {
  int N, o_state, o_rec, o_misc, words;
  // o_misc: 2 words, the position of xinit (the slab's positions are relative to it; the road boundaries are absolute)
  MPC_HD explicit FLayout(int N_, bool rb = false) : N(N_) {
    o_state = 0; o_rec = FST * N; o_misc = o_rec + forces_rec_stride(rb) * N; words = (o_misc + 2 + 3) & ~3;
  }
};

// road boundaries (optional rows): vertex lists [n][2] in absolute coordinates, device memory
template <typename T>
struct RoadBounds {
  const T* left; const T* right;
  int nl, nr;
  T r_min;          // radius_ego (optimizer.py:115)
};

template <typename T>
struct FParams {
  ParamsT<T> P;     // horizon, bounds, stage weights, solver options (a_max: symmetric acceleration bound AND friction radius)
  T Pt[5];          // terminal weights
};

struct FLaneTab {
  int i, j, own;
  int hidx;         // x_k Hessian / gradient term of entry (i, j)
  int c[5];         // [A d][t][j]
  int s1[5];        // source lanes of P[i][t]
  int ai[5];        // A[l][i]
  int so0, so1;     // S[c][j] (gradient r_c on the affine column)
  int si1;          // S[1][i]
  int aij, bi0, bi1;
  int exw;          // dynamics-curvature addition: 0 none, 1 (4,4), 2 (3,4), 3 (2,2), 4 (2,3)
  MPC_HD explicit FLaneTab(int lane) {
    const int l = lane < 30 ? lane : 0;
    i = l / 6; j = l % 6;
    own = (j == 5) ? 1 : 0;
    hidx = FR_ZERO;
    if (j == 5) hidx = FR_GX + i;
    else if (i <= j) {
      if (i == 0 && j == 0) hidx = FR_H + FH00;
      if (i == 0 && j == 1) hidx = FR_H + FH01;
      if (i == 0 && j == 4) hidx = FR_H + FH04;
      if (i == 1 && j == 1) hidx = FR_H + FH11;
      if (i == 1 && j == 4) hidx = FR_H + FH14;
      if (i == 4 && j == 4) hidx = FR_H + FH44;
      if (i == 2 && j == 2) hidx = FR_H + FH22;
      if (i == 2 && j == 3) hidx = FR_H + FH23;
      if (i == 3 && j == 3) hidx = FR_H + FH33;
    }
#pragma unroll
    for (int t = 0; t < 5; ++t) {
      c[t] = (j == 5) ? (FR_D + t) : (FR_A + 5 * t + j);
      s1[t] = (i <= t) ? (6 * i + t) : (6 * t + i);
      ai[t] = FR_A + 5 * t + i;
    }
    so0 = (j == 5) ? FR_RG : FR_ZERO;
    so1 = (j == 2) ? FR_SX : (j == 3) ? (FR_SX + 1) : (j == 5) ? (FR_RG + 1) : FR_ZERO;
    si1 = (i == 2) ? FR_SX : (i == 3) ? (FR_SX + 1) : FR_ZERO;
    aij = (j == 5) ? (FR_D + i) : (FR_A + 5 * i + j);
    bi0 = FR_B + 2 * i; bi1 = FR_B + 2 * i + 1;
    exw = 0;
    if (j < 5 && i <= j) { if (i == 4 && j == 4) exw = 1; if (i == 3 && j == 4) exw = 2; if (i == 2 && j == 2) exw = 3; if (i == 2 && j == 3) exw = 4; }
  }
};

// RB: the six road-boundary rows per stage are part of the problem (compile-time: a kernel without them carries none of their code)
template <typename T, bool RB = false>
struct ForcesSolver {
  const ParamsT<T>& P;
  const T* Pt;
  const RoadBounds<T> rb;
  static constexpr int FREC = RB ? FREC_RB : FREC_PLAIN;
  const FLayout L;
  const SlabRef<T> sl;
  const WarpCtx& w;
  const int lane;
  const FLaneTab tb;
  ForcesConsts<T> FC;
  const T a2max, r2, irows;

  MPC_HD ForcesSolver(const FParams<T>& fp, const SlabRef<T>& slab, const WarpCtx& w_, const RoadBounds<T>& rb_ = RoadBounds<T>{nullptr, nullptr, 0, 0, T(0)})
      : P(fp.P), Pt(fp.Pt), rb(rb_), L(fp.P.N, RB), sl(slab), w(w_), lane(w_.lane()), tb(w_.lane()), a2max(fp.P.a_max * fp.P.a_max),
        r2(fp.P.r_sum * fp.P.r_sum), irows(T(1) / T(18 * fp.P.N - 13 + (RB ? 6 * (fp.P.N - 1) : 0))) {
    FC.dt = P.dt; FC.l_wb = P.l_wb; FC.l_fric = P.l_fric; FC.ego_off = P.ego_off;
    for (int q = 0; q < 5; ++q) { FC.Q[q] = P.Q[q]; FC.Pt[q] = fp.Pt[q]; }
    FC.R[0] = P.R[0]; FC.R[1] = P.R[1];
  }

  MPC_HD T& sx(int k, int f) const { return sl[L.o_state + FST * k + f]; }
  MPC_HD T& rc(int k, int f) const { return sl[L.o_rec + FREC * k + f]; }
  MPC_HD static T x0_tol() { return sizeof(T) == 4 ? T(2e-5) : T(1e-9); }

  // ------------------------------------------------------------------ nonlinear rows
  // friction circle as c = a_max^2 - aLong^2 - q^2 >= 0 with q = v^2 tan(delta) / wheelbase (optimizer.py:131, 145)
  struct Fric { T c, ga, gde, gv, qd, qv; };
  MPC_HD Fric fric(T a, T de, T v) const {
    const T tn = m_tan(de);
    const T il = T(1) / P.l_fric;
    const T q = v * v * tn * il;
    Fric f;
    f.qd = v * v * (T(1) + tn * tn) * il; f.qv = T(2) * v * tn * il;
    f.c = a2max - a * a - q * q;
    f.ga = T(-2) * a; f.gde = T(-2) * q * f.qd; f.gv = T(-2) * q * f.qv;
    return f;
  }
  // circle pair (ego circle e in {centre, front, rear}, obstacle circle o).  The model states the row on the SQUARED distance,
  // D^2 >= r_sum^2 (optimizer.py:110, 146-154); the solver works on the equivalent row c = D - r_sum >= 0 -- same feasible set,
  // same KKT points (the gradients are parallel, D > 0), but nearly linear in the position, so the l1 merit accepts long steps
  // (on D^2 the linearisation error of a 10 m step is 100 m^2 and the line search crawls: measured, 1/64 steps).
  MPC_HD void pair(int e, int o, int k, T px, T py, T sn, T cs, T& c, T& gx, T& gy, T& gp) const {
    const T off = (e == 0) ? T(0) : (e == 1 ? P.ego_off : -P.ego_off);
    const T dx = px + off * cs - sx(k, FS_OB + 2 * o), dy = py + off * sn - sx(k, FS_OB + 2 * o + 1);
    const T d2 = m_max(dx * dx + dy * dy, T(1e-24));
    const T ih = m_rsqrt(d2);
    c = d2 * ih - P.r_sum;
    gx = dx * ih; gy = dy * ih;
    gp = off * (gy * cs - gx * sn);
  }
  // Row screening (same idea as warp_core.cuh): a circle row whose barrier curvature (nu/s) |grad c|^2 ~ mu / (D - r)^2 is
  // below 1/screen_inv_curv is skipped for the iteration.  Conservative on the centre distance D_c.
  MPC_HD bool is_far(int k, T px, T py, T mu) const {
    if (!(P.screen_inv_curv > T(0))) return false;
    const T ox = sx(k, FS_OB), oy = sx(k, FS_OB + 1);
    const T s1x = sx(k, FS_OB + 2) - ox, s1y = sx(k, FS_OB + 3) - oy, s2x = sx(k, FS_OB + 4) - ox, s2y = sx(k, FS_OB + 5) - oy;
    const T spread = m_sqrt_fast(m_max(s1x * s1x + s1y * s1y, s2x * s2x + s2y * s2y));
    const T reach = P.r_sum + P.ego_off + spread + m_sqrt_fast(mu * P.screen_inv_curv);
    const T dx = px - ox, dy = py - oy;
    return dx * dx + dy * dy > reach * reach;
  }

  // road-boundary row (side 0 left / 1 right, ego circle e): c = distance of the circle centre to the CLOSEST VERTEX of the
  // boundary polyline - radius_ego >= 0 (find_closest_distance_with_road_boundary, optimizer.py:18-30: ca.mmin over the vertex
  // distances; rows :156-161, bound :115).  Closest vertex by a coarse pass over every 8th vertex and a fine pass around the
  // best one (road boundaries are smooth polylines with ~1 m spacing); the gradient is that of the distance to this vertex.
  MPC_HD void brow(int side, int e, T px, T py, T sn, T cs, T& c, T& gx, T& gy, T& gp) const {
    const T off = (e == 0) ? T(0) : (e == 1 ? P.ego_off : -P.ego_off);
    const T ax = px + off * cs + sl[L.o_misc], ay = py + off * sn + sl[L.o_misc + 1];
    const T* b = side ? rb.right : rb.left;
    const int n = side ? rb.nr : rb.nl;
    int best = 0; T bd = T(3e38);
    for (int i = 0; i < n; i += 8) {
      const T dx = ax - b[2 * i], dy = ay - b[2 * i + 1], d2 = dx * dx + dy * dy;
      if (d2 < bd) { bd = d2; best = i; }
    }
    const int lo = best - 7 > 0 ? best - 7 : 0, hi = best + 7 < n - 1 ? best + 7 : n - 1;
    for (int i = lo; i <= hi; ++i) {
      const T dx = ax - b[2 * i], dy = ay - b[2 * i + 1], d2 = dx * dx + dy * dy;
      if (d2 < bd) { bd = d2; best = i; }
    }
    const T dx = ax - b[2 * best], dy = ay - b[2 * best + 1];
    const T d2 = m_max(dx * dx + dy * dy, T(1e-24));
    const T ih = m_rsqrt(d2);
    c = d2 * ih - rb.r_min;
    gx = dx * ih; gy = dy * ih;
    gp = off * (gy * cs - gx * sn);
  }

  // ------------------------------------------------------------------ problem I/O (float64 arrays of ONE problem)
  // xinit [5]; params [N][10] (FORCESNLPsolver_params.all_parameters, stage major: path_x, path_y, v_des, psi_ref, 3 circle
  // centres); Zin [N][7] warm start (FORCESNLPsolver_params.x0; may be null: xinit tiled, zero inputs).  lane = stage.
  MPC_HD void load(const double* xinit, const double* par, const double* Zin) const {
    const int N = P.N;
    const double ox = xinit[0], oy = xinit[1];
    if (lane == 0) { sl[L.o_misc] = (T)ox; sl[L.o_misc + 1] = (T)oy; }
    for (int k = lane; k < N; k += 32) {
      const double* p = par + 10 * k;
      const double ref[5] = {p[0], p[1], 0.0, p[2], p[3]};
      sx(k, FS_XR) = (T)(ref[0] - ox); sx(k, FS_XR + 1) = (T)(ref[1] - oy); sx(k, FS_XR + 2) = T(0);
      sx(k, FS_XR + 3) = (T)ref[3]; sx(k, FS_XR + 4) = (T)ref[4];
      for (int q = 0; q < 3; ++q) { sx(k, FS_OB + 2 * q) = (T)(p[4 + 2 * q] - ox); sx(k, FS_OB + 2 * q + 1) = (T)(p[5 + 2 * q] - oy); }
      const double* xs = (k == 0 || !Zin) ? xinit : (Zin + 7 * k + 2);
      for (int q = 0; q < 5; ++q) sx(k, FS_XT + q) = (T)(xs[q] - ref[q]);
      if (k + 1 < N) { sx(k, FS_CP) = (T)(p[0] - p[10]); sx(k, FS_CP + 1) = (T)(p[1] - p[11]); }
      else { sx(k, FS_CP) = T(0); sx(k, FS_CP + 1) = T(0); }
      rc(k, FR_U) = Zin ? (T)Zin[7 * k] : T(0); rc(k, FR_U + 1) = Zin ? (T)Zin[7 * k + 1] : T(0);
      rc(k, FR_ZERO) = T(0);
    }
    w.sync();
  }
  MPC_HD void store(const double* xinit, const double* par, double* Z) const {
    const int N = P.N;
    w.sync();
    for (int k = lane; k < N; k += 32) {
      const double* p = par + 10 * k;
      const double ref[5] = {p[0], p[1], 0.0, p[2], p[3]};
      Z[7 * k] = (double)rc(k, FR_U); Z[7 * k + 1] = (double)rc(k, FR_U + 1);
      for (int q = 0; q < 5; ++q) Z[7 * k + 2 + q] = (k == 0) ? xinit[q] : ((double)sx(k, FS_XT + q) + ref[q]);
    }
    w.sync();
  }

  // ------------------------------------------------------------------ initialisation (lane = stage)
  MPC_HD void init(ProbState<T>& st) const {
    const int N = P.N;
    st.mu = P.mu0; st.rho = T(1); st.status = ST_MAXIT; st.iters = 0; st.done = 0; st.nfail = 0; st.nsoc = 0; st.nacc = 0; st.centered = 0; st.nstall = 0; st.best = T(1e30); st.kkt = T(0);
    st.a0_lo = st.a0_hi = T(0);
    st.d_al = st.d_ap = st.d_ad = st.d_c1 = st.d_dphi = T(0); st.d_blk = 0;
    const T kp = P.bound_push, mu = st.mu;
    bool bad = false;
    for (int k = lane; k < N; k += 32) {
      // inputs strictly inside their box
      const T pdd = m_min(kp, kp * (P.dd_max - P.dd_min));
      const T dd = m_min(m_max(rc(k, FR_U), P.dd_min + pdd), P.dd_max - pdd);
      const T pa = kp * m_max(T(1), P.a_max);
      T a = m_min(m_max(rc(k, FR_U + 1), -P.a_max + pa), P.a_max - pa);
      T xa[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) xa[q] = sx(k, FS_XT + q) + sx(k, FS_XR + q);
      if (k >= 1) {
        const T pde = m_min(kp * m_max(T(1), m_abs(P.de_max)), kp * (P.de_max - P.de_min));
        const T pv = m_min(kp * m_max(T(1), m_abs(P.v_max)), kp * (P.v_max - P.v_min));
        xa[2] = m_min(m_max(xa[2], P.de_min + pde), P.de_max - pde);
        xa[3] = m_min(m_max(xa[3], P.v_min + pv), P.v_max - pv);
        sx(k, FS_XT + 2) = xa[2] - sx(k, FS_XR + 2);
        sx(k, FS_XT + 3) = xa[3] - sx(k, FS_XR + 3);
      }
      // friction row: pull aLong inside the circle of this stage's (delta, v) if the guess is outside
      Fric f = fric(a, xa[2], xa[3]);
      if (k == 0 && !(f.c + a * a > x0_tol())) bad = true;                  // lateral acceleration of xinit alone exceeds a_max
      if (!(f.c > kp * a2max)) {
        const T room = m_max(f.c + a * a - T(2) * kp * a2max, T(0));
        const T am = m_sqrt(room);
        a = m_min(m_max(a, -am), am);
        f = fric(a, xa[2], xa[3]);
      }
      rc(k, FR_U) = dd; rc(k, FR_U + 1) = a;
      const T sf = m_max(f.c, kp * a2max);
      rc(k, FR_S) = sf; rc(k, FR_V + FV_FR) = mu * m_rcp(sf);
      rc(k, FR_V + FV_DD_LO) = mu * m_rcp(dd - P.dd_min); rc(k, FR_V + FV_DD_HI) = mu * m_rcp(P.dd_max - dd);
      rc(k, FR_V + FV_A_LO) = mu * m_rcp(a + P.a_max); rc(k, FR_V + FV_A_HI) = mu * m_rcp(P.a_max - a);
      if (k >= 1) {
        rc(k, FR_V + FV_DE_LO) = mu * m_rcp(xa[2] - P.de_min); rc(k, FR_V + FV_DE_HI) = mu * m_rcp(P.de_max - xa[2]);
        rc(k, FR_V + FV_V_LO) = mu * m_rcp(xa[3] - P.v_min); rc(k, FR_V + FV_V_HI) = mu * m_rcp(P.v_max - xa[3]);
      } else {
        rc(k, FR_V + FV_DE_LO) = T(0); rc(k, FR_V + FV_DE_HI) = T(0); rc(k, FR_V + FV_V_LO) = T(0); rc(k, FR_V + FV_V_HI) = T(0);
        if (xa[2] < P.de_min - x0_tol() || xa[2] > P.de_max + x0_tol() || xa[3] < P.v_min - x0_tol() || xa[3] > P.v_max + x0_tol()) bad = true;
      }
      T sn, cs; m_sincos(xa[4], &sn, &cs);
#pragma unroll
      for (int e = 0; e < 3; ++e) {
#pragma unroll
        for (int o = 0; o < 3; ++o) {
          T c, gx, gy, gp; pair(e, o, k, xa[0], xa[1], sn, cs, c, gx, gy, gp);
          if (k == 0 && c < -x0_tol()) bad = true;
          const T s = m_max(c, kp * m_max(T(1), P.r_sum));
          rc(k, FR_S + 1 + 3 * e + o) = s;
          rc(k, FR_V + FV_OB0 + 3 * e + o) = mu * m_rcp(s);
        }
      }
      if (RB) {
        for (int q = 0; q < 6; ++q) {
          T c, gx, gy, gp; brow(q / 3, q % 3, xa[0], xa[1], sn, cs, c, gx, gy, gp);
          if (k == 0 && c < -x0_tol()) bad = true;
          const T s = m_max(c, kp * m_max(T(1), rb.r_min));
          rc(k, FR_BS + q) = s;
          rc(k, FR_BV + q) = mu * m_rcp(s);
        }
      }
#pragma unroll
      for (int q = 0; q < 5; ++q) { rc(k, FR_DX + q) = T(0); sx(k, FS_XTT + q) = sx(k, FS_XT + q); }
      rc(k, FR_DU) = T(0); rc(k, FR_DU + 1) = T(0);
      rc(k, FR_FAR) = T(0);
    }
    if (w.any(bad)) { st.status = ST_INFEASIBLE_X0; st.done = 1; }
    w.sync();
  }

  // ------------------------------------------------------------------ phase A: stage KKT blocks (lane = stage)
  MPC_HD void linearize(const ProbState<T>& st) const {
    const int N = P.N;
    const T mu = st.mu;
    for (int k = lane; k < N; k += 32) {
      T xd[5], xa[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) { xd[q] = sx(k, FS_XT + q); xa[q] = xd[q] + sx(k, FS_XR + q); }
      const T u0 = rc(k, FR_U), u1 = rc(k, FR_U + 1);
      const bool term = (k == N - 1);
      if (!term) {
        // RK4 step and its 5x7 Jacobian; the position rows of the state do not enter f, so they are passed as 0 and the
        // "next state" of those rows IS the increment
        const T z[7] = {u0, u1, T(0), T(0), xa[2], xa[3], xa[4]};
        T xn[5], inc[5], dc[5][7];
        forces_dynamics(FC, z, xn, dc, inc);
#pragma unroll
        for (int r = 0; r < 5; ++r) {
#pragma unroll
          for (int c = 0; c < 5; ++c) rc(k, FR_A + 5 * r + c) = dc[r][2 + c];
          rc(k, FR_B + 2 * r) = dc[r][0]; rc(k, FR_B + 2 * r + 1) = dc[r][1];
        }
        rc(k, FR_D + 0) = (xd[0] - sx(k + 1, FS_XT + 0)) + inc[0] + sx(k, FS_CP);
        rc(k, FR_D + 1) = (xd[1] - sx(k + 1, FS_XT + 1)) + inc[1] + sx(k, FS_CP + 1);
#pragma unroll
        for (int q = 2; q < 5; ++q) rc(k, FR_D + q) = (xd[q] - sx(k + 1, FS_XT + q)) + inc[q] + (sx(k, FS_XR + q) - sx(k + 1, FS_XR + q));
      } else {
#pragma unroll
        for (int q = 0; q < 25; ++q) rc(k, FR_A + q) = T(0);
#pragma unroll
        for (int q = 0; q < 10; ++q) rc(k, FR_B + q) = T(0);
#pragma unroll
        for (int q = 0; q < 5; ++q) rc(k, FR_D + q) = T(0);
      }
      // x_k terms (k >= 1; x_0 is fixed)
      T hd[5] = {T(0), T(0), T(0), T(0), T(0)}, g[5] = {T(0), T(0), T(0), T(0), T(0)}, lc[5] = {T(0), T(0), T(0), T(0), T(0)};
      T h01 = T(0), h04 = T(0), h14 = T(0), h23 = T(0);
      bool far = true;
      T sn, cs; m_sincos(xa[4], &sn, &cs);
      sx(k, FS_TR) = sn; sx(k, FS_TR + 1) = cs; sx(k, FS_TR + 2) = m_tan(xa[2]);
      if (k >= 1) {
        const T* wq = term ? Pt : P.Q;
#pragma unroll
        for (int q = 0; q < 5; ++q) { hd[q] = T(2) * wq[q]; g[q] = T(2) * wq[q] * xd[q]; lc[q] = g[q]; }
        {
          const T ilo = m_rcp(m_slack(xa[2] - P.de_min)), ihi = m_rcp(m_slack(P.de_max - xa[2]));
          hd[2] += rc(k, FR_V + FV_DE_LO) * ilo + rc(k, FR_V + FV_DE_HI) * ihi;
          g[2] += mu * (ihi - ilo);
          lc[2] += rc(k, FR_V + FV_DE_HI) - rc(k, FR_V + FV_DE_LO);
        }
        {
          const T ilo = m_rcp(m_slack(xa[3] - P.v_min)), ihi = m_rcp(m_slack(P.v_max - xa[3]));
          hd[3] += rc(k, FR_V + FV_V_LO) * ilo + rc(k, FR_V + FV_V_HI) * ihi;
          g[3] += mu * (ihi - ilo);
          lc[3] += rc(k, FR_V + FV_V_HI) - rc(k, FR_V + FV_V_LO);
        }
        far = is_far(k, xa[0], xa[1], mu);
        if (!far) {
#pragma unroll
          for (int e = 0; e < 3; ++e) {
#pragma unroll
            for (int o = 0; o < 3; ++o) {
              T c, gx, gy, gp; pair(e, o, k, xa[0], xa[1], sn, cs, c, gx, gy, gp);
              T s = rc(k, FR_S + 1 + 3 * e + o);
              if (c > s) { s = c; rc(k, FR_S + 1 + 3 * e + o) = s; }           // slack reset (Nocedal & Wright 19.3)
              const T nu = rc(k, FR_V + FV_OB0 + 3 * e + o);
              const T is = m_rcp(s);
              const T wgt = nu * is, r = c - s;
              const T cg = -(mu * is - wgt * r);
              hd[0] += wgt * gx * gx; h01 += wgt * gx * gy; h04 += wgt * gx * gp;
              hd[1] += wgt * gy * gy; h14 += wgt * gy * gp; hd[4] += wgt * gp * gp;
              g[0] += cg * gx; g[1] += cg * gy; g[4] += cg * gp;
              lc[0] -= nu * gx; lc[1] -= nu * gy; lc[4] -= nu * gp;
            }
          }
        }
        if (RB) {
          for (int q = 0; q < 6; ++q) {
            T c, gx, gy, gp; brow(q / 3, q % 3, xa[0], xa[1], sn, cs, c, gx, gy, gp);
            T s = rc(k, FR_BS + q);
            if (c > s) { s = c; rc(k, FR_BS + q) = s; }
            const T nu = rc(k, FR_BV + q);
            const T is = m_rcp(s);
            const T wgt = nu * is, r = c - s;
            const T cg = -(mu * is - wgt * r);
            hd[0] += wgt * gx * gx; h01 += wgt * gx * gy; h04 += wgt * gx * gp;
            hd[1] += wgt * gy * gy; h14 += wgt * gy * gp; hd[4] += wgt * gp * gp;
            g[0] += cg * gx; g[1] += cg * gy; g[4] += cg * gp;
            lc[0] -= nu * gx; lc[1] -= nu * gy; lc[4] -= nu * gp;
          }
        }
      }
      rc(k, FR_FAR) = far ? T(1) : T(0);
      // input terms
      T Ru0, Ru1, ru0, ru1;
      {
        const T ilo = m_rcp(m_slack(u0 - P.dd_min)), ihi = m_rcp(m_slack(P.dd_max - u0));
        Ru0 = T(2) * P.R[0] + rc(k, FR_V + FV_DD_LO) * ilo + rc(k, FR_V + FV_DD_HI) * ihi;
        ru0 = T(2) * P.R[0] * u0 + mu * (ihi - ilo);
        const T jlo = m_rcp(m_slack(u1 + P.a_max)), jhi = m_rcp(m_slack(P.a_max - u1));
        Ru1 = T(2) * P.R[1] + rc(k, FR_V + FV_A_LO) * jlo + rc(k, FR_V + FV_A_HI) * jhi;
        ru1 = T(2) * P.R[1] * u1 + mu * (jhi - jlo);
      }
      // friction row: Gauss-Newton term (nu/s) grad grad' plus the positive semi-definite part of -nu * hess c
      // (2 nu on aLong, 2 nu grad q grad q' on (delta, v)); the indefinite part -2 nu q hess q is dropped
      T sx0 = T(0), sx1 = T(0);
      {
        const Fric f = fric(u1, xa[2], xa[3]);
        T s = rc(k, FR_S);
        if (f.c > s) { s = f.c; rc(k, FR_S) = s; }
        const T nu = rc(k, FR_V + FV_FR);
        const T is = m_rcp(s);
        const T wgt = nu * is, r = f.c - s;
        const T cg = -(mu * is - wgt * r);
        Ru1 += wgt * f.ga * f.ga + T(2) * nu;
        ru1 += cg * f.ga;
        if (k >= 1) {
          hd[2] += wgt * f.gde * f.gde + T(2) * nu * f.qd * f.qd;
          h23 += wgt * f.gde * f.gv + T(2) * nu * f.qd * f.qv;
          hd[3] += wgt * f.gv * f.gv + T(2) * nu * f.qv * f.qv;
          g[2] += cg * f.gde; g[3] += cg * f.gv;
          lc[2] -= nu * f.gde; lc[3] -= nu * f.gv;
          sx0 = wgt * f.ga * f.gde; sx1 = wgt * f.ga * f.gv;
        }
      }
#pragma unroll
      for (int q = 0; q < 5; ++q) rc(k, FR_LC + q) = lc[q];
      rc(k, FR_H + FH00) = hd[0]; rc(k, FR_H + FH01) = h01; rc(k, FR_H + FH04) = h04; rc(k, FR_H + FH11) = hd[1];
      rc(k, FR_H + FH14) = h14; rc(k, FR_H + FH44) = hd[4]; rc(k, FR_H + FH22) = hd[2]; rc(k, FR_H + FH23) = h23; rc(k, FR_H + FH33) = hd[3];
#pragma unroll
      for (int q = 0; q < 5; ++q) rc(k, FR_GX + q) = g[q];
      rc(k, FR_RU) = Ru0; rc(k, FR_RU + 1) = Ru1; rc(k, FR_RG) = ru0; rc(k, FR_RG + 1) = ru1;
      rc(k, FR_SX) = sx0; rc(k, FR_SX + 1) = sx1;
    }
    w.sync();
  }

  // ------------------------------------------------------------------ phase C: backward Riccati sweep (lane = entry of [P | p])
  // EX: the curvature of the dynamics rows, sum_i lam_{k+1,i} hess c_i(z_k), is added to the state block with the adjoint
  // multipliers lam_k = lc_k + A_k' lam_{k+1} (every lane keeps the 5-vector) and the CONTINUOUS-TIME second derivatives times dt
  // (the RK4 step's own second derivative differs by O(dt^2); a Hessian approximation only has to be good enough for the
  // iteration to contract -- measured: with position weights of 200 against a heading weight of 1 (USA_Lanker) the pure
  // Gauss-Newton iteration DIVERGES at rate 1.1 near the solution).  Returns false (uniformly) if an input block is not positive
  // definite; the caller then repeats the sweep without the term.
  template <bool EX>
  MPC_HD bool backward_t() const {
    const int N = P.N;
    T Pij = T(0);
    T lam0 = T(0), lam1 = T(0), lam2 = T(0), lam3 = T(0), lam4 = T(0);       // lam_{k+1} (uniform)
    const T ownf = (T)tb.own;
    bool pd = true;                                   // every input block so far positive definite (uniform; no early return:
                                                      // a branch out of the loop would make every shuffle a guarded collective)
    for (int k = N - 1; k >= 0; --k) {
      const int o = L.o_rec + FREC * k;
      T b0[5], b1[5], q[5];
      // round 1: row i of P
#pragma unroll
      for (int t = 0; t < 5; ++t) q[t] = w.shfl(Pij, tb.s1[t]);
#pragma unroll
      for (int t = 0; t < 5; ++t) { b0[t] = sl[o + FR_B + 2 * t]; b1[t] = sl[o + FR_B + 2 * t + 1]; }
      T M = ownf * Pij, MB0 = T(0), MB1 = T(0);
#pragma unroll
      for (int t = 0; t < 5; ++t) { M += sl[o + tb.c[t]] * q[t]; MB0 += b0[t] * q[t]; MB1 += b1[t] * q[t]; }
      // round 2: column j and column i of M, the two columns of P B
      T F0j = sl[o + tb.so0], F1j = sl[o + tb.so1], F0i = T(0), F1i = sl[o + tb.si1];
      T G00 = sl[o + FR_RU], G01 = T(0), G11 = sl[o + FR_RU + 1];
      T Fxx = sl[o + tb.hidx];
      T a_i[5];
#pragma unroll
      for (int l = 0; l < 5; ++l) a_i[l] = sl[o + tb.ai[l]];
#pragma unroll
      for (int l = 0; l < 5; ++l) {
        const T mlj = w.shfl(M, 6 * l + tb.j), mli = w.shfl(M, 6 * l + tb.i);
        const T mb0 = w.shfl(MB0, 6 * l), mb1 = w.shfl(MB1, 6 * l);
        F0j += b0[l] * mlj; F1j += b1[l] * mlj;
        F0i += b0[l] * mli; F1i += b1[l] * mli;
        G00 += b0[l] * mb0; G01 += b0[l] * mb1; G11 += b1[l] * mb1;
        Fxx += a_i[l] * mlj;
      }
      const T det = G00 * G11 - G01 * G01;
      if (EX) pd = pd && (G00 > T(0)) && (det > T(1e-8) * G00 * G11);
      const T cdet = m_rcp(det);
      const T J00 = -cdet * G11, J01 = cdet * G01, J11 = -cdet * G00;
      const T T0 = J00 * F0j + J01 * F1j, T1 = J01 * F0j + J11 * F1j;      // gains [K | kff] column j
      if (EX) {
        const T dt = P.dt, il = T(1) / P.l_wb;
        const T v = sx(k, FS_XT + 3) + sx(k, FS_XR + 3);
        const T sn = sx(k, FS_TR), cs = sx(k, FS_TR + 1), tn = sx(k, FS_TR + 2);
        const T sec2 = T(1) + tn * tn;
        T ex = T(0);
        if (tb.exw == 1) ex = -dt * v * (lam0 * cs + lam1 * sn);
        if (tb.exw == 2) ex = dt * (lam1 * cs - lam0 * sn);
        if (tb.exw == 3) ex = dt * lam4 * T(2) * v * il * sec2 * tn;
        if (tb.exw == 4) ex = dt * lam4 * sec2 * il;
        Fxx += ex;
        // lam_k: lane (i, .) forms component i from the A[l][i] it already holds, five broadcasts hand the vector to every lane
        const T li = sl[o + FR_LC + tb.i] + (a_i[0] * lam0 + a_i[1] * lam1) + (a_i[2] * lam2 + a_i[3] * lam3) + a_i[4] * lam4;
        lam0 = w.shfl(li, 0); lam1 = w.shfl(li, 6); lam2 = w.shfl(li, 12); lam3 = w.shfl(li, 18); lam4 = w.shfl(li, 24);
      }
      const T Pn = Fxx + F0i * T0 + F1i * T1;
      if (lane < 6) { sl[o + FR_KK + tb.j] = T0; sl[o + FR_KK + 6 + tb.j] = T1; }
      if (lane < 30) sl[o + FR_ACL + lane] = sl[o + tb.aij] + sl[o + tb.bi0] * T0 + sl[o + tb.bi1] * T1;
      Pij = Pn;
    }
    w.sync();
    return pd;
  }
  MPC_HD void backward() const {
    if (P.hessian == HESS_EXACT) { if (backward_t<true>()) return; }
    backward_t<false>();
  }

  // ------------------------------------------------------------------ phase D: forward sweep
  // lanes 0..4: dx_{k+1} = (A + BK) dx_k + (d + B kff); lanes 5, 6: du_k = K dx_k + kff; 5 broadcast shuffles per stage
  MPC_HD void forward_sweep() const {
    const int N = P.N;
    const int r = lane < 7 ? lane : 0;
    const int fo = (r < 5) ? (FR_ACL + 6 * r) : (FR_KK + 6 * (r - 5));
    T dx0 = T(0), dx1 = T(0), dx2 = T(0), dx3 = T(0), dx4 = T(0);
    for (int k = 0; k < N; ++k) {
      const int o = L.o_rec + FREC * k + fo;
      const T nx = (sl[o + 5] + sl[o] * dx0) + (sl[o + 1] * dx1 + sl[o + 2] * dx2) + (sl[o + 3] * dx3 + sl[o + 4] * dx4);
      if (lane == 5 || lane == 6) rc(k, FR_DU + lane - 5) = nx;
      if (lane < 5 && k + 1 < N) rc(k + 1, FR_DX + lane) = nx;
      dx0 = w.shfl(nx, 0); dx1 = w.shfl(nx, 1); dx2 = w.shfl(nx, 2); dx3 = w.shfl(nx, 3); dx4 = w.shfl(nx, 4);
    }
    w.sync();
  }

  // ------------------------------------------------------------------ phase E: step-length limits, merit slope (lane = stage)
  struct FwdOut { T a_p, a_d, dphi, c1, step_inf, mag; int blk, cur; };
  MPC_HD void row_limits(T s, T nu, T ds, T mu, FwdOut& o) const {
    const T is = m_rcp(s), inu = m_rcp(nu);
    const T t = ds * is;
    const T q = (mu * is) * inu - T(1) - t;
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
    const T mu = st.mu;
    const T tau = m_max(P.tau_min, T(1) - mu);
    FwdOut o; o.a_p = T(0); o.a_d = T(0); o.dphi = T(0); o.c1 = T(0); o.step_inf = T(0); o.mag = T(0); o.blk = -1; o.cur = 0;
    for (int k = lane; k < N; k += 32) {
      o.cur = 32 * k;
      T xd[5], xa[5], dx[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) { xd[q] = sx(k, FS_XT + q); xa[q] = xd[q] + sx(k, FS_XR + q); dx[q] = rc(k, FR_DX + q); }
      const T u0 = rc(k, FR_U), u1 = rc(k, FR_U + 1), du0 = rc(k, FR_DU), du1 = rc(k, FR_DU + 1);
      const bool term = (k == N - 1);
      if (!term) {
#pragma unroll
        for (int q = 0; q < 5; ++q) o.c1 += m_abs(rc(k, FR_D + q));
      }
#pragma unroll
      for (int q = 0; q < 5; ++q) o.mag += T(2) * m_abs(xd[q]);
      o.mag += P.dt * (T(2) * m_abs(xa[3]) + m_abs(u0) + m_abs(u1)) + m_abs(sx(k, FS_CP)) + m_abs(sx(k, FS_CP + 1));
      o.dphi += T(2) * P.R[0] * u0 * du0 + T(2) * P.R[1] * u1 * du1;
      row_limits(m_slack(u0 - P.dd_min), rc(k, FR_V + FV_DD_LO), du0, mu, o);
      row_limits(m_slack(P.dd_max - u0), rc(k, FR_V + FV_DD_HI), -du0, mu, o);
      row_limits(m_slack(u1 + P.a_max), rc(k, FR_V + FV_A_LO), du1, mu, o);
      row_limits(m_slack(P.a_max - u1), rc(k, FR_V + FV_A_HI), -du1, mu, o);
      const Fric f = fric(u1, xa[2], xa[3]);
      T dsf = f.ga * du1 + (f.c - rc(k, FR_S));
      if (k >= 1) {
        const T* wq = term ? Pt : P.Q;
#pragma unroll
        for (int q = 0; q < 5; ++q) o.dphi += T(2) * wq[q] * xd[q] * dx[q];
        row_limits(m_slack(xa[2] - P.de_min), rc(k, FR_V + FV_DE_LO), dx[2], mu, o);
        row_limits(m_slack(P.de_max - xa[2]), rc(k, FR_V + FV_DE_HI), -dx[2], mu, o);
        row_limits(m_slack(xa[3] - P.v_min), rc(k, FR_V + FV_V_LO), dx[3], mu, o);
        row_limits(m_slack(P.v_max - xa[3]), rc(k, FR_V + FV_V_HI), -dx[3], mu, o);
        dsf += f.gde * dx[2] + f.gv * dx[3];
        if (rc(k, FR_FAR) == T(0)) {
          T sn, cs; m_sincos(xa[4], &sn, &cs);
#pragma unroll
          for (int e = 0; e < 3; ++e) {
#pragma unroll
            for (int ob = 0; ob < 3; ++ob) {
              T c, gx, gy, gp; pair(e, ob, k, xa[0], xa[1], sn, cs, c, gx, gy, gp);
              const T s = rc(k, FR_S + 1 + 3 * e + ob);
              const T r = c - s;
              o.c1 += m_resid(r, c + P.r_sum);
              row_limits(s, rc(k, FR_V + FV_OB0 + 3 * e + ob), gx * dx[0] + gy * dx[1] + gp * dx[4] + r, mu, o);
            }
          }
        }
        if (RB) {
          T sn, cs; m_sincos(xa[4], &sn, &cs);
          for (int q = 0; q < 6; ++q) {
            T c, gx, gy, gp; brow(q / 3, q % 3, xa[0], xa[1], sn, cs, c, gx, gy, gp);
            const T s = rc(k, FR_BS + q);
            const T r = c - s;
            o.c1 += m_resid(r, c + rb.r_min);
            row_limits(s, rc(k, FR_BV + q), gx * dx[0] + gy * dx[1] + gp * dx[4] + r, mu, o);
          }
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) o.step_inf = m_max(o.step_inf, m_abs(dx[q]));
      }
      o.c1 += m_resid(f.c - rc(k, FR_S), a2max);
      row_limits(rc(k, FR_S), rc(k, FR_V + FV_FR), dsf, mu, o);
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
    o.a_p = (o.a_p > tau) ? tau * m_rcp(o.a_p) : T(1);
    o.a_d = (o.a_d > tau) ? tau * m_rcp(o.a_d) : T(1);
    o.step_inf = w.max_nonneg(fin ? o.step_inf : T(0));
    o.dphi = w.sum(o.dphi); o.c1 = w.sum(o.c1); o.mag = w.sum(o.mag);
    if (!allfin) o.step_inf = T(NAN);
    return o;
  }

  // ------------------------------------------------------------------ phase F: merit difference phi(alpha) - phi(0) (lane = stage)
  MPC_HD void trial_points(T al) const {
    const int N = P.N;
    for (int k = lane; k < N; k += 32) {
#pragma unroll
      for (int q = 0; q < 5; ++q) sx(k, FS_XTT + q) = sx(k, FS_XT + q) + al * rc(k, FR_DX + q);
    }
    w.sync();
  }
  // (A second-order correction -- re-simulating the trial states from the trial inputs as warp_core.cuh does -- was built and
  // measured: never accepted on the lane-following batches, no help on the crawling collision-avoidance instances (their l1
  // infeasibility is in the circle-row slacks, not in the defects), +20 registers and -9 % throughput.  Not in the build.)
  MPC_HD bool trial_merit(const ProbState<T>& st, T al, T& dphi, T& c1, T& nz) const {
    const int N = P.N;
    const T mu = st.mu;
    dphi = T(0); c1 = T(0); nz = T(0);
    bool ok = true;
    T lg = T(0), lga = T(0);
    for (int k0 = 0; k0 < N; k0 += 32) {
      const int k = k0 + lane;
      T rts[18];
#pragma unroll
      for (int q = 0; q < 18; ++q) rts[q] = T(0);
      T prodb = T(1), xsb = T(0);            // road-boundary rows: their (1 + x) product / log sum is accumulated in the loop
      if (k < N) {
        const bool term = (k == N - 1);
        const T u0 = rc(k, FR_U), u1 = rc(k, FR_U + 1);
        const T du0 = al * rc(k, FR_DU), du1 = al * rc(k, FR_DU + 1);
        const T nu0 = u0 + du0, nu1 = u1 + du1;
        T xd[5], xa[5], xbd[5], xba[5], dxa[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          xd[q] = sx(k, FS_XT + q); xa[q] = xd[q] + sx(k, FS_XR + q);
          xbd[q] = sx(k, FS_XTT + q); xba[q] = xbd[q] + sx(k, FS_XR + q);
          dxa[q] = al * rc(k, FR_DX + q);
        }
        if (!term) {
          const T z[7] = {nu0, nu1, T(0), T(0), xba[2], xba[3], xba[4]};
          T inc[5];
          forces_rk4_increment(FC, z, inc);
          c1 += m_abs((xbd[0] - sx(k + 1, FS_XTT + 0)) + inc[0] + sx(k, FS_CP));
          c1 += m_abs((xbd[1] - sx(k + 1, FS_XTT + 1)) + inc[1] + sx(k, FS_CP + 1));
#pragma unroll
          for (int q = 2; q < 5; ++q) c1 += m_abs((xbd[q] - sx(k + 1, FS_XTT + q)) + inc[q] + (sx(k, FS_XR + q) - sx(k + 1, FS_XR + q)));
        }
        {
          const T t0 = P.R[0] * du0 * (T(2) * u0 + du0), t1 = P.R[1] * du1 * (T(2) * u1 + du1);
          dphi += t0 + t1; nz += m_abs(t0) + m_abs(t1);
        }
        rts[0] = du0 * m_rcp(m_slack(u0 - P.dd_min)); rts[1] = -du0 * m_rcp(m_slack(P.dd_max - u0));
        rts[2] = du1 * m_rcp(m_slack(u1 + P.a_max)); rts[3] = -du1 * m_rcp(m_slack(P.a_max - u1));
        const Fric f = fric(u1, xa[2], xa[3]);
        const T sf = rc(k, FR_S);
        T dsf = f.ga * rc(k, FR_DU + 1) + (f.c - sf);
        if (k >= 1) {
          const T* wq = term ? Pt : P.Q;
#pragma unroll
          for (int q = 0; q < 5; ++q) {
            const T t0 = wq[q] * dxa[q] * (T(2) * xd[q] + dxa[q]);
            dphi += t0; nz += m_abs(t0);
          }
          rts[4] = dxa[2] * m_rcp(m_slack(xa[2] - P.de_min)); rts[5] = -dxa[2] * m_rcp(m_slack(P.de_max - xa[2]));
          rts[6] = dxa[3] * m_rcp(m_slack(xa[3] - P.v_min)); rts[7] = -dxa[3] * m_rcp(m_slack(P.v_max - xa[3]));
          dsf += f.gde * rc(k, FR_DX + 2) + f.gv * rc(k, FR_DX + 3);
          if (rc(k, FR_FAR) == T(0)) {
            T sn, cs, snb, csb; m_sincos(xa[4], &sn, &cs); m_sincos(xba[4], &snb, &csb);
#pragma unroll
            for (int e = 0; e < 3; ++e) {
#pragma unroll
              for (int ob = 0; ob < 3; ++ob) {
                T c, gx, gy, gp; pair(e, ob, k, xa[0], xa[1], sn, cs, c, gx, gy, gp);
                const T s = rc(k, FR_S + 1 + 3 * e + ob);
                const T ds = al * (gx * rc(k, FR_DX) + gy * rc(k, FR_DX + 1) + gp * rc(k, FR_DX + 4) + (c - s));
                rts[9 + 3 * e + ob] = ds * m_rcp(s);
                T cb, g1, g2, g3; pair(e, ob, k, xba[0], xba[1], snb, csb, cb, g1, g2, g3);
                c1 += m_resid(cb - (s + ds), cb + P.r_sum);
              }
            }
          }
          if (RB) {
            T sn, cs, snb, csb; m_sincos(xa[4], &sn, &cs); m_sincos(xba[4], &snb, &csb);
            for (int q = 0; q < 6; ++q) {
              T c, gx, gy, gp; brow(q / 3, q % 3, xa[0], xa[1], sn, cs, c, gx, gy, gp);
              const T s = rc(k, FR_BS + q);
              const T ds = al * (gx * rc(k, FR_DX) + gy * rc(k, FR_DX + 1) + gp * rc(k, FR_DX + 4) + (c - s));
              {
                const T x = ds * m_rcp(s);
                ok = ok && (x > T(-1));
                const T xc = m_max(x, T(-0.999999));
                if (sizeof(T) == 4) { prodb += prodb * xc; xsb += m_abs(xc); }
                else { const T l = m_log1p(xc); lg += l; lga += m_abs(l); }
              }
              T cb, g1, g2, g3; brow(q / 3, q % 3, xba[0], xba[1], snb, csb, cb, g1, g2, g3);
              c1 += m_resid(cb - (s + ds), cb + rb.r_min);
            }
          }
        }
        rts[8] = al * dsf * m_rcp(sf);
        const Fric fb = fric(nu1, xba[2], xba[3]);
        c1 += m_resid(fb.c - (sf + al * dsf), a2max);
      }
      // two logarithms per stage: log of the product of the (1 + x_i) of the 9 bound / friction rows and of the 9 circle rows
      if (sizeof(T) == 4) {
#pragma unroll
        for (int half = 0; half < (RB ? 3 : 2); ++half) {
          T prod = (half == 2) ? prodb : T(1), xs = (half == 2) ? xsb : T(0);
#pragma unroll
          for (int q = 0; q < (half == 2 ? 0 : 9); ++q) {
            const T x = rts[9 * half + q];
            ok = ok && (x > T(-1));
            const T xc = m_max(x, T(-0.999999));
            prod += prod * xc;
            xs += m_abs(xc);
          }
          const T d = prod - T(1);
          const T ser = d * (T(1) + d * (T(-0.5) + d * (T(1) / T(3) - T(0.25) * d)));
          lg += (m_abs(d) < T(0.02)) ? ser : m_fastlog(prod);
          lga += xs + T(2);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 18; ++q) {
          ok = ok && (rts[q] > T(-1));
          const T l = m_log1p(m_max(rts[q], T(-0.999999)));
          lg += l; lga += m_abs(l);
        }
      }
    }
    dphi -= mu * lg;
    nz += mu * lga;
    ok = w.all(ok);
    dphi = w.sum(dphi); c1 = w.sum(c1); nz = w.sum(nz);
    w.sync();
    return ok;
  }

  // ------------------------------------------------------------------ phase G: commit the step + complementarity statistics
  MPC_HD void commit(ProbState<T>& st, T al, T ad, T& avg, T& cmax, T& smin_nl) const {
    const int N = P.N;
    const T mu = st.mu, ikap = m_rcp(P.kappa_sigma);
    T sum = T(0); cmax = T(0);
    T smin = T(1e30);
    for (int k = lane; k < N; k += 32) {
      const T u0 = rc(k, FR_U), u1 = rc(k, FR_U + 1), du0 = rc(k, FR_DU), du1 = rc(k, FR_DU + 1);
      T xa[5], dx[5], xn[5];
#pragma unroll
      for (int q = 0; q < 5; ++q) { xa[q] = sx(k, FS_XT + q) + sx(k, FS_XR + q); dx[q] = rc(k, FR_DX + q); xn[q] = sx(k, FS_XTT + q); }
      const T nu0 = u0 + al * du0, nu1 = u1 + al * du1;
      auto upd = [&](int slot, T s, T ds, T snew) {
        const T nu = rc(k, FR_V + slot);
        const T dnu = (mu - nu * s - nu * ds) * m_rcp(s);
        T nn = m_max(nu + ad * dnu, T(1e-30));
        nn = m_max(nn, mu * m_rcp(snew) * ikap);
        rc(k, FR_V + slot) = nn;
        const T c = snew * nn; sum += c; cmax = m_max(cmax, c);
      };
      upd(FV_DD_LO, m_slack(u0 - P.dd_min), du0, m_slack(nu0 - P.dd_min));
      upd(FV_DD_HI, m_slack(P.dd_max - u0), -du0, m_slack(P.dd_max - nu0));
      upd(FV_A_LO, m_slack(u1 + P.a_max), du1, m_slack(nu1 + P.a_max));
      upd(FV_A_HI, m_slack(P.a_max - u1), -du1, m_slack(P.a_max - nu1));
      const Fric f = fric(u1, xa[2], xa[3]);
      const T sf = rc(k, FR_S);
      T dsf = f.ga * du1 + (f.c - sf);
      if (k >= 1) {
        const T nde = xn[2] + sx(k, FS_XR + 2), nvv = xn[3] + sx(k, FS_XR + 3);
        upd(FV_DE_LO, m_slack(xa[2] - P.de_min), dx[2], m_slack(nde - P.de_min));
        upd(FV_DE_HI, m_slack(P.de_max - xa[2]), -dx[2], m_slack(P.de_max - nde));
        upd(FV_V_LO, m_slack(xa[3] - P.v_min), dx[3], m_slack(nvv - P.v_min));
        upd(FV_V_HI, m_slack(P.v_max - xa[3]), -dx[3], m_slack(P.v_max - nvv));
        dsf += f.gde * dx[2] + f.gv * dx[3];
        if (rc(k, FR_FAR) == T(0)) {
          T sn, cs; m_sincos(xa[4], &sn, &cs);
#pragma unroll
          for (int e = 0; e < 3; ++e) {
#pragma unroll
            for (int ob = 0; ob < 3; ++ob) {
              T c, gx, gy, gp; pair(e, ob, k, xa[0], xa[1], sn, cs, c, gx, gy, gp);
              const T s = rc(k, FR_S + 1 + 3 * e + ob);
              const T ds = gx * dx[0] + gy * dx[1] + gp * dx[4] + (c - s);
              const T snew = m_slack(s + al * ds);
              upd(FV_OB0 + 3 * e + ob, s, ds, snew);
              rc(k, FR_S + 1 + 3 * e + ob) = snew;
              smin = m_min(smin, snew);
            }
          }
        } else {
          sum += T(9) * mu;
        }
        if (RB) {
          T sn, cs; m_sincos(xa[4], &sn, &cs);
          for (int q = 0; q < 6; ++q) {
            T c, gx, gy, gp; brow(q / 3, q % 3, xa[0], xa[1], sn, cs, c, gx, gy, gp);
            const T s = rc(k, FR_BS + q);
            const T ds = gx * dx[0] + gy * dx[1] + gp * dx[4] + (c - s);
            const T snew = m_slack(s + al * ds);
            // the boundary multipliers live outside FR_V: same update as upd()
            const T nu = rc(k, FR_BV + q);
            const T dnu = (mu - nu * s - nu * ds) * m_rcp(s);
            T nn = m_max(nu + ad * dnu, T(1e-30));
            nn = m_max(nn, mu * m_rcp(snew) * ikap);
            rc(k, FR_BV + q) = nn;
            const T cc = snew * nn; sum += cc; cmax = m_max(cmax, cc);
            rc(k, FR_BS + q) = snew;
            smin = m_min(smin, snew);
          }
        }
      }
      {
        const T snew = m_slack(sf + al * dsf);
        upd(FV_FR, sf, dsf, snew);
        rc(k, FR_S) = snew;
        smin = m_min(smin, snew * m_rcp(a2max));
      }
      rc(k, FR_U) = nu0; rc(k, FR_U + 1) = nu1;
      if (k >= 1) {
#pragma unroll
        for (int q = 0; q < 5; ++q) sx(k, FS_XT + q) = xn[q];
      }
    }
    sum = w.sum(sum);
    cmax = w.max_nonneg(m_max(cmax, T(0)));
    smin_nl = w.min_nonneg(m_max(smin, T(0)));
    avg = sum * irows;
    w.sync();
  }

  // ------------------------------------------------------------------ one SQP / interior-point iteration (uniform control flow)
  MPC_HD void iterate(ProbState<T>& st) const {
    if (st.done) return;
    linearize(st);
    backward();
    forward_sweep();
    FwdOut f = forward_stats(st);
    if (!m_finite(f.step_inf) || !m_finite(f.dphi)) { st.status = ST_NAN; st.done = 1; return; }
    const T epsm = m_eps(T(0));
    if (f.c1 > T(8) * epsm * f.mag) {
      const T need = f.dphi * m_rcp(T(0.5) * f.c1);
      if (need > st.rho) st.rho = need * T(1.5) + T(1);
      // the penalty parameter also comes DOWN again when the current step needs much less: one huge early step (a blocked
      // iterate far from the central path) otherwise leaves rho in the thousands, and every later step that bends the trajectory
      // around an obstacle is cut to 1/16 .. 1/32 by the l1 merit (measured: collision avoidance / road-boundary crawlers)
      else if (rho_decay() && need * T(1.5) + T(1) < T(0.5) * st.rho) st.rho = m_max(T(1), m_max(need * T(1.5) + T(1), T(0.5) * st.rho));
    }
    const T slope = f.dphi - st.rho * f.c1;
    const T cfloor = T(8) * epsm * f.mag;
    const bool trust = (f.c1 <= cfloor) && (f.step_inf <= P.trust_step);
    T al = f.a_p;
    bool accepted = false;
    T c1_new = f.c1;                                       // l1 infeasibility at the point the step leads to
    for (int t = 0; t < P.ls_max; ++t) {
      T dphi, c1, nz;
      trial_points(al);
      const bool ok = trial_merit(st, al, dphi, c1, nz);
      if (ok && m_finite(c1)) c1_new = c1;
      const T dm = dphi + st.rho * (c1 - f.c1);
      const T noise = T(8) * epsm * (nz + st.rho * f.mag);
      if (ok && m_finite(dm) && dm <= T(1e-4) * al * slope + noise) { accepted = true; break; }
      if (trust && ok && m_finite(dm) && c1 <= T(2) * cfloor) { accepted = true; break; }
      al *= T(0.5);
    }
    if (!accepted) {
      st.nfail++;
      if (st.nfail >= 3) { st.status = ST_NOPROGRESS; st.done = 1; return; }
      trial_points(al);
    } else {
      st.nfail = 0;
    }
    T avg, cmax, smin_nl;
    commit(st, al, f.a_d, avg, cmax, smin_nl);
    st.d_al = al; st.d_ap = f.a_p; st.d_ad = f.a_d; st.d_c1 = f.c1; st.d_dphi = f.dphi; st.d_blk = f.blk;
    st.iters++;
    st.kkt = f.step_inf;
    if (st.mu <= P.mu_min * T(1.0001) && f.c1 <= P.tol_feas) {
      if (al >= T(0.5) && al * f.step_inf <= P.tol_step) { st.status = ST_OPTIMAL; st.done = 1; return; }
      if (al * f.step_inf <= P.acc_factor * P.tol_step) {
        if (++st.nacc >= P.acc_iters) { st.status = (sizeof(T) == 4 && smin_nl < P.stiff_slack) ? ST_STALLED : ST_OPTIMAL; st.done = 1; return; }
      }
      else st.nacc = 0;
      const T sz = al * f.step_inf;
      if (sz < T(0.5) * st.best) { st.best = sz; st.nstall = 0; }
      else if (++st.nstall >= P.stall_iters) { st.status = ST_STALLED; st.done = 1; return; }
    }
    if (!st.centered) {
      if (al >= T(0.5)) st.centered = 1;
      else if (al < P.mu_up_alpha && st.mu * P.mu_up_factor <= P.mu_max) {
        st.mu *= P.mu_up_factor;
        for (int k = lane; k < P.N; k += 32) {
#pragma unroll
          for (int q = 0; q < FNV; ++q) rc(k, FR_V + q) *= P.mu_up_factor;
          if (RB) {
#pragma unroll
            for (int q = 0; q < 6; ++q) rc(k, FR_BV + q) *= P.mu_up_factor;
          }
        }
        w.sync();
        return;
      }
    }
    if (al >= P.mu_min_alpha) {
      const T fac = (al >= T(1) && f.a_d >= T(1)) ? P.mu_factor_full : P.mu_factor;
      T mu_new = m_max(P.mu_min, m_min(st.mu, m_min(fac * avg, avg * m_sqrt_fast(avg))));
      // the barrier parameter does not run ahead of feasibility (IPOPT lowers mu only once the barrier problem's error --
      // which includes the primal infeasibility -- is below kappa_eps mu): with mu ~ 1e-7 and defects still O(1) the next
      // steps are cut to 1e-3 by the fraction-to-the-boundary rule and the iteration jams on the bounds (measured: the tail
      // of the perturbed batches, 15 - 70 iterations instead of 6)
      mu_new = m_max(mu_new, m_min(st.mu, mu_feas() * c1_new));
      if (mu_new < st.mu) st.rho = m_max(T(1), st.rho * T(0.5));
      st.mu = mu_new;
    }
  }
  MPC_HD static bool rho_decay() {
#if !defined(__CUDACC__) && defined(MPC_DIAG)
    static const bool v = getenv("FORCES_RHO_DECAY") ? atoi(getenv("FORCES_RHO_DECAY")) != 0 : true;
    return v;
#else
    return true;
#endif
  }
  MPC_HD static T mu_feas() {
#if !defined(__CUDACC__) && defined(MPC_DIAG)
    static const T v = getenv("FORCES_MU_FEAS") ? (T)atof(getenv("FORCES_MU_FEAS")) : T(1e-3);
    return v;
#else
    return T(1e-3);
#endif
  }
};

}  // namespace mpcb200
