// forces_model.cuh -- stage functions of the reference's FORCESPRO formulation (/root/reference/MPC_Planner/optimizer.py:90-195)
// and their first derivatives, hand-derived (no AD): the linearisation a stage of an SQP iteration on that formulation needs.
//
//   stage variable   z = [deltaDot, aLong, xPos, yPos, delta, v, psi]                                    optimizer.py:93
//   stage parameters p = [path_x, path_y, v_des, psi_ref, obst_c_x, obst_c_y, obst_f_x, obst_f_y, obst_r_x, obst_r_y]   :108-111
//   dynamics   c(z)  = ONE explicit RK4 step of the kinematic single-track model over dt (forcespro.nlp.integrate) :97-98
//   inequalities h(z,p) (10): friction circle aLong^2 + (v * psiDot)^2 with psiDot = v tan(delta) / wheelbase, then the nine
//                  SQUARED distances ego circle (centre, front, rear) x obstacle circle (centre, front, rear)       :119-155
//   objective  f(z,p): least squares on path, steering angle, speed, heading and inputs :163-179; terminal: states only :181-195
//
// The reference ships these very functions as CasADi-generated C (test/FORCESNLPsolver/FORCESNLPsolver_model.c: dynamics_0,
// ddynamics_0, inequalities_0, dinequalities_0, objective_0/1, dobjective_0/1); oracle/Makefile compiles that file where it lies
// into oracle/_ref/libforces_model.so, and tests/test_forces_model_gpu.py compares this code with it value by value.
// The same source compiles for the host (tests/host_sim) like warp_core.cuh.
#pragma once
#include "mpc_types.cuh"

namespace mpcb200 {

template <typename T>
struct ForcesConsts {
  T dt;           // integrator step (0.1, optimizer.py:97)
  T l_wb;         // p.a + p.b of the dynamics (configuration.py:362-363)
  T l_fric;       // configuration.wheelbase used by the friction row (2.578, optimizer.py:119)
  T ego_off;      // ego circle-centre offset along the heading (0.75, configuration.py:80-91)
  T Q[5], R[2];   // stage weights: x, y, steering angle, velocity, heading | steering rate, acceleration
  T Pt[5];        // terminal weights (weight_*_terminate)
};

// f(x,u) of the kinematic single-track model (configuration.py:364-368)
template <typename T>
MPC_HD void ks_rhs(const T* x, const T* u, T l_wb, T* f) {
  T s, c; m_sincos(x[4], &s, &c);
  f[0] = x[3] * c; f[1] = x[3] * s; f[2] = u[0]; f[3] = u[1]; f[4] = x[3] / l_wb * m_tan(x[2]);
}
// J = df/dx at x: six structural non-zeros (rows 0,1 wrt v, psi; row 4 wrt delta, v); df/du is the constant selector e_delta, e_v
template <typename T>
struct KsJac { T f03, f04, f13, f14, f42, f43; };
template <typename T>
MPC_HD KsJac<T> ks_jac(const T* x, T l_wb) {
  T s, c; m_sincos(x[4], &s, &c);
  const T tn = m_tan(x[2]);
  KsJac<T> j;
  j.f03 = c; j.f04 = -x[3] * s; j.f13 = s; j.f14 = x[3] * c;
  j.f42 = x[3] / l_wb * (T(1) + tn * tn); j.f43 = tn / l_wb;
  return j;
}
// Y (5x7) = J * X (5x7), J with the sparsity of KsJac
template <typename T>
MPC_HD void ks_jac_mul(const KsJac<T>& j, const T X[5][7], T Y[5][7]) {
  for (int c = 0; c < 7; ++c) {
    Y[0][c] = j.f03 * X[3][c] + j.f04 * X[4][c];
    Y[1][c] = j.f13 * X[3][c] + j.f14 * X[4][c];
    Y[2][c] = T(0); Y[3][c] = T(0);
    Y[4][c] = j.f42 * X[2][c] + j.f43 * X[3][c];
  }
}

// One RK4 step x+ = c(z) and its Jacobian dc/dz (5x7, z = [u; x]) by the chain rule through the four stages.
// `inc` (optional): the increment x+ - x = h/6 (k1 + 2 k2 + 2 k3 + k4) itself, free of the cancellation in (x + inc) - x.
template <typename T>
MPC_HD void forces_dynamics(const ForcesConsts<T>& C, const T* z, T* xn, T dc[5][7], T* inc = nullptr) {
  const T h = C.dt;
  const T* u = z; const T* x = z + 2;
  T k1[5], k2[5], k3[5], k4[5], xs[5];
  T D[5][7], K[5][7], A[5][7];          // D = d(stage point)/dz, K = d(k_i)/dz, A = accumulated sum
  // stage 1
  ks_rhs(x, u, C.l_wb, k1);
  for (int r = 0; r < 5; ++r) for (int c = 0; c < 7; ++c) D[r][c] = (c == r + 2) ? T(1) : T(0);
  KsJac<T> j = ks_jac(x, C.l_wb);
  ks_jac_mul(j, D, K); K[2][0] += T(1); K[3][1] += T(1);
  for (int r = 0; r < 5; ++r) for (int c = 0; c < 7; ++c) A[r][c] = K[r][c];
  // stage 2
  for (int r = 0; r < 5; ++r) { xs[r] = x[r] + T(0.5) * h * k1[r]; for (int c = 0; c < 7; ++c) D[r][c] = ((c == r + 2) ? T(1) : T(0)) + T(0.5) * h * K[r][c]; }
  ks_rhs(xs, u, C.l_wb, k2);
  j = ks_jac(xs, C.l_wb);
  ks_jac_mul(j, D, K); K[2][0] += T(1); K[3][1] += T(1);
  for (int r = 0; r < 5; ++r) for (int c = 0; c < 7; ++c) A[r][c] += T(2) * K[r][c];
  // stage 3
  for (int r = 0; r < 5; ++r) { xs[r] = x[r] + T(0.5) * h * k2[r]; for (int c = 0; c < 7; ++c) D[r][c] = ((c == r + 2) ? T(1) : T(0)) + T(0.5) * h * K[r][c]; }
  ks_rhs(xs, u, C.l_wb, k3);
  j = ks_jac(xs, C.l_wb);
  ks_jac_mul(j, D, K); K[2][0] += T(1); K[3][1] += T(1);
  for (int r = 0; r < 5; ++r) for (int c = 0; c < 7; ++c) A[r][c] += T(2) * K[r][c];
  // stage 4
  for (int r = 0; r < 5; ++r) { xs[r] = x[r] + h * k3[r]; for (int c = 0; c < 7; ++c) D[r][c] = ((c == r + 2) ? T(1) : T(0)) + h * K[r][c]; }
  ks_rhs(xs, u, C.l_wb, k4);
  j = ks_jac(xs, C.l_wb);
  ks_jac_mul(j, D, K); K[2][0] += T(1); K[3][1] += T(1);
  for (int r = 0; r < 5; ++r) {
    const T dxr = h / T(6) * (k1[r] + T(2) * k2[r] + T(2) * k3[r] + k4[r]);
    xn[r] = x[r] + dxr;
    if (inc) inc[r] = dxr;
    for (int c = 0; c < 7; ++c) dc[r][c] = ((c == r + 2) ? T(1) : T(0)) + h / T(6) * (A[r][c] + K[r][c]);
  }
}

// the RK4 increment alone (no Jacobian): trial points of the line search
template <typename T>
MPC_HD void forces_rk4_increment(const ForcesConsts<T>& C, const T* z, T* inc) {
  const T h = C.dt;
  const T* u = z; const T* x = z + 2;
  T k1[5], k2[5], k3[5], k4[5], xs[5];
  ks_rhs(x, u, C.l_wb, k1);
  for (int r = 0; r < 5; ++r) xs[r] = x[r] + T(0.5) * h * k1[r];
  ks_rhs(xs, u, C.l_wb, k2);
  for (int r = 0; r < 5; ++r) xs[r] = x[r] + T(0.5) * h * k2[r];
  ks_rhs(xs, u, C.l_wb, k3);
  for (int r = 0; r < 5; ++r) xs[r] = x[r] + h * k3[r];
  ks_rhs(xs, u, C.l_wb, k4);
  for (int r = 0; r < 5; ++r) inc[r] = h / T(6) * (k1[r] + T(2) * k2[r] + T(2) * k3[r] + k4[r]);
}

// h(z,p) (10) and dh/dz (10x7): friction circle, then ego circle i x obstacle circle j squared distances, i outer, j inner
template <typename T>
MPC_HD void forces_inequalities(const ForcesConsts<T>& C, const T* z, const T* p, T* hv, T dh[10][7]) {
  for (int r = 0; r < 10; ++r) for (int c = 0; c < 7; ++c) dh[r][c] = T(0);
  const T a = z[1], de = z[4], v = z[5], psi = z[6];
  const T tn = m_tan(de);
  const T psid = v * tn / C.l_fric;                 // psi_dot with the configuration's wheelbase literal
  const T q = v * psid;                             // lateral acceleration v^2 tan(delta) / wheelbase
  hv[0] = a * a + q * q;
  dh[0][1] = T(2) * a;
  dh[0][4] = T(2) * q * (v * v * (T(1) + tn * tn) / C.l_fric);
  dh[0][5] = T(2) * q * (T(2) * v * tn / C.l_fric);
  T s, c; m_sincos(psi, &s, &c);
  for (int i = 0; i < 3; ++i) {
    const T o = (i == 0) ? T(0) : ((i == 1) ? C.ego_off : -C.ego_off);
    const T ex = z[2] + o * c, ey = z[3] + o * s;
    for (int jj = 0; jj < 3; ++jj) {
      const T dx = ex - p[4 + 2 * jj], dy = ey - p[5 + 2 * jj];
      const int r = 1 + 3 * i + jj;
      hv[r] = dx * dx + dy * dy;
      dh[r][2] = T(2) * dx; dh[r][3] = T(2) * dy;
      dh[r][6] = T(2) * (dx * (-o * s) + dy * (o * c));
    }
  }
}

// stage objective and gradient (optimizer.py:163-179); terminal variant: states only with the *_terminate weights (:181-195)
template <typename T>
MPC_HD T forces_objective(const ForcesConsts<T>& C, const T* z, const T* p, bool terminal, T* g) {
  const T* w = terminal ? C.Pt : C.Q;
  const T e[5] = {z[2] - p[0], z[3] - p[1], z[4], z[5] - p[2], z[6] - p[3]};
  T f = T(0);
  for (int i = 0; i < 5; ++i) { f += w[i] * e[i] * e[i]; g[2 + i] = T(2) * w[i] * e[i]; }
  g[0] = g[1] = T(0);
  if (!terminal) {
    f += C.R[0] * z[0] * z[0] + C.R[1] * z[1] * z[1];
    g[0] = T(2) * C.R[0] * z[0]; g[1] = T(2) * C.R[1] * z[1];
  }
  return f;
}

// everything one SQP stage linearisation of the formulation produces, packed: c(5) dc(35) h(10) dh(70) f(1) df(7) fN(1) dfN(7) = 136
enum : int { FORCES_OUT_WORDS = 136 };
template <typename T>
MPC_HD void forces_stage_eval(const ForcesConsts<T>& C, const T* z, const T* p, T* out) {
  T xn[5], dc[5][7], hv[10], dh[10][7], g[7], gN[7];
  forces_dynamics(C, z, xn, dc);
  forces_inequalities(C, z, p, hv, dh);
  const T f = forces_objective(C, z, p, false, g);
  const T fN = forces_objective(C, z, p, true, gN);
  int o = 0;
  for (int r = 0; r < 5; ++r) out[o++] = xn[r];
  for (int r = 0; r < 5; ++r) for (int c = 0; c < 7; ++c) out[o++] = dc[r][c];
  for (int r = 0; r < 10; ++r) out[o++] = hv[r];
  for (int r = 0; r < 10; ++r) for (int c = 0; c < 7; ++c) out[o++] = dh[r][c];
  out[o++] = f;
  for (int c = 0; c < 7; ++c) out[o++] = g[c];
  out[o++] = fN;
  for (int c = 0; c < 7; ++c) out[o++] = gN[c];
}

}  // namespace mpcb200
