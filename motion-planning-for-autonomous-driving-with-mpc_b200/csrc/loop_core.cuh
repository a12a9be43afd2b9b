// loop_core.cuh -- the receding-horizon loop of CasadiOptimizer.optimize() (/root/reference/MPC_Planner/optimizer.py:596-631)
// for ONE ego handled by one warp: solve, record u_0, plant step + warm-start shift (shift_movement, :645-655), next reference
// window (desired_command_and_trajectory, :657-702).  Shared by the device closed-loop kernel (mpcb200.cu) and the host-side
// algorithm tests (tests/host_sim), like warp_core.cuh.
#pragma once
#include "warp_core.cuh"

namespace mpcb200 {

MPC_HD void plant_euler(double* x, double u0, double u1, double dt, double l_wb) {
  // shift_movement: st = x0 + delta_t * f(x0, u[:,0])   (optimizer.py:649-650; KS model configuration.py:364-368)
  double s, c;
#if defined(__CUDA_ARCH__)
  sincos(x[4], &s, &c);
#else
  s = sin(x[4]); c = cos(x[4]);
#endif
  const double v = x[3], tn = tan(x[2]);
  x[0] += dt * v * c; x[1] += dt * v * s; x[2] += dt * u0; x[3] += dt * u1; x[4] += dt * v / l_wb * tn;
}

// row k+1 of the X_ref block of MPC step i (desired_command_and_trajectory, optimizer.py:657-702, quirk Q8)
MPC_HD void ref_window_row(int i, int k, int N, int Tlen, const double* path, const double* orient, double vdes, double* r) {
  const int idx = (i >= Tlen - N) ? (k + (Tlen - N)) : (i + k + 1);
  r[0] = path[2 * idx]; r[1] = path[2 * idx + 1]; r[2] = 0.0; r[3] = vdes; r[4] = orient[idx];
}
MPC_HD void ref_window_rows(int i, int N, int Tlen, const double* path, const double* orient, double vdes,
                            const double* x_now, double* xref /* [N+1][5] */) {
  for (int j = 0; j < 5; ++j) xref[j] = x_now[j];
  for (int k = 0; k < N; ++k) ref_window_row(i, k, N, Tlen, path, orient, vdes, xref + 5 * (k + 1));
}

#if defined(MPC_DIAG) && !defined(__CUDACC__)
static int mpc_trace_step = -1;      // host emulator only: MPC step whose SQP iterations are printed
#endif

struct LoopData {
  double obstacle[6];
  const double* path;      // [Tlen][2]
  const double* orient;    // [Tlen]
  const double* x0;        // [B][5]
  double* traj;            // [B][Tlen][5]
  double* ctrl;            // [B][Tlen][2]
  int* status;             // [B][Tlen]
  int* iters;              // [B][Tlen]
  double desired_velocity;
  double l_wb, dt;
  int B, Tlen;
  int warm_duals;          // 1: slacks / multipliers / barrier parameter carried across MPC steps (shifted one stage)
};

// The whole loop for ego `b`.  my_xref / my_X / my_U: float64 blocks of this warp ([N+1][5], [N+1][5], [N][2]; shared memory on
// the device) -- the reference's parameter block and warm-start arrays, shifted in float64 exactly like shift_movement does.
template <typename T, int HM>
MPC_HD void closed_loop_ego(const WarpSolver<T, HM>& S, const LoopData& a, int b, double* my_xref, double* my_X, double* my_U, T* obs,
                            int max_iter) {
  const WarpCtx& w = S.w;
  const int lane = S.lane, N = S.P.N, nu = 2 * N;
  ProbState<T> st;
  double x[5];
  for (int j = 0; j < 5; ++j) x[j] = a.x0[(size_t)b * 5 + j];
  // first parameter block and warm start: the initial state tiled, controls zero (optimizer.py:578-583, quirk Q4)
  for (int k = lane; k <= N; k += 32)
    for (int j = 0; j < 5; ++j) { my_xref[5 * k + j] = x[j]; my_X[5 * k + j] = x[j]; }
  for (int k = lane; k < nu; k += 32) my_U[k] = 0.0;
  w.sync();
  for (int i = 0; i < a.Tlen; ++i) {
    if (lane == 0 && a.traj) for (int j = 0; j < 5; ++j) a.traj[((size_t)b * a.Tlen + i) * 5 + j] = x[j];   // quirk Q12
    const bool warm = a.warm_duals && i > 0 && st.status == ST_OPTIMAL;
    if (warm) {
      // the slab still holds the previous step's solution: new parameter block, primal guess from the previous controls
      // (shifted or not, whichever rolls out better), slacks / multipliers carried, barrier restart at mu_warm
      S.load_reference(my_xref, a.obstacle, obs);
      S.warm_primal(st);
      S.init_warm(st);
    } else {
      S.load(my_xref, my_X, my_U, a.obstacle, obs);
      S.init(st);
    }
    for (int it = 0; it < max_iter && !st.done; ++it) {
      S.iterate(st);
#if defined(MPC_DIAG) && !defined(__CUDACC__)
      if (mpc_trace_step == i && lane == 0)
        printf("step %d it %3d mu %.2e step %.3e rho %.2e al %.3e ap %.3e ad %.3e c1 %.3e dphi %.3e blk %d/%d status %d\n", i, st.iters, (double)st.mu,
               (double)st.kkt, (double)st.rho, (double)st.d_al, (double)st.d_ap, (double)st.d_ad, (double)st.d_c1, (double)st.d_dphi, st.d_blk / 16, st.d_blk % 16, st.status);
#endif
    }
    S.store(my_xref, my_X, my_U);
    const double u0 = my_U[0], u1 = my_U[1];
    if (lane == 0) {
      if (a.ctrl) { a.ctrl[((size_t)b * a.Tlen + i) * 2] = u0; a.ctrl[((size_t)b * a.Tlen + i) * 2 + 1] = u1; }
      if (a.status) a.status[(size_t)b * a.Tlen + i] = st.status;
      if (a.iters) a.iters[(size_t)b * a.Tlen + i] = st.iters;
    }
    plant_euler(x, u0, u1, a.dt, a.l_wb);
    w.sync();
    // shift the warm start one stage, repeating the last (optimizer.py:652-653); lane-strided with a register hop
    for (int k0 = 0; k0 < N; k0 += 32) {
      const int k = k0 + lane;
      double nxt[7];
      if (k < N) {
        const int ks = (k + 1 < N) ? k + 1 : N - 1;
        nxt[5] = my_U[2 * ks]; nxt[6] = my_U[2 * ks + 1];
        for (int j = 0; j < 5; ++j) nxt[j] = my_X[5 * (k + 1) + j];
      }
      w.sync();
      if (k < N) {
        my_U[2 * k] = nxt[5]; my_U[2 * k + 1] = nxt[6];
        for (int j = 0; j < 5; ++j) my_X[5 * k + j] = nxt[j];
      }
      w.sync();
    }
    // next window from the new state (optimizer.py:628)
    if (lane == 0) for (int j = 0; j < 5; ++j) my_xref[j] = x[j];
    for (int k = lane; k < N; k += 32) ref_window_row(i, k, N, a.Tlen, a.path, a.orient, a.desired_velocity, my_xref + 5 * (k + 1));
    w.sync();
  }
}

}  // namespace mpcb200
