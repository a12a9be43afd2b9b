// host_sim.cpp -- TEST TOOLING ONLY.  Compiles csrc/warp_core.cuh (the device code of the warp-per-problem CUDA
// kernels) with g++ on the 32-fiber lock-step warp emulator of csrc/warp_ctx.cuh, so the algorithm and its lane
// mappings can be debugged against the oracle in a container without a GPU.  Never loaded by the mpc_b200 package;
// libmpcb200.so has no CPU path and fails loudly without CUDA.
#include <vector>
#include <cstring>
#include <cstdio>
#include <cmath>
#include "../../motion-planning-for-autonomous-driving-with-mpc_b200/csrc/config_params.h"
#include "../../motion-planning-for-autonomous-driving-with-mpc_b200/csrc/warp_core.cuh"
#include "../../motion-planning-for-autonomous-driving-with-mpc_b200/csrc/loop_core.cuh"
#include "../../motion-planning-for-autonomous-driving-with-mpc_b200/csrc/forces_model.cuh"
#include "../../motion-planning-for-autonomous-driving-with-mpc_b200/csrc/forces_core.cuh"

using namespace mpcb200;

template <typename T>
struct Job {
  const mpcb200_config* cfg;
  ParamsT<T> P;
  T* slab;
  const double* xref; double* X; double* U;
  double* lam;          // dual block in / out (mpcb200_solve_dual), or null
  HostWarp* hw;
  int status, iters, nsoc, trace;
  double kkt;
};

template <typename T>
static void lane_body(int lane, void* arg) {
  Job<T>& J = *(Job<T>*)arg;
  WarpCtx w(J.hw, lane);
  T obs[6];
  WarpSolver<T> S(J.P, SlabRef<T>{J.slab, 0}, obs, w);
  S.load(J.xref, J.X, J.U, J.cfg->obstacle, obs);
  ProbState<T> st;
  if (J.lam && S.duals_valid(J.lam)) { S.load_duals(J.lam); S.init_warm(st); } else S.init(st);
  for (int it = 0; it < J.cfg->max_iter && !st.done; ++it) {
    S.iterate(st);
    if ((J.trace == 1 || J.trace == 2) && lane == 0)
      printf("it %3d mu %.2e step %.3e rho %.2e al %.3e ap %.3e ad %.3e c1 %.3e dphi %.3e blk %d/%d status %d\n", st.iters, (double)st.mu,
             (double)st.kkt, (double)st.rho, (double)st.d_al, (double)st.d_ap, (double)st.d_ad, (double)st.d_c1, (double)st.d_dphi, st.d_blk / 16, st.d_blk % 16, st.status);
  }
  S.store(J.xref, J.X, J.U);
  if (J.lam) S.store_duals(J.lam, st.mu);
  if (lane == 0) { J.status = st.status; J.iters = st.iters; J.kkt = (double)st.kkt; J.nsoc = st.nsoc; }
}

template <typename T>
static void run(const mpcb200_config& cfg, const double* xref, double* Xio, double* Uio, int* status, int* iters,
                double* kkt, int B, int trace, double* lam = nullptr) {
  const int N = cfg.N;
  WLayout L(N);
  std::vector<T> buf(L.words + REC_STRIDE + 4);   // + one record: the forward sweep prefetches one record past the end
  HostWarp hw;
  for (int b = 0; b < B; ++b) {
    Job<T> J;
    J.cfg = &cfg; J.P = params_from_config<T>(cfg); J.slab = buf.data();
    J.xref = xref + (size_t)b * 5 * (N + 1); J.X = Xio + (size_t)b * 5 * (N + 1); J.U = Uio + (size_t)b * 2 * N;
    J.lam = lam ? lam + (size_t)b * (14 * N + 2) : nullptr;
    J.hw = &hw; J.trace = trace; J.status = 0; J.iters = 0; J.kkt = 0; J.nsoc = 0;
    for (auto& v : buf) v = T(NAN);        // catch reads of never-written slab words
    hw.run(&lane_body<T>, &J);
    if (status) status[b] = J.status;
    if (iters) iters[b] = J.iters;
    if (kkt) kkt[b] = trace == 3 ? (double)J.nsoc : J.kkt;
  }
}

// ---- the closed loop (loop_core.cuh: the body of mpc_warp_closed_loop_kernel) for one ego per emulated warp
template <typename T>
struct LoopJob {
  const mpcb200_config* cfg;
  ParamsT<T> P;
  T* slab;
  LoopData d;
  int b;
  double* stg;
  HostWarp* hw;
};
template <typename T>
static void loop_lane_body(int lane, void* arg) {
  LoopJob<T>& J = *(LoopJob<T>*)arg;
  WarpCtx w(J.hw, lane);
  T obs[6];
  WarpSolver<T> S(J.P, SlabRef<T>{J.slab, 0}, obs, w);
  const int nx = 5 * (J.P.N + 1);
  closed_loop_ego<T, HESS_RUNTIME>(S, J.d, J.b, J.stg, J.stg + nx, J.stg + 2 * nx, obs, J.cfg->max_iter);
}
template <typename T>
static void run_loop(const mpcb200_config& cfg, const LoopData& d) {
  const int N = cfg.N;
  WLayout L(N);
  std::vector<T> buf(L.words + REC_STRIDE + 4);
  std::vector<double> stg(12 * N + 10);
  HostWarp hw;
  for (int b = 0; b < d.B; ++b) {
    LoopJob<T> J;
    J.cfg = &cfg; J.P = params_from_config<T>(cfg); J.slab = buf.data(); J.d = d; J.b = b; J.stg = stg.data(); J.hw = &hw;
    for (auto& v : buf) v = T(NAN);
    hw.run(&loop_lane_body<T>, &J);
  }
}

// ---- the FORCESPRO-formulation solver core (forces_core.cuh: the device code of mpc_forces_solve_kernel)
template <typename T>
struct FJob {
  const mpcb200_config* cfg;
  FParams<T> fp;
  T* slab;
  const double* xinit; const double* par; const double* Zin; double* Z;
  RoadBounds<T> rb;
  HostWarp* hw;
  int status, iters, trace;
};
template <typename T, bool RB>
static void forces_lane_body(int lane, void* arg) {
  FJob<T>& J = *(FJob<T>*)arg;
  WarpCtx w(J.hw, lane);
  ForcesSolver<T, RB> S(J.fp, SlabRef<T>{J.slab, 0}, w, J.rb);
  S.load(J.xinit, J.par, J.Zin);
  ProbState<T> st;
  S.init(st);
  for (int it = 0; it < J.cfg->max_iter && !st.done; ++it) {
    S.iterate(st);
    if (J.trace && lane == 0)
      printf("it %3d mu %.2e step %.3e rho %.2e al %.3e ap %.3e ad %.3e c1 %.3e dphi %.3e blk %d/%d status %d\n", st.iters, (double)st.mu,
             (double)st.kkt, (double)st.rho, (double)st.d_al, (double)st.d_ap, (double)st.d_ad, (double)st.d_c1, (double)st.d_dphi, st.d_blk / 32, st.d_blk % 32, st.status);
  }
  S.store(J.xinit, J.par, J.Z);
  if (lane == 0) { J.status = st.status; J.iters = st.iters; }
}
template <typename T>
static void run_forces(const mpcb200_config& cfg, const double* Pt, const double* xinit, const double* par, const double* Zin, double* Z,
                       int* status, int* iters, int B, int trace, const double* bl = nullptr, int nl = 0, const double* br = nullptr, int nr = 0,
                       double rmin = 0.0) {
  const int N = cfg.N;
  FLayout L(N, nl > 0 && nr > 0);
  std::vector<T> buf(L.words + 4);
  std::vector<T> bnd(2 * (size_t)(nl + nr) + 2);
  for (int i = 0; i < 2 * nl; ++i) bnd[i] = (T)bl[i];
  for (int i = 0; i < 2 * nr; ++i) bnd[2 * nl + i] = (T)br[i];
  HostWarp hw;
  for (int b = 0; b < B; ++b) {
    FJob<T> J;
    J.cfg = &cfg; J.fp.P = params_from_config<T>(cfg);
    for (int i = 0; i < 5; ++i) J.fp.Pt[i] = (T)Pt[i];
    J.slab = buf.data(); J.xinit = xinit + 5 * (size_t)b; J.par = par + (size_t)b * 10 * N; J.Zin = Zin ? Zin + (size_t)b * 7 * N : nullptr;
    J.Z = Z + (size_t)b * 7 * N; J.hw = &hw; J.trace = trace; J.status = 0; J.iters = 0;
    J.rb.left = bnd.data(); J.rb.right = bnd.data() + 2 * nl; J.rb.nl = nl; J.rb.nr = nr; J.rb.r_min = (T)rmin;
    for (auto& v : buf) v = T(NAN);
    if (nl > 0 && nr > 0) hw.run(&forces_lane_body<T, true>, &J); else hw.run(&forces_lane_body<T, false>, &J);
    if (status) status[b] = J.status;
    if (iters) iters[b] = J.iters;
  }
}

extern "C" {
int hostsim_forces_solve(const mpcb200_config* cfg, const double* Pt, const double* xinit, const double* par, const double* Zin, double* Z,
                         int* status, int* iters, int B, int trace) {
  if (cfg->precision == MPCB200_F64) run_forces<double>(*cfg, Pt, xinit, par, Zin, Z, status, iters, B, trace);
  else run_forces<float>(*cfg, Pt, xinit, par, Zin, Z, status, iters, B, trace);
  return 0;
}
int hostsim_forces_solve_rb(const mpcb200_config* cfg, const double* Pt, const double* xinit, const double* par, const double* Zin, double* Z,
                            int* status, int* iters, int B, int trace, const double* bl, int nl, const double* br, int nr, double rmin) {
  if (cfg->precision == MPCB200_F64) run_forces<double>(*cfg, Pt, xinit, par, Zin, Z, status, iters, B, trace, bl, nl, br, nr, rmin);
  else run_forces<float>(*cfg, Pt, xinit, par, Zin, Z, status, iters, B, trace, bl, nl, br, nr, rmin);
  return 0;
}
int hostsim_solve_dual(const mpcb200_config* cfg, const double* xref, double* X, double* U, double* lam, int* status, int* iters, int B) {
  if (cfg->precision == MPCB200_F64) run<double>(*cfg, xref, X, U, status, iters, nullptr, B, 0, lam);
  else run<float>(*cfg, xref, X, U, status, iters, nullptr, B, 0, lam);
  return 0;
}
int hostsim_closed_loop(const mpcb200_config* cfg, int iter_length, const double* path, const double* orient, double vdes, const double* x0,
                        double* traj, double* ctrl, int* status, int* iters, int B) {
  LoopData d;
  for (int i = 0; i < 6; ++i) d.obstacle[i] = cfg->obstacle[i];
  d.path = path; d.orient = orient; d.x0 = x0; d.traj = traj; d.ctrl = ctrl; d.status = status; d.iters = iters;
  d.desired_velocity = vdes; d.l_wb = cfg->l_wb; d.dt = cfg->dt; d.B = B; d.Tlen = iter_length; d.warm_duals = cfg->warm_duals;
  if (cfg->precision == MPCB200_F64) run_loop<double>(*cfg, d); else run_loop<float>(*cfg, d);
  return 0;
}
// FORCESPRO-formulation stage functions (csrc/forces_model.cuh) in float64: consts = [dt, l_wb, l_fric, ego_off, Q5, R2, Pt5]
void hostsim_forces_eval(const double* consts, const double* z, const double* p, double* out, int n) {
  ForcesConsts<double> C;
  C.dt = consts[0]; C.l_wb = consts[1]; C.l_fric = consts[2]; C.ego_off = consts[3];
  for (int i = 0; i < 5; ++i) { C.Q[i] = consts[4 + i]; C.Pt[i] = consts[11 + i]; }
  C.R[0] = consts[9]; C.R[1] = consts[10];
  for (int i = 0; i < n; ++i) forces_stage_eval<double>(C, z + 7 * i, p + 10 * i, out + FORCES_OUT_WORDS * i);
}
void hostsim_trace_step(int i) { mpc_trace_step = i; }
void hostsim_default_config(mpcb200_config* c, int N, int precision) { default_config(c, N, precision); }
int hostsim_solve(const mpcb200_config* cfg, const double* xref, double* X, double* U, int* status, int* iters,
                  double* kkt, int B, int trace) {
  if (cfg->precision == MPCB200_F64) run<double>(*cfg, xref, X, U, status, iters, kkt, B, trace);
  else run<float>(*cfg, xref, X, U, status, iters, kkt, B, trace);
  return 0;
}
}
