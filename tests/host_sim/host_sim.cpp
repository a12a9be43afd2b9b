// host_sim.cpp -- TEST TOOLING ONLY.  Compiles csrc/sqp_core.cuh (the device code of the CUDA kernels) with g++ so the
// algorithm can be debugged against the oracle in a container without a GPU.  Never loaded by the mpc_b200 package;
// libmpcb200.so has no CPU path and fails loudly without CUDA.
#include <vector>
#include <cstring>
#include <cstdio>
#include <cmath>
#include "../../motion-planning-for-autonomous-driving-with-mpc_b200/csrc/config_params.h"

using namespace mpcb200;

template <typename T>
static void run(const mpcb200_config& cfg, const double* xref, double* Xio, double* Uio, int* status, int* iters,
                double* kkt, int B, int trace) {
  ParamsT<T> P = params_from_config<T>(cfg);
  const int N = cfg.N;
  Layout L(N);
  std::vector<T> buf(L.words);
  for (int b = 0; b < B; ++b) {
    const double* xr = xref + (size_t)b * 5 * (N + 1);
    double* Xb = Xio + (size_t)b * 5 * (N + 1);
    double* Ub = Uio + (size_t)b * 2 * N;
    Ws<T, 1> ws{buf.data()};
    T obs[6];
    Solver<T, 1> S(P, ws, obs);
    S.load(xr, Xb, Ub, cfg.obstacle, obs);
    ProbState<T> st;
    S.init(st);
    for (int it = 0; it < cfg.max_iter && !st.done; ++it) {
      S.iterate(st);
      if (trace == 2) {
        int bk = 0, bj = 0; double bv = 0;
        for (int k = 0; k < N; ++k) for (int j = 0; j < 7; ++j) { double v = fabs((double)S.DX(k, j)); if (v > bv) { bv = v; bk = k; bj = j; } }
        printf("   max step comp: stage %d comp %d val %.3e | du0 %.3e %.3e | dx_N %.2e %.2e %.2e %.2e %.2e\n", bk, bj, bv, (double)S.DU(0,0), (double)S.DU(0,1),
               (double)S.DX(N-1,0),(double)S.DX(N-1,1),(double)S.DX(N-1,2),(double)S.DX(N-1,3),(double)S.DX(N-1,4));
      }
      if (trace) printf("it %3d mu %.2e step %.3e rho %.2e al %.3e ap %.3e ad %.3e c1 %.3e dphi %.3e blk %d/%d status %d\n", st.iters, (double)st.mu, (double)st.kkt, (double)st.rho, (double)st.d_al, (double)st.d_ap, (double)st.d_ad, (double)st.d_c1, (double)st.d_dphi, st.d_blk / 16, st.d_blk % 16, st.status);
    }
    S.store(xr, Xb, Ub);
    if (status) status[b] = st.status;
    if (iters) iters[b] = st.iters;
    if (kkt) kkt[b] = (double)st.kkt;
  }
}

extern "C" {
void hostsim_default_config(mpcb200_config* c, int N, int precision) { default_config(c, N, precision); }
int hostsim_solve(const mpcb200_config* cfg, const double* xref, double* X, double* U, int* status, int* iters,
                  double* kkt, int B, int trace) {
  if (cfg->precision == MPCB200_F64) run<double>(*cfg, xref, X, U, status, iters, kkt, B, trace);
  else run<float>(*cfg, xref, X, U, status, iters, kkt, B, trace);
  return 0;
}
}
