// mpcb200.cu -- sm_100a kernels + the C ABI of libmpcb200.so (include/mpcb200.h).
//
// Execution model (B200-first, not a translation of anything in the reference -- the reference has no GPU code):
//   * ONE WARP PER EGO INSTANCE (one NLP); a CTA is WPC warps = WPC problems.  Inside a problem the lanes are stages
//     (linearisation, residuals, merit, commit), entries of the 5x6 Riccati block [P | p] (KKT factor sweep) or state
//     components (forward sweep) -- see warp_core.cuh;
//   * the problem's whole KKT working set ("slab": iterate, reference, multipliers, slacks, stage KKT blocks, Riccati
//     gains, step; 90N+21 words = 10.9 KB at N = 30 in fp32, 94N+21 with the exact-Hessian adjoint words) lives in shared memory
//     for the whole solve;
//   * problem data moves HBM <-> shared memory with TMA bulk copies (cp.async.bulk + mbarrier): the CTA's float64
//     xref/X/U rows are one contiguous chunk per array, and in the launch-per-iteration mode each warp's slab is one
//     bulk load + one bulk store per launch;
//   * all SQP iterations of a problem run inside one launch (problems are independent, so no grid-wide sync is ever
//     needed) and a warp stops as soon as ITS problem has converged;
//   * batch 1024 = 512 CTAs x 2 warps over 148 SMs (6-8 warps per SM, every SM busy); larger batches run in waves
//     scheduled by the hardware (up to 8 CTAs of 27.7 KB resident per SM in the Gauss-Newton kernels).
// Tensor cores are not used: the factorisation works on 5x5/2x2 stage blocks along a length-N dependency chain.
#include "mpcb200_internal.cuh"
#include "loop_core.cuh"
#include "forces_model.cuh"

// resident warps per SM the float32 kernels are compiled for: 16 = 4 per sub-partition = 128 registers per thread at most (the
// kernel needs ~120); 18-20 would need <= 96 registers (a sub-partition then holds 5 warps) and spills
#ifndef MPC_WARPS_PER_SM
#define MPC_WARPS_PER_SM 16
#endif
template <typename T, int WPC> struct MinBlocks { static constexpr int v = (sizeof(T) == 4 ? MPC_WARPS_PER_SM : 8) / WPC; };

// ===================================================================================================== solve kernel
// One warp per ego instance, PERSISTENT: a warp solves problem (global warp index), then claims further problems from a
// device counter until the batch is empty -- problems need 4 ... 60 SQP iterations, so static pairing would idle a finished
// warp until its CTA partner is done.  All SQP iterations of a problem run inside the launch (MODE_ONESHOT), or `n_iter` of
// them with the slab round-tripping HBM <-> shared memory by TMA (stepwise modes).
// DUAL = 1: the instantiation behind mpcb200_solve_dual (multipliers / slacks in and out).  A separate instantiation because the
// dual warm start (init_warm) is a second copy of the initialisation code: inlined into the plain kernel it cost 5 registers and
// 2 % of the batch-1024 time without ever being executed there.
template <typename T, int WPC, int HM, int DUAL>
__global__ void __launch_bounds__(32 * WPC, MinBlocks<T, WPC>::v) mpc_warp_solve_kernel(const __grid_constant__ SolveArgs<T> a) {
  unsigned char* const smem_raw = mpc_dyn_smem;
  __shared__ __align__(8) uint64_t bar_x[WPC];     // xref staging
  __shared__ __align__(8) uint64_t bar_w[WPC];     // slab image (stepwise modes)
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.P.N;
  const WLayout L(N, rec_stride_for(HM));
  const Smem<T, WPC> sm(smem_raw, N, L.words);
  const int nx = sm.nx, nu = sm.nu;
  const int total_warps = gridDim.x * WPC;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < WPC; ++i) { mbar_init(&bar_x[i], 1); mbar_init(&bar_w[i], 1); }
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();                                  // the only CTA-wide barrier: from here on the warps are independent
  // Programmatic dependent launch: the float32 pass lets the refinement grid become resident right away (its launch latency
  // then hides under this grid); the refinement grid waits here until the float32 grid has completed and flushed its results.
  if (a.pdl_primary) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (a.refine) asm volatile("griddepcontrol.wait;" ::: "memory");

  const WarpCtx w;
  T obs[6];
  WarpSolver<T, HM> S(a.P, SlabRef<T>{wid * L.words}, obs, w);
  const bool need_xref = (a.mode != MODE_ITER);
  const bool need_init = (a.mode == MODE_ONESHOT || a.mode == MODE_BEGIN);
  const bool need_warm = need_init && !a.cold;
  const int nwork = a.refine ? a.ctr->q_count : a.B;
  uint32_t ph_x = 0, ph_w = 0;

  for (int item = blockIdx.x * WPC + wid; item < nwork;) {
    const int b = a.refine ? a.q_list[item] : item;
    ProbState<T> st;
    const double* xr = nullptr;
    if (need_xref) xr = fetch_xref(a.xref + (size_t)b * nx, sm.xstg(wid), nx, &bar_x[wid], ph_x, lane);
    if (need_init) {
      S.load(xr, need_warm ? a.Xin + (size_t)b * nx : nullptr, need_warm ? a.Uin + (size_t)b * nu : nullptr, a.obstacle, obs);
      bool warm_duals = false;
      if (DUAL) {
        const double* const lam_b = a.lam + (size_t)b * WarpSolver<T, HM>::lam_words(N);
        warm_duals = need_warm && !a.refine && S.duals_valid(lam_b);
        if (warm_duals) { S.load_duals(lam_b); S.init_warm(st); }                                    // dual warm start
      }
      if (!warm_duals) S.init(st);
    } else {
      // resume: the slab image comes back by one TMA bulk copy, the per-problem scalars by plain loads
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        const uint32_t bytes = (uint32_t)((size_t)L.words * sizeof(T));
        mbar_expect_tx(&bar_w[wid], bytes);
        tma_load_1d(sm.slab(wid), a.slab + (size_t)b * L.words, bytes, &bar_w[wid]);
      }
      mbar_wait(&bar_w[wid], ph_w);
      ph_w ^= 1u;
      st = a.state[b];
#pragma unroll
      for (int j = 0; j < 6; ++j) obs[j] = a.obs_shift[(size_t)b * 6 + j];
    }

    if (a.mode != MODE_END) {
      for (int it = 0; it < a.n_iter && !st.done; ++it) S.iterate(st);
    }

    if (a.mode == MODE_ONESHOT || a.mode == MODE_END) {
      // solution back to float64 row-major (rho is added back in float64): coalesced stores straight from the slab
      S.store(xr, a.X + (size_t)b * nx, a.U + (size_t)b * nu);
      if (DUAL) S.store_duals(a.lam + (size_t)b * WarpSolver<T, HM>::lam_words(N), st.mu);
      if (lane == 0) {
        if (a.status) a.status[b] = st.status;
        if (a.iters) a.iters[b] = st.iters + (a.refine ? a.iters[b] : 0);
        // a float32 pass with a refinement queue hands every problem it did not converge to the float64 pass
        if (a.q_list && !a.refine && st.status != ST_OPTIMAL && st.status != ST_INFEASIBLE_X0) a.q_list[atomicAdd(&a.ctr->q_count, 1)] = b;
      }
    } else {
      // keep the slab + scalars for the next launch
      __syncwarp();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_1d(a.slab + (size_t)b * L.words, sm.slab(wid), (uint32_t)((size_t)L.words * sizeof(T)));
        tma_store_commit_wait();
        a.state[b] = st;
#pragma unroll
        for (int j = 0; j < 6; ++j) a.obs_shift[(size_t)b * 6 + j] = obs[j];
      }
      __syncwarp();
    }
    if (!a.dynamic) break;
    int nxt = 0;
    if (lane == 0) nxt = total_warps + atomicAdd(&a.ctr->next, 1);
    item = __shfl_sync(0xffffffffu, nxt, 0);
  }
  // self-resetting counters: the last warp to leave zeroes what this launch used (an empty refinement queue used nothing)
  if ((a.dynamic || a.refine) && lane == 0 && !(a.refine && nwork == 0)) {
    __threadfence();
    if (atomicAdd(&a.ctr->done, 1) == total_warps - 1) {
      a.ctr->next = 0; a.ctr->done = 0;
      if (a.refine) a.ctr->q_count = 0;
    }
  }
}

// ===================================================================================================== per-problem scenarios
// BASELINE configs[4] ("all scenarios x random inits"): ONE launch over problems that belong to DIFFERENT scenarios -- weights,
// time step and obstacle come from a device table indexed by the problem's scenario id instead of the launch constants.  A
// separate (Gauss-Newton, cold start, fused) kernel so that the single-scenario kernel keeps its constants in the constant bank:
// here the warp keeps the problem's ParamsT in shared memory and the solver core reads it from there.
template <typename T>
struct ScnArgs {
  ParamsT<T> P;                     // everything that is not per scenario (bounds, solver options, N)
  const mpcb200_scenario* table;    // [n_scn] device copy of the scenario table
  const int* scn_id;                // [B]
  const double* xref; double* X; double* U;
  int* status; int* iters;
  WorkCtr* ctr; int* q_list;
  int B, n_scn, refine, dynamic, pdl_primary;
};

template <typename T, int WPC>
__global__ void __launch_bounds__(32 * WPC, MinBlocks<T, WPC>::v) mpc_warp_solve_scenarios_kernel(const __grid_constant__ ScnArgs<T> a) {
  unsigned char* const smem_raw = mpc_dyn_smem;
  __shared__ __align__(8) uint64_t bar_x[WPC];
  __shared__ ParamsT<T> sp[WPC];
  __shared__ double sobst[WPC][6];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.P.N;
  const WLayout L(N, rec_stride_for(HESS_GN));
  const Smem<T, WPC> sm(smem_raw, N, L.words);
  const int nx = sm.nx, nu = sm.nu;
  const int total_warps = gridDim.x * WPC;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < WPC; ++i) mbar_init(&bar_x[i], 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (a.pdl_primary) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (a.refine) asm volatile("griddepcontrol.wait;" ::: "memory");
  const WarpCtx w;
  T obs[6];
  const int nwork = a.refine ? a.ctr->q_count : a.B;
  uint32_t ph_x = 0;
  for (int item = blockIdx.x * WPC + wid; item < nwork;) {
    const int b = a.refine ? a.q_list[item] : item;
    // this problem's constants: launch-wide ones + its scenario's row of the table
    __syncwarp();
    if (lane == 0) {
      int sid = a.scn_id[b];
      sid = sid < 0 ? 0 : (sid >= a.n_scn ? a.n_scn - 1 : sid);
      const mpcb200_scenario& sc = a.table[sid];
      ParamsT<T> p = a.P;
      p.dt = (T)sc.dt; p.r_sum = (T)sc.r_sum;
      for (int i = 0; i < 5; ++i) p.Q[i] = (T)sc.Q[i];
      p.R[0] = (T)sc.R[0]; p.R[1] = (T)sc.R[1];
      sp[wid] = p;
      for (int i = 0; i < 6; ++i) sobst[wid][i] = sc.obstacle[i];
    }
    __syncwarp();
    WarpSolver<T, HESS_GN> S(sp[wid], SlabRef<T>{wid * L.words}, obs, w);
    ProbState<T> st;
    const double* xr = fetch_xref(a.xref + (size_t)b * nx, sm.xstg(wid), nx, &bar_x[wid], ph_x, lane);
    if (a.refine) S.load(xr, a.X + (size_t)b * nx, a.U + (size_t)b * nu, sobst[wid], obs);      // warm start = the float32 result
    else S.load(xr, nullptr, nullptr, sobst[wid], obs);
    S.init(st);
    for (int it = 0; it < a.P.max_iter && !st.done; ++it) S.iterate(st);
    S.store(xr, a.X + (size_t)b * nx, a.U + (size_t)b * nu);
    if (lane == 0) {
      if (a.status) a.status[b] = st.status;
      if (a.iters) a.iters[b] = st.iters + (a.refine ? a.iters[b] : 0);
      if (a.q_list && !a.refine && st.status != ST_OPTIMAL && st.status != ST_INFEASIBLE_X0) a.q_list[atomicAdd(&a.ctr->q_count, 1)] = b;
    }
    if (!a.dynamic) break;
    int nxt = 0;
    if (lane == 0) nxt = total_warps + atomicAdd(&a.ctr->next, 1);
    item = __shfl_sync(0xffffffffu, nxt, 0);
  }
  if ((a.dynamic || a.refine) && lane == 0 && !(a.refine && nwork == 0)) {
    __threadfence();
    if (atomicAdd(&a.ctr->done, 1) == total_warps - 1) {
      a.ctr->next = 0; a.ctr->done = 0;
      if (a.refine) a.ctr->q_count = 0;
    }
  }
}

// ===================================================================================================== closed loop
// The whole receding-horizon loop of CasadiOptimizer.optimize() (optimizer.py:596-631), one ego per warp, no host round trip
// between MPC steps (loop body: loop_core.cuh).  Persistent warps like the solve kernel.  Shared memory per warp: the KKT slab
// + the float64 parameter block and warm-start arrays the reference shifts every step ([N+1][5], [N+1][5], [N][2]).
template <typename T>
struct LoopArgs {
  ParamsT<T> P;
  LoopData d;
  WorkCtr* ctr;
  int dynamic;
};
static size_t loop_smem_bytes_for(int N, int words, size_t elem, int wpc) {
  return (size_t)wpc * ((size_t)words * elem + (size_t)(12 * N + 10) * sizeof(double));
}

template <typename T, int WPC, int HM>
__global__ void __launch_bounds__(32 * WPC) mpc_warp_closed_loop_kernel(const __grid_constant__ LoopArgs<T> a) {
  unsigned char* const smem_raw = mpc_dyn_smem;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.P.N;
  const WLayout L(N, rec_stride_for(HM));
  const int nx = 5 * (N + 1), nu = 2 * N;
  double* const stg = reinterpret_cast<double*>(smem_raw + (size_t)WPC * L.words * sizeof(T)) + (size_t)wid * (2 * nx + nu);
  const int total_warps = gridDim.x * WPC;
  const WarpCtx w;
  T obs[6];
  WarpSolver<T, HM> S(a.P, SlabRef<T>{wid * L.words}, obs, w);
  for (int b = blockIdx.x * WPC + wid; b < a.d.B;) {
    closed_loop_ego<T, HM>(S, a.d, b, stg, stg + nx, stg + 2 * nx, obs, a.P.max_iter);
    if (!a.dynamic) break;
    int nxt = 0;
    if (lane == 0) nxt = total_warps + atomicAdd(&a.ctr->next, 1);
    b = __shfl_sync(0xffffffffu, nxt, 0);
  }
  if (a.dynamic && lane == 0) {
    __threadfence();
    if (atomicAdd(&a.ctr->done, 1) == total_warps - 1) { a.ctr->next = 0; a.ctr->done = 0; }
  }
}

// ===================================================================================================== small kernels
__global__ void plant_step_shift_kernel(double* x, double* U, double* X, double* u_applied, int B, int N, double dt, double l_wb) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double* xb = x + (size_t)b * 5;
  double* Ub = U + (size_t)b * 2 * N;
  double* Xb = X + (size_t)b * 5 * (N + 1);
  const double u0 = Ub[0], u1 = Ub[1];
  if (u_applied) { u_applied[2 * b] = u0; u_applied[2 * b + 1] = u1; }
  double xs[5];
  for (int j = 0; j < 5; ++j) xs[j] = xb[j];
  plant_euler(xs, u0, u1, dt, l_wb);
  for (int j = 0; j < 5; ++j) xb[j] = xs[j];
  for (int k = 0; k < N - 1; ++k) { Ub[2 * k] = Ub[2 * k + 2]; Ub[2 * k + 1] = Ub[2 * k + 3]; }
  for (int k = 0; k < N; ++k) for (int j = 0; j < 5; ++j) Xb[5 * k + j] = Xb[5 * (k + 1) + j];
}

__global__ void build_ref_window_kernel(int i, int Tlen, const double* path, const double* orient, double vdes, const double* x,
                                        double* xref, int B, int N) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  ref_window_rows(i, N, Tlen, path, orient, vdes, x + (size_t)b * 5, xref + (size_t)b * 5 * (N + 1));
}

// FORCESPRO-formulation stage linearisation (forces_model.cuh): one thread per (z, p) point, float64, one warp per CTA.  The 136
// output words of a point are staged through shared memory so that the global stores are coalesced.
__global__ void __launch_bounds__(32) forces_stage_eval_kernel(ForcesConsts<double> C, const double* __restrict__ z, const double* __restrict__ p,
                                                                 double* __restrict__ out, int n) {
  __shared__ double stage[32 * (FORCES_OUT_WORDS + 1)];
  const int i = blockIdx.x * 32 + threadIdx.x;
  if (i < n) {
    double zz[7], pp[10], o[FORCES_OUT_WORDS];
    for (int j = 0; j < 7; ++j) zz[j] = z[(size_t)i * 7 + j];
    for (int j = 0; j < 10; ++j) pp[j] = p[(size_t)i * 10 + j];
    forces_stage_eval<double>(C, zz, pp, o);
    for (int j = 0; j < FORCES_OUT_WORDS; ++j) stage[threadIdx.x * (FORCES_OUT_WORDS + 1) + j] = o[j];
  }
  __syncthreads();
  const int base = blockIdx.x * 32;
  const int cnt = min(32, n - base) * FORCES_OUT_WORDS;
  for (int e = threadIdx.x; e < cnt; e += 32) {
    const int r = e / FORCES_OUT_WORDS, c = e - r * FORCES_OUT_WORDS;
    out[(size_t)base * FORCES_OUT_WORDS + e] = stage[r * (FORCES_OUT_WORDS + 1) + c];
  }
}

template <typename T, int WPC, int HM>
static cudaError_t launch_solve(mpcb200_handle* h, SolveArgs<T>& a, cudaStream_t s, const KernelPlan& k, int nwork) {
  const int ctas = grid_for(k, nwork);
  a.dynamic = (a.refine || nwork > ctas * WPC) ? 1 : 0;
  h->launches++;
  if (a.refine) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(32 * WPC); cfg.dynamicSmemBytes = k.smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (a.lam) return cudaLaunchKernelEx(&cfg, mpc_warp_solve_kernel<T, WPC, HM, 1>, a);
    return cudaLaunchKernelEx(&cfg, mpc_warp_solve_kernel<T, WPC, HM, 0>, a);
  }
  if (a.lam) mpc_warp_solve_kernel<T, WPC, HM, 1><<<ctas, 32 * WPC, k.smem, s>>>(a);
  else mpc_warp_solve_kernel<T, WPC, HM, 0><<<ctas, 32 * WPC, k.smem, s>>>(a);
  return cudaGetLastError();
}
template <typename T, int WPC, int HM>
static cudaError_t launch_loop(mpcb200_handle* h, LoopArgs<T>& a, cudaStream_t s) {
  const int ctas = grid_for(h->loop, a.d.B);
  a.dynamic = (a.d.B > ctas * WPC) ? 1 : 0;
  mpc_warp_closed_loop_kernel<T, WPC, HM><<<ctas, 32 * WPC, h->loop.smem, s>>>(a);
  h->launches++;
  return cudaGetLastError();
}

// kernel instantiations: arithmetic type x problems per CTA (1, 2, 4) x Hessian mode
#define MPC_DISPATCH_WPC(WPCV, ...)                          \
  switch (WPCV) {                                            \
    case 4: { constexpr int W = 4; __VA_ARGS__; } break;     \
    case 2: { constexpr int W = 2; __VA_ARGS__; } break;     \
    default: { constexpr int W = 1; __VA_ARGS__; } break;    \
  }
template <typename T>
static cudaError_t dispatch_solve(mpcb200_handle* h, SolveArgs<T>& a, cudaStream_t s, const KernelPlan& k, int nwork) {
  const bool ex = a.P.hessian == HESS_EXACT;
  cudaError_t e = cudaSuccess;
  MPC_DISPATCH_WPC(k.wpc, e = ex ? launch_solve<T, W, HESS_EXACT>(h, a, s, k, nwork) : launch_solve<T, W, HESS_GN>(h, a, s, k, nwork));
  return e;
}
template <typename T>
static cudaError_t dispatch_loop(mpcb200_handle* h, LoopArgs<T>& a, cudaStream_t s) {
  const bool ex = a.P.hessian == HESS_EXACT;
  cudaError_t e = cudaSuccess;
  MPC_DISPATCH_WPC(h->loop.wpc, e = ex ? launch_loop<T, W, HESS_EXACT>(h, a, s) : launch_loop<T, W, HESS_GN>(h, a, s));
  return e;
}

template <typename T>
static cudaError_t plan_solve(KernelPlan& k, bool exact, int optin, int sms) {
  cudaError_t e = cudaSuccess;
  int dual_ctas = 0;
  MPC_DISPATCH_WPC(k.wpc, e = exact ? plan_kernel(mpc_warp_solve_kernel<T, W, HESS_EXACT, 0>, W, k.smem, optin, sms, &k.max_ctas)
                                    : plan_kernel(mpc_warp_solve_kernel<T, W, HESS_GN, 0>, W, k.smem, optin, sms, &k.max_ctas));
  if (e == cudaSuccess) {
    MPC_DISPATCH_WPC(k.wpc, e = exact ? plan_kernel(mpc_warp_solve_kernel<T, W, HESS_EXACT, 1>, W, k.smem, optin, sms, &dual_ctas)
                                      : plan_kernel(mpc_warp_solve_kernel<T, W, HESS_GN, 1>, W, k.smem, optin, sms, &dual_ctas));
    if (e == cudaSuccess && dual_ctas < k.max_ctas) k.max_ctas = dual_ctas;      // one grid shape for both instantiations
  }
  return e;
}
template <typename T>
static cudaError_t plan_loop(KernelPlan& k, bool exact, int optin, int sms) {
  cudaError_t e = cudaSuccess;
  MPC_DISPATCH_WPC(k.wpc, e = exact ? plan_kernel(mpc_warp_closed_loop_kernel<T, W, HESS_EXACT>, W, k.smem, optin, sms, &k.max_ctas)
                                    : plan_kernel(mpc_warp_closed_loop_kernel<T, W, HESS_GN>, W, k.smem, optin, sms, &k.max_ctas));
  return e;
}
// problems per CTA: the preferred count, else the largest of 4 / 2 / 1 whose CTA fits the opt-in shared memory
static bool choose_wpc(KernelPlan& k, int pref, size_t (*bytes)(int, int, size_t, int), int N, int words, size_t elem, size_t smem_max) {
  const int cand[4] = {pref, 4, 2, 1};
  for (int i = 0; i < 4; ++i) {
    const int wpc = cand[i];
    if (wpc != 1 && wpc != 2 && wpc != 4) continue;
    const size_t need = bytes(N, words, elem, wpc);
    if (need + 1024 <= smem_max) { k.wpc = wpc; k.smem = need; return true; }
  }
  return false;
}

static int ensure_stepwise_scratch(mpcb200_handle* h) {
  if (h->slab) return 0;
  const size_t mb = (size_t)h->cfg.max_batch;
  CK(cudaMalloc(&h->slab, mb * h->words * h->elem));
  CK(cudaMalloc(&h->state, mb * sizeof(ProbState<double>)));
  CK(cudaMalloc(&h->obs_shift, mb * 6 * h->elem));
  return 0;
}

template <typename T>
static int do_solve(mpcb200_handle* h, int mode, int n_iter, const double* xref, double* X, double* U, int* status, int* iters,
                    int B, cudaStream_t s, int cold = 0, const double* Xin = nullptr, const double* Uin = nullptr, double* lam = nullptr) {
  if (B <= 0) return 0;
  if (B > h->cfg.max_batch) { h->err = "B exceeds cfg.max_batch"; return -2; }
  if (((uintptr_t)xref | (uintptr_t)X | (uintptr_t)U | (uintptr_t)Xin | (uintptr_t)Uin) & 7u) { h->err = "float64 arrays must be 8-byte aligned"; return -2; }
  if (mode != MODE_ONESHOT) { int rc = ensure_stepwise_scratch(h); if (rc) return rc; }
  SolveArgs<T> a;
  a.P = params_from_config<T>(h->cfg);
  for (int i = 0; i < 6; ++i) a.obstacle[i] = h->cfg.obstacle[i];
  a.xref = xref; a.X = X; a.U = U; a.status = status; a.iters = iters; a.lam = lam;
  a.Xin = Xin ? Xin : X; a.Uin = Uin ? Uin : U;
  a.slab = (T*)h->slab; a.state = (ProbState<T>*)h->state; a.obs_shift = (T*)h->obs_shift;
  a.ctr = h->ctr;
  // a float32 handle with cfg.refine_f64 queues what it did not converge for the float64 pass that follows (fused mode only)
  a.q_list = (mode == MODE_ONESHOT && sizeof(T) == 4 && h->cfg.refine_f64 && status) ? h->q_list : nullptr;
  a.B = B; a.mode = mode; a.n_iter = n_iter; a.cold = cold; a.refine = 0; a.dynamic = 0; a.pdl_primary = a.q_list ? 1 : 0;
  cudaError_t e;
  // cfg.warps_per_cta = 8 | 16: the phase-aligned kernel (aligned_solver.cu) for fused float32 Gauss-Newton solves
  if (sizeof(T) == 4 && mode == MODE_ONESHOT && !lam && a.P.hessian == HESS_GN && (h->cfg.warps_per_cta == 8 || h->cfg.warps_per_cta == 16))
    e = launch_solve_aligned(h, reinterpret_cast<SolveArgs<float>&>(a), s, B);
  else e = dispatch_solve<T>(h, a, s, h->solve, B);
  if (e != cudaSuccess) return fail(h, "mpc_warp_solve_kernel launch", e);
  return 0;
}

// second pass of a float32 handle with cfg.refine_f64: float64 arithmetic (float32 tolerances) on the instances the float32
// pass queued (status other than 1), warm-started from their float32 result.  A small persistent grid: the queue is usually
// short or empty (then every warp leaves after one load).
static int refine_pass(mpcb200_handle* h, const double* xref, double* X, double* U, int* status, int* iters, int B, cudaStream_t s,
                       double* lam = nullptr) {
  if (!status) return 0;                                // without a status buffer nothing was queued
  SolveArgs<double> a;
  a.P = params_from_config<double>(h->cfg);
  for (int i = 0; i < 6; ++i) a.obstacle[i] = h->cfg.obstacle[i];
  a.xref = xref; a.X = X; a.U = U; a.status = status; a.iters = iters; a.lam = lam;
  a.Xin = X; a.Uin = U;
  a.slab = nullptr; a.state = nullptr; a.obs_shift = nullptr;
  a.ctr = h->ctr; a.q_list = h->q_list;
  a.B = B; a.mode = MODE_ONESHOT; a.n_iter = h->cfg.max_iter; a.cold = 0; a.refine = 1; a.dynamic = 1; a.pdl_primary = 0;
  cudaError_t e = dispatch_solve<double>(h, a, s, h->refine, B);
  if (e != cudaSuccess) return fail(h, "mpc_warp_solve_kernel<double> (refinement) launch", e);
  return 0;
}

// one solve (+ the float64 refinement pass of a refine_f64 handle) with separate warm-start-in and result-out arrays
static int solve_io(mpcb200_handle* h, const double* xref, const double* Xin, const double* Uin, double* X, double* U,
                    int32_t* status, int32_t* iters, int32_t B, cudaStream_t s, int cold, double* lam = nullptr) {
  if (h->cfg.precision == MPCB200_F64) return do_solve<double>(h, MODE_ONESHOT, h->cfg.max_iter, xref, X, U, status, iters, B, s, cold, Xin, Uin, lam);
  int rc = do_solve<float>(h, MODE_ONESHOT, h->cfg.max_iter, xref, X, U, status, iters, B, s, cold, Xin, Uin, lam);
  if (rc == 0 && h->cfg.refine_f64 && B > 0) rc = refine_pass(h, xref, X, U, status, iters, B, s, lam);
  return rc;
}

template <typename T>
static cudaError_t closed_loop_t(mpcb200_handle* h, int32_t iter_length, const double* d_path, const double* d_orientation, double desired_velocity,
                                 const double* d_x0, double* d_traj, double* d_ctrl, int32_t* d_status, int32_t* d_iters, int32_t B, cudaStream_t s) {
  LoopArgs<T> a;
  a.P = params_from_config<T>(h->cfg);
  for (int i = 0; i < 6; ++i) a.d.obstacle[i] = h->cfg.obstacle[i];
  a.d.path = d_path; a.d.orient = d_orientation; a.d.x0 = d_x0; a.d.traj = d_traj; a.d.ctrl = d_ctrl; a.d.status = d_status; a.d.iters = d_iters;
  a.d.desired_velocity = desired_velocity; a.d.l_wb = h->cfg.l_wb; a.d.dt = h->cfg.dt; a.d.B = B; a.d.Tlen = iter_length;
  a.d.warm_duals = h->cfg.warm_duals;
  a.ctr = h->ctr; a.dynamic = 0;
  return dispatch_loop<T>(h, a, s);
}

template <typename T>
static cudaError_t plan_scn(KernelPlan& k, int optin, int sms) {
  cudaError_t e = cudaSuccess;
  switch (k.wpc) {
    case 2: e = plan_kernel(mpc_warp_solve_scenarios_kernel<T, 2>, 2, k.smem, optin, sms, &k.max_ctas); break;
    default: k.wpc = 1; e = plan_kernel(mpc_warp_solve_scenarios_kernel<T, 1>, 1, k.smem, optin, sms, &k.max_ctas); break;
  }
  return e;
}
template <typename T>
static cudaError_t launch_scn(mpcb200_handle* h, ScnArgs<T>& a, cudaStream_t s, const KernelPlan& k, int nwork) {
  const int ctas = grid_for(k, nwork);
  a.dynamic = (a.refine || nwork > ctas * k.wpc) ? 1 : 0;
  h->launches++;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(32 * k.wpc); cfg.dynamicSmemBytes = k.smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = a.refine ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (k.wpc == 2) return cudaLaunchKernelEx(&cfg, mpc_warp_solve_scenarios_kernel<T, 2>, a);
  return cudaLaunchKernelEx(&cfg, mpc_warp_solve_scenarios_kernel<T, 1>, a);
}

template <typename T>
static void fill_scn_args(mpcb200_handle* h, ScnArgs<T>& a, const double* xref, const int32_t* sid, double* X, double* U, int32_t* status,
                          int32_t* iters, int32_t B) {
  a.P = params_from_config<T>(h->cfg);
  a.table = h->scn_table; a.scn_id = sid; a.n_scn = h->n_scn;
  a.xref = xref; a.X = X; a.U = U; a.status = status; a.iters = iters;
  a.ctr = h->ctr; a.q_list = nullptr; a.B = B; a.refine = 0; a.dynamic = 0; a.pdl_primary = 0;
}

extern "C" {

int32_t mpcb200_abi_version(void) { return MPCB200_ABI_VERSION; }

void mpcb200_default_config(mpcb200_config* cfg, int32_t N, int32_t precision) { default_config(cfg, N, precision); }

const char* mpcb200_last_error(const mpcb200_handle* h) { return h ? h->err.c_str() : create_err().c_str(); }

int mpcb200_create(const mpcb200_config* cfg, mpcb200_handle** out) {
  mpcb200_handle* h = nullptr;
  if (!cfg || !out) { create_err() = "null argument"; return -2; }
  if (cfg->abi_version != MPCB200_ABI_VERSION) { create_err() = "abi_version mismatch"; return -2; }
  if (cfg->N < 4 || cfg->N > 128 || cfg->max_batch < 1) { create_err() = "N must be in [4, 128] (tested range), max_batch >= 1"; return -2; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { fail(nullptr, "no CUDA device (libmpcb200 has no CPU path)", e); return -3; }
  if (cfg->device < 0 || cfg->device >= ndev) { create_err() = "cfg.device out of range"; return -2; }
  DeviceGuard guard(cfg->device);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 9) { create_err() = "libmpcb200 needs TMA bulk copies (sm_90+); built for sm_100a"; return -3; }
  h = new (std::nothrow) mpcb200_handle();
  if (!h) { create_err() = "out of host memory"; return -4; }
  h->cfg = *cfg;
  h->launches = 0; h->slab = h->state = h->obs_shift = nullptr; h->ctr = nullptr; h->q_list = nullptr; h->scn_table = nullptr; h->n_scn = 0; h->forces_planned = 0; h->rb_f32 = h->rb_f64 = nullptr; h->rb_nl = h->rb_nr = 0; h->rb_rmin = 0.0;
  h->d_xref = h->d_X = h->d_U = nullptr; h->d_status = h->d_iters = nullptr; h->h_pin = nullptr;
  h->sw_xref = nullptr; h->sw_B = 0;
  const bool exact = cfg->hessian == MPCB200_HESS_EXACT;
  const WLayout L(cfg->N, rec_stride_for(exact ? HESS_EXACT : HESS_GN));
  h->words = L.words;
  const bool f64 = cfg->precision == MPCB200_F64;
  h->elem = f64 ? 8 : 4;
  const size_t smem_max = prop.sharedMemPerBlockOptin;   // 227 KB on B200
  const int optin = (int)prop.sharedMemPerBlockOptin, sms = prop.multiProcessorCount;
  const int pref = (cfg->warps_per_cta == 1 || cfg->warps_per_cta == 2 || cfg->warps_per_cta == 4) ? cfg->warps_per_cta : 2;
  if (!choose_wpc(h->solve, pref, smem_bytes_for, cfg->N, L.words, h->elem, smem_max) ||
      !choose_wpc(h->loop, pref, loop_smem_bytes_for, cfg->N, L.words, h->elem, smem_max)) {
    create_err() = "horizon too long: the per-problem KKT slab does not fit shared memory"; delete h; return -2;
  }
  e = f64 ? plan_solve<double>(h->solve, exact, optin, sms) : plan_solve<float>(h->solve, exact, optin, sms);
  if (e == cudaSuccess) e = f64 ? plan_loop<double>(h->loop, exact, optin, sms) : plan_loop<float>(h->loop, exact, optin, sms);
  if (e != cudaSuccess) { fail(nullptr, "kernel configuration (shared memory opt-in / occupancy)", e); delete h; return -1; }
  h->refine.wpc = 0;
  if (!f64 && cfg->refine_f64) {
    if (!choose_wpc(h->refine, 1, smem_bytes_for, cfg->N, L.words, 8, smem_max)) {
      create_err() = "horizon too long for the float64 refinement pass"; delete h; return -2;
    }
    e = plan_solve<double>(h->refine, exact, optin, sms);
    if (e != cudaSuccess) { fail(nullptr, "kernel configuration (float64 refinement)", e); delete h; return -1; }
    if (h->refine.max_ctas > sms) h->refine.max_ctas = sms;          // the queue is short: one CTA per SM is plenty
    if (cudaMalloc(&h->q_list, (size_t)cfg->max_batch * sizeof(int)) != cudaSuccess) { fail(nullptr, "cudaMalloc(refinement queue)", cudaGetLastError()); delete h; return -1; }
  }
  if (cudaMalloc(&h->ctr, sizeof(WorkCtr)) != cudaSuccess || cudaMemset(h->ctr, 0, sizeof(WorkCtr)) != cudaSuccess) {
    fail(nullptr, "cudaMalloc(work counters)", cudaGetLastError()); cudaFree(h->q_list); delete h; return -1;
  }
  *out = h;
  return 0;
}

void mpcb200_destroy(mpcb200_handle* h) {
  if (!h) return;
  DeviceGuard guard(h->cfg.device);
  cudaFree(h->slab); cudaFree(h->state); cudaFree(h->obs_shift); cudaFree(h->ctr); cudaFree(h->q_list); cudaFree(h->scn_table); cudaFree(h->rb_f32); cudaFree(h->rb_f64);
  if (h->h_pin) for (int i = 0; i < MPCB200_HOST_STREAMS; ++i) cudaStreamDestroy(h->hs[i]);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  cudaFree(h->d_xref); cudaFree(h->d_X); cudaFree(h->d_U); cudaFree(h->d_status); cudaFree(h->d_iters);
  delete h;
}

int mpcb200_solve(mpcb200_handle* h, const double* d_xref, double* d_X, double* d_U, int32_t* d_status, int32_t* d_iters,
                  int32_t B, void* stream) {
  if (!h) return -2;
  DeviceGuard guard(h->cfg.device);
  return solve_io(h, d_xref, nullptr, nullptr, d_X, d_U, d_status, d_iters, B, (cudaStream_t)stream, 0);
}

int mpcb200_solve_dual(mpcb200_handle* h, const double* d_xref, double* d_X, double* d_U, double* d_lam, int32_t* d_status, int32_t* d_iters,
                       int32_t B, void* stream) {
  if (!h) return -2;
  if (!d_lam) { h->err = "null dual block"; return -2; }
  if ((uintptr_t)d_lam & 7u) { h->err = "float64 arrays must be 8-byte aligned"; return -2; }
  DeviceGuard guard(h->cfg.device);
  return solve_io(h, d_xref, nullptr, nullptr, d_X, d_U, d_status, d_iters, B, (cudaStream_t)stream, 0, d_lam);
}

int32_t mpcb200_lam_words(const mpcb200_handle* h) { return h ? 14 * h->cfg.N + 2 : 0; }

int mpcb200_solve_cold(mpcb200_handle* h, const double* d_xref, double* d_X, double* d_U, int32_t* d_status, int32_t* d_iters,
                       int32_t B, void* stream) {
  if (!h) return -2;
  DeviceGuard guard(h->cfg.device);
  return solve_io(h, d_xref, nullptr, nullptr, d_X, d_U, d_status, d_iters, B, (cudaStream_t)stream, 1);
}

int mpcb200_sqp_begin(mpcb200_handle* h, const double* d_xref, const double* d_X, const double* d_U, int32_t B, void* stream) {
  if (!h) return -2;
  DeviceGuard guard(h->cfg.device);
  h->sw_xref = d_xref; h->sw_B = B;
  cudaStream_t s = (cudaStream_t)stream;
  if (h->cfg.precision == MPCB200_F64) return do_solve<double>(h, MODE_BEGIN, 0, d_xref, (double*)d_X, (double*)d_U, nullptr, nullptr, B, s);
  return do_solve<float>(h, MODE_BEGIN, 0, d_xref, (double*)d_X, (double*)d_U, nullptr, nullptr, B, s);
}

int mpcb200_sqp_iter(mpcb200_handle* h, int32_t n_iter, void* stream) {
  if (!h || !h->sw_xref) { if (h) h->err = "sqp_iter without sqp_begin"; return -2; }
  DeviceGuard guard(h->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  if (h->cfg.precision == MPCB200_F64) return do_solve<double>(h, MODE_ITER, n_iter, h->sw_xref, nullptr, nullptr, nullptr, nullptr, h->sw_B, s);
  return do_solve<float>(h, MODE_ITER, n_iter, h->sw_xref, nullptr, nullptr, nullptr, nullptr, h->sw_B, s);
}

int mpcb200_sqp_end(mpcb200_handle* h, double* d_X, double* d_U, int32_t* d_status, int32_t* d_iters, void* stream) {
  if (!h || !h->sw_xref) { if (h) h->err = "sqp_end without sqp_begin"; return -2; }
  DeviceGuard guard(h->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  int rc;
  if (h->cfg.precision == MPCB200_F64) rc = do_solve<double>(h, MODE_END, 0, h->sw_xref, d_X, d_U, d_status, d_iters, h->sw_B, s);
  else rc = do_solve<float>(h, MODE_END, 0, h->sw_xref, d_X, d_U, d_status, d_iters, h->sw_B, s);
  h->sw_xref = nullptr;
  return rc;
}

int mpcb200_plant_step_shift(mpcb200_handle* h, double* d_x, double* d_U, double* d_X, double* d_u_applied, int32_t B, void* stream) {
  if (!h) return -2;
  if (B <= 0) return 0;
  DeviceGuard guard(h->cfg.device);
  plant_step_shift_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_x, d_U, d_X, d_u_applied, B, h->cfg.N, h->cfg.dt, h->cfg.l_wb);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, "plant_step_shift launch", e);
  return 0;
}

int mpcb200_build_ref_window(mpcb200_handle* h, int32_t i, int32_t iter_length, const double* d_path, const double* d_orientation,
                             double desired_velocity, const double* d_x, double* d_xref, int32_t B, void* stream) {
  if (!h) return -2;
  if (B <= 0) return 0;
  if (h->cfg.N > iter_length) { h->err = "predict_horizon exceeds iter_length"; return -2; }
  DeviceGuard guard(h->cfg.device);
  build_ref_window_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(i, iter_length, d_path, d_orientation, desired_velocity, d_x,
                                                                             d_xref, B, h->cfg.N);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, "build_ref_window launch", e);
  return 0;
}

int mpcb200_closed_loop(mpcb200_handle* h, int32_t iter_length, const double* d_path, const double* d_orientation, double desired_velocity,
                        const double* d_x0, double* d_traj, double* d_ctrl, int32_t* d_status, int32_t* d_iters, int32_t B, void* stream) {
  if (!h) return -2;
  if (B <= 0) return 0;
  if (B > h->cfg.max_batch) { h->err = "B exceeds cfg.max_batch"; return -2; }
  if (h->cfg.N > iter_length) { h->err = "predict_horizon exceeds iter_length"; return -2; }
  DeviceGuard guard(h->cfg.device);
  const cudaError_t e = (h->cfg.precision == MPCB200_F64)
      ? closed_loop_t<double>(h, iter_length, d_path, d_orientation, desired_velocity, d_x0, d_traj, d_ctrl, d_status, d_iters, B, (cudaStream_t)stream)
      : closed_loop_t<float>(h, iter_length, d_path, d_orientation, desired_velocity, d_x0, d_traj, d_ctrl, d_status, d_iters, B, (cudaStream_t)stream);
  if (e != cudaSuccess) return fail(h, "mpc_closed_loop_kernel launch", e);
  return 0;
}

// Device-visible alias of a pinned host pointer (cudaHostAlloc / cudaHostRegister memory under unified addressing), or
// nullptr for pageable memory.
static void* mapped_alias(const void* p) {
  if (!p) return nullptr;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return (at.type == cudaMemoryTypeHost) ? at.devicePointer : nullptr;
}

int mpcb200_solve_host(mpcb200_handle* h, const double* h_xref, const double* h_X, const double* h_U, double* h_X_out, double* h_U_out,
                       int32_t* h_status, int32_t* h_iters, int32_t B) {
  if (!h) return -2;
  if (B <= 0) return 0;
  if (B > h->cfg.max_batch) { h->err = "B exceeds cfg.max_batch"; return -2; }
  if (!h_xref || !h_X_out || !h_U_out || (!h_X != !h_U)) { h->err = "null host buffer"; return -2; }
  DeviceGuard guard(h->cfg.device);
  const bool cold = !h_X;                               // no warm start: nothing but xref is uploaded
  const int N = h->cfg.N;
  const size_t nx = (size_t)5 * (N + 1), nu = (size_t)2 * N, mb = h->cfg.max_batch;
  if (!h->h_pin) {
    CK(cudaMallocHost(&h->h_pin, 2 * mb * 4));
    for (int i = 0; i < MPCB200_HOST_STREAMS; ++i) CK(cudaStreamCreateWithFlags(&h->hs[i], cudaStreamNonBlocking));
  }
  // ---- zero-copy route: every data buffer is pinned host memory the device can address.  ONE launch (two with the float64
  // refinement pass); the warps' TMA bulk copies read xref from, and their coalesced stores write X / U to, host memory
  // directly over PCIe, so problem b's transfer overlaps the other problems' iterations inside the kernel and no staging
  // copy or extra launch is on the critical path.
  if (h->cfg.host_route != 1) {
    const double* m_xref = (const double*)mapped_alias(h_xref);
    double* m_Xo = (double*)mapped_alias(h_X_out);
    double* m_Uo = (double*)mapped_alias(h_U_out);
    const double* m_Xi = cold ? nullptr : (const double*)mapped_alias(h_X);
    const double* m_Ui = cold ? nullptr : (const double*)mapped_alias(h_U);
    int* m_pin = (int*)mapped_alias(h->h_pin);
    if (m_xref && m_Xo && m_Uo && m_pin && (cold || (m_Xi && m_Ui))) {
      cudaStream_t s = h->hs[0];
      int rc = solve_io(h, m_xref, m_Xi, m_Ui, m_Xo, m_Uo, m_pin, m_pin + mb, B, s, cold ? 1 : 0);
      if (rc) return rc;
      CK(cudaStreamSynchronize(s));
      if (h_status) memcpy(h_status, h->h_pin, (size_t)B * 4);
      if (h_iters) memcpy(h_iters, h->h_pin + mb, (size_t)B * 4);
      return 0;
    }
  }
  // ---- staged route (pageable buffers): device staging arrays + chunked copy / solve / copy pipeline
  if (!h->d_xref) {
    CK(cudaMalloc(&h->d_xref, mb * nx * 8)); CK(cudaMalloc(&h->d_X, mb * nx * 8)); CK(cudaMalloc(&h->d_U, mb * nu * 8));
    CK(cudaMalloc(&h->d_status, mb * 4)); CK(cudaMalloc(&h->d_iters, mb * 4));
  }
  // Chunked pipeline over a few streams: the H2D copy of chunk c+1 and the D2H copy of chunk c-1 run under the solve
  // of chunk c (with pinned host buffers; pageable ones still work, the copies just serialise).  The work counters and the
  // refinement queue belong to ONE launch at a time, so a handle whose launches use them (refine_f64, or chunks larger than
  // the resident grid) runs its chunks on one stream.
  int nchunk = (B >= 4096) ? MPCB200_HOST_STREAMS : (B >= 512 ? 2 : 1);   // small batches are latency-bound: fewer, larger chunks
  if (h->cfg.host_chunks >= 1 && h->cfg.host_chunks <= MPCB200_HOST_STREAMS) nchunk = h->cfg.host_chunks;
  const int per = (B + nchunk - 1) / nchunk;
  const bool one_stream = (h->cfg.precision == MPCB200_F32 && h->cfg.refine_f64) || per > h->solve.max_ctas * h->solve.wpc;
  for (int c = 0, lo = 0; lo < B; ++c, lo += per) {
    const int n = (B - lo < per) ? (B - lo) : per;
    cudaStream_t s = h->hs[one_stream ? 0 : c % MPCB200_HOST_STREAMS];
    CK(cudaMemcpyAsync(h->d_xref + lo * nx, h_xref + lo * nx, n * nx * 8, cudaMemcpyHostToDevice, s));
    if (!cold) {
      CK(cudaMemcpyAsync(h->d_X + lo * nx, h_X + lo * nx, n * nx * 8, cudaMemcpyHostToDevice, s));
      CK(cudaMemcpyAsync(h->d_U + lo * nu, h_U + lo * nu, n * nu * 8, cudaMemcpyHostToDevice, s));
    }
    int rc = solve_io(h, h->d_xref + lo * nx, nullptr, nullptr, h->d_X + lo * nx, h->d_U + lo * nu, h->d_status + lo, h->d_iters + lo, n, s,
                      cold ? 1 : 0);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h_X_out + lo * nx, h->d_X + lo * nx, n * nx * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h_U_out + lo * nu, h->d_U + lo * nu, n * nu * 8, cudaMemcpyDeviceToHost, s));
    if (h_status) CK(cudaMemcpyAsync(h->h_pin + lo, h->d_status + lo, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    if (h_iters) CK(cudaMemcpyAsync(h->h_pin + mb + lo, h->d_iters + lo, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
  }
  for (int i = 0; i < MPCB200_HOST_STREAMS; ++i) CK(cudaStreamSynchronize(h->hs[i]));
  if (h_status) memcpy(h_status, h->h_pin, (size_t)B * 4);
  if (h_iters) memcpy(h_iters, h->h_pin + mb, (size_t)B * 4);
  return 0;
}

int mpcb200_set_scenarios(mpcb200_handle* h, const mpcb200_scenario* table, int32_t n) {
  if (!h) return -2;
  if (!table || n < 1 || n > 4096) { h->err = "scenario table: need 1 .. 4096 rows"; return -2; }
  if (h->cfg.hessian == MPCB200_HESS_EXACT) { h->err = "per-problem scenarios are built for the Gauss-Newton Hessian only"; return -2; }
  DeviceGuard guard(h->cfg.device);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->cfg.device));
  const int optin = (int)prop.sharedMemPerBlockOptin, sms = prop.multiProcessorCount;
  const bool f64 = h->cfg.precision == MPCB200_F64;
  h->scn.wpc = h->solve.wpc == 4 ? 2 : h->solve.wpc; h->scn.smem = smem_bytes_for(h->cfg.N, h->words, h->elem, h->scn.wpc);
  cudaError_t e = f64 ? plan_scn<double>(h->scn, optin, sms) : plan_scn<float>(h->scn, optin, sms);
  if (e != cudaSuccess) return fail(h, "kernel configuration (per-problem scenarios)", e);
  h->scn_refine.wpc = 0;
  if (!f64 && h->cfg.refine_f64) {
    h->scn_refine.wpc = 1; h->scn_refine.smem = smem_bytes_for(h->cfg.N, h->words, 8, 1);
    e = plan_scn<double>(h->scn_refine, optin, sms);
    if (e != cudaSuccess) return fail(h, "kernel configuration (per-problem scenarios, float64 refinement)", e);
    if (h->scn_refine.max_ctas > sms) h->scn_refine.max_ctas = sms;
  }
  cudaFree(h->scn_table); h->scn_table = nullptr;
  CK(cudaMalloc(&h->scn_table, (size_t)n * sizeof(mpcb200_scenario)));
  CK(cudaMemcpy(h->scn_table, table, (size_t)n * sizeof(mpcb200_scenario), cudaMemcpyHostToDevice));
  h->n_scn = n;
  return 0;
}

int mpcb200_solve_scenarios(mpcb200_handle* h, const double* d_xref, const int32_t* d_scenario_id, double* d_X, double* d_U,
                            int32_t* d_status, int32_t* d_iters, int32_t B, void* stream) {
  if (!h) return -2;
  if (B <= 0) return 0;
  if (!h->scn_table) { h->err = "mpcb200_solve_scenarios without mpcb200_set_scenarios"; return -2; }
  if (B > h->cfg.max_batch) { h->err = "B exceeds cfg.max_batch"; return -2; }
  if (!d_xref || !d_scenario_id || !d_X || !d_U) { h->err = "null argument"; return -2; }
  if (((uintptr_t)d_xref | (uintptr_t)d_X | (uintptr_t)d_U) & 7u) { h->err = "float64 arrays must be 8-byte aligned"; return -2; }
  DeviceGuard guard(h->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e;
  if (h->cfg.precision == MPCB200_F64) {
    ScnArgs<double> a; fill_scn_args(h, a, d_xref, d_scenario_id, d_X, d_U, d_status, d_iters, B);
    e = launch_scn<double>(h, a, s, h->scn, B);
  } else {
    const bool refine = h->cfg.refine_f64 && d_status && h->scn_refine.wpc;
    ScnArgs<float> a; fill_scn_args(h, a, d_xref, d_scenario_id, d_X, d_U, d_status, d_iters, B);
    a.q_list = refine ? h->q_list : nullptr; a.pdl_primary = refine ? 1 : 0;
    e = launch_scn<float>(h, a, s, h->scn, B);
    if (e == cudaSuccess && refine) {
      ScnArgs<double> r; fill_scn_args(h, r, d_xref, d_scenario_id, d_X, d_U, d_status, d_iters, B);
      r.q_list = h->q_list; r.refine = 1;
      e = launch_scn<double>(h, r, s, h->scn_refine, B);
    }
  }
  if (e != cudaSuccess) return fail(h, "mpc_warp_solve_scenarios_kernel launch", e);
  return 0;
}

int mpcb200_forces_stage_eval(mpcb200_handle* h, const double* weights_terminal, const double* d_z, const double* d_p, double* d_out, int32_t n,
                              void* stream) {
  if (!h) return -2;
  if (n <= 0) return 0;
  if (!weights_terminal || !d_z || !d_p || !d_out) { h->err = "null argument"; return -2; }
  DeviceGuard guard(h->cfg.device);
  ForcesConsts<double> C;
  C.dt = h->cfg.dt; C.l_wb = h->cfg.l_wb; C.l_fric = h->cfg.l_fric; C.ego_off = h->cfg.ego_offset;
  for (int i = 0; i < 5; ++i) { C.Q[i] = h->cfg.Q[i]; C.Pt[i] = weights_terminal[i]; }
  C.R[0] = h->cfg.R[0]; C.R[1] = h->cfg.R[1];
  forces_stage_eval_kernel<<<(n + 31) / 32, 32, 0, (cudaStream_t)stream>>>(C, d_z, d_p, d_out, n);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, "forces_stage_eval launch", e);
  return 0;
}

int64_t mpcb200_launch_count(const mpcb200_handle* h) { return h ? h->launches : 0; }
int32_t mpcb200_workspace_words(const mpcb200_handle* h) { return h ? h->words : 0; }
int32_t mpcb200_slab_in_smem(const mpcb200_handle* h) { return h ? h->solve.wpc : 0; }

}  // extern "C"
