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
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <string>
#include <new>

#include "config_params.h"
#include "warp_core.cuh"

using namespace mpcb200;

// ===================================================================================================== PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// TMA 1-D bulk copy shared -> global (bulk-group completion)
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ===================================================================================================== kernel args
enum : int { MODE_ONESHOT = 0, MODE_BEGIN = 1, MODE_ITER = 2, MODE_END = 3 };

template <typename T>
struct SolveArgs {
  ParamsT<T> P;
  double obstacle[6];
  const double* xref;   // [B][N+1][5]
  double* X;            // [B][N+1][5]  optimal states out
  double* U;            // [B][N][2]    optimal controls out
  const double* Xin;    // warm start in (may alias X / U)
  const double* Uin;
  int* status;          // [B]
  int* iters;           // [B]
  T* slab;              // global image of the slabs [B][words] (stepwise mode)
  ProbState<T>* state;  // [B] (stepwise mode)
  T* obs_shift;         // [B][6] shifted obstacle centres (stepwise mode)
  int B;
  int mode;
  int n_iter;
  int cold;             // 1: X / U are outputs only (cold start: X_0 tiled, zero controls)
  int refine;           // 1: second pass -- only instances whose status is not 1 are solved (warm start = their X / U)
};

// shared-memory carve-up of one CTA: [WPC slabs of T][float64 staging: xref | X | U for WPC problems]
template <typename T, int WPC>
struct Smem {
  int nx, nu;
  size_t slab_bytes;
  unsigned char* raw;
  __device__ Smem(unsigned char* raw_, int N, int words) : nx(5 * (N + 1)), nu(2 * N), slab_bytes((size_t)words * sizeof(T)), raw(raw_) {}
  __device__ T* slab(int w) const { return reinterpret_cast<T*>(raw + (size_t)w * slab_bytes); }
  __device__ double* xref(int w = 0) const { return reinterpret_cast<double*>(raw + (size_t)WPC * slab_bytes) + (size_t)w * nx; }
  __device__ double* X(int w = 0) const { return xref(0) + (size_t)WPC * nx + (size_t)w * nx; }
  __device__ double* U(int w = 0) const { return xref(0) + (size_t)2 * WPC * nx + (size_t)w * nu; }
};
static size_t smem_bytes_for(int N, int words, size_t elem, int wpc) {
  return (size_t)wpc * ((size_t)words * elem + (size_t)(12 * N + 10) * sizeof(double));
}

// CTA-cooperative tile copies HBM <-> float64 staging.  A full tile whose byte count is a multiple of 16 moves as TMA
// bulk copies (one elected thread, completion on an mbarrier / bulk group); a ragged last tile uses plain loops.
template <int WPC>
__device__ __forceinline__ bool tile_is_bulk(int nvalid, int per_problem) {
  return nvalid == WPC && ((WPC * per_problem * (int)sizeof(double)) % 16) == 0;
}

// ===================================================================================================== solve kernel
// One warp per ego instance; WPC warps (problems) per CTA.  All SQP iterations of a problem run inside the launch
// (MODE_ONESHOT), or `n_iter` of them with the slab round-tripping HBM <-> shared memory by TMA (stepwise modes).
template <typename T, int WPC, int HM>
__global__ void __launch_bounds__(32 * WPC, 16 / WPC) mpc_warp_solve_kernel(const __grid_constant__ SolveArgs<T> a) {
  unsigned char* const smem_raw = mpc_dyn_smem;
  __shared__ __align__(8) uint64_t bar_io;
  __shared__ __align__(8) uint64_t bar_w[WPC];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.P.N;
  const WLayout L(N, rec_stride_for(HM));
  const Smem<T, WPC> sm(smem_raw, N, L.words);
  const int nx = sm.nx, nu = sm.nu;
  const int base = blockIdx.x * WPC;
  const int nvalid = min(WPC, a.B - base);
  const bool valid = wid < nvalid;
  const int b = base + (valid ? wid : 0);

  if (threadIdx.x == 0) {
    mbar_init(&bar_io, 1);
#pragma unroll
    for (int i = 0; i < WPC; ++i) mbar_init(&bar_w[i], 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const WarpCtx w;
  T obs[6];
  WarpSolver<T, HM> S(a.P, SlabRef<T>{wid * L.words}, obs, w);
  ProbState<T> st;
  const bool need_xref = (a.mode != MODE_ITER);
  const bool need_init = (a.mode == MODE_ONESHOT || a.mode == MODE_BEGIN);
  const bool need_warm = need_init && !a.cold;

  // ---- problem data in: xref (+ warm start) of the CTA's tile, HBM -> staging
  if (need_xref) {
    if (tile_is_bulk<WPC>(nvalid, nx) && tile_is_bulk<WPC>(nvalid, nu)) {
      if (threadIdx.x == 0) {
        const uint32_t bx = (uint32_t)(WPC * nx * sizeof(double)), bu = (uint32_t)(WPC * nu * sizeof(double));
        mbar_expect_tx(&bar_io, need_warm ? (2 * bx + bu) : bx);
        tma_load_1d(sm.xref(), a.xref + (size_t)base * nx, bx, &bar_io);
        if (need_warm) {
          tma_load_1d(sm.X(), a.Xin + (size_t)base * nx, bx, &bar_io);
          tma_load_1d(sm.U(), a.Uin + (size_t)base * nu, bu, &bar_io);
        }
      }
      mbar_wait(&bar_io, 0);
    } else {
      for (int i = threadIdx.x; i < nvalid * nx; i += 32 * WPC) {
        sm.xref()[i] = a.xref[(size_t)base * nx + i];
        if (need_warm) sm.X()[i] = a.Xin[(size_t)base * nx + i];
      }
      if (need_warm) for (int i = threadIdx.x; i < nvalid * nu; i += 32 * WPC) sm.U()[i] = a.Uin[(size_t)base * nu + i];
      __syncthreads();
    }
  }

  // refinement pass: a problem that already converged keeps its result (its staging rows are written back untouched)
  const bool skip = a.refine && valid && a.status[b] == ST_OPTIMAL;
  if (need_init) {
    if (skip) { st.done = 1; st.status = ST_OPTIMAL; st.iters = a.iters ? a.iters[b] : 0; }
    else if (valid) { S.load(sm.xref(wid), need_warm ? sm.X(wid) : nullptr, need_warm ? sm.U(wid) : nullptr, a.obstacle, obs); S.init(st); }
    else { st.done = 1; st.status = ST_MAXIT; st.iters = 0; }
  } else {
    // resume: the slab image comes back by one TMA bulk copy per warp, the per-problem scalars by plain loads
    if (valid) {
      if (lane == 0) {
        const uint32_t bytes = (uint32_t)((size_t)L.words * sizeof(T));
        mbar_expect_tx(&bar_w[wid], bytes);
        tma_load_1d(sm.slab(wid), a.slab + (size_t)b * L.words, bytes, &bar_w[wid]);
      }
      mbar_wait(&bar_w[wid], 0);
      st = a.state[b];
#pragma unroll
      for (int j = 0; j < 6; ++j) obs[j] = a.obs_shift[(size_t)b * 6 + j];
    } else { st.done = 1; st.status = ST_MAXIT; st.iters = 0; }
  }

  if (a.mode != MODE_END && valid && !skip) {
    for (int it = 0; it < a.n_iter && !st.done; ++it) S.iterate(st);
  }

  if (a.mode == MODE_ONESHOT || a.mode == MODE_END) {
    // solution back to float64 row-major (rho is added back in float64), staging -> HBM
    if (valid && !skip) {
      S.store(sm.xref(wid), sm.X(wid), sm.U(wid));
      if (lane == 0) {
        if (a.status) a.status[b] = st.status;
        if (a.iters) a.iters[b] = st.iters + (a.refine && a.iters ? a.iters[b] : 0);
      }
    }
    if (tile_is_bulk<WPC>(nvalid, nx) && tile_is_bulk<WPC>(nvalid, nu)) {
      fence_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        tma_store_1d(a.X + (size_t)base * nx, sm.X(), (uint32_t)(WPC * nx * sizeof(double)));
        tma_store_1d(a.U + (size_t)base * nu, sm.U(), (uint32_t)(WPC * nu * sizeof(double)));
        tma_store_commit_wait();
      }
    } else {
      __syncthreads();
      for (int i = threadIdx.x; i < nvalid * nx; i += 32 * WPC) a.X[(size_t)base * nx + i] = sm.X()[i];
      for (int i = threadIdx.x; i < nvalid * nu; i += 32 * WPC) a.U[(size_t)base * nu + i] = sm.U()[i];
    }
  } else if (valid) {
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
  }
}

// ===================================================================================================== closed loop
template <typename T>
struct LoopArgs {
  ParamsT<T> P;
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
};

__device__ __forceinline__ void plant_euler(double* x, double u0, double u1, double dt, double l_wb) {
  // shift_movement: st = x0 + delta_t * f(x0, u[:,0])   (optimizer.py:649-650; KS model configuration.py:364-368)
  double s, c;
  sincos(x[4], &s, &c);
  const double v = x[3], tn = tan(x[2]);
  x[0] += dt * v * c; x[1] += dt * v * s; x[2] += dt * u0; x[3] += dt * u1; x[4] += dt * v / l_wb * tn;
}

// row k+1 of the X_ref block of MPC step i (desired_command_and_trajectory, optimizer.py:657-702, quirk Q8)
__device__ __forceinline__ void ref_window_row(int i, int k, int N, int Tlen, const double* path, const double* orient, double vdes, double* r) {
  const int idx = (i >= Tlen - N) ? (k + (Tlen - N)) : (i + k + 1);
  r[0] = path[2 * idx]; r[1] = path[2 * idx + 1]; r[2] = 0.0; r[3] = vdes; r[4] = orient[idx];
}
__device__ __forceinline__ void ref_window_rows(int i, int N, int Tlen, const double* path, const double* orient, double vdes,
                                                const double* x_now, double* xref /* [N+1][5] */) {
  for (int j = 0; j < 5; ++j) xref[j] = x_now[j];
  for (int k = 0; k < N; ++k) ref_window_row(i, k, N, Tlen, path, orient, vdes, xref + 5 * (k + 1));
}

// The whole receding-horizon loop of CasadiOptimizer.optimize() (optimizer.py:596-631) for one ego per warp, no host
// round trip between MPC steps: solve, record u_0, plant step + warm-start shift, next reference window.
template <typename T, int WPC, int HM>
__global__ void __launch_bounds__(32 * WPC) mpc_warp_closed_loop_kernel(const __grid_constant__ LoopArgs<T> a) {
  unsigned char* const smem_raw = mpc_dyn_smem;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.P.N;
  const WLayout L(N, rec_stride_for(HM));
  const Smem<T, WPC> sm(smem_raw, N, L.words);
  const int nu = sm.nu;
  const int b = blockIdx.x * WPC + wid;
  if (b >= a.B) return;                                   // whole warp leaves; no CTA-wide barrier below
  double* my_xref = sm.xref(wid);
  double* my_X = sm.X(wid);
  double* my_U = sm.U(wid);
  const WarpCtx w;
  T obs[6];
  WarpSolver<T, HM> S(a.P, SlabRef<T>{wid * L.words}, obs, w);
  ProbState<T> st;
  double x[5];
  for (int j = 0; j < 5; ++j) x[j] = a.x0[(size_t)b * 5 + j];
  // first parameter block and warm start: the initial state tiled, controls zero (optimizer.py:578-583, quirk Q4)
  for (int k = lane; k <= N; k += 32)
    for (int j = 0; j < 5; ++j) { my_xref[5 * k + j] = x[j]; my_X[5 * k + j] = x[j]; }
  for (int k = lane; k < nu; k += 32) my_U[k] = 0.0;
  __syncwarp();
  for (int i = 0; i < a.Tlen; ++i) {
    if (lane == 0 && a.traj) for (int j = 0; j < 5; ++j) a.traj[((size_t)b * a.Tlen + i) * 5 + j] = x[j];   // quirk Q12
    S.load(my_xref, my_X, my_U, a.obstacle, obs);
    S.init(st);
    for (int it = 0; it < a.P.max_iter && !st.done; ++it) S.iterate(st);
    S.store(my_xref, my_X, my_U);
    const double u0 = my_U[0], u1 = my_U[1];
    if (lane == 0) {
      if (a.ctrl) { a.ctrl[((size_t)b * a.Tlen + i) * 2] = u0; a.ctrl[((size_t)b * a.Tlen + i) * 2 + 1] = u1; }
      if (a.status) a.status[(size_t)b * a.Tlen + i] = st.status;
      if (a.iters) a.iters[(size_t)b * a.Tlen + i] = st.iters;
    }
    plant_euler(x, u0, u1, a.dt, a.l_wb);
    __syncwarp();
    // shift the warm start one stage, repeating the last (optimizer.py:652-653); lane-strided with a register hop
    for (int k0 = 0; k0 < N; k0 += 32) {
      const int k = k0 + lane;
      double nxt[7];
      if (k < N) {
        const int ks = (k + 1 < N) ? k + 1 : N - 1;
        nxt[5] = my_U[2 * ks]; nxt[6] = my_U[2 * ks + 1];
        for (int j = 0; j < 5; ++j) nxt[j] = my_X[5 * (k + 1) + j];
      }
      __syncwarp();
      if (k < N) {
        my_U[2 * k] = nxt[5]; my_U[2 * k + 1] = nxt[6];
        for (int j = 0; j < 5; ++j) my_X[5 * k + j] = nxt[j];
      }
      __syncwarp();
    }
    // next window from the new state (optimizer.py:628)
    if (lane == 0) for (int j = 0; j < 5; ++j) my_xref[j] = x[j];
    for (int k = lane; k < N; k += 32) ref_window_row(i, k, N, a.Tlen, a.path, a.orient, a.desired_velocity, my_xref + 5 * (k + 1));
    __syncwarp();
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

// ===================================================================================================== handle
#define MPCB200_HOST_STREAMS 4
struct mpcb200_handle {
  mpcb200_config cfg;
  int wpc;              // warps (= problems) per CTA
  size_t smem_bytes;
  int wpc64;            // same for the float64 refinement pass of a float32 handle (cfg.refine_f64)
  size_t smem64;
  int words;            // slab words per problem
  void* slab;           // global slab image [max_batch][words] (stepwise mode), allocated on first use
  void* state;
  void* obs_shift;
  size_t elem;          // sizeof(T)
  int64_t launches;
  // stepwise-mode context
  const double* sw_xref;
  int sw_B;
  // host-path staging
  double *d_xref, *d_X, *d_U;
  int *d_status, *d_iters;
  cudaStream_t hs[MPCB200_HOST_STREAMS];
  int* h_pin;           // pinned host staging for status + iters (a pageable D2H target would serialise the chunk pipeline)
  std::string err;
};

static thread_local std::string g_create_err;

static int fail(mpcb200_handle* h, const char* what, cudaError_t e) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s: %s", what, e == cudaSuccess ? "" : cudaGetErrorString(e));
  if (h) h->err = buf; else g_create_err = buf;
  return -1;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(h, #call, e_); } while (0)

template <typename T, int WPC, int HM>
static cudaError_t launch_solve(mpcb200_handle* h, const SolveArgs<T>& a, cudaStream_t s, size_t smem) {
  const int ctas = (a.B + WPC - 1) / WPC;
  mpc_warp_solve_kernel<T, WPC, HM><<<ctas, 32 * WPC, smem, s>>>(a);
  h->launches++;
  return cudaGetLastError();
}
template <typename T, int WPC, int HM>
static cudaError_t launch_loop(mpcb200_handle* h, const LoopArgs<T>& a, cudaStream_t s) {
  const int ctas = (a.B + WPC - 1) / WPC;
  mpc_warp_closed_loop_kernel<T, WPC, HM><<<ctas, 32 * WPC, h->smem_bytes, s>>>(a);
  h->launches++;
  return cudaGetLastError();
}

// kernel instantiations: arithmetic type x problems per CTA (2, or 1 when the slab of a long horizon leaves no room for
// two) x Hessian mode
template <typename T>
static cudaError_t dispatch_solve(mpcb200_handle* h, SolveArgs<T>& a, cudaStream_t s, int wpc, size_t smem) {
  const bool ex = a.P.hessian == HESS_EXACT;
  if (wpc == 2) return ex ? launch_solve<T, 2, HESS_EXACT>(h, a, s, smem) : launch_solve<T, 2, HESS_GN>(h, a, s, smem);
  return ex ? launch_solve<T, 1, HESS_EXACT>(h, a, s, smem) : launch_solve<T, 1, HESS_GN>(h, a, s, smem);
}
template <typename T>
static cudaError_t dispatch_loop(mpcb200_handle* h, LoopArgs<T>& a, cudaStream_t s) {
  const bool ex = a.P.hessian == HESS_EXACT;
  if (h->wpc == 2) return ex ? launch_loop<T, 2, HESS_EXACT>(h, a, s) : launch_loop<T, 2, HESS_GN>(h, a, s);
  return ex ? launch_loop<T, 1, HESS_EXACT>(h, a, s) : launch_loop<T, 1, HESS_GN>(h, a, s);
}

// opt the handle's kernel instantiations in to their dynamic shared memory size (once, at create)
template <typename T, int WPC>
static cudaError_t configure_kernels_t(size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(mpc_warp_solve_kernel<T, WPC, HESS_GN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mpc_warp_solve_kernel<T, WPC, HESS_EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mpc_warp_closed_loop_kernel<T, WPC, HESS_GN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(mpc_warp_closed_loop_kernel<T, WPC, HESS_EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  return e;
}
template <typename T>
static cudaError_t configure_kernels(int wpc, size_t smem) {
  if (wpc == 2) return configure_kernels_t<T, 2>(smem);
  return configure_kernels_t<T, 1>(smem);
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
                    int B, cudaStream_t s, int cold = 0, const double* Xin = nullptr, const double* Uin = nullptr) {
  if (B <= 0) return 0;
  if (B > h->cfg.max_batch) { h->err = "B exceeds cfg.max_batch"; return -2; }
  if (mode != MODE_ONESHOT) { int rc = ensure_stepwise_scratch(h); if (rc) return rc; }
  SolveArgs<T> a;
  a.P = params_from_config<T>(h->cfg);
  for (int i = 0; i < 6; ++i) a.obstacle[i] = h->cfg.obstacle[i];
  a.xref = xref; a.X = X; a.U = U; a.status = status; a.iters = iters;
  a.Xin = Xin ? Xin : X; a.Uin = Uin ? Uin : U;
  a.slab = (T*)h->slab; a.state = (ProbState<T>*)h->state; a.obs_shift = (T*)h->obs_shift;
  a.B = B; a.mode = mode; a.n_iter = n_iter; a.cold = cold; a.refine = 0;
  cudaError_t e = dispatch_solve<T>(h, a, s, h->wpc, h->smem_bytes);
  if (e != cudaSuccess) return fail(h, "mpc_warp_solve_kernel launch", e);
  return 0;
}

// second pass of a float32 handle with cfg.refine_f64: float64 arithmetic (float32 tolerances) on the instances that did
// not reach status 1, warm-started from their float32 result
static int refine_pass(mpcb200_handle* h, const double* xref, double* X, double* U, int* status, int* iters, int B, cudaStream_t s) {
  if (!status) { h->err = "refine_f64 needs a status buffer"; return -2; }
  SolveArgs<double> a;
  a.P = params_from_config<double>(h->cfg);
  for (int i = 0; i < 6; ++i) a.obstacle[i] = h->cfg.obstacle[i];
  a.xref = xref; a.X = X; a.U = U; a.status = status; a.iters = iters;
  a.Xin = X; a.Uin = U;
  a.slab = nullptr; a.state = nullptr; a.obs_shift = nullptr;
  a.B = B; a.mode = MODE_ONESHOT; a.n_iter = h->cfg.max_iter; a.cold = 0; a.refine = 1;
  cudaError_t e = dispatch_solve<double>(h, a, s, h->wpc64, h->smem64);
  if (e != cudaSuccess) return fail(h, "mpc_warp_solve_kernel<double> (refinement) launch", e);
  return 0;
}

extern "C" {

int32_t mpcb200_abi_version(void) { return MPCB200_ABI_VERSION; }

void mpcb200_default_config(mpcb200_config* cfg, int32_t N, int32_t precision) { default_config(cfg, N, precision); }

const char* mpcb200_last_error(const mpcb200_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int mpcb200_create(const mpcb200_config* cfg, mpcb200_handle** out) {
  mpcb200_handle* h = nullptr;
  if (!cfg || !out) { g_create_err = "null argument"; return -2; }
  if (cfg->abi_version != MPCB200_ABI_VERSION) { g_create_err = "abi_version mismatch"; return -2; }
  if (cfg->N < 4 || cfg->N > 512 || cfg->max_batch < 1) { g_create_err = "N must be in [4, 512], max_batch >= 1"; return -2; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { fail(nullptr, "no CUDA device (libmpcb200 has no CPU path)", e); return -3; }
  CK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 9) { g_create_err = "libmpcb200 needs TMA bulk copies (sm_90+); built for sm_100a"; return -3; }
  h = new (std::nothrow) mpcb200_handle();
  if (!h) { g_create_err = "out of host memory"; return -4; }
  h->cfg = *cfg;
  h->launches = 0; h->slab = h->state = h->obs_shift = nullptr;
  h->d_xref = h->d_X = h->d_U = nullptr; h->d_status = h->d_iters = nullptr; h->h_pin = nullptr;
  h->sw_xref = nullptr; h->sw_B = 0;
  const WLayout L(cfg->N, rec_stride_for(cfg->hessian == MPCB200_HESS_EXACT ? HESS_EXACT : HESS_GN));
  h->words = L.words;
  h->elem = cfg->precision == MPCB200_F64 ? 8 : 4;
  const size_t smem_max = prop.sharedMemPerBlockOptin;   // 227 KB on B200
  const size_t reserve = 1024;
  h->wpc = 0;
  int wpc_pref = 2;
  if (const char* ev = getenv("MPCB200_WPC")) { const int v = atoi(ev); if (v == 1 || v == 2) wpc_pref = v; }   // tuning knob
  for (int wpc : {wpc_pref, 2, 1}) {
    const size_t need = smem_bytes_for(cfg->N, L.words, h->elem, wpc);
    if (need + reserve <= smem_max) { h->wpc = wpc; h->smem_bytes = need; break; }
  }
  if (!h->wpc) { g_create_err = "horizon too long: the per-problem KKT slab does not fit shared memory"; delete h; return -2; }
  e = (cfg->precision == MPCB200_F64) ? configure_kernels<double>(h->wpc, h->smem_bytes) : configure_kernels<float>(h->wpc, h->smem_bytes);
  if (e != cudaSuccess) { fail(nullptr, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", e); delete h; return -1; }
  h->wpc64 = 0; h->smem64 = 0;
  if (cfg->precision == MPCB200_F32 && cfg->refine_f64) {
    for (int wpc : {2, 1}) {
      const size_t need = smem_bytes_for(cfg->N, L.words, 8, wpc);
      if (need + reserve <= smem_max) { h->wpc64 = wpc; h->smem64 = need; break; }
    }
    if (!h->wpc64) { g_create_err = "horizon too long for the float64 refinement pass"; delete h; return -2; }
    e = configure_kernels<double>(h->wpc64, h->smem64);
    if (e != cudaSuccess) { fail(nullptr, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", e); delete h; return -1; }
  }
  *out = h;
  return 0;
}

void mpcb200_destroy(mpcb200_handle* h) {
  if (!h) return;
  cudaFree(h->slab); cudaFree(h->state); cudaFree(h->obs_shift);
  if (h->h_pin) for (int i = 0; i < MPCB200_HOST_STREAMS; ++i) cudaStreamDestroy(h->hs[i]);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  cudaFree(h->d_xref); cudaFree(h->d_X); cudaFree(h->d_U); cudaFree(h->d_status); cudaFree(h->d_iters);
  delete h;
}

int mpcb200_solve(mpcb200_handle* h, const double* d_xref, double* d_X, double* d_U, int32_t* d_status, int32_t* d_iters,
                  int32_t B, void* stream) {
  if (!h) return -2;
  cudaStream_t s = (cudaStream_t)stream;
  if (h->cfg.precision == MPCB200_F64) return do_solve<double>(h, MODE_ONESHOT, h->cfg.max_iter, d_xref, d_X, d_U, d_status, d_iters, B, s);
  int rc = do_solve<float>(h, MODE_ONESHOT, h->cfg.max_iter, d_xref, d_X, d_U, d_status, d_iters, B, s);
  if (rc == 0 && h->cfg.refine_f64 && B > 0) rc = refine_pass(h, d_xref, d_X, d_U, d_status, d_iters, B, s);
  return rc;
}

int mpcb200_solve_cold(mpcb200_handle* h, const double* d_xref, double* d_X, double* d_U, int32_t* d_status, int32_t* d_iters,
                       int32_t B, void* stream) {
  if (!h) return -2;
  cudaStream_t s = (cudaStream_t)stream;
  if (h->cfg.precision == MPCB200_F64) return do_solve<double>(h, MODE_ONESHOT, h->cfg.max_iter, d_xref, d_X, d_U, d_status, d_iters, B, s, 1);
  int rc = do_solve<float>(h, MODE_ONESHOT, h->cfg.max_iter, d_xref, d_X, d_U, d_status, d_iters, B, s, 1);
  if (rc == 0 && h->cfg.refine_f64 && B > 0) rc = refine_pass(h, d_xref, d_X, d_U, d_status, d_iters, B, s);
  return rc;
}

int mpcb200_sqp_begin(mpcb200_handle* h, const double* d_xref, const double* d_X, const double* d_U, int32_t B, void* stream) {
  if (!h) return -2;
  h->sw_xref = d_xref; h->sw_B = B;
  cudaStream_t s = (cudaStream_t)stream;
  if (h->cfg.precision == MPCB200_F64) return do_solve<double>(h, MODE_BEGIN, 0, d_xref, (double*)d_X, (double*)d_U, nullptr, nullptr, B, s);
  return do_solve<float>(h, MODE_BEGIN, 0, d_xref, (double*)d_X, (double*)d_U, nullptr, nullptr, B, s);
}

int mpcb200_sqp_iter(mpcb200_handle* h, int32_t n_iter, void* stream) {
  if (!h || !h->sw_xref) { if (h) h->err = "sqp_iter without sqp_begin"; return -2; }
  cudaStream_t s = (cudaStream_t)stream;
  if (h->cfg.precision == MPCB200_F64) return do_solve<double>(h, MODE_ITER, n_iter, h->sw_xref, nullptr, nullptr, nullptr, nullptr, h->sw_B, s);
  return do_solve<float>(h, MODE_ITER, n_iter, h->sw_xref, nullptr, nullptr, nullptr, nullptr, h->sw_B, s);
}

int mpcb200_sqp_end(mpcb200_handle* h, double* d_X, double* d_U, int32_t* d_status, int32_t* d_iters, void* stream) {
  if (!h || !h->sw_xref) { if (h) h->err = "sqp_end without sqp_begin"; return -2; }
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
  cudaError_t e;
  if (h->cfg.precision == MPCB200_F64) {
    LoopArgs<double> a;
    a.P = params_from_config<double>(h->cfg);
    for (int i = 0; i < 6; ++i) a.obstacle[i] = h->cfg.obstacle[i];
    a.path = d_path; a.orient = d_orientation; a.x0 = d_x0; a.traj = d_traj; a.ctrl = d_ctrl; a.status = d_status; a.iters = d_iters;
    a.desired_velocity = desired_velocity; a.l_wb = h->cfg.l_wb; a.dt = h->cfg.dt; a.B = B; a.Tlen = iter_length;
    e = dispatch_loop<double>(h, a, (cudaStream_t)stream);
  } else {
    LoopArgs<float> a;
    a.P = params_from_config<float>(h->cfg);
    for (int i = 0; i < 6; ++i) a.obstacle[i] = h->cfg.obstacle[i];
    a.path = d_path; a.orient = d_orientation; a.x0 = d_x0; a.traj = d_traj; a.ctrl = d_ctrl; a.status = d_status; a.iters = d_iters;
    a.desired_velocity = desired_velocity; a.l_wb = h->cfg.l_wb; a.dt = h->cfg.dt; a.B = B; a.Tlen = iter_length;
    e = dispatch_loop<float>(h, a, (cudaStream_t)stream);
  }
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

// one solve (+ optional float64 refinement) with separate warm-start-in and result-out arrays
static int solve_io(mpcb200_handle* h, const double* xref, const double* Xin, const double* Uin, double* X, double* U,
                    int32_t* status, int32_t* iters, int32_t B, cudaStream_t s, int cold) {
  if (h->cfg.precision == MPCB200_F64) return do_solve<double>(h, MODE_ONESHOT, h->cfg.max_iter, xref, X, U, status, iters, B, s, cold, Xin, Uin);
  int rc = do_solve<float>(h, MODE_ONESHOT, h->cfg.max_iter, xref, X, U, status, iters, B, s, cold, Xin, Uin);
  if (rc == 0 && h->cfg.refine_f64 && B > 0) rc = refine_pass(h, xref, X, U, status, iters, B, s);
  return rc;
}

int mpcb200_solve_host(mpcb200_handle* h, const double* h_xref, const double* h_X, const double* h_U, double* h_X_out, double* h_U_out,
                       int32_t* h_status, int32_t* h_iters, int32_t B) {
  if (!h) return -2;
  if (B <= 0) return 0;
  if (B > h->cfg.max_batch) { h->err = "B exceeds cfg.max_batch"; return -2; }
  if (!h_xref || !h_X_out || !h_U_out || (!h_X != !h_U)) { h->err = "null host buffer"; return -2; }
  const bool cold = !h_X;                               // no warm start: nothing but xref is uploaded
  const int N = h->cfg.N;
  const size_t nx = (size_t)5 * (N + 1), nu = (size_t)2 * N, mb = h->cfg.max_batch;
  if (!h->h_pin) {
    CK(cudaMallocHost(&h->h_pin, 2 * mb * 4));
    for (int i = 0; i < MPCB200_HOST_STREAMS; ++i) CK(cudaStreamCreateWithFlags(&h->hs[i], cudaStreamNonBlocking));
  }
  // ---- zero-copy route: every data buffer is pinned host memory the device can address.  ONE launch; the kernel's TMA
  // bulk copies read xref (+ warm start) from and write X / U to host memory directly over PCIe, so problem b's transfer
  // overlaps the other problems' iterations inside the kernel and no staging copy or extra launch is on the critical path.
  const char* staged_env = getenv("MPCB200_HOST_STAGED");   // tuning knob: force the staged pipeline below
  if (!(staged_env && atoi(staged_env))) {
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
  // of chunk c (with pinned host buffers; pageable ones still work, the copies just serialise).  Chunks are even-sized
  // so every CTA keeps a full 2-problem tile.
  int nchunk = (B >= 4096) ? MPCB200_HOST_STREAMS : (B >= 512 ? 2 : 1);   // small batches are latency-bound: fewer, larger chunks
  if (const char* ev = getenv("MPCB200_HOST_CHUNKS")) { const int v = atoi(ev); if (v >= 1 && v <= MPCB200_HOST_STREAMS) nchunk = v; }   // tuning knob
  int per = ((B + nchunk - 1) / nchunk + 1) & ~1;
  for (int c = 0, lo = 0; lo < B; ++c, lo += per) {
    const int n = (B - lo < per) ? (B - lo) : per;
    cudaStream_t s = h->hs[c % MPCB200_HOST_STREAMS];
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

int64_t mpcb200_launch_count(const mpcb200_handle* h) { return h ? h->launches : 0; }
int32_t mpcb200_workspace_words(const mpcb200_handle* h) { return h ? h->words : 0; }
int32_t mpcb200_slab_in_smem(const mpcb200_handle* h) { return h ? h->wpc : 0; }

}  // extern "C"
