// aligned_solver.cu -- PHASE-ALIGNED variant of the fused solve kernel (float32, Gauss-Newton), selected with cfg.warps_per_cta = 8 | 16.
//
// Why: at large batches the per-instruction stall samples of mpc_warp_solve_kernel (profiles/r02_summary.md section 12) put
// `no_instruction` first in every straight-line phase (linearisation 43 %, trial merit 39 %, commit 36 %, step statistics 30 % of
// their samples), clustered on 128-byte line boundaries: the four warps of a sub-partition are in four different phases of four
// different problems and evict each other's instruction lines, so the ~2 000 straight-line instructions of an iteration are
// fetched from the SM-level cache again every time.  Here one CTA owns ALL resident warps of an SM (8 or 16 problems) and the
// warps meet at a CTA barrier once per SQP iteration, so the warps that share a sub-partition walk through the same text together.
// Everything else is the original kernel: one warp per problem, slab in shared memory, per-warp TMA input, dynamic work claiming
// (a warp whose problem has converged stores, claims the next one, loads and joins the next round).
#include "mpcb200_internal.cuh"

template <int WPC>
__global__ void __launch_bounds__(32 * WPC, 16 / WPC) mpc_warp_solve_aligned_kernel(const __grid_constant__ SolveArgs<float> a) {
  using T = float;
  unsigned char* const smem_raw = mpc_dyn_smem;
  __shared__ __align__(8) uint64_t bar_x[WPC];
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
  const WarpCtx w;
  T obs[6];
  WarpSolver<T, HESS_GN> S(a.P, SlabRef<T>{wid * L.words}, obs, w);
  const bool warm = !a.cold;
  uint32_t ph_x = 0;
  int item = blockIdx.x * WPC + wid;
  bool more = item < a.B, active = false;
  int b = -1, nit = 0;
  const double* xr = nullptr;
  ProbState<T> st;
  for (;;) {
    if (!active && more) {                         // claim: load + initialise, then join the round
      b = item;
      xr = fetch_xref(a.xref + (size_t)b * nx, sm.xstg(wid), nx, &bar_x[wid], ph_x, lane);
      S.load(xr, warm ? a.Xin + (size_t)b * nx : nullptr, warm ? a.Uin + (size_t)b * nu : nullptr, a.obstacle, obs);
      S.init(st);
      active = true; nit = 0;
    }
    if (active) {
      if (!st.done) { S.iterate(st); ++nit; }
      if (st.done || nit >= a.n_iter) {
        S.store(xr, a.X + (size_t)b * nx, a.U + (size_t)b * nu);
        if (lane == 0) {
          if (a.status) a.status[b] = st.status;
          if (a.iters) a.iters[b] = st.iters;
          if (a.q_list && st.status != ST_OPTIMAL && st.status != ST_INFEASIBLE_X0) a.q_list[atomicAdd(&a.ctr->q_count, 1)] = b;
        }
        active = false;
        if (a.dynamic) {
          int nxt = 0;
          if (lane == 0) nxt = total_warps + atomicAdd(&a.ctr->next, 1);
          item = __shfl_sync(0xffffffffu, nxt, 0);
          more = item < a.B;
        } else more = false;
      }
    }
    if (!__syncthreads_or((active || more) ? 1 : 0)) break;       // the per-iteration meeting point (and the exit test)
  }
  if (a.dynamic && lane == 0) {
    __threadfence();
    if (atomicAdd(&a.ctr->done, 1) == total_warps - 1) { a.ctr->next = 0; a.ctr->done = 0; }
  }
}

template <int WPC>
static cudaError_t launch_aligned_t(mpcb200_handle* h, SolveArgs<float>& a, cudaStream_t s, int nwork) {
  static int max_ctas = 0;      // per process and instantiation: occupancy is a property of the function on the device
  const size_t smem = smem_bytes_for(h->cfg.N, h->words, 4, WPC);
  if (!max_ctas) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, h->cfg.device);
    if (e != cudaSuccess) return e;
    if (smem + 1024 > prop.sharedMemPerBlockOptin) return cudaErrorInvalidConfiguration;
    e = plan_kernel(mpc_warp_solve_aligned_kernel<WPC>, WPC, smem, (int)prop.sharedMemPerBlockOptin, prop.multiProcessorCount, &max_ctas);
    if (e != cudaSuccess) return e;
  }
  const int want = (nwork + WPC - 1) / WPC;
  const int ctas = want < max_ctas ? want : max_ctas;
  a.dynamic = (nwork > ctas * WPC) ? 1 : 0;
  h->launches++;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(32 * WPC); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  return cudaLaunchKernelEx(&cfg, mpc_warp_solve_aligned_kernel<WPC>, a);
}

cudaError_t launch_solve_aligned(mpcb200_handle* h, SolveArgs<float>& a, cudaStream_t s, int nwork) {
  return h->cfg.warps_per_cta == 16 ? launch_aligned_t<16>(h, a, s, nwork) : launch_aligned_t<8>(h, a, s, nwork);
}
