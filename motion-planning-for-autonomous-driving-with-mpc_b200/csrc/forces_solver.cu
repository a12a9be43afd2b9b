// forces_solver.cu -- the FORCESPRO-formulation solve kernel and its C-ABI entry point (include/mpcb200.h, mpcb200_forces_solve).
// A translation unit of its own: the solver core (forces_core.cuh) is independent of the CasADi-formulation kernels in mpcb200.cu.
#include "mpcb200_internal.cuh"
#include "forces_core.cuh"

// ===================================================================================================== FORCESPRO formulation
// One warp per ego instance on the reference's FORCESPRO formulation of the MPC problem (forces_core.cuh), persistent warps with
// dynamic work claiming like the CasADi-formulation kernel.  The problem's parameter block ([N][10] float64 = 80 N bytes, always a
// multiple of 16) comes HBM -> shared memory by ONE TMA bulk copy per problem when its address is 16-byte aligned (plain loads
// otherwise); the KKT slab (164 N words) lives in shared memory for the whole solve; Z goes back lane = stage.
template <typename T>
struct ForcesArgs {
  FParams<T> fp;
  const double* xinit;   // [B][5]
  const double* par;     // [B][N][10]
  const double* Zin;     // [B][N][7] warm start or null
  double* Z;             // [B][N][7]
  int* status; int* iters;
  WorkCtr* ctr; int* q_list;
  RoadBounds<T> rb;      // road-boundary vertex lists (kernels instantiated with RB = true)
  int B, refine, dynamic, pdl_primary;
};
static size_t forces_smem_bytes_for(int N, int words, size_t elem, int wpc) {
  return (size_t)wpc * ((size_t)words * elem + (size_t)10 * N * sizeof(double));
}

template <typename T, int WPC, bool RB>
__global__ void __launch_bounds__(32 * WPC) mpc_forces_solve_kernel(const __grid_constant__ ForcesArgs<T> a) {
  unsigned char* const smem_raw = mpc_dyn_smem;
  __shared__ __align__(8) uint64_t bar_p[WPC];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.fp.P.N;
  const FLayout L(N);
  const int total_warps = gridDim.x * WPC;
  double* const stg = reinterpret_cast<double*>(smem_raw + (size_t)WPC * L.words * sizeof(T)) + (size_t)wid * 10 * N;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < WPC; ++i) mbar_init(&bar_p[i], 1);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (a.pdl_primary) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (a.refine) asm volatile("griddepcontrol.wait;" ::: "memory");
  const WarpCtx w;
  ForcesSolver<T, RB> S(a.fp, SlabRef<T>{wid * L.words}, w, a.rb);
  const int nwork = a.refine ? a.ctr->q_count : a.B;
  uint32_t ph = 0;
  for (int item = blockIdx.x * WPC + wid; item < nwork;) {
    const int b = a.refine ? a.q_list[item] : item;
    const double* gpar = a.par + (size_t)b * 10 * N;
    const double* par = gpar;
    if ((((uintptr_t)gpar) & 15u) == 0) {
      fence_async_smem();                        // the staging was last read through the generic proxy (previous problem)
      __syncwarp();
      if (lane == 0) {
        mbar_expect_tx(&bar_p[wid], (uint32_t)(80 * N));
        tma_load_1d(stg, gpar, (uint32_t)(80 * N), &bar_p[wid]);
      }
      mbar_wait(&bar_p[wid], ph);
      ph ^= 1u;
      __syncwarp();
      par = stg;
    }
    const double* xin = a.xinit + (size_t)b * 5;
    // refinement pass: warm start = the float32 result
    const double* zin = a.refine ? (a.Z + (size_t)b * 7 * N) : (a.Zin ? a.Zin + (size_t)b * 7 * N : nullptr);
    ProbState<T> st;
    S.load(xin, par, zin);
    S.init(st);
    for (int it = 0; it < a.fp.P.max_iter && !st.done; ++it) S.iterate(st);
    S.store(xin, par, a.Z + (size_t)b * 7 * N);
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

template <typename T>
static cudaError_t plan_forces(KernelPlan& k, int optin, int sms) {
  // one grid shape for the kernels with and without the road-boundary rows
  int c0 = 0, c1 = 0;
  cudaError_t e;
  if (k.wpc == 2) {
    e = plan_kernel(mpc_forces_solve_kernel<T, 2, false>, 2, k.smem, optin, sms, &c0);
    if (e == cudaSuccess) e = plan_kernel(mpc_forces_solve_kernel<T, 2, true>, 2, k.smem, optin, sms, &c1);
  } else {
    k.wpc = 1;
    e = plan_kernel(mpc_forces_solve_kernel<T, 1, false>, 1, k.smem, optin, sms, &c0);
    if (e == cudaSuccess) e = plan_kernel(mpc_forces_solve_kernel<T, 1, true>, 1, k.smem, optin, sms, &c1);
  }
  k.max_ctas = c0 < c1 ? c0 : c1;
  return e;
}
template <typename T>
static cudaError_t launch_forces(mpcb200_handle* h, ForcesArgs<T>& a, cudaStream_t s, const KernelPlan& k, int nwork) {
  const int ctas = grid_for(k, nwork);
  a.dynamic = (a.refine || nwork > ctas * k.wpc) ? 1 : 0;
  h->launches++;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(32 * k.wpc); cfg.dynamicSmemBytes = k.smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = a.refine ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  const bool rb = a.rb.nl > 0 && a.rb.nr > 0;
  if (k.wpc == 2) return rb ? cudaLaunchKernelEx(&cfg, mpc_forces_solve_kernel<T, 2, true>, a) : cudaLaunchKernelEx(&cfg, mpc_forces_solve_kernel<T, 2, false>, a);
  return rb ? cudaLaunchKernelEx(&cfg, mpc_forces_solve_kernel<T, 1, true>, a) : cudaLaunchKernelEx(&cfg, mpc_forces_solve_kernel<T, 1, false>, a);
}
template <typename T>
static void fill_forces_args(mpcb200_handle* h, ForcesArgs<T>& a, const double* wt, const double* xinit, const double* par, const double* Zin,
                             double* Z, int32_t* status, int32_t* iters, int32_t B) {
  a.fp.P = params_from_config<T>(h->cfg);
  for (int i = 0; i < 5; ++i) a.fp.Pt[i] = (T)wt[i];
  a.xinit = xinit; a.par = par; a.Zin = Zin; a.Z = Z; a.status = status; a.iters = iters;
  a.ctr = h->ctr; a.q_list = nullptr; a.B = B; a.refine = 0; a.dynamic = 0; a.pdl_primary = 0;
  const size_t nl = (size_t)h->rb_nl, nr = (size_t)h->rb_nr;
  const T* base = (const T*)(sizeof(T) == 4 ? h->rb_f32 : h->rb_f64);         // [left (nl x 2) | right (nr x 2)]
  a.rb.left = base; a.rb.right = base ? base + 2 * nl : nullptr;
  a.rb.nl = base ? (int)nl : 0; a.rb.nr = base ? (int)nr : 0; a.rb.r_min = (T)h->rb_rmin;
}
static int ensure_forces_plan(mpcb200_handle* h) {
  if (h->forces_planned) return 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->cfg.device));
  const int optin = (int)prop.sharedMemPerBlockOptin, sms = prop.multiProcessorCount;
  const size_t smem_max = prop.sharedMemPerBlockOptin;
  const bool f64 = h->cfg.precision == MPCB200_F64;
  const FLayout L(h->cfg.N);
  const int pref = (h->cfg.warps_per_cta == 1) ? 1 : 2;
  const int cand[2] = {pref, 1};
  bool fits = false;
  for (int i = 0; i < 2 && !fits; ++i) {
    const size_t need = forces_smem_bytes_for(h->cfg.N, L.words, h->elem, cand[i]);
    if (need + 1024 <= smem_max) { h->forces.wpc = cand[i]; h->forces.smem = need; fits = true; }
  }
  if (!fits) { h->err = "horizon too long: the FORCESPRO-formulation KKT slab does not fit shared memory"; return -2; }
  cudaError_t e = f64 ? plan_forces<double>(h->forces, optin, sms) : plan_forces<float>(h->forces, optin, sms);
  if (e != cudaSuccess) return fail(h, "kernel configuration (FORCESPRO formulation)", e);
  h->forces_refine.wpc = 0;
  if (!f64 && h->cfg.refine_f64 && h->q_list) {
    h->forces_refine.wpc = 1; h->forces_refine.smem = forces_smem_bytes_for(h->cfg.N, L.words, 8, 1);
    if (h->forces_refine.smem + 1024 > smem_max) { h->err = "horizon too long for the float64 refinement pass (FORCESPRO formulation)"; return -2; }
    e = plan_forces<double>(h->forces_refine, optin, sms);
    if (e != cudaSuccess) return fail(h, "kernel configuration (FORCESPRO formulation, float64 refinement)", e);
    if (h->forces_refine.max_ctas > sms) h->forces_refine.max_ctas = sms;
  }
  h->forces_planned = 1;
  return 0;
}

extern "C" {

int mpcb200_forces_set_road_boundaries(mpcb200_handle* h, const double* left, int32_t n_left, const double* right, int32_t n_right, double r_min) {
  if (!h) return -2;
  DeviceGuard guard(h->cfg.device);
  cudaFree(h->rb_f32); cudaFree(h->rb_f64); h->rb_f32 = nullptr; h->rb_f64 = nullptr; h->rb_nl = h->rb_nr = 0;
  if (n_left == 0 && n_right == 0) return 0;                            // rows off
  if (!left || !right || n_left < 1 || n_right < 1 || n_left > 65536 || n_right > 65536 || !(r_min >= 0.0)) { h->err = "road boundaries: need two vertex lists of 1 .. 65536 points and r_min >= 0"; return -2; }
  const size_t n = 2 * ((size_t)n_left + (size_t)n_right);
  double* h64 = (double*)malloc(n * sizeof(double));
  float* h32 = (float*)malloc(n * sizeof(float));
  if (!h64 || !h32) { free(h64); free(h32); h->err = "out of host memory"; return -4; }
  memcpy(h64, left, 2 * (size_t)n_left * sizeof(double));
  memcpy(h64 + 2 * (size_t)n_left, right, 2 * (size_t)n_right * sizeof(double));
  for (size_t i = 0; i < n; ++i) h32[i] = (float)h64[i];
  cudaError_t e = cudaMalloc(&h->rb_f64, n * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&h->rb_f32, n * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(h->rb_f64, h64, n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(h->rb_f32, h32, n * sizeof(float), cudaMemcpyHostToDevice);
  free(h64); free(h32);
  if (e != cudaSuccess) { cudaFree(h->rb_f32); cudaFree(h->rb_f64); h->rb_f32 = nullptr; h->rb_f64 = nullptr; return fail(h, "road boundaries upload", e); }
  h->rb_nl = n_left; h->rb_nr = n_right; h->rb_rmin = r_min;
  return 0;
}

int mpcb200_forces_solve(mpcb200_handle* h, const double* weights_terminal, const double* d_xinit, const double* d_params, const double* d_z_init,
                         double* d_z, int32_t* d_status, int32_t* d_iters, int32_t B, void* stream) {
  if (!h) return -2;
  if (B <= 0) return 0;
  if (B > h->cfg.max_batch) { h->err = "B exceeds cfg.max_batch"; return -2; }
  if (!weights_terminal || !d_xinit || !d_params || !d_z) { h->err = "null argument"; return -2; }
  if (((uintptr_t)d_xinit | (uintptr_t)d_params | (uintptr_t)d_z_init | (uintptr_t)d_z) & 7u) { h->err = "float64 arrays must be 8-byte aligned"; return -2; }
  DeviceGuard guard(h->cfg.device);
  int rc = ensure_forces_plan(h);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e;
  if (h->cfg.precision == MPCB200_F64) {
    ForcesArgs<double> a; fill_forces_args(h, a, weights_terminal, d_xinit, d_params, d_z_init, d_z, d_status, d_iters, B);
    e = launch_forces<double>(h, a, s, h->forces, B);
  } else {
    const bool refine = h->cfg.refine_f64 && d_status && h->forces_refine.wpc;
    ForcesArgs<float> a; fill_forces_args(h, a, weights_terminal, d_xinit, d_params, d_z_init, d_z, d_status, d_iters, B);
    a.q_list = refine ? h->q_list : nullptr; a.pdl_primary = refine ? 1 : 0;
    e = launch_forces<float>(h, a, s, h->forces, B);
    if (e == cudaSuccess && refine) {
      ForcesArgs<double> r; fill_forces_args(h, r, weights_terminal, d_xinit, d_params, d_z_init, d_z, d_status, d_iters, B);
      r.q_list = h->q_list; r.refine = 1;
      e = launch_forces<double>(h, r, s, h->forces_refine, B);
    }
  }
  if (e != cudaSuccess) return fail(h, "mpc_forces_solve_kernel launch", e);
  return 0;
}

}  // extern "C"
