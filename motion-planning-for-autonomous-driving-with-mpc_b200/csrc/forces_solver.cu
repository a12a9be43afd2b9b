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
  const FLayout L(N, RB);
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

// ===================================================================================================== closed loop
// ForcesproOptimizer.optimize()'s loop (optimizer.py:286-362) for one ego per warp, no host round trip between the MPC steps: the
// parameter rows of step k (next N path points / headings replenished with the last one, the desired-velocity profile, the obstacle
// circle centres tiled, :288-317), one solve (warm start = the previous solution shifted one stage), the first input applied to
// the RK4 plant in float64 (model.eq, :359).  Shared memory per warp: the KKT slab + float64 parameter block [N][10] + stage
// variables [N][7] + xinit.
template <typename T>
struct ForcesLoopArgs {
  FParams<T> fp;
  RoadBounds<T> rb;
  double obstacle[6];
  const double* path;     // [Tlen][2]
  const double* orient;   // [Tlen]
  const double* vel;      // [Tlen] desired-velocity profile (optimizer.py:291-294)
  const double* x0;       // [B][5]
  double* traj;           // [B][Tlen][5]
  double* ctrl;           // [B][Tlen][2]
  int* status; int* iters;   // [B][Tlen]
  WorkCtr* ctr;
  double l_wb, dt;
  int B, Tlen, dynamic;
};
static size_t forces_loop_smem_bytes_for(int N, int words, size_t elem, int wpc) {
  return (size_t)wpc * ((size_t)words * elem + (size_t)(17 * N + 6) * sizeof(double));
}
__device__ __forceinline__ void ks_rhs_f64(const double* x, double u0, double u1, double l_wb, double* f) {
  double s, c; sincos(x[4], &s, &c);
  f[0] = x[3] * c; f[1] = x[3] * s; f[2] = u0; f[3] = u1; f[4] = x[3] / l_wb * tan(x[2]);
}

template <typename T, int WPC, bool RB>
__global__ void __launch_bounds__(32 * WPC) mpc_forces_closed_loop_kernel(const __grid_constant__ ForcesLoopArgs<T> a) {
  unsigned char* const smem_raw = mpc_dyn_smem;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.fp.P.N, Tlen = a.Tlen;
  const FLayout L(N, RB);
  const int total_warps = gridDim.x * WPC;
  double* const par = reinterpret_cast<double*>(smem_raw + (size_t)WPC * L.words * sizeof(T)) + (size_t)wid * (17 * N + 6);
  double* const Zw = par + 10 * N;
  double* const xs = Zw + 7 * N;         // xinit of the current step
  const WarpCtx w;
  ForcesSolver<T, RB> S(a.fp, SlabRef<T>{wid * L.words}, w, a.rb);
  for (int b = blockIdx.x * WPC + wid; b < a.B;) {
    double x[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) x[j] = a.x0[(size_t)b * 5 + j];
    for (int k = 0; k < Tlen; ++k) {
      __syncwarp();
      if (lane < 5) { xs[lane] = x[lane]; a.traj[((size_t)b * Tlen + k) * 5 + lane] = x[lane]; }
      for (int j = lane; j < N; j += 32) {
        const int idx = (k + 1 + j < Tlen) ? (k + 1 + j) : (Tlen - 1);
        double* p = par + 10 * j;
        p[0] = a.path[2 * idx]; p[1] = a.path[2 * idx + 1]; p[2] = a.vel[idx]; p[3] = a.orient[idx];
#pragma unroll
        for (int q = 0; q < 6; ++q) p[4 + q] = a.obstacle[q];
      }
      __syncwarp();
      ProbState<T> st;
      S.load(xs, par, k > 0 ? Zw : nullptr);
      S.init(st);
      for (int it = 0; it < a.fp.P.max_iter && !st.done; ++it) S.iterate(st);
      S.store(xs, par, Zw);
      const double u0 = Zw[0], u1 = Zw[1];
      if (lane == 0) {
        a.ctrl[((size_t)b * Tlen + k) * 2] = u0; a.ctrl[((size_t)b * Tlen + k) * 2 + 1] = u1;
        a.status[(size_t)b * Tlen + k] = st.status; a.iters[(size_t)b * Tlen + k] = st.iters;
      }
      // plant: one RK4 step in float64 (every lane redundantly)
      {
        const double h = a.dt;
        double k1[5], k2[5], k3[5], k4[5], t[5];
        ks_rhs_f64(x, u0, u1, a.l_wb, k1);
        for (int j = 0; j < 5; ++j) t[j] = x[j] + 0.5 * h * k1[j];
        ks_rhs_f64(t, u0, u1, a.l_wb, k2);
        for (int j = 0; j < 5; ++j) t[j] = x[j] + 0.5 * h * k2[j];
        ks_rhs_f64(t, u0, u1, a.l_wb, k3);
        for (int j = 0; j < 5; ++j) t[j] = x[j] + h * k3[j];
        ks_rhs_f64(t, u0, u1, a.l_wb, k4);
        for (int j = 0; j < 5; ++j) x[j] += h / 6.0 * (k1[j] + 2.0 * k2[j] + 2.0 * k3[j] + k4[j]);
      }
      // warm start of the next step: the solution shifted one stage, the last stage repeated
      __syncwarp();
      for (int j0 = 0; j0 < N; j0 += 32) {
        const int j = j0 + lane;
        double z[7];
        if (j < N) { const int src = (j + 1 < N) ? j + 1 : N - 1; for (int q = 0; q < 7; ++q) z[q] = Zw[7 * src + q]; }
        __syncwarp();
        if (j < N) { for (int q = 0; q < 7; ++q) Zw[7 * j + q] = z[q]; }
        __syncwarp();
      }
    }
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

template <typename T>
static cudaError_t plan_forces(KernelPlan& k, bool rb, int optin, int sms) {
  if (k.wpc == 2) return rb ? plan_kernel(mpc_forces_solve_kernel<T, 2, true>, 2, k.smem, optin, sms, &k.max_ctas)
                            : plan_kernel(mpc_forces_solve_kernel<T, 2, false>, 2, k.smem, optin, sms, &k.max_ctas);
  k.wpc = 1;
  return rb ? plan_kernel(mpc_forces_solve_kernel<T, 1, true>, 1, k.smem, optin, sms, &k.max_ctas)
            : plan_kernel(mpc_forces_solve_kernel<T, 1, false>, 1, k.smem, optin, sms, &k.max_ctas);
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
// launch shapes of the kernel without (v = 0) and with (v = 1) the road-boundary rows: their slabs differ (FLayout), so do the
// shared-memory sizes and the resident grids.  One problem per CTA unless cfg.warps_per_cta asks for two: shared memory binds the
// residency (11 problems per SM at N = 30 in float32), and single-warp CTAs pack it without a remainder.
static int ensure_forces_plan(mpcb200_handle* h) {
  if (h->forces_planned) return 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->cfg.device));
  const int optin = (int)prop.sharedMemPerBlockOptin, sms = prop.multiProcessorCount;
  const size_t smem_max = prop.sharedMemPerBlockOptin;
  const bool f64 = h->cfg.precision == MPCB200_F64;
  for (int v = 0; v < 2; ++v) {
    const FLayout L(h->cfg.N, v == 1);
    const int cand[2] = {h->cfg.warps_per_cta == 2 ? 2 : 1, 1};
    bool fits = false;
    for (int i = 0; i < 2 && !fits; ++i) {
      const size_t need = forces_smem_bytes_for(h->cfg.N, L.words, h->elem, cand[i]);
      if (need + 1024 <= smem_max) { h->forces[v].wpc = cand[i]; h->forces[v].smem = need; fits = true; }
    }
    if (!fits) { h->err = "horizon too long: the FORCESPRO-formulation KKT slab does not fit shared memory"; return -2; }
    cudaError_t e = f64 ? plan_forces<double>(h->forces[v], v == 1, optin, sms) : plan_forces<float>(h->forces[v], v == 1, optin, sms);
    if (e != cudaSuccess) return fail(h, "kernel configuration (FORCESPRO formulation)", e);
    h->forces_refine[v].wpc = 0;
    if (!f64 && h->cfg.refine_f64 && h->q_list) {
      h->forces_refine[v].wpc = 1; h->forces_refine[v].smem = forces_smem_bytes_for(h->cfg.N, L.words, 8, 1);
      if (h->forces_refine[v].smem + 1024 > smem_max) { h->err = "horizon too long for the float64 refinement pass (FORCESPRO formulation)"; return -2; }
      e = plan_forces<double>(h->forces_refine[v], v == 1, optin, sms);
      if (e != cudaSuccess) return fail(h, "kernel configuration (FORCESPRO formulation, float64 refinement)", e);
      if (h->forces_refine[v].max_ctas > sms) h->forces_refine[v].max_ctas = sms;
    }
  }
  h->forces_planned = 1;
  return 0;
}

template <typename T>
static int forces_closed_loop_t(mpcb200_handle* h, const double* wt, int32_t Tlen, const double* d_path, const double* d_orient, const double* d_vel,
                                const double* d_x0, double* d_traj, double* d_ctrl, int32_t* d_status, int32_t* d_iters, int32_t B, cudaStream_t s) {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->cfg.device));
  const int optin = (int)prop.sharedMemPerBlockOptin, sms = prop.multiProcessorCount;
  ForcesLoopArgs<T> a;
  ForcesArgs<T> tmp; fill_forces_args(h, tmp, wt, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, B);
  a.fp = tmp.fp; a.rb = tmp.rb;
  const FLayout L(h->cfg.N, a.rb.nl > 0 && a.rb.nr > 0);
  const size_t smem = forces_loop_smem_bytes_for(h->cfg.N, L.words, sizeof(T), 1);
  if (smem + 1024 > prop.sharedMemPerBlockOptin) { h->err = "horizon too long for the FORCESPRO-formulation closed loop"; return -2; }
  for (int i = 0; i < 6; ++i) a.obstacle[i] = h->cfg.obstacle[i];
  a.path = d_path; a.orient = d_orient; a.vel = d_vel; a.x0 = d_x0; a.traj = d_traj; a.ctrl = d_ctrl; a.status = d_status; a.iters = d_iters;
  a.ctr = h->ctr; a.l_wb = h->cfg.l_wb; a.dt = h->cfg.dt; a.B = B; a.Tlen = Tlen;
  const bool rb = a.rb.nl > 0 && a.rb.nr > 0;
  int max_ctas = 0;
  cudaError_t e = rb ? plan_kernel(mpc_forces_closed_loop_kernel<T, 1, true>, 1, smem, optin, sms, &max_ctas)
                     : plan_kernel(mpc_forces_closed_loop_kernel<T, 1, false>, 1, smem, optin, sms, &max_ctas);
  if (e != cudaSuccess) return fail(h, "kernel configuration (FORCESPRO-formulation closed loop)", e);
  const int ctas = B < max_ctas ? B : max_ctas;
  a.dynamic = (B > ctas) ? 1 : 0;
  if (rb) mpc_forces_closed_loop_kernel<T, 1, true><<<ctas, 32, smem, s>>>(a);
  else mpc_forces_closed_loop_kernel<T, 1, false><<<ctas, 32, smem, s>>>(a);
  h->launches++;
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(h, "mpc_forces_closed_loop_kernel launch", e);
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
  const int v = (h->rb_f64 && h->rb_nl > 0 && h->rb_nr > 0) ? 1 : 0;
  if (h->cfg.precision == MPCB200_F64) {
    ForcesArgs<double> a; fill_forces_args(h, a, weights_terminal, d_xinit, d_params, d_z_init, d_z, d_status, d_iters, B);
    e = launch_forces<double>(h, a, s, h->forces[v], B);
  } else {
    const bool refine = h->cfg.refine_f64 && d_status && h->forces_refine[v].wpc;
    ForcesArgs<float> a; fill_forces_args(h, a, weights_terminal, d_xinit, d_params, d_z_init, d_z, d_status, d_iters, B);
    a.q_list = refine ? h->q_list : nullptr; a.pdl_primary = refine ? 1 : 0;
    e = launch_forces<float>(h, a, s, h->forces[v], B);
    if (e == cudaSuccess && refine) {
      ForcesArgs<double> r; fill_forces_args(h, r, weights_terminal, d_xinit, d_params, d_z_init, d_z, d_status, d_iters, B);
      r.q_list = h->q_list; r.refine = 1;
      e = launch_forces<double>(h, r, s, h->forces_refine[v], B);
    }
  }
  if (e != cudaSuccess) return fail(h, "mpc_forces_solve_kernel launch", e);
  return 0;
}

int mpcb200_forces_closed_loop(mpcb200_handle* h, const double* weights_terminal, int32_t iter_length, const double* d_path, const double* d_orientation,
                               const double* d_velocity, const double* d_x0, double* d_traj, double* d_ctrl, int32_t* d_status, int32_t* d_iters,
                               int32_t B, void* stream) {
  if (!h) return -2;
  if (B <= 0) return 0;
  if (B > h->cfg.max_batch) { h->err = "B exceeds cfg.max_batch"; return -2; }
  if (!weights_terminal || !d_path || !d_orientation || !d_velocity || !d_x0 || !d_traj || !d_ctrl || !d_status || !d_iters || iter_length < 1) { h->err = "null argument"; return -2; }
  DeviceGuard guard(h->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  if (h->cfg.precision == MPCB200_F64) return forces_closed_loop_t<double>(h, weights_terminal, iter_length, d_path, d_orientation, d_velocity, d_x0, d_traj, d_ctrl, d_status, d_iters, B, s);
  return forces_closed_loop_t<float>(h, weights_terminal, iter_length, d_path, d_orientation, d_velocity, d_x0, d_traj, d_ctrl, d_status, d_iters, B, s);
}

}  // extern "C"
