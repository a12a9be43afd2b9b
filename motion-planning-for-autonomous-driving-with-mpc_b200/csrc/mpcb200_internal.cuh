// mpcb200_internal.cuh -- what the translation units of libmpcb200.so share: PTX helpers (mbarrier, TMA bulk copies), the handle,
// launch planning.  Not part of the C ABI (include/mpcb200.h).
#pragma once
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

// device-side work counters of a handle (self-resetting: the last warp to leave a launch that used them zeroes them)
struct WorkCtr {
  int next;       // next unclaimed work item beyond the statically assigned first wave
  int done;       // warps that have left the launch
  int q_count;    // refinement queue: problems the float32 pass did not bring to status 1
  int q_pad;
};

// ===================================================================================================== handle
#define MPCB200_HOST_STREAMS 4
struct KernelPlan {     // launch shape of one kernel family (solve / refinement / closed loop) for this handle
  int wpc;              // warps (= problems) per CTA
  size_t smem;          // dynamic shared memory per CTA
  int max_ctas;         // resident CTAs on the device (SM count x occupancy): the persistent grid never exceeds it
};
struct mpcb200_handle {
  mpcb200_config cfg;
  KernelPlan solve, refine, loop;
  int words;            // slab words per problem
  void* slab;           // global slab image [max_batch][words] (stepwise mode), allocated on first use
  void* state;
  void* obs_shift;
  WorkCtr* ctr;         // device work counters (zeroed at create, self-resetting afterwards)
  int* q_list;          // [max_batch] refinement queue (refine_f64 handles)
  mpcb200_scenario* scn_table;   // device copy of the scenario table (mpcb200_set_scenarios)
  int n_scn;
  KernelPlan scn, scn_refine;    // launch shapes of the per-problem-scenario kernels (planned at set_scenarios)
  KernelPlan forces[2], forces_refine[2];   // FORCESPRO-formulation kernels without / with the road-boundary rows (planned at the first mpcb200_forces_solve)
  int forces_planned;
  void *rb_f32, *rb_f64;              // road-boundary vertex lists on the device, [left | right], both precisions (mpcb200_forces_set_road_boundaries)
  int rb_nl, rb_nr;
  double rb_rmin;
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

inline std::string& create_err() { static thread_local std::string e; return e; }

inline int fail(mpcb200_handle* h, const char* what, cudaError_t e) {
  char buf[512];
  snprintf(buf, sizeof buf, "%s: %s", what, e == cudaSuccess ? "" : cudaGetErrorString(e));
  if (h) h->err = buf; else create_err() = buf;
  return -1;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(h, #call, e_); } while (0)

// Every entry point runs on the handle's device and leaves the caller's current device as it found it.
struct DeviceGuard {
  int prev; bool switched;
  explicit DeviceGuard(int dev) : prev(-1), switched(false) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = (cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

inline int grid_for(const KernelPlan& k, int nwork) {
  const int want = (nwork + k.wpc - 1) / k.wpc;
  return want < k.max_ctas ? want : k.max_ctas;
}

// Opt every instantiation this handle can launch in to the device's FULL opt-in shared memory (a property of the function on
// the device, shared by all handles: never a per-handle size, so one handle cannot lower another's limit), and ask the
// occupancy calculator how many CTAs of the handle's size stay resident.
template <typename K>
inline cudaError_t plan_kernel(K kern, int wpc, size_t smem, int optin, int sms, int* max_ctas) {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, kern);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
  if (e != cudaSuccess) return e;
  int occ = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * wpc, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorInvalidConfiguration;
  *max_ctas = occ * sms;
  return cudaSuccess;
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
  double* lam;          // [B][14N+2] inequality multipliers / obstacle slacks in and out (mpcb200_solve_dual), or null
  T* slab;              // global image of the slabs [B][words] (stepwise mode)
  ProbState<T>* state;  // [B] (stepwise mode)
  T* obs_shift;         // [B][6] shifted obstacle centres (stepwise mode)
  WorkCtr* ctr;         // work counters (dynamic scheduling beyond the first wave, refinement queue)
  int* q_list;          // [max_batch] refinement queue: written by the float32 pass, consumed by the float64 pass
  int B;
  int mode;
  int n_iter;
  int cold;             // 1: X / U are outputs only (cold start: X_0 tiled, zero controls)
  int refine;           // 1: float64 refinement pass -- the work list is q_list[0 .. q_count), warm start = the float32 X / U
  int dynamic;          // 1: B exceeds the launch's warps: warps claim further problems from ctr->next
  int pdl_primary;      // 1: a dependent launch (the refinement pass) follows: let it get resident while this grid runs
};

// shared-memory carve-up of one CTA: [WPC slabs of T][WPC float64 xref staging blocks of (5(N+1) + 1 rounded up to even) doubles]
// The staging block has one spare double in front: a row whose global address is 8 (mod 16) is placed 8 bytes in, so that the
// 16-byte-aligned interior the TMA bulk copy moves is 16-byte aligned on both sides.
MPC_HD int stg_doubles(int N) { return (5 * (N + 1) + 2) & ~1; }
template <typename T, int WPC>
struct Smem {
  int nx, nu;
  size_t slab_bytes;
  unsigned char* raw;
  __device__ Smem(unsigned char* raw_, int N, int words) : nx(5 * (N + 1)), nu(2 * N), slab_bytes((size_t)words * sizeof(T)), raw(raw_) {}
  __device__ T* slab(int w) const { return reinterpret_cast<T*>(raw + (size_t)w * slab_bytes); }
  __device__ double* xstg(int w) const { return reinterpret_cast<double*>(raw + (size_t)WPC * slab_bytes) + (size_t)w * stg_doubles(nu / 2); }
};
inline size_t smem_bytes_for(int N, int words, size_t elem, int wpc) {
  return (size_t)wpc * ((size_t)words * elem + (size_t)stg_doubles(N) * sizeof(double));
}

// One problem's xref block HBM (or pinned host memory) -> the warp's staging: the 16-byte-aligned interior by ONE TMA bulk copy
// issued by lane 0 (completion on the warp's mbarrier), the 8-byte head / tail words a misaligned row leaves by plain loads.
// Returns the staging address of the row.  Works for every 8-byte-aligned address (odd problem index at even N, sliced tensors).
__device__ __forceinline__ const double* fetch_xref(const double* g, double* stg, int nx, uint64_t* bar, uint32_t& phase, int lane) {
  const uint32_t head = (uint32_t)((uintptr_t)g & 8u);                 // 0 or 8 bytes in front of the aligned interior
  const uint32_t bytes = (uint32_t)nx * 8u;
  const uint32_t interior = (bytes - head) & ~15u;
  double* row = stg + (head >> 3);
  // the staging was last read through the generic proxy (previous problem): order those reads before the async-proxy write
  fence_async_smem();
  __syncwarp();
  if (lane == 0) {
    mbar_expect_tx(bar, interior);
    tma_load_1d(reinterpret_cast<unsigned char*>(row) + head, reinterpret_cast<const unsigned char*>(g) + head, interior, bar);
  }
  if (lane == 1 && head) row[0] = g[0];
  if (lane == 2 && head + interior < bytes) row[nx - 1] = g[nx - 1];
  mbar_wait(bar, phase);
  phase ^= 1u;
  __syncwarp();
  return row;
}


// phase-aligned variant of the fused float32 Gauss-Newton solve (aligned_solver.cu)
cudaError_t launch_solve_aligned(mpcb200_handle* h, SolveArgs<float>& a, cudaStream_t s, int nwork);
