// warp_ctx.cuh -- the 32-lane execution context the solver core (warp_core.cuh) is written against.
//
// On the device (nvcc, __CUDA_ARCH__) `WarpCtx` is an empty struct whose members are the warp intrinsics
// (__shfl_sync, __syncwarp, redux / butterfly reductions): zero overhead.
//
// On the host (g++, tests/host_sim only -- TEST TOOLING, never part of the product) the same member functions run on
// a 32-fiber lock-step emulator: every lane of a "warp" is a ucontext fiber, a shuffle / sync point yields to the
// next lane round-robin, so when lane 0 resumes every lane has reached the same point (exactly the guarantee the
// *_sync intrinsics give).  This lets the CUDA kernels' device code be run and debugged against the float64 oracle
// in a container without a GPU.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MPC_HD __host__ __device__ __forceinline__
#define MPC_D __device__ __forceinline__
#else
#define MPC_HD inline
#define MPC_D inline
#endif

namespace mpcb200 {

#if defined(__CUDACC__)
// ------------------------------------------------------------------------------------------------ device (nvcc)
// The members are __host__ __device__ only so that nvcc's host pass can parse the kernels; the host bodies are never run.
#if defined(__CUDA_ARCH__)
#define MPC_WARP_DEV(dev_expr, host_expr) dev_expr
#else
#define MPC_WARP_DEV(dev_expr, host_expr) host_expr
#endif
struct WarpCtx {
  static constexpr unsigned FULL = 0xffffffffu;
  MPC_HD int lane() const { return MPC_WARP_DEV((int)(threadIdx.x & 31u), 0); }
  MPC_HD void sync() const { MPC_WARP_DEV(__syncwarp(), (void)0); }
  MPC_HD float shfl(float v, int src) const { return MPC_WARP_DEV(__shfl_sync(FULL, v, src), v); }
  MPC_HD double shfl(double v, int src) const { return MPC_WARP_DEV(__shfl_sync(FULL, v, src), v); }
  MPC_HD int shfl(int v, int src) const { return MPC_WARP_DEV(__shfl_sync(FULL, v, src), v); }
  MPC_HD float shfl_xor(float v, int m) const { return MPC_WARP_DEV(__shfl_xor_sync(FULL, v, m), v); }
  MPC_HD double shfl_xor(double v, int m) const { return MPC_WARP_DEV(__shfl_xor_sync(FULL, v, m), v); }
  MPC_HD int shfl_xor(int v, int m) const { return MPC_WARP_DEV(__shfl_xor_sync(FULL, v, m), v); }
  MPC_HD bool all(bool p) const { return MPC_WARP_DEV(__all_sync(FULL, p) != 0, p); }
  MPC_HD bool any(bool p) const { return MPC_WARP_DEV(__any_sync(FULL, p) != 0, p); }
  // min / max of NON-NEGATIVE floats: their IEEE bit patterns order like unsigned integers -> one REDUX instruction
  MPC_HD float min_nonneg(float v) const { return MPC_WARP_DEV(__uint_as_float(__reduce_min_sync(FULL, __float_as_uint(v))), v); }
  MPC_HD float max_nonneg(float v) const { return MPC_WARP_DEV(__uint_as_float(__reduce_max_sync(FULL, __float_as_uint(v))), v); }
  MPC_HD double min_nonneg(double v) const {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int m = 16; m; m >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, m));
#endif
    return v;
  }
  MPC_HD double max_nonneg(double v) const {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int m = 16; m; m >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, m));
#endif
    return v;
  }
  template <typename T> MPC_HD T sum(T v) const {
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int m = 16; m; m >>= 1) v += __shfl_xor_sync(FULL, v, m);
#endif
    return v;
  }
};

#else
// ------------------------------------------------------------------------------------------------ host emulator
}  // namespace mpcb200
#include <ucontext.h>
#include <math.h>
#include <stdlib.h>
#include <vector>
namespace mpcb200 {

struct HostWarp {
  static constexpr int NL = 32;
  static constexpr size_t STACK = 512 * 1024;
  ucontext_t main_ctx, lane_ctx[NL];
  std::vector<char> stacks;
  uint64_t xbuf[2][NL];
  bool finished[NL];
  int cur;
  void (*body)(int lane, void* arg);
  void* arg;
  HostWarp() : stacks(STACK * NL) {}
  static HostWarp*& active() { static thread_local HostWarp* p = nullptr; return p; }
  static void tramp() {
    HostWarp* w = active();
    const int lane = w->cur;
    w->body(lane, w->arg);
    w->finished[lane] = true;
    w->yield_from(lane);      // never returns
  }
  // hand control to the next unfinished lane (round-robin); to main when all are finished
  void yield_from(int lane) {
    for (int s = 1; s <= NL; ++s) {
      const int nxt = (lane + s) % NL;
      if (!finished[nxt]) {
        if (nxt == lane) return;
        cur = nxt;
        swapcontext(&lane_ctx[lane], &lane_ctx[nxt]);
        return;
      }
    }
    swapcontext(&lane_ctx[lane], &main_ctx);
  }
  void run(void (*f)(int, void*), void* a) {
    body = f; arg = a; active() = this;
    for (int l = 0; l < NL; ++l) {
      finished[l] = false;
      getcontext(&lane_ctx[l]);
      lane_ctx[l].uc_stack.ss_sp = stacks.data() + STACK * l;
      lane_ctx[l].uc_stack.ss_size = STACK;
      lane_ctx[l].uc_link = &main_ctx;
      makecontext(&lane_ctx[l], (void (*)())tramp, 0);
    }
    cur = 0;
    swapcontext(&main_ctx, &lane_ctx[0]);
    active() = nullptr;
  }
};

struct WarpCtx {
  HostWarp* hw;
  int lane_;
  mutable int par;
  WarpCtx(HostWarp* h, int l) : hw(h), lane_(l), par(0) {}
  int lane() const { return lane_; }
  void sync() const { hw->yield_from(lane_); }
  template <typename T> T xchg(T v, int src) const {
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    hw->xbuf[par][lane_] = bits;
    hw->yield_from(lane_);
    T out;
    const uint64_t b = hw->xbuf[par][src & 31];
    memcpy(&out, &b, sizeof(T));
    par ^= 1;
    return out;
  }
  float shfl(float v, int src) const { return xchg(v, src); }
  double shfl(double v, int src) const { return xchg(v, src); }
  int shfl(int v, int src) const { return xchg(v, src); }
  float shfl_xor(float v, int m) const { return xchg(v, lane_ ^ m); }
  double shfl_xor(double v, int m) const { return xchg(v, lane_ ^ m); }
  int shfl_xor(int v, int m) const { return xchg(v, lane_ ^ m); }
  float shfl_up(float v, int d) const { return xchg(v, lane_ >= d ? lane_ - d : lane_); }
  double shfl_up(double v, int d) const { return xchg(v, lane_ >= d ? lane_ - d : lane_); }
  bool all(bool p) const { int v = p ? 1 : 0; for (int m = 16; m; m >>= 1) v &= xchg(v, lane_ ^ m); return v != 0; }
  bool any(bool p) const { int v = p ? 1 : 0; for (int m = 16; m; m >>= 1) v |= xchg(v, lane_ ^ m); return v != 0; }
  float min_nonneg(float v) const { for (int m = 16; m; m >>= 1) v = fminf(v, xchg(v, lane_ ^ m)); return v; }
  float max_nonneg(float v) const { for (int m = 16; m; m >>= 1) v = fmaxf(v, xchg(v, lane_ ^ m)); return v; }
  double min_nonneg(double v) const { for (int m = 16; m; m >>= 1) v = fmin(v, xchg(v, lane_ ^ m)); return v; }
  double max_nonneg(double v) const { for (int m = 16; m; m >>= 1) v = fmax(v, xchg(v, lane_ ^ m)); return v; }
  template <typename T> T sum(T v) const { for (int m = 16; m; m >>= 1) v += xchg(v, lane_ ^ m); return v; }
};
#endif

}  // namespace mpcb200
