// Shared definitions of libcggp_b200: context, error handling, dtype dispatch, reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>
#include <vector>

#include "../../include/cggp_b200.h"

struct cggp_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;
  int sm_count = 148;
  int cc_major = 0, cc_minor = 0;
  // scratch owned by the ctx (grown on demand, freed at destroy)
  void* ws = nullptr;
  size_t ws_bytes = 0;
  // device-side CG loop state: [0]=active, [1]=iteration, [2]=ticket
  int* cg_state = nullptr;
  int* cg_state_host = nullptr;  // pinned
  void* exp_tab = nullptr;  // exp tables of the pipelined matvec (built on first use)
  // pre-scaled, duplicated row norms of the pipelined kernels (kpipe::dup_scaled_norms); valid for the current solve
  void* xa2 = nullptr;
  size_t xa2_bytes = 0;
  const void* xa2_key = nullptr;
  int64_t xa2_n = 0;
  double xa2_alpha = 0.0;
  int64_t solve_epoch = 0;  // > 0 while cggp_cg_solve runs (a fresh value per solve), 0 outside
  int64_t xa2_epoch = -1;
  int64_t solve_counter = 0;
  // NCCL (dlopen'ed)
  void* comm = nullptr;
  int rank = 0, world = 1;
  // one-shot all-reduce over NVLink peer memory for the small per-iteration vectors (cggp_peer_*): every rank's
  // buffer = two data slots + two flags, mapped into every other rank through CUDA IPC
  void* peer_local = nullptr;        // this rank's buffer (cudaMalloc)
  void* peer_ptrs_host[16] = {};     // all ranks' buffers as seen from this process
  void** peer_ptrs_dev = nullptr;    // the same table on the device
  int peer_world = 0;
  int64_t peer_slot_bytes = 0;
  unsigned peer_seq = 0;
  int* peer_counter = nullptr;
  // optional per-section device timing (cggp_profile_*): event pairs recorded on the ctx stream
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev[CGGP_PROF_SECTIONS];  // start, stop, start, stop, ...
  size_t prof_used[CGGP_PROF_SECTIONS] = {0, 0, 0, 0};
};

// RAII section timer: no-op unless profiling is enabled on the ctx
struct ProfScope {
  cggp_ctx* ctx;
  int sec;
  bool live;
  ProfScope(cggp_ctx* c, int s);
  ~ProfScope();
};

// A process may drive several devices (one ctx each): every ABI entry runs on the device of its ctx and restores the
// caller's current device on return.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(const cggp_ctx* c) {
    if (c && cudaGetDevice(&prev) == cudaSuccess && prev != c->device) switched = cudaSetDevice(c->device) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};
#define CGGP_DEVICE_GUARD(ctx) DeviceGuard _cggp_device_guard(ctx)

#define CGGP_FAIL(ctx, code, ...)                       \
  do {                                                  \
    char _buf[512];                                     \
    snprintf(_buf, sizeof(_buf), __VA_ARGS__);          \
    if (ctx) (ctx)->err = _buf;                         \
    return (code);                                      \
  } while (0)

#define CGGP_CUDA(ctx, expr)                                                                          \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess)                                                                            \
      CGGP_FAIL(ctx, CGGP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                __LINE__);                                                                            \
  } while (0)

#define CGGP_LAUNCH_CHECK(ctx)                \
  do {                                        \
    (ctx)->launches += 1;                     \
    CGGP_CUDA(ctx, cudaPeekAtLastError());    \
  } while (0)

int cggp_ws_reserve(cggp_ctx* ctx, size_t bytes);

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block-wide sum (fixed order); every thread gets the result.  `red` has >= 33 elements.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from a previous use
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    T s = lane < nw ? red[lane] : T(0);
    s = warp_sum(s);
    if (lane == 0) red[32] = s;
  }
  __syncthreads();
  return red[32];
}

// Kernels inside the CG loop skip their work once the device-side loop has terminated.
__device__ __forceinline__ bool cg_inactive(const int* active) { return active != nullptr && *active == 0; }

struct LsParam {
  double v[128];
  int count;
};
