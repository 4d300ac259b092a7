// Conjugate gradient: fused vector update kernels + device-resident loop (cggp/conjugate_gradient.py:24-122).
//
// Per iteration the reference runs ~10 un-fused TF kernels over the [B, n] state plus a host round trip for the
// while_loop predicate.  Here one kernel (one CTA per right-hand side; all dot products of that row are CTA-local,
// deterministic) does:  denom = sum p*pA; gamma; v += gamma p; r -= gamma pA; (z, rz') = precond(r);
// p = z + p rz'/rz; 0.5|r|^2  -- and the last CTA to finish evaluates the stopping condition for the whole batch,
// appends the residual history row and advances the device-side iteration counter / active flag.  Kernels enqueued
// after termination see active == 0 and return immediately, so the host never synchronises inside the loop.
//
// Rounding follows the reference op by op where it is cheap to do so (separate multiply and add, (p*rz')/rz with a
// true division, guards `<= 1e-16 -> 0`), so trajectories differ from the TF path only through summation order.
#include <cstdlib>
#include <vector>

#include "common.cuh"

int cggp_symm_matmul_ex(cggp_ctx* ctx, int dtype, const void* A, int64_t lda, int64_t n, const void* V, int64_t ldv,
                        int B, void* Y, int64_t ldy, const void* addend, int64_t ldadd, double scale,
                        const int* active);
int cggp_matvec_dispatch(cggp_ctx* ctx, int dtype, int kind, double variance, const void* PX, const void* nX,
                         int64_t n, const void* PZ, const void* nZ, int64_t m, int D, int64_t ldp, const void* V,
                         int64_t ldv, int B, void* W, int64_t ldw, int variant, const int* active);
int cggp_matvec_tf32(cggp_ctx* ctx, int kind, double variance, const float* Xb, const float* Xs, const float* xn,
                     int64_t n, const float* Zb, const float* Zs, const float* zn, int64_t m, int D, const float* V,
                     int64_t ldv, int B, float* W, int64_t ldw, int nsplit, const int* active);
int cggp_ws2_reserve(cggp_ctx* ctx, size_t bytes);
void* cggp_ws2_ptr(cggp_ctx* ctx);

__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }

enum { MODE_STEP = 0, MODE_PRE = 1, MODE_POST = 2, MODE_INIT = 3 };

template <typename T>
struct StepArgs {
  int mode;
  int B;
  int64_t n;
  const T* q;    // pA (MODE_STEP / MODE_PRE) or vA (MODE_POST / MODE_INIT)
  const T* rhs;  // b (MODE_POST / MODE_INIT)
  T* v;
  T* r;
  T* p;
  T* z;          // scratch for the block preconditioner (nullptr for Eye)
  T* rz;         // [B]
  T* half_rr;    // [B] 0.5 |r|^2
  T* half_rz;    // [B] 0.5 rz (stats_error), may be nullptr
  T* history;    // [cap, B] or nullptr
  int64_t history_cap;
  T threshold;
  int max_iterations;
  int* state;    // [0] active [1] iteration [2] ticket; nullptr = stand-alone step (no loop control)
  // block preconditioner
  int num_blocks, block_size;
  const int64_t* block_idx;
  const T* chol;
  // dense preconditioner: the step kernel stops after r / 0.5|r|^2 (`defer` != 0); z = r @ Pinv is a separate
  // symmetric product and cg_finish_kernel completes the iteration
  int defer;
  const T* pinv;
  int64_t ldpinv;
};

constexpr int MAX_BS = 64;

// z[blk] = (L L^T)^-1 r[blk] for the blocks of this row: the batched triangular solve of the block-Jacobi
// preconditioner.  One WARP per block (blocks are independent, the CTA's warps take them round-robin): the lanes hold
// the block's right-hand side (up to two entries per lane, block_size <= 64); every substitution step is one
// lane-parallel multiply + a warp-shuffle reduction, so a 64 x 64 block costs 2 x 64 short steps instead of 2 x 2016
// dependent FMAs of a single thread.
template <typename T>
__device__ void block_precond_apply(const StepArgs<T>& a, const T* __restrict__ r, T* __restrict__ z) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int bs = a.block_size;
  for (int blk = warp; blk < a.num_blocks; blk += nwarps) {
    const int64_t* idx = a.block_idx + (int64_t)blk * bs;
    const T* L = a.chol + (int64_t)blk * bs * bs;
    // y[k] lives in lane k % 32, slot k / 32
    T y0 = lane < bs ? r[idx[lane]] : T(0);
    T y1 = lane + 32 < bs ? r[idx[lane + 32]] : T(0);
    // forward: L y = r.  After step i, y[i] is final; every later row k > i subtracts L[k, i] y[i] (column sweep)
    for (int i = 0; i < bs; ++i) {
      const T yi_raw = __shfl_sync(0xffffffffu, i < 32 ? y0 : y1, i & 31);
      const T yi = yi_raw / L[i * bs + i];
      if (lane == (i & 31)) { if (i < 32) y0 = yi; else y1 = yi; }
      if (lane > i && lane < bs) y0 -= L[lane * bs + i] * yi;
      if (lane + 32 > i && lane + 32 < bs) y1 -= L[(lane + 32) * bs + i] * yi;
    }
    // backward: L^T x = y.  Row i of L^T is column i of L: x[k] for k < i subtracts L[i, k] x[i]
    for (int i = bs - 1; i >= 0; --i) {
      const T xi_raw = __shfl_sync(0xffffffffu, i < 32 ? y0 : y1, i & 31);
      const T xi = xi_raw / L[i * bs + i];
      if (lane == (i & 31)) { if (i < 32) y0 = xi; else y1 = xi; }
      if (lane < i) y0 -= L[i * bs + lane] * xi;
      if (lane + 32 < i) y1 -= L[i * bs + lane + 32] * xi;
    }
    if (lane < bs) z[idx[lane]] = y0;
    if (lane + 32 < bs) z[idx[lane + 32]] = y1;
  }
}

// Loop control by the last CTA of a launch (stopping condition, conjugate_gradient.py:59-62): history row, iteration
// counter and the device-side `active` flag.
template <typename T>
__device__ __forceinline__ void cg_loop_control(const StepArgs<T>& a, int& s_last) {
  if (!a.state) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&a.state[2], 1) == a.B - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int it = (a.mode == MODE_INIT) ? 0 : a.state[1] + 1;
  int over = 0;
  const volatile T* hr = a.half_rr;
  for (int k = threadIdx.x; k < a.B; k += blockDim.x) {
    const T h = hr[k];
    if (h > a.threshold) over = 1;
    if (a.history && it < a.history_cap) a.history[(int64_t)it * a.B + k] = h;
  }
  over = __syncthreads_or(over);
  if (threadIdx.x == 0) {
    a.state[1] = it;
    a.state[2] = 0;
    a.state[0] = (over && it < a.max_iterations) ? 1 : 0;
  }
}

// EPT > 0: the row of the four vectors lives in REGISTERS (EPT elements per thread, n <= EPT * blockDim.x; plain
// iteration, Eye preconditioner): p, pA, v, r are read from HBM once and v, r, p written once - the algorithmic
// 7 B n sizeof bytes - instead of the three passes of the general form (11 vector transfers, 4 of them from L2).
// Same arithmetic in the same order as the general form (element e of thread t is index t + e * blockDim.x, exactly
// the order of its strided loops), so the results are bit-identical.
template <typename T, int EPT>
__device__ __forceinline__ void cg_step_body(const StepArgs<T>& a, T* red, int& s_last) {
  const int b = blockIdx.x;
  const int64_t n = a.n;
  const T* q = a.q + (int64_t)b * n;
  T* v = a.v + (int64_t)b * n;
  T* r = a.r + (int64_t)b * n;
  T* p = a.p + (int64_t)b * n;
  T* z = a.z ? a.z + (int64_t)b * n : nullptr;
  const T min_float = T(1e-16);  // conjugate_gradient.py:50 (1e-16 also for float32)
  T gamma = T(0);
  const T rz_old = (a.mode == MODE_STEP || a.mode == MODE_PRE) ? a.rz[b] : T(0);
  if constexpr (EPT > 0) {
    T pk[EPT], qk[EPT], vk[EPT], rk[EPT];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int64_t k = threadIdx.x + (int64_t)e * blockDim.x;
      const bool ok = k < n;
      pk[e] = ok ? p[k] : T(0);
      qk[e] = ok ? q[k] : T(0);
      vk[e] = ok ? v[k] : T(0);
      rk[e] = ok ? r[k] : T(0);
    }
    T part = T(0);
#pragma unroll
    for (int e = 0; e < EPT; ++e)
      if (threadIdx.x + (int64_t)e * blockDim.x < n) part = add_rn(part, mul_rn(pk[e], qk[e]));
    const T denom = block_sum(part, red);                       // :66
    gamma = denom <= min_float ? T(0) : div_rn(rz_old, denom);  // :67-68
    T rr_part = T(0);
#pragma unroll
    for (int e = 0; e < EPT; ++e)
      if (threadIdx.x + (int64_t)e * blockDim.x < n) {
        vk[e] = add_rn(vk[e], mul_rn(gamma, pk[e]));   // :69
        rk[e] = add_rn(rk[e], -mul_rn(gamma, qk[e]));  // :75
        rr_part = add_rn(rr_part, mul_rn(rk[e], rk[e]));
      }
    const T rr = block_sum(rr_part, red);
    const T rz_new = rr;  // EyePreconditioner: z = r
    const bool dead = rz_old <= min_float;  // :79
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int64_t k = threadIdx.x + (int64_t)e * blockDim.x;
      if (k < n) {
        const T upd = dead ? T(0) : div_rn(mul_rn(pk[e], rz_new), rz_old);  // :78
        v[k] = vk[e];
        r[k] = rk[e];
        p[k] = add_rn(rk[e], upd);  // :83
      }
    }
    if (threadIdx.x == 0) {
      a.rz[b] = rz_new;
      a.half_rr[b] = T(0.5) * rr;
      if (a.half_rz) a.half_rz[b] = T(0.5) * rz_new;  // :97
    }
    cg_loop_control(a, s_last);
    return;
  }

  if (a.mode == MODE_STEP || a.mode == MODE_PRE) {
    T part = T(0);
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) part = add_rn(part, mul_rn(p[k], q[k]));
    const T denom = block_sum(part, red);              // :66
    gamma = denom <= min_float ? T(0) : div_rn(rz_old, denom);  // :67-68
  }
  if (a.mode == MODE_PRE) {  // reset iteration, first half: only v is updated (:69); r is recomputed from b - v@A
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) v[k] = add_rn(v[k], mul_rn(gamma, p[k]));
    return;
  }

  T rr_part = T(0);
  if (a.mode == MODE_STEP) {
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
      const T pk = p[k], qk = q[k];
      v[k] = add_rn(v[k], mul_rn(gamma, pk));          // :69
      const T rk = add_rn(r[k], -mul_rn(gamma, qk));   // :75
      r[k] = rk;
      rr_part = add_rn(rr_part, mul_rn(rk, rk));
    }
  } else {  // MODE_POST / MODE_INIT: r = b - v@A (:74, :88)
    const T* rhs = a.rhs + (int64_t)b * n;
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
      const T rk = add_rn(rhs[k], -q[k]);
      r[k] = rk;
      rr_part = add_rn(rr_part, mul_rn(rk, rk));
    }
  }
  const T rr = block_sum(rr_part, red);
  if (a.defer) {
    if (threadIdx.x == 0) a.half_rr[b] = T(0.5) * rr;
    return;
  }
  T rz_new = rr;  // EyePreconditioner (:131-134): z = r, rz = sum r^2
  const T* zsrc = r;
  if (a.num_blocks > 0) {  // BlockPreconditioner (intent of :137-157)
    __syncthreads();       // r of this row is complete in global memory (same CTA wrote it)
    block_precond_apply(a, r, z);
    __syncthreads();
    T zp = T(0);
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) zp = add_rn(zp, mul_rn(z[k], r[k]));
    rz_new = block_sum(zp, red);  // :157
    zsrc = z;
  }
  if (a.mode == MODE_STEP) {
    const bool dead = rz_old <= min_float;  // :79
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
      const T upd = dead ? T(0) : div_rn(mul_rn(p[k], rz_new), rz_old);  // :78  (p * new_rz) / rz
      p[k] = add_rn(zsrc[k], upd);                                       // :83
    }
  } else {
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) p[k] = zsrc[k];  // :82 / :90
  }
  const T half = T(0.5) * rr;
  if (threadIdx.x == 0) {
    a.rz[b] = rz_new;
    a.half_rr[b] = half;
    if (a.half_rz) a.half_rz[b] = T(0.5) * rz_new;  // :97
  }
  cg_loop_control(a, s_last);
}

template <typename T, int EPT>
__global__ void __launch_bounds__(512) cg_step_kernel(const StepArgs<T> a) {
  if (a.state && a.state[0] == 0) return;
  __shared__ T red[33];
  __shared__ int s_last;
  cg_step_body<T, EPT>(a, red, s_last);
}

// ---------------------------------------------------------------------------------------------------------
extern "C" int cggp_peer_enabled(cggp_ctx* ctx);

// Fused tail of an iteration on the matrix-free operator (Eye preconditioner, B <= 8): ONE kernel does
//   (1) the all-reduce of this rank's partial product over NVLink peer memory: every rank copies its vector into its
//       own IPC-shared slot, publishes a sequence number, waits for the other ranks' numbers and sums all slots in
//       RANK ORDER straight out of peer memory (bit-identical on all ranks; the protocol of peer_allreduce_kernel),
//   (2) q = p Kuu + scale * w   (the ranks' shares of p Kuu ride in the same sum), and
//   (3) the CG vector update with its dot products, stopping condition and history row (cg_step_body).
// Per iteration this replaces ncclAllReduce + the replicated Kuu product + the step kernel (profiles/README.md).
// ---------------------------------------------------------------------------------------------------------
template <typename T>
struct TailArgs {
  const T* w;       // [B, n] this rank's partial Kuf Kfu product (already all-reduced when world == 1 or by NCCL)
  T* q;             // [B, n] in: p @ Kuu (world == 1: all columns; else this rank's columns [col_lo, col_hi)), out: p Sigma
  T scale, inv_scale;
  int64_t col_lo, col_hi;  // world > 1: the slice of p @ Kuu this rank computed and contributes to the sum
  int shard;               // world > 1: 1 = p @ Kuu sharded over ranks (above), 0 = every rank holds all of it in q
  // peer exchange (world > 1)
  char* const* peers;
  int rank, world;
  unsigned seq;
  int64_t slot_bytes;
  int* counter;
};

__device__ __forceinline__ unsigned tail_ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void tail_st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(512) cg_tail_kernel(const StepArgs<T> a, const TailArgs<T> t) {
  if (a.state && a.state[0] == 0) return;  // identical on every rank: all ranks skip together
  __shared__ T red[33];
  __shared__ int s_last;
  const int b = blockIdx.x;
  const int64_t n = a.n;
  const T* w = t.w + (int64_t)b * n;
  T* q = t.q + (int64_t)b * n;
  if (t.world > 1) {
    const int slot = (int)(t.seq & 1u);
    T* mine = reinterpret_cast<T*>(t.peers[t.rank] + slot * t.slot_bytes) + (int64_t)b * n;
    // this rank's share of the replicated term rides in the same sum: every rank computes 1 / world of p @ Kuu
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x)
      mine[k] = (t.shard && k >= t.col_lo && k < t.col_hi) ? fma(t.inv_scale, q[k], w[k]) : w[k];
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      if (atomicAdd(t.counter, 1) == (int)gridDim.x - 1) {  // the last CTA of this rank: all rows are in the slot
        *t.counter = 0;
        __threadfence_system();
        tail_st_release_sys(reinterpret_cast<unsigned*>(t.peers[t.rank] + 2 * t.slot_bytes) + slot, t.seq);
      }
      const long long t0 = clock64();
      for (int r = 0; r < t.world; ++r) {
        const unsigned* flag = reinterpret_cast<const unsigned*>(t.peers[r] + 2 * t.slot_bytes) + slot;
        while ((int)(tail_ld_acquire_sys(flag) - t.seq) < 0) {
          if (clock64() - t0 > (1ll << 34)) {  // ~8 s: a peer died; report instead of hanging the device
            if (a.state) a.state[3] = 1;
            break;
          }
        }
      }
    }
    __syncthreads();
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
      T part[16];
#pragma unroll
      for (int r = 0; r < 16; ++r)
        part[r] = r < t.world
                      ? __ldcv(reinterpret_cast<const T*>(t.peers[r] + slot * t.slot_bytes) + (int64_t)b * n + k)
                      : T(0);
      T v = T(0);
#pragma unroll
      for (int r = 0; r < 16; ++r)
        if (r < t.world) v += part[r];
      q[k] = t.shard ? t.scale * v : fma(t.scale, v, q[k]);
    }
  } else {
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) q[k] = fma(t.scale, w[k], q[k]);
  }
  __syncthreads();  // q of this row is complete (written by this CTA)
  cg_step_body<T, 0>(a, red, s_last);
}

// Second half of an iteration with the dense preconditioner: given z = r @ Pinv, rz' = sum z*r (:157-style),
// p = z + p rz'/rz (:78-83) or p = z (init / refresh), stats and loop control as in cg_step_kernel.
template <typename T>
__global__ void __launch_bounds__(512) cg_finish_kernel(const StepArgs<T> a) {
  if (a.state && a.state[0] == 0) return;
  __shared__ T red[33];
  __shared__ int s_last;
  const int b = blockIdx.x;
  const int64_t n = a.n;
  const T* r = a.r + (int64_t)b * n;
  const T* z = a.z + (int64_t)b * n;
  T* p = a.p + (int64_t)b * n;
  const T min_float = T(1e-16);
  T zp = T(0);
  for (int64_t k = threadIdx.x; k < n; k += blockDim.x) zp = add_rn(zp, mul_rn(z[k], r[k]));
  const T rz_new = block_sum(zp, red);
  if (a.mode == MODE_STEP) {
    const T rz_old = a.rz[b];
    const bool dead = rz_old <= min_float;
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
      const T upd = dead ? T(0) : div_rn(mul_rn(p[k], rz_new), rz_old);
      p[k] = add_rn(z[k], upd);
    }
  } else {
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) p[k] = z[k];
  }
  __syncthreads();  // every thread has read rz_old before it is overwritten
  if (threadIdx.x == 0) {
    a.rz[b] = rz_new;
    if (a.half_rz) a.half_rz[b] = T(0.5) * rz_new;
  }
  if (!a.state) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&a.state[2], 1) == a.B - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int it = (a.mode == MODE_INIT) ? 0 : a.state[1] + 1;
  int over = 0;
  const volatile T* hr = a.half_rr;
  for (int k = threadIdx.x; k < a.B; k += blockDim.x) {
    const T h = hr[k];
    if (h > a.threshold) over = 1;
    if (a.history && it < a.history_cap) a.history[(int64_t)it * a.B + k] = h;
  }
  over = __syncthreads_or(over);
  if (threadIdx.x == 0) {
    a.state[1] = it;
    a.state[2] = 0;
    a.state[0] = (over && it < a.max_iterations) ? 1 : 0;
  }
}

template <typename T>
static int launch_step(cggp_ctx* ctx, const StepArgs<T>& a) {
  {
    ProfScope prof(ctx, 2);
    // plain iteration with the Eye preconditioner and a row that fits the registers of one CTA: single-pass form
    // (256 threads up to n = 2048: two or three CTAs per SM keep more rows in flight than one CTA of 512)
    const bool regs = a.mode == MODE_STEP && !a.defer && a.num_blocks == 0 && a.n <= 8 * 512;
    const int threads = (regs && a.n <= 2048) ? 256 : 512;
    const int64_t ept = (a.n + threads - 1) / threads;
    if (regs && ept <= 1) cg_step_kernel<T, 1><<<a.B, threads, 0, ctx->stream>>>(a);
    else if (regs && ept <= 2) cg_step_kernel<T, 2><<<a.B, threads, 0, ctx->stream>>>(a);
    else if (regs && ept <= 4) cg_step_kernel<T, 4><<<a.B, threads, 0, ctx->stream>>>(a);
    else if (regs) cg_step_kernel<T, 8><<<a.B, threads, 0, ctx->stream>>>(a);
    else cg_step_kernel<T, 0><<<a.B, 512, 0, ctx->stream>>>(a);
    CGGP_LAUNCH_CHECK(ctx);
  }
  if (a.defer && a.mode != MODE_PRE) {
    // z = r @ Pinv (symmetric, HBM-bound for few right-hand sides), then the second half of the iteration
    int rc = cggp_symm_matmul_ex(ctx, sizeof(T) == 8 ? CGGP_F64 : CGGP_F32, a.pinv, a.ldpinv, a.n, a.r, a.n, a.B, a.z,
                                 a.n, nullptr, 0, 0.0, a.state);
    if (rc) return rc;
    ProfScope prof(ctx, 2);
    cg_finish_kernel<T><<<a.B, 512, 0, ctx->stream>>>(a);
    CGGP_LAUNCH_CHECK(ctx);
  }
  return CGGP_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Block-Jacobi factorisation: one CTA per block, in-place right-looking Cholesky in shared memory.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void block_cholesky_kernel(const T* __restrict__ A, int64_t lda, const int64_t* __restrict__ idx, int bs,
                                      T* __restrict__ chol) {
  __shared__ T s[MAX_BS][MAX_BS + 1];
  const int blk = blockIdx.x;
  const int64_t* id = idx + (int64_t)blk * bs;
  for (int e = threadIdx.x; e < bs * bs; e += blockDim.x) {
    const int i = e / bs, j = e % bs;
    s[i][j] = A[id[i] * lda + id[j]];
  }
  __syncthreads();
  for (int k = 0; k < bs; ++k) {
    if (threadIdx.x == 0) s[k][k] = sqrt(s[k][k]);
    __syncthreads();
    for (int i = k + 1 + threadIdx.x; i < bs; i += blockDim.x) s[i][k] /= s[k][k];
    __syncthreads();
    for (int e = threadIdx.x; e < (bs - k - 1) * (bs - k - 1); e += blockDim.x) {
      const int i = k + 1 + e / (bs - k - 1), j = k + 1 + e % (bs - k - 1);
      if (j <= i) s[i][j] -= s[i][k] * s[j][k];
    }
    __syncthreads();
  }
  T* L = chol + (int64_t)blk * bs * bs;
  for (int e = threadIdx.x; e < bs * bs; e += blockDim.x) {
    const int i = e / bs, j = e % bs;
    L[e] = j <= i ? s[i][j] : T(0);
  }
}

extern "C" int cggp_block_cholesky(cggp_ctx* ctx, int dtype, const void* A, int64_t lda, int64_t n,
                                   const int64_t* idx, int num_blocks, int block_size, void* chol) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (block_size < 1 || block_size > MAX_BS)
    CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "block_size %d outside [1, %d]", block_size, MAX_BS);
  if ((int64_t)num_blocks * block_size != n)
    CGGP_FAIL(ctx, CGGP_ERR_INVALID, "blocks (%d x %d) must partition n=%lld", num_blocks, block_size, (long long)n);
  if (dtype == CGGP_F64)
    block_cholesky_kernel<double><<<num_blocks, 256, 0, ctx->stream>>>((const double*)A, lda, idx, block_size,
                                                                      (double*)chol);
  else
    block_cholesky_kernel<float><<<num_blocks, 256, 0, ctx->stream>>>((const float*)A, lda, idx, block_size,
                                                                     (float*)chol);
  CGGP_LAUNCH_CHECK(ctx);
  return CGGP_OK;
}

// ---------------------------------------------------------------------------------------------------------
template <typename T>
static void fill_precond(StepArgs<T>& a, const cggp_precond* pc) {
  a.num_blocks = 0;
  a.block_size = 0;
  a.block_idx = nullptr;
  a.chol = nullptr;
  a.defer = 0;
  a.pinv = nullptr;
  a.ldpinv = 0;
  if (pc && pc->type == CGGP_PRECOND_DENSE) {
    a.defer = 1;
    a.pinv = (const T*)pc->dev_pinv;
    a.ldpinv = pc->ldpinv;
  }
  if (pc && pc->type == CGGP_PRECOND_BLOCK) {
    a.num_blocks = pc->num_blocks;
    a.block_size = pc->block_size;
    a.block_idx = pc->dev_block_indices;
    a.chol = (const T*)pc->dev_chol;
  }
}

static int check_precond(cggp_ctx* ctx, const cggp_precond* pc, int64_t n) {
  if (!pc || pc->type == CGGP_PRECOND_EYE) return CGGP_OK;
  if (pc->type == CGGP_PRECOND_DENSE) {
    if (!pc->dev_pinv || pc->ldpinv < n) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "dense preconditioner needs a [n, n] matrix");
    return CGGP_OK;
  }
  if (pc->type != CGGP_PRECOND_BLOCK) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "unknown preconditioner type %d", pc->type);
  if (pc->block_size < 1 || pc->block_size > MAX_BS)
    CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "block_size %d outside [1, %d]", pc->block_size, MAX_BS);
  if ((int64_t)pc->num_blocks * pc->block_size != n)
    CGGP_FAIL(ctx, CGGP_ERR_INVALID, "block preconditioner must partition n");
  if (!pc->dev_block_indices || !pc->dev_chol) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "block preconditioner buffers missing");
  return CGGP_OK;
}

template <typename T>
static int fused_step_impl(cggp_ctx* ctx, int B, int64_t n, const void* pA, void* v, void* r, void* p, void* rz,
                           void* half_rr, const cggp_precond* pc) {
  StepArgs<T> a{};
  a.mode = MODE_STEP;
  a.B = B;
  a.n = n;
  a.q = (const T*)pA;
  a.v = (T*)v;
  a.r = (T*)r;
  a.p = (T*)p;
  a.rz = (T*)rz;
  a.half_rr = (T*)half_rr;
  fill_precond(a, pc);
  if (a.num_blocks > 0 || a.defer) {
    int rc = cggp_ws2_reserve(ctx, sizeof(T) * (size_t)B * n);
    if (rc) return rc;
    a.z = (T*)cggp_ws2_ptr(ctx);
  }
  return launch_step(ctx, a);
}

// z = blockdiag(A)^-1 r for every row of r [B, n] (the preconditioner protocol `__call__(vec, mat) -> (z, rz)` of
// cggp/conjugate_gradient.py:125-128 outside the solve loop)
template <typename T>
__global__ void __launch_bounds__(512) block_precond_kernel(StepArgs<T> a) {
  const int b = blockIdx.x;
  block_precond_apply(a, a.r + (int64_t)b * a.n, a.z + (int64_t)b * a.n);
}

extern "C" int cggp_block_precond_apply(cggp_ctx* ctx, int dtype, int B, int64_t n, const void* r,
                                        const cggp_precond* pc, void* z) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (B <= 0 || n <= 0) return CGGP_OK;
  if (!pc || pc->type != CGGP_PRECOND_BLOCK) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "block preconditioner expected");
  int rc = check_precond(ctx, pc, n);
  if (rc) return rc;
  if (dtype == CGGP_F64) {
    StepArgs<double> a{};
    a.B = B; a.n = n; a.r = (double*)const_cast<void*>(r); a.z = (double*)z;
    fill_precond(a, pc);
    block_precond_kernel<double><<<B, 512, 0, ctx->stream>>>(a);
  } else {
    StepArgs<float> a{};
    a.B = B; a.n = n; a.r = (float*)const_cast<void*>(r); a.z = (float*)z;
    fill_precond(a, pc);
    block_precond_kernel<float><<<B, 512, 0, ctx->stream>>>(a);
  }
  CGGP_LAUNCH_CHECK(ctx);
  return CGGP_OK;
}

extern "C" int cggp_cg_fused_step(cggp_ctx* ctx, int dtype, int B, int64_t n, const void* pA, void* v, void* r,
                                  void* p, void* rz, void* half_rr, const cggp_precond* pc) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (B <= 0 || n <= 0) return CGGP_OK;
  int rc = check_precond(ctx, pc, n);
  if (rc) return rc;
  if (dtype == CGGP_F64) return fused_step_impl<double>(ctx, B, n, pA, v, r, p, rz, half_rr, pc);
  return fused_step_impl<float>(ctx, B, n, pA, v, r, p, rz, half_rr, pc);
}

// ---------------------------------------------------------------------------------------------------------
// Operator application  Y = V @ A  for the two operator forms
// ---------------------------------------------------------------------------------------------------------
// W = V @ (Kuf_r Kfu_r) on this rank's shard (not all-reduced)
static int apply_kuf_kfu(cggp_ctx* ctx, const cggp_operator* op, const void* V, int B, void* wbuf, const int* active) {
  const int64_t n = op->n;
  if (op->dtype == CGGP_F32 && op->dev_X32_big) {
    ProfScope prof(ctx, 0);
    return cggp_matvec_tf32(ctx, op->kind, op->variance, (const float*)op->dev_X32_big, (const float*)op->dev_X32_small,
                            (const float*)op->dev_x32_norms, op->n_local, (const float*)op->dev_Z32_big,
                            (const float*)op->dev_Z32_small, (const float*)op->dev_z32_norms, n, op->D, (const float*)V,
                            n, B, (float*)wbuf, n, (op->tf32_nsplit == 1 || op->tf32_nsplit == 16) ? op->tf32_nsplit : 3,
                            active);
  }
  return cggp_matvec_dispatch(ctx, op->dtype, op->kind, op->variance, op->dev_PX, op->dev_normsX, op->n_local,
                              op->dev_PZ, op->dev_normsZ, n, op->D, op->ldp, V, n, B, wbuf, n, op->variant, active);
}

static int apply_operator(cggp_ctx* ctx, const cggp_operator* op, const void* V, int B, void* Y, void* wbuf,
                          const int* active) {
  const int64_t n = op->n;
  if (op->type == CGGP_OP_DENSE)
    return cggp_symm_matmul_ex(ctx, op->dtype, op->dev_A, op->lda, n, V, n, B, Y, n, nullptr, 0, 0.0, active);
  // W = V @ (Kuf_r Kfu_r) on this rank's shard
  int rc;
  if (op->dtype == CGGP_F32 && op->dev_X32_big) {
    ProfScope prof(ctx, 0);
    rc = cggp_matvec_tf32(ctx, op->kind, op->variance, (const float*)op->dev_X32_big, (const float*)op->dev_X32_small,
                          (const float*)op->dev_x32_norms, op->n_local, (const float*)op->dev_Z32_big,
                          (const float*)op->dev_Z32_small, (const float*)op->dev_z32_norms, n, op->D, (const float*)V,
                          n, B, (float*)wbuf, n, (op->tf32_nsplit == 1 || op->tf32_nsplit == 16) ? op->tf32_nsplit : 3, active);
  } else
    rc = cggp_matvec_dispatch(ctx, op->dtype, op->kind, op->variance, op->dev_PX, op->dev_normsX, op->n_local,
                                op->dev_PZ, op->dev_normsZ, n, op->D, op->ldp, V, n, B, wbuf, n, op->variant, active);
  if (rc) return rc;
  // the one collective of the path: sum the partial [B, M] products over ranks
  rc = cggp_allreduce_sum(ctx, op->dtype, wbuf, (int64_t)B * n);
  if (rc) return rc;
  // Y = V @ Kuu + scale * W
  return cggp_symm_matmul_ex(ctx, op->dtype, op->dev_A, op->lda, n, V, n, B, Y, n, wbuf, n, op->scale, active);
}


// One non-refresh iteration on the matrix-free operator with the fused tail (see cg_tail_kernel).  The replicated term
// p @ Kuu is SHARDED over the ranks when the tail all-reduces over peer memory: rank r computes the columns
// [r n / W, (r + 1) n / W) only (1 / W of the HBM traffic of Kuu) and adds them, divided by sigma^-2, to its partial
// product before the exchange, so the one sum over ranks delivers p @ Sigma / sigma^-2.  (Overlapping the full product
// with the Kuf Kfu kernel on a side stream was measured first: it slowed that kernel by as much as it hid, 2.636 vs
// 2.607 ms at 8 GPUs.)
int cggp_symm_matmul_rows(cggp_ctx* ctx, int dtype, const void* A, int64_t lda, int64_t n, const void* V, int64_t ldv,
                          int B, void* Y, int64_t ldy, int64_t row_lo, int64_t row_hi, const int* active);

template <typename T>
static int fused_tail_iteration(cggp_ctx* ctx, const cggp_operator* op, StepArgs<T>& a, T* q, T* w, bool peer) {
  const int64_t n = op->n;
  const int B = a.B;
  TailArgs<T> t{};
  t.col_lo = 0;
  t.col_hi = n;
  static const int shard_env = getenv("CGGP_TAIL_SHARD") ? atoi(getenv("CGGP_TAIL_SHARD")) : 1;  // tuning knob
  t.shard = (peer && shard_env) ? 1 : 0;
  if (t.shard) {
    t.col_lo = n * ctx->rank / ctx->world;
    t.col_hi = n * (ctx->rank + 1) / ctx->world;
  }
  int rc = cggp_symm_matmul_rows(ctx, op->dtype, op->dev_A, op->lda, n, a.p, n, B, q, n, t.col_lo, t.col_hi, a.state);
  if (rc) return rc;
  rc = apply_kuf_kfu(ctx, op, a.p, B, w, a.state);
  if (rc) return rc;
  if (ctx->world > 1 && !peer) {
    rc = cggp_allreduce_sum(ctx, op->dtype, w, (int64_t)B * n);
    if (rc) return rc;
  }
  t.w = w;
  t.q = q;
  t.scale = (T)op->scale;
  t.inv_scale = (T)(1.0 / op->scale);
  t.world = peer ? ctx->world : 1;
  t.rank = ctx->rank;
  if (peer) {
    t.peers = (char* const*)ctx->peer_ptrs_dev;
    t.seq = ++ctx->peer_seq;
    t.slot_bytes = ctx->peer_slot_bytes;
    t.counter = ctx->peer_counter;
  }
  a.mode = MODE_STEP;
  a.q = q;
  ProfScope prof(ctx, 2);
  cg_tail_kernel<T><<<B, 512, 0, ctx->stream>>>(a, t);
  CGGP_LAUNCH_CHECK(ctx);
  return CGGP_OK;
}

template <typename T>
static int cg_solve_impl(cggp_ctx* ctx, const cggp_operator* op, const void* rhs, const void* x0, int B,
                         double error_threshold, int max_iterations, int max_steps_cycle, const cggp_precond* pc,
                         int check_every, void* solution, int32_t* host_steps, void* half_rz, void* history,
                         int64_t history_cap) {
  const int64_t n = op->n;
  const size_t vec = sizeof(T) * (size_t)B * n;
  const bool block = pc && (pc->type == CGGP_PRECOND_BLOCK || pc->type == CGGP_PRECOND_DENSE);  // needs a z buffer
  const bool sgpr = op->type == CGGP_OP_SGPR;
  // r, p, q (+ z) (+ w) and the per-row scalars
  size_t need = vec * (3 + (block ? 1 : 0) + (sgpr ? 1 : 0)) + sizeof(T) * 3 * (size_t)B + 256;
  int rc = cggp_ws2_reserve(ctx, need);
  if (rc) return rc;
  char* base = (char*)cggp_ws2_ptr(ctx);
  T* r = (T*)base; base += vec;
  T* p = (T*)base; base += vec;
  T* q = (T*)base; base += vec;
  T* z = nullptr;
  if (block) { z = (T*)base; base += vec; }
  T* w = nullptr;
  if (sgpr) { w = (T*)base; base += vec; }
  T* rz = (T*)base; base += sizeof(T) * B;
  T* half_rr = (T*)base; base += sizeof(T) * B;
  T* half_rz_int = (T*)base; base += sizeof(T) * B;
  T* v = (T*)solution;
  T* hrz = half_rz ? (T*)half_rz : half_rz_int;

  // fused tail: matrix-free operator, Eye preconditioner, few right-hand sides; over NVLink peer memory when the
  // peer buffers are mapped and the vector fits a slot, else after ncclAllReduce (CGGP_FUSED_TAIL=0 switches it off)
  static const int tail_env = getenv("CGGP_FUSED_TAIL") ? atoi(getenv("CGGP_FUSED_TAIL")) : 1;  // tuning knob
  const bool eye = !pc || pc->type == CGGP_PRECOND_EYE;
  const bool fused_tail = tail_env && sgpr && eye && B <= 8;
  const bool peer_tail = fused_tail && ctx->world > 1 && cggp_peer_enabled(ctx) && (int64_t)vec <= ctx->peer_slot_bytes;

  int* st = ctx->cg_state;
  ctx->cg_state_host[0] = 1;
  ctx->cg_state_host[1] = 0;
  ctx->cg_state_host[2] = 0;
  ctx->cg_state_host[3] = 0;
  CGGP_CUDA(ctx, cudaMemcpyAsync(st, ctx->cg_state_host, 4 * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CGGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the pinned staging buffer is reused below

  StepArgs<T> a{};
  a.B = B;
  a.n = n;
  a.rhs = (const T*)rhs;
  a.v = v;
  a.r = r;
  a.p = p;
  a.z = z;
  a.rz = rz;
  a.half_rr = half_rr;
  a.half_rz = hrz;
  a.history = (T*)history;
  a.history_cap = history ? history_cap : 0;
  a.threshold = (T)error_threshold;
  a.max_iterations = max_iterations;
  a.state = st;
  fill_precond(a, pc);

  // init (:87-92): r = b - v@A, (z, rz) = precond(r), p = z, i = 0
  if (x0) {
    CGGP_CUDA(ctx, cudaMemcpyAsync(v, x0, vec, cudaMemcpyDeviceToDevice, ctx->stream));
    rc = apply_operator(ctx, op, v, B, q, w, nullptr);
    if (rc) return rc;
  } else {
    // v = 0: v@A is exactly 0 for finite A, so the product is skipped (b - 0 == b bit for bit)
    CGGP_CUDA(ctx, cudaMemsetAsync(v, 0, vec, ctx->stream));
    CGGP_CUDA(ctx, cudaMemsetAsync(q, 0, vec, ctx->stream));
  }
  a.mode = MODE_INIT;
  a.q = q;
  rc = launch_step(ctx, a);
  if (rc) return rc;

  if (check_every < 1) check_every = 16;
  cudaEvent_t ev[2];
  CGGP_CUDA(ctx, cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
  CGGP_CUDA(ctx, cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
  int* hstage = ctx->cg_state_host;  // [0..2] chunk A, [4..6] chunk B
  int enq = 0;                       // iterations enqueued so far
  int chunk = 0;
  bool done = false;
  int pending[2] = {0, 0};
  while (!done) {
    const int slot = chunk & 1;
    // enqueue one chunk of iterations (skipped on the device once inactive)
    for (int c = 0; c < check_every && enq < max_iterations; ++c, ++enq) {
      const bool reset = (enq % max_steps_cycle) == (max_steps_cycle - 1);  // :71, pre-increment i
      if (fused_tail && !reset) {
        rc = fused_tail_iteration<T>(ctx, op, a, q, w, peer_tail);
        if (rc) goto fail;
        continue;
      }
      rc = apply_operator(ctx, op, p, B, q, w, st);
      if (rc) goto fail;
      a.q = q;
      if (!reset) {
        a.mode = MODE_STEP;
        rc = launch_step(ctx, a);
        if (rc) goto fail;
      } else {
        a.mode = MODE_PRE;
        rc = launch_step(ctx, a);
        if (rc) goto fail;
        rc = apply_operator(ctx, op, v, B, q, w, st);
        if (rc) goto fail;
        a.mode = MODE_POST;
        rc = launch_step(ctx, a);
        if (rc) goto fail;
      }
    }
    {
      cudaError_t e = cudaMemcpyAsync(hstage + 4 * slot, st, 3 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
      if (e == cudaSuccess) e = cudaEventRecord(ev[slot], ctx->stream);
      if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = CGGP_ERR_CUDA; goto fail; }
      pending[slot] = 1;
    }
    // look at the PREVIOUS chunk's flag while this one runs (keeps the GPU fed)
    const int prev = slot ^ 1;
    if (pending[prev]) {
      cudaError_t e = cudaEventSynchronize(ev[prev]);
      if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = CGGP_ERR_CUDA; goto fail; }
      pending[prev] = 0;
      if (hstage[4 * prev] == 0) done = true;
    }
    if (enq >= max_iterations) done = true;
    ++chunk;
  }
  {
    cudaError_t e = cudaMemcpyAsync(hstage, st, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = CGGP_ERR_CUDA; goto fail; }
    if (host_steps) *host_steps = hstage[1];
    if (hstage[3] != 0) {
      ctx->err = "fused CG tail: a peer rank did not publish its partial product in time (NVLink peer all-reduce)";
      rc = CGGP_ERR_COMM;
      goto fail;
    }
  }
  rc = CGGP_OK;
fail:
  cudaEventDestroy(ev[0]);
  cudaEventDestroy(ev[1]);
  return rc;
}

extern "C" int cggp_cg_solve(cggp_ctx* ctx, const cggp_operator* op, const void* rhs, const void* x0, int B,
                             double error_threshold, int max_iterations, int max_steps_cycle, const cggp_precond* pc,
                             int check_every, void* solution, int32_t* host_steps, void* half_rz, void* history,
                             int64_t history_cap) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (!op || !rhs || !solution) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "null operator / rhs / solution");
  if (op->struct_size != (uint32_t)sizeof(cggp_operator))
    CGGP_FAIL(ctx, CGGP_ERR_INVALID, "cggp_operator.struct_size is %u, this library expects %u (set it to sizeof of the "
              "header you compiled against; a shorter struct would be read past its end)", op->struct_size,
              (unsigned)sizeof(cggp_operator));
  if (B <= 0 || op->n <= 0) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "empty system (B=%d, n=%lld)", B, (long long)op->n);
  if (max_steps_cycle < 1) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "max_steps_cycle must be >= 1");
  if (max_iterations < 0) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "max_iterations must be >= 0");
  if (op->type != CGGP_OP_DENSE && op->type != CGGP_OP_SGPR) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "bad operator type");
  int rc = check_precond(ctx, pc, op->n);
  if (rc) return rc;
  ctx->solve_epoch = ++ctx->solve_counter;  // per-solve caches of the product kernels are valid until the return
  if (op->dtype == CGGP_F64)
    rc = cg_solve_impl<double>(ctx, op, rhs, x0, B, error_threshold, max_iterations, max_steps_cycle, pc,
                               check_every, solution, host_steps, half_rz, history, history_cap);
  else
    rc = cg_solve_impl<float>(ctx, op, rhs, x0, B, error_threshold, max_iterations, max_steps_cycle, pc, check_every,
                              solution, host_steps, half_rz, history, history_cap);
  ctx->solve_epoch = 0;
  return rc;
}
