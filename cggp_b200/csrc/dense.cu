// Y[B, n] = V[B, n] @ A[n, n] (+ scale * addend) for symmetric A: CG's `state.p @ A`
// (cggp/conjugate_gradient.py:65,74,87).  Symmetry lets every output column read a contiguous row of A:
//   Y[b, j] = sum_k V[b, k] * A[j, k].
// Small B (<= 8, the headline B = 1 and the 5-probe case) is HBM-bound on A: one warp per row of A, 16-byte loads,
// B accumulators, warp-shuffle reduction.  Larger B goes through the DMMA tile GEMM (dmma_gemm.cuh).
#include "common.cuh"
#include "dmma_gemm.cuh"

template <typename T>
struct Vec2;
template <>
struct Vec2<double> { using type = double2; };
template <>
struct Vec2<float> { using type = float2; };

// One warp per ROWS consecutive rows of A (ROWS x BB accumulators): the right-hand sides are read once per ROWS rows
// instead of once per row, and ROWS independent 16-byte loads per lane are in flight - short rows (M = 4096: 32 KB)
// are latency-bound otherwise.
template <typename T, int BB, int ROWS>
__global__ void __launch_bounds__(256)
symm_gemv_kernel(const T* __restrict__ A, int64_t lda, int64_t n, const T* __restrict__ V, int64_t ldv,
                 T* __restrict__ Y, int64_t ldy, const T* __restrict__ addend, int64_t ldadd, T scale,
                 const int* __restrict__ active) {
  if (cg_inactive(active)) return;
  const int lane = threadIdx.x & 31;
  const int64_t j0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ROWS;
  if (j0 >= n) return;
  const T* __restrict__ row[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) row[r] = A + (j0 + r < n ? j0 + r : n - 1) * lda;  // clamp: tail rows recompute n-1
  T acc[ROWS][BB];
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int b = 0; b < BB; ++b) acc[r][b] = T(0);
  using V2 = typename Vec2<T>::type;
  const bool vec_ok = ((lda | ldv) % 2 == 0) && ((((uintptr_t)A) | ((uintptr_t)V)) % (2 * sizeof(T)) == 0);
  if (vec_ok) {
    const int64_t n2 = n / 2;
    for (int64_t k = lane; k < n2; k += 32) {
      V2 a[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) a[r] = reinterpret_cast<const V2*>(row[r])[k];
#pragma unroll
      for (int b = 0; b < BB; ++b) {
        const V2 v = reinterpret_cast<const V2*>(V + b * ldv)[k];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          acc[r][b] = fma(a[r].x, v.x, acc[r][b]);
          acc[r][b] = fma(a[r].y, v.y, acc[r][b]);
        }
      }
    }
    if ((n & 1) && lane == 0) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int b = 0; b < BB; ++b) acc[r][b] = fma(row[r][n - 1], V[b * ldv + n - 1], acc[r][b]);
    }
  } else {
    for (int64_t k = lane; k < n; k += 32) {
#pragma unroll
      for (int b = 0; b < BB; ++b) {
        const T v = V[b * ldv + k];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc[r][b] = fma(row[r][k], v, acc[r][b]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
#pragma unroll
    for (int b = 0; b < BB; ++b) {
      T s = warp_sum(acc[r][b]);
      if (lane == 0 && j0 + r < n) {
        if (addend) s += scale * addend[b * ldadd + j0 + r];
        Y[b * ldy + j0 + r] = s;
      }
    }
}

template <typename T>
static int symm_matmul_impl(cggp_ctx* ctx, const T* A, int64_t lda, int64_t n, const T* V, int64_t ldv, int B, T* Y,
                            int64_t ldy, const T* addend, int64_t ldadd, T scale, const int* active) {
  int b0 = 0;
  if (B > 8) {
    // DMMA tile GEMM, NT form: C[b, j] = sum_k V[b, k] A[j, k]
    int rc = dmma_gemm_nt<T>(ctx, V, ldv, B, A, lda, n, n, Y, ldy, addend, ldadd, scale, active);
    return rc;
  }
  // rows per warp: enough warps to fill the machine first (n / ROWS >= ~8 warps per SM), then amortise the reads of V
  const int warps = 8;
  while (b0 < B) {
    const int bb = B - b0 < 8 ? B - b0 : 8;
    // measured on B200 (tools/bench_dense.py): one right-hand side streams best with 1 row per warp once there are
    // enough rows to fill the machine; several right-hand sides want their reads of V amortised over 2-4 rows
    const int rows = bb == 1 ? ((n >= 2048 && n < 8192) ? 2 : 1) : ((n >= 8192 && bb <= 4) ? 4 : (n >= 2048 ? 2 : 1));
    const int64_t nw = (n + rows - 1) / rows;
    const unsigned grid = (unsigned)((nw + warps - 1) / warps);
    const T* Vb = V + (int64_t)b0 * ldv;
    T* Yb = Y + (int64_t)b0 * ldy;
    const T* Ab = addend ? addend + (int64_t)b0 * ldadd : nullptr;
#define GEMV_R(BBV, R)                                                                                          \
  symm_gemv_kernel<T, BBV, R><<<grid, warps * 32, 0, ctx->stream>>>(A, lda, n, Vb, ldv, Yb, ldy, Ab, ldadd, scale, \
                                                                    active)
#define GEMV(BBV)                         \
  do {                                    \
    if (rows == 4 && BBV <= 4) {          \
      GEMV_R((BBV <= 4 ? BBV : 1), 4);    \
    } else if (rows >= 2) {               \
      GEMV_R(BBV, 2);                     \
    } else {                              \
      GEMV_R(BBV, 1);                     \
    }                                     \
  } while (0)
    switch (bb) {
      case 1: GEMV(1); break;
      case 2: GEMV(2); break;
      case 3: GEMV(3); break;
      case 4: GEMV(4); break;
      case 5: GEMV(5); break;
      case 6: GEMV(6); break;
      case 7: GEMV(7); break;
      default: GEMV(8); break;
    }
#undef GEMV
#undef GEMV_R
    CGGP_LAUNCH_CHECK(ctx);
    b0 += bb;
  }
  return CGGP_OK;
}

// internal entry used by the CG driver (adds the optional scaled addend and the loop-active flag)
int cggp_symm_matmul_ex(cggp_ctx* ctx, int dtype, const void* A, int64_t lda, int64_t n, const void* V, int64_t ldv,
                        int B, void* Y, int64_t ldy, const void* addend, int64_t ldadd, double scale,
                        const int* active) {
  ProfScope prof(ctx, 1);
  if (dtype == CGGP_F64)
    return symm_matmul_impl<double>(ctx, (const double*)A, lda, n, (const double*)V, ldv, B, (double*)Y, ldy,
                                    (const double*)addend, ldadd, scale, active);
  return symm_matmul_impl<float>(ctx, (const float*)A, lda, n, (const float*)V, ldv, B, (float*)Y, ldy,
                                 (const float*)addend, ldadd, (float)scale, active);
}

extern "C" int cggp_symm_matmul(cggp_ctx* ctx, int dtype, const void* A, int64_t lda, int64_t n, const void* V,
                                int64_t ldv, int B, void* Y, int64_t ldy) {
  if (!ctx) return CGGP_ERR_INVALID;
  if (B <= 0 || n <= 0) return CGGP_OK;
  return cggp_symm_matmul_ex(ctx, dtype, A, lda, n, V, ldv, B, Y, ldy, nullptr, 0, 0.0, nullptr);
}
