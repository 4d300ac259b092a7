// Y[B, n] = V[B, n] @ A[n, n] (+ scale * addend) for symmetric A: CG's `state.p @ A`
// (cggp/conjugate_gradient.py:65,74,87).  Symmetry lets every output column read a contiguous row of A:
//   Y[b, j] = sum_k V[b, k] * A[j, k].
// Small B (<= 8, the headline B = 1 and the 5-probe case) is HBM-bound on A: one warp per row of A, 16-byte loads,
// B accumulators, warp-shuffle reduction.  Larger B goes through the DMMA tile GEMM (dmma_gemm.cuh).
#include <cstdlib>

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
symm_gemv_kernel(const T* __restrict__ A, int64_t lda, int64_t n, int64_t nrows, const T* __restrict__ V, int64_t ldv,
                 T* __restrict__ Y, int64_t ldy, const T* __restrict__ addend, int64_t ldadd, T scale,
                 const int* __restrict__ active) {
  // A points at the first of `nrows` rows (all n columns of each are contracted); Y / addend at their first column
  if (cg_inactive(active)) return;
  const int lane = threadIdx.x & 31;
  const int64_t j0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * ROWS;
  if (j0 >= nrows) return;
  const T* __restrict__ row[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
    row[r] = A + (j0 + r < nrows ? j0 + r : nrows - 1) * lda;  // clamp: tail rows recompute the last one
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
      if (lane == 0 && j0 + r < nrows) {
        if (addend) s = fma(scale, addend[b * ldadd + j0 + r], s);
        Y[b * ldy + j0 + r] = s;
      }
    }
}

template <typename T>
static int symm_matmul_impl(cggp_ctx* ctx, const T* A, int64_t lda, int64_t n, const T* V, int64_t ldv, int B, T* Y,
                            int64_t ldy, const T* addend, int64_t ldadd, T scale, const int* active,
                            int64_t row_lo = 0, int64_t row_hi = -1) {
  // [row_lo, row_hi): only these output columns Y[:, j] (= rows j of the symmetric A) are computed (B <= 8 path)
  if (row_hi < 0) row_hi = n;
  const int64_t nrows = row_hi - row_lo;
  if (nrows <= 0) return CGGP_OK;
  A += row_lo * lda;
  Y += row_lo;
  if (addend) addend += row_lo;
  int b0 = 0;
  if (B > 8) {
    if (nrows != n) CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "row-range symmetric product needs B <= 8");
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
    const int64_t nw = (nrows + rows - 1) / rows;
    const unsigned grid = (unsigned)((nw + warps - 1) / warps);
    const T* Vb = V + (int64_t)b0 * ldv;
    T* Yb = Y + (int64_t)b0 * ldy;
    const T* Ab = addend ? addend + (int64_t)b0 * ldadd : nullptr;
#define GEMV_R(BBV, R)                                                                                          \
  symm_gemv_kernel<T, BBV, R><<<grid, warps * 32, 0, ctx->stream>>>(A, lda, n, nrows, Vb, ldv, Yb, ldy, Ab, ldadd,  \
                                                                    scale, active)
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
int cggp_symm_matmul_rows(cggp_ctx* ctx, int dtype, const void* A, int64_t lda, int64_t n, const void* V, int64_t ldv,
                          int B, void* Y, int64_t ldy, int64_t row_lo, int64_t row_hi, const int* active) {
  ProfScope prof(ctx, 1);
  if (dtype == CGGP_F64)
    return symm_matmul_impl<double>(ctx, (const double*)A, lda, n, (const double*)V, ldv, B, (double*)Y, ldy, nullptr,
                                    0, 0.0, active, row_lo, row_hi);
  return symm_matmul_impl<float>(ctx, (const float*)A, lda, n, (const float*)V, ldv, B, (float*)Y, ldy, nullptr, 0,
                                 0.0f, active, row_lo, row_hi);
}

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
  CGGP_DEVICE_GUARD(ctx);
  if (B <= 0 || n <= 0) return CGGP_OK;
  return cggp_symm_matmul_ex(ctx, dtype, A, lda, n, V, ldv, B, Y, ldy, nullptr, 0, 0.0, nullptr);
}

// ---------------------------------------------------------------------------------------------------------
// G (+)= Kuf Kfu over this rank's rows: the [M, M] Gram matrix GPflow's SGPR materialises as A A^T.
// Kuf is evaluated in row chunks sized to stay L2-resident (cggp_kernel_matrix -> [M, nc], never more than one chunk
// alive) and contracted by the DMMA GEMM of this library as a symmetric rank-k update: only tiles touching the lower
// triangle are computed, the result is mirrored once at the end.  (Evaluating the Gram entries INSIDE the GEMM tiles
// would re-evaluate every entry M / 256 times: 27 FP64 slots per evaluation against 128 FMAs of use per tile pass is
// +21 % on the binding pipe, against +1 % for the chunked form - DESIGN.md 4.3.)
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void mirror_lower_kernel(T* __restrict__ G, int64_t ldg, int64_t m) {
  __shared__ T tile[32][33];
  const int64_t bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;  // read lower tile (bi, bj), write upper tile (bj, bi)
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = bi * 32 + r, j = bj * 32 + tx;
    tile[r][tx] = (i < m && j < m) ? G[i * ldg + j] : T(0);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t i = bj * 32 + r, j = bi * 32 + tx;  // upper element (i, j) = lower element (j, i)
    if (i < m && j < m && j > i) G[i * ldg + j] = tile[tx][r];
  }
}

int cggp_ws2_reserve(cggp_ctx* ctx, size_t bytes);
void* cggp_ws2_ptr(cggp_ctx* ctx);

extern "C" int cggp_kuf_gram(cggp_ctx* ctx, int dtype, int kind, double variance, const void* PX, const void* nX,
                             int64_t n, const void* PZ, const void* nZ, int64_t m, int D, int64_t ldp, void* G,
                             int64_t ldg, int accumulate) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (m <= 0) return CGGP_OK;
  if (!G || ldg < m) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "gram: output must be [m, m] with ldg >= m");
  const size_t es = dtype == CGGP_F64 ? 8 : 4;
  if (!accumulate) CGGP_CUDA(ctx, cudaMemset2DAsync(G, (size_t)ldg * es, 0, (size_t)m * es, (size_t)m, ctx->stream));
  if (n <= 0) return CGGP_OK;
  // rows per chunk: 64 MB of Kuf, but never fewer than 4096 rows - the rank-k update then runs 128 k-tiles per output
  // tile and the GEMM's prologue / epilogue is amortised (M = 16384, executed TFLOP/s with 512 / 2048 / 4096-row chunks:
  // 23.6 / 25.7 / 28.9, gpurun_out/r2_gram10.log; 537 MB of scratch at that size)
  static const int64_t nc_env = getenv("CGGP_GRAM_ROWS") ? atoll(getenv("CGGP_GRAM_ROWS")) : 0;  // tuning knob
  int64_t nc = (int64_t)((size_t)(1u << 26) / ((size_t)m * es));
  nc = nc < 4096 ? 4096 : nc;
  if (nc_env > 0) nc = nc_env;
  nc = (nc + 31) / 32 * 32;
  if (nc > n) nc = (n + 1) / 2 * 2;
  int rc = cggp_ws2_reserve(ctx, (size_t)m * (size_t)nc * es);
  if (rc) return rc;
  char* Kc = (char*)cggp_ws2_ptr(ctx);
  for (int64_t s = 0; s < n; s += nc) {
    const int64_t rows = n - s < nc ? n - s : nc;
    rc = cggp_kernel_matrix(ctx, dtype, kind, variance, CGGP_OUT_KERNEL, 0, PZ, nZ, m, (const char*)PX + (size_t)s * ldp * es,
                            (const char*)nX + (size_t)s * es, rows, D, ldp, 0.0, Kc, nc);
    if (rc) return rc;
    if (dtype == CGGP_F64)
      rc = dmma_gemm_nt<double>(ctx, (const double*)Kc, nc, m, (const double*)Kc, nc, m, rows, (double*)G, ldg,
                                (const double*)G, ldg, 1.0, nullptr, 1);
    else
      rc = dmma_gemm_nt<float>(ctx, (const float*)Kc, nc, m, (const float*)Kc, nc, m, rows, (float*)G, ldg,
                               (const float*)G, ldg, 1.0f, nullptr, 0);
    if (rc) return rc;
  }
  if (dtype == CGGP_F64) {
    dim3 grid((unsigned)((m + 31) / 32), (unsigned)((m + 31) / 32));
    mirror_lower_kernel<double><<<grid, 256, 0, ctx->stream>>>((double*)G, ldg, m);
    CGGP_LAUNCH_CHECK(ctx);
  }
  return CGGP_OK;
}
