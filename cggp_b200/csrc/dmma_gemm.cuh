// FP64 "NT" GEMM on the DMMA path:  C[i, j] = sum_k A[i, k] * Bm[j, k]  (+ scale * addend[i, j])
// Both operands are row-major with K contiguous, which is exactly what mma.sync.m8n8k4.f64 wants for its
// row (A) and col (B) fragments.  Used for CG's multi-RHS `p @ A` on the dense Kuu + Lambda system
// (cggp/conjugate_gradient.py:65 with B = batch size / M right-hand sides; cggp/models.py:305,340).
// B200 facts this is built on (tools/microbench.cu, measured): every f64 mma shape lowers to DMMA.8x8x4,
// DMMA and DFMA share one pipe at 64 FMA/clk/SM (37 TFLOP/s), cuBLAS DGEMM reaches 35.4 TFLOP/s.
//
// CTA = 256 threads, tile (WM*16) x 128 x 16, warps 2 (rows) x 4 (cols), cp.async double buffering.
// float32 falls back to the FFMA tile engine (the tcgen05 TF32 path is the float32 successor, not in this round).
#pragma once
#include "common.cuh"
#include "tile.cuh"

namespace dg {
constexpr int BN = 128, BK = 32, LDS = 36;  // LDS: smem row stride in doubles, = 4 mod 8 -> conflict-free frags
constexpr int STAGES = 3;

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// rows: number of tile rows to load (BM or BN); 256 threads, 8-byte cp.async, zero fill outside the matrix
template <int ROWS>
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ G, int64_t ld, int64_t row0,
                                          int64_t nrows, int64_t k0, int64_t K) {
#pragma unroll
  for (int e = threadIdx.x; e < ROWS * BK; e += 256) {
    const int r = e / BK, k = e % BK;
    const int64_t gr = row0 + r, gk = k0 + k;
    const bool ok = gr < nrows && gk < K;
    cp_async8(s + r * LDS + k, ok ? (G + gr * ld + gk) : G, ok);
  }
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}

// Per-thread loader state for the 16-byte path: every thread owns ROWS / 32 fixed (row, k-pair) slots of the tile; the
// global pointers advance by BK per k-tile, so the steady state is one LDGSTS per slot with no address arithmetic or
// bounds logic (the issue port, not the DMMA pipe, is what the 8-byte loader was spending: 3.2 other instructions per
// DMMA = 81 % tensor-pipe utilisation, profiles/r01_final_kernels.md).
template <int ROWS>
struct TileLoader {
  static constexpr int SLOTS = ROWS * (BK / 2) / 256;
  const double* g[SLOTS];
  unsigned soff[SLOTS];   // offset in doubles inside a stage
  bool row_ok[SLOTS];
  int kk[SLOTS];
  __device__ __forceinline__ void init(const double* G, int64_t ld, int64_t row0, int64_t nrows) {
#pragma unroll
    for (int i = 0; i < SLOTS; ++i) {
      const int c = threadIdx.x + 256 * i;
      const int r = c / (BK / 2);
      kk[i] = (c % (BK / 2)) * 2;
      const int64_t gr = row0 + r;
      row_ok[i] = gr < nrows;
      g[i] = G + (row_ok[i] ? gr : 0) * ld + kk[i];
      soff[i] = r * LDS + kk[i];
    }
  }
  // k0: first k of this tile; K: total.  Full tiles copy 16 bytes, the ragged last one zero-fills past K.
  __device__ __forceinline__ void issue(double* stage, int64_t k0, int64_t K) {
#pragma unroll
    for (int i = 0; i < SLOTS; ++i) {
      int bytes = 16;
      if (k0 + BK > K) {
        const int64_t left = K - (k0 + kk[i]);
        bytes = left >= 2 ? 16 : (left == 1 ? 8 : 0);
      }
      if (!row_ok[i]) bytes = 0;
      cp_async16(stage + soff[i], bytes ? (const void*)(g[i] + k0) : (const void*)g[i], bytes);
    }
  }
};

template <int WMB>  // m8-blocks per warp along rows: 8 (BM = 128) or 4 (BM = 64)
__global__ void __launch_bounds__(256)
dmma_gemm_nt_kernel(const double* __restrict__ A, int64_t lda, int64_t M, const double* __restrict__ Bm, int64_t ldb,
                    int64_t N, int64_t K, double* __restrict__ C, int64_t ldc, const double* __restrict__ addend,
                    int64_t ldadd, double scale, const int* __restrict__ active, const int vec16, const int64_t k_chunk,
                    double* __restrict__ partial, const int lower) {
  if (cg_inactive(active)) return;
  constexpr int BM = WMB * 16;
  // symmetric rank-k update (C = A A^T): tiles entirely above the diagonal are skipped, the caller mirrors
  if (lower && (int64_t)blockIdx.x * BN >= ((int64_t)blockIdx.y + 1) * BM) return;
  if (gridDim.z > 1) {
    // split-K: this CTA contracts k in [z * k_chunk, min(K, (z + 1) * k_chunk)) and writes its partial [M, N] tile
    // set; a fixed-order reduction adds the partials (and the addend) afterwards
    const int64_t kb = (int64_t)blockIdx.z * k_chunk;
    A += kb;
    Bm += kb;
    K = (K - kb) < k_chunk ? (K - kb) : k_chunk;
    C = partial + (int64_t)blockIdx.z * M * N;
    ldc = N;
    addend = nullptr;
  }
  extern __shared__ __align__(16) double smem[];
  double* sA = smem;                        // [STAGES][BM][LDS]
  double* sB = smem + STAGES * BM * LDS;    // [STAGES][BN][LDS]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wi = warp >> 2, wj = warp & 3;  // 2 x 4 warps
  const int64_t row0 = (int64_t)blockIdx.y * BM, col0 = (int64_t)blockIdx.x * BN;
  const int fr = lane >> 2, fk = lane & 3;

  double acc[WMB][4][2];
#pragma unroll
  for (int a = 0; a < WMB; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  const int64_t nk = (K + BK - 1) / BK;
  TileLoader<BM> la;
  TileLoader<BN> lb;
  if (vec16) {
    la.init(A, lda, row0, M);
    lb.init(Bm, ldb, col0, N);
  }
  auto load_stage = [&](int s, int64_t k0) {
    if (vec16) {
      la.issue(sA + s * BM * LDS, k0, K);
      lb.issue(sB + s * BN * LDS, k0, K);
    } else {
      load_tile<BM>(sA + s * BM * LDS, A, lda, row0, M, k0, K);
      load_tile<BN>(sB + s * BN * LDS, Bm, ldb, col0, N, k0, K);
    }
  };
  // prologue
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage(s, (int64_t)s * BK);
    cp_async_commit();
  }
  for (int64_t kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {  // prefetch tile kt + STAGES - 1 into the slot freed at iteration kt - 1
      const int64_t nt = kt + STAGES - 1;
      if (nt < nk) load_stage((int)(nt % STAGES), nt * BK);
      cp_async_commit();
    }
    const int s = (int)(kt % STAGES);
    const double* a_s = sA + s * BM * LDS + (wi * WMB * 8 + fr) * LDS + fk;
    const double* b_s = sB + s * BN * LDS + (wj * 32 + fr) * LDS + fk;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      double af[WMB], bf[4];
#pragma unroll
      for (int a = 0; a < WMB; ++a) af[a] = a_s[a * 8 * LDS + ks * 4];
#pragma unroll
      for (int b = 0; b < 4; ++b) bf[b] = b_s[b * 8 * LDS + ks * 4];
#pragma unroll
      for (int a = 0; a < WMB; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
  }
  cp_async_wait<0>();
  // epilogue
#pragma unroll
  for (int a = 0; a < WMB; ++a) {
    const int64_t r = row0 + wi * WMB * 8 + a * 8 + fr;
    if (r >= M) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int64_t c = col0 + wj * 32 + b * 8 + 2 * fk;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (c + q < N) {
          double v = acc[a][b][q];
          if (addend) v += scale * addend[r * ldadd + c + q];
          C[r * ldc + c + q] = v;
        }
      }
    }
  }
}
}  // namespace dg

// FFMA / DFMA tile-engine GEMM (float32 path and cross-check)
template <typename T>
__global__ void __launch_bounds__(TILE_THREADS)
tile_gemm_nt_kernel(const T* __restrict__ A, int64_t lda, int64_t M, const T* __restrict__ Bm, int64_t ldb, int64_t N,
                    int64_t K, T* __restrict__ C, int64_t ldc, const T* __restrict__ addend, int64_t ldadd, T scale,
                    const int* __restrict__ active) {
  if (cg_inactive(active)) return;
  __shared__ TileSmem<T> s;
  const int64_t row0 = (int64_t)blockIdx.y * TILE, col0 = (int64_t)blockIdx.x * TILE;
  T acc[4][4];
  // K can exceed int range only for absurd sizes; the CG system size is an int
  tile_compute<T, 0>(acc, s, A, lda, row0, M, Bm, ldb, col0, N, (int)K);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = row0 + ty * 4 + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t c = col0 + tx * 4 + j;
      if (c >= N) continue;
      T v = acc[i][j];
      if (addend) v += scale * addend[r * ldadd + c];
      C[r * ldc + c] = v;
    }
  }
}

__global__ void splitk_reduce_kernel(const double* __restrict__ partial, int splits, int64_t M, int64_t N,
                                     double* __restrict__ C, int64_t ldc, const double* __restrict__ addend,
                                     int64_t ldadd, double scale, const int* __restrict__ active) {
  if (cg_inactive(active)) return;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= M * N) return;
  const int64_t r = e / N, c = e % N;
  double v = 0.0;
  for (int s = 0; s < splits; ++s) v += partial[(int64_t)s * M * N + e];
  if (addend) v += scale * addend[r * ldadd + c];
  C[r * ldc + c] = v;
}

// lower != 0 (float64 only): only the tiles touching the lower triangle are computed (A == Bm, M == N)
template <typename T>
int dmma_gemm_nt(cggp_ctx* ctx, const T* A, int64_t lda, int64_t M, const T* Bm, int64_t ldb, int64_t N, int64_t K,
                 T* C, int64_t ldc, const T* addend, int64_t ldadd, T scale, const int* active, int lower = 0);

template <>
inline int dmma_gemm_nt<float>(cggp_ctx* ctx, const float* A, int64_t lda, int64_t M, const float* Bm, int64_t ldb,
                               int64_t N, int64_t K, float* C, int64_t ldc, const float* addend, int64_t ldadd,
                               float scale, const int* active, int /*lower*/) {
  dim3 grid((unsigned)((N + TILE - 1) / TILE), (unsigned)((M + TILE - 1) / TILE));
  tile_gemm_nt_kernel<float><<<grid, TILE_THREADS, 0, ctx->stream>>>(A, lda, M, Bm, ldb, N, K, C, ldc, addend, ldadd,
                                                                      scale, active);
  CGGP_LAUNCH_CHECK(ctx);
  return CGGP_OK;
}

template <>
inline int dmma_gemm_nt<double>(cggp_ctx* ctx, const double* A, int64_t lda, int64_t M, const double* Bm, int64_t ldb,
                                int64_t N, int64_t K, double* C, int64_t ldc, const double* addend, int64_t ldadd,
                                double scale, const int* active, int lower) {
  using namespace dg;
  // 16-byte cp.async needs 16-byte aligned rows on both operands
  const int vec16 = ((lda | ldb) % 2 == 0) && ((((uintptr_t)A) | ((uintptr_t)Bm)) % 16 == 0) ? 1 : 0;
  // 64-row tiles when there are few rows, or when 128-row tiles would leave most of the machine idle (measured:
  // at half a wave and above the 128-row tile wins despite the quantisation, tools/bench_dense.py)
  const int64_t ctas128 = ((N + BN - 1) / BN) * ((M + 127) / 128);
  const bool small = M <= 64 || 2 * ctas128 < ctx->sm_count;
  const int BM = small ? 64 : 128;
  const size_t smem = (size_t)STAGES * (BM + BN) * LDS * sizeof(double);
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM));
  // split-K when the output tiles alone leave most SMs idle (9..~256 right-hand sides): k chunks of >= 256
  int64_t k_chunk = K;
  double* partial = nullptr;
  {
    const int64_t tiles = (int64_t)grid.x * grid.y;
    int64_t splits = tiles * 2 <= ctx->sm_count ? ctx->sm_count / tiles : 1;
    if (splits > K / 256) splits = K / 256;
    if (splits > 1) {
      k_chunk = ((K + splits - 1) / splits + BK - 1) / BK * BK;
      splits = (K + k_chunk - 1) / k_chunk;
    }
    if (splits > 1) {
      int rc = cggp_ws_reserve(ctx, sizeof(double) * (size_t)splits * (size_t)M * (size_t)N);
      if (rc) return rc;
      partial = (double*)ctx->ws;
      grid.z = (unsigned)splits;
    }
  }
  // the opt-in is per device and a process may drive several (one ctx each): set it on every launch (cheap)
  if (small) {
    CGGP_CUDA(ctx, cudaFuncSetAttribute(dmma_gemm_nt_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dmma_gemm_nt_kernel<4><<<grid, 256, smem, ctx->stream>>>(A, lda, M, Bm, ldb, N, K, C, ldc, addend, ldadd, scale,
                                                             active, vec16, k_chunk, partial, lower);
  } else {
    CGGP_CUDA(ctx, cudaFuncSetAttribute(dmma_gemm_nt_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dmma_gemm_nt_kernel<8><<<grid, 256, smem, ctx->stream>>>(A, lda, M, Bm, ldb, N, K, C, ldc, addend, ldadd, scale,
                                                             active, vec16, k_chunk, partial, lower);
  }
  CGGP_LAUNCH_CHECK(ctx);
  if (grid.z > 1) {
    const int64_t total = M * N;
    splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(partial, (int)grid.z, M, N, C, ldc,
                                                                                  addend, ldadd, scale, active);
    CGGP_LAUNCH_CHECK(ctx);
  }
  return CGGP_OK;
}
