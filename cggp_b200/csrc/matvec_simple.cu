// Simple two-sweep matrix-free product W = V @ (Kuf Kfu) (general shapes / float32 / cross-check for the fused
// DMMA kernel).  Sweep 1: T[b, i] = sum_j k(x_i, z_j) V[b, j].  Sweep 2: W[b, j] = sum_i k(x_i, z_j) T[b, i].
// Every Gram entry is evaluated twice (once per sweep) with library exp / sqrt; nothing of size N x M is stored.
// Deterministic: sweep 2 writes per-split partials that a second kernel sums in fixed order.
#include "common.cuh"
#include "kmath.cuh"
#include "tile.cuh"

constexpr int SB = 4;  // right-hand sides per sweep

template <typename T, int KIND>
__global__ void __launch_bounds__(TILE_THREADS)
kfu_sweep1_kernel(const T* __restrict__ PX, const T* __restrict__ nX, int64_t n, const T* __restrict__ PZ,
                  const T* __restrict__ nZ, int64_t m, int D, int64_t ldp, T variance, const T* __restrict__ V,
                  int64_t ldv, int nb, T* __restrict__ Tb /*[SB, n]*/, const int* __restrict__ active) {
  if (cg_inactive(active)) return;
  __shared__ TileSmem<T> s;
  __shared__ T vs[SB][TILE];
  __shared__ T nzs[TILE];
  const int64_t row0 = (int64_t)blockIdx.x * TILE;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  T na[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = row0 + ty * 4 + i;
    na[i] = r < n ? nX[r] : T(0);
  }
  T tp[4][SB];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int b = 0; b < SB; ++b) tp[i][b] = T(0);
  for (int64_t col0 = 0; col0 < m; col0 += TILE) {
    T acc[4][4];
    tile_compute<T, 0>(acc, s, PX, ldp, row0, n, PZ, ldp, col0, m, D);
    __syncthreads();
    for (int e = threadIdx.x; e < SB * TILE; e += TILE_THREADS) {
      const int b = e / TILE, c = e % TILE;
      vs[b][c] = (b < nb && col0 + c < m) ? V[(int64_t)b * ldv + col0 + c] : T(0);
    }
    if (threadIdx.x < TILE) nzs[threadIdx.x] = (col0 + threadIdx.x < m) ? nZ[col0 + threadIdx.x] : T(0);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = tx * 4 + j;
      const bool ok = col0 + c < m;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const T r2 = (T(-2) * acc[i][j]) + (na[i] + nzs[c]);
        const T k = ok ? kernel_value<T, KIND>(r2, variance) : T(0);
#pragma unroll
        for (int b = 0; b < SB; ++b) tp[i][b] = fma(k, vs[b][c], tp[i][b]);
      }
    }
  }
  // reduce over the 16 tx-threads of each row (one half-warp)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int b = 0; b < SB; ++b) {
      T v = tp[i][b];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const int64_t r = row0 + ty * 4 + i;
      if (tx == 0 && r < n && b < nb) Tb[(int64_t)b * n + r] = v;
    }
}

template <typename T, int KIND>
__global__ void __launch_bounds__(TILE_THREADS)
kfu_sweep2_kernel(const T* __restrict__ PX, const T* __restrict__ nX, int64_t n, const T* __restrict__ PZ,
                  const T* __restrict__ nZ, int64_t m, int D, int64_t ldp, T variance, const T* __restrict__ Tb,
                  int nb, int64_t rows_per_split, T* __restrict__ Wp /*[splits, SB, m]*/,
                  const int* __restrict__ active) {
  if (cg_inactive(active)) return;
  __shared__ TileSmem<T> s;
  __shared__ T ts[SB][TILE];
  __shared__ T nxs[TILE];
  __shared__ T red[8][SB][TILE + 1];
  const int64_t col0 = (int64_t)blockIdx.x * TILE;
  const int split = blockIdx.y;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  T nb_[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t c = col0 + tx * 4 + j;
    nb_[j] = c < m ? nZ[c] : T(0);
  }
  T wp[4][SB];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int b = 0; b < SB; ++b) wp[j][b] = T(0);
  const int64_t rbeg = (int64_t)split * rows_per_split;
  int64_t rend = rbeg + rows_per_split;
  if (rend > n) rend = n;
  for (int64_t row0 = rbeg; row0 < rend; row0 += TILE) {
    T acc[4][4];
    // A-rows = X rows (limited to this split), B-rows = Z rows
    tile_compute<T, 0>(acc, s, PX, ldp, row0, rend, PZ, ldp, col0, m, D);
    __syncthreads();
    for (int e = threadIdx.x; e < SB * TILE; e += TILE_THREADS) {
      const int b = e / TILE, r = e % TILE;
      ts[b][r] = (b < nb && row0 + r < rend) ? Tb[(int64_t)b * n + row0 + r] : T(0);
    }
    if (threadIdx.x < TILE) nxs[threadIdx.x] = (row0 + threadIdx.x < rend) ? nX[row0 + threadIdx.x] : T(0);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const bool ok = row0 + r < rend;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const T r2 = (T(-2) * acc[i][j]) + (nxs[r] + nb_[j]);
        const T k = ok ? kernel_value<T, KIND>(r2, variance) : T(0);
#pragma unroll
        for (int b = 0; b < SB; ++b) wp[j][b] = fma(k, ts[b][r], wp[j][b]);
      }
    }
  }
  // reduce over ty (16 threads per column): the two ty of a warp by shuffle, the 8 warps through shared memory
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int b = 0; b < SB; ++b) {
      T v = wp[j][b];
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((threadIdx.x & 16) == 0) red[threadIdx.x >> 5][b][tx * 4 + j] = v;
    }
  __syncthreads();
  for (int e = threadIdx.x; e < SB * TILE; e += TILE_THREADS) {
    const int b = e / TILE, c = e % TILE;
    T v = T(0);
#pragma unroll
    for (int y = 0; y < 8; ++y) v += red[y][b][c];
    if (b < nb && col0 + c < m) Wp[((int64_t)split * SB + b) * m + col0 + c] = v;
  }
}

template <typename T>
__global__ void reduce_splits_kernel(const T* __restrict__ Wp, int splits, int nb, int64_t m, T* __restrict__ W,
                                     int64_t ldw, const int* __restrict__ active) {
  if (cg_inactive(active)) return;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= m || b >= nb) return;
  T v = T(0);
  for (int s = 0; s < splits; ++s) v += Wp[((int64_t)s * SB + b) * m + c];
  W[(int64_t)b * ldw + c] = v;
}

template <typename T>
static int matvec_simple_impl(cggp_ctx* ctx, int kind, double variance, const T* PX, const T* nX, int64_t n,
                              const T* PZ, const T* nZ, int64_t m, int D, int64_t ldp, const T* V, int64_t ldv, int B,
                              T* W, int64_t ldw, const int* active) {
  if (n == 0) {  // an empty shard contributes exactly zero
    for (int b = 0; b < B; ++b) CGGP_CUDA(ctx, cudaMemsetAsync(W + (int64_t)b * ldw, 0, sizeof(T) * m, ctx->stream));
    return CGGP_OK;
  }
  const int64_t row_tiles = (n + TILE - 1) / TILE;
  const int64_t col_tiles = (m + TILE - 1) / TILE;
  // enough (col tile, split) CTAs to fill the machine a few times over
  int64_t splits = (4LL * ctx->sm_count + col_tiles - 1) / col_tiles;
  if (splits > row_tiles) splits = row_tiles;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t rows_per_split = ((row_tiles + splits - 1) / splits) * TILE;
  splits = (n + rows_per_split - 1) / rows_per_split;
  if (splits < 1) splits = 1;
  const size_t need = sizeof(T) * ((size_t)SB * (size_t)(n > 0 ? n : 1) + (size_t)splits * SB * (size_t)m);
  int rc = cggp_ws_reserve(ctx, need);
  if (rc) return rc;
  T* Tb = (T*)ctx->ws;
  T* Wp = Tb + (size_t)SB * (size_t)(n > 0 ? n : 1);
  for (int b0 = 0; b0 < B; b0 += SB) {
    const int nb = (B - b0) < SB ? (B - b0) : SB;
    const T* Vb = V + (int64_t)b0 * ldv;
    T* Wb = W + (int64_t)b0 * ldw;
    if (n == 0) {
      for (int b = 0; b < nb; ++b) CGGP_CUDA(ctx, cudaMemsetAsync(Wb + (int64_t)b * ldw, 0, sizeof(T) * m, ctx->stream));
      continue;
    }
#define SWEEPS(KV)                                                                                                  \
  kfu_sweep1_kernel<T, KV><<<(unsigned)row_tiles, TILE_THREADS, 0, ctx->stream>>>(PX, nX, n, PZ, nZ, m, D, ldp,     \
                                                                                  (T)variance, Vb, ldv, nb, Tb,     \
                                                                                  active);                          \
  CGGP_LAUNCH_CHECK(ctx);                                                                                           \
  kfu_sweep2_kernel<T, KV><<<dim3((unsigned)col_tiles, (unsigned)splits), TILE_THREADS, 0, ctx->stream>>>(          \
      PX, nX, n, PZ, nZ, m, D, ldp, (T)variance, Tb, nb, rows_per_split, Wp, active);                               \
  CGGP_LAUNCH_CHECK(ctx);
    switch (kind) {
      case CGGP_SE: SWEEPS(CGGP_SE) break;
      case CGGP_MATERN12: SWEEPS(CGGP_MATERN12) break;
      case CGGP_MATERN32: SWEEPS(CGGP_MATERN32) break;
      case CGGP_MATERN52: SWEEPS(CGGP_MATERN52) break;
      default: CGGP_FAIL(ctx, CGGP_ERR_INVALID, "unknown kernel kind %d", kind);
    }
#undef SWEEPS
    reduce_splits_kernel<T><<<dim3((unsigned)((m + 255) / 256), (unsigned)nb), 256, 0, ctx->stream>>>(
        Wp, (int)splits, nb, m, Wb, ldw, active);
    CGGP_LAUNCH_CHECK(ctx);
  }
  return CGGP_OK;
}

int cggp_matvec_simple(cggp_ctx* ctx, int dtype, int kind, double variance, const void* PX, const void* nX, int64_t n,
                       const void* PZ, const void* nZ, int64_t m, int D, int64_t ldp, const void* V, int64_t ldv,
                       int B, void* W, int64_t ldw, const int* active) {
  if (dtype == CGGP_F64)
    return matvec_simple_impl<double>(ctx, kind, variance, (const double*)PX, (const double*)nX, n, (const double*)PZ,
                                      (const double*)nZ, m, D, ldp, (const double*)V, ldv, B, (double*)W, ldw, active);
  return matvec_simple_impl<float>(ctx, kind, variance, (const float*)PX, (const float*)nX, n, (const float*)PZ,
                                   (const float*)nZ, m, D, ldp, (const float*)V, ldv, B, (float*)W, ldw, active);
}
