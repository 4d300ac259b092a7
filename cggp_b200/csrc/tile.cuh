// Generic 64x64 register-tiled engine on "prepared" point sets (row-major, K-contiguous rows): every thread of a
// 256-thread CTA owns a 4x4 micro-tile of pairwise quantities between 64 A-rows and 64 B-rows, accumulated over
// the feature axis in chunks of 16 staged (transposed) through shared memory.  Used by the general-purpose
// kernels (kernel matrix, distances, nearest centre, the simple two-sweep matvec, large-B dense products); the
// headline fused matvec has its own DMMA pipeline (matvec_pipe.cu, matvec_pipe8.cu).
#pragma once
#include "common.cuh"

constexpr int TILE = 64;
constexpr int TILE_DC = 16;
constexpr int TILE_LD = 66;  // padded: transposed stores are <= 2-way bank conflicted, float4/double2 reads aligned
constexpr int TILE_THREADS = 256;

template <typename T>
struct TileSmem {
  T a[TILE_DC][TILE_LD];
  T b[TILE_DC][TILE_LD];
};

template <typename T>
__device__ __forceinline__ void tile_load(T (*s)[TILE_LD], const T* __restrict__ P, int64_t ld, int64_t row0,
                                          int64_t nrows, int d0, int dmax) {
  for (int e = threadIdx.x; e < TILE * TILE_DC; e += TILE_THREADS) {
    const int r = e >> 4, d = e & 15;
    const int64_t gr = row0 + r;
    T v = T(0);
    if (gr < nrows && d0 + d < dmax) v = P[gr * ld + d0 + d];
    s[d][r] = v;
  }
}

// MODE 0: acc[i][j] += sum_d a_i[d] * b_j[d]     MODE 1: acc[i][j] += sum_d (a_i[d] - b_j[d])^2
template <typename T, int MODE>
__device__ __forceinline__ void tile_accumulate(T acc[4][4], const TileSmem<T>& s, int dc) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  for (int d = 0; d < dc; ++d) {
    T av[4], bv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) av[i] = s.a[d][ty * 4 + i];
#pragma unroll
    for (int j = 0; j < 4; ++j) bv[j] = s.b[d][tx * 4 + j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (MODE == 0) {
          acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        } else {
          T df = av[i] - bv[j];
          acc[i][j] = fma(df, df, acc[i][j]);
        }
      }
  }
}

// Full feature-axis accumulation for the tile (rowA0.., rowB0..).
template <typename T, int MODE>
__device__ __forceinline__ void tile_compute(T acc[4][4], TileSmem<T>& s, const T* __restrict__ PA, int64_t lda,
                                             int64_t rowA0, int64_t nA, const T* __restrict__ PB, int64_t ldb,
                                             int64_t rowB0, int64_t nB, int D) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  for (int d0 = 0; d0 < D; d0 += TILE_DC) {
    __syncthreads();
    tile_load(s.a, PA, lda, rowA0, nA, d0, D);
    tile_load(s.b, PB, ldb, rowB0, nB, d0, D);
    __syncthreads();
    const int dc = (D - d0) < TILE_DC ? (D - d0) : TILE_DC;
    tile_accumulate<T, MODE>(acc, s, dc);
  }
}
