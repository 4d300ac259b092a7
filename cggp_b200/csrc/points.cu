// Prepared points, kernel / distance matrices, nearest-centre assignment and cluster statistics.
// Reference behaviour: GPflow Stationary.scale + square_distance + K_r2 (call sites cggp/models.py:300,333-335),
// cggp/distance.py:9-34, cggp/selection.py:14-32, cggp/optimize.py:50-67,88-96.
#include "common.cuh"
#include "kmath.cuh"
#include "tile.cuh"

// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void prepare_points_kernel(const T* __restrict__ X, int64_t n, int D, int64_t ldx, LsParam ls,
                                      T* __restrict__ P, int64_t ldp, T* __restrict__ norms) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T s = T(0);
  for (int d = 0; d < D; ++d) {
    const T l = (T)ls.v[ls.count == 1 ? 0 : d];
    const T v = X[i * ldx + d] / l;  // true division, as GPflow's X / lengthscales
    P[i * ldp + d] = v;
    s += v * v;                      // reduce_sum(square(.)) in feature order
  }
  // spare column D holds 1 (it pairs with alpha |z|^2 in the fused matvec's DMMA), the rest of the padding is 0
  for (int d = D; d < ldp; ++d) P[i * ldp + d] = d == D ? T(1) : T(0);
  norms[i] = s;
}

template <typename T, int KIND>
__device__ __forceinline__ T pair_value(int output, int distance, T acc, T na, T nb, T variance) {
  if (output == CGGP_OUT_DISTANCE && distance == CGGP_DIST_EUCLIDEAN) return xsqrt(acc);  // acc = sum (a-b)^2
  const T r2 = (T(-2) * acc) + (na + nb);  // GPflow: dist = -2 a.b; dist += |a|^2 + |b|^2
  if (output == CGGP_OUT_DISTANCE && distance == CGGP_DIST_SQEUCLIDEAN) return r2;
  const T k = kernel_value<T, KIND>(r2, variance);
  if (output == CGGP_OUT_KERNEL) return k;
  if (distance == CGGP_DIST_COVARIANCE) return (variance + variance) - T(2) * k;  // distance.py:21
  return T(1) - k / xsqrt(variance * variance);                                      // distance.py:30
}

template <typename T, int KIND, int MODE>
__global__ void __launch_bounds__(TILE_THREADS)
kernel_matrix_kernel(const T* __restrict__ PA, const T* __restrict__ nA, int64_t n, const T* __restrict__ PB,
                     const T* __restrict__ nB, int64_t m, int D, int64_t ldp, T variance, int output, int distance,
                     T jitter, T* __restrict__ out, int64_t ldo) {
  __shared__ TileSmem<T> s;
  const int64_t row0 = (int64_t)blockIdx.y * TILE, col0 = (int64_t)blockIdx.x * TILE;
  T acc[4][4];
  tile_compute<T, MODE>(acc, s, PA, ldp, row0, n, PB, ldp, col0, m, D);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = row0 + ty * 4 + i;
    if (r >= n) continue;
    const T na = nA[r];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t c = col0 + tx * 4 + j;
      if (c >= m) continue;
      T v = pair_value<T, KIND>(output, distance, acc[i][j], na, nB[c], variance);
      if (r == c) v += jitter;
      out[r * ldo + c] = v;
    }
  }
}

// One CTA per 64 data rows, sweeping all centres; running (min, argmin) per row, first minimum wins (tf.argmin).
template <typename T, int KIND, int MODE>
__global__ void __launch_bounds__(TILE_THREADS)
nearest_center_kernel(const T* __restrict__ PX, const T* __restrict__ nX, int64_t n, const T* __restrict__ PZ,
                      const T* __restrict__ nZ, int64_t m, int D, int64_t ldp, T variance, int distance,
                      int64_t* __restrict__ idx, T* __restrict__ dist) {
  __shared__ TileSmem<T> s;
  const int64_t row0 = (int64_t)blockIdx.x * TILE;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  T best[4];
  int64_t bidx[4];
  T na[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    best[i] = T(INFINITY);
    bidx[i] = 0;
    const int64_t r = row0 + ty * 4 + i;
    na[i] = r < n ? nX[r] : T(0);
  }
  for (int64_t col0 = 0; col0 < m; col0 += TILE) {
    T acc[4][4];
    tile_compute<T, MODE>(acc, s, PX, ldp, row0, n, PZ, ldp, col0, m, D);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t c = col0 + tx * 4 + j;
      if (c >= m) continue;
      const T nb = nZ[c];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const T v = pair_value<T, KIND>(CGGP_OUT_DISTANCE, distance, acc[i][j], na[i], nb, variance);
        if (v < best[i]) {  // columns are visited in increasing order per thread: strict < keeps the first minimum
          best[i] = v;
          bidx[i] = c;
        }
      }
    }
  }
  // merge the 16 tx-threads of a row (they sit in one half-warp): smaller value wins, ties -> smaller index
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const T ov = __shfl_xor_sync(0xffffffffu, best[i], o);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, bidx[i], o);
      if (ov < best[i] || (ov == best[i] && oi < bidx[i])) {
        best[i] = ov;
        bidx[i] = oi;
      }
    }
    const int64_t r = row0 + ty * 4 + i;
    if (tx == 0 && r < n) {
      idx[r] = bidx[i];
      dist[r] = best[i];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Backward of the kernel matrix w.r.t. the hyper-parameters (what TF autodiff does behind the
// reference's Adam loop, cggp/optimize.py:198-254): given G = dL/dK [n, m],
//   dL/dvariance = sum_ij G_ij f(r2_ij),            K = variance * f(r2),  r2 = sum_d (a_d - b_d)^2,  a = x / l
//   dL/dl_d      = sum_ij G_ij variance f'(r2_ij) * (-2 (a_id - b_jd)^2 / l_d)
// Per-CTA partials [blocks][1 + D] are summed in fixed order by a second kernel (deterministic).
// f' = df/dr2:  SE -f/2;  Matern-1/2 -exp(-r)/(2 r);  3/2 -(3/2) exp(-s);  5/2 -(5/6)(1 + s) exp(-s)   (s = sqrt(nu2) r);
// 0 where GPflow's max(r2, 1e-36) clamps (autodiff through the max gives no gradient there).
template <typename T, int KIND>
__device__ __forceinline__ void kernel_f_and_dr2(T r2, T& f, T& fp) {
  if (KIND == CGGP_SE) {
    f = xexp(T(-0.5) * r2);
    fp = T(-0.5) * f;
    return;
  }
  const bool clamped = !(r2 > KConst<T>::clamp());
  const T r = xsqrt(xmax(r2, KConst<T>::clamp()));
  if (KIND == CGGP_MATERN12) {
    f = xexp(-r);
    fp = clamped ? T(0) : -f / (T(2) * r);
  } else if (KIND == CGGP_MATERN32) {
    const T s = KConst<T>::sqrt3() * r;
    const T e = xexp(-s);
    f = (T(1) + s) * e;
    fp = clamped ? T(0) : T(-1.5) * e;
  } else {
    const T s = KConst<T>::sqrt5() * r;
    const T e = xexp(-s);
    f = (T(1) + s + KConst<T>::c53() * (r * r)) * e;
    fp = clamped ? T(0) : T(-5.0 / 6.0) * (T(1) + s) * e;
  }
}

template <typename T, int KIND>
__global__ void __launch_bounds__(TILE_THREADS)
kernel_matrix_backward_kernel(const T* __restrict__ PA, int64_t n, const T* __restrict__ PB, int64_t m, int D,
                              int64_t ldp, T variance, const T* __restrict__ G, int64_t ldg,
                              T* __restrict__ partials /* [blocks][1 + D] */) {
  __shared__ TileSmem<T> s;
  __shared__ T red[33];
  const int64_t row0 = (int64_t)blockIdx.y * TILE, col0 = (int64_t)blockIdx.x * TILE;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  T acc[4][4];
  tile_compute<T, 1>(acc, s, PA, ldp, row0, n, PB, ldp, col0, m, D);  // difference form: r2 = sum (a - b)^2
  T w[4][4];
  T gv = T(0);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t r = row0 + ty * 4 + i, c = col0 + tx * 4 + j;
      T g = T(0);
      if (r < n && c < m) g = G[r * ldg + c];
      T f, fp;
      kernel_f_and_dr2<T, KIND>(acc[i][j], f, fp);
      gv += g * f;
      w[i][j] = T(-2) * g * variance * fp;
    }
  T* out = partials + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * (1 + D);
  const T gvs = block_sum(gv, red);
  if (threadIdx.x == 0) out[0] = gvs;
  for (int d0 = 0; d0 < D; d0 += TILE_DC) {
    __syncthreads();
    tile_load(s.a, PA, ldp, row0, n, d0, D);
    tile_load(s.b, PB, ldp, col0, m, d0, D);
    __syncthreads();
    const int dc = (D - d0) < TILE_DC ? (D - d0) : TILE_DC;
    for (int d = 0; d < dc; ++d) {
      T gl = T(0);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const T df = s.a[d][ty * 4 + i] - s.b[d][tx * 4 + j];
          gl = fma(w[i][j], df * df, gl);
        }
      const T gls = block_sum(gl, red);
      if (threadIdx.x == 0) out[1 + d0 + d] = gls;  // still to be divided by l_d
    }
  }
}

template <typename T>
__global__ void kernel_matrix_backward_reduce(const T* __restrict__ partials, int64_t blocks, int D, LsParam ls,
                                              T* __restrict__ g_variance, T* __restrict__ g_ls) {
  const int k = blockIdx.x;  // 0: variance, 1 + d: lengthscale d
  __shared__ T red[33];
  T v = T(0);
  for (int64_t b = threadIdx.x; b < blocks; b += blockDim.x) v += partials[b * (1 + D) + k];
  const T sum = block_sum(v, red);
  if (threadIdx.x == 0) {
    if (k == 0) *g_variance = sum;
    else g_ls[k - 1] = sum / (T)ls.v[ls.count == 1 ? 0 : k - 1];
  }
}

// float64 squared-Euclidean assignment (cggp/optimize.py:50-51, `argmin(square_distance(iv, inputs), axis=0)`) on the
// DMMA path: the expanded distance |x|^2 + |z|^2 - 2 x.z of an 8 x 8 block comes out of mma.sync.m8n8k4.f64 (|z|^2 rides
// in the spare feature slot, |x|^2 initialises the accumulator), 12 FMA-slots per pair instead of a DFMA each plus the
// shared-memory traffic of the register-tile engine.  A warp owns 32 rows (4 A-fragment sets in registers) and sweeps
// the centres, which all 8 warps of the CTA stage through shared memory in chunks of 256.  Each lane keeps the first
// minimum of its column subset (columns ascend per lane, strict <), the 4 lanes of a row merge with ties going to the
// smaller index: the result is tf.argmin's first minimum.
namespace ncd {
constexpr int ZC = 256, RBW = 4, WARPS = 8;
__host__ __device__ constexpr int ldz_for(int KS) { return ((KS * 4) % 8 == 4) ? KS * 4 : KS * 4 + 4; }
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int KS>
__global__ void __launch_bounds__(WARPS * 32)
nearest_center_dmma_kernel(const double* __restrict__ PX, const double* __restrict__ nX, int64_t n,
                           const double* __restrict__ PZ, const double* __restrict__ nZ, int64_t m, int D, int64_t ldp,
                           int64_t* __restrict__ idx, double* __restrict__ dist) {
  constexpr int LDZ = ldz_for(KS);
  extern __shared__ __align__(16) double zt[];  // [ZC][LDZ]: -2 z | |z|^2 | 0
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lk = lane & 3;
  const int64_t row0 = ((int64_t)blockIdx.x * WARPS + warp) * (RBW * 8);
  double af[RBW][KS], xn[RBW], best[RBW];
  int bidx[RBW];
#pragma unroll
  for (int rb = 0; rb < RBW; ++rb) {
    const int64_t r = row0 + rb * 8 + lr;
    const bool valid = r < n;
    xn[rb] = valid ? nX[r] : 0.0;
    best[rb] = INFINITY;
    bidx[rb] = 0;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int k = ks * 4 + lk;
      af[rb][ks] = valid ? (k < D ? PX[r * ldp + k] : (k == D ? 1.0 : 0.0)) : 0.0;
    }
  }
  for (int64_t c0 = 0; c0 < m; c0 += ZC) {
    __syncthreads();
    for (int e = tid; e < ZC * LDZ; e += WARPS * 32) {
      const int c = e / LDZ, k = e % LDZ;
      const int64_t gc = c0 + c;
      double v = 0.0;
      if (gc < m) {
        if (k < D) v = -2.0 * PZ[gc * ldp + k];
        else if (k == D) v = nZ[gc];
      } else if (k == D) {
        v = INFINITY;  // past the last centre: never the minimum
      }
      zt[e] = v;
    }
    __syncthreads();
    const double* zw = zt + lr * LDZ + lk;
#pragma unroll 4
    for (int cb = 0; cb < ZC / 8; ++cb) {
      double bf[KS];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) bf[ks] = zw[cb * 8 * LDZ + ks * 4];
      const int col = (int)c0 + cb * 8 + 2 * lk;
#pragma unroll
      for (int rb = 0; rb < RBW; ++rb) {
        double d0 = xn[rb], d1 = xn[rb];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) dmma884(d0, d1, af[rb][ks], bf[ks]);
        if (d0 < best[rb]) { best[rb] = d0; bidx[rb] = col; }
        if (d1 < best[rb]) { best[rb] = d1; bidx[rb] = col + 1; }
      }
    }
  }
#pragma unroll
  for (int rb = 0; rb < RBW; ++rb) {
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, best[rb], o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx[rb], o);
      if (ov < best[rb] || (ov == best[rb] && oi < bidx[rb])) {
        best[rb] = ov;
        bidx[rb] = oi;
      }
    }
    const int64_t r = row0 + rb * 8 + lr;
    if (lk == 0 && r < n) {
      idx[r] = bidx[rb];
      dist[r] = best[rb];
    }
  }
}

template <int KS>
static void launch(cggp_ctx* ctx, const double* PX, const double* nX, int64_t n, const double* PZ, const double* nZ,
                   int64_t m, int D, int64_t ldp, int64_t* idx, double* dist) {
  const size_t smem = sizeof(double) * ZC * ldz_for(KS);
  const unsigned grid = (unsigned)((n + WARPS * RBW * 8 - 1) / (WARPS * RBW * 8));
  nearest_center_dmma_kernel<KS><<<grid, WARPS * 32, smem, ctx->stream>>>(PX, nX, n, PZ, nZ, m, D, ldp, idx, dist);
}
}  // namespace ncd

template <typename T>
__global__ void cluster_stats_kernel(const int64_t* __restrict__ idx, const T* __restrict__ y, int64_t n, int64_t m,
                                     T* __restrict__ counts, T* __restrict__ sums) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t j = idx[i];
  if (j < 0 || j >= m) return;
  atomicAdd(&counts[j], T(1));
  atomicAdd(&sums[j], y[i]);
}

// ---------------------------------------------------------------------------------------------------------
#define DISPATCH_KIND(T, MODE, CALL)                                          \
  switch (kind) {                                                             \
    case CGGP_SE: { constexpr int K = CGGP_SE; CALL; } break;                 \
    case CGGP_MATERN12: { constexpr int K = CGGP_MATERN12; CALL; } break;     \
    case CGGP_MATERN32: { constexpr int K = CGGP_MATERN32; CALL; } break;     \
    case CGGP_MATERN52: { constexpr int K = CGGP_MATERN52; CALL; } break;     \
    default: CGGP_FAIL(ctx, CGGP_ERR_INVALID, "unknown kernel kind %d", kind); \
  }

extern "C" int64_t cggp_prepared_ld(int D) { return ((int64_t)D + 1 + 3) / 4 * 4; }

extern "C" int cggp_prepare_points(cggp_ctx* ctx, int dtype, const void* X, int64_t n, int D, int64_t ldx,
                                   const double* ls, int ls_count, void* P, int64_t ldp, void* norms) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (D < 1 || D > 128) CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "D=%d outside [1,128]", D);
  if (ls_count != 1 && ls_count != D) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "lengthscales count %d != 1 or D", ls_count);
  if (ldp < cggp_prepared_ld(D)) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "ldp=%lld < %lld", (long long)ldp,
                                           (long long)cggp_prepared_ld(D));
  if (n == 0) return CGGP_OK;
  LsParam lp;
  lp.count = ls_count;
  for (int d = 0; d < ls_count; ++d) lp.v[d] = ls[d];
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  if (dtype == CGGP_F64)
    prepare_points_kernel<double><<<blocks, threads, 0, ctx->stream>>>((const double*)X, n, D, ldx, lp, (double*)P,
                                                                        ldp, (double*)norms);
  else
    prepare_points_kernel<float><<<blocks, threads, 0, ctx->stream>>>((const float*)X, n, D, ldx, lp, (float*)P, ldp,
                                                                       (float*)norms);
  CGGP_LAUNCH_CHECK(ctx);
  return CGGP_OK;
}

template <typename T>
static int kernel_matrix_impl(cggp_ctx* ctx, int kind, double variance, int output, int distance, const void* PA,
                              const void* nA, int64_t n, const void* PB, const void* nB, int64_t m, int D,
                              int64_t ldp, double jitter, void* out, int64_t ldo) {
  dim3 grid((unsigned)((m + TILE - 1) / TILE), (unsigned)((n + TILE - 1) / TILE));
  const bool diff = (output == CGGP_OUT_DISTANCE && distance == CGGP_DIST_EUCLIDEAN);
#define KM_CALL(MODE)                                                                                             \
  kernel_matrix_kernel<T, K, MODE><<<grid, TILE_THREADS, 0, ctx->stream>>>(                                       \
      (const T*)PA, (const T*)nA, n, (const T*)PB, (const T*)nB, m, D, ldp, (T)variance, output, distance,        \
      (T)jitter, (T*)out, ldo)
  if (diff) {
    DISPATCH_KIND(T, 1, KM_CALL(1));
  } else {
    DISPATCH_KIND(T, 0, KM_CALL(0));
  }
#undef KM_CALL
  CGGP_LAUNCH_CHECK(ctx);
  return CGGP_OK;
}

extern "C" int cggp_kernel_matrix(cggp_ctx* ctx, int dtype, int kind, double variance, int output, int distance,
                                  const void* PA, const void* nA, int64_t n, const void* PB, const void* nB, int64_t m,
                                  int D, int64_t ldp, double jitter, void* out, int64_t ldo) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (n == 0 || m == 0) return CGGP_OK;
  if (n > 65535LL * TILE) CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "n too large for one kernel_matrix call; batch it");
  if (ldo < m) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "ldo < m");
  if (dtype == CGGP_F64)
    return kernel_matrix_impl<double>(ctx, kind, variance, output, distance, PA, nA, n, PB, nB, m, D, ldp, jitter, out,
                                      ldo);
  return kernel_matrix_impl<float>(ctx, kind, variance, output, distance, PA, nA, n, PB, nB, m, D, ldp, jitter, out,
                                   ldo);
}

template <typename T>
static int nearest_center_impl(cggp_ctx* ctx, int kind, double variance, int distance, const void* PX, const void* nX,
                               int64_t n, const void* PZ, const void* nZ, int64_t m, int D, int64_t ldp, int64_t* idx,
                               void* dist) {
  if (sizeof(T) == 8 && distance == CGGP_DIST_SQEUCLIDEAN && D + 1 <= 16 && m < (1LL << 31) - 512) {
    const int ks = (D + 1 + 3) / 4;
    const double *px = (const double*)PX, *nx = (const double*)nX, *pz = (const double*)PZ, *nz = (const double*)nZ;
    switch (ks) {
      case 1: ncd::launch<1>(ctx, px, nx, n, pz, nz, m, D, ldp, idx, (double*)dist); break;
      case 2: ncd::launch<2>(ctx, px, nx, n, pz, nz, m, D, ldp, idx, (double*)dist); break;
      case 3: ncd::launch<3>(ctx, px, nx, n, pz, nz, m, D, ldp, idx, (double*)dist); break;
      default: ncd::launch<4>(ctx, px, nx, n, pz, nz, m, D, ldp, idx, (double*)dist); break;
    }
    CGGP_LAUNCH_CHECK(ctx);
    return CGGP_OK;
  }
  const unsigned grid = (unsigned)((n + TILE - 1) / TILE);
  const bool diff = distance == CGGP_DIST_EUCLIDEAN;
#define NC_CALL(MODE)                                                                                       \
  nearest_center_kernel<T, K, MODE><<<grid, TILE_THREADS, 0, ctx->stream>>>(                                \
      (const T*)PX, (const T*)nX, n, (const T*)PZ, (const T*)nZ, m, D, ldp, (T)variance, distance, idx,     \
      (T*)dist)
  if (diff) {
    DISPATCH_KIND(T, 1, NC_CALL(1));
  } else {
    DISPATCH_KIND(T, 0, NC_CALL(0));
  }
#undef NC_CALL
  CGGP_LAUNCH_CHECK(ctx);
  return CGGP_OK;
}

extern "C" int cggp_nearest_center(cggp_ctx* ctx, int dtype, int kind, double variance, int distance, const void* PX,
                                   const void* nX, int64_t n, const void* PZ, const void* nZ, int64_t m, int D,
                                   int64_t ldp, int64_t* idx, void* dist) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (n == 0) return CGGP_OK;
  if (m == 0) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "nearest_center needs at least one centre");
  if (dtype == CGGP_F64)
    return nearest_center_impl<double>(ctx, kind, variance, distance, PX, nX, n, PZ, nZ, m, D, ldp, idx, dist);
  return nearest_center_impl<float>(ctx, kind, variance, distance, PX, nX, n, PZ, nZ, m, D, ldp, idx, dist);
}

template <typename T>
static int kernel_matrix_backward_impl(cggp_ctx* ctx, int kind, double variance, const void* PA, int64_t n,
                                       const void* PB, int64_t m, int D, int64_t ldp, const double* ls, int ls_count,
                                       const void* G, int64_t ldg, void* g_variance, void* g_ls) {
  const dim3 grid((unsigned)((m + TILE - 1) / TILE), (unsigned)((n + TILE - 1) / TILE));
  const int64_t blocks = (int64_t)grid.x * grid.y;
  int rc = cggp_ws_reserve(ctx, sizeof(T) * (size_t)blocks * (1 + D));
  if (rc) return rc;
  T* partials = (T*)ctx->ws;
  LsParam lp;
  lp.count = ls_count;
  for (int d = 0; d < ls_count; ++d) lp.v[d] = ls[d];
#define KMB_CALL(MODE)                                                                                          \
  kernel_matrix_backward_kernel<T, K><<<grid, TILE_THREADS, 0, ctx->stream>>>(                                  \
      (const T*)PA, n, (const T*)PB, m, D, ldp, (T)variance, (const T*)G, ldg, partials)
  DISPATCH_KIND(T, 0, KMB_CALL(0));
#undef KMB_CALL
  CGGP_LAUNCH_CHECK(ctx);
  kernel_matrix_backward_reduce<T><<<1 + D, 256, 0, ctx->stream>>>(partials, blocks, D, lp, (T*)g_variance, (T*)g_ls);
  CGGP_LAUNCH_CHECK(ctx);
  return CGGP_OK;
}

extern "C" int cggp_kernel_matrix_backward(cggp_ctx* ctx, int dtype, int kind, double variance, const void* PA,
                                           int64_t n, const void* PB, int64_t m, int D, int64_t ldp,
                                           const double* host_lengthscales, int ls_count, const void* G, int64_t ldg,
                                           void* g_variance, void* g_ls) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (D < 1 || D > 128) CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "D=%d outside [1,128]", D);
  if (ls_count != 1 && ls_count != D) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "lengthscales count %d != 1 or D", ls_count);
  if ((int64_t)((n + TILE - 1) / TILE) > 65535) CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "too many rows for one call");
  const size_t es = dtype == CGGP_F64 ? 8 : 4;
  if (n == 0 || m == 0) {
    CGGP_CUDA(ctx, cudaMemsetAsync(g_variance, 0, es, ctx->stream));
    CGGP_CUDA(ctx, cudaMemsetAsync(g_ls, 0, es * D, ctx->stream));
    return CGGP_OK;
  }
  if (dtype == CGGP_F64)
    return kernel_matrix_backward_impl<double>(ctx, kind, variance, PA, n, PB, m, D, ldp, host_lengthscales, ls_count,
                                               G, ldg, g_variance, g_ls);
  return kernel_matrix_backward_impl<float>(ctx, kind, variance, PA, n, PB, m, D, ldp, host_lengthscales, ls_count, G,
                                            ldg, g_variance, g_ls);
}

extern "C" int cggp_cluster_stats(cggp_ctx* ctx, int dtype, const int64_t* idx, const void* y, int64_t n, int64_t m,
                                  void* counts, void* sums) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  const size_t es = dtype == CGGP_F64 ? 8 : 4;
  CGGP_CUDA(ctx, cudaMemsetAsync(counts, 0, es * m, ctx->stream));
  CGGP_CUDA(ctx, cudaMemsetAsync(sums, 0, es * m, ctx->stream));
  if (n == 0) return CGGP_OK;
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  if (dtype == CGGP_F64)
    cluster_stats_kernel<double><<<blocks, threads, 0, ctx->stream>>>(idx, (const double*)y, n, m, (double*)counts,
                                                                       (double*)sums);
  else
    cluster_stats_kernel<float><<<blocks, threads, 0, ctx->stream>>>(idx, (const float*)y, n, m, (float*)counts,
                                                                      (float*)sums);
  CGGP_LAUNCH_CHECK(ctx);
  return CGGP_OK;
}
