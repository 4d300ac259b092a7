// Model-level entry points of the C ABI: the prediction and objective chains of cggp/models.py as single calls.
//   cggp_predict_f   <-> CGGP.predict_f, cggp/models.py:333-352 (Kmn, the B-RHS solve, fvar and fmu reductions)
//   cggp_elbo_terms  <-> the data term of LpSVGP.elbo, cggp/models.py:131-133 (GPflow Gaussian variational_expectations)
// Everything stays on the device; the reductions are deterministic (fixed order, no atomics).
#include <cmath>
#include <cstring>

#include "common.cuh"

// mean[b] = sum_m Knm[b, m] a[m]          (models.py:351, fmu = Kmn^T a)
// var[b]  = knn - sum_m Knm[b, m] S[b, m] (models.py:343-345, fvar = Knn - sum_m Kmn * S)
// One warp per test point: both rows are streamed once (HBM-bound: 2 * 8 * B * M bytes).
template <typename T>
__global__ void __launch_bounds__(256)
predict_reduce_kernel(const T* __restrict__ Knm, int64_t ldk, const T* __restrict__ S, int64_t lds,
                      const T* __restrict__ a, int64_t m, int64_t nb, T knn, T* __restrict__ mean,
                      T* __restrict__ var) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= nb) return;
  const T* kr = Knm + b * ldk;
  const T* sr = S + b * lds;
  T dm = T(0), dv = T(0);
  for (int64_t k = lane; k < m; k += 32) {
    const T kv = kr[k];
    dm = fma(kv, a[k], dm);
    dv = fma(kv, sr[k], dv);
  }
  dm = warp_sum(dm);
  dv = warp_sum(dv);
  if (lane == 0) {
    mean[b] = dm;
    var[b] = knn - dv;
  }
}

// sum_i [ -1/2 log(2 pi) - 1/2 log(s2) - 1/2 ((y_i - mu_i)^2 + v_i) / s2 ]   (GPflow Gaussian.variational_expectations)
// stage 1: one partial per CTA (fixed grid), stage 2: one CTA adds the partials in order.
template <typename T>
__global__ void __launch_bounds__(256)
gauss_expect_partial_kernel(const T* __restrict__ y, const T* __restrict__ mu, const T* __restrict__ v, int64_t n,
                            double c0, double inv_s2, double* __restrict__ partial) {
  __shared__ double red[33];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double e = (double)y[i] - (double)mu[i];
    acc += c0 - 0.5 * (e * e + (double)v[i]) * inv_s2;
  }
  const double s = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
template <typename T>
__global__ void gauss_expect_final_kernel(const double* __restrict__ partial, int np, T* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < np; ++i) s += partial[i];
    out[0] = (T)s;
  }
}

extern "C" int cggp_elbo_terms(cggp_ctx* ctx, int dtype, const void* y, const void* mean, const void* var, int64_t n,
                               double noise_variance, void* out) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (!out || n < 0 || !(noise_variance > 0.0)) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "elbo_terms: bad arguments");
  const int np = 256;
  int rc = cggp_ws_reserve(ctx, sizeof(double) * np);
  if (rc) return rc;
  double* partial = (double*)ctx->ws;
  const double c0 = -0.5 * log(2.0 * 3.14159265358979323846) - 0.5 * log(noise_variance);
  const double inv = 1.0 / noise_variance;
  if (dtype == CGGP_F64) {
    gauss_expect_partial_kernel<double><<<np, 256, 0, ctx->stream>>>((const double*)y, (const double*)mean,
                                                                    (const double*)var, n, c0, inv, partial);
    CGGP_LAUNCH_CHECK(ctx);
    gauss_expect_final_kernel<double><<<1, 32, 0, ctx->stream>>>(partial, np, (double*)out);
  } else {
    gauss_expect_partial_kernel<float><<<np, 256, 0, ctx->stream>>>((const float*)y, (const float*)mean,
                                                                   (const float*)var, n, c0, inv, partial);
    CGGP_LAUNCH_CHECK(ctx);
    gauss_expect_final_kernel<float><<<1, 32, 0, ctx->stream>>>(partial, np, (float*)out);
  }
  CGGP_LAUNCH_CHECK(ctx);
  return CGGP_OK;
}

extern "C" int cggp_predict_f(cggp_ctx* ctx, int dtype, int kind, double variance, const void* PZ, const void* nZ,
                              int64_t m, const void* Pnew, const void* nNew, int64_t nb, int D, int64_t ldp,
                              const void* A, int64_t lda, const void* a, double error_threshold, int max_iterations,
                              int max_steps_cycle, void* Knm_work, void* S_work, void* mean, void* var,
                              int32_t* host_steps) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (host_steps) *host_steps = 0;
  if (nb <= 0) return CGGP_OK;
  if (!A || !a || !Knm_work || !S_work || !mean || !var || m <= 0)
    CGGP_FAIL(ctx, CGGP_ERR_INVALID, "predict_f: null buffer or empty system");
  if (nb > 0x7fffffff) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "predict_f: batch too large, predict in batches");
  // Knm = K(Xnew, Z) [nb, m]: the rows are the right-hand sides of the solve (the reference's Kmn, transposed)
  int rc = cggp_kernel_matrix(ctx, dtype, kind, variance, CGGP_OUT_KERNEL, 0, Pnew, nNew, nb, PZ, nZ, m, D, ldp, 0.0,
                              Knm_work, m);
  if (rc) return rc;
  cggp_operator op;
  memset(&op, 0, sizeof(op));
  op.struct_size = (uint32_t)sizeof(op);
  op.type = CGGP_OP_DENSE;
  op.dtype = dtype;
  op.n = m;
  op.dev_A = A;
  op.lda = lda;
  rc = cggp_cg_solve(ctx, &op, Knm_work, nullptr, (int)nb, error_threshold, max_iterations, max_steps_cycle, nullptr, 16,
                     S_work, host_steps, nullptr, nullptr, 0);  // models.py:340
  if (rc) return rc;
  const unsigned grid = (unsigned)((nb + 7) / 8);
  if (dtype == CGGP_F64)
    predict_reduce_kernel<double><<<grid, 256, 0, ctx->stream>>>((const double*)Knm_work, m, (const double*)S_work, m,
                                                                 (const double*)a, m, nb, variance, (double*)mean,
                                                                 (double*)var);
  else
    predict_reduce_kernel<float><<<grid, 256, 0, ctx->stream>>>((const float*)Knm_work, m, (const float*)S_work, m,
                                                                (const float*)a, m, nb, (float)variance, (float*)mean,
                                                                (float*)var);
  CGGP_LAUNCH_CHECK(ctx);
  return CGGP_OK;
}
