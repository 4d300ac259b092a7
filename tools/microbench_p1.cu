// Micro-benchmark of phase 1 of the pipelined Kuf*Kfu kernel (DMMA distance tile + Matern-5/2 epilogue) in isolation:
// no shared-memory parking, no exchange, no phase 2.  Which part of the FP64-pipe budget is lost where?
// Build: nvcc --cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench_p1 tools/microbench_p1.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double fsqrt5(double u) {  // 1 DMUL + 4 DFMA
  double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(u));
  double h = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));
  double g = u * y;
  double e = fma(-g, g, u); g = fma(e, h, g);
  e = fma(-g, g, u); g = fma(e, h, g);
  return g;
}
__device__ __forceinline__ double fsqrt3(double u) {  // 1 DMUL + 2 DFMA (2^-47)
  double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(u));
  double h = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));
  double g = u * y;
  double e = fma(-g, g, u); g = fma(e, h, g);
  return g;
}
// EXPMODE 0: table by SHFL, 1: table by LDS, 2: no table (timing only)
template <int EXPMODE>
__device__ __forceinline__ double fexpneg(double a, int thi, int tlo, const double* stab) {
  const double L2E32 = 46.16624130844682903551758979206054839765, MAGIC = 6755399441055744.0;
  const double LN2_32 = 0.02166084939249829091928849858592451515688;
  double t = fma(a, -L2E32, MAGIC);
  int n = __double2loint(t);
  double nf = t - MAGIC;
  double d = fma(nf, -LN2_32, -a);
  double q = fma(d, 8.33337406147829918e-03, 4.16668703096581480e-02);
  q = fma(q, d, 1.66666666664448626e-01);
  q = fma(q, d, 4.99999999994028277e-01);
  q = fma(q, d, 1.0);
  q = fma(q, d, 1.0);
  if (EXPMODE == 2) return q;
  int hi, lo;
  if (EXPMODE == 0) { hi = __shfl_sync(0xffffffffu, thi, n); lo = __shfl_sync(0xffffffffu, tlo, n); }
  else if (EXPMODE == 3) { hi = __shfl_sync(0xffffffffu, thi, n); lo = 0; }
  else { double tv = stab[n & 31]; hi = __double2hiint(tv); lo = __double2loint(tv); }
  hi += n << 15;  // biased table
  return __hiloint2double(hi, lo) * q;
}

template <bool DMMA, int SQRT, bool EXP, int EXPMODE, int RB, bool TPRED = false>
__global__ void __launch_bounds__(1024) p1_kernel(double* out, int iters, double s) {
  __shared__ double stab[32];
  if (threadIdx.x < 32) stab[threadIdx.x] = exp2(threadIdx.x / 32.0);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const double tv = exp2(lane / 32.0);
  const int thi = __double2hiint(tv), tlo = __double2loint(tv);
  double bf[2][3], vv[2][2], w[2][2] = {{0, 0}, {0, 0}};
  for (int cb = 0; cb < 2; ++cb) {
    for (int ks = 0; ks < 3; ++ks) bf[cb][ks] = -1e-3 * (lane + cb + ks + 1) * s;
    vv[cb][0] = 1.0 + 1e-3 * lane; vv[cb][1] = 1.0 - 1e-3 * lane;
  }
  double xbase = 1e-2 * (lane + 1), tsum = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      double af[3];
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) af[ks] = xbase + 1e-3 * (ks + rb) + 1e-7 * it;
      const double xa = 5.0 * (2.0 + xbase + 1e-6 * it + rb);
      double c[2][2];
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        c[cb][0] = xa; c[cb][1] = xa + 0.5;
        if (DMMA) {
#pragma unroll
          for (int ks = 0; ks < 3; ++ks) dmma884(c[cb][0], c[cb][1], af[ks], bf[cb][ks]);
        } else {
          c[cb][0] += af[0] * bf[cb][0]; c[cb][1] += af[1] * bf[cb][1];
        }
      }
      double tp = 0.0;
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        double k[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int h = min(max(__double2hiint(c[cb][q]), 0x389a95a5), 0x411e9840);
          const double qc = __hiloint2double(h, __double2loint(c[cb][q]));
          const double a = SQRT == 5 ? fsqrt5(qc) : (SQRT == 3 ? fsqrt3(qc) : qc * 0.1);
          const double e = EXP ? fexpneg<EXPMODE>(a, thi, tlo, stab) : a;
          k[q] = fma(qc, 1.0 / 3.0, 1.0 + a) * e;
        }
        tp = fma(k[0], vv[cb][0], fma(k[1], vv[cb][1], tp));
        w[cb][0] = fma(k[0], xa, w[cb][0]);  // stands in for phase 2 (1 DFMA per entry)
        w[cb][1] = fma(k[1], xa, w[cb][1]);
      }
      if (TPRED) {
        tp += __shfl_xor_sync(0xffffffffu, tp, 1);
        tp += __shfl_xor_sync(0xffffffffu, tp, 2);
      }
      tsum += tp;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = tsum + w[0][0] + w[0][1] + w[1][0] + w[1][1];
}

template <typename F> float run(F f, int reps = 3) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
  return best;
}

template <bool DMMA, int SQRT, bool EXP, int EXPMODE, bool TPRED = false>
void bench(const char* name, double* out, int sms, double fp64_per_entry) {
  const int iters = 2000;
  for (int threads : {512, 1024}) {
    float ms = run([&] { p1_kernel<DMMA, SQRT, EXP, EXPMODE, 6, TPRED><<<sms, threads>>>(out, iters, 0.999); });
    const double entries = (double)sms * threads * iters * 6 * 4;
    const double slots = fp64_per_entry + (DMMA ? 12.0 : 2.0);
    const double peak = sms * 64.0 * 1.965e9;  // DFMA-equivalent slots / s at the max clock
    printf("%-34s warps/SM %2d: %7.1f Gentry/s  -> %.1f%% of the FP64 pipe (%.0f slots/entry)\n", name, threads / 32,
           entries / ms / 1e6, 100.0 * entries * slots / (ms * 1e-3) / peak, slots);
  }
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 1024));
  // FP64 (non-DMMA) instructions per entry: sqrt5 5 / sqrt3 3, poly 2 (DADD + DFMA), exp 9 (8 without table multiply),
  // K multiply 1, contractions 2
  bench<true, 5, true, 0>("full (DMMA, sqrt5, exp SHFL)", out, sms, 19);
  bench<true, 5, true, 0, true>("full + quad reduction of t", out, sms, 19.5);
  bench<true, 5, true, 3>("full, ONE SHFL (timing only)", out, sms, 19);
  bench<true, 5, true, 1>("full (DMMA, sqrt5, exp LDS)", out, sms, 19);
  bench<true, 5, true, 2>("full (DMMA, sqrt5, exp no table)", out, sms, 18);
  bench<true, 3, true, 0>("DMMA, sqrt3, exp SHFL", out, sms, 17);
  bench<false, 5, true, 0>("no DMMA, sqrt5, exp SHFL", out, sms, 19);
  bench<true, 0, true, 0>("DMMA, no sqrt, exp SHFL", out, sms, 15);
  bench<true, 5, false, 0>("DMMA, sqrt5, no exp", out, sms, 10);
  bench<true, 0, false, 0>("DMMA only (+poly)", out, sms, 6);
  bench<false, 0, false, 0>("neither (poly + contractions)", out, sms, 6);
  return 0;
}
