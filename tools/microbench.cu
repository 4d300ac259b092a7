// FP64 pipe micro-benchmarks for B200 (sm_100a): what bounds the fused Kuf*Kfu matvec kernel.
// Build: nvcc --cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double* c, const double* a, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

__global__ void k_dfma(double* out, int iters, double s) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], s, 1e-9);
  }
  double r = 0; for (int i = 0; i < 8; ++i) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_dmma884(double* out, int iters, double s) {
  double c[16]; for (int i = 0; i < 16; ++i) c[i] = 0;
  double a = threadIdx.x * 1e-3, b = s;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
  }
  double r = 0; for (int i = 0; i < 16; ++i) r += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_dmma1684(double* out, int iters, double s) {
  double c[32]; for (int i = 0; i < 32; ++i) c[i] = 0;
  double a[2] = {threadIdx.x * 1e-3, 0.5}, b = s;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma1684(c + 4 * i, a, b);
  }
  double r = 0; for (int i = 0; i < 32; ++i) r += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_dmma1688(double* out, int iters, double s) {
  double c[32]; for (int i = 0; i < 32; ++i) c[i] = 0;
  double a[4] = {threadIdx.x * 1e-3, 0.5, 0.25, 0.125}, b[2] = {s, s * 0.5};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma1688(c + 4 * i, a, b);
  }
  double r = 0; for (int i = 0; i < 32; ++i) r += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_dmma16816(double* out, int iters, double s) {
  double c[32]; for (int i = 0; i < 32; ++i) c[i] = 0;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int i = 0; i < 4; ++i) b[i] = s * (i + 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma16816(c + 4 * i, a, b);
  }
  double r = 0; for (int i = 0; i < 32; ++i) r += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// DMMA + DFMA interleaved: do they share the pipe?
__global__ void k_mix(double* out, int iters, double s) {
  double c[16]; for (int i = 0; i < 16; ++i) c[i] = 0;
  double f[8]; for (int i = 0; i < 8; ++i) f[i] = threadIdx.x * 1e-3 + i;
  double a = threadIdx.x * 1e-3, b = s;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { dmma884(c[2 * i], c[2 * i + 1], a, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fma(f[j], s, 1e-9); }
  }
  double r = 0; for (int i = 0; i < 16; ++i) r += c[i]; for (int i = 0; i < 8; ++i) r += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_exp(double* out, int iters, double s) {
  double a[4]; for (int i = 0; i < 4; ++i) a[i] = -(threadIdx.x * 1e-2 + i);
  double acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc += exp(a[i]); a[i] += s; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_sqrt(double* out, int iters, double s) {
  double a[4]; for (int i = 0; i < 4; ++i) a[i] = (threadIdx.x * 1e-2 + i + 1);
  double acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc += sqrt(a[i]); a[i] += s; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__device__ __forceinline__ double my_fsqrt(double u) {
  double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(u));
  double h = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));
  double g = u * y;
  double e = fma(-g, g, u); g = fma(e, h, g);
  e = fma(-g, g, u); g = fma(e, h, g);
  return g;
}
__global__ void k_fsqrt(double* out, int iters, double s) {
  double a[4]; for (int i = 0; i < 4; ++i) a[i] = (threadIdx.x * 1e-2 + i + 1);
  double acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc += my_fsqrt(a[i]); a[i] += s; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_rsq(double* out, int iters, double s) {
  double a[4]; for (int i = 0; i < 4; ++i) a[i] = (threadIdx.x * 1e-2 + i + 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a[i])); a[i] = y; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a[0] + a[1] + a[2] + a[3];
}
// fast exp: 32-entry table in lanes via shuffle, degree-5 polynomial
__device__ __forceinline__ double my_fexp(double x, double tab) {
  const double L2E32 = 46.16624130844682903551758979206054839765;
  const double MAGIC = 6755399441055744.0;
  const double LN2_32 = 0.02166084939249829091928849858592451515688;
  double t = fma(x, L2E32, MAGIC);
  int ki = __double2loint(t);
  double kf = t - MAGIC;
  double d = fma(kf, -LN2_32, x);
  double q = fma(d, 8.33333333333333e-3, 4.16666666666667e-2);
  q = fma(q, d, 1.66666666666667e-1);
  q = fma(q, d, 0.5);
  q = fma(q, d, 1.0);
  q = fma(q, d, 1.0);
  int j = ki & 31;
  int hi = __shfl_sync(0xffffffffu, __double2hiint(tab), j);
  int lo = __shfl_sync(0xffffffffu, __double2loint(tab), j);
  hi += (ki >> 5) << 20;
  return __hiloint2double(hi, lo) * q;
}
__global__ void k_fexp(double* out, int iters, double s) {
  double tab = exp2((threadIdx.x & 31) / 32.0);
  double a[4]; for (int i = 0; i < 4; ++i) a[i] = -(threadIdx.x * 1e-2 + i);
  double acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc += my_fexp(a[i], tab); a[i] += s; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// DFMA + shuffle co-issue
__global__ void k_dfma_shfl(double* out, int iters, double s) {
  double a[8]; int v = threadIdx.x;
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fma(a[i], s, 1e-9);
    v = __shfl_sync(0xffffffffu, v, (v + it) & 31);
    v = __shfl_sync(0xffffffffu, v, (v + 3) & 31);
  }
  double r = v; for (int i = 0; i < 8; ++i) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <typename F> float run(F f, int reps = 3) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
  return best;
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sms=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  int nb = p.multiProcessorCount * 8, nt = 256, iters = 20000;
  double* out; CK(cudaMalloc(&out, sizeof(double) * nb * nt));
  double total_threads = (double)nb * nt;
  float ms;
  ms = run([&] { k_dfma<<<nb, nt>>>(out, iters, 0.999); });
  printf("DFMA            : %.2f TFLOP/s (%.3f ms)\n", total_threads * iters * 8 * 2 / ms / 1e9, ms);
  ms = run([&] { k_dmma884<<<nb, nt>>>(out, iters / 4, 0.999); });
  printf("DMMA m8n8k4     : %.2f TFLOP/s (%.3f ms)\n", (total_threads / 32) * (iters / 4) * 8 * (8 * 8 * 4 * 2.0) / ms / 1e9, ms);
  ms = run([&] { k_dmma1684<<<nb, nt>>>(out, iters / 4, 0.999); });
  printf("DMMA m16n8k4    : %.2f TFLOP/s (%.3f ms)\n", (total_threads / 32) * (iters / 4) * 8 * (16 * 8 * 4 * 2.0) / ms / 1e9, ms);
  ms = run([&] { k_dmma1688<<<nb, nt>>>(out, iters / 4, 0.999); });
  printf("DMMA m16n8k8    : %.2f TFLOP/s (%.3f ms)\n", (total_threads / 32) * (iters / 4) * 8 * (16 * 8 * 8 * 2.0) / ms / 1e9, ms);
  ms = run([&] { k_dmma16816<<<nb, nt>>>(out, iters / 8, 0.999); });
  printf("DMMA m16n8k16   : %.2f TFLOP/s (%.3f ms)\n", (total_threads / 32) * (iters / 8) * 8 * (16 * 8 * 16 * 2.0) / ms / 1e9, ms);
  ms = run([&] { k_mix<<<nb, nt>>>(out, iters / 8, 0.999); });
  { double fl = (total_threads / 32) * (iters / 8) * 8 * (8 * 8 * 4 * 2.0) + total_threads * (iters / 8) * 64 * 2.0;
    printf("DMMA+DFMA mix   : %.2f TFLOP/s total (%.3f ms) [dmma part %.1f%%]\n", fl / ms / 1e9, ms, 100.0 * (total_threads / 32) * (iters / 8) * 8 * 512.0 / fl); }
  ms = run([&] { k_exp<<<nb, nt>>>(out, iters / 8, 1e-6); });
  printf("libm exp        : %.2f Gexp/s (%.3f ms)\n", total_threads * (iters / 8) * 4 / ms / 1e6, ms);
  ms = run([&] { k_fexp<<<nb, nt>>>(out, iters / 8, 1e-6); });
  printf("fast exp (shfl) : %.2f Gexp/s (%.3f ms)\n", total_threads * (iters / 8) * 4 / ms / 1e6, ms);
  ms = run([&] { k_sqrt<<<nb, nt>>>(out, iters / 8, 1e-6); });
  printf("libm sqrt       : %.2f Gsqrt/s (%.3f ms)\n", total_threads * (iters / 8) * 4 / ms / 1e6, ms);
  ms = run([&] { k_fsqrt<<<nb, nt>>>(out, iters / 8, 1e-6); });
  printf("fast sqrt       : %.2f Gsqrt/s (%.3f ms)\n", total_threads * (iters / 8) * 4 / ms / 1e6, ms);
  ms = run([&] { k_rsq<<<nb, nt>>>(out, iters / 8, 1e-6); });
  printf("MUFU.RSQ64H     : %.2f Gop/s (%.3f ms)\n", total_threads * (iters / 8) * 4 / ms / 1e6, ms);
  ms = run([&] { k_dfma_shfl<<<nb, nt>>>(out, iters, 0.999); });
  printf("DFMA(8)+2 SHFL  : %.2f TFLOP/s (%.3f ms)\n", total_threads * iters * 8 * 2 / ms / 1e9, ms);
  // accuracy of fast exp / sqrt vs libm on host-visible sample is checked in tests; here only rates.
  return 0;
}
