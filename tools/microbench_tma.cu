// Micro-benchmark: how fast can every SM pull L2-resident data into shared memory with 1-D TMA bulk copies
// (cp.async.bulk + mbarrier), the feed of the tcgen05 gram kernel?  One producer thread per CTA keeps `stages` copies
// of `chunk` bytes in flight from a `footprint`-byte buffer that all CTAs read (optionally each from its own offset).
// Build: nvcc --cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench_tma tools/microbench_tma.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(128, 1) feed_kernel(const char* src, size_t footprint, int chunk, int stages, int iters,
                                                      int rotate, int producers) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar[16];
  if (threadIdx.x == 0) {
    for (int s = 0; s < 16; ++s) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) != 0 || w >= producers) return;
  // producer w handles stages s with s % producers == w
  const size_t nchunks = footprint / chunk;
  size_t pos = rotate ? ((size_t)blockIdx.x * 37) % nchunks : 0;
  for (int it = 0; it < iters; ++it) {
    for (int s = w; s < stages; s += producers) {
      if (it > 0) mbar_wait(&bar[s], (unsigned)((it - 1) & 1));
      mbar_expect_tx(&bar[s], (unsigned)chunk);
      tma_bulk_g2s(smem + (size_t)s * chunk, src + ((pos + s) % nchunks) * (size_t)chunk, (unsigned)chunk, &bar[s]);
    }
    pos = (pos + stages) % nchunks;
  }
  for (int s = w; s < stages; s += producers) mbar_wait(&bar[s], (unsigned)((iters - 1) & 1));
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const double clk = prop.clockRate * 1e3;
  const size_t maxfoot = 512ull << 20;
  char* src;
  CK(cudaMalloc(&src, maxfoot));
  CK(cudaMemset(src, 1, maxfoot));
  CK(cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  printf("%s sms=%d\n", prop.name, sms);
  struct Cfg { size_t foot; int chunk, stages, rotate, producers, ctas; };
  const Cfg cfgs[] = {
      {6u << 20, 16384, 6, 0, 1, 0}, {6u << 20, 16384, 6, 1, 1, 0}, {6u << 20, 16384, 12, 1, 1, 0},
      {6u << 20, 8192, 12, 1, 1, 0},  {6u << 20, 32768, 6, 1, 1, 0}, {6u << 20, 16384, 12, 1, 2, 0},
      {6u << 20, 16384, 12, 1, 4, 0}, {64u << 20, 16384, 12, 1, 1, 0}, {512ull << 20, 16384, 12, 1, 1, 0},
      {6u << 20, 16384, 12, 1, 1, 74}, {6u << 20, 16384, 12, 1, 1, 37}, {6u << 20, 16384, 12, 1, 1, 8},
  };
  for (const Cfg& c : cfgs) {
    const int ctas = c.ctas ? c.ctas : sms;
    const int iters = 400;
    const size_t smem = (size_t)c.chunk * c.stages;
    feed_kernel<<<ctas, 128, smem>>>(src, c.foot, c.chunk, c.stages, 20, c.rotate, c.producers);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    feed_kernel<<<ctas, 128, smem>>>(src, c.foot, c.chunk, c.stages, iters, c.rotate, c.producers);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double bytes = (double)ctas * iters * c.stages * c.chunk;
    printf("footprint %4zu MB chunk %5d stages %2d rotate %d producers %d ctas %3d: %7.1f GB/s total, %5.1f B/clk/SM\n",
           c.foot >> 20, c.chunk, c.stages, c.rotate, c.producers, ctas, bytes / ms / 1e6, bytes / (ms * 1e-3) / clk / ctas);
  }
  return 0;
}
