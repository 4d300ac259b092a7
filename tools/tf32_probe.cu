// Probe for the float32 path: one 128 x 128 x K TF32 tile through tcgen05.mma (A, B K-major in shared memory with the
// canonical no-swizzle "interleave" layout, accumulator in TMEM, read back with tcgen05.ld) against a CPU reference.
// Validates the shared-memory / instruction descriptors that csrc/matvec_tf32.cu builds on.
// Build: nvcc --cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tf32_probe tools/tf32_probe.cu
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int M = 128, N = 128;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// element (r, k) of a K-major operand tile [rows x KP] in the canonical no-swizzle layout, in floats:
// [r / 8][k / 4][r % 8][k % 4]  -> core matrix = 8 rows x 16 bytes = 128 contiguous bytes;
// LBO (next 16-byte chunk along K) = 128 B, SBO (next 8-row group) = (KP / 4) * 128 B
__host__ __device__ inline int canon_off(int r, int k, int KP) { return ((r >> 3) * (KP >> 2) + (k >> 2)) * 32 + (r & 7) * 4 + (k & 3); }

__device__ __forceinline__ uint64_t make_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) |
         (1ull << 46);
}

template <int KP, bool A_TMEM>
__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* B, float* D) {
  extern __shared__ __align__(128) float smem[];
  float* sA = smem;
  float* sB = smem + M * KP;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < M * KP; e += 128) { const int r = e / KP, k = e % KP; sA[canon_off(r, k, KP)] = A[e]; }
  for (int e = tid; e < N * KP; e += 128) { const int r = e / KP, k = e % KP; sB[canon_off(r, k, KP)] = B[e]; }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // generic-proxy writes of the operands -> visible to the async proxy (tensor core reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tb = tmem_base;
  if (A_TMEM) {
    // A operand in tensor memory: row m -> lane m, feature k -> column 128 + k (one 32-bit column per TF32 element);
    // every thread stores its own row with tcgen05.st (thread = lane), 8 columns at a time
    const int row = warp * 32 + lane;
    for (int k0 = 0; k0 < KP; k0 += 8) {
      uint32_t v[8];
      for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(A[row * KP + k0 + i]);
      const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + 128 + k0;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
                   "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const unsigned lbo = 128, sbo = (KP / 4) * 128;
    for (int j = 0; j < KP / 8; ++j) {
      const uint64_t da = make_desc(smem_u32(sA) + j * 256, lbo, sbo);
      const uint64_t db = make_desc(smem_u32(sB) + j * 256, lbo, sbo);
      const uint32_t acc = j > 0;
      if (A_TMEM) {
        const uint32_t ta = tb + 128 + j * 8;
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tb), "r"(ta), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
      } else {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tb), "l"(da), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // wait for the MMAs
  asm volatile(
      "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra DN;\nbra W;\nDN:\n}\n" ::"r"(smem_u32(&mbar))
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;");
  // epilogue: warp w reads TMEM lanes 32 w .. 32 w + 31 (= rows), 32 columns at a time
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    const uint32_t taddr = tb + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int row = warp * 32 + lane;
    for (int c = 0; c < 32; ++c) D[row * N + c0 + c] = __uint_as_float(v[c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tb));
}

static float tf32_trunc(float x) { unsigned u; memcpy(&u, &x, 4); u &= 0xffffe000u; float y; memcpy(&y, &u, 4); return y; }

template <int KP, bool A_TMEM>
int run() {
  std::vector<float> A(M * KP), B(N * KP), D(M * N), R(M * N), Rt(M * N);
  srand(1);
  for (auto& x : A) x = (rand() / (float)RAND_MAX - 0.5f);
  for (auto& x : B) x = (rand() / (float)RAND_MAX - 0.5f);
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < N; ++j) {
      double s = 0, st = 0;
      for (int k = 0; k < KP; ++k) { s += (double)A[i * KP + k] * B[j * KP + k]; st += (double)tf32_trunc(A[i * KP + k]) * tf32_trunc(B[j * KP + k]); }
      R[i * N + j] = (float)s; Rt[i * N + j] = (float)st;
    }
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, D.size() * 4));
  const size_t smem = (size_t)(M + N) * KP * 4;
  CK(cudaFuncSetAttribute(probe_kernel<KP, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<KP, A_TMEM><<<1, 128, smem>>>(dA, dB, dD);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double e_full = 0, e_trunc = 0, scale = 0;
  for (int i = 0; i < M * N; ++i) { e_full = fmax(e_full, fabs(D[i] - R[i])); e_trunc = fmax(e_trunc, fabs(D[i] - Rt[i])); scale = fmax(scale, fabs(R[i])); }
  printf("%s K=%3d: max|D - fp32 ref| = %.3e   max|D - tf32-truncated ref| = %.3e   (scale %.3f)  D[0..3] = %f %f %f %f  ref %f %f %f %f\n", A_TMEM ? "A in TMEM:" : "A in smem:", KP, e_full,
         e_trunc, scale, D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
  return (e_full < 5e-3 * scale) ? 0 : 1;
}

int main() {
  int bad = run<8, false>();
  bad += run<32, false>();
  bad += run<96, false>();
  bad += run<8, true>();
  bad += run<32, true>();
  bad += run<96, true>();
  printf(bad ? "PROBE FAIL\n" : "PROBE OK\n");
  return bad;
}
