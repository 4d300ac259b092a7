// float32 matrix-free product  W = V @ (Kuf Kfu)  on the 5th-generation tensor cores (tcgen05, TF32 inputs, FP32
// accumulators in TMEM), the float32 counterpart of matvec_pipe.cu (BASELINE configs[4]: N = 2M, D = 90, M = 8192).
//
// At D = 90 the distance contraction dominates (2 N M D flop per sweep), so it runs as a TF32 GEMM:
//   * the scaled points are converted ONCE (cggp_tf32_prepare) into the order the tensor cores read: 128-row tiles x
//     32-feature chunks, each chunk a contiguous 16 KB block in the UMMA "canonical K-major, no swizzle" order
//     [row / 8][k / 4][row % 8][k % 4] (LBO = 128 B, SBO = 1 KB), split  x = x_big + x_small  (both round-to-nearest
//     TF32).  One TMA bulk copy (cp.async.bulk) per chunk drops it into shared memory exactly as the tcgen05
//     shared-memory descriptors describe it;
//   * 3xTF32:  x.z ~ xb.zb + xs.zb + xb.zs  (three accumulating MMA groups; the dropped xs.zs term is 2^-22 relative),
//     which keeps the expanded squared distance at float32 accuracy - a single TF32 pass (NSPLIT = 1) loses ~3 digits
//     to the cancellation in |x|^2 + |z|^2 - 2 x.z;
//   * generic "gram contraction"  out[b, p] = sum_q k(P_p, Q_q) U[b, q]:  a CTA owns 128 P rows (= the 128 TMEM lanes;
//     both parts of the P tile live in TENSOR MEMORY as the A operand of every MMA, next to the accumulators) and
//     loops over 128-column Q tiles that stream through a ring of K-chunk stages in shared memory.  Warp 8 is the TMA producer, warp 9 issues the tcgen05.mma chain (one elected lane each) and commits
//     onto mbarriers, warp 10 stages the per-column scalars (|q|^2, U); warps 0-7 are the epilogue: tcgen05.ld of a finished accumulator (thread = row, registers = 64
//     columns), r2 = |p|^2 + |q|^2 - 2 p.q, kernel value with MUFU ex2 / sqrt, dot with U.  Two TMEM accumulators
//     (2 x 128 columns) let the MMAs of tile j+1 run under the epilogue of tile j.
//   The product is two such sweeps:  T = gram(X, Z, V)  then  W = gram(Z, X, T)  (roles swapped, X split over
//   grid.y with a fixed-order reduction), i.e. every Gram entry is evaluated twice - parking a 128 x 128 FP32 tile per
//   block would be the next step, as matvec_pipe.cu does for float64.
#include <cuda_fp16.h>

#include <cstdlib>
#include <cuda_runtime.h>

#include "common.cuh"
#include "kmath.cuh"

namespace tf32 {
constexpr int BM = 128, BN = 128;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE, version 1 (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand in tensor memory (row m = lane m, feature k = column k of the given TMEM address)
__device__ __forceinline__ void umma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 (FP16 inputs, FP32 accumulate) with the A operand in tensor memory: two halfs per 32-bit TMEM column
__device__ __forceinline__ void umma_f16_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void umma_commit(void* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

// offset (in floats) of element (r, k): [r / 128][k / 32] chunks of 4096 floats, inside a chunk the UMMA canonical
// K-major order [(r % 128) / 8][(k % 32) / 4][r % 8][k % 4]
__host__ __device__ inline int64_t canon_off(int64_t r, int k, int KP) {
  const int64_t chunk = (r >> 7) * (KP >> 5) + (k >> 5);
  return chunk * 4096 + ((((r & 127) >> 3) * 8 + ((k & 31) >> 2)) * 32) + (r & 7) * 4 + (k & 3);
}
// same for the FP16 arrays (offset in halfs): core matrices are 8 rows x 8 halfs (16 bytes per row), a 128-row x
// 32-feature chunk is 8 KB: [(r % 128) / 8][(k % 32) / 8][r % 8][k % 8] (LBO = 128 B, SBO = 512 B)
__host__ __device__ inline int64_t canon_off_h(int64_t r, int k, int KP) {
  const int64_t chunk = (r >> 7) * (KP >> 5) + (k >> 5);
  return chunk * 4096 + ((((r & 127) >> 3) * 4 + ((k & 31) >> 3)) * 64) + (r & 7) * 8 + (k & 7);
}
constexpr int CHUNK_FLOATS = 4096;                      // 128 rows x 32 features
constexpr unsigned CHUNK_BYTES = CHUNK_FLOATS * 4;      // 16 KB
constexpr unsigned CHUNK_BYTES_H = 4096 * 2;            // the same chunk in FP16: 8 KB
constexpr int MAX_STAGES = 8;
constexpr int F16X3 = 16;                               // `nsplit` code of the 3xFP16 mode

// ---------------------------------------------------------------------------------------------------------
// one-time conversion of prepared points into the canonical TF32 big / small arrays (rows padded to 128)
// ---------------------------------------------------------------------------------------------------------
__global__ void prepare_kernel(const float* __restrict__ P, const float* __restrict__ norms, int64_t n, int D,
                               int64_t ldp, int KP, int64_t n_pad, float* __restrict__ big, float* __restrict__ small,
                               float* __restrict__ norms_pad) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_pad) norms_pad[e] = e < n ? norms[e] : 0.f;
  if (e >= n_pad * KP) return;
  const int64_t r = e / KP;
  const int k = (int)(e % KP);
  float x = 0.f;
  if (r < n && k < D) x = P[r * ldp + k];
  uint32_t bu, su;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bu) : "f"(x));
  const float b = __uint_as_float(bu);
  const float rem = x - b;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(su) : "f"(rem));
  const float s = __uint_as_float(su);
  const int64_t o = canon_off(r, k, KP);
  big[o] = b;
  small[o] = s;
}

// 3xFP16 mode: the tensor cores run FP16 inputs (FP32 accumulate) at twice the TF32 rate, and FP16 carries the same
// 11 significant bits as TF32 - what it lacks is exponent range, which a per-row power-of-two scale restores:
//   xs = x 2^s (row maximum in [2^14, 2^15)),  H = fp16(xs),  rem = xs - H  (exact, |rem| <= 2^-11 |xs|)
//   row role (tile resident in tensor memory):  H, L' = fp16(rem), H' = fp16(H 2^-11)
//   column role (streamed through shared memory): H, L = fp16(rem 2^11)      - only TWO arrays travel per tile
//   x.z 2^(sx + sz) = sum H_x H_z + L'_x H_z + H'_x L_z   (+ the dropped rem_x rem_z term, 2^-22, as in 3xTF32)
// (values 2^-17 below their row maximum lose bits of L' / H' to FP16 subnormals: an absolute error of 2^-40 of the
// row-norm product).  One warp per row: row maximum -> scale, then the four arrays in the canonical FP16 order;
// 1 / 2^s goes to `rinv`.
__global__ void prepare_f16_kernel(const float* __restrict__ P, const float* __restrict__ norms, int64_t n, int D,
                                   int64_t ldp, int KP, int64_t n_pad, __half* __restrict__ H, __half* __restrict__ L,
                                   __half* __restrict__ Lp, __half* __restrict__ Hp, float* __restrict__ rinv,
                                   float* __restrict__ norms_pad) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n_pad) return;
  float mx = 0.f;
  if (r < n)
    for (int k = lane; k < D; k += 32) mx = fmaxf(mx, fabsf(P[r * ldp + k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  int s = 0;
  if (mx > 0.f && mx < 3.0e38f) {
    int e;
    frexpf(mx, &e);  // mx = f 2^e, f in [0.5, 1)  ->  mx 2^(15 - e) in [2^14, 2^15)
    s = 15 - e;
    s = s > 100 ? 100 : (s < -100 ? -100 : s);
  }
  const float up = exp2f((float)s);
  if (lane == 0) {
    rinv[r] = exp2f((float)-s);
    norms_pad[r] = r < n ? norms[r] : 0.f;
  }
  for (int k = lane; k < KP; k += 32) {
    float x = 0.f;
    if (r < n && k < D) x = P[r * ldp + k] * up;
    const __half h = __float2half_rn(x);
    const float rem = x - __half2float(h);
    const int64_t o = canon_off_h(r, k, KP);
    H[o] = h;
    L[o] = __float2half_rn(rem * 2048.f);
    Lp[o] = __float2half_rn(rem);
    Hp[o] = __float2half_rn(__half2float(h) * (1.f / 2048.f));
  }
}

// ---------------------------------------------------------------------------------------------------------
// float32 kernel values (GPflow formulas; float32 constants as GPflow builds them in the default float)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {  // MUFU.EX2, 2 ulp, flushes results below 2^-126 to 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {  // MUFU.SQRT
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Matern families from r2 (SE takes the fused path in the epilogue: its exponent is linear in the accumulator)
template <int KIND>
__device__ __forceinline__ float kval32(float r2) {
  const float r = sqrt_approx(fmaxf(r2, 1e-36f));
  if constexpr (KIND == CGGP_MATERN12) {
    return ex2_approx(-1.44269504088896340736f * r);
  } else if constexpr (KIND == CGGP_MATERN32) {
    const float s = 1.7320508075688772f * r;
    return (1.f + s) * ex2_approx(-1.44269504088896340736f * s);
  } else {
    const float s = 2.23606797749979f * r;
    return (1.f + s + (float)(5.0 / 3.0) * (r * r)) * ex2_approx(-1.44269504088896340736f * s);
  }
}
constexpr float HALF_LOG2E = 0.72134752044448170368f;  // exp(-r2 / 2) = 2^(-HALF_LOG2E r2)

struct Args {
  const float* Pb;   // canonical big / small parts of the row set (TMEM lanes)
  const float* Ps;
  const float* pn;   // padded norms of the row set
  int64_t np;        // valid rows
  const float* Qb;   // column set
  const float* Qs;
  const float* qn;
  int64_t nq;        // valid columns
  const float* U;    // [NB, ldu] weights over the columns
  int64_t ldu;
  float* out;        // [gridDim.y][NB][ldo] partial results over the rows
  int64_t ldo;
  int64_t q_tiles_per_split;
  float variance;
  int64_t np_pad, nq_pad;  // rows padded to 128 (3xFP16: offsets of the L array and of the row scales)
  int stages;        // Q chunk ring depth (what fits next to the resident P tile)
  int dbg;           // timing experiments only (env CGGP_TF32_DBG; results are WRONG when set): 1 = no epilogue math,
                     // 2 = no MMAs, 4 = no TMA copies, 8 = no global loads of the column scalars, 16 = no tcgen05.ld
  long long* stamps;  // timing experiments: [4 roles][128] clock64 stamps of CTA (0, 0) at tile boundaries, or NULL
  const int* active;
};

template <int KIND, int NSPLIT, int NB>
__global__ void __launch_bounds__(352, 1) gram_contract_kernel(const Args a, const int KP) {
  if (cg_inactive(a.active)) return;
  constexpr bool F16 = NSPLIT == F16X3;
  constexpr int PARTS = NSPLIT > 1 ? 2 : 1;  // arrays of the column set that travel through shared memory
  constexpr unsigned CHB = F16 ? CHUNK_BYTES_H : CHUNK_BYTES;  // bytes of one part of one K chunk
  constexpr int AUXR = (F16 ? 2 : 1) + NB;                     // per-column scalars: |q|^2 term, (scale,) weights
  constexpr int BN = 128;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nchunk = KP >> 5;
  const int STAGES = a.stages;
  unsigned char* sQ = smem_raw;                                          // [STAGES][PARTS][CHB bytes]
  float* aux = reinterpret_cast<float*>(sQ + (size_t)STAGES * PARTS * CHB);  // [2][AUXR * BN]
  // TMEM: columns [0, 256) two accumulators; [256, 256 + PARTS * KP) the P tile as the A operand (row = lane)
  constexpr uint32_t TM_P = 256;
  __shared__ uint64_t bar_p, bar_qfull[MAX_STAGES], bar_qfree[MAX_STAGES], bar_full[2], bar_empty[2], bar_aux[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float comb[BM * NB];  // partial sums of the second column half, combined at the end

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t p0 = (int64_t)blockIdx.x * BM;
  const int64_t q_tiles_total = (a.nq + BN - 1) / BN;
  const int64_t jt0 = (int64_t)blockIdx.y * a.q_tiles_per_split;
  int64_t njt = q_tiles_total - jt0;
  if (njt > a.q_tiles_per_split) njt = a.q_tiles_per_split;
  if (njt < 0) njt = 0;

  if (tid == 0) {
    mbar_init(&bar_p, 8);  // one elected lane per epilogue warp arrives (256 arrivals on one mbarrier serialise)
    for (int b = 0; b < MAX_STAGES; ++b) {
      mbar_init(&bar_qfull[b], 1);
      mbar_init(&bar_qfree[b], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_full[b], 1);
      mbar_init(&bar_empty[b], 8);
      mbar_init(&bar_aux[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 8) {
    // =============================== TMA producer warp ===============================
    if (lane == 0) {
      int64_t g = 0;  // running chunk counter over the whole loop
      for (int64_t j = 0; j < njt; ++j) {
        const int64_t q0 = (jt0 + j) * BN;
        for (int c = 0; c < nchunk; ++c, ++g) {
          const int s = (int)(g % STAGES);
          if (g >= STAGES) mbar_wait(&bar_qfree[s], (unsigned)(((g / STAGES) - 1) & 1));  // MMAs done with this stage
          unsigned char* dst = sQ + (size_t)s * PARTS * CHB;
          const int64_t src = (q0 >> 7) * (int64_t)nchunk * CHUNK_FLOATS + (int64_t)c * CHUNK_FLOATS;  // elements
          if (a.dbg & 4) {
            mbar_arrive(&bar_qfull[s]);
            continue;
          }
          mbar_expect_tx(&bar_qfull[s], PARTS * CHB);
          if constexpr (F16) {
            const __half* QH = reinterpret_cast<const __half*>(a.Qb);
            const __half* QL = QH + a.nq_pad * KP;
            tma_bulk_g2s(dst, QH + src, CHB, &bar_qfull[s]);
            tma_bulk_g2s(dst + CHB, QL + src, CHB, &bar_qfull[s]);
          } else {
            tma_bulk_g2s(dst, a.Qb + src, CHB, &bar_qfull[s]);
            if (PARTS > 1) tma_bulk_g2s(dst + CHB, a.Qs + src, CHB, &bar_qfull[s]);
          }
        }
        if (a.stamps && blockIdx.x == 0 && blockIdx.y == 0 && j < 128) a.stamps[j] = clock64();
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    // =============================== column-scalar warp ===============================
    // per-column scalars of every tile for the epilogue: |q|^2 (SE: its share of the exponent) and the weights U
    // (zero past the end); kept off the TMA warp so that the global-load latency never delays a bulk copy
    for (int64_t j = 0; j < njt; ++j) {
      const int buf = (int)(j & 1);
      const int64_t q0 = (jt0 + j) * BN;
      float vals[BN / 32][AUXR];
      constexpr int UO = AUXR - NB;  // first weight row
#pragma unroll
      for (int i = 0; i < BN / 32; ++i) {
        const int64_t q = q0 + i * 32 + lane;
        if (a.dbg & 8) {
#pragma unroll
          for (int b = 0; b < AUXR; ++b) vals[i][b] = 0.f;
          continue;
        }
        vals[i][0] = KIND == CGGP_SE ? -HALF_LOG2E * a.qn[q] : a.qn[q];  // padded array
        if constexpr (F16) {
          // the column's share of the accumulator scale, folded with the constant the kernel family multiplies by
          const float rq = (a.Qs + a.nq_pad * KP)[q];
          vals[i][1] = (KIND == CGGP_SE ? 2.f * HALF_LOG2E : -2.f) * rq;
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) vals[i][UO + b] = q < a.nq ? a.U[(int64_t)b * a.ldu + q] : 0.f;
      }
      if (j >= 2) mbar_wait(&bar_empty[buf], (unsigned)(((j >> 1) - 1) & 1));  // epilogue of tile j-2 left aux[buf]
      float* ax = aux + buf * AUXR * BN;
#pragma unroll
      for (int i = 0; i < BN / 32; ++i)
#pragma unroll
        for (int b = 0; b < AUXR; ++b) ax[b * BN + i * 32 + lane] = vals[i][b];
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_aux[buf]);
      if (a.stamps && lane == 0 && blockIdx.x == 0 && blockIdx.y == 0 && j < 128) a.stamps[128 + j] = clock64();
    }
  } else if (warp == 9) {
    // =============================== MMA issuer warp ===============================
    if (lane == 0 && njt > 0) {
      const unsigned lbo = 128, sbo = F16 ? 512 : 1024;
      // instruction descriptor: FP32 accumulate, A / B format TF32 (2) resp. F16 (0), K-major, N, M
      const uint32_t idesc = (1u << 4) | ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      mbar_wait(&bar_p, 0);  // the epilogue warps have stored the P tile into tensor memory
      asm volatile("tcgen05.fence::after_thread_sync;");
      int64_t g = 0;
      for (int64_t j = 0; j < njt; ++j) {
        const int buf = (int)(j & 1);
        if (j >= 2) mbar_wait(&bar_empty[buf], (unsigned)(((j >> 1) - 1) & 1));  // accumulator drained by the epilogue
        const uint32_t d = tmem_base + (uint32_t)(buf * BN);
        uint32_t acc = 0;
        for (int c = 0; c < nchunk; ++c, ++g) {
          const int s = (int)(g % STAGES);
          mbar_wait(&bar_qfull[s], (unsigned)((g / STAGES) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;");
          const unsigned qb = smem_u32(sQ) + (unsigned)s * PARTS * CHB, qs = qb + CHB;
          if (a.dbg & 2) {
            umma_commit(&bar_qfree[s]);
            continue;
          }
          if constexpr (F16) {
            // A operand: H at TMEM columns [TM_P, + KP / 2), L' and H' behind it (two features per column); a K = 16
            // step is 8 columns of A and two 128-byte core matrices of B; stage order: H | L
            const uint32_t ph = tmem_base + TM_P + (uint32_t)c * 16, pl = ph + (uint32_t)(KP / 2),
                           php = pl + (uint32_t)(KP / 2);
#pragma unroll
            for (int k = 0; k < 2; ++k) {  // D (+)= H_p H_q^T
              umma_f16_ta(d, ph + k * 8, smem_desc(qb + k * 256, lbo, sbo), idesc, acc);
              acc = 1;
            }
#pragma unroll
            for (int k = 0; k < 2; ++k)  // + L'_p H_q^T
              umma_f16_ta(d, pl + k * 8, smem_desc(qb + k * 256, lbo, sbo), idesc, 1);
#pragma unroll
            for (int k = 0; k < 2; ++k)  // + H'_p L_q^T
              umma_f16_ta(d, php + k * 8, smem_desc(qs + k * 256, lbo, sbo), idesc, 1);
            umma_commit(&bar_qfree[s]);
            continue;
          }
          const uint32_t pb = tmem_base + TM_P + (uint32_t)c * 32, ps = pb + (uint32_t)KP;  // A operand: TMEM columns
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // D (+)= Pb Qb^T
            umma_tf32_ta(d, pb + k * 8, smem_desc(qb + k * 256, lbo, sbo), idesc, acc);
            acc = 1;
          }
          if (PARTS > 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k)  // + Ps Qb^T
              umma_tf32_ta(d, ps + k * 8, smem_desc(qb + k * 256, lbo, sbo), idesc, 1);
#pragma unroll
            for (int k = 0; k < 4; ++k)  // + Pb Qs^T
              umma_tf32_ta(d, pb + k * 8, smem_desc(qs + k * 256, lbo, sbo), idesc, 1);
          }
          umma_commit(&bar_qfree[s]);  // the stage may be refilled once these MMAs are done
        }
        umma_commit(&bar_full[buf]);  // accumulator ready for the epilogue
        if (a.stamps && blockIdx.x == 0 && blockIdx.y == 0 && j < 128) a.stamps[256 + j] = clock64();
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue warps ===============================
    // 8 warps: warp w reads TMEM lanes 32 (w % 4) .. + 31 (thread = row) and the column half w / 4 of every tile
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    const int64_t p = p0 + row;
    const float pn = a.pn[p];  // padded array
    const float cp = -HALF_LOG2E * pn;
    float rp = 1.f;            // 3xFP16: this row's share of the accumulator scale
    if constexpr (F16) rp = (a.Ps + a.np_pad * KP)[p];
    if (njt > 0) {
      // P tile -> tensor memory (A operand of every MMA of this CTA): thread = row = TMEM lane; the two column halves
      // of the epilogue split the features.  Halves the shared-memory operand traffic of the MMAs (only Q is read
      // from shared memory) and frees 96 KB for a deeper Q ring.
      const int kper = KP / 2;  // KP is a multiple of 32
      if constexpr (F16) {
        const __half* PH = reinterpret_cast<const __half*>(a.Pb);
        const __half* PLp = reinterpret_cast<const __half*>(a.Ps);
        for (int part = 0; part < 3; ++part) {
          const __half* src = part == 0 ? PH : (part == 1 ? PLp : PLp + a.np_pad * KP);
          for (int k0 = half * kper; k0 < (half + 1) * kper; k0 += 16) {
            const uint4 x0 = *reinterpret_cast<const uint4*>(src + canon_off_h(p, k0, KP));
            const uint4 x1 = *reinterpret_cast<const uint4*>(src + canon_off_h(p, k0 + 8, KP));
            const uint32_t v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            tmem_st8(tmem_base + ((uint32_t)(quarter * 32) << 16) + TM_P + (uint32_t)(part * (KP / 2) + k0 / 2), v);
          }
        }
      } else
      for (int part = 0; part < PARTS; ++part) {
        const float* src = part == 0 ? a.Pb : a.Ps;
        for (int k0 = half * kper; k0 < (half + 1) * kper; k0 += 8) {
          const float4 x0 = *reinterpret_cast<const float4*>(src + canon_off(p, k0, KP));
          const float4 x1 = *reinterpret_cast<const float4*>(src + canon_off(p, k0 + 4, KP));
          const uint32_t v[8] = {__float_as_uint(x0.x), __float_as_uint(x0.y), __float_as_uint(x0.z),
                                 __float_as_uint(x0.w), __float_as_uint(x1.x), __float_as_uint(x1.y),
                                 __float_as_uint(x1.z), __float_as_uint(x1.w)};
          tmem_st8(tmem_base + ((uint32_t)(quarter * 32) << 16) + TM_P + (uint32_t)(part * KP + k0), v);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_p);
    }
    float acc[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = 0.f;
    for (int64_t j = 0; j < njt; ++j) {
      const int buf = (int)(j & 1);
      mbar_wait(&bar_aux[buf], (unsigned)((j >> 1) & 1));
      mbar_wait(&bar_full[buf], (unsigned)((j >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;");
      const float* ax = aux + buf * AUXR * BN + half * 64;
      constexpr int UO = AUXR - NB;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + half * 64);
      uint32_t v[2][32];
      if (!(a.dbg & 16)) {
        tmem_ld32_nowait(taddr, v[0]);
        tmem_ld32_nowait(taddr + 32, v[1]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
        v[0][0] = v[1][31] = 0;
      }
      // the accumulator is in registers: hand the TMEM buffer back before the math
      asm volatile("tcgen05.fence::before_thread_sync;");
      float part[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) part[b] = 0.f;
      if (a.dbg & 1) {
        part[0] = __uint_as_float(v[0][0] ^ v[1][31]);
      } else
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int c4 = 0; c4 < 32; c4 += 4) {
          const float4 qv = *reinterpret_cast<const float4*>(ax + h * 32 + c4);
          float4 uv[NB];
#pragma unroll
          for (int b = 0; b < NB; ++b) uv[b] = *reinterpret_cast<const float4*>(ax + (UO + b) * BN + h * 32 + c4);
          const float qs[4] = {qv.x, qv.y, qv.z, qv.w};
          float aq[4] = {0.f, 0.f, 0.f, 0.f};
          if constexpr (F16) {
            const float4 av = *reinterpret_cast<const float4*>(ax + BN + h * 32 + c4);
            aq[0] = av.x; aq[1] = av.y; aq[2] = av.z; aq[3] = av.w;
          }
          float kv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float d = __uint_as_float(v[h][c4 + i]);
            if constexpr (F16) {
              // accumulator = 2^(sp + sq) p.q; aq carries 2^-sq and the family's factor on the cross term
              if constexpr (KIND == CGGP_SE) kv[i] = ex2_approx(fmaf(d * aq[i], rp, cp + qs[i]));
              else kv[i] = kval32<KIND>(fmaf(d * aq[i], rp, pn + qs[i]));
            } else if constexpr (KIND == CGGP_SE) {
              // exp(-r2 / 2) with r2 = |p|^2 + |q|^2 - 2 p.q: one FADD + one FFMA + MUFU.EX2
              kv[i] = ex2_approx(fmaf(2.f * HALF_LOG2E, d, cp + qs[i]));
            } else {
              kv[i] = kval32<KIND>(fmaf(-2.f, d, pn + qs[i]));  // GPflow: dist = -2 p.q; dist += |p|^2 + |q|^2
            }
          }
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            part[b] = fmaf(kv[0], uv[b].x, part[b]);
            part[b] = fmaf(kv[1], uv[b].y, part[b]);
            part[b] = fmaf(kv[2], uv[b].z, part[b]);
            part[b] = fmaf(kv[3], uv[b].w, part[b]);
          }
        }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_empty[buf]);  // the warp has read its accumulator slice and aux[buf]
      if (a.stamps && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && j < 128) a.stamps[384 + j] = clock64();
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[b] += part[b];
    }
    if (half == 1) {
#pragma unroll
      for (int b = 0; b < NB; ++b) comb[row * NB + b] = acc[b];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 epilogue warps only
    if (half == 0 && p < a.np) {
#pragma unroll
      for (int b = 0; b < NB; ++b)
        a.out[((int64_t)blockIdx.y * NB + b) * a.ldo + p] = a.variance * (acc[b] + comb[row * NB + b]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

__global__ void reduce_splits_kernel(const float* __restrict__ part, int splits, int NB, int64_t ld, int64_t n,
                                     float* __restrict__ out, int64_t ldo, const int* __restrict__ active) {
  if (cg_inactive(active)) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += part[((int64_t)k * NB + b) * ld + i];
  out[(int64_t)b * ldo + i] = s;
}

using KernelFn = void (*)(const Args, const int);
template <int KIND, int NSPLIT>
static KernelFn pick_nb(int nb) {
  return nb == 1 ? gram_contract_kernel<KIND, NSPLIT, 1> : gram_contract_kernel<KIND, NSPLIT, 2>;
}
template <int KIND>
static KernelFn pick_split(int nsplit, int nb) {
  return nsplit == F16X3 ? pick_nb<KIND, F16X3>(nb) : (nsplit > 1 ? pick_nb<KIND, 3>(nb) : pick_nb<KIND, 1>(nb));
}
static KernelFn pick(int kind, int nsplit, int nb) {
  switch (kind) {
    case CGGP_SE: return pick_split<CGGP_SE>(nsplit, nb);
    case CGGP_MATERN12: return pick_split<CGGP_MATERN12>(nsplit, nb);
    case CGGP_MATERN32: return pick_split<CGGP_MATERN32>(nsplit, nb);
    default: return pick_split<CGGP_MATERN52>(nsplit, nb);
  }
}

constexpr size_t SMEM_BUDGET = 227 * 1024 - 1024;  // leave room for the static barriers
// ring depth that fits next to the resident P tile (0 = does not fit)
static size_t stage_bytes(int nsplit) {
  return nsplit == F16X3 ? 2 * (size_t)CHUNK_BYTES_H : (nsplit > 1 ? 2 : 1) * (size_t)CHUNK_BYTES;
}
static size_t aux_bytes(int nsplit, int nb) {
  return (size_t)2 * ((nsplit == F16X3 ? 2 : 1) + nb) * 128 * sizeof(float) + 128;
}
static int stages_for(int KP, int nsplit, int nb) {
  // the P tile must fit in the 256 tensor-memory columns next to the accumulators (FP16: two features per column)
  const size_t pcols = nsplit == F16X3 ? 3 * (size_t)KP / 2 : (nsplit > 1 ? 2 : 1) * (size_t)KP;
  if (pcols > 256) return 0;
  const size_t fixed = aux_bytes(nsplit, nb);
  if (fixed >= SMEM_BUDGET) return 0;
  size_t st = (SMEM_BUDGET - fixed) / stage_bytes(nsplit);
  if (st > MAX_STAGES) st = MAX_STAGES;
  return st >= 2 ? (int)st : 0;
}
static size_t smem_bytes(int KP, int nsplit, int nb, int stages) {
  return (size_t)stages * stage_bytes(nsplit) + aux_bytes(nsplit, nb);
}
}  // namespace tf32

extern "C" int cggp_tf32_kp(int D) { return (D + 31) / 32 * 32; }
extern "C" int64_t cggp_tf32_rows(int64_t n) { return (n + 127) / 128 * 128; }

extern "C" int cggp_tf32_prepare(cggp_ctx* ctx, const void* P, const void* norms, int64_t n, int D, int64_t ldp,
                                 void* big, void* small, void* norms_pad) {
  if (!ctx) return CGGP_ERR_INVALID;
  if (D < 1 || n < 0) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "bad shape");
  const int KP = cggp_tf32_kp(D);
  const int64_t n_pad = cggp_tf32_rows(n);
  if (n_pad == 0) return CGGP_OK;
  const int64_t total = n_pad * KP;
  tf32::prepare_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
      (const float*)P, (const float*)norms, n, D, ldp, KP, n_pad, (float*)big, (float*)small, (float*)norms_pad);
  CGGP_LAUNCH_CHECK(ctx);
  return CGGP_OK;
}

// 3xFP16 arrays in the buffers cggp_tf32_prepare fills, with `small` longer by the row scales (every entry point
// below stays as it is; pass nsplit = 16 to the products):  big = [H | L] (2 x rows x KP halfs, the column role),
// small = [L' | H' (2 x rows x KP halfs, the row role) | 1 / row scale (rows floats)]
extern "C" int cggp_f16x3_prepare(cggp_ctx* ctx, const void* P, const void* norms, int64_t n, int D, int64_t ldp,
                                  void* big, void* small, void* norms_pad) {
  if (!ctx) return CGGP_ERR_INVALID;
  if (D < 1 || n < 0) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "bad shape");
  const int KP = cggp_tf32_kp(D);
  const int64_t n_pad = cggp_tf32_rows(n);
  if (n_pad == 0) return CGGP_OK;
  __half* H = (__half*)big;
  __half* Lp = (__half*)small;
  float* rinv = (float*)small + n_pad * KP;
  const int warps = 8;
  tf32::prepare_f16_kernel<<<(unsigned)((n_pad + warps - 1) / warps), warps * 32, 0, ctx->stream>>>(
      (const float*)P, (const float*)norms, n, D, ldp, KP, n_pad, H, H + n_pad * KP, Lp, Lp + n_pad * KP, rinv, (float*)norms_pad);
  CGGP_LAUNCH_CHECK(ctx);
  return CGGP_OK;
}

extern "C" int cggp_tf32_supported(cggp_ctx* ctx, int D, int nsplit);
bool cggp_matvec_tf32_supported(cggp_ctx* ctx, int D, int nsplit) {
  const int KP = cggp_tf32_kp(D);
  return ctx->cc_major >= 10 && tf32::stages_for(KP, nsplit, 2) >= 2;
}

extern "C" int cggp_tf32_supported(cggp_ctx* ctx, int D, int nsplit) {
  if (!ctx || D < 1 || (nsplit != 1 && nsplit != 3 && nsplit != tf32::F16X3)) return 0;
  return cggp_matvec_tf32_supported(ctx, D, nsplit) ? 1 : 0;
}

// out[b, p] = variance * sum_q k(P_p, Q_q) U[b, q] for b < B: one sweep, the Q tiles split over grid.y when the row set
// alone does not fill the machine (fixed-order reduction of the partials)
static int gram_sweep(cggp_ctx* ctx, int kind, double variance, int nsplit, int KP, const float* Pb, const float* Ps,
                      const float* pn, int64_t np, const float* Qb, const float* Qs, const float* qn, int64_t nq,
                      const float* U, int64_t ldu, int B, float* out, int64_t ldo, float* scratch, const int* active) {
  using namespace tf32;
  const int64_t p_blocks = (np + BM - 1) / BM, q_tiles = (nq + BN - 1) / BN;
  // grid.y: 1 when the row set alone fills the machine; else the split count (<= ~6 waves) whose CTA total wastes the
  // least of its last wave (1 CTA per SM: 64 row blocks x 7 splits = 448 CTAs would leave a 4-CTA fourth wave)
  int64_t splits = 1;
  if (p_blocks < 2LL * ctx->sm_count) {
    const int64_t sms = ctx->sm_count;
    const int64_t lo = (sms + p_blocks - 1) / p_blocks, hi = (6 * sms + p_blocks - 1) / p_blocks;
    double best = -1.0;
    for (int64_t sp = lo; sp <= hi && sp <= q_tiles; ++sp) {
      const int64_t total = p_blocks * sp;
      const double eff = (double)total / (double)(((total + sms - 1) / sms) * sms);
      if (eff > best + 1e-9) { best = eff; splits = sp; }
    }
    if (splits > q_tiles) splits = q_tiles;
    if (splits < 1) splits = 1;
  }
  const int64_t tiles_per_split = (q_tiles + splits - 1) / splits;
  splits = (q_tiles + tiles_per_split - 1) / tiles_per_split;
  for (int b0 = 0; b0 < B; b0 += 2) {
    const int nb = (B - b0) >= 2 ? 2 : 1;
    KernelFn fn = pick(kind, nsplit, nb);
    const int stages = stages_for(KP, nsplit, nb);
    const size_t smem = smem_bytes(KP, nsplit, nb, stages);
    CGGP_CUDA(ctx, cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Args a;
    a.Pb = Pb; a.Ps = Ps; a.pn = pn; a.np = np;
    a.Qb = Qb; a.Qs = Qs; a.qn = qn; a.nq = nq;
    a.U = U + (int64_t)b0 * ldu; a.ldu = ldu;
    a.out = splits == 1 ? out + (int64_t)b0 * ldo : scratch;
    a.ldo = splits == 1 ? ldo : np;
    a.np_pad = (np + BM - 1) / BM * BM;
    a.nq_pad = (nq + BN - 1) / BN * BN;
    a.q_tiles_per_split = tiles_per_split;
    a.variance = (float)variance;
    a.stages = stages;
    a.dbg = getenv("CGGP_TF32_DBG") ? atoi(getenv("CGGP_TF32_DBG")) : 0;
    a.stamps = nullptr;
    if (getenv("CGGP_TF32_STAMPS")) {
      static long long* dev_stamps = nullptr;
      if (!dev_stamps) cudaMalloc(&dev_stamps, 512 * sizeof(long long));
      cudaMemsetAsync(dev_stamps, 0, 512 * sizeof(long long), ctx->stream);
      a.stamps = dev_stamps;
    }
    a.active = active;
    fn<<<dim3((unsigned)p_blocks, (unsigned)splits), 352, smem, ctx->stream>>>(a, KP);
    CGGP_LAUNCH_CHECK(ctx);
    if (a.stamps) {
      static int dumped = 0;
      if (dumped < 2) {
        ++dumped;
        long long h[512];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(h, a.stamps, sizeof(h), cudaMemcpyDeviceToHost);
        const char* names[4] = {"tma", "aux", "mma", "epi"};
        long long t0 = h[0];
        for (int r = 0; r < 4; ++r)
          for (int j = 0; j < 128; ++j)
            if (h[r * 128 + j] && h[r * 128 + j] < t0) t0 = h[r * 128 + j];
        for (int r = 0; r < 4; ++r) {
          printf("stamps %s:", names[r]);
          for (int j = 0; j < 40; ++j) printf(" %lld", h[r * 128 + j] ? h[r * 128 + j] - t0 : -1);
          printf("\n");
        }
        fflush(stdout);
      }
    }
    if (splits > 1) {
      reduce_splits_kernel<<<dim3((unsigned)((np + 255) / 256), (unsigned)nb), 256, 0, ctx->stream>>>(
          scratch, (int)splits, nb, np, np, out + (int64_t)b0 * ldo, ldo, active);
      CGGP_LAUNCH_CHECK(ctx);
    }
  }
  return CGGP_OK;
}
// scratch for the split partials of a sweep over `np` rows (only row sets that do not fill the machine are split)
static size_t gram_scratch_bytes(cggp_ctx* ctx, int64_t np) {
  const int64_t p_blocks = (np + tf32::BM - 1) / tf32::BM;
  if (p_blocks >= 2LL * ctx->sm_count) return 0;
  const int64_t splits = (6LL * ctx->sm_count + p_blocks - 1) / (p_blocks > 0 ? p_blocks : 1) + 1;
  return sizeof(float) * (size_t)splits * 2 * (size_t)np;
}

// W[B, m] = V[B, m] @ (Kuf Kfu): T = gram(X; Z, V), W = gram(Z; X, T)
int cggp_matvec_tf32(cggp_ctx* ctx, int kind, double variance, const float* Xb, const float* Xs, const float* xn,
                     int64_t n, const float* Zb, const float* Zs, const float* zn, int64_t m, int D, const float* V,
                     int64_t ldv, int B, float* W, int64_t ldw, int nsplit, const int* active) {
  const int KP = cggp_tf32_kp(D);
  if (!cggp_matvec_tf32_supported(ctx, D, nsplit))
    CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "tcgen05 TF32 matvec: D=%d with nsplit=%d does not fit shared memory", D, nsplit);
  if (n == 0) {
    for (int b = 0; b < B; ++b) CGGP_CUDA(ctx, cudaMemsetAsync(W + (int64_t)b * ldw, 0, sizeof(float) * m, ctx->stream));
    return CGGP_OK;
  }
  const int64_t n_pad = cggp_tf32_rows(n);
  // scratch: T [2][n_pad], then the split partials of whichever sweep needs them
  const size_t t_bytes = sizeof(float) * 2 * (size_t)n_pad;
  const size_t sa = gram_scratch_bytes(ctx, n), sb = gram_scratch_bytes(ctx, m);
  const size_t s_bytes = sa > sb ? sa : sb;
  int rc = cggp_ws_reserve(ctx, t_bytes + s_bytes + 256);
  if (rc) return rc;
  float* T = (float*)ctx->ws;
  float* scratch = (float*)((char*)ctx->ws + t_bytes);
  for (int b0 = 0; b0 < B; b0 += 2) {
    const int nb = (B - b0) >= 2 ? 2 : 1;
    rc = gram_sweep(ctx, kind, variance, nsplit, KP, Xb, Xs, xn, n, Zb, Zs, zn, m, V + (int64_t)b0 * ldv, ldv, nb, T,
                    n_pad, scratch, active);
    if (rc) return rc;
    rc = gram_sweep(ctx, kind, variance, nsplit, KP, Zb, Zs, zn, m, Xb, Xs, xn, n, T, n_pad, nb, W + (int64_t)b0 * ldw,
                    ldw, scratch, active);
    if (rc) return rc;
  }
  return CGGP_OK;
}

// W[p, j] = sum_i k(z_j, x_i) Yt[p, i]  (Kuf @ Y with Y given transposed, [P, ldy])
extern "C" int cggp_kuf_times_tf32(cggp_ctx* ctx, int kind, double variance, const void* Xb, const void* Xs,
                                   const void* xn, int64_t n, const void* Zb, const void* Zs, const void* zn,
                                   int64_t m, int D, const void* Yt, int64_t ldy, int P, void* W, int64_t ldw,
                                   int nsplit) {
  if (!ctx) return CGGP_ERR_INVALID;
  if (P <= 0 || m <= 0) return CGGP_OK;
  if (kind < CGGP_SE || kind > CGGP_MATERN52) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "unknown kernel kind %d", kind);
  if (nsplit != 1 && nsplit != 3 && nsplit != tf32::F16X3)
    CGGP_FAIL(ctx, CGGP_ERR_INVALID, "nsplit must be 1, 3 (TF32) or 16 (3xFP16)");
  if (!cggp_matvec_tf32_supported(ctx, D, nsplit))
    CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "tcgen05 TF32 path: D=%d with nsplit=%d does not fit shared memory", D, nsplit);
  if (n == 0) {
    for (int b = 0; b < P; ++b)
      CGGP_CUDA(ctx, cudaMemsetAsync((float*)W + (int64_t)b * ldw, 0, sizeof(float) * m, ctx->stream));
    return CGGP_OK;
  }
  int rc = cggp_ws_reserve(ctx, gram_scratch_bytes(ctx, m) + 256);
  if (rc) return rc;
  ProfScope prof(ctx, 0);
  return gram_sweep(ctx, kind, variance, nsplit, cggp_tf32_kp(D), (const float*)Zb, (const float*)Zs, (const float*)zn,
                    m, (const float*)Xb, (const float*)Xs, (const float*)xn, n, (const float*)Yt, ldy, P, (float*)W, ldw,
                    (float*)ctx->ws, nullptr);
}
