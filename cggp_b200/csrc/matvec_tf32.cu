// float32 matrix-free product  W = V @ (Kuf Kfu)  on the 5th-generation tensor cores (tcgen05, FP32 accumulators in
// TMEM), the float32 counterpart of matvec_pipe.cu (BASELINE configs[4]: N = 2M, D = 90, M = 8192).
//
// At D = 90 the distance contraction dominates (2 N M D flop per sweep), so it runs as a tensor-core GEMM:
//   * cggp_tf32_prepare converts the scaled points ONCE into what the tensor cores read: 128-row tiles x 32-feature
//     chunks in the UMMA "canonical K-major, no swizzle" order, all parts of all chunks of a tile contiguous and
//     followed by the tile's per-row scalars, so that a tile (or any run of its chunks) is ONE TMA bulk copy;
//   * arithmetic (`nsplit`): 3xFP16 (default) - x 2^s = H + R with a per-row power-of-two scale, x.z = H.H + R.H + H.R
//     (kind::f16, the same 11 significant bits per part as TF32 at twice the rate); 3xTF32 - x = big + small,
//     b.b + s.b + b.s (kind::tf32); both keep the expanded squared distance at float32 accuracy (dropped term 2^-22);
//     a single TF32 pass loses ~3 digits to the cancellation in |x|^2 + |z|^2 - 2 x.z;
//   * generic "gram contraction"  out[b, p] = sum_q k(P_p, Q_q) U[b, q]:  a CTA owns 128 P rows (= the 128 TMEM lanes;
//     the parts of the P tile live in TENSOR MEMORY as the A operand of every MMA, next to the accumulators) and walks
//     128-column Q tiles that stream through a ring of stages in shared memory.  Roles (one elected lane each unless
//     noted): warp 8 operand producer (one bulk copy per stage), warp 10 weights producer (the per-call U of a tile,
//     behind the tile's scalars in the same stage), warps 9 and 11 MMA issuers (tiles dealt round-robin; ONE
//     tcgen05.commit per tile), warps 0-7 epilogue in two groups of four that take every other tile (tcgen05.ld,
//     thread = row, r2 = |p|^2 + |q|^2 - 2 p.q, kernel value with MUFU ex2 / sqrt, dot with U; the group releases the
//     tile's stages and its accumulator).  Two (TF32) or three (FP16) TMEM accumulators.
//   What the clock stamps showed: a tcgen05.commit holds its thread ~550 cycles, a bulk copy ~400, an
//   MMA issue ~135 (the 128 x 128 x 32-byte instruction itself runs 64) - the kernel is bound by what single threads
//   can issue per tile, hence few large copies, one commit per tile and two issuers.
//   The product is two such sweeps:  T = gram(X, Z, V)  then  W = gram(Z, X, T)  (roles swapped, X split over
//   grid.y with a fixed-order reduction), i.e. every Gram entry is evaluated twice.
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

// Inside a 128-row x 32-feature chunk (4096 elements) the operands sit in the UMMA canonical K-major order:
// TF32 [(r % 128) / 8][(k % 32) / 4][r % 8][k % 4] (LBO = 128 B, SBO = 1 KB), FP16 [(r % 128) / 8][(k % 32) / 8][r % 8]
// [k % 8] (core matrices of 8 rows x 16 bytes; LBO = 128 B, SBO = 512 B).
// The STREAMED arrays keep everything the products need of a 128-row tile in ONE contiguous block:
//   [K chunk 0: part 0 | part 1] [K chunk 1: ...] ... [trailer: |x|^2 of the 128 rows | 1 / row scale (FP16 mode)]
// so that any run of consecutive K chunks - all parts, and with the last chunk the per-row scalars - is one TMA bulk
// copy (a bulk copy costs its issuing thread ~400 cycles whatever its size: tools/microbench_tma.cu).
constexpr int TRAILER_BYTES = 1024;  // 2 x 128 floats
__host__ __device__ inline int64_t tile_bytes(int KP, int parts, int esz) {
  return (int64_t)(KP >> 5) * parts * 4096 * esz + TRAILER_BYTES;
}
__host__ __device__ inline int64_t stream_off(int64_t r, int k, int KP, int parts, int part) {  // TF32 (floats)
  return (r >> 7) * (tile_bytes(KP, parts, 4) / 4) + ((int64_t)(k >> 5) * parts + part) * 4096 +
         ((((r & 127) >> 3) * 8 + ((k & 31) >> 2)) * 32) + (r & 7) * 4 + (k & 3);
}
__host__ __device__ inline int64_t stream_off_h(int64_t r, int k, int KP, int parts, int part) {  // FP16 (halfs)
  return (r >> 7) * (tile_bytes(KP, parts, 2) / 2) + ((int64_t)(k >> 5) * parts + part) * 4096 +
         ((((r & 127) >> 3) * 4 + ((k & 31) >> 3)) * 64) + (r & 7) * 8 + (k & 7);
}
// float index (from the start of the stream) of scalar `which` (0: |x|^2, 1: 1 / row scale) of row r
__host__ __device__ inline int64_t trailer_off(int64_t r, int KP, int parts, int esz, int which) {
  return ((r >> 7) * tile_bytes(KP, parts, esz) + (tile_bytes(KP, parts, esz) - TRAILER_BYTES)) / 4 + which * 128 +
         (r & 127);
}
constexpr int CHUNK_FLOATS = 4096;                      // 128 rows x 32 features
constexpr unsigned CHUNK_BYTES = CHUNK_FLOATS * 4;      // 16 KB
constexpr unsigned CHUNK_BYTES_H = 4096 * 2;            // the same chunk in FP16: 8 KB
constexpr int MAX_STAGES = 12;
constexpr int F16X3 = 16;                               // `nsplit` code of the 3xFP16 mode

// ---------------------------------------------------------------------------------------------------------
// one-time conversion of prepared points into the canonical TF32 big / small arrays (rows padded to 128)
// ---------------------------------------------------------------------------------------------------------
__global__ void prepare_kernel(const float* __restrict__ P, const float* __restrict__ norms, int64_t n, int D,
                               int64_t ldp, int KP, int64_t n_pad, int parts, float* __restrict__ stream,
                               float* __restrict__ norms_pad) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_pad) {
    const float nv = e < n ? norms[e] : 0.f;
    norms_pad[e] = nv;
    stream[trailer_off(e, KP, parts, 4, 0)] = nv;
    stream[trailer_off(e, KP, parts, 4, 1)] = 1.f;
  }
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
  stream[stream_off(r, k, KP, parts, 0)] = b;
  if (parts > 1) stream[stream_off(r, k, KP, parts, 1)] = s;
}

// 3xFP16 mode: the tensor cores run FP16 inputs (FP32 accumulate) at twice the TF32 rate, and FP16 carries the same
// 11 significant bits as TF32 - what it lacks is exponent range, which a per-row power-of-two scale restores:
//   xs = x 2^s (row maximum in [2^14, 2^15)),  H = fp16(xs),  R = fp16(xs - H)   (xs - H is exact, <= 2^-11 |xs|)
//   x.z 2^(sx + sz) = sum H_x H_z + R_x H_z + H_x R_z   (+ the dropped R_x R_z term, 2^-22, as in 3xTF32)
// Elements 2^-18 below their row maximum have R in the FP16 subnormals (absolute error 2^-25): 2^-40 of the product of
// the row maxima, far below the float32 rounding of the distance.  One warp per row: row maximum -> scale, H | R into
// the interleaved stream (both roles of the point set read it), 1 / 2^s into `rinv`.
__global__ void prepare_f16_kernel(const float* __restrict__ P, const float* __restrict__ norms, int64_t n, int D,
                                   int64_t ldp, int KP, int64_t n_pad, __half* __restrict__ HR,
                                   float* __restrict__ rinv, float* __restrict__ norms_pad) {
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
    const float nv = r < n ? norms[r] : 0.f, ri = exp2f((float)-s);
    rinv[r] = ri;
    norms_pad[r] = nv;
    float* tr = reinterpret_cast<float*>(HR);
    tr[trailer_off(r, KP, 2, 2, 0)] = nv;
    tr[trailer_off(r, KP, 2, 2, 1)] = ri;
  }
  for (int k = lane; k < KP; k += 32) {
    float x = 0.f;
    if (r < n && k < D) x = P[r * ldp + k] * up;
    const __half h = __float2half_rn(x);
    HR[stream_off_h(r, k, KP, 2, 0)] = h;
    HR[stream_off_h(r, k, KP, 2, 1)] = __float2half_rn(x - __half2float(h));
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
  const float* Pb;   // row set (TMEM lanes): interleaved stream (TF32 big | small; FP16 H | R)
  const float* Ps;   // FP16 only: 1 / row scale
  const float* pn;   // padded norms of the row set
  int64_t np;        // valid rows
  const float* Qb;   // column set: interleaved stream
  const float* Qs;   // (the column set's scales and norms travel in the trailers of its stream)
  const float* qn;
  int64_t nq;        // valid columns
  const float* U;    // [NB, ldu] weights over the columns, zero-padded to whole tiles (ldu >= 128 * tiles)
  int64_t ldu;
  float* out;        // [gridDim.y][NB][ldo] partial results over the rows
  int64_t ldo;
  int64_t q_tiles_per_split;
  float variance;
  int stages;        // ring depth in stages (what fits next to the resident P tile)
  int gc;            // K chunks per stage (divides KP / 32): one TMA bulk copy, one full barrier
  int niss;          // MMA-issuing threads in use (1 or 2); the ring holds a multiple of `niss` tiles, so that a stage
                     // is always consumed by the same issuer (which then sees its barrier's phases in order)
  int dbg;           // timing experiments only (env CGGP_TF32_DBG; results are WRONG when set): 1 = no epilogue math,
                     // 2 = no MMAs, 4 = no operand copies, 8 = no weights copies, 16 = no tcgen05.ld, 64 = all CTAs walk
                     // the tiles in the same order
  const int* active;
};

template <int KIND, int NSPLIT, int NB>
__global__ void __launch_bounds__(384, 1) gram_contract_kernel(const Args a, const int KP) {
  if (cg_inactive(a.active)) return;
  constexpr bool F16 = NSPLIT == F16X3;
  constexpr int PARTS = NSPLIT > 1 ? 2 : 1;  // arrays of the column set that travel through shared memory
  constexpr unsigned CHB = F16 ? CHUNK_BYTES_H : CHUNK_BYTES;  // bytes of one part of one K chunk
  // per-column scalars of a tile, behind the last K chunk of its last stage: |q|^2 | 1 / scale | NB x weights
  constexpr unsigned AUX_BYTES = TRAILER_BYTES + NB * 512;
  constexpr int BN = 128;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int nchunk = KP >> 5;
  const int STAGES = a.stages;
  const int GC = a.gc, NG = nchunk / GC;  // chunks per stage, stages per tile
  const unsigned stage_bytes = (unsigned)GC * PARTS * CHB;     // operand bytes of a stage
  const unsigned stage_stride = stage_bytes + AUX_BYTES;
  unsigned char* sQ = smem_raw;                                // [STAGES][GC x PARTS x CHB operands | AUX_BYTES]
  // TMEM: NBUF accumulators of 128 columns, then the P tile as the A operand (row = lane).  FP16 packs two features
  // per column (P tile = KP columns), which leaves room for a THIRD accumulator: an issuing thread can start the next
  // tile while both of the previous ones are still being issued / drained.
  constexpr int NBUF = F16 ? 3 : 2;
  const int NISS = a.niss;  // MMA-issuing threads in use (warps 9, 11)
  // Tiles are dealt round-robin to NBUF accumulators, NISS issuers and 2 epilogue groups.  The per-tile barriers are
  // indexed j % NBAR with NBAR a common multiple of all three, so that every barrier always has the same producer and
  // the same consumer, which therefore sees its phases strictly in order (a parity wait cannot tell phase k from
  // phase k - 2).
  constexpr int NBAR = 6;
  constexpr uint32_t TM_P = NBUF * 128;
  __shared__ uint64_t bar_p, bar_qfull[MAX_STAGES], bar_qfree[MAX_STAGES], bar_full[NBAR], bar_empty[NBAR];
  __shared__ uint32_t tmem_base_s;
  __shared__ float comb[BM * NB];  // partial sums of the second epilogue group, combined at the end

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t p0 = (int64_t)blockIdx.x * BM;
  const int64_t q_tiles_total = (a.nq + BN - 1) / BN;
  const int64_t jt0 = (int64_t)blockIdx.y * a.q_tiles_per_split;
  int64_t njt = q_tiles_total - jt0;
  if (njt > a.q_tiles_per_split) njt = a.q_tiles_per_split;
  if (njt < 0) njt = 0;
  // Every CTA of a sweep needs the same column tiles (operands and per-column scalars); started together and walking
  // them in the same order, all SMs would pull the same lines from the same L2 slices at the same moment.  Each CTA
  // therefore starts its walk at a different tile (a fixed function of its index: results stay reproducible).
  const int64_t rot = (njt > 0 && !(a.dbg & 64)) ? (int64_t)((blockIdx.x * 7u + blockIdx.y * 3u) % (unsigned)njt) : 0;
  auto tile_of = [&](int64_t j) {
    int64_t jj = j + rot;
    if (jj >= njt) jj -= njt;
    return jt0 + jj;
  };

  if (tid == 0) {
    mbar_init(&bar_p, 8);  // one elected lane per epilogue warp arrives (256 arrivals on one mbarrier serialise)
    for (int b = 0; b < MAX_STAGES; ++b) {
      mbar_init(&bar_qfull[b], 2);  // operand producer + weights producer (each with its own expect_tx)
      mbar_init(&bar_qfree[b], 1);
    }
    for (int b = 0; b < NBAR; ++b) {
      mbar_init(&bar_full[b], 1);
      mbar_init(&bar_empty[b], 4);  // one elected lane of each of the 4 warps of the epilogue group
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
      int64_t g = 0;  // running stage counter over the whole loop
      constexpr int ESZ = F16 ? 2 : 4;
      const unsigned char* qstream = reinterpret_cast<const unsigned char*>(a.Qb);
      for (int64_t j = 0; j < njt; ++j) {
        const int64_t tile = tile_of(j);
        for (int gi = 0; gi < NG; ++gi, ++g) {
          const int s = (int)(g % STAGES);
          // the epilogue releases a tile's stages once the MMAs that read them are done
          if (g >= STAGES) mbar_wait(&bar_qfree[s], (unsigned)(((g / STAGES) - 1) & 1));
          if (a.dbg & 4) {
            mbar_arrive(&bar_qfull[s]);
            continue;
          }
          const int64_t src = tile * tile_bytes(KP, PARTS, ESZ) + (int64_t)gi * stage_bytes;
          const unsigned bytes = stage_bytes + (gi == NG - 1 ? (unsigned)TRAILER_BYTES : 0u);  // + the tile's scalars
          mbar_expect_tx(&bar_qfull[s], bytes);
          tma_bulk_g2s(sQ + (size_t)s * stage_stride, qstream + src, bytes, &bar_qfull[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 10) {
    // =============================== weights producer ===============================
    // the per-call weights U of every tile (NB x 512 bytes, from the zero-padded copy) land behind the tile's static
    // scalars; a second thread, so that the operand producer stays at one bulk copy per stage
    if (lane == 0) {
      int64_t g = 0;
      for (int64_t j = 0; j < njt; ++j) {
        const int64_t tile = tile_of(j);
        for (int gi = 0; gi < NG; ++gi, ++g) {
          const int s = (int)(g % STAGES);
          if (g >= STAGES) mbar_wait(&bar_qfree[s], (unsigned)(((g / STAGES) - 1) & 1));
          if (gi != NG - 1 || (a.dbg & 8)) {
            mbar_arrive(&bar_qfull[s]);
            continue;
          }
          mbar_expect_tx(&bar_qfull[s], NB * 512u);
          unsigned char* dst = sQ + (size_t)s * stage_stride + stage_bytes + TRAILER_BYTES;
#pragma unroll
          for (int b = 0; b < NB; ++b)
            tma_bulk_g2s(dst + b * 512, a.U + (int64_t)b * a.ldu + tile * BN, 512u, &bar_qfull[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 9 || warp == 11) {
    // =============================== MMA issuer warps ===============================
    // NISS issuing threads (tiles j = t mod NISS): a single thread issues a 128 x 128 MMA only every ~135 cycles
    // (measured; the instruction itself takes 64) and stalls ~500 more on its commit, so one issuer leaves the tensor
    // pipe half idle.  The threads' MMAs interleave in the pipe; every tile accumulates into buffer j % NBUF and is
    // committed by its issuer.
    const int mt = warp == 9 ? 0 : warp - 10;
    if (lane == 0 && njt > mt && mt < NISS) {
      const unsigned lbo = 128, sbo = F16 ? 512 : 1024;
      // instruction descriptor: FP32 accumulate, A / B format TF32 (2) resp. F16 (0), K-major, N, M
      const uint32_t idesc = (1u << 4) | ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      mbar_wait(&bar_p, 0);  // the epilogue warps have stored the P tile into tensor memory
      asm volatile("tcgen05.fence::after_thread_sync;");
      for (int64_t j = mt; j < njt; j += NISS) {
        const int buf = (int)(j % NBUF);
        int64_t g = j * NG;  // running stage counter of the producer
        if (j >= NBUF) {  // accumulator drained by the epilogue of tile j - NBUF
          mbar_wait(&bar_empty[(j - NBUF) % NBAR], (unsigned)(((j - NBUF) / NBAR) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;");           // its tcgen05.ld before our MMAs
        }
        const uint32_t d = tmem_base + (uint32_t)(buf * BN);
        uint32_t acc = 0;
        for (int gi = 0; gi < NG; ++gi, ++g) {
          const int s = (int)(g % STAGES);
          mbar_wait(&bar_qfull[s], (unsigned)((g / STAGES) & 1));  // TMA bytes landed (async proxy -> async proxy)
          if (a.dbg & 2) continue;
          for (int cc = 0; cc < GC; ++cc) {
            const int c = gi * GC + cc;
            const unsigned qb = smem_u32(sQ) + (unsigned)s * stage_stride + (unsigned)cc * PARTS * CHB, qs = qb + CHB;
            if constexpr (F16) {
              // A operand: H at TMEM columns [TM_P, + KP / 2), R behind it (two features per column); a K = 16 step
              // is 8 columns of A and two 128-byte core matrices of B; chunk order in the stage: H | R
              const uint32_t ph = tmem_base + TM_P + (uint32_t)c * 16, pr = ph + (uint32_t)(KP / 2);
#pragma unroll
              for (int k = 0; k < 2; ++k) {  // D (+)= H_p H_q^T
                umma_f16_ta(d, ph + k * 8, smem_desc(qb + k * 256, lbo, sbo), idesc, acc);
                acc = 1;
              }
#pragma unroll
              for (int k = 0; k < 2; ++k)  // + R_p H_q^T
                umma_f16_ta(d, pr + k * 8, smem_desc(qb + k * 256, lbo, sbo), idesc, 1);
#pragma unroll
              for (int k = 0; k < 2; ++k)  // + H_p R_q^T
                umma_f16_ta(d, ph + k * 8, smem_desc(qs + k * 256, lbo, sbo), idesc, 1);
            } else {
              const uint32_t pb = tmem_base + TM_P + (uint32_t)c * 32, ps = pb + (uint32_t)KP;  // A: TMEM columns
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
            }
          }
        }
        // ONE commit per tile (a tcgen05.commit holds the issuing thread for ~550 cycles; one per K chunk made this
        // thread the bottleneck of the kernel): it publishes the accumulator, and the epilogue releases the stages
        umma_commit(&bar_full[j % NBAR]);  // accumulator ready for the epilogue
      }
    }
    __syncwarp();
  } else {
    // =============================== epilogue warps ===============================
    // 8 warps in two groups of 4: group w / 4 owns accumulator buffer w / 4, i.e. every other tile, and warp w reads
    // TMEM lanes 32 (w % 4) .. + 31 of it (thread = row, all 128 columns in two passes of 64).  While one group waits
    // for its accumulator / its tcgen05.ld, the other group's math keeps the issue slots and the MUFU pipe busy (all 8
    // warps on the same tile left both idle for the ~1000 cycles of hand-offs and loads of every tile).
    const int quarter = warp & 3, half = warp >> 2;  // `half`: group (tile loop) / feature half (P tile gather)
    const int row = quarter * 32 + lane;
    const int64_t p = p0 + row;
    const float pn = a.pn[p];  // padded array
    const float cp = -HALF_LOG2E * pn;
    // 3xFP16: accumulator = 2^(sp + sq) p.q; the row's 2^-sp, folded with the factor of the cross term
    float rp = 1.f;
    if constexpr (F16) rp = (KIND == CGGP_SE ? 2.f * HALF_LOG2E : -2.f) * a.Ps[p];
    if (njt > 0) {
      // P tile -> tensor memory (A operand of every MMA of this CTA): thread = row = TMEM lane; the two column halves
      // of the epilogue split the features.  Halves the shared-memory operand traffic of the MMAs (only Q is read
      // from shared memory) and frees 96 KB for a deeper Q ring.
      const int kper = KP / 2;  // KP is a multiple of 32
      if constexpr (F16) {
        // all loads of the thread's row first (one round trip to memory instead of one per 16 features: the tile
        // loop cannot start before the P tile is in place, and a CTA lives for only ~64 tiles), then the stores
        const __half* PH = reinterpret_cast<const __half*>(a.Pb);
        const int cnt = kper / 16;  // <= 4: KP <= 128 in this mode
        uint4 x[2][8];
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int it = 0; it < 4; ++it)
            if (it < cnt) {
              const int k0 = half * kper + it * 16;
              x[part][2 * it] = *reinterpret_cast<const uint4*>(PH + stream_off_h(p, k0, KP, 2, part));
              x[part][2 * it + 1] = *reinterpret_cast<const uint4*>(PH + stream_off_h(p, k0 + 8, KP, 2, part));
            }
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int it = 0; it < 4; ++it)
            if (it < cnt) {
              const int k0 = half * kper + it * 16;
              const uint4 x0 = x[part][2 * it], x1 = x[part][2 * it + 1];
              const uint32_t v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
              tmem_st8(tmem_base + ((uint32_t)(quarter * 32) << 16) + TM_P + (uint32_t)(part * (KP / 2) + k0 / 2), v);
            }
      } else
      for (int part = 0; part < PARTS; ++part) {
        const float* src = a.Pb;
        for (int k0 = half * kper; k0 < (half + 1) * kper; k0 += 8) {
          const float4 x0 = *reinterpret_cast<const float4*>(src + stream_off(p, k0, KP, PARTS, part));
          const float4 x1 = *reinterpret_cast<const float4*>(src + stream_off(p, k0 + 4, KP, PARTS, part));
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
    for (int64_t j = half; j < njt; j += 2) {
      const int buf = (int)(j % NBUF);
      mbar_wait(&bar_full[j % NBAR], (unsigned)((j / NBAR) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;");
      // The MMAs of tile j are done, so its bytes have landed and its stages cannot have been refilled: the phase of
      // the last stage's barrier is exactly the tile's.  Waiting on it (it completes at once) makes the scalars the
      // TMA wrote behind the operands visible to this thread.
      const int64_t glast = j * NG + NG - 1;
      const int slast = (int)(glast % STAGES);
      mbar_wait(&bar_qfull[slast], (unsigned)((glast / STAGES) & 1));
      if (quarter == 0 && lane == 0) {  // all stages but the one holding the scalars may be refilled
        for (int gi = 0; gi < NG - 1; ++gi) mbar_arrive(&bar_qfree[(int)((j * NG + gi) % STAGES)]);
      }
      const float* axt = reinterpret_cast<const float*>(sQ + (size_t)slast * stage_stride + stage_bytes);
      constexpr int UO = 2;  // trailer rows: |q|^2, 1 / scale, weights
      float part[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) part[b] = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {  // the two 64-column halves of the tile
        const float* ax = axt + ch * 64;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BN + ch * 64);
        uint32_t v[2][32];
        if (!(a.dbg & 16)) {
          tmem_ld32_nowait(taddr, v[0]);
          tmem_ld32_nowait(taddr + 32, v[1]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
          v[0][0] = v[1][31] = 0;
        }
        if (a.dbg & 1) {
          part[0] += __uint_as_float(v[0][0] ^ v[1][31]);
          continue;
        }
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
                // aq = 2^-sq of the column, rp = 2^-sp of the row times the factor of the cross term
                if constexpr (KIND == CGGP_SE) kv[i] = ex2_approx(fmaf(d * aq[i], rp, fmaf(-HALF_LOG2E, qs[i], cp)));
                else kv[i] = kval32<KIND>(fmaf(d * aq[i], rp, pn + qs[i]));
              } else if constexpr (KIND == CGGP_SE) {
                // exp(-r2 / 2) with r2 = |p|^2 + |q|^2 - 2 p.q: two FFMA + MUFU.EX2
                kv[i] = ex2_approx(fmaf(2.f * HALF_LOG2E, d, fmaf(-HALF_LOG2E, qs[i], cp)));
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
      }
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_empty[j % NBAR]);  // the warp has read its accumulator slice
      asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");  // the 4 warps of the group have read the scalars
      if (quarter == 0 && lane == 0) mbar_arrive(&bar_qfree[slast]);
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

// zero-padded copy of the weights, so that every 128-column tile of them is one aligned 512-byte bulk copy
__global__ void pad_weights_kernel(const float* __restrict__ U, int64_t ldu, int64_t nq, int64_t nq_pad,
                                   float* __restrict__ out, const int* __restrict__ active) {
  if (cg_inactive(active)) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nq_pad) out[(int64_t)blockIdx.y * nq_pad + i] = i < nq ? U[(int64_t)blockIdx.y * ldu + i] : 0.f;
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
static size_t chunk_bytes(int nsplit) {  // all streamed parts of one K chunk of a 128-row tile
  return nsplit == F16X3 ? 2 * (size_t)CHUNK_BYTES_H : (nsplit > 1 ? 2 : 1) * (size_t)CHUNK_BYTES;
}
static size_t aux_bytes(int nb) {  // per stage: the tile's static scalars + nb rows of weights
  return (size_t)TRAILER_BYTES + (size_t)nb * 512;
}
struct RingPlan {
  int gc = 0, stages = 0, niss = 0;  // K chunks per stage, ring depth (0 = does not fit), issuing threads
};
// Stage = `gc` consecutive K chunks (one bulk copy).  A whole tile per stage keeps the producer at one copy per tile
// (its per-copy cost is what matters when the MMAs of a tile are short: FP16 and single-pass TF32); 3xTF32 tiles are
// long enough for one copy per chunk, which overlaps better.  A tile's stages are released together, so the ring
// must hold at least one tile - two for overlap.
static RingPlan ring_for(int KP, int nsplit, int nb) {
  RingPlan r;
  // the P tile must fit in the 256 tensor-memory columns next to the accumulators (FP16: two features per column)
  const size_t pcols = nsplit == F16X3 ? (size_t)KP : (nsplit > 1 ? 2 : 1) * (size_t)KP;
  if (pcols > (nsplit == F16X3 ? 128u : 256u)) return r;
  const int nchunk = KP / 32;
  const size_t room = SMEM_BUDGET;
  // two issuers where the tiles are long enough to profit (a third one did not pay: the 13th warp caps the epilogue
  // at 128 registers), then one; for each, the coarsest stage that still leaves `niss` whole tiles (and at least
  // two) in the ring
  // measured at the c5 shape (ms per product, 1 / 2 issuers): 3xFP16 5.02 / 4.17, 3xTF32 8.63 / 7.88, 1xTF32 4.14 / 5.08
  static const int niss_env = getenv("CGGP_TF32_NISS") ? atoi(getenv("CGGP_TF32_NISS")) : 0;  // tuning knob
  const int niss_max = niss_env > 0 ? niss_env : (nsplit == 1 ? 1 : 2);
  for (int niss = niss_max >= 2 ? 2 : 1; niss >= 1; --niss) {
    for (int gc = (nsplit == 3 ? 1 : nchunk); gc >= 1; --gc) {
      if (nchunk % gc) continue;
      const int ng = nchunk / gc;
      size_t st = room / (chunk_bytes(nsplit) * gc + aux_bytes(nb));
      if (st > MAX_STAGES) st = MAX_STAGES;
      st = st / (size_t)(ng * niss) * (size_t)(ng * niss);  // whole groups of `niss` tiles
      if ((int)st >= 2 * ng) {
        r.gc = gc;
        r.stages = (int)st;
        r.niss = niss;
        return r;
      }
    }
  }
  // last resort (wide tiles): one issuer and a single tile in the ring, in the finest stages that hold it
  for (int gc = 1; gc <= nchunk; ++gc) {
    if (nchunk % gc) continue;
    const int ng = nchunk / gc;
    size_t st = room / (chunk_bytes(nsplit) * gc + aux_bytes(nb));
    if (st > MAX_STAGES) st = MAX_STAGES;
    st = st / (size_t)ng * (size_t)ng;
    if ((int)st >= ng) {
      r.gc = gc;
      r.stages = (int)st;
      r.niss = 1;
      return r;
    }
  }
  return r;
}
static size_t smem_bytes(int nsplit, int nb, const RingPlan& r) {
  return (size_t)r.stages * (r.gc * chunk_bytes(nsplit) + aux_bytes(nb));
}
}  // namespace tf32

extern "C" int cggp_tf32_kp(int D) { return (D + 31) / 32 * 32; }
extern "C" int64_t cggp_tf32_rows(int64_t n) { return (n + 127) / 128 * 128; }

// Buffer sizes in floats for a point set of n rows (header): dev_stream holds the interleaved arrays the products
// stream (TF32 big [| small]; FP16 H | R), dev_rows the 1 / row scales of the 3xFP16 mode.
extern "C" int cggp_tf32_sizes(int nsplit, int64_t n, int D, int64_t* stream_floats, int64_t* rows_floats) {
  if (nsplit != 1 && nsplit != 3 && nsplit != tf32::F16X3) return CGGP_ERR_INVALID;
  const int64_t rows = cggp_tf32_rows(n), kp = cggp_tf32_kp(D);
  if (stream_floats)
    *stream_floats = (rows / 128) * (tf32::tile_bytes((int)kp, nsplit == 1 ? 1 : 2, nsplit == tf32::F16X3 ? 2 : 4) / 4);
  if (rows_floats) *rows_floats = nsplit == tf32::F16X3 ? (rows > 0 ? rows : 1) : 1;
  return CGGP_OK;
}

extern "C" int cggp_tf32_prepare(cggp_ctx* ctx, int nsplit, const void* P, const void* norms, int64_t n, int D,
                                 int64_t ldp, void* stream, void* rows_buf, void* norms_pad) {
  if (!ctx) return CGGP_ERR_INVALID;
  CGGP_DEVICE_GUARD(ctx);
  if (D < 1 || n < 0) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "bad shape");
  if (nsplit != 1 && nsplit != 3 && nsplit != tf32::F16X3)
    CGGP_FAIL(ctx, CGGP_ERR_INVALID, "nsplit must be 1, 3 (TF32) or 16 (3xFP16)");
  const int KP = cggp_tf32_kp(D);
  const int64_t n_pad = cggp_tf32_rows(n);
  if (n_pad == 0) return CGGP_OK;
  if (nsplit == tf32::F16X3) {
    const int warps = 8;
    tf32::prepare_f16_kernel<<<(unsigned)((n_pad + warps - 1) / warps), warps * 32, 0, ctx->stream>>>(
        (const float*)P, (const float*)norms, n, D, ldp, KP, n_pad, (__half*)stream, (float*)rows_buf,
        (float*)norms_pad);
  } else {
    const int64_t total = n_pad * KP;
    tf32::prepare_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
        (const float*)P, (const float*)norms, n, D, ldp, KP, n_pad, nsplit == 3 ? 2 : 1, (float*)stream,
        (float*)norms_pad);
  }
  CGGP_LAUNCH_CHECK(ctx);
  return CGGP_OK;
}

// The shared-memory ring the kernel would run with (host-side query, no GPU needed; tests check its invariants):
// K chunks per stage, stages, MMA-issuing threads, dynamic shared memory.  0 stages = does not fit.
extern "C" int cggp_tf32_ring_plan(int nsplit, int D, int nb, int* gc, int* stages, int* niss, int64_t* smem_bytes_out) {
  if (D < 1 || nb < 1 || nb > 2 || (nsplit != 1 && nsplit != 3 && nsplit != tf32::F16X3)) return CGGP_ERR_INVALID;
  const tf32::RingPlan r = tf32::ring_for(cggp_tf32_kp(D), nsplit, nb);
  if (gc) *gc = r.gc;
  if (stages) *stages = r.stages;
  if (niss) *niss = r.niss;
  if (smem_bytes_out) *smem_bytes_out = r.stages ? (int64_t)tf32::smem_bytes(nsplit, nb, r) : 0;
  return CGGP_OK;
}

extern "C" int cggp_tf32_supported(cggp_ctx* ctx, int D, int nsplit);
bool cggp_matvec_tf32_supported(cggp_ctx* ctx, int D, int nsplit) {
  const int KP = cggp_tf32_kp(D);
  return ctx->cc_major >= 10 && tf32::ring_for(KP, nsplit, 2).stages >= 1;
}

extern "C" int cggp_tf32_supported(cggp_ctx* ctx, int D, int nsplit) {
  if (!ctx || D < 1 || (nsplit != 1 && nsplit != 3 && nsplit != tf32::F16X3)) return 0;
  return cggp_matvec_tf32_supported(ctx, D, nsplit) ? 1 : 0;
}

// out[b, p] = variance * sum_q k(P_p, Q_q) U[b, q] for b < B: one sweep, the Q tiles split over grid.y when the row set
// alone does not fill the machine (fixed-order reduction of the partials)
static int gram_sweep(cggp_ctx* ctx, int kind, double variance, int nsplit, int KP, const float* Pb, const float* Ps,
                      const float* pn, int64_t np, const float* Qb, const float* Qs, const float* qn, int64_t nq,
                      const float* U, int64_t ldu, int B, float* out, int64_t ldo, float* scratch, float* upad,
                      const int* active) {
  using namespace tf32;
  const int64_t nq_pad = (nq + BN - 1) / BN * BN;  // `upad`: [2][nq_pad] floats
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
    const RingPlan ring = ring_for(KP, nsplit, nb);
    const size_t smem = smem_bytes(nsplit, nb, ring);
    CGGP_CUDA(ctx, cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Args a;
    a.Pb = Pb; a.Ps = Ps; a.pn = pn; a.np = np;
    a.Qb = Qb; a.Qs = Qs; a.qn = qn; a.nq = nq;
    pad_weights_kernel<<<dim3((unsigned)((nq_pad + 255) / 256), (unsigned)nb), 256, 0, ctx->stream>>>(
        U + (int64_t)b0 * ldu, ldu, nq, nq_pad, upad, active);
    CGGP_LAUNCH_CHECK(ctx);
    a.U = upad; a.ldu = nq_pad;
    a.out = splits == 1 ? out + (int64_t)b0 * ldo : scratch;
    a.ldo = splits == 1 ? ldo : np;
    a.q_tiles_per_split = tiles_per_split;
    a.variance = (float)variance;
    a.stages = ring.stages;
    a.gc = ring.gc;
    a.niss = ring.niss;
#ifdef CGGP_DEBUG_KNOBS
    // timing experiments that switch parts of the product OFF (wrong results): only in builds with -DCGGP_DEBUG_KNOBS
    static const int dbg_env = getenv("CGGP_TF32_DBG") ? atoi(getenv("CGGP_TF32_DBG")) : 0;
    a.dbg = dbg_env;
#else
    a.dbg = 0;  // the shipped library ignores CGGP_TF32_DBG
#endif
    a.active = active;
    fn<<<dim3((unsigned)p_blocks, (unsigned)splits), 384, smem, ctx->stream>>>(a, KP);
    CGGP_LAUNCH_CHECK(ctx);
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
  const size_t s_bytes = ((sa > sb ? sa : sb) + 255) / 256 * 256;
  const int64_t m_pad = cggp_tf32_rows(m);
  const size_t u_bytes = sizeof(float) * 2 * (size_t)(n_pad > m_pad ? n_pad : m_pad);
  int rc = cggp_ws_reserve(ctx, t_bytes + s_bytes + u_bytes + 512);
  if (rc) return rc;
  float* T = (float*)ctx->ws;
  float* scratch = (float*)((char*)ctx->ws + t_bytes);
  float* upad = (float*)((char*)ctx->ws + t_bytes + s_bytes);
  for (int b0 = 0; b0 < B; b0 += 2) {
    const int nb = (B - b0) >= 2 ? 2 : 1;
    rc = gram_sweep(ctx, kind, variance, nsplit, KP, Xb, Xs, xn, n, Zb, Zs, zn, m, V + (int64_t)b0 * ldv, ldv, nb, T,
                    n_pad, scratch, upad, active);
    if (rc) return rc;
    rc = gram_sweep(ctx, kind, variance, nsplit, KP, Zb, Zs, zn, m, Xb, Xs, xn, n, T, n_pad, nb, W + (int64_t)b0 * ldw,
                    ldw, scratch, upad, active);
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
  CGGP_DEVICE_GUARD(ctx);
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
  const size_t s_bytes = (gram_scratch_bytes(ctx, m) + 255) / 256 * 256;
  int rc = cggp_ws_reserve(ctx, s_bytes + sizeof(float) * 2 * (size_t)cggp_tf32_rows(n) + 512);
  if (rc) return rc;
  ProfScope prof(ctx, 0);
  return gram_sweep(ctx, kind, variance, nsplit, cggp_tf32_kp(D), (const float*)Zb, (const float*)Zs, (const float*)zn,
                    m, (const float*)Xb, (const float*)Xs, (const float*)xn, n, (const float*)Yt, ldy, P, (float*)W, ldw,
                    (float*)ctx->ws, (float*)((char*)ctx->ws + s_bytes), nullptr);
}
