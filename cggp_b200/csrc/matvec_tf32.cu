// float32 matrix-free product  W = V @ (Kuf Kfu)  on the 5th-generation tensor cores (tcgen05, TF32 inputs, FP32
// accumulators in TMEM), the float32 counterpart of matvec_pipe.cu (BASELINE configs[4]: N = 2M, D = 90, M = 8192).
//
// At D = 90 the distance contraction dominates (2 N M D flop per sweep), so it runs as a TF32 GEMM:
//   * the scaled points are converted ONCE (cggp_tf32_prepare) into the UMMA "canonical K-major, no swizzle" order
//     [row / 8][k / 4][row % 8][k % 4] and split  x = x_big + x_small  (both round-to-nearest TF32), so that a
//     128-row tile is one contiguous block that a single TMA bulk copy (cp.async.bulk) drops into shared memory in
//     exactly the layout the tcgen05 shared-memory descriptors describe (LBO = 128 B, SBO = KP / 4 * 128 B);
//   * 3xTF32:  x.z ~ xb.zb + xs.zb + xb.zs  (three accumulating MMAs; the dropped xs.zs term is 2^-22 relative), which
//     keeps the expanded squared distance at float32 accuracy - a single TF32 pass (NSPLIT = 1) loses ~3 digits to the
//     cancellation in |x|^2 + |z|^2 - 2 x.z;
//   * generic "gram contraction"  out[b, p] = sum_q k(P_p, Q_q) U[b, q]:  a CTA owns 128 P rows (= the 128 TMEM lanes)
//     and loops over 128-column Q tiles.  Warp 4 is the producer (one elected lane: TMA of the next Q tile, the
//     tcgen05.mma chain, tcgen05.commit onto mbarriers); warps 0-3 are the epilogue: tcgen05.ld of a finished
//     accumulator (thread = row, registers = columns), r2 = |p|^2 + |q|^2 - 2 p.q, kernel value with MUFU ex2 / sqrt,
//     dot with U.  Two TMEM accumulators (2 x 128 columns) let the MMAs of tile j+1 run under the epilogue of tile j.
//   The product is two such sweeps:  T = gram(X, Z, V)  then  W = gram(Z, X, T)  (roles swapped, X split over
//   grid.y with a fixed-order reduction), i.e. every Gram entry is evaluated twice - parking a 128 x 128 FP32 tile per
//   block would be the next step, as matvec_pipe.cu does for float64.
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

// offset (in floats) of element (r, k) in the canonical order for a point set with KP padded features
__host__ __device__ inline int64_t canon_off(int64_t r, int k, int KP) {
  return ((r >> 3) * (KP >> 2) + (k >> 2)) * 32 + (r & 7) * 4 + (k & 3);
}

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

// ---------------------------------------------------------------------------------------------------------
// float32 kernel values (GPflow formulas; float32 constants as GPflow builds them in the default float)
// ---------------------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ float kval32(float r2) {
  if constexpr (KIND == CGGP_SE) {
    return exp2f(-0.72134752044448170368f * r2);  // exp(-r2 / 2), MUFU.EX2
  } else {
    const float r = sqrtf(fmaxf(r2, 1e-36f));
    if constexpr (KIND == CGGP_MATERN12) {
      return exp2f(-1.44269504088896340736f * r);
    } else if constexpr (KIND == CGGP_MATERN32) {
      const float s = 1.7320508075688772f * r;
      return (1.f + s) * exp2f(-1.44269504088896340736f * s);
    } else {
      const float s = 2.23606797749979f * r;
      return (1.f + s + (float)(5.0 / 3.0) * (r * r)) * exp2f(-1.44269504088896340736f * s);
    }
  }
}

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
  const int* active;
};

template <int KIND, int NSPLIT, int NB>
__global__ void __launch_bounds__(160, 1) gram_contract_kernel(const Args a, const int KP) {
  if (cg_inactive(a.active)) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const unsigned tile_bytes = (unsigned)BM * KP * sizeof(float);
  float* sPb = reinterpret_cast<float*>(smem_raw);
  float* sPs = reinterpret_cast<float*>(smem_raw + (NSPLIT > 1 ? tile_bytes : 0));
  float* sQb = reinterpret_cast<float*>(smem_raw + (NSPLIT > 1 ? 2 : 1) * tile_bytes);
  float* sQs = reinterpret_cast<float*>(smem_raw + (NSPLIT > 1 ? 3 : 1) * tile_bytes);
  float* aux = reinterpret_cast<float*>(smem_raw + (NSPLIT > 1 ? 4 : 2) * tile_bytes);  // [2][(1 + NB) * BN]
  __shared__ uint64_t bar_p, bar_q, bar_qfree, bar_full[2], bar_empty[2], bar_aux[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t p0 = (int64_t)blockIdx.x * BM;
  const int64_t q_tiles_total = (a.nq + BN - 1) / BN;
  const int64_t jt0 = (int64_t)blockIdx.y * a.q_tiles_per_split;
  int64_t njt = q_tiles_total - jt0;
  if (njt > a.q_tiles_per_split) njt = a.q_tiles_per_split;
  if (njt < 0) njt = 0;

  if (tid == 0) {
    mbar_init(&bar_p, 1);
    mbar_init(&bar_q, 1);
    mbar_init(&bar_qfree, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_full[b], 1);
      mbar_init(&bar_empty[b], 128);
      mbar_init(&bar_aux[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 4) {
    // =============================== producer warp ===============================
    const unsigned lbo = 128, sbo = (unsigned)(KP / 4) * 128;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    if (lane == 0) {
      mbar_expect_tx(&bar_p, NSPLIT > 1 ? 2 * tile_bytes : tile_bytes);
      tma_bulk_g2s(sPb, a.Pb + p0 * KP, tile_bytes, &bar_p);
      if (NSPLIT > 1) tma_bulk_g2s(sPs, a.Ps + p0 * KP, tile_bytes, &bar_p);
    }
    for (int64_t j = 0; j < njt; ++j) {
      const int buf = (int)(j & 1);
      const int64_t q0 = (jt0 + j) * BN;
      if (j >= 2) mbar_wait(&bar_empty[buf], (unsigned)(((j >> 1) - 1) & 1));  // epilogue of tile j-2 left buf / aux
      // per-column scalars of this tile for the epilogue: |q|^2 and the weights U (zero past the end)
      float* ax = aux + buf * (1 + NB) * BN;
      for (int c = lane; c < BN; c += 32) {
        const int64_t q = q0 + c;
        ax[c] = a.qn[q];  // padded array
#pragma unroll
        for (int b = 0; b < NB; ++b) ax[(1 + b) * BN + c] = q < a.nq ? a.U[(int64_t)b * a.ldu + q] : 0.f;
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bar_aux[buf]);
        if (j >= 1) mbar_wait(&bar_qfree, (unsigned)((j - 1) & 1));  // the MMAs of tile j-1 have read the Q tile
        mbar_expect_tx(&bar_q, NSPLIT > 1 ? 2 * tile_bytes : tile_bytes);
        tma_bulk_g2s(sQb, a.Qb + q0 * KP, tile_bytes, &bar_q);
        if (NSPLIT > 1) tma_bulk_g2s(sQs, a.Qs + q0 * KP, tile_bytes, &bar_q);
        if (j == 0) mbar_wait(&bar_p, 0);
        mbar_wait(&bar_q, (unsigned)(j & 1));
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t d = tmem_base + (uint32_t)(buf * BN);
        uint32_t acc = 0;
        for (int k = 0; k < KP / 8; ++k) {  // D (+)= Pb Qb^T
          umma_tf32(d, smem_desc(smem_u32(sPb) + k * 256, lbo, sbo), smem_desc(smem_u32(sQb) + k * 256, lbo, sbo), idesc,
                    acc);
          acc = 1;
        }
        if (NSPLIT > 1) {
          for (int k = 0; k < KP / 8; ++k)  // + Ps Qb^T
            umma_tf32(d, smem_desc(smem_u32(sPs) + k * 256, lbo, sbo), smem_desc(smem_u32(sQb) + k * 256, lbo, sbo),
                      idesc, 1);
          for (int k = 0; k < KP / 8; ++k)  // + Pb Qs^T
            umma_tf32(d, smem_desc(smem_u32(sPb) + k * 256, lbo, sbo), smem_desc(smem_u32(sQs) + k * 256, lbo, sbo),
                      idesc, 1);
        }
        umma_commit(&bar_qfree);      // Q tile may be overwritten
        umma_commit(&bar_full[buf]);  // accumulator ready for the epilogue
      }
      __syncwarp();
    }
  } else {
    // =============================== epilogue warps (thread = row) ===============================
    const int64_t p = p0 + tid;
    const float pn = a.pn[p];  // padded array
    float acc[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = 0.f;
    for (int64_t j = 0; j < njt; ++j) {
      const int buf = (int)(j & 1);
      mbar_wait(&bar_aux[buf], (unsigned)((j >> 1) & 1));
      mbar_wait(&bar_full[buf], (unsigned)((j >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;");
      const float* ax = aux + buf * (1 + NB) * BN;
      float part[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) part[b] = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * BN + c0), v);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          // GPflow: dist = -2 p.q; dist += |p|^2 + |q|^2
          const float r2 = fmaf(-2.f, __uint_as_float(v[c]), pn + ax[c0 + c]);
          const float kv = kval32<KIND>(r2);
#pragma unroll
          for (int b = 0; b < NB; ++b) part[b] = fmaf(kv, ax[(1 + b) * BN + c0 + c], part[b]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;");
      mbar_arrive(&bar_empty[buf]);
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[b] += part[b];
    }
    if (p < a.np) {
#pragma unroll
      for (int b = 0; b < NB; ++b) a.out[((int64_t)blockIdx.y * NB + b) * a.ldo + p] = a.variance * acc[b];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base));
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
  return nsplit > 1 ? pick_nb<KIND, 3>(nb) : pick_nb<KIND, 1>(nb);
}
static KernelFn pick(int kind, int nsplit, int nb) {
  switch (kind) {
    case CGGP_SE: return pick_split<CGGP_SE>(nsplit, nb);
    case CGGP_MATERN12: return pick_split<CGGP_MATERN12>(nsplit, nb);
    case CGGP_MATERN32: return pick_split<CGGP_MATERN32>(nsplit, nb);
    default: return pick_split<CGGP_MATERN52>(nsplit, nb);
  }
}

static size_t smem_bytes(int KP, int nsplit, int nb) {
  return (size_t)(nsplit > 1 ? 4 : 2) * BM * KP * sizeof(float) + (size_t)2 * (1 + nb) * BN * sizeof(float) + 128;
}
}  // namespace tf32

extern "C" int cggp_tf32_kp(int D) { return (D + 7) / 8 * 8; }
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

bool cggp_matvec_tf32_supported(cggp_ctx* ctx, int D, int nsplit) {
  const int KP = cggp_tf32_kp(D);
  return ctx->cc_major >= 10 && tf32::smem_bytes(KP, nsplit, 2) <= 227 * 1024;
}

// W[B, m] = V[B, m] @ (Kuf Kfu): T = gram(X; Z, V), W = gram(Z; X, T)
int cggp_matvec_tf32(cggp_ctx* ctx, int kind, double variance, const float* Xb, const float* Xs, const float* xn,
                     int64_t n, const float* Zb, const float* Zs, const float* zn, int64_t m, int D, const float* V,
                     int64_t ldv, int B, float* W, int64_t ldw, int nsplit, const int* active) {
  using namespace tf32;
  const int KP = cggp_tf32_kp(D);
  if (!cggp_matvec_tf32_supported(ctx, D, nsplit))
    CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "tcgen05 TF32 matvec: D=%d with nsplit=%d does not fit shared memory", D, nsplit);
  if (n == 0) {
    for (int b = 0; b < B; ++b) CGGP_CUDA(ctx, cudaMemsetAsync(W + (int64_t)b * ldw, 0, sizeof(float) * m, ctx->stream));
    return CGGP_OK;
  }
  const int64_t p_blocks_x = (n + BM - 1) / BM, p_blocks_z = (m + BM - 1) / BM;
  const int64_t x_tiles = (n + BN - 1) / BN;
  // sweep 2 splits the X tiles over grid.y so that ~3 waves of CTAs are in flight
  int64_t splits = (3LL * ctx->sm_count + p_blocks_z - 1) / p_blocks_z;
  if (splits > x_tiles) splits = x_tiles;
  if (splits < 1) splits = 1;
  const int64_t tiles_per_split = (x_tiles + splits - 1) / splits;
  splits = (x_tiles + tiles_per_split - 1) / tiles_per_split;
  const int64_t n_pad = cggp_tf32_rows(n);
  // scratch: T [2][n_pad] and the sweep-2 partials [splits][2][m]
  const size_t t_bytes = sizeof(float) * 2 * (size_t)n_pad, wp_bytes = sizeof(float) * (size_t)splits * 2 * (size_t)m;
  int rc = cggp_ws_reserve(ctx, t_bytes + wp_bytes + 256);
  if (rc) return rc;
  float* T = (float*)ctx->ws;
  float* Wp = (float*)((char*)ctx->ws + t_bytes);
  for (int b0 = 0; b0 < B; b0 += 2) {
    const int nb = (B - b0) >= 2 ? 2 : 1;
    KernelFn fn = pick(kind, nsplit, nb);
    const size_t smem = smem_bytes(KP, nsplit, nb);
    CGGP_CUDA(ctx, cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Args a1;
    a1.Pb = Xb; a1.Ps = Xs; a1.pn = xn; a1.np = n;
    a1.Qb = Zb; a1.Qs = Zs; a1.qn = zn; a1.nq = m;
    a1.U = V + (int64_t)b0 * ldv; a1.ldu = ldv;
    a1.out = T; a1.ldo = n_pad;
    a1.q_tiles_per_split = (m + BN - 1) / BN;
    a1.variance = (float)variance;
    a1.active = active;
    fn<<<dim3((unsigned)p_blocks_x, 1), 160, smem, ctx->stream>>>(a1, KP);
    CGGP_LAUNCH_CHECK(ctx);
    Args a2;
    a2.Pb = Zb; a2.Ps = Zs; a2.pn = zn; a2.np = m;
    a2.Qb = Xb; a2.Qs = Xs; a2.qn = xn; a2.nq = n;
    a2.U = T; a2.ldu = n_pad;
    a2.out = Wp; a2.ldo = m;
    a2.q_tiles_per_split = tiles_per_split;
    a2.variance = (float)variance;
    a2.active = active;
    fn<<<dim3((unsigned)p_blocks_z, (unsigned)splits), 160, smem, ctx->stream>>>(a2, KP);
    CGGP_LAUNCH_CHECK(ctx);
    reduce_splits_kernel<<<dim3((unsigned)((m + 255) / 256), (unsigned)nb), 256, 0, ctx->stream>>>(
        Wp, (int)splits, nb, m, m, W + (int64_t)b0 * ldw, ldw, active);
    CGGP_LAUNCH_CHECK(ctx);
  }
  return CGGP_OK;
}
