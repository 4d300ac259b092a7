// Software-pipelined fused matrix-free product  W = V @ (Kuf Kfu)  in float64 (variant 3, the default).
//
//   t_i = sum_j K_ij v_j   (phase 1: needs all M columns of row i)      w_j = sum_i K_ij t_i   (phase 2)
//
// Every Gram entry k(x_i, z_j) is evaluated ONCE per application and is never written to global memory.
//
// What binds on B200 (profiles/r01_fused_v2_summary.md): the FP64 pipe (64 FMA/clk/SM, shared by DFMA and DMMA).
// The first fused kernel of round 1 (K tile in registers, 2 x 4 warps per SM; removed) kept that pipe only 45 % busy:
// 8 warps/SM cannot cover the DFMA dependency chains of sqrt / exp, and every row block ends in an exposed L2
// round trip for the exchange of the partial t.  This kernel fixes both:
//   * 1 CTA of 16 warps per SM (4 warps per scheduler, 4 independent epilogue chains each); the K values a thread
//     produced in phase 1 are parked in a thread-private slice of shared memory (conflict-free 16-byte slots,
//     196 KB per SM) instead of 128 registers, and read back by the same thread in phase 2;
//   * the M inducing points are split over a GROUP of C CTAs (256 columns each; the warp's Z fragments, its v
//     entries and its w accumulators stay in registers for the whole kernel, Z is never re-read);
//   * phases are software pipelined: P1(block i+1) runs between publishing the partial t of block i and consuming
//     the group's sum, so the L2 exchange (release/acquire counter per group) is hidden behind ~3 us of math;
//   * X row tiles (48 rows x ldp doubles + norms) are staged by TMA bulk copies (cp.async.bulk + mbarrier) two
//     blocks ahead; the scaled squared distance  a2 = alpha |x|^2 + alpha |z|^2 + beta x.z  comes straight out of
//     DMMA m8n8k4 (alpha |z|^2 rides in the spare feature column, alpha |x|^2 initialises the accumulator), with
//     alpha / beta chosen per kernel family so that no per-entry scaling multiply is left (kmath.cuh).
// Determinism: all reductions run in a fixed order (no atomics on data); every CTA of a group sums the C partials
// in the same order, so all ranks hold bit-identical t.
#include <cooperative_groups.h>

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "pipe_common.cuh"

// Timing experiments that switch parts of the product OFF (wrong results) exist only in builds with
// -DCGGP_DEBUG_KNOBS; the shipped library ignores CGGP_PIPE_DBG.
#ifdef CGGP_DEBUG_KNOBS
#define CGGP_DBG(x) (x)
#else
#define CGGP_DBG(x) false
#endif

namespace kpipe {
constexpr int XS = 2;     // X-tile ring stages
constexpr int SLOTS = 4;  // exchange slot ring.  Blocks j and j + 2 belong to the same exchange warp in every CTA, which
                          // publishes j + 2 only after it gathered j; so once a warp has gathered j + 2 (all CTAs
                          // published it) every CTA is done reading slot j, and the warp may publish j + 4 into it

struct Args {
  const double* PX;
  const double* nX;
  int64_t n;
  const double* PZ;
  const double* nZ;
  int64_t m;
  int D;
  int64_t ldp;
  const double* V;
  int64_t ldv;
  double variance2;
  double* Wp;     // [G][NB][m] per-group partial results
  double* part;   // [G][SLOTS][C][BM*NB] exchanged partial t
  int* counters;  // [G][SLOTS] one arrival counter per slot of the ring (monotonic over the launch)
  int C, G;
  int64_t nblocks;
  int tma_ok;     // PX / nX are 16-byte aligned and ldp == KS * 4: full tiles go through cp.async.bulk
  const double* Tin;  // "t given" mode (Kuf @ Y): per-row weights [n, ldt] instead of the phase-1 contraction
  int64_t ldt;
  int dbg;        // timing experiments only (env CGGP_PIPE_DBG; results are WRONG when set): 1 = skip phase 2,
                  // 2 = skip the L2 exchange
  const int* active;
  const int2* etab;  // exp table of the shared-memory variant (ET = 10: 1024 entries)
  const double2* xa2;  // DUP: (alpha |x_i|^2, alpha |x_i|^2) per row - the initial DMMA accumulator pair, one LDS.128
};

template <int WARPS, int RB, int CBW, int NB, int KS, int NBUF, int ET, int DUP = 0>
struct Layout {
  static constexpr int THREADS = WARPS * 32;
  static constexpr int BM = RB * 8;
  static constexpr int WN = CBW * 8;
  static constexpr int BN = WARPS * WN;
  static constexpr int LDX = KS * 4;
  static constexpr size_t kbuf_bytes = (size_t)NBUF * RB * CBW * THREADS * sizeof(double2);
  static constexpr size_t xt_bytes = (size_t)XS * BM * LDX * sizeof(double);
  static constexpr int XNW = DUP ? 2 : 1;  // doubles per row of the staged norms (DUP: the accumulator pair)
  static constexpr size_t xn_bytes = (size_t)XS * BM * XNW * sizeof(double);
  static constexpr size_t tred_bytes = (size_t)NBUF * WARPS * BM * NB * sizeof(double);
  static constexpr size_t tfull_bytes = (size_t)NBUF * BM * NB * sizeof(double);
  static constexpr size_t etab_bytes = ET ? ((size_t)sizeof(int2) << ET) : 0;
  static constexpr size_t bar_bytes = (XS + NBUF + 1) / 2 * 2 * sizeof(uint64_t);  // keeps the table 16-byte aligned
  static constexpr size_t total = kbuf_bytes + xt_bytes + xn_bytes + tred_bytes + tfull_bytes + bar_bytes + etab_bytes;
};

constexpr int BAR_T = 1;  // + parity of the block

template <int KIND, int KS, int WARPS, int RB, int CBW, int NB, int NBUF, int ET, int SQ, int DUP>
__global__ void __launch_bounds__(WARPS * 32 + 64, 1) kfu_pipe_kernel(const Args a) {
  if (cg_inactive(a.active)) return;
  using L = Layout<WARPS, RB, CBW, NB, KS, NBUF, ET, DUP>;
  constexpr int XNW = L::XNW;
  constexpr int LAG = NBUF - 1;  // phase 2 of block j runs after phase 1 of block j + LAG
  constexpr int THREADS = L::THREADS, BM = L::BM, WN = L::WN, BN = L::BN, LDX = L::LDX;
  constexpr int ALL = THREADS + 32;  // compute warps + the exchange warp of a block (hand-off barrier T)
  const int g = blockIdx.x / a.C, rank = blockIdx.x % a.C;
  if (g >= a.G) return;  // CTAs beyond the last full group stay idle
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* kbuf = reinterpret_cast<double2*>(smem_raw);
  double* xt = reinterpret_cast<double*>(smem_raw + L::kbuf_bytes);
  double* xn = reinterpret_cast<double*>(smem_raw + L::kbuf_bytes + L::xt_bytes);
  double* tred = reinterpret_cast<double*>(smem_raw + L::kbuf_bytes + L::xt_bytes + L::xn_bytes);  // [2][WARPS][BM*NB]
  double* tfull = tred + NBUF * WARPS * BM * NB;                                                       // [2][BM*NB]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(tfull + NBUF * BM * NB);
  const int2* etab = reinterpret_cast<const int2*>(reinterpret_cast<unsigned char*>(mbar) + L::bar_bytes);
  if constexpr (ET != 0) {
    int2* et = const_cast<int2*>(etab);
    for (int j = tid; j < (1 << ET); j += THREADS + 64) et[j] = a.etab[j];
  }

  const int64_t nit = a.nblocks > g ? (a.nblocks - g + a.G - 1) / a.G : 0;  // row blocks of this group
  auto row0_of = [&](int64_t it) { return (g + it * (int64_t)a.G) * BM; };
  auto is_manual = [&](int64_t it) { return !a.tma_ok || row0_of(it) + BM > a.n; };

  uint64_t* mbarF = mbar + XS;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < XS; ++s) mbar_init(&mbar[s], 1);
#pragma unroll
    for (int s = 0; s < NBUF; ++s) mbar_init(&mbarF[s], 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= WARPS) {
    // =============================== exchange warps ===============================
    // Stage X tiles (TMA bulk copies, two blocks ahead), reduce the 16 per-warp partial t, publish them to the
    // group through L2, wait for the other CTAs, sum the C partials in rank order and hand t to the compute
    // warps - all while those are already busy with phase 1 of the next block.  TWO such warps take alternate
    // blocks: the chain of one block (two bulk copies, fence + atomic + poll, C loads from L2: ~3 us) is longer than
    // phase 1 of a cheap block (D <= 3: ~2 us), and a single warp made it the pace of the kernel.
    const int ew = warp - WARPS;
    auto stage_tile = [&](int64_t it) {
      if (it >= nit) return;
      const int s = (int)(it % XS);
      const int64_t r0 = row0_of(it);
      if (!is_manual(it)) {
        if (lane == 0) {
          constexpr unsigned xb = BM * LDX * sizeof(double), nb = BM * XNW * sizeof(double);
          mbar_expect_tx(&mbar[s], xb + nb);
          tma_bulk_g2s(xt + s * BM * LDX, a.PX + r0 * a.ldp, xb, &mbar[s]);
          if constexpr (DUP) tma_bulk_g2s(xn + s * BM * 2, a.xa2 + r0, nb, &mbar[s]);
          else tma_bulk_g2s(xn + s * BM, a.nX + r0, nb, &mbar[s]);
        }
      } else {  // ragged last block / unaligned caller: guarded loads, completed on the same mbarrier as a TMA tile
        for (int e = lane; e < BM * LDX; e += 32) {
          const int r = e / LDX, k = e % LDX;
          double x = 0.0;
          if (r0 + r < a.n) x = (k < a.D) ? a.PX[(r0 + r) * a.ldp + k] : (k == a.D ? 1.0 : 0.0);
          xt[s * BM * LDX + e] = x;
        }
        for (int e = lane; e < BM; e += 32) {
          const double v = (r0 + e < a.n) ? a.nX[r0 + e] : 0.0;
          if constexpr (DUP) xn[(s * BM + e) * 2] = xn[(s * BM + e) * 2 + 1] = Fam<KIND>::alpha * v;
          else xn[s * BM + e] = v;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&mbar[s]);  // release: the tile written by the 32 lanes is visible to the waiters
      }
    };
    static_assert(XS == 2, "exchange warp w refills X stage w");
    stage_tile(ew);
    __threadfence_block();
    __syncthreads();  // (S) first tiles staged (manual ones visible)
    double* slots_g = a.part + (int64_t)g * SLOTS * a.C * (BM * NB);
    for (int64_t it = ew; it < nit; it += 2) {
      const int par = (int)(it % NBUF);
      bar_sync(BAR_T + par, ALL);  // every compute warp finished phase 1 of block it
      stage_tile(it + XS);         // X stage it % XS is free again
      const double* tr = tred + par * WARPS * BM * NB;
      double* tf = tfull + par * BM * NB;
      const int64_t r0 = row0_of(it);
      double sum[(BM * NB + 31) / 32];
#pragma unroll
      for (int q = 0; q < (BM * NB + 31) / 32; ++q) {
        const int e = q * 32 + lane;
        double v = 0.0;
        if (e < BM * NB) {
          v = tr[e];
#pragma unroll
          for (int w = 1; w < WARPS; ++w) v += tr[w * BM * NB + e];
        }
        sum[q] = v;
      }
      if (a.Tin) {
        // Kuf @ Y: the row weights are given, no contraction / exchange; K carries one factor `variance`
#pragma unroll
        for (int q = 0; q < (BM * NB + 31) / 32; ++q) {
          const int e = q * 32 + lane;
          if (e < BM * NB) {
            const int64_t row = r0 + e / NB;
            tf[e] = row < a.n ? a.Tin[row * a.ldt + e % NB] * a.variance2 : 0.0;
          }
        }
        mbar_arrive(&mbarF[par]);
        continue;
      }
      if (a.C > 1 && !CGGP_DBG(a.dbg & 2)) {
        double* mine = slots_g + ((int64_t)(it % SLOTS) * a.C + rank) * (BM * NB);
#pragma unroll
        for (int q = 0; q < (BM * NB + 31) / 32; ++q)
          if (q * 32 + lane < BM * NB) __stcg(&mine[q * 32 + lane], sum[q]);
        __syncwarp();
        if (lane == 0) {
          __threadfence();
          int* cnt = &a.counters[g * SLOTS + (int)(it % SLOTS)];
          atomicAdd(cnt, 1);
          // block `it` is the (it / SLOTS + 1)-th user of its slot
          const int target = a.C * (int)(it / SLOTS + 1);
          while (ld_acquire(cnt) < target) {
          }
        }
        __syncwarp();
        const double* sl = slots_g + (int64_t)(it % SLOTS) * a.C * (BM * NB);
#pragma unroll
        for (int q = 0; q < (BM * NB + 31) / 32; ++q) {
          const int e = q * 32 + lane;
          if (e < BM * NB) {
            // the C partials in rank order (same order on every CTA); 16 loads in flight at a time, so that a large
            // group (M = 16384: 64 CTAs) costs 4 L2 round trips per block instead of one per few partials
            double v = 0.0;
            for (int c0 = 0; c0 < a.C; c0 += 16) {
              double tmp[16];
#pragma unroll
              for (int u = 0; u < 16; ++u)
                tmp[u] = (c0 + u < a.C) ? __ldcg(&sl[(int64_t)(c0 + u) * BM * NB + e]) : 0.0;
#pragma unroll
              for (int u = 0; u < 16; ++u) v += tmp[u];
            }
            sum[q] = v;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < (BM * NB + 31) / 32; ++q) {
        const int e = q * 32 + lane;
        // rows past the end contribute nothing
        if (e < BM * NB) tf[e] = (r0 + e / NB < a.n) ? sum[q] * a.variance2 : 0.0;
      }
      mbar_arrive(&mbarF[par]);  // release: this lane's tfull entries (and a manually staged X tile) are visible
    }
    return;
  }

  // =============================== compute warps ===============================
  const int lr = lane >> 2, lk = lane & 3;
  const int64_t col0 = (int64_t)rank * BN + warp * WN;  // first column of this warp
  // per-warp constants held in registers for the whole kernel: Z fragments, v entries, w accumulators
  double bf[CBW][KS];
  double2 vv[CBW][NB];
#pragma unroll
  for (int cb = 0; cb < CBW; ++cb) {
    const int64_t zc = col0 + cb * 8 + lr;  // B fragment: column lr of the 8-block, features ks*4 + lk
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int k = ks * 4 + lk;
      double val = 0.0;
      if (zc < a.m) {
        if (k < a.D) val = Fam<KIND>::beta * a.PZ[zc * a.ldp + k];
        else if (k == a.D) val = Fam<KIND>::alpha * a.nZ[zc];
      }
      bf[cb][ks] = val;
    }
    const int64_t vc = col0 + cb * 8 + 2 * lk;  // C fragment: columns 2 lk, 2 lk + 1
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      vv[cb][b].x = (a.V && vc < a.m) ? a.V[(int64_t)b * a.ldv + vc] : 0.0;
      vv[cb][b].y = (a.V && vc + 1 < a.m) ? a.V[(int64_t)b * a.ldv + vc + 1] : 0.0;
    }
  }
  double wacc[CBW][2][NB];
#pragma unroll
  for (int cb = 0; cb < CBW; ++cb)
#pragma unroll
    for (int b = 0; b < NB; ++b) wacc[cb][0][b] = wacc[cb][1][b] = 0.0;
  const FastExpTable tab = fast_exp_table_biased();

  // phase 2 of block `jt`: w += K^T t from the K values this thread parked in shared memory.  (Interleaving these
  // steps into phase 1 of the next block was measured slower than running them back to back: 22.8-25.6 vs 20.5 ms.)
  auto phase2 = [&](int64_t jt) {
    const int par = (int)(jt % NBUF);
    mbar_wait(&mbarF[par], (unsigned)((jt / NBUF) & 1));  // t of block jt is complete (normally long before)
    const double* tf = tfull + par * BM * NB;
    const double2* kb = kbuf + (size_t)par * RB * CBW * THREADS + tid;
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      double t[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) t[b] = tf[(rb * 8 + lr) * NB + b];
#pragma unroll
      for (int cb = 0; cb < CBW; ++cb) {
        const double2 k = kb[(rb * CBW + cb) * THREADS];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          wacc[cb][0][b] = fma(k.x, t[b], wacc[cb][0][b]);
          wacc[cb][1][b] = fma(k.y, t[b], wacc[cb][1][b]);
        }
      }
    }
  };

  __syncthreads();  // (S)
  for (int64_t it = 0; it < nit; ++it) {
    // ------------------------------- phase 1 of block `it` -------------------------------
    const int s = (int)(it % XS), par = (int)(it % NBUF);
    mbar_wait(&mbar[s], (unsigned)((it / XS) & 1));  // X tile (TMA or guarded loads) has landed
    const double* xs = xt + s * BM * LDX + lr * LDX + lk;
    const double* xns = xn + (s * BM + lr) * XNW;
    double2* kb = kbuf + (size_t)par * RB * CBW * THREADS + tid;
    double* tr = tred + (par * WARPS + warp) * BM * NB;
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      double af[KS];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) af[ks] = xs[rb * 8 * LDX + ks * 4];
      double c[CBW][2];
      if constexpr (DUP) {
        // one LDS.128 delivers the accumulator pair (alpha |x|^2 twice): no FP64 multiply, no register moves
        const double2 xa2 = *reinterpret_cast<const double2*>(xns + rb * 8 * 2);
#pragma unroll
        for (int cb = 0; cb < CBW; ++cb) c[cb][0] = xa2.x, c[cb][1] = xa2.y;
      } else {
        const double xa = Fam<KIND>::alpha * xns[rb * 8];
#pragma unroll
        for (int cb = 0; cb < CBW; ++cb) c[cb][0] = c[cb][1] = xa;
      }
#pragma unroll
      for (int cb = 0; cb < CBW; ++cb)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) dmma884(c[cb][0], c[cb][1], af[ks], bf[cb][ks]);
      double tp[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) tp[b] = 0.0;
#pragma unroll
      for (int cb = 0; cb < CBW; ++cb) {
        const double k0 = kval<KIND, ET, SQ>(c[cb][0], tab, etab);
        const double k1 = kval<KIND, ET, SQ>(c[cb][1], tab, etab);
        kb[(rb * CBW + cb) * THREADS] = make_double2(k0, k1);
#pragma unroll
        for (int b = 0; b < NB; ++b) tp[b] = fma(k0, vv[cb][b].x, fma(k1, vv[cb][b].y, tp[b]));
      }
      // partial t of this warp's columns: the 4 lanes of a row
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        double v = tp[b];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (lk == 0) tr[(rb * 8 + lr) * NB + b] = v;
      }
    }
    __threadfence_block();
    bar_arrive(BAR_T + par, ALL);  // hand the partials (and the X stage) to the exchange warp; do not wait
    // ------------------------------- phase 2 of block `it - 1` -------------------------------
    if (it >= LAG && !CGGP_DBG(a.dbg & 1)) phase2(it - LAG);
  }
  if (!CGGP_DBG(a.dbg & 1))
    for (int64_t jt = nit > LAG ? nit - LAG : 0; jt < nit; ++jt) phase2(jt);

  // reduce the 8 row-lanes of every column, write this group's partial
#pragma unroll
  for (int cb = 0; cb < CBW; ++cb)
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        double v = wacc[cb][q][b];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        const int64_t col = col0 + cb * 8 + 2 * lk + q;
        if (lr == 0 && col < a.m) a.Wp[((int64_t)g * NB + b) * a.m + col] = v;
      }
}

__global__ void reduce_groups_kernel(const double* __restrict__ Wp, int G, int NB, int64_t m, double* __restrict__ W,
                                     int64_t ldw, const int* __restrict__ active) {
  if (cg_inactive(active)) return;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= m) return;
  double v = 0.0;
  for (int g = 0; g < G; ++g) v += Wp[((int64_t)g * NB + b) * m + c];
  W[(int64_t)b * ldw + c] = v;
}

}  // namespace kpipe

// (alpha |x_i|^2, alpha |x_i|^2): the initial DMMA accumulator pair of row i, read by the kernels with one LDS.128.
__global__ void scale_dup_kernel(const double* __restrict__ nX, int64_t n, double alpha, double2* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double v = alpha * nX[i];
    out[i] = make_double2(v, v);
  }
}
// Built per launch, or once per cggp_cg_solve (the operator's points cannot change inside a solve: the buffer is
// keyed on pointer, row count, family and the solve epoch).
int kpipe::dup_scaled_norms(cggp_ctx* ctx, int kind, const double* nX, int64_t n, const int* active,
                            const double2** out) {
  const double alpha = kind == CGGP_SE ? Fam<CGGP_SE>::alpha : kind == CGGP_MATERN12 ? Fam<CGGP_MATERN12>::alpha
                       : kind == CGGP_MATERN32 ? Fam<CGGP_MATERN32>::alpha : Fam<CGGP_MATERN52>::alpha;
  const size_t need = sizeof(double2) * (size_t)(n > 0 ? n : 1);
  const bool hit = ctx->solve_epoch > 0 && ctx->xa2_epoch == ctx->solve_epoch && ctx->xa2_key == (const void*)nX &&
                   ctx->xa2_n == n && ctx->xa2_alpha == alpha && ctx->xa2;
  if (!hit) {
    if (need > ctx->xa2_bytes) {
      CGGP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      if (ctx->xa2) cudaFree(ctx->xa2);
      ctx->xa2 = nullptr;
      ctx->xa2_bytes = 0;
      CGGP_CUDA(ctx, cudaMalloc(&ctx->xa2, need + need / 8));
      ctx->xa2_bytes = need + need / 8;
    }
    (void)active;  // built unconditionally: a later launch of the same solve may reuse it
    if (n > 0) {
      scale_dup_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(nX, n, alpha, (double2*)ctx->xa2);
      CGGP_LAUNCH_CHECK(ctx);
    }
    ctx->xa2_key = nX;
    ctx->xa2_n = n;
    ctx->xa2_alpha = alpha;
    ctx->xa2_epoch = ctx->solve_epoch;
  }
  *out = (const double2*)ctx->xa2;
  return CGGP_OK;
}

namespace kpipe {
struct Plan {
  const void* fn;
  int threads, BM, BN, NB, dup;
  size_t smem;
};

template <int KIND, int KS, int WARPS, int RB, int CBW, int NB, int NBUF, int ET = 0, int SQ = 0, int DUP = 0>
static Plan make_plan() {
  using L = Layout<WARPS, RB, CBW, NB, KS, NBUF, ET, DUP>;
  static_assert(L::total <= 227 * 1024, "plan exceeds the shared memory of an SM");
  Plan p;
  p.fn = (const void*)kfu_pipe_kernel<KIND, KS, WARPS, RB, CBW, NB, NBUF, ET, SQ, DUP>;
  p.dup = DUP;
  p.threads = L::THREADS + 64;  // + the two exchange warps
  p.BM = L::BM;
  p.BN = L::BN;
  p.NB = NB;
  p.smem = L::total;
  return p;
}

// `deep`: three K buffers of 32 rows instead of two of 48, i.e. the group exchange of a block has TWO phase-1 periods
// to complete.  Chosen where the group is large (M = 16384: 64 CTAs to wait for) or phase 1 is cheap (D <= 3).
template <int KIND, int KS>
static bool plan_for_nb(int nb, bool deep, bool wide, Plan& p) {
  // `wide` (D <= 15, one right-hand side): 32 columns per warp = 512 per CTA.  The X fragments and the cross-lane
  // reduction of t are amortised over twice the columns, and the group halves (M = 4096: 8 CTAs, M = 16384: 32): the
  // exchange of a block gathers half as many partials for the same number of Gram entries per block.
  if constexpr (KS <= 4) {
    if (nb == 1 && wide) {
      // KS <= 2: 16-row blocks, three buffers (phase 1 is cheap: the exchange gets two periods); KS = 3, 4: 24-row
      // blocks, two buffers by default (c3: 19.24 ms against 20.79 ms with 16 columns per warp)
      if (KS <= 2 || deep) p = make_plan<KIND, KS, 16, 2, 4, 1, 3, 10, 3, 1>();
      else p = make_plan<KIND, KS, 16, 3, 4, 1, 2, 10, 3, 1>();
      return true;
    }
  }
  // one right-hand side: 1024-entry shared-memory exp table (degree-3 polynomial) + third-order sqrt step + the
  // pre-scaled accumulator pair per row + early range reduction (kmath.cuh fast_matern_early); two: 32-entry shuffle table (degree 5) + two Newton steps (no room for the table)
  if (nb == 1) {
    if constexpr (KS <= 4)
      p = deep ? make_plan<KIND, KS, 16, 4, 2, 1, 3, 10, 3, 1>() : make_plan<KIND, KS, 16, 6, 2, 1, 2, 10, 3, 1>();
    else  // wider X tiles: smaller row blocks keep the kernel inside the 227 KB of shared memory
      p = deep ? make_plan<KIND, KS, 16, 3, 2, 1, 3, 10, 3, 1>() : make_plan<KIND, KS, 16, 5, 2, 1, 2, 10, 3, 1>();
    return true;
  }
  if constexpr (KS <= 4) {  // 16 <= D <= 31: two and more right-hand sides go through the 8-wide kernel (matvec_pipe8.cu)
    if (nb == 2) {
      p = deep ? make_plan<KIND, KS, 16, 3, 2, 2, 3>() : make_plan<KIND, KS, 16, 5, 2, 2, 2>();
      return true;
    }
  }
  return false;
}
template <int KIND>
static bool plan_for_ks(int ks, int nb, bool deep, bool wide, Plan& p) {
  switch (ks) {
    case 1: return plan_for_nb<KIND, 1>(nb, deep, wide, p);
    case 2: return plan_for_nb<KIND, 2>(nb, deep, wide, p);
    case 3: return plan_for_nb<KIND, 3>(nb, deep, wide, p);
    case 4: return plan_for_nb<KIND, 4>(nb, deep, wide, p);
    case 5: return plan_for_nb<KIND, 5>(nb, deep, wide, p);
    case 6: return plan_for_nb<KIND, 6>(nb, deep, wide, p);
    case 7: return plan_for_nb<KIND, 7>(nb, deep, wide, p);
    case 8: return plan_for_nb<KIND, 8>(nb, deep, wide, p);
    default: return false;
  }
}
static bool plan_for(int kind, int ks, int nb, bool deep, bool wide, Plan& p) {
  switch (kind) {
    case CGGP_SE: return plan_for_ks<CGGP_SE>(ks, nb, deep, wide, p);
    case CGGP_MATERN12: return plan_for_ks<CGGP_MATERN12>(ks, nb, deep, wide, p);
    case CGGP_MATERN32: return plan_for_ks<CGGP_MATERN32>(ks, nb, deep, wide, p);
    case CGGP_MATERN52: return plan_for_ks<CGGP_MATERN52>(ks, nb, deep, wide, p);
    default: return false;
  }
}
}  // namespace kpipe

bool cggp_matvec_pipe_supported(cggp_ctx* ctx, int dtype, int64_t m, int D, int B) {
  if (dtype != CGGP_F64 || B < 1) return false;
  const int ks = (D + 1 + 3) / 4;
  if (ks > 8) return false;  // D <= 31
  const int64_t C = (m + 255) / 256;  // one co-resident CTA per 256 columns
  return C <= ctx->sm_count;
}

static int pipe_launch(cggp_ctx* ctx, int kind, double variance, const double* PX, const double* nX, int64_t n,
                       const double* PZ, const double* nZ, int64_t m, int D, int64_t ldp, const double* V, int64_t ldv,
                       int B, double* W, int64_t ldw, const int* active, const double* Tin, int64_t ldt);
int cggp_pipe8_launch(cggp_ctx* ctx, int kind, double variance, const double* PX, const double* nX, int64_t n,
                      const double* PZ, const double* nZ, int64_t m, int D, int64_t ldp, const double* V, int64_t ldv,
                      int B, double* W, int64_t ldw, const int* active, const double* Tin, int64_t ldt);

int cggp_matvec_pipe(cggp_ctx* ctx, int kind, double variance, const double* PX, const double* nX, int64_t n,
                     const double* PZ, const double* nZ, int64_t m, int D, int64_t ldp, const double* V, int64_t ldv,
                     int B, double* W, int64_t ldw, const int* active) {
  return pipe_launch(ctx, kind, variance, PX, nX, n, PZ, nZ, m, D, ldp, V, ldv, B, W, ldw, active, nullptr, 0);
}

// W[b, :] = sum_i k(z_:, x_i) Y[i, b]  (Kuf @ Y for the local shard): the phase-2 contraction of the pipelined kernel
// with the row weights given.  V is a dummy (any [B, m] buffer of finite numbers, e.g. W itself is NOT allowed).
int cggp_kuf_times_pipe(cggp_ctx* ctx, int kind, double variance, const double* PX, const double* nX, int64_t n,
                        const double* PZ, const double* nZ, int64_t m, int D, int64_t ldp, const double* Y, int64_t ldy,
                        int P, double* W, int64_t ldw) {
  return pipe_launch(ctx, kind, variance, PX, nX, n, PZ, nZ, m, D, ldp, nullptr, 0, P, W, ldw, nullptr, Y, ldy);
}

// 2^(j / 1024), j < 1024, correctly rounded from long double on the host; the high word of entry j is pre-decremented
// by j << (20 - TBITS) (kmath.cuh).  Built once per ctx.
int kpipe::exp_table_device(cggp_ctx* ctx, const int2** out) {
  if (!ctx->exp_tab) {
    std::vector<int2> h(1024);
    auto fill = [&](int off, int bits) {
      for (int j = 0; j < (1 << bits); ++j) {
        const double t = (double)exp2l((long double)j / (long double)(1 << bits));
        int64_t u;
        memcpy(&u, &t, 8);
        h[off + j] = make_int2((int)(u & 0xffffffff), (int)(u >> 32) - (j << (20 - bits)));
      }
    };
    fill(0, 10);
    CGGP_CUDA(ctx, cudaMalloc(&ctx->exp_tab, h.size() * sizeof(int2)));
    CGGP_CUDA(ctx, cudaMemcpy(ctx->exp_tab, h.data(), h.size() * sizeof(int2), cudaMemcpyHostToDevice));
  }
  *out = (const int2*)ctx->exp_tab;
  return CGGP_OK;
}

static int pipe_launch(cggp_ctx* ctx, int kind, double variance, const double* PX, const double* nX, int64_t n,
                       const double* PZ, const double* nZ, int64_t m, int D, int64_t ldp, const double* V, int64_t ldv,
                       int B, double* W, int64_t ldw, const int* active, const double* Tin, int64_t ldt) {
  using namespace kpipe;
  const int ks = (D + 1 + 3) / 4;
  int b0 = 0;
  // from 3 right-hand sides on, sweeps of up to 8 with both contractions on DMMA (matvec_pipe8.cu): one sweep costs
  // ~1.4 x a single-RHS sweep; below that the FMA-contraction kernels of this file (1 or 2 per sweep)
  static const int p8_min = getenv("CGGP_PIPE8_MIN") ? atoi(getenv("CGGP_PIPE8_MIN")) : 3;  // tuning knob
  while (b0 < B) {
    if (B - b0 >= (ks > 4 ? 2 : p8_min)) {
      const int nb8 = (B - b0) >= 8 ? 8 : (B - b0);
      int rc8 = cggp_pipe8_launch(ctx, kind, variance, PX, nX, n, PZ, nZ, m, D, ldp, V ? V + (int64_t)b0 * ldv : nullptr,
                                  ldv, nb8, W + (int64_t)b0 * ldw, ldw, active, Tin ? Tin + b0 : nullptr, ldt);
      if (rc8) return rc8;
      b0 += nb8;
      continue;
    }
    const int nb = (B - b0) >= 2 ? 2 : 1;
    Plan p;
    static const int deep_env = getenv("CGGP_PIPE_DEEP") ? atoi(getenv("CGGP_PIPE_DEEP")) : -1;  // tuning knob
    // measured with the two exchange warps (one right-hand side, ms deep / not): c2 (D = 3) 1.56 / 1.65, c4 shape
    // (M = 16384) wins big; c3 (D = 11) 20.39 / 20.40 at N = 2M but 2.600 / 2.562 at an 8-GPU shard of 250k rows:
    // three buffers where phase 1 is cheap (one DMMA k-step) or the group is large, two otherwise
    const bool deep = deep_env >= 0 ? deep_env != 0 : ((nb == 1 && ks <= 1) || m > 8192);
    const int et = nb == 1 ? 10 : 0;
    static const int wide_env = getenv("CGGP_PIPE_WIDE") ? atoi(getenv("CGGP_PIPE_WIDE")) : -1;  // tuning knob
    // default: the wide plan where it keeps at least as many SMs busy as 256 columns per CTA and the problem is not
    // tiny (measured: c2, M = 2048: 37 groups of 4 CTAs = 148 SMs, 1.345 vs 1.586 ms; c4 share, M = 16384: 33.0 vs
    // 40.6 ms; c1, M = 500: 0.062 vs 0.055 ms - stays narrow)
    const int64_t c_narrow = (m + 255) / 256, c_wide = (m + 511) / 512;
    const int64_t use_narrow = c_narrow <= ctx->sm_count ? (ctx->sm_count / c_narrow) * c_narrow : 0;
    const int64_t use_wide = c_wide <= ctx->sm_count ? (ctx->sm_count / c_wide) * c_wide : 0;
    const bool wide = wide_env >= 0 ? wide_env != 0 : (ks <= 4 && m >= 1024 && use_wide >= use_narrow);
    if (!plan_for(kind, ks, nb, deep, wide, p)) CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "pipelined matvec: no plan for D=%d", D);
    CGGP_CUDA(ctx, cudaFuncSetAttribute(p.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    int occ = 0;
    CGGP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, p.fn, p.threads, p.smem));
    if (occ < 1) CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "pipelined matvec kernel does not fit on an SM");
    const int grid = ctx->sm_count;  // one persistent CTA per SM
    const int C = (int)((m + p.BN - 1) / p.BN);
    if (C > grid)
      CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "pipelined matvec: M=%lld needs %d co-resident CTAs", (long long)m, C);
    const int G = grid / C;
    const int64_t nblocks = (n + p.BM - 1) / p.BM;
    const size_t wp_bytes = sizeof(double) * (size_t)G * p.NB * (size_t)m;
    const size_t part_bytes = sizeof(double) * (size_t)G * SLOTS * C * p.BM * p.NB;
    const size_t cnt_bytes = sizeof(int) * (size_t)G * SLOTS;
    int rc = cggp_ws_reserve(ctx, wp_bytes + part_bytes + cnt_bytes + 256);
    if (rc) return rc;
    char* base = (char*)ctx->ws;
    Args a;
    a.PX = PX; a.nX = nX; a.n = n; a.PZ = PZ; a.nZ = nZ; a.m = m; a.D = D; a.ldp = ldp;
    a.V = V ? V + (int64_t)b0 * ldv : nullptr; a.ldv = ldv;
    a.variance2 = Tin ? variance : variance * variance;
    a.Tin = Tin ? Tin + b0 : nullptr;
    a.ldt = ldt;
    a.Wp = (double*)base;
    a.part = (double*)(base + wp_bytes);
    a.counters = (int*)(base + wp_bytes + part_bytes);
    a.C = C; a.G = G; a.nblocks = nblocks; a.active = active;
#ifdef CGGP_DEBUG_KNOBS
    static const int dbg_env = getenv("CGGP_PIPE_DBG") ? atoi(getenv("CGGP_PIPE_DBG")) : 0;  // timing experiments
    a.dbg = dbg_env;
#else
    a.dbg = 0;
#endif
    a.xa2 = nullptr;
    if (p.dup) {
      rc = dup_scaled_norms(ctx, kind, nX, n, active, &a.xa2);
      if (rc) return rc;
    }
    a.etab = nullptr;
    if (et == 10) {
      rc = exp_table_device(ctx, &a.etab);
      if (rc) return rc;
    }
    a.tma_ok = (ldp == ks * 4) && (((uintptr_t)PX | (uintptr_t)nX) % 16 == 0) ? 1 : 0;
    CGGP_CUDA(ctx, cudaMemsetAsync(a.counters, 0, cnt_bytes, ctx->stream));
    void* kargs[] = {(void*)&a};
    CGGP_CUDA(ctx, cudaLaunchCooperativeKernel(p.fn, dim3(grid), dim3(p.threads), kargs, p.smem, ctx->stream));
    ctx->launches += 1;
    reduce_groups_kernel<<<dim3((unsigned)((m + 255) / 256), (unsigned)p.NB), 256, 0, ctx->stream>>>(
        a.Wp, G, p.NB, m, W + (int64_t)b0 * ldw, ldw, active);
    CGGP_LAUNCH_CHECK(ctx);
    b0 += nb;
  }
  return CGGP_OK;
}
