// Fused matrix-free product  W = V @ (Kuf Kfu)  for UP TO EIGHT right-hand sides per sweep, float64, with BOTH tile
// contractions on the FP64 tensor cores (variant 3 of cggp_kuf_kfu_matvec for B >= 3).
//
//   t[i, b] = sum_j K_ij V[b, j]   (phase 1)          w[j, b] = sum_i K_ij t[i, b]   (phase 2)
//
// For B right-hand sides the two contractions are [rows x cols] . [cols x B] and [cols x rows] . [rows x B] products;
// with B = 8 they are exact DMMA m8n8k4 shapes, so the kernel evaluation (DMMA distance + FP64 sqrt / exp, the part
// that binds the single-RHS kernel) is amortised over 8 right-hand sides:  per 8 x 8 block of Gram
// entries  KS DMMAs (distance) + 2 DMMAs (t) + 2 DMMAs (w)  instead of 2 B FMAs + shuffles per entry.  Used by the
// multi-RHS solves of the models (cggp/models.py:308-314 probes, :340 predictive variance).
//
// Structure = matvec_pipe.cu (persistent cooperative CTAs, column groups, TMA-staged X tiles, software-pipelined
// phases, exchange through L2 in a fixed order), with these differences:
//   * a Gram block leaves phase 1 in the DMMA accumulator layout (lane (r, q): row r, columns 2q, 2q + 1) and is the
//     A operand of the t-product in exactly that layout (k index q <-> column 2q, then 2q + 1): no data movement;
//   * the block is parked in shared memory as a row-major 8 x 8 tile (one conflict-free STS.128 per lane) and read
//     back in phase 2 as the A operand of the TRANSPOSED product (lane (c, i) reads row 4 s + i, column c: one
//     conflict-free LDS.64 per k-step);
//   * the DMMA of phase 2 sums over the rows, so the result needs no cross-lane reduction at all;
//   * the partial t of a block is 8 x larger: every block is handled by EWC exchange warps that split the values,
//     two such sets take alternate blocks.
// Determinism as in matvec_pipe.cu: fixed summation orders everywhere, no atomics on data.
#include <cooperative_groups.h>

#include <cstdlib>

#include "pipe_common.cuh"

namespace kpipe8 {
using namespace kpipe;

constexpr int XS = 2;     // X-tile ring stages
constexpr int SLOTS = 4;  // exchange slot ring (see matvec_pipe.cu)
constexpr int NB = 8;     // right-hand sides per sweep (= the n of m8n8k4)
constexpr int WARPS = 16, CBW = 2;
constexpr int EWC = 2;    // exchange warps cooperating on one block
constexpr int EWA = 2;    // sets of exchange warps taking alternate blocks
constexpr int EWARPS = EWC * EWA;

struct Args {
  const double* PX;
  const double2* xa2;  // (alpha |x|^2, alpha |x|^2) per row
  const double* nX;
  int64_t n;
  const double* PZ;
  const double* nZ;
  int64_t m;
  int D;
  int64_t ldp;
  const double* V;  // [B, ldv]
  int64_t ldv;
  int B;            // 1..8 valid right-hand sides (the rest of the 8 are zero)
  double variance2;
  double* Wp;       // [G][8][m]
  double* part;     // [G][SLOTS][C][BM*8]
  int* counters;    // [G][SLOTS][EWC]
  int C, G;
  int64_t nblocks;
  int tma_ok;
  const double* Tin;  // "t given" mode (Kuf @ Y): [n, ldt], P = B columns
  int64_t ldt;
  const int* active;
  const int2* etab;
};

template <int RB, int KS, int NBUF>
struct Layout {
  static constexpr int THREADS = WARPS * 32;
  static constexpr int BM = RB * 8;
  static constexpr int BN = WARPS * CBW * 8;
  static constexpr int LDX = KS * 4;
  static constexpr int TV = BM * NB;  // values of one block's t
  static constexpr size_t kbuf_bytes = (size_t)NBUF * WARPS * RB * CBW * 64 * sizeof(double);
  static constexpr size_t xt_bytes = (size_t)XS * BM * LDX * sizeof(double);
  static constexpr size_t xa_bytes = (size_t)XS * BM * 2 * sizeof(double);
  static constexpr size_t tred_bytes = (size_t)2 * WARPS * TV * sizeof(double);  // indexed by the block's parity
  static constexpr size_t tfull_bytes = (size_t)NBUF * TV * sizeof(double);
  static constexpr size_t etab_bytes = sizeof(int2) << 10;
  static constexpr size_t bar_bytes = (XS + NBUF + 2 + 1) / 2 * 2 * sizeof(uint64_t);
  static constexpr size_t total = kbuf_bytes + xt_bytes + xa_bytes + tred_bytes + tfull_bytes + bar_bytes + etab_bytes;
};

constexpr int BAR_T = 1;  // named barriers BAR_T + (block % NBUF)

template <int KIND, int KS, int RB, int NBUF>
__global__ void __launch_bounds__((WARPS + EWARPS) * 32, 1) kfu_pipe8_kernel(const Args a) {
  if (cg_inactive(a.active)) return;
  using L = Layout<RB, KS, NBUF>;
  constexpr int LAG = NBUF - 1;
  constexpr int THREADS = L::THREADS, BM = L::BM, BN = L::BN, LDX = L::LDX, TV = L::TV;
  constexpr int ALL = THREADS + EWC * 32;  // compute warps + the exchange set of a block
  constexpr int QV = (TV + EWC * 32 - 1) / (EWC * 32);  // values per exchange lane
  const int g = blockIdx.x / a.C, rank = blockIdx.x % a.C;
  if (g >= a.G) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* kbuf = reinterpret_cast<double*>(smem_raw);
  double* xt = reinterpret_cast<double*>(smem_raw + L::kbuf_bytes);
  double* xa = reinterpret_cast<double*>(smem_raw + L::kbuf_bytes + L::xt_bytes);
  double* tred = reinterpret_cast<double*>(smem_raw + L::kbuf_bytes + L::xt_bytes + L::xa_bytes);
  double* tfull = tred + 2 * WARPS * TV;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(tfull + NBUF * TV);
  const int2* etab = reinterpret_cast<const int2*>(reinterpret_cast<unsigned char*>(mbar) + L::bar_bytes);
  {
    int2* et = const_cast<int2*>(etab);
    for (int j = tid; j < 1024; j += (WARPS + EWARPS) * 32) et[j] = a.etab[j];
  }

  const int64_t nit = a.nblocks > g ? (a.nblocks - g + a.G - 1) / a.G : 0;
  auto row0_of = [&](int64_t it) { return (g + it * (int64_t)a.G) * BM; };
  auto is_manual = [&](int64_t it) { return !a.tma_ok || row0_of(it) + BM > a.n; };

  // mbar[0..XS): X tile landed; F[par]: t of the block is in tfull[par]; R[parity]: the exchange set has read the
  // per-warp partials tred[parity] of the block (two buffers: block it + 2 may overwrite them)
  uint64_t* mbarF = mbar + XS;
  uint64_t* mbarR = mbarF + NBUF;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < XS; ++s) mbar_init(&mbar[s], 1);
#pragma unroll
    for (int s = 0; s < NBUF; ++s) mbar_init(&mbarF[s], EWC * 32);
#pragma unroll
    for (int s = 0; s < 2; ++s) mbar_init(&mbarR[s], EWC * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= WARPS) {
    // =============================== exchange warps ===============================
    const int ew = warp - WARPS;
    const int set = ew / EWC, part_id = ew % EWC;  // set takes blocks it % EWA == set; part = slice of the values
    const int elane = part_id * 32 + lane;         // 0 .. EWC*32-1
    auto stage_tile = [&](int64_t it) {            // by part 0 of the set that owns block it - XS ... see below
      if (it >= nit) return;
      const int s = (int)(it % XS);
      const int64_t r0 = row0_of(it);
      if (!is_manual(it)) {
        if (lane == 0) {
          constexpr unsigned xb = BM * LDX * sizeof(double), nb = BM * 2 * sizeof(double);
          mbar_expect_tx(&mbar[s], xb + nb);
          tma_bulk_g2s(xt + s * BM * LDX, a.PX + r0 * a.ldp, xb, &mbar[s]);
          tma_bulk_g2s(xa + s * BM * 2, a.xa2 + r0, nb, &mbar[s]);
        }
      } else {
        for (int e = lane; e < BM * LDX; e += 32) {
          const int r = e / LDX, k = e % LDX;
          double x = 0.0;
          if (r0 + r < a.n) x = (k < a.D) ? a.PX[(r0 + r) * a.ldp + k] : (k == a.D ? 1.0 : 0.0);
          xt[s * BM * LDX + e] = x;
        }
        for (int e = lane; e < BM; e += 32) {
          const double v = (r0 + e < a.n) ? Fam<KIND>::alpha * a.nX[r0 + e] : 0.0;
          xa[(s * BM + e) * 2] = xa[(s * BM + e) * 2 + 1] = v;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&mbar[s]);
      }
    };
    static_assert(XS == 2 && EWA == 2, "exchange set s refills X stage s");
    if (part_id == 0) stage_tile(set);
    __threadfence_block();
    __syncthreads();  // (S)
    double* slots_g = a.part + (int64_t)g * SLOTS * a.C * TV;
    for (int64_t it = set; it < nit; it += EWA) {
      const int par = (int)(it % NBUF), par2 = (int)(it & 1);
      bar_sync(BAR_T + par, ALL);  // every compute warp finished phase 1 of block it
      if (part_id == 0) stage_tile(it + XS);
      const double* tr = tred + (size_t)par2 * WARPS * TV;
      double* tf = tfull + par * TV;
      const int64_t r0 = row0_of(it);
      double sum[QV];
#pragma unroll
      for (int q = 0; q < QV; ++q) {
        const int e = q * (EWC * 32) + elane;
        double v = 0.0;
        if (e < TV) {
          v = tr[e];
#pragma unroll
          for (int w = 1; w < WARPS; ++w) v += tr[w * TV + e];
        }
        sum[q] = v;
      }
      mbar_arrive(&mbarR[par2]);  // tred[par2] may be overwritten (block it + 2)
      if (a.Tin) {
#pragma unroll
        for (int q = 0; q < QV; ++q) {
          const int e = q * (EWC * 32) + elane;
          if (e < TV) {
            const int64_t row = r0 + e / NB;
            const int b = e % NB;
            tf[e] = (row < a.n && b < a.B) ? a.Tin[row * a.ldt + b] * a.variance2 : 0.0;
          }
        }
        mbar_arrive(&mbarF[par]);
        continue;
      }
      if (a.C > 1) {
        const int slot = (int)(it % SLOTS);
        double* mine = slots_g + ((int64_t)slot * a.C + rank) * TV;
#pragma unroll
        for (int q = 0; q < QV; ++q) {
          const int e = q * (EWC * 32) + elane;
          if (e < TV) __stcg(&mine[e], sum[q]);
        }
        __syncwarp();
        if (lane == 0) {
          __threadfence();
          int* cnt = &a.counters[(g * SLOTS + slot) * EWC + part_id];
          atomicAdd(cnt, 1);
          const int target = a.C * (int)(it / SLOTS + 1);
          while (ld_acquire(cnt) < target) {
          }
        }
        __syncwarp();
        const double* sl = slots_g + (int64_t)slot * a.C * TV;
#pragma unroll
        for (int q = 0; q < QV; ++q) {
          const int e = q * (EWC * 32) + elane;
          if (e < TV) {
            double v = 0.0;
            for (int c0 = 0; c0 < a.C; c0 += 16) {
              double tmp[16];
#pragma unroll
              for (int u = 0; u < 16; ++u) tmp[u] = (c0 + u < a.C) ? __ldcg(&sl[(int64_t)(c0 + u) * TV + e]) : 0.0;
#pragma unroll
              for (int u = 0; u < 16; ++u) v += tmp[u];
            }
            sum[q] = v;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < QV; ++q) {
        const int e = q * (EWC * 32) + elane;
        if (e < TV) tf[e] = (r0 + e / NB < a.n) ? sum[q] * a.variance2 : 0.0;
      }
      mbar_arrive(&mbarF[par]);
    }
    return;
  }

  // =============================== compute warps ===============================
  const int lr = lane >> 2, lk = lane & 3;
  const int64_t col0 = (int64_t)rank * BN + warp * (CBW * 8);
  double bf[CBW][KS];
  double ve[CBW], vo[CBW];  // B fragments of the t-product: V[b = lr][col 2 lk] and V[b = lr][col 2 lk + 1]
#pragma unroll
  for (int cb = 0; cb < CBW; ++cb) {
    const int64_t zc = col0 + cb * 8 + lr;
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
    const int64_t vc = col0 + cb * 8 + 2 * lk;
    ve[cb] = (a.V && lr < a.B && vc < a.m) ? a.V[(int64_t)lr * a.ldv + vc] : 0.0;
    vo[cb] = (a.V && lr < a.B && vc + 1 < a.m) ? a.V[(int64_t)lr * a.ldv + vc + 1] : 0.0;
  }
  double wacc[CBW][2];  // w[col cb*8 + lr][b = 2 lk, 2 lk + 1]
#pragma unroll
  for (int cb = 0; cb < CBW; ++cb) wacc[cb][0] = wacc[cb][1] = 0.0;
  const FastExpTable tab = fast_exp_table_biased();  // unused by the shared-memory exp (kept for kval's signature)

  double* kwarp = kbuf + (size_t)warp * RB * CBW * 64;  // + par * WARPS * RB * CBW * 64
  constexpr size_t KPAR = (size_t)WARPS * RB * CBW * 64;

  auto phase2 = [&](int64_t jt) {
    const int par = (int)(jt % NBUF);
    mbar_wait(&mbarF[par], (unsigned)((jt / NBUF) & 1));
    const double* tf = tfull + par * TV;
    const double* kb = kwarp + par * KPAR;
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      // B fragments: t[row rb*8 + 4 s + lk][b = lr]
      const double t0 = tf[(rb * 8 + lk) * NB + lr];
      const double t1 = tf[(rb * 8 + 4 + lk) * NB + lr];
#pragma unroll
      for (int cb = 0; cb < CBW; ++cb) {
        const double* blk = kb + (rb * CBW + cb) * 64;
        const double a0 = blk[lk * 8 + lr];        // K[row lk][col lr]      (A of the transposed product, k = row)
        const double a1 = blk[(4 + lk) * 8 + lr];  // K[row 4 + lk][col lr]
        dmma884(wacc[cb][0], wacc[cb][1], a0, t0);
        dmma884(wacc[cb][0], wacc[cb][1], a1, t1);
      }
    }
  };

  __syncthreads();  // (S)
  for (int64_t it = 0; it < nit; ++it) {
    const int s = (int)(it % XS), par = (int)(it % NBUF), par2 = (int)(it & 1);
    mbar_wait(&mbar[s], (unsigned)((it / XS) & 1));
    if (it >= 2) mbar_wait(&mbarR[par2], (unsigned)(((it - 2) >> 1) & 1));  // partials of block it - 2 consumed
    const double* xs = xt + s * BM * LDX + lr * LDX + lk;
    const double* xas = xa + (s * BM + lr) * 2;
    double* kb = kwarp + par * KPAR;
    double* tr = tred + ((size_t)par2 * WARPS + warp) * TV;
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      double af[KS];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) af[ks] = xs[rb * 8 * LDX + ks * 4];
      const double2 xa2 = *reinterpret_cast<const double2*>(xas + rb * 8 * 2);
      double c[CBW][2];
#pragma unroll
      for (int cb = 0; cb < CBW; ++cb) {
        c[cb][0] = xa2.x, c[cb][1] = xa2.y;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) dmma884(c[cb][0], c[cb][1], af[ks], bf[cb][ks]);
      }
      double t0 = 0.0, t1 = 0.0;  // t[row lr][b = 2 lk, 2 lk + 1] over this warp's 16 columns
#pragma unroll
      for (int cb = 0; cb < CBW; ++cb) {
        const double k0 = kval<KIND, 10, 3>(c[cb][0], tab, etab);
        const double k1 = kval<KIND, 10, 3>(c[cb][1], tab, etab);
        *reinterpret_cast<double2*>(kb + (rb * CBW + cb) * 64 + lr * 8 + 2 * lk) = make_double2(k0, k1);
        dmma884(t0, t1, k0, ve[cb]);  // k index lk <-> column 2 lk
        dmma884(t0, t1, k1, vo[cb]);  //                   column 2 lk + 1
      }
      *reinterpret_cast<double2*>(tr + (rb * 8 + lr) * NB + 2 * lk) = make_double2(t0, t1);
    }
    __threadfence_block();
    bar_arrive(BAR_T + par, ALL);
    if (it >= LAG) phase2(it - LAG);
  }
  for (int64_t jt = nit > LAG ? nit - LAG : 0; jt < nit; ++jt) phase2(jt);

  // the DMMA summed over the rows: lane (lr, lk) holds w[col lr][b = 2 lk, 2 lk + 1] of this group
#pragma unroll
  for (int cb = 0; cb < CBW; ++cb) {
    const int64_t col = col0 + cb * 8 + lr;
    if (col < a.m) {
      a.Wp[((int64_t)g * NB + 2 * lk) * a.m + col] = wacc[cb][0];
      a.Wp[((int64_t)g * NB + 2 * lk + 1) * a.m + col] = wacc[cb][1];
    }
  }
}

__global__ void reduce_groups8_kernel(const double* __restrict__ Wp, int G, int B, int64_t m, double* __restrict__ W,
                                      int64_t ldw, const int* __restrict__ active) {
  if (cg_inactive(active)) return;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (c >= m || b >= B) return;
  double v = 0.0;
  for (int g = 0; g < G; ++g) v += Wp[((int64_t)g * NB + b) * m + c];
  W[(int64_t)b * ldw + c] = v;
}

struct Plan {
  const void* fn;
  int threads, BM, BN;
  size_t smem;
};
template <int KIND, int KS, int RB, int NBUF>
static Plan make_plan() {
  using L = Layout<RB, KS, NBUF>;
  static_assert(L::total <= 227 * 1024, "plan exceeds the shared memory of an SM");
  Plan p;
  p.fn = (const void*)kfu_pipe8_kernel<KIND, KS, RB, NBUF>;
  p.threads = (WARPS + EWARPS) * 32;
  p.BM = L::BM;
  p.BN = L::BN;
  p.smem = L::total;
  return p;
}
// three K buffers of 24 rows: the group exchange of a block (8 x the single-RHS volume) has two phase-1 periods
template <int KIND>
static bool plan_for_ks(int ks, Plan& p) {
  switch (ks) {
    case 1: p = make_plan<KIND, 1, 3, 3>(); return true;
    case 2: p = make_plan<KIND, 2, 3, 3>(); return true;
    case 3: p = make_plan<KIND, 3, 3, 3>(); return true;
    case 4: p = make_plan<KIND, 4, 3, 3>(); return true;
    case 5: p = make_plan<KIND, 5, 3, 3>(); return true;
    case 6: p = make_plan<KIND, 6, 3, 3>(); return true;
    case 7: p = make_plan<KIND, 7, 3, 3>(); return true;
    case 8: p = make_plan<KIND, 8, 3, 3>(); return true;
    default: return false;
  }
}
static bool plan_for(int kind, int ks, Plan& p) {
  switch (kind) {
    case CGGP_SE: return plan_for_ks<CGGP_SE>(ks, p);
    case CGGP_MATERN12: return plan_for_ks<CGGP_MATERN12>(ks, p);
    case CGGP_MATERN32: return plan_for_ks<CGGP_MATERN32>(ks, p);
    case CGGP_MATERN52: return plan_for_ks<CGGP_MATERN52>(ks, p);
    default: return false;
  }
}
}  // namespace kpipe8

// One sweep for B <= 8 right-hand sides (V != nullptr) or P <= 8 given weight columns (Tin != nullptr).
int cggp_pipe8_launch(cggp_ctx* ctx, int kind, double variance, const double* PX, const double* nX, int64_t n,
                      const double* PZ, const double* nZ, int64_t m, int D, int64_t ldp, const double* V, int64_t ldv,
                      int B, double* W, int64_t ldw, const int* active, const double* Tin, int64_t ldt) {
  using namespace kpipe8;
  if (B < 1 || B > NB) CGGP_FAIL(ctx, CGGP_ERR_INVALID, "pipe8: B=%d outside [1, 8]", B);
  const int ks = (D + 1 + 3) / 4;
  Plan p;
  if (!plan_for(kind, ks, p)) CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "pipelined matvec (8 RHS): no plan for D=%d", D);
  CGGP_CUDA(ctx, cudaFuncSetAttribute(p.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  int occ = 0;
  CGGP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, p.fn, p.threads, p.smem));
  if (occ < 1) CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "pipelined matvec (8 RHS) does not fit on an SM");
  const int grid = ctx->sm_count;
  const int C = (int)((m + p.BN - 1) / p.BN);
  if (C > grid) CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "pipelined matvec: M=%lld needs %d co-resident CTAs", (long long)m, C);
  const int G = grid / C;
  const int64_t nblocks = (n + p.BM - 1) / p.BM;
  const size_t wp_bytes = (sizeof(double) * (size_t)G * NB * (size_t)m + 15) / 16 * 16;
  const size_t part_bytes = sizeof(double) * (size_t)G * SLOTS * C * p.BM * NB;
  const size_t cnt_bytes = sizeof(int) * (size_t)G * SLOTS * EWC;
  int rc = cggp_ws_reserve(ctx, wp_bytes + part_bytes + cnt_bytes + 256);
  if (rc) return rc;
  char* base = (char*)ctx->ws;
  Args a;
  a.PX = PX; a.nX = nX; a.n = n; a.PZ = PZ; a.nZ = nZ; a.m = m; a.D = D; a.ldp = ldp;
  a.V = V; a.ldv = ldv; a.B = B;
  a.variance2 = Tin ? variance : variance * variance;
  a.Tin = Tin; a.ldt = ldt;
  a.Wp = (double*)base;
  a.part = (double*)(base + wp_bytes);
  a.counters = (int*)(base + wp_bytes + part_bytes);
  a.C = C; a.G = G; a.nblocks = nblocks; a.active = active;
  rc = dup_scaled_norms(ctx, kind, nX, n, active, &a.xa2);
  if (rc) return rc;
  rc = exp_table_device(ctx, &a.etab);
  if (rc) return rc;
  a.tma_ok = (ldp == ks * 4) && (((uintptr_t)PX) % 16 == 0) ? 1 : 0;
  CGGP_CUDA(ctx, cudaMemsetAsync(a.counters, 0, cnt_bytes, ctx->stream));
  void* kargs[] = {(void*)&a};
  CGGP_CUDA(ctx, cudaLaunchCooperativeKernel(p.fn, dim3(grid), dim3(p.threads), kargs, p.smem, ctx->stream));
  ctx->launches += 1;
  reduce_groups8_kernel<<<dim3((unsigned)((m + 255) / 256), (unsigned)B), 256, 0, ctx->stream>>>(a.Wp, G, B, m, W, ldw,
                                                                                                 active);
  CGGP_LAUNCH_CHECK(ctx);
  return CGGP_OK;
}
