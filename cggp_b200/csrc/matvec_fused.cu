// Fused matrix-free product  W = V @ (Kuf Kfu)  in float64: every Gram entry k(x_i, z_j) is evaluated ONCE per
// application, lives in registers between the two contractions, and is never written anywhere.
//
//   t_i = sum_j K_ij v_j        (needs all M columns of row i)
//   w_j = sum_i K_ij t_i        (needs t_i complete)
//
// Layout of the work (B200: 148 SMs, FP64 pipe 64 FMA/clk/SM shared by DFMA and DMMA -- tools/microbench.cu):
//   * persistent cooperative launch, 2 CTAs of 128 threads per SM (each thread may use ~250 registers);
//   * the M inducing points are split over a GROUP of C consecutive CTAs; CTA `rank` keeps its 256-column chunk of
//     Z (pre-multiplied by -2, |z|^2 appended as an extra feature) resident in shared memory for the whole kernel;
//   * groups stride over row blocks of BM = 32 training rows; per block each warp computes a 32 x 64 sub-tile:
//       r2   = |x|^2 (accumulator init) + [x, 1] . [-2 z, |z|^2]    -> DMMA m8n8k4, K = D + 1 padded to 4
//       K    = Matern / SE of r2 with the FP64-pipe-lean sqrt / exp of kmath.cuh
//       t    += K v   (quad shuffle, cross-warp through shared memory)
//     the C partial t vectors are exchanged through L2 (release/acquire counter per group; every CTA sums the C
//     partials in the same order, so all ranks hold bit-identical t), then  w += K^T t  from the register tile;
//   * w accumulators persist in registers over all row blocks; one cross-lane reduction at the end; per-group
//     partials are summed in fixed order by a tiny second kernel (deterministic, no atomics on data).
// While one CTA of an SM waits for its group's exchange the other one computes, which hides the L2 round trip.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kmath.cuh"

namespace fused {
constexpr int WARPS = 4;
constexpr int THREADS = WARPS * 32;

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
  double* Wp;        // [G][NB][m] per-group partial results
  double* part;      // [G][2][C][BM*NB] exchanged partial t
  int* counters;     // [G]
  int C, G;
  int64_t nblocks;
  const int* active;
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__host__ __device__ constexpr int ldz_for(int KS) { return ((KS * 4) % 8 == 4) ? KS * 4 : KS * 4 + 4; }

template <int KIND, int KS, int RB, int CBW, int NB>
__global__ void __launch_bounds__(THREADS, 2) kfu_fused_kernel(const Args a) {
  if (cg_inactive(a.active)) return;
  constexpr int BM = RB * 8, WN = CBW * 8, BN = WARPS * WN, LDZ = ldz_for(KS);
  const int g = blockIdx.x / a.C, rank = blockIdx.x % a.C;
  if (g >= a.G) return;  // CTAs beyond the last full group stay idle
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lk = lane & 3;
  const int64_t col0 = (int64_t)rank * BN;

  extern __shared__ __align__(16) double sm[];
  double* zt = sm;                      // [BN][LDZ]   -2 z | |z|^2 | 0
  double* vs = zt + BN * LDZ;           // [NB][BN]
  double* tred = vs + NB * BN;          // [WARPS][BM*NB]
  double* tfull = tred + WARPS * BM * NB;  // [BM*NB]

  for (int e = tid; e < BN * LDZ; e += THREADS) {
    const int c = e / LDZ, k = e % LDZ;
    const int64_t gc = col0 + c;
    double val = 0.0;
    if (gc < a.m) {
      if (k < a.D) val = -2.0 * a.PZ[gc * a.ldp + k];
      else if (k == a.D) val = a.nZ[gc];
    }
    zt[e] = val;
  }
  for (int e = tid; e < NB * BN; e += THREADS) {
    const int b = e / BN, c = e % BN;
    const int64_t gc = col0 + c;
    vs[e] = gc < a.m ? a.V[(int64_t)b * a.ldv + gc] : 0.0;
  }
  __syncthreads();

  const FastExpTable tab = fast_exp_table();
  double wacc[CBW][2][NB];
#pragma unroll
  for (int cb = 0; cb < CBW; ++cb)
#pragma unroll
    for (int b = 0; b < NB; ++b) wacc[cb][0][b] = wacc[cb][1][b] = 0.0;

  const double* zw = zt + (warp * WN + lr) * LDZ + lk;
  const double* vw = vs + warp * WN + 2 * lk;
  int it = 0;
  for (int64_t blk = g; blk < a.nblocks; blk += a.G, ++it) {
    const int64_t row0 = blk * BM;
    double af[RB][KS], xn[RB];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      const int64_t r = row0 + rb * 8 + lr;
      const bool valid = r < a.n;
      xn[rb] = valid ? a.nX[r] : 0.0;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int k = ks * 4 + lk;
        double x = 0.0;
        if (valid) x = (k == a.D) ? 1.0 : a.PX[r * a.ldp + k];  // padding columns of PX are zero
        af[rb][ks] = x;
      }
    }
    double kf[RB][CBW][2];
    double tp[RB][NB];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
#pragma unroll
      for (int b = 0; b < NB; ++b) tp[rb][b] = 0.0;

#pragma unroll
    for (int cb = 0; cb < CBW; ++cb) {
      double bf[KS];
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) bf[ks] = zw[cb * 8 * LDZ + ks * 4];
      double2 vv[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) vv[b] = *reinterpret_cast<const double2*>(vw + b * BN + cb * 8);
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) {
        double c0 = xn[rb], c1 = xn[rb];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) dmma884(c0, c1, af[rb][ks], bf[ks]);
        const double k0 = kernel_value_fast_unit<KIND>(c0, tab);
        const double k1 = kernel_value_fast_unit<KIND>(c1, tab);
        kf[rb][cb][0] = k0;
        kf[rb][cb][1] = k1;
#pragma unroll
        for (int b = 0; b < NB; ++b) tp[rb][b] = fma(k0, vv[b].x, fma(k1, vv[b].y, tp[rb][b]));
      }
    }
    // t partials: quad reduction (the 4 lanes of a row), then across warps through shared memory
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        double v = tp[rb][b];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (lk == 0) tred[warp * BM * NB + (rb * 8 + lr) * NB + b] = v;
      }
    __syncthreads();
    const int parity = it & 1;
    double* slots = a.part + ((int64_t)(g * 2 + parity) * a.C) * (BM * NB);
    if (tid < BM * NB) {
      double s = tred[tid];
#pragma unroll
      for (int w = 1; w < WARPS; ++w) s += tred[w * BM * NB + tid];
      if (a.C > 1) {
        __stcg(&slots[(int64_t)rank * BM * NB + tid], s);
        __threadfence();
      } else {
        const int64_t r = row0 + tid / NB;
        tfull[tid] = r < a.n ? s * a.variance2 : 0.0;
      }
    }
    if (a.C > 1) {
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        atomicAdd(&a.counters[g], 1);
        const int target = a.C * (it + 1);
        while (ld_acquire(&a.counters[g]) < target) {
        }
      }
      __syncthreads();
      if (tid < BM * NB) {
        double s = 0.0;
        for (int c = 0; c < a.C; ++c) s += __ldcg(&slots[(int64_t)c * BM * NB + tid]);  // same order on every rank
        const int64_t r = row0 + tid / NB;
        tfull[tid] = r < a.n ? s * a.variance2 : 0.0;  // rows past the end contribute nothing
      }
    }
    __syncthreads();
    // w += K^T t from the register tile
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const double t = tfull[(rb * 8 + lr) * NB + b];
#pragma unroll
        for (int cb = 0; cb < CBW; ++cb) {
          wacc[cb][0][b] = fma(kf[rb][cb][0], t, wacc[cb][0][b]);
          wacc[cb][1][b] = fma(kf[rb][cb][1], t, wacc[cb][1][b]);
        }
      }
  }
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
        const int64_t col = col0 + warp * WN + cb * 8 + 2 * lk + q;
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

struct Plan {
  const void* fn;
  int BM, BN, NB;
  size_t smem;
};

template <int KIND, int KS, int RB, int CBW, int NB>
static Plan make_plan() {
  Plan p;
  p.fn = (const void*)kfu_fused_kernel<KIND, KS, RB, CBW, NB>;
  p.BM = RB * 8;
  p.BN = WARPS * CBW * 8;
  p.NB = NB;
  p.smem = sizeof(double) * ((size_t)p.BN * ldz_for(KS) + (size_t)NB * p.BN + (size_t)WARPS * p.BM * NB + p.BM * NB);
  return p;
}

template <int KIND, int KS>
static bool plan_for_nb(int nb, Plan& p) {
  switch (nb) {
    case 1: p = make_plan<KIND, KS, 4, 8, 1>(); return true;
    case 2: p = make_plan<KIND, KS, 4, 4, 2>(); return true;
    default: return false;
  }
}
template <int KIND>
static bool plan_for_ks(int ks, int nb, Plan& p) {
  switch (ks) {
    case 1: return plan_for_nb<KIND, 1>(nb, p);
    case 2: return plan_for_nb<KIND, 2>(nb, p);
    case 3: return plan_for_nb<KIND, 3>(nb, p);
    case 4: return plan_for_nb<KIND, 4>(nb, p);
    default: return false;
  }
}
static bool plan_for(int kind, int ks, int nb, Plan& p) {
  switch (kind) {
    case CGGP_SE: return plan_for_ks<CGGP_SE>(ks, nb, p);
    case CGGP_MATERN12: return plan_for_ks<CGGP_MATERN12>(ks, nb, p);
    case CGGP_MATERN32: return plan_for_ks<CGGP_MATERN32>(ks, nb, p);
    case CGGP_MATERN52: return plan_for_ks<CGGP_MATERN52>(ks, nb, p);
    default: return false;
  }
}
}  // namespace fused

bool cggp_matvec_fused_supported(cggp_ctx* ctx, int dtype, int64_t m, int D, int B) {
  if (dtype != CGGP_F64 || B < 1) return false;
  const int ks = (D + 1 + 3) / 4;
  if (ks > 4) return false;
  // the column groups must fit in one co-resident wave: C = ceil(m / 256) CTAs per group
  const int64_t C = (m + 255) / 256;
  return C <= ctx->sm_count;  // conservative (occupancy is checked again at launch)
}

int cggp_matvec_fused(cggp_ctx* ctx, int kind, double variance, const double* PX, const double* nX, int64_t n,
                      const double* PZ, const double* nZ, int64_t m, int D, int64_t ldp, const double* V, int64_t ldv,
                      int B, double* W, int64_t ldw, const int* active) {
  using namespace fused;
  const int ks = (D + 1 + 3) / 4;
  int b0 = 0;
  while (b0 < B) {
    const int nb = (B - b0) >= 2 ? 2 : 1;
    Plan p;
    if (!plan_for(kind, ks, nb, p)) CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "fused matvec: no plan for D=%d", D);
    int occ = 0;
    CGGP_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, p.fn, THREADS, p.smem));
    if (occ < 1) CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "fused matvec kernel does not fit on an SM");
    if (occ > 2) occ = 2;
    const int grid = occ * ctx->sm_count;
    const int C = (int)((m + p.BN - 1) / p.BN);
    if (C > grid) CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "fused matvec: M=%lld needs %d co-resident CTAs", (long long)m, C);
    const int G = grid / C;
    const int64_t nblocks = (n + p.BM - 1) / p.BM;
    // scratch: per-group partial W, exchange slots, counters
    const size_t wp_bytes = sizeof(double) * (size_t)G * p.NB * (size_t)m;
    const size_t part_bytes = sizeof(double) * (size_t)G * 2 * C * p.BM * p.NB;
    const size_t cnt_bytes = sizeof(int) * (size_t)G;
    int rc = cggp_ws_reserve(ctx, wp_bytes + part_bytes + cnt_bytes + 256);
    if (rc) return rc;
    char* base = (char*)ctx->ws;
    Args a;
    a.PX = PX; a.nX = nX; a.n = n; a.PZ = PZ; a.nZ = nZ; a.m = m; a.D = D; a.ldp = ldp;
    a.V = V + (int64_t)b0 * ldv; a.ldv = ldv;
    a.variance2 = variance * variance;
    a.Wp = (double*)base;
    a.part = (double*)(base + wp_bytes);
    a.counters = (int*)(base + wp_bytes + part_bytes);
    a.C = C; a.G = G; a.nblocks = nblocks; a.active = active;
    CGGP_CUDA(ctx, cudaMemsetAsync(a.counters, 0, cnt_bytes, ctx->stream));
    void* kargs[] = {(void*)&a};
    CGGP_CUDA(ctx, cudaLaunchCooperativeKernel(p.fn, dim3(grid), dim3(THREADS), kargs, p.smem, ctx->stream));
    ctx->launches += 1;
    reduce_groups_kernel<<<dim3((unsigned)((m + 255) / 256), (unsigned)p.NB), 256, 0, ctx->stream>>>(
        a.Wp, G, p.NB, m, W + (int64_t)b0 * ldw, ldw, active);
    CGGP_LAUNCH_CHECK(ctx);
    b0 += nb;
  }
  return CGGP_OK;
}
