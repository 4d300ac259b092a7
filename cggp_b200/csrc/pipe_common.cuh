// Pieces shared by the pipelined float64 kernels (matvec_pipe.cu, matvec_pipe8.cu, gram.cu): DMMA / mbarrier / TMA
// wrappers, the per-family scaling of the DMMA distance and the FP64-lean kernel value.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "kmath.cuh"

namespace kpipe {

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
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// TMA 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// a2 = alpha (|x|^2 + |z|^2) + beta x.z : the argument the kernel family wants, produced directly by the DMMA
template <int KIND>
struct Fam;
template <>
struct Fam<CGGP_SE> {  // K = exp(-r2 / 2): the DMMA delivers the exponent itself
  static constexpr double alpha = -0.5, beta = 1.0;
};
template <>
struct Fam<CGGP_MATERN12> {  // a = r
  static constexpr double alpha = 1.0, beta = -2.0;
  static constexpr double clampv = 1e-36;
  static constexpr int clamp_hi = 0x38754484;
};
template <>
struct Fam<CGGP_MATERN32> {  // a = sqrt(3) r
  static constexpr double alpha = 3.0, beta = -6.0;
  static constexpr double clampv = 3e-36;
  static constexpr int clamp_hi = 0x388fe6c6;
};
template <>
struct Fam<CGGP_MATERN52> {  // a = sqrt(5) r
  static constexpr double alpha = 5.0, beta = -10.0;
  static constexpr double clampv = 5e-36;
  static constexpr int clamp_hi = 0x389a95a5;
};

// Unit-variance kernel value from the scaled argument q.  FP64-pipe instructions: SE 9, Matern-1/2 14, 3/2 16,
// 5/2 17.  Range handling costs two integer min / max on the high word (ALU pipe), no FP64 compare / select:
//   Matern: hi(q) -> clamp to [hi(alpha 1e-36), hi(708^2)] as SIGNED ints: negative q (rounding noise at x == z) and
//           q below GPflow's max(r2, 1e-36) land on the lower clamp, q beyond (708 lengthscales)^2 on the upper one
//           (exp(-708) = 3e-308 instead of an underflowed 0: absolute error 3e-308);
//   SE:     hi(q) -> min with hi(-708) as UNSIGNED ints (more negative = larger).
template <int KIND, int ET, int SQ = 0>
__device__ __forceinline__ double kval(double q, const FastExpTable& tab, const int2* etab) {
  if constexpr (KIND == CGGP_SE) {
    const unsigned h = min((unsigned)__double2hiint(q), 0xC0862000u);
    const double x = __hiloint2double((int)h, __double2loint(q));
    if constexpr (ET == 0) return fast_exp_core(x, tab);
    else return fast_exp_core_smem<ET>(x, etab);
  } else {
    // in place on the register pair (the compiler otherwise copies the low word to a fresh pair)
    double qc = q;
    asm("{\n"
        ".reg .b32 lo, hi;\n"
        "mov.b64 {lo, hi}, %0;\n"
        "max.s32 hi, hi, %1;\n"
        "min.s32 hi, hi, 0x411e9840;\n"
        "mov.b64 %0, {lo, hi};\n"
        "}\n"
        : "+d"(qc)
        : "n"(Fam<KIND>::clamp_hi));
    double a, e;
    // SQ: 0 = rsqrt seed + two Newton steps, 1 = one third-order step, 2 = 1 with the exp range reduction started
    // from the seed-accurate sqrt, 3 = 2 with the Matern polynomial folded into the table value (the shipped form:
    // 19.23 -> 18.80 -> 18.50 ms at c3 for 1 -> 2 -> 3; same FP64 instruction count, shorter dependent chain)
    if constexpr (SQ == 3 && ET != 0) {
      return fast_matern_early<ET, KIND == CGGP_MATERN12 ? 0 : (KIND == CGGP_MATERN32 ? 1 : 2)>(qc, etab);
    } else if constexpr (SQ == 2 && ET != 0) {
      e = fast_sqrt_exp_neg_early<ET>(qc, etab, a);  // range reduction started from the seed-accurate sqrt
    } else {
      a = SQ == 1 ? fast_sqrt_pos_cubic(qc) : fast_sqrt_pos_lean(qc);
      if constexpr (ET == 0) e = fast_exp_neg_core(a, tab);
      else e = fast_exp_neg_core_smem<ET>(a, etab);
    }
    if constexpr (KIND == CGGP_MATERN12) {
      return e;
    } else if constexpr (KIND == CGGP_MATERN32) {
      return (1.0 + a) * e;
    } else {
      return fma(qc, 1.0 / 3.0, 1.0 + a) * e;  // 1 + sqrt5 r + 5/3 r^2, with a^2 = 5 r2 (= qc up to one rounding)
    }
  }
}

// Hand-offs inside a CTA (matvec_pipe.cu).  T (named barrier 1 + parity): "the partial t of block j is in tred[j & 1] and X stage
// j % XS is free" - the compute warps arrive without waiting, the exchange warp waits.  F (mbarrier per parity):
// "t of block j is in tfull[j & 1]" - the 32 exchange lanes arrive, every compute warp waits on its own, so the
// compute warps are never synchronised with each other and drift apart by up to a phase.
__device__ __forceinline__ void bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 2^(j / 1024), j < 1024 (biased high words, kmath.cuh), built once per ctx on the host (matvec_pipe.cu)
int exp_table_device(cggp_ctx* ctx, const int2** out);
// (alpha |x_i|^2, alpha |x_i|^2) per row in a ctx-owned buffer (matvec_pipe.cu)
int dup_scaled_norms(cggp_ctx* ctx, int kind, const double* nX, int64_t n, const int* active, const double2** out);

}  // namespace kpipe
