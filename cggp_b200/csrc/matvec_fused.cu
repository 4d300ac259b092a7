// placeholder until the fused DMMA kernel lands
#include "common.cuh"
bool cggp_matvec_fused_supported(cggp_ctx*, int, int64_t, int, int) { return false; }
int cggp_matvec_fused(cggp_ctx* ctx, int, double, const double*, const double*, int64_t, const double*, const double*,
                      int64_t, int, int64_t, const double*, int64_t, int, double*, int64_t, const int*) {
  CGGP_FAIL(ctx, CGGP_ERR_UNSUPPORTED, "fused matvec not built");
}
