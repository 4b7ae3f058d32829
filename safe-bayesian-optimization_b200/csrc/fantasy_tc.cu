// fantasy_tc.cu -- TF32 tcgen05/TMEM fantasy-expander GEMM (placeholder until the kernel lands).
#include "common.cuh"
struct FantasyConsts;
int fantasy_tc_run(sbo_ctx* ctx, const FantasyConsts& fc, long long nx, long long nz, long long nxp, long long nzp,
                   const float* Vx, const float* Vz, const double* aux_x, const double* aux_z, int* counts_c) {
  return sbo_fail(ctx, SBO_ERR_INVALID, "TF32 fantasy kernel not built yet");
}
