// Register-tier kernel instantiations, fp32 / complex64, 2^LQ lanes per sample:
//   (5, LQ)  n = 6..10 — the lane-distributed layout for mid-size circuits (the shared-memory tier is the
//            default there; kept for A/B runs, QON_SMEM_FIRST_N).
// The one-amplitude-per-lane small-batch layout lives in hea_warp.cuh.
#include "hea_reg_inst.cuh"

namespace qon {

#define QON_F32_LANE_COMBOS(X) X(5, 1) X(5, 2) X(5, 3) X(5, 4) X(5, 5)

RegLaunchInfo reg_info_f32_lanes(int nl, int lq, int mode) {
#define X(NL, LQ) if (nl == NL && lq == LQ) return reg_info_t<float, NL, LQ>(mode);
    QON_F32_LANE_COMBOS(X)
#undef X
    return RegLaunchInfo{0, 0, 0, false};
}

cudaError_t reg_launch_f32_lanes(int nl, int lq, int mode, int grid, const HeaParams<float>& p, cudaStream_t st) {
#define X(NL, LQ) if (nl == NL && lq == LQ) return reg_launch_t<float, NL, LQ>(mode, grid, p, st);
    QON_F32_LANE_COMBOS(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace qon
