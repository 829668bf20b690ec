// Register-tier kernel instantiations, fp32 / complex64, one thread per sample (n <= 5).
#include "hea_reg_inst.cuh"

namespace qon {

RegLaunchInfo reg_info_f32_lanes(int nl, int lq, int mode);
cudaError_t reg_launch_f32_lanes(int nl, int lq, int mode, int grid, const HeaParams<float>& p, cudaStream_t st);

#define QON_F32_COMBOS(X) X(1) X(2) X(3) X(4) X(5)

RegLaunchInfo reg_info_f32(int nl, int lq, int mode) {
    if (lq != 0) return reg_info_f32_lanes(nl, lq, mode);
#define X(NL) if (nl == NL) return reg_info_t<float, NL, 0>(mode);
    QON_F32_COMBOS(X)
#undef X
    return RegLaunchInfo{0, 0, 0, false};
}

cudaError_t reg_launch_f32(int nl, int lq, int mode, int grid, const HeaParams<float>& p, cudaStream_t st) {
    if (lq != 0) return reg_launch_f32_lanes(nl, lq, mode, grid, p, st);
#define X(NL) if (nl == NL) return reg_launch_t<float, NL, 0>(mode, grid, p, st);
    QON_F32_COMBOS(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace qon
