// Register-tier kernel instantiations, fp64 / complex128.
#include "hea_reg_inst.cuh"

namespace qon {

#define QON_F64_COMBOS(X) X(1, 0) X(2, 0) X(3, 0) X(4, 0) X(4, 1) X(4, 2) X(4, 3) X(4, 4) X(4, 5)

RegLaunchInfo reg_info_f64(int nl, int lq, int mode) {
#define X(NL, LQ) if (nl == NL && lq == LQ) return reg_info_t<double, NL, LQ>(mode);
    QON_F64_COMBOS(X)
#undef X
    return RegLaunchInfo{0, 0, 0, false};
}

cudaError_t reg_launch_f64(int nl, int lq, int mode, int grid, const HeaParams<double>& p, cudaStream_t st) {
#define X(NL, LQ) if (nl == NL && lq == LQ) return reg_launch_t<double, NL, LQ>(mode, grid, p, st);
    QON_F64_COMBOS(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace qon
