// Register-tier kernel instantiations, fp32 / complex64.
#include "hea_reg_inst.cuh"

namespace qon {

#define QON_F32_COMBOS(X) X(1, 0) X(2, 0) X(3, 0) X(4, 0) X(5, 0) X(5, 1) X(5, 2) X(5, 3) X(5, 4) X(5, 5)

RegLaunchInfo reg_info_f32(int nl, int lq, int mode) {
#define X(NL, LQ)                                                        \
    if (nl == NL && lq == LQ) {                                          \
        if (mode == 0) return RegK<float, NL, LQ, 0>::info();            \
        if (mode == 1) return RegK<float, NL, LQ, 1>::info();            \
        return RegK<float, NL, LQ, 2>::info();                           \
    }
    QON_F32_COMBOS(X)
#undef X
    return RegLaunchInfo{0, 0, 0, false};
}

cudaError_t reg_launch_f32(int nl, int lq, int mode, int grid, const HeaParams<float>& p, cudaStream_t st) {
#define X(NL, LQ)                                                                  \
    if (nl == NL && lq == LQ) {                                                    \
        if (mode == 0) return RegK<float, NL, LQ, 0>::launch(grid, p, st);         \
        if (mode == 1) return RegK<float, NL, LQ, 1>::launch(grid, p, st);         \
        return RegK<float, NL, LQ, 2>::launch(grid, p, st);                        \
    }
    QON_F32_COMBOS(X)
#undef X
    return cudaErrorInvalidValue;
}

}  // namespace qon
