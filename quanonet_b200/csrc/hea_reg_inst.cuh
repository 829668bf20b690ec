// Instantiation helper for the register tier: included by hea_reg_f32.cu / hea_reg_f64.cu.
//
// mode: 0 forward | 1 fwd+grad with dL/dx | 2 fwd+grad without dL/dx
//       3 forward, fused encoding | 4 fwd+grad, fused encoding, fixed frequencies
//       5 fwd+grad, fused encoding + frequency-layer gradients
// Fused-encoding modes are instantiated for one-thread-per-sample layouts (LQ == 0) only.
#pragma once
#include "hea_dispatch.cuh"
#include "hea_reg.cuh"

#ifndef QON_GRAD_THREADS
#define QON_GRAD_THREADS 256
#endif
#ifndef QON_FWD_THREADS
#define QON_FWD_THREADS 128
#endif
#ifndef QON_FWD_MINB
#define QON_FWD_MINB 3
#endif

namespace qon {

template <typename T, int NL, int LQ, int MODE>
struct RegK {
    static constexpr bool GRAD = MODE == 1 || MODE == 2 || MODE == 4 || MODE == 5;
    static constexpr bool GX = MODE == 1;
    static constexpr int ENC = MODE < 3 ? 0 : (MODE == 5 ? 2 : 1);
    static constexpr int THREADS = GRAD ? QON_GRAD_THREADS : QON_FWD_THREADS;
#ifndef QON_GRAD_MINB
#define QON_GRAD_MINB (256 / QON_GRAD_THREADS)
#endif
    static constexpr int MINB = GRAD ? QON_GRAD_MINB : QON_FWD_MINB;
    static void (*kernel())(const HeaParams<T>) { return hea_reg_kernel<T, NL, LQ, GRAD, GX, ENC, THREADS, MINB>; }
    static RegLaunchInfo info() {
        RegLaunchInfo r{THREADS, 0, 0, true};
        cudaFuncAttributes a;
        if (cudaFuncGetAttributes(&a, kernel()) != cudaSuccess) { r.ok = false; return r; }
        r.regs = a.numRegs;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&r.blocks_per_sm, kernel(), THREADS, 0) != cudaSuccess ||
            r.blocks_per_sm < 1)
            r.ok = false;
        return r;
    }
    static cudaError_t launch(int grid, const HeaParams<T>& p, cudaStream_t st) {
        kernel()<<<grid, THREADS, 0, st>>>(p);
        return cudaGetLastError();
    }
};

template <typename T, int NL, int LQ>
RegLaunchInfo reg_info_t(int mode) {
    switch (mode) {
        case 0: return RegK<T, NL, LQ, 0>::info();
        case 1: return RegK<T, NL, LQ, 1>::info();
        case 2: return RegK<T, NL, LQ, 2>::info();
        default: break;
    }
    if constexpr (LQ == 0) {
        switch (mode) {
            case 3: return RegK<T, NL, LQ, 3>::info();
            case 4: return RegK<T, NL, LQ, 4>::info();
            case 5: return RegK<T, NL, LQ, 5>::info();
            default: break;
        }
    }
    return RegLaunchInfo{0, 0, 0, false};
}

template <typename T, int NL, int LQ>
cudaError_t reg_launch_t(int mode, int grid, const HeaParams<T>& p, cudaStream_t st) {
    switch (mode) {
        case 0: return RegK<T, NL, LQ, 0>::launch(grid, p, st);
        case 1: return RegK<T, NL, LQ, 1>::launch(grid, p, st);
        case 2: return RegK<T, NL, LQ, 2>::launch(grid, p, st);
        default: break;
    }
    if constexpr (LQ == 0) {
        switch (mode) {
            case 3: return RegK<T, NL, LQ, 3>::launch(grid, p, st);
            case 4: return RegK<T, NL, LQ, 4>::launch(grid, p, st);
            case 5: return RegK<T, NL, LQ, 5>::launch(grid, p, st);
            default: break;
        }
    }
    return cudaErrorInvalidValue;
}

}  // namespace qon
