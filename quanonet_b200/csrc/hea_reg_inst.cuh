// Instantiation helper for the register tier: included by hea_reg_f32.cu / hea_reg_f64.cu with
// QON_REAL, QON_SUFFIX and QON_COMBOS(X) defined.
#pragma once
#include "hea_dispatch.cuh"
#include "hea_reg.cuh"

namespace qon {

template <typename T, int NL, int LQ, int MODE>
struct RegK {
    static constexpr bool GRAD = MODE != 0;
    static constexpr bool GX = MODE == 1;
    static constexpr int THREADS = 128;
    static constexpr int MINB = GRAD ? 2 : 3;
    static void (*kernel())(const HeaParams<T>) { return hea_reg_kernel<T, NL, LQ, GRAD, GX, THREADS, MINB>; }
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

}  // namespace qon
