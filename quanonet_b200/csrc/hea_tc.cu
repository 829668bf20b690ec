// Tensor-core tier (hea_tc.cuh, hea_tc2.cuh): launch glue.
#include <cstdlib>

#include "hea_dispatch.cuh"
#include "hea_tc2.cuh"

namespace qon {

size_t tc_workspace_bytes(int K, int S, int64_t B) {
    // operand images + error flag + (split training step) one 256-byte state row per sample
    return (size_t)(K + S) * kTcImgBytes + 256 + (size_t)(B > 0 ? B : 0) * 256;
}

template <bool GRAD, bool GX, int ENC, bool DBG, bool SPLIT = false>
static cudaError_t tc_launch_t(int grid, const HeaParams<float>& p, const unsigned char* img, float* dbg, int* err,
                               float* state, cudaStream_t st) {
    using G = TcGeom<GRAD>;
    auto kern = hea_tc_kernel<GRAD, GX, ENC, DBG, SPLIT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
    if (e != cudaSuccess) return e;
    static const int flags = [] { const char* e = getenv("QON_TC_FLAGS"); return e ? atoi(e) : 0; }();
    kern<<<grid, G::THREADS, G::SMEM, st>>>(p, img, dbg, err, state, flags);
    return cudaGetLastError();
}

// mode: hea_reg_inst.cuh (0 fwd | 1 grad + dL/dx | 2 grad | 3 fwd, fused encoding | 4 grad, fused encoding |
// 5 grad, fused encoding + frequency-layer gradients).  version 3 = the training step as ONE kernel, kept for A/B runs
// (default: split into a forward-only and a reverse-only kernel).
cudaError_t tc_launch(int mode, int version, int sms, const HeaParams<float>& p, const float* w, const DepthPack& dp,
                      char* tc_ws, float* dbg, int* err_user, cudaStream_t st) {
    unsigned char* img = reinterpret_cast<unsigned char*>(tc_ws);
    const bool grad = mode == 1 || mode == 2 || mode == 4 || mode == 5;
    int* err = err_user ? err_user : reinterpret_cast<int*>(tc_ws + (size_t)(p.K + p.S) * kTcImgBytes);
    if (!err_user) {
        cudaError_t e0 = cudaMemsetAsync(err, 0, sizeof(int), st);
        if (e0 != cudaSuccess) return e0;
    }
    tc_prep_kernel<<<p.K, 32, 0, st>>>(w, p.K, dp, img);
    if (grad) tc_prep_rev_kernel<<<p.S, 32, 0, st>>>(w, p.K, p.S, dp, img + (size_t)p.K * kTcImgBytes);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int64_t ntiles = (p.B + 127) / 128;
    auto grid_for = [&](int nt) {
        int64_t grid = (ntiles + nt - 1) / nt;
        if (grid > sms) grid = sms;
        return (int)(grid < 1 ? 1 : grid);
    };
    float* state = reinterpret_cast<float*>(tc_ws + (size_t)(p.K + p.S) * kTcImgBytes + 256);
    if (!grad) {
        const int g = grid_for(4);
        if (mode == 0) return dbg ? tc_launch_t<false, false, 0, true>(g, p, img, dbg, err, nullptr, st)
                                  : tc_launch_t<false, false, 0, false>(g, p, img, dbg, err, nullptr, st);
        return tc_launch_t<false, false, 1, false>(g, p, img, dbg, err, nullptr, st);
    }
    const int g = grid_for(2);
    if (version == 3 || dbg) {      // the whole step in ONE kernel (forward sweep on the gradient kernel's 2 tiles)
        switch (mode) {
            case 1: return dbg ? tc_launch_t<true, true, 0, true>(g, p, img, dbg, err, nullptr, st)
                               : tc_launch_t<true, true, 0, false>(g, p, img, dbg, err, nullptr, st);
            case 2: return tc_launch_t<true, false, 0, false>(g, p, img, dbg, err, nullptr, st);
            case 4: return tc_launch_t<true, false, 1, false>(g, p, img, dbg, err, nullptr, st);
            case 5: return tc_launch_t<true, false, 2, false>(g, p, img, dbg, err, nullptr, st);
            default: return cudaErrorInvalidValue;
        }
    }
    // split step: forward-only kernel (4 tiles / 16 warps) leaves the final states, the gradient kernel sweeps back
    e = (mode == 1 || mode == 2) ? tc_launch_t<false, false, 0, false>(grid_for(4), p, img, nullptr, err, state, st)
                                 : tc_launch_t<false, false, 1, false>(grid_for(4), p, img, nullptr, err, state, st);
    if (e != cudaSuccess) return e;
    switch (mode) {
        case 1: return tc_launch_t<true, true, 0, false, true>(g, p, img, nullptr, err, state, st);
        case 2: return tc_launch_t<true, false, 0, false, true>(g, p, img, nullptr, err, state, st);
        case 4: return tc_launch_t<true, false, 1, false, true>(g, p, img, nullptr, err, state, st);
        case 5: return tc_launch_t<true, false, 2, false, true>(g, p, img, nullptr, err, state, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace qon
