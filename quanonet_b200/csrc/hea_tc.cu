// Tensor-core tier (hea_tc.cuh, hea_tc2.cuh): launch glue.
#include "hea_dispatch.cuh"
#include "hea_tc2.cuh"

namespace qon {

size_t tc_workspace_bytes(int K, int S) { return (size_t)(K + S) * kTcImgBytes + 256; }

template <bool GRAD, bool GX, int ENC, bool DBG>
static cudaError_t tc_launch_t(int grid, const HeaParams<float>& p, const unsigned char* img, float* dbg, int* err,
                               cudaStream_t st) {
    using G = TcGeom<GRAD>;
    auto kern = hea_tc_kernel<GRAD, GX, ENC, DBG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
    if (e != cudaSuccess) return e;
    kern<<<grid, G::THREADS, G::SMEM, st>>>(p, img, dbg, err);
    return cudaGetLastError();
}

// mode: hea_reg_inst.cuh (0 fwd | 1 grad + dL/dx | 2 grad | 3 fwd, fused encoding | 4 grad, fused encoding |
// 5 grad, fused encoding + frequency-layer gradients).  version 1 = the first forward kernel (hea_tc.cuh), kept for A/B.
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
    const int nt = grad ? 2 : 4;
    int64_t grid = (ntiles + nt - 1) / nt;
    if (grid > sms) grid = sms;
    if (grid < 1) grid = 1;
    const int g = (int)grid;
    if (version == 1 && !grad) {
        const int smem = kTcStages * kTcImgBytes;
        if (mode == 0) {
            auto k1 = hea_tc_fwd_kernel<0, false>;
            if ((e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
            k1<<<g, kTcThreads, smem, st>>>(p, img, nullptr, err);
        } else {
            auto k1 = hea_tc_fwd_kernel<1, false>;
            if ((e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
            k1<<<g, kTcThreads, smem, st>>>(p, img, nullptr, err);
        }
        return cudaGetLastError();
    }
    switch (mode) {
        case 0: return dbg ? tc_launch_t<false, false, 0, true>(g, p, img, dbg, err, st)
                           : tc_launch_t<false, false, 0, false>(g, p, img, dbg, err, st);
        case 3: return tc_launch_t<false, false, 1, false>(g, p, img, dbg, err, st);
        case 1: return dbg ? tc_launch_t<true, true, 0, true>(g, p, img, dbg, err, st)
                           : tc_launch_t<true, true, 0, false>(g, p, img, dbg, err, st);
        case 2: return tc_launch_t<true, false, 0, false>(g, p, img, dbg, err, st);
        case 4: return tc_launch_t<true, false, 1, false>(g, p, img, dbg, err, st);
        case 5: return tc_launch_t<true, false, 2, false>(g, p, img, dbg, err, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace qon
