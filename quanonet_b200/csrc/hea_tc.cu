// Tensor-core tier (hea_tc.cuh): launch glue.
#include "hea_dispatch.cuh"
#include "hea_tc.cuh"

namespace qon {

size_t tc_workspace_bytes(int K) { return (size_t)K * kTcImgBytes + 256; }

template <int ENC, bool DBG>
static cudaError_t tc_fwd_launch_t(int grid, const HeaParams<float>& p, const unsigned char* bimg, float* dbg, int* err,
                                   cudaStream_t st) {
    auto kern = hea_tc_fwd_kernel<ENC, DBG>;
    const int smem = kTcStages * kTcImgBytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kTcThreads, smem, st>>>(p, bimg, dbg, err);
    return cudaGetLastError();
}

cudaError_t tc_forward_launch(int mode, int sms, const HeaParams<float>& p, const float* w, const DepthPack& dp,
                              char* tc_ws, float* dbg, int* err_user, cudaStream_t st) {
    unsigned char* bimg = reinterpret_cast<unsigned char*>(tc_ws);
    int* err = err_user ? err_user : reinterpret_cast<int*>(tc_ws + (size_t)p.K * kTcImgBytes);
    if (!err_user) {
        cudaError_t e0 = cudaMemsetAsync(err, 0, sizeof(int), st);
        if (e0 != cudaSuccess) return e0;
    }
    tc_prep_kernel<<<p.K, 32, 0, st>>>(w, p.K, dp, bimg);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int64_t ntiles = (p.B + 127) / 128;
    int64_t grid = (ntiles + 3) / 4;
    if (grid > sms) grid = sms;
    if (grid < 1) grid = 1;
    if (mode == 0) return dbg ? tc_fwd_launch_t<0, true>((int)grid, p, bimg, dbg, err, st)
                              : tc_fwd_launch_t<0, false>((int)grid, p, bimg, dbg, err, st);
    if (mode == 3) return dbg ? tc_fwd_launch_t<1, true>((int)grid, p, bimg, dbg, err, st)
                              : tc_fwd_launch_t<1, false>((int)grid, p, bimg, dbg, err, st);
    return cudaErrorInvalidValue;
}

}  // namespace qon
