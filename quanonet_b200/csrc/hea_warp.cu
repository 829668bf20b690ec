// Small-batch latency tier: instantiations, planning and launch (fp32 and fp64, n = 1..5, modes 0..5 as in
// hea_reg_inst.cuh).
#include "hea_dispatch.cuh"
#include "hea_warp_wide.cuh"

namespace qon {

namespace {
template <typename T>
using WarpKern = void (*)(const HeaParams<T>, const DepthPack);

template <typename T, int N>
WarpKern<T> warp_kernel_n(int mode) {
    switch (mode) {
        case 0: return hea_warp_kernel<T, N, false, false, 0, kWarpThreads>;
        case 1: return hea_warp_kernel<T, N, true, true, 0, kWarpThreads>;
        case 2: return hea_warp_kernel<T, N, true, false, 0, kWarpThreads>;
        case 3: return hea_warp_kernel<T, N, false, false, 1, kWarpThreads>;
        case 4: return hea_warp_kernel<T, N, true, false, 1, kWarpThreads>;
        case 5: return hea_warp_kernel<T, N, true, false, 2, kWarpThreads>;
        default: return nullptr;
    }
}

template <typename T, int N>
WarpKern<T> warp_wide_kernel_n(int mode) {
    switch (mode) {
        case 0: return hea_warp_wide_kernel<T, N, false, false, 0, kWarpThreads>;
        case 1: return hea_warp_wide_kernel<T, N, true, true, 0, kWarpThreads>;
        case 2: return hea_warp_wide_kernel<T, N, true, false, 0, kWarpThreads>;
        default: break;
    }
    if constexpr (sizeof(T) == 4 && N <= 9) {      // fused-encoding modes: fp32, n = 6..9
        switch (mode) {
            case 3: return hea_warp_wide_kernel<T, N, false, false, 1, kWarpThreads>;
            case 4: return hea_warp_wide_kernel<T, N, true, false, 1, kWarpThreads>;
            case 5: return hea_warp_wide_kernel<T, N, true, false, 2, kWarpThreads>;
            default: break;
        }
    }
    return nullptr;
}

template <typename T>
WarpKern<T> warp_kernel(int n, int mode) {
    // n = 6..10 (fp64: 6..9): several amplitudes per lane (hea_warp_wide.cuh), x given
    switch (n) {
        case 6: return warp_wide_kernel_n<T, 6>(mode);
        case 7: return warp_wide_kernel_n<T, 7>(mode);
        case 8: return warp_wide_kernel_n<T, 8>(mode);
        case 9: return warp_wide_kernel_n<T, 9>(mode);
        case 10:
            if constexpr (sizeof(T) == 4) return warp_wide_kernel_n<T, 10>(mode);
            else return nullptr;
        default: break;
    }
    switch (n) {
        case 1: return warp_kernel_n<T, 1>(mode);
        case 2: return warp_kernel_n<T, 2>(mode);
        case 3: return warp_kernel_n<T, 3>(mode);
        case 4: return warp_kernel_n<T, 4>(mode);
        case 5: return warp_kernel_n<T, 5>(mode);
        default: return nullptr;
    }
}

template <typename T>
WarpPlan plan_t(int n, int K, int S, int mode) {
    WarpPlan wp{};
    WarpKern<T> k = warp_kernel<T>(n, mode);
    if (!k) return wp;
    wp.threads = kWarpThreads;
    wp.smem_bytes = warp_smem_bytes<T>(n, K, S, mode == 1 || mode == 5, mode == 5, kWarpThreads);
    if (wp.smem_bytes > 200 * 1024) return wp;             // very long circuits: the throughput layout serves them
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wp.smem_bytes) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&wp.blocks_per_sm, k, kWarpThreads, wp.smem_bytes) != cudaSuccess ||
        wp.blocks_per_sm < 1) {
        cudaGetLastError();
        return wp;
    }
    wp.ok = true;
    return wp;
}

template <typename T>
cudaError_t launch_t(int n, int mode, int grid, const WarpPlan& wp, const HeaParams<T>& p, const DepthPack& dp,
                     cudaStream_t st) {
    WarpKern<T> k = warp_kernel<T>(n, mode);
    if (!k) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wp.smem_bytes);
    if (e != cudaSuccess) return e;
    k<<<grid, wp.threads, wp.smem_bytes, st>>>(p, dp);
    return cudaGetLastError();
}
}  // namespace

WarpPlan warp_plan(int n, int K, int S, int dtype_bytes, int mode) {
    return dtype_bytes == 4 ? plan_t<float>(n, K, S, mode) : plan_t<double>(n, K, S, mode);
}
cudaError_t warp_launch_f32(int n, int mode, int grid, const WarpPlan& wp, const HeaParams<float>& p,
                            const DepthPack& dp, cudaStream_t st) {
    return launch_t<float>(n, mode, grid, wp, p, dp, st);
}
cudaError_t warp_launch_f64(int n, int mode, int grid, const WarpPlan& wp, const HeaParams<double>& p,
                            const DepthPack& dp, cudaStream_t st) {
    return launch_t<double>(n, mode, grid, wp, p, dp, st);
}

}  // namespace qon
