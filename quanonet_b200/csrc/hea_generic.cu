// Generic (shared-memory / HBM-streamed) tier instantiations.
#include "hea_dispatch.cuh"
#include "hea_generic.cuh"

namespace qon {

namespace {
constexpr size_t kSmemBudget = 200 * 1024;   // dynamic smem per CTA we are willing to ask for (<= 227 KB)

template <typename T, int MODE, bool SG>
void (*gkernel())(const HeaParams<T>, int, int, T*) {
    return hea_generic_kernel<T, MODE != 0, MODE == 1, SG>;
}

template <typename T>
void (*pick(int mode, bool sg))(const HeaParams<T>, int, int, T*) {
    if (sg) return mode == 0 ? gkernel<T, 0, true>() : mode == 1 ? gkernel<T, 1, true>() : gkernel<T, 2, true>();
    return mode == 0 ? gkernel<T, 0, false>() : mode == 1 ? gkernel<T, 1, false>() : gkernel<T, 2, false>();
}

template <typename T>
GenericPlan plan_t(int n, int mode) {
    GenericPlan gp{};
    const size_t N = (size_t)1 << n;
    const size_t state = (mode ? 4 : 2) * N * sizeof(T);
    gp.state_global = state > kSmemBudget;
    gp.smem_bytes = gp.state_global ? 0 : state;
    gp.threads = N >= 2048 ? 512 : (N >= 512 ? 256 : 128);
    if (gp.state_global) gp.threads = 512;       // (the kernel is built for at most 512 threads: two-gate passes hold 8 amplitudes per thread)
    auto k = pick<T>(mode, gp.state_global);
    if (gp.smem_bytes > 48 * 1024)
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gp.smem_bytes);
    int bps = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, gp.threads, gp.smem_bytes) != cudaSuccess) bps = 0;
    gp.blocks_per_sm = bps;
    return gp;
}

template <typename T>
cudaError_t launch_t(int n, int mode, int grid, const GenericPlan& gp, const HeaParams<T>& p, int vp, T* gstate,
                     cudaStream_t st) {
    auto k = pick<T>(mode, gp.state_global);
    if (gp.smem_bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gp.smem_bytes);
        if (e != cudaSuccess) return e;
    }
    k<<<grid, gp.threads, gp.smem_bytes, st>>>(p, n, vp, gstate);
    return cudaGetLastError();
}
}  // namespace

GenericPlan generic_plan(int n, int dtype, int mode) {
    return dtype == 0 ? plan_t<float>(n, mode) : plan_t<double>(n, mode);
}
cudaError_t generic_launch_f32(int n, int mode, int grid, const GenericPlan& gp, const HeaParams<float>& p, int vp,
                               float* gstate, cudaStream_t st) {
    return launch_t<float>(n, mode, grid, gp, p, vp, gstate, st);
}
cudaError_t generic_launch_f64(int n, int mode, int grid, const GenericPlan& gp, const HeaParams<double>& p, int vp,
                               double* gstate, cudaStream_t st) {
    return launch_t<double>(n, mode, grid, gp, p, vp, gstate, st);
}

}  // namespace qon
