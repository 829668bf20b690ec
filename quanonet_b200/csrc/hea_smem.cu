// fp32 shared-memory tier: planning and launch.
#include "hea_dispatch.cuh"
#include "hea_smem.cuh"

namespace qon {

namespace {
constexpr int kThreads = 256;
using Kern = void (*)(const HeaParams<float>, const SmemGeom);

template <int GL>
Kern pick_gl(int mode) {
    if (mode == 0) return hea_smem_kernel<false, false, GL, kThreads>;
    if (mode == 1) return hea_smem_kernel<true, true, GL, kThreads>;
    return hea_smem_kernel<true, false, GL, kThreads>;
}

Kern pick(int mode, int gl) {
    switch (gl) {
        case 1: return pick_gl<1>(mode);
        case 2: return pick_gl<2>(mode);
        case 3: return pick_gl<3>(mode);
        case 4: return pick_gl<4>(mode);
        default: return pick_gl<5>(mode);
    }
}
}  // namespace

SmemPlan smem_plan(int n, int mode) {
    SmemPlan sp{};
    sp.ok = false;
    if (n < kSmemMinN || n > kSmemMaxN || mode < 0 || mode > 2) return sp;
    SmemGeom& g = sp.geo;
    g.n = n;
    g.P = smem_passes(n);
    for (int p = 0; p < kSmemMaxP; ++p) {
        g.lo[p] = p < g.P ? smem_lo(n, p) : 0;
        g.gm[p] = 0;
    }
    g.tps_log2 = n - kSmemW;
    g.spc = kThreads >> g.tps_log2;
    g.region_bytes = 8 << n;
    g.vp = (3 * n + 3) / 4 * 4;
    sp.threads = kThreads;
    const size_t regions = (size_t)g.spc * g.region_bytes * (mode ? 2 : 1);
    sp.smem_bytes = regions + g.region_bytes;        // + one region of slack to align the base to the region size
    Kern k = pick(mode, smem_last_gates(n));
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp.smem_bytes) != cudaSuccess) {
        cudaGetLastError();
        return sp;
    }
    int bps = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, kThreads, sp.smem_bytes) != cudaSuccess || bps < 1) {
        cudaGetLastError();
        return sp;
    }
    sp.blocks_per_sm = bps;
    sp.ok = true;
    return sp;
}

cudaError_t smem_launch(int mode, int grid, const SmemPlan& sp, const HeaParams<float>& p, cudaStream_t st) {
    Kern k = pick(mode, smem_last_gates(sp.geo.n));
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp.smem_bytes);
    if (e != cudaSuccess) return e;
    k<<<grid, sp.threads, sp.smem_bytes, st>>>(p, sp.geo);
    return cudaGetLastError();
}

}  // namespace qon
