// Internal launch interface between the C-ABI translation unit (qon_capi.cu) and the kernel
// instantiation units (hea_reg_f32.cu, hea_reg_f64.cu, hea_generic.cu).
#pragma once
#include <cuda_runtime.h>
#include "hea_common.cuh"

namespace qon {

struct RegLaunchInfo {
    int threads;         // CTA size
    int blocks_per_sm;   // resident CTAs per SM (occupancy API)
    int regs;            // registers per thread
    bool ok;             // the (NL, LQ) combination is instantiated
};

// mode: 0 = forward only, 1 = forward+backward with dL/dx, 2 = forward+backward without dL/dx
RegLaunchInfo reg_info_f32(int nl, int lq, int mode);
RegLaunchInfo reg_info_f64(int nl, int lq, int mode);
cudaError_t reg_launch_f32(int nl, int lq, int mode, int grid, const HeaParams<float>& p, cudaStream_t st);
cudaError_t reg_launch_f64(int nl, int lq, int mode, int grid, const HeaParams<double>& p, cudaStream_t st);

struct GenericPlan {
    int threads;
    size_t smem_bytes;       // dynamic shared memory (0 when the state lives in HBM)
    bool state_global;
    int blocks_per_sm;
};
GenericPlan generic_plan(int n, int dtype, int mode);
cudaError_t generic_launch_f32(int n, int mode, int grid, const GenericPlan& gp, const HeaParams<float>& p, int vp,
                               float* gstate, cudaStream_t st);
cudaError_t generic_launch_f64(int n, int mode, int grid, const GenericPlan& gp, const HeaParams<double>& p, int vp,
                               double* gstate, cudaStream_t st);

// fp32 shared-memory tier (hea_smem.cu): n in [kSmemMinN, kSmemMaxN], modes 0 / 1 / 2
struct SmemPlan {
    bool ok;
    int threads, blocks_per_sm;
    size_t smem_bytes;
    SmemGeom geo;
};
SmemPlan smem_plan(int n, int mode);
cudaError_t smem_launch(int mode, int grid, const SmemPlan& sp, const HeaParams<float>& p, cudaStream_t st);

// small-batch latency tier (hea_warp.cu): one amplitude per lane, n in [1, 5], fp32 and fp64, modes 0 / 1 / 2
struct WarpPlan {
    bool ok;
    int threads, blocks_per_sm;
    size_t smem_bytes;
};
WarpPlan warp_plan(int n, int K, int S, int dtype_bytes, int mode);
cudaError_t warp_launch_f32(int n, int mode, int grid, const WarpPlan& wp, const HeaParams<float>& p,
                            const DepthPack& dp, cudaStream_t st);
cudaError_t warp_launch_f64(int n, int mode, int grid, const WarpPlan& wp, const HeaParams<double>& p,
                            const DepthPack& dp, cudaStream_t st);

// fp32 HBM-streamed tier (hea_hbm.cu): n in [kHbmMinN, kHbmMaxN], modes 0 / 1 / 2
constexpr int kHbmMinN = 14, kHbmMaxN = 22;
struct HbmPlan {
    bool ok;
    int n;
    int tb_fwd, tb_rev;            // tile size (log2 amplitudes) of the forward / reverse pass kernels: 12 or 13
    int64_t Sc;                    // samples resident in the HBM workspace per chunk
    int grid_fwd, grid_rev, rows;  // persistent grids; rows = per-warp partial rows the reverse kernels write
    size_t smem_fwd, smem_rev;
    size_t off_psi, off_lam, off_epart, off_gval, off_mx, bytes;   // layout of the state area
};
HbmPlan hbm_plan(int64_t B, int n, int K, int mode);
cudaError_t hbm_run(const HeaParams<float>& p, const int* depth_host, int n, int K, int mode, const HbmPlan& pl,
                    char* state_ws, cudaStream_t st);

// fp32 tensor-core tier (hea_tc.cu): n = 5, diagonal observables, every mode of hea_reg_inst.cuh
size_t tc_workspace_bytes(int K, int S, int64_t B, int sms);
bool tc_outer_supported(int K);
cudaError_t tc_launch(int mode, int version, int sms, const HeaParams<float>& p, const float* w, const DepthPack& dp,
                      char* tc_ws, float* dbg, int* err_user, cudaStream_t st);

}  // namespace qon
