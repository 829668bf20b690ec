// Tensor-core tier (hea_tc.cuh, hea_tc2.cuh): launch glue.
#include <cstdlib>

#include "hea_dispatch.cuh"
#include "hea_tc3.cuh"

namespace qon {

// One CTA per SM as soon as there are that many tiles: a tile slot without a tile costs nothing (it stops after its last
// live tile), and a lone tile has the SM to itself — so a mid-size batch spreads over all SMs instead of filling a few.
static int tc_grid(int64_t B, int sms) {
    static const int cap = [] { const char* e = getenv("QON_TC_GRID"); return e ? atoi(e) : 0; }();     // experiments only
    if (cap > 0 && cap < sms) sms = cap;
    int64_t grid = (B + 127) / 128;
    if (grid > sms) grid = sms;
    return (int)(grid < 1 ? 1 : grid);
}

// The GEMM-form gradients keep 2 x grid accumulators of K x 8 KB: beyond this many blocks (2.4 GB at K = 1,024) the
// per-sublayer-moment step is used instead.
bool tc_outer_supported(int K) { return K <= kTcOuterMaxBlocks; }

size_t tc_workspace_bytes(int K, int S, int64_t B, int sms) {
    // operand images + flags (error word, max |g| bits) + (training step) one 256-byte state row per sample + the
    // outer-product accumulators of the GEMM-form weight gradients: one per (CTA, tile slot) and block, + their fp64 sums
    return (size_t)(K + S) * kTcImgBytes + 256 +
           (B > 0 ? (size_t)B * 256 + (tc_outer_supported(K) ? (size_t)2 * tc_grid(B, sms) * K * kTcAccLen * sizeof(float) +
                                                                    (size_t)K * kTcAccLen * sizeof(double) : 0) : 0);
}

static int tc_flags() {
    static const int flags = [] { const char* e = getenv("QON_TC_FLAGS"); return e ? atoi(e) : 0; }();
    return flags;
}

template <bool GRAD, bool GX, int ENC, bool DBG, bool SPLIT = false>
static cudaError_t tc_launch_t(int grid, const HeaParams<float>& p, const unsigned char* img, float* dbg, int* err,
                               float* state, cudaStream_t st, unsigned* gmax = nullptr) {
    using G = TcGeom<GRAD>;
    auto kern = hea_tc_kernel<GRAD, GX, ENC, DBG, SPLIT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
    if (e != cudaSuccess) return e;
    kern<<<grid, G::THREADS, G::SMEM, st>>>(p, img, dbg, err, state, tc_flags(), gmax);
    return cudaGetLastError();
}

template <bool GX, int ENC>
static cudaError_t tc_launch_rev(int grid, const HeaParams<float>& p, const unsigned char* img, int* err, const float* state,
                                 const unsigned* gmax, float* gacc, float* dbg, cudaStream_t st) {
    auto kern = hea_tc_rev_kernel<GX, ENC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcRev::SMEM);
    if (e != cudaSuccess) return e;
    kern<<<grid, TcRev::THREADS, TcRev::SMEM, st>>>(p, img, err, state, gmax, gacc, dbg, tc_flags());
    return cudaGetLastError();
}

// mode: hea_reg_inst.cuh (0 fwd | 1 grad + dL/dx | 2 grad | 3 fwd, fused encoding | 4 grad, fused encoding |
// 5 grad, fused encoding + frequency-layer gradients).  The training step, by version:
//   4 (default)  forward-only kernel, reverse kernel with GEMM-form weight gradients (hea_tc3.cuh), moment kernel
//   2            forward-only kernel + reverse kernel with per-sublayer Pauli-string moments (hea_tc2.cuh)
//   3            the whole step in ONE kernel (hea_tc2.cuh)                        — 2 and 3 are kept for A/B runs
// dbg: version 4 dumps the raw outer-product accumulator of its first tile / last block; versions 2, 3 run the
// one-kernel step with its state dumps.
cudaError_t tc_launch(int mode, int version, int sms, const HeaParams<float>& p, const float* w, const DepthPack& dp,
                      char* tc_ws, float* dbg, int* err_user, cudaStream_t st) {
    unsigned char* img = reinterpret_cast<unsigned char*>(tc_ws);
    const bool grad = mode == 1 || mode == 2 || mode == 4 || mode == 5;
    char* flag_block = tc_ws + (size_t)(p.K + p.S) * kTcImgBytes;
    int* err = err_user ? err_user : reinterpret_cast<int*>(flag_block);
    unsigned* gmax = reinterpret_cast<unsigned*>(flag_block + 64);
    float* state = reinterpret_cast<float*>(flag_block + 256);
    float* gacc = reinterpret_cast<float*>(flag_block + 256 + (size_t)(p.B > 0 ? p.B : 0) * 256);
    const bool outer = grad && version == 4 && tc_outer_supported(p.K);
    if (grad && version == 4 && !outer) version = 2;
    cudaError_t e = cudaMemsetAsync(flag_block, 0, 256, st);      // (a caller-provided error word is the caller's to clear)
    if (e != cudaSuccess) return e;
    if (grad) tc_prep_all_kernel<<<p.K + (outer ? p.K : p.S), 512, 0, st>>>(w, p.K, p.S, dp, img, img + (size_t)p.K * kTcImgBytes, outer ? 1 : 0);
    else tc_prep_kernel<<<p.K, 512, 0, st>>>(w, p.K, dp, img);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int64_t ntiles = (p.B + 127) / 128;
    auto grid_for = [&](int) { return tc_grid(p.B, sms); };
    if (!grad) {
        const int g = grid_for(4);
        if (mode == 0) return dbg ? tc_launch_t<false, false, 0, true>(g, p, img, dbg, err, nullptr, st)
                                  : tc_launch_t<false, false, 0, false>(g, p, img, dbg, err, nullptr, st);
        return tc_launch_t<false, false, 1, false>(g, p, img, dbg, err, nullptr, st);
    }
    const int g = grid_for(2);
    if (outer) {
        const unsigned char* rimg = img + (size_t)p.K * kTcImgBytes;
        e = (mode == 1 || mode == 2) ? tc_launch_t<false, false, 0, false>(grid_for(4), p, img, nullptr, err, state, st, gmax)
                                     : tc_launch_t<false, false, 1, false>(grid_for(4), p, img, nullptr, err, state, st, gmax);
        if (e != cudaSuccess) return e;
        switch (mode) {
            case 1: e = tc_launch_rev<true, 0>(g, p, rimg, err, state, gmax, gacc, dbg, st); break;
            case 2: e = tc_launch_rev<false, 0>(g, p, rimg, err, state, gmax, gacc, dbg, st); break;
            case 4: e = tc_launch_rev<false, 1>(g, p, rimg, err, state, gmax, gacc, dbg, st); break;
            case 5: e = tc_launch_rev<false, 2>(g, p, rimg, err, state, gmax, gacc, dbg, st); break;
            default: return cudaErrorInvalidValue;
        }
        if (e != cudaSuccess) return e;
        double* ysum = reinterpret_cast<double*>(gacc + (size_t)2 * g * p.K * kTcAccLen);
        const int nslots = (int)(ntiles < 2 * (int64_t)g ? ntiles : 2 * (int64_t)g);      // slots with a tile: a prefix (slot = t * grid + CTA)
        tc_slot_reduce_kernel<<<p.K * 8, 256, 0, st>>>(gacc, nslots, p.K, ysum);
        tc_moment_kernel<<<p.K, 1024, 0, st>>>(ysum, gmax, p.hdiag, w, p.K, dp, p.mpart);
        return cudaGetLastError();
    }
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
