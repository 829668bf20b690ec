// fp32 HBM-streamed tier: planning and the host-side launch sequence (two tile passes per sublayer).
#include "hea_dispatch.cuh"
#include "hea_hbm.cuh"

namespace qon {

namespace {
inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

unsigned ring_host(unsigned k, int n) {
    for (int i = 0; i < n; ++i) k ^= ((k >> (i + 1 == n ? 0 : i + 1)) & 1u) << i;
    return k;
}
unsigned gidx_host(unsigned l, unsigned t, int c) {
    return (l & ((1u << c) - 1u)) | (t << c) | ((l >> c) << kTileBits);
}

void make_pass(HbmPass& hp, int n, bool passB, bool reverse) {
    hp.n = n;
    hp.c = passB ? 2 * kTileBits - n : kTileBits;
    hp.qoff = passB ? kTileBits - hp.c : 0;
    int wins[3], masks[3], cnt = 0;
    for (int pw = 0; pw < 3; ++pw) {
        const int lo = pw == 0 ? 0 : (pw == 1 ? 5 : 8);
        int m = 0;
        for (int r = 0; r < 5; ++r) {
            const int l = lo + r;
            bool gated;
            if (!passB) gated = pw < 2 ? true : l >= 10;                       // pass A: every local bit once
            else gated = l >= hp.c && (pw == 2 || l < (pw == 1 ? 8 : 5));      // pass B: local bits [c, 13) once
            if (gated) m |= 1 << r;
        }
        if (m) { wins[cnt] = pw; masks[cnt] = m; ++cnt; }
    }
    hp.nwin = cnt;
    for (int i = 0; i < cnt; ++i) {
        const int src = reverse ? cnt - 1 - i : i;
        hp.win[i] = wins[src];
        hp.mask[i] = masks[src];
    }
    for (int bit = 0; bit < kTileBits; ++bit) hp.ringp[bit] = ring_host(gidx_host(1u << bit, 0u, hp.c), n);
}
}  // namespace

HbmPlan hbm_plan(int64_t B, int n, int K, int mode) {
    HbmPlan pl{};
    pl.ok = false;
    if (n < kHbmMinN || n > kHbmMaxN || mode < 0 || mode > 2) return pl;
    const bool grad = mode != 0;
    pl.n = n;
    pl.tiles_log2 = n - kTileBits;
    const size_t state = (size_t)8 << n;
    // chunk of samples resident in HBM: ~2 GB of state, at least 8 samples, at most B
    int64_t sc = (int64_t)(((size_t)2 << 30) / (state * (grad ? 2 : 1)));
    if (sc < 8) sc = 8;
    if (sc > 4096) sc = 4096;
    if (sc > B) sc = B > 0 ? B : 1;
    pl.Sc = sc;
    const int64_t T = (int64_t)1 << pl.tiles_log2;
    pl.smem_fwd = (size_t)(8 << kTileBits);          // psi tile: 64 KB -> two forward CTAs per SM
    pl.smem_rev = (size_t)(8 << kTileBits) * 2;      // psi + lam tiles
    auto kf = hea_hbm_pass_kernel<false>;
    auto kr = hea_hbm_pass_kernel<true>;
    if (cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_fwd) != cudaSuccess ||
        cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_rev) != cudaSuccess) {
        cudaGetLastError();
        return pl;
    }
    int of = 0, orv = 0, sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&of, kf, kHbmThreads, pl.smem_fwd) != cudaSuccess || of < 1 ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&orv, kr, kHbmThreads, pl.smem_rev) != cudaSuccess || orv < 1) {
        cudaGetLastError();
        return pl;
    }
    pl.grid_fwd = sms * of;
    pl.grid_rev = sms * orv;
    pl.rows = pl.grid_rev * (kHbmThreads / 32);
    size_t off = 0;
    pl.off_psi = off; off = up256(off + (size_t)sc * state);
    pl.off_lam = off; if (grad) off = up256(off + (size_t)sc * state);
    pl.off_epart = off; off = up256(off + (size_t)sc * T * sizeof(float));
    pl.off_gval = off; off = up256(off + (size_t)sc * sizeof(float));
    pl.off_mx = off; if (mode == 1) off = up256(off + (size_t)sc * T * 3 * n * K * sizeof(float));
    pl.bytes = off;
    pl.ok = true;
    return pl;
}

cudaError_t hbm_run(const HeaParams<float>& p, const int* depth, int n, int K, int mode, const HbmPlan& pl, char* ws,
                    cudaStream_t st) {
    const bool grad = mode != 0, need_gx = mode == 1;
    HbmBuffers hb{};
    hb.psi = (u64*)(ws + pl.off_psi);
    hb.lam = grad ? (u64*)(ws + pl.off_lam) : nullptr;
    hb.epart = (float*)(ws + pl.off_epart);
    hb.gval = (float*)(ws + pl.off_gval);
    hb.mxpart = need_gx ? (float*)(ws + pl.off_mx) : nullptr;
    const int64_t N = (int64_t)1 << n, T = (int64_t)1 << pl.tiles_log2;
    auto kf = hea_hbm_pass_kernel<false>;
    auto kr = hea_hbm_pass_kernel<true>;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_fwd)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_rev)) != cudaSuccess) return e;

    HbmPass A_f, B_f, A_r, B_r;
    make_pass(A_f, n, false, false);
    make_pass(B_f, n, true, false);
    make_pass(A_r, n, false, true);
    make_pass(B_r, n, true, true);

    for (int64_t b0 = 0; b0 < p.B; b0 += pl.Sc) {
        const int64_t nb = p.B - b0 < pl.Sc ? p.B - b0 : pl.Sc;
        const int64_t tiles = nb * T;
        auto grid_for = [&](int cap) { return (int)(tiles < cap ? tiles : cap); };
        auto fill = [&](HbmPass hp, int s, int k, int j, bool reverse) {
            hp.s = s; hp.kblk = k; hp.fold = j == 0; hp.reverse = reverse;
            hp.ring_store = 0; hp.ring_load = 0; hp.scale_lam = 0; hp.tiles_log2 = pl.tiles_log2;
            hp.need_gx = need_gx; hp.b0 = b0; hp.nb = nb;
            return hp;
        };
        {
            int64_t blocks = (N * nb + 255) / 256;
            if (blocks > 4096) blocks = 4096;
            hea_hbm_init_kernel<<<(int)blocks, 256, 0, st>>>(hb.psi, N, nb);
        }
        int s = 0;
        for (int k = 0; k < K; ++k)
            for (int j = 0; j < depth[k]; ++j, ++s) {
                HbmPass a = fill(A_f, s, k, j, false);
                kf<<<grid_for(pl.grid_fwd), kHbmThreads, pl.smem_fwd, st>>>(p, a, hb);
                HbmPass b = fill(B_f, s, k, j, false);
                b.ring_store = 1;
                kf<<<grid_for(pl.grid_fwd), kHbmThreads, pl.smem_fwd, st>>>(p, b, hb);
            }
        {
            const int g = (int)(tiles < 4 * 148 ? tiles : 4 * 148);
            if (grad) hea_hbm_measure_kernel<true><<<g, 256, 0, st>>>(p, n, pl.tiles_log2, nb, hb);
            else hea_hbm_measure_kernel<false><<<g, 256, 0, st>>>(p, n, pl.tiles_log2, nb, hb);
            hea_hbm_seed_kernel<<<(int)((nb + 127) / 128), 128, 0, st>>>(p, pl.tiles_log2, b0, nb, hb, grad ? 1 : 0);
        }
        if (grad) {
            bool first = true;
            for (int k = K - 1; k >= 0; --k)
                for (int j = depth[k] - 1; j >= 0; --j) {
                    --s;
                    HbmPass b = fill(B_r, s, k, j, true);
                    b.ring_load = 1;
                    b.scale_lam = first ? 1 : 0;
                    first = false;
                    kr<<<grid_for(pl.grid_rev), kHbmThreads, pl.smem_rev, st>>>(p, b, hb);
                    HbmPass a = fill(A_r, s, k, j, true);
                    kr<<<grid_for(pl.grid_rev), kHbmThreads, pl.smem_rev, st>>>(p, a, hb);
                }
            if (need_gx) {
                int64_t blocks = (nb * n * K + 255) / 256;
                if (blocks > 2048) blocks = 2048;
                hea_hbm_gx_kernel<<<(int)blocks, 256, 0, st>>>(p, n, pl.tiles_log2, b0, nb, hb);
            }
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace qon
