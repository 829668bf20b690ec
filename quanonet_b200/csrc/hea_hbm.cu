// fp32 HBM-streamed tier: planning and the host-side launch sequence (two tile passes per sublayer).
#include <cstdlib>

#include "hea_dispatch.cuh"
#include "hea_hbm.cuh"

namespace qon {

namespace {
inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

using PassKern = void (*)(const HeaParams<float>, const HbmPass, const HbmBuffers);

// bulk = pass A (contiguous tiles, cp.async.bulk); QON_HBM_BULK=0 keeps the per-element gather for A/B runs
bool bulk_enabled() {
    static const bool v = [] { const char* e = getenv("QON_HBM_BULK"); return !(e && atoi(e) == 0); }();
    return v;
}
PassKern pass_kernel(bool reverse, int tb, bool bulk) {
    if (bulk) {
        if (reverse) return tb == 13 ? hea_hbm_pass_kernel<true, 13, true> : hea_hbm_pass_kernel<true, 12, true>;
        return tb == 13 ? hea_hbm_pass_kernel<false, 13, true> : hea_hbm_pass_kernel<false, 12, true>;
    }
    if (reverse) return tb == 13 ? hea_hbm_pass_kernel<true, 13, false> : hea_hbm_pass_kernel<true, 12, false>;
    return tb == 13 ? hea_hbm_pass_kernel<false, 13, false> : hea_hbm_pass_kernel<false, 12, false>;
}

// tile size per direction; QON_HBM_TB_FWD / QON_HBM_TB_REV override (12 or 13) for A/B runs
int tile_bits(bool reverse) {
    static const int f = [] { const char* e = getenv("QON_HBM_TB_FWD"); const int v = e ? atoi(e) : 12; return v == 13 ? 13 : 12; }();
    static const int r = [] { const char* e = getenv("QON_HBM_TB_REV"); const int v = e ? atoi(e) : 12; return v == 13 ? 13 : 12; }();
    return reverse ? r : f;
}

// first qubit of pass B: pass A takes all TB low qubits.  A balanced split (QON_HBM_SPLIT=8 at n = 16: 8 fused gates
// per pass instead of 13 + 3) was measured and is SLOWER (forward 53.6 vs 49.7 ms at n = 16, B = 1,184; equal at
// n = 14 / 18): pass B's tiles then gather 128-byte chunks and the streaming phases, not the gates, bound both passes
// (profiles/r2_hbm_split_ab.md).  The override stays for experiments; the tile index is general.
int split_qubit(int n, int tb) {
    static const int ov = [] { const char* e = getenv("QON_HBM_SPLIT"); return e ? atoi(e) : 0; }();
    int qa = ov > 0 ? ov : tb;
    if (qa > tb) qa = tb;                 // pass A's tile holds qubits 0..TB-1
    if (qa < n - tb) qa = n - tb;         // pass B's tile holds at most TB qubits
    if (qa < 1) qa = 1;
    return qa;
}

void make_pass(HbmPass& hp, int n, int tb, bool passB, bool reverse) {
    const int qa = split_qubit(n, tb);
    hp.n = n;
    hp.c = passB ? tb - (n - qa) : tb;
    hp.qoff = passB ? n - tb : 0;     // qubit of local bit l >= c in pass B: l - c + qa = l + (n - tb)
    const int lo2 = smem_lo(tb, 2);
    int wins[3], masks[3], cnt = 0;
    for (int pw = 0; pw < 3; ++pw) {
        const int lo = smem_lo(tb, pw);
        int m = 0;
        for (int r = 0; r < 5; ++r) {
            const int l = lo + r;
            bool gated;
            if (!passB) gated = (pw < 2 ? true : l >= 10) && l < qa;            // pass A: local bits [0, qa) once
            else gated = l >= hp.c && (pw == 2 || l < (pw == 1 ? lo2 : 5));     // pass B: local bits [c, TB) once
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
    for (int bit = 0; bit < kMaxTileBits; ++bit) {
        unsigned g = 0;
        if (bit < tb) g = tb == 13 ? hbm_gidx<13>(1u << bit, 0u, hp.c, n) : hbm_gidx<12>(1u << bit, 0u, hp.c, n);
        // plain and ring-permuted position increments; hbm_run() picks per pass (ring_load / ring_store)
        hp.plainp[bit] = bit < tb ? hbm_pos(g) : 0u;
        hp.ringp[bit] = bit < tb ? hbm_pos(hbm_ring(g, n)) : 0u;
        hp.ldp[bit] = hp.stp[bit] = hp.plainp[bit];
    }
}
}  // namespace

HbmPlan hbm_plan(int64_t B, int n, int K, int mode) {
    HbmPlan pl{};
    pl.ok = false;
    if (n < kHbmMinN || n > kHbmMaxN || mode < 0 || mode > 2) return pl;
    const bool grad = mode != 0;
    pl.n = n;
    pl.tb_fwd = tile_bits(false);
    pl.tb_rev = tile_bits(true);
    const size_t state = (size_t)8 << n;
    // chunk of samples resident in HBM: ~2 GB of state, at least 8 samples, at most B
    int64_t sc = (int64_t)(((size_t)2 << 30) / (state * (grad ? 2 : 1)));
    if (sc < 8) sc = 8;
    if (sc > 4096) sc = 4096;
    if (sc > B) sc = B > 0 ? B : 1;
    pl.Sc = sc;
    pl.smem_fwd = (size_t)8 << pl.tb_fwd;              // psi tile
    pl.smem_rev = (size_t)16 << pl.tb_rev;             // psi + lam tiles
    PassKern kf = pass_kernel(false, pl.tb_fwd, false), kr = pass_kernel(true, pl.tb_rev, false);
    for (int bulk = 0; bulk < 2; ++bulk) {     // both variants of each direction need the opt-in (same footprint)
        cudaFuncSetAttribute(pass_kernel(false, pl.tb_fwd, bulk != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_fwd);
        cudaFuncSetAttribute(pass_kernel(true, pl.tb_rev, bulk != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_rev);
    }
    if (cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_fwd) != cudaSuccess ||
        cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_rev) != cudaSuccess) {
        cudaGetLastError();
        return pl;
    }
    int of = 0, orv = 0, sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&of, kf, 1 << (pl.tb_fwd - 5), pl.smem_fwd) != cudaSuccess || of < 1 ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&orv, kr, 1 << (pl.tb_rev - 5), pl.smem_rev) != cudaSuccess || orv < 1) {
        cudaGetLastError();
        return pl;
    }
    pl.grid_fwd = sms * of;
    pl.grid_rev = sms * orv;
    pl.rows = pl.grid_rev * ((1 << (pl.tb_rev - 5)) / 32);
    const int64_t Tm = (int64_t)1 << (n - kMeasureBits), Tr = (int64_t)1 << (n - pl.tb_rev);
    size_t off = 0;
    pl.off_psi = off; off = up256(off + (size_t)sc * state);
    pl.off_lam = off; if (grad) off = up256(off + (size_t)sc * state);
    pl.off_epart = off; off = up256(off + (size_t)sc * Tm * sizeof(float));
    pl.off_gval = off; off = up256(off + (size_t)sc * sizeof(float));
    pl.off_mx = off; if (mode == 1) off = up256(off + (size_t)sc * Tr * 3 * n * K * sizeof(float));
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
    const int64_t N = (int64_t)1 << n;
    const int tbf = pl.tb_fwd, tbr = pl.tb_rev;
    const int thr_f = 1 << (tbf - 5), thr_r = 1 << (tbr - 5);
    const bool bulk = bulk_enabled();
    PassKern kfA = pass_kernel(false, tbf, bulk), kfB = pass_kernel(false, tbf, false);
    PassKern krA = pass_kernel(true, tbr, bulk), krB = pass_kernel(true, tbr, false);
    cudaError_t e;
    for (PassKern k : {kfA, kfB})
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_fwd)) != cudaSuccess) return e;
    for (PassKern k : {krA, krB})
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_rev)) != cudaSuccess) return e;

    HbmPass A_f, B_f, A_r, B_r;
    make_pass(A_f, n, tbf, false, false);
    make_pass(B_f, n, tbf, true, false);
    make_pass(A_r, n, tbr, false, true);
    make_pass(B_r, n, tbr, true, true);

    for (int64_t b0 = 0; b0 < p.B; b0 += pl.Sc) {
        const int64_t nb = p.B - b0 < pl.Sc ? p.B - b0 : pl.Sc;
        auto grid_for = [&](int tb, int cap) {
            const int64_t tiles = nb << (n - tb);
            return (int)(tiles < cap ? tiles : cap);
        };
        auto fill = [&](HbmPass hp, int tb, int s, int k, int j) {
            hp.s = s; hp.kblk = k; hp.fold = j == 0;
            hp.ring_store = 0; hp.ring_load = 0; hp.tiles_log2 = n - tb;
            hp.need_gx = need_gx; hp.b0 = b0; hp.nb = nb;
            return hp;
        };
        auto route = [&](HbmPass& hp, bool ring_load, bool ring_store) {    // which position tables the pass streams through
            hp.ring_load = ring_load; hp.ring_store = ring_store;
            for (int bit = 0; bit < kMaxTileBits; ++bit) {
                hp.ldp[bit] = ring_load ? hp.ringp[bit] : hp.plainp[bit];
                hp.stp[bit] = ring_store ? hp.ringp[bit] : hp.plainp[bit];
            }
        };
        {
            int64_t blocks = (N * nb + 255) / 256;
            if (blocks > 4096) blocks = 4096;
            hea_hbm_init_kernel<<<(int)blocks, 256, 0, st>>>(hb.psi, N, nb);
        }
        int s = 0;
        for (int k = 0; k < K; ++k)
            for (int j = 0; j < depth[k]; ++j, ++s) {
                HbmPass a = fill(A_f, tbf, s, k, j);
                kfA<<<grid_for(tbf, pl.grid_fwd), thr_f, pl.smem_fwd, st>>>(p, a, hb);
                HbmPass b = fill(B_f, tbf, s, k, j);
                route(b, false, true);
                kfB<<<grid_for(tbf, pl.grid_fwd), thr_f, pl.smem_fwd, st>>>(p, b, hb);
            }
        {
            const int64_t mt = nb << (n - kMeasureBits);
            const int g = (int)(mt < 4 * 148 ? mt : 4 * 148);
            if (grad) hea_hbm_measure_kernel<true><<<g, 256, 0, st>>>(p, n, nb, hb);
            else hea_hbm_measure_kernel<false><<<g, 256, 0, st>>>(p, n, nb, hb);
            hea_hbm_seed_kernel<<<(int)((nb + 127) / 128), 128, 0, st>>>(p, n, b0, nb, hb, grad ? 1 : 0);
        }
        if (grad) {
            for (int k = K - 1; k >= 0; --k)
                for (int j = depth[k] - 1; j >= 0; --j) {
                    --s;
                    HbmPass b = fill(B_r, tbr, s, k, j);
                    route(b, true, false);
                    krB<<<grid_for(tbr, pl.grid_rev), thr_r, pl.smem_rev, st>>>(p, b, hb);
                    HbmPass a = fill(A_r, tbr, s, k, j);
                    krA<<<grid_for(tbr, pl.grid_rev), thr_r, pl.smem_rev, st>>>(p, a, hb);
                }
            if (need_gx) {
                int64_t blocks = (nb * n * K + 255) / 256;
                if (blocks > 2048) blocks = 2048;
                hea_hbm_gx_kernel<<<(int)blocks, 256, 0, st>>>(p, n, n - tbr, b0, nb, hb);
            }
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace qon
