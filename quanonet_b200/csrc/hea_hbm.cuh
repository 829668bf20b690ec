// HBM-streamed tier (fp32, n = 14 ... 22): the states of a chunk of samples live in the HBM workspace and
// are streamed through shared memory in TILES of 2^TB amplitudes (TB = 12 or 13: 32 / 64 KB of psi, twice
// that with lam in the reverse sweep).  One ansatz sublayer = TWO passes over HBM, each a kernel launch
// over all tiles of the chunk:
//
//   pass A  tile = 2^TB contiguous amplitudes           -> gates on qubits 0..qa-1 (windows [0,5) [5,10) [TB-5,TB))
//   pass B  tile = 2^c-amplitude contiguous chunks x the TB-c highest qubits (c = TB - (n - qa))
//                                                       -> gates on qubits qa..n-1, then the CNOT ring as a
//                                                          GF(2)-linear scatter on the way back to HBM
// The split point qa balances the two passes (round 1 gave pass A all TB low qubits: at n = 16 that is 13 fused
// gates per 1 MB of traffic — FP32-bound at any efficiency — against 3 gates in the HBM-bound pass B; with qa = 8
// both passes carry 8 gates).
//
// Inside a tile the register-blocked FFMA2 window passes of the shared-memory tier (hea_smem.cuh) are reused
// (same swizzle, same constant-memory offset tables, n = TB geometry, 2^(TB-5) threads per tile); windows
// always run five gates and take identity coefficients for bits that carry no gate in this pass (pass B is
// HBM-bound, so the padding is free).  Tiles travel HBM -> shared memory with 8-byte cp.async straight into
// the swizzled slots.  Smaller tiles (TB = 12) put 4 forward / 2 reverse CTAs on an SM, so one CTA's HBM
// phase overlaps another's FP32 phase.  Algorithmic HBM traffic: 2 x (read + write) x 8 B x 2^n per sublayer
// forward, 2 x that in the reverse sweep (psi and lam); the encoding layers cost none (RX folded).
//
// The reverse sweep runs the same two passes backwards on (psi, lam): pass B first (gathers through the
// ring permutation), then pass A; lam stays unscaled (= H psi) and the per-sample upstream gradient g
// multiplies the moments where they are consumed.  Shared-parameter moments go to the per-warp rows as in
// the other tiers, per-sample dL/dx moments to a per-(sample, tile) partial buffer summed in fixed order.
// Reference semantics: core/quantum_circuits_tq.py:65-127.
#pragma once
#include "hea_smem.cuh"

namespace qon {

constexpr int kMaxTileBits = 13;
constexpr int kMeasureBits = 13;          // granularity of the expectation-value partials (independent of TB)

struct HbmPass {
    int n, c;                     // c = contiguous low bits of a tile chunk (TB for pass A, 2 TB - n for pass B)
    int nwin;                     // windows to run, in execution order
    int win[3], mask[3];          // window index into the n = TB tables and its gate mask
    int qoff;                     // qubit of local bit l (for gated bits) = l + qoff
    int s, kblk, fold;            // sublayer, block (x columns kblk*n ..), RX folded into this sublayer
    int ring_store, ring_load;    // scatter / gather through the CNOT-ring permutation
    int tiles_log2;               // log2(tiles per sample) = n - TB
    int need_gx;
    int64_t b0, nb;               // first sample of the chunk, samples in the chunk
    // HBM position of local index l = base(t) ^ XOR over the set bits b of l of tbl[b]  (GF(2)-linear: the tile
    // index map, the CNOT-ring permutation when asked, and the storage swizzle of hbm_pos() all are)
    unsigned ldp[kMaxTileBits], stp[kMaxTileBits];
    unsigned plainp[kMaxTileBits], ringp[kMaxTileBits];   // host side: the two candidate tables (hbm_run copies one into ldp / stp)
};

struct HbmBuffers {
    u64* psi;          // [Sc][2^n]
    u64* lam;          // [Sc][2^n]            (reverse only)
    float* epart;      // [Sc][2^(n-13)]
    float* gval;       // [Sc]                 upstream gradient per sample
    float* mxpart;     // [Sc][T][3*n*K]       per-tile Pauli moments of the folded RX gates (T = reverse tiles)
};

__device__ __forceinline__ void cp_async8(unsigned smem_addr, const void* gptr) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_addr), "l"(gptr) : "memory");
}

// window group <-> shared-memory tile; the tile base is ADDED (not XORed as in hea_smem.cuh) so the tile
// needs no size alignment
__device__ __forceinline__ void tile_load(SmemState& st, unsigned region, unsigned gb, const int* off) {
#pragma unroll
    for (int i = 0; i < 32; ++i) st.a[i] = lds64(region + (gb ^ (unsigned)off[i]));
}
__device__ __forceinline__ void tile_store(const SmemState& st, unsigned region, unsigned gb, const int* off) {
#pragma unroll
    for (int i = 0; i < 32; ++i) sts64(region + (gb ^ (unsigned)off[i]), st.a[i]);
}

// Storage order of a state in HBM: amplitude k lives at position k ^ ((k >> 5) & 31) — the shared-memory swizzle of
// hea_smem.cuh applied to the global index (an involution touching bits 0..4 only).  A pass-A tile (2^TB contiguous
// amplitudes) is then ALREADY in its shared-memory layout in HBM and moves with one bulk copy each way.
__host__ __device__ __forceinline__ unsigned hbm_pos(unsigned k) { return k ^ ((k >> 5) & 31u); }

__host__ __device__ __forceinline__ unsigned hbm_ring(unsigned k, int n) {
    for (int i = 0; i < n; ++i) k ^= ((k >> (i + 1 == n ? 0 : i + 1)) & 1u) << i;
    return k;
}
// global amplitude index of local index l in tile t (TB-bit tiles of an n-qubit state): the tile holds the c lowest
// index bits and the TB - c highest; the n - TB bits in between enumerate the tiles
template <int TB>
__host__ __device__ __forceinline__ unsigned hbm_gidx(unsigned l, unsigned t, int c, int n) {
    return (l & ((1u << c) - 1u)) | (t << c) | ((l >> c) << (c + n - TB));
}

// BULK (pass A: contiguous tiles): the tile travels with cp.async.bulk (TMA engine, one instruction per direction,
// completion on an mbarrier / bulk group) instead of 32 eight-byte cp.async + index arithmetic per thread.
#ifndef QON_HBM_FWD_BLOCKS
#define QON_HBM_FWD_BLOCKS 5      // resident CTAs per SM the forward pass kernels (TB = 12) are compiled for: 5 x 128 threads at 96 registers measured +3 % over 4 at 123 (6 at 80: -9 %)
#endif
#ifndef QON_HBM_REV_BLOCKS
#define QON_HBM_REV_BLOCKS 3      // the same for the reverse pass kernels (psi + lam tiles: 64 KB per CTA): 3 CTAs at 168 registers measured -7 % time at n = 16 against 2 at 247
#endif
template <bool REVERSE, int TB, bool BULK>
__global__ void __launch_bounds__(1 << (TB - 5), REVERSE ? (TB == 13 ? 1 : QON_HBM_REV_BLOCKS) : (TB == 13 ? 2 : QON_HBM_FWD_BLOCKS))
hea_hbm_pass_kernel(const HeaParams<float> p, const HbmPass hp, const HbmBuffers hb) {
    using State = SmemState;
    constexpr int THREADS = 1 << (TB - 5);
    constexpr int WARPS = THREADS / 32;
    constexpr int TIDB = TB - 5;                       // local bits that come from the thread index when streaming
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float s_red[WARPS][16];
    const int n = hp.n, c = hp.c;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned psi_base = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned lam_base = psi_base + (8u << TB);
    const int VP = (3 * n + 3) / 4 * 4;
    const int64_t gwarp = (int64_t)blockIdx.x * WARPS + warp;
    float* mrow = REVERSE ? p.mpart + gwarp * p.rowlen : nullptr;
    const int64_t T = (int64_t)1 << hp.tiles_log2;
    const int64_t ntiles = hp.nb * T;
    const int64_t N = (int64_t)1 << n;

    auto group_base = [&](int lo) -> unsigned {
        const int kb = ((tid >> lo) << (lo + kSmemW)) | (tid & ((1 << lo) - 1));
        return (unsigned)(8 * smem_swz(kb));
    };
    // HBM position of local index l = tid + THREADS * r: tile part + thread part once per tile, r part = constants
    auto pos_base = [&](unsigned t, bool ring, const unsigned* tbl) -> unsigned {
        unsigned g = hbm_gidx<TB>(0u, t, c, n);
        if (ring) g = hbm_ring(g, n);
        unsigned v = hbm_pos(g);
#pragma unroll
        for (int bit = 0; bit < TIDB; ++bit)
            if ((tid >> bit) & 1) v ^= tbl[bit];
        return v;
    };
    __shared__ __align__(8) unsigned long long bulk_bar;
    unsigned bulk_par = 0;
    if constexpr (BULK) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&bulk_bar)) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t sl = tile >> hp.tiles_log2;            // sample slot in the chunk
        const unsigned t = (unsigned)(tile & (T - 1));
        const int64_t b = hp.b0 + sl;
        const u64* gpsi = hb.psi + sl * N;
        u64* gpsi_w = hb.psi + sl * N;
        const u64* glam = REVERSE ? hb.lam + sl * N : nullptr;
        u64* glam_w = REVERSE ? hb.lam + sl * N : nullptr;
        const float gsample = REVERSE ? hb.gval[sl] : 1.f;

        // ---------------- HBM -> shared memory ----------------
        if constexpr (BULK) {
            const unsigned bar = (unsigned)__cvta_generic_to_shared(&bulk_bar);
            if (tid == 0) {
                constexpr unsigned kBytes = 8u << TB;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(REVERSE ? 2 * kBytes : kBytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(psi_base),
                             "l"(gpsi + ((int64_t)t << TB)), "r"(kBytes), "r"(bar) : "memory");
                if constexpr (REVERSE)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(lam_base),
                                 "l"(glam + ((int64_t)t << TB)), "r"(kBytes), "r"(bar) : "memory");
            }
            unsigned ok = 0;
            for (int it = 0; it < 200000 && !ok; ++it)     // bounded: ~4 s; a copy that never lands must not hang the GPU
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(bar), "r"(bulk_par), "r"(20000u) : "memory");
            bulk_par ^= 1u;
        } else {
            const unsigned rbase = pos_base(t, hp.ring_load != 0, hp.ldp);
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const unsigned l = (unsigned)(tid + THREADS * r);      // consecutive lanes -> consecutive amplitudes
                unsigned src = rbase;
#pragma unroll
                for (int bit = 0; bit < 5; ++bit)
                    if ((r >> bit) & 1) src ^= hp.ldp[TIDB + bit];
                const unsigned slot = (unsigned)(8 * smem_swz((int)l));
                cp_async8(psi_base + slot, gpsi + src);
                if constexpr (REVERSE) cp_async8(lam_base + slot, glam + src);
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            __syncthreads();
        }

        // ---------------- window passes on the tile ----------------
        const float* xk = p.x + b * p.ldx + (int64_t)hp.kblk * n;
#pragma unroll 1
        for (int wdx = 0; wdx < hp.nwin; ++wdx) {
            const int pw = hp.win[wdx], mask = hp.mask[wdx];
            const int lo = smem_lo(TB, pw);
            const unsigned gb = group_base(lo);
            const int* off = c_smem_tbl.off[TB][pw];
            State st;
            tile_load(st, psi_base, gb, off);
            if constexpr (!REVERSE) {
                static_for<kSmemW>([&](auto Rc) {
                    constexpr int R = decltype(Rc)::value;
                    const bool gated = (mask >> R) & 1;
                    const int q = lo + R + hp.qoff;
                    Vec4<float> u{1.f, 0.f, 0.f, 0.f};
                    if (gated) u = ldg4(p.ucoef + (int64_t)hp.s * n + q);
                    float ar = u.x, ai = u.y, br = u.z, bi = u.w;
                    if (gated && hp.fold) fold_rx_coef(u, __ldg(xk + q), ar, ai, br, bi);
                    apply_u<R, false>(st, ar, ai, br, bi, lane);
                });
                tile_store(st, psi_base, gb, off);
            } else {
                State lm;
                tile_load(lm, lam_base, gb, off);
                float mv[15];
                static_for<kSmemW>([&](auto Rc) {
                    constexpr int R = kSmemW - 1 - decltype(Rc)::value;
                    const bool gated = (mask >> R) & 1;
                    const int q = lo + R + hp.qoff;
                    Vec4<float> u{1.f, 0.f, 0.f, 0.f};
                    if (gated) u = ldg4(p.ucoef + (int64_t)hp.s * n + q);
                    float ar = u.x, ai = u.y, br = u.z, bi = u.w;
                    if (gated && hp.fold) fold_rx_coef(u, __ldg(xk + q), ar, ai, br, bi);
                    bwd_group<R>(st, lm, ar, ai, br, bi, lane, mv[3 * R], mv[3 * R + 1], mv[3 * R + 2]);
                });
                tile_store(st, psi_base, gb, off);
                tile_store(lm, lam_base, gb, off);
                // shared-parameter moments -> this warp's partial row; keep warp totals for the per-tile dL/dx part
                float bv[16];
#pragma unroll
                for (int i = 0; i < 15; ++i) bv[i] = mv[i];
                bv[15] = 0.f;
                const float tot = gsample * butterfly_reduce<float, 16>(bv, lane);
                const int slot = lane >> 1, R = slot / 3;
                const bool mine = (lane & 1) == 0 && slot < 15 && ((mask >> R) & 1);
                if (mine) atomicAdd(mrow + (int64_t)hp.s * VP + 3 * (lo + R + hp.qoff) + slot % 3, tot);
                if (hp.fold && hp.need_gx) {
                    if ((lane & 1) == 0 && slot < 15) s_red[warp][slot] = tot;
                    __syncthreads();
                    if (tid < 15 && ((mask >> (tid / 3)) & 1)) {
                        float acc = 0.f;
                        for (int w = 0; w < WARPS; ++w) acc += s_red[w][tid];
                        const int q = lo + tid / 3 + hp.qoff;
                        hb.mxpart[(sl * T + t) * (int64_t)(3 * n * p.K) + ((int64_t)hp.kblk * n + q) * 3 + tid % 3] = acc;
                    }
                }
            }
            __syncthreads();
        }

        // ---------------- shared memory -> HBM (scatter through the ring when asked) ----------------
        if constexpr (BULK) {
            // the window passes ended with a barrier; make the generic-proxy writes visible to the async proxy, then
            // ONE bulk store per state; the tile buffer is reused only after the stores have read it
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                constexpr unsigned kBytes = 8u << TB;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gpsi_w + ((int64_t)t << TB)), "r"(psi_base),
                             "r"(kBytes) : "memory");
                if constexpr (REVERSE)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(glam_w + ((int64_t)t << TB)),
                                 "r"(lam_base), "r"(kBytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncthreads();
        } else {
            const unsigned rbase = pos_base(t, hp.ring_store != 0, hp.stp);
#pragma unroll 8
            for (int r = 0; r < 32; ++r) {
                const unsigned l = (unsigned)(tid + THREADS * r);
                unsigned dst = rbase;
#pragma unroll
                for (int bit = 0; bit < 5; ++bit)
                    if ((r >> bit) & 1) dst ^= hp.stp[TIDB + bit];
                const unsigned slot = (unsigned)(8 * smem_swz((int)l));
                gpsi_w[dst] = lds64(psi_base + slot);
                if constexpr (REVERSE) glam_w[dst] = lds64(lam_base + slot);
            }
            __syncthreads();
        }
    }
    if constexpr (BULK) {      // the last tile's stores must complete before the CTA's shared memory goes away
        if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// |0...0> for every sample slot of the chunk
__global__ void hea_hbm_init_kernel(u64* psi, int64_t N, int64_t nb) {
    const int64_t total = N * nb;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        psi[i] = (i % N) == 0 ? pack2(1.f, 0.f) : 0ull;
}

// E partials per (sample, chunk of 2^13 contiguous amplitudes); lam = H psi for the reverse sweep
template <bool GRAD>
__global__ void __launch_bounds__(256) hea_hbm_measure_kernel(const HeaParams<float> p, int n, int64_t nb,
                                                              const HbmBuffers hb) {
    __shared__ float red[8];
    const int64_t T = (int64_t)1 << (n - kMeasureBits), N = (int64_t)1 << n;
    for (int64_t tile = blockIdx.x; tile < nb * T; tile += gridDim.x) {
        const int64_t sl = tile / T, t = tile % T;
        const u64* ps = hb.psi + sl * N;
        float e = 0.f;
        for (int i = threadIdx.x; i < (1 << kMeasureBits); i += blockDim.x) {
            const int64_t pos = t * (1 << kMeasureBits) + i;         // storage position; amplitude index k = hbm_pos(pos)
            const int64_t k = (int64_t)hbm_pos((unsigned)pos);
            const u64 v = ps[pos];
            u64 h;
            if (p.pauli == 0) {
                h = mul2<0>(__ldg(p.hdiag + k), v);
            } else {
                h = mul2<0>(p.offset, v);
                for (int q = 0; q < n; ++q) {
                    const u64 f = ps[hbm_pos((unsigned)(k ^ ((int64_t)1 << q)))];
                    if (p.pauli == 1) h = fma2<0>(p.coeff, f, h);
                    else h = fma2<2>(((k >> q) & 1) ? p.coeff : -p.coeff, f, h);
                }
            }
            float vr, vi, hr, hi;
            unpack2(v, vr, vi);
            unpack2(h, hr, hi);
            e = fmaf(vr, hr, e);
            e = fmaf(vi, hi, e);
            if constexpr (GRAD) hb.lam[sl * N + pos] = h;
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) e += __shfl_xor_sync(0xffffffffu, e, m);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = e;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int w = 0; w < 8; ++w) s += red[w];
            hb.epart[sl * T + t] = s;
        }
    }
}

// per sample: E = sum of the partials (fixed order), out, upstream gradient g
__global__ void hea_hbm_seed_kernel(const HeaParams<float> p, int n, int64_t b0, int64_t nb, const HbmBuffers hb,
                                    int grad) {
    const int64_t T = (int64_t)1 << (n - kMeasureBits);
    for (int64_t sl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; sl < nb; sl += (int64_t)gridDim.x * blockDim.x) {
        float e = 0.f;
        for (int64_t t = 0; t < T; ++t) e += hb.epart[sl * T + t];
        const int64_t b = b0 + sl;
        p.out[b] = e;
        if (grad) {
            float g;
            if (p.target) {
                g = p.gscale * (e + (p.bias ? p.bias[0] : 0.f) - p.target[b]);
                if (p.gbuf) p.gbuf[b] = g;
            } else {
                g = p.gout[b];
            }
            hb.gval[sl] = g;
        }
    }
}

// dL/dx[b, col] = r(col) . sum over tiles of the per-tile moments (fixed order)
__global__ void hea_hbm_gx_kernel(const HeaParams<float> p, int n, int tiles_log2, int64_t b0, int64_t nb,
                                  const HbmBuffers hb) {
    const int64_t T = (int64_t)1 << tiles_log2;
    const int64_t cols = (int64_t)n * p.K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nb * cols; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t sl = i / cols, col = i % cols;
        float m[3] = {0.f, 0.f, 0.f};
        for (int64_t t = 0; t < T; ++t) {
            const float* src = hb.mxpart + (sl * T + t) * (3 * cols) + col * 3;
            m[0] += src[0]; m[1] += src[1]; m[2] += src[2];
        }
        const int k = (int)(col / n), q = (int)(col % n);
        int s0 = 0;                                       // first sublayer of block k
        for (int kk = 0; kk < k; ++kk) s0 += p.depth[kk];
        const Vec4<float> rc = ldg4(p.rcoef + (int64_t)s0 * n + q);
        p.gx[(b0 + sl) * p.ldgx + col] = fmaf(rc.z, m[2], fmaf(rc.y, m[1], rc.x * m[0]));
    }
}

}  // namespace qon
