// Shared-memory tier (fp32, n = 6 ... 13): the state of a sample lives in shared memory; threads sweep it
// in register-blocked PASSES.  In a pass every thread pulls the 32 amplitudes that differ in a window of
// 5 consecutive qubits [lo, lo+5) into registers (a PackedState<5>), applies the fused gates of the
// window's qubits with exactly the FFMA2 code of the register tier (hea_reg.cuh), and writes them back.
// A sublayer of an n-qubit circuit is ceil(n/5) passes, so shared-memory traffic is 2 x 8 B per
// amplitude per 5 gates (FP : LDS/STS instruction ratio ~10 : 1) and the FP32 pipe stays the limiter.
//
//  * The window position is a run-time value, so ONE pass body serves every full window of every qubit
//    count; the per-register address offsets come from a constexpr-built table in constant memory and an
//    address is one LOP3 (XOR) away: slot(k) = k ^ ((k >> 5) & 31) is GF(2)-linear, hence
//    slot(base | i << lo) = slot(base) ^ slot(i << lo).  The XOR swizzle makes both the low window
//    (lanes differ in bits >= 5) and the high windows (lanes differ in bits < 5) bank-conflict free;
//    regions are aligned to their size so the region base folds into the same XOR.
//  * When n is not a multiple of 5 the last window overlaps the previous one and only its top GL register
//    bits carry gates; GL is a template parameter (a second, smaller pass body) so that no gate sits
//    behind a run-time branch — a conditional gate costs 64 MOVs at the join (measured: 14 % of samples).
//  * The CNOT ring is a GF(2)-linear index permutation: the last pass of a sublayer stores to
//    slot(ring(k)) (again base ^ table[i]); the reverse sweep's first pass loads from there.
//  * Reverse sweep: psi and lam both in shared memory, 64 + 64 register pairs per thread per pass, same
//    moments / butterfly / finalize machinery as the register tier.
//
// Capacity: 2 x 2^n x 8 B <= 128 KB (+ one region of alignment slack) => n <= 13.
// Reference semantics: core/quantum_circuits_tq.py:65-127.
#pragma once
#include "hea_reg.cuh"

namespace qon {

__host__ __device__ constexpr int smem_swz(int k) { return k ^ ((k >> 5) & 31); }
__host__ __device__ constexpr int smem_ring(int k, int n) {
    for (int i = 0; i < n; ++i) k ^= ((k >> ((i + 1) % n)) & 1) << i;   // CNOT control (i+1)%n -> target i
    return k;
}
__host__ __device__ constexpr int smem_passes(int n) { return (n + kSmemW - 1) / kSmemW; }
__host__ __device__ constexpr int smem_lo(int n, int p) { return p * kSmemW + kSmemW <= n ? p * kSmemW : n - kSmemW; }
// number of gated register bits of the last window (its top bits); 5 when n is a multiple of 5
__host__ __device__ constexpr int smem_last_gates(int n) { return n % kSmemW == 0 ? kSmemW : n % kSmemW; }

struct SmemTables {
    int off[kSmemMaxN + 1][kSmemMaxP][32];    // 8 * slot(i << lo)            (byte offsets)
    int poff[kSmemMaxN + 1][32];              // 8 * slot(ring(i << lo_last)) (ring-permuted, last pass)
};
constexpr SmemTables make_smem_tables() {
    SmemTables t{};
    for (int n = kSmemMinN; n <= kSmemMaxN; ++n) {
        for (int p = 0; p < smem_passes(n); ++p)
            for (int i = 0; i < 32; ++i) t.off[n][p][i] = 8 * smem_swz(i << smem_lo(n, p));
        for (int i = 0; i < 32; ++i) t.poff[n][i] = 8 * smem_swz(smem_ring(i << smem_lo(n, smem_passes(n) - 1), n));
    }
    return t;
}
__constant__ SmemTables c_smem_tbl = make_smem_tables();

__device__ __forceinline__ u64 lds64(unsigned addr) {
    u64 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(unsigned addr, u64 v) {
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

using SmemState = PackedState<5>;

__device__ __forceinline__ void smem_load(SmemState& st, unsigned base, const int* off) {
#pragma unroll
    for (int i = 0; i < 32; ++i) st.a[i] = lds64(base ^ (unsigned)off[i]);
}
__device__ __forceinline__ void smem_store(const SmemState& st, unsigned base, const int* off) {
#pragma unroll
    for (int i = 0; i < 32; ++i) sts64(base ^ (unsigned)off[i], st.a[i]);
}

// gates of one window on a register-resident group: register bits [5-G, 5) carry qubits lo+5-G .. lo+4
template <int G>
__device__ __forceinline__ void smem_fwd_gates(SmemState& st, const HeaParams<float>& p, int n, int s, int lo, bool fold,
                                               const float* xk, int lane) {
    float th[G];
#pragma unroll
    for (int r = 0; r < G; ++r) th[r] = fold ? __ldg(xk + lo + (kSmemW - G) + r) : 0.f;
    static_for<G>([&](auto Rc) {
        constexpr int R = kSmemW - G + decltype(Rc)::value;
        const Vec4<float> u = ldg4(p.ucoef + (int64_t)s * n + lo + R);
        float ar = u.x, ai = u.y, br = u.z, bi = u.w;
        if (fold) fold_rx_coef(u, th[R - (kSmemW - G)], ar, ai, br, bi);
        apply_u<R, false>(st, ar, ai, br, bi, lane);
    });
}

// reverse: moments + un-apply; mv[3*(R-(5-G)) + {0,1,2}] = (mX, mY, mZ) of register bit R
template <int G>
__device__ __forceinline__ void smem_bwd_gates(SmemState& st, SmemState& lm, const HeaParams<float>& p, int n, int s,
                                               int lo, bool fold, const float* xk, int lane, float (&mv)[3 * G]) {
    float th[G];
#pragma unroll
    for (int r = 0; r < G; ++r) th[r] = fold ? __ldg(xk + lo + (kSmemW - G) + r) : 0.f;
    static_for<G>([&](auto Rc) {
        constexpr int I = G - 1 - decltype(Rc)::value;
        constexpr int R = kSmemW - G + I;
        const Vec4<float> u = ldg4(p.ucoef + (int64_t)s * n + lo + R);
        float ar = u.x, ai = u.y, br = u.z, bi = u.w;
        if (fold) fold_rx_coef(u, th[I], ar, ai, br, bi);
        bwd_group<R>(st, lm, ar, ai, br, bi, lane, mv[3 * I], mv[3 * I + 1], mv[3 * I + 2]);
    });
}

// GL = gated bits of the last window (1..5).  Full windows use G = 5.
template <bool GRAD, bool NEED_GX, int GL, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) hea_smem_kernel(const HeaParams<float> p, const SmemGeom geo) {
    using State = SmemState;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float s_red[THREADS / 32][16];          // cross-warp partials (per-sample reductions, n >= 11)
    constexpr int WARPS = THREADS / 32;
    const int n = geo.n, P = geo.P, VP = geo.vp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tps = 1 << geo.tps_log2;
    const int s_local = tid >> geo.tps_log2;            // sample slot of this thread within the CTA
    const int t_in = tid & (tps - 1);                   // thread index within the sample = group index
    // region bases (shared-space byte addresses), aligned to the region size so that addr = base ^ offset
    const unsigned raw = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned rb = (unsigned)geo.region_bytes;
    const unsigned aligned = (raw + rb - 1) & ~(rb - 1);
    const unsigned sxor = (unsigned)((s_local << geo.tps_log2) & 31) * 8u;   // per-sample bank shift (n < 10)
    const unsigned psi_base = (aligned + (unsigned)s_local * rb) ^ sxor;
    const unsigned lam_base = psi_base + (unsigned)geo.spc * rb;
    const int64_t gwarp = (int64_t)blockIdx.x * WARPS + warp;
    float* mrow = GRAD ? p.mpart + gwarp * p.rowlen : nullptr;

    // byte offset of this thread's group for a window starting at lo (amplitude bits outside the window)
    auto group_base = [&](int lo) -> unsigned {
        const int kb = ((t_in >> lo) << (lo + kSmemW)) | (t_in & ((1 << lo) - 1));
        return (unsigned)(8 * smem_swz(kb));
    };
    auto group_base_ring = [&](int lo) -> unsigned {
        int kb = ((t_in >> lo) << (lo + kSmemW)) | (t_in & ((1 << lo) - 1));
        for (int i = 0; i < n; ++i) kb ^= ((kb >> (i + 1 == n ? 0 : i + 1)) & 1) << i;
        return (unsigned)(8 * smem_swz(kb));
    };
    // The forward kernel lets warps run free when a sample fits in one warp (n <= 10).  With gradients the
    // CTA stays in lock-step: the reverse pass body is ~40 KB of SASS and drifting warps thrash the
    // instruction cache (measured: 41.9 -> 37.5 TFLOP/s at n = 10 with warp-level sync).
    const bool warp_local = !GRAD && tps <= 32;
    auto sync_sample = [&]() {
        if (warp_local) __syncwarp();
        else __syncthreads();
    };
    // moments of one pass: warp butterfly over 16 slots (3*G used), RED into this warp's partial row
    auto reduce_moments = [&](const float* mvp, int g, int s, int q0) {
        float bv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) bv[i] = i < 3 * g ? mvp[i < 3 * g ? i : 0] : 0.f;
        const float tot = butterfly_reduce<float, 16>(bv, lane);
        if ((lane & 1) == 0 && (lane >> 1) < 3 * g) atomicAdd(mrow + (int64_t)s * VP + 3 * q0 + (lane >> 1), tot);
    };

    const int64_t nrounds = (p.B + geo.spc - 1) / geo.spc;
    for (int64_t round = blockIdx.x; round < nrounds; round += gridDim.x) {
        const int64_t b = round * geo.spc + s_local;
        const bool valid = b < p.B;
        const float* xrow = p.x + (valid ? b : p.B - 1) * p.ldx;

        // ---------------- |0...0> ----------------
        {
            const unsigned gb = psi_base ^ group_base(0);
#pragma unroll
            for (int i = 0; i < 32; ++i) sts64(gb ^ (unsigned)c_smem_tbl.off[n][0][i], 0ull);
        }
        sync_sample();
        if (t_in == 0) sts64(psi_base, pack2(1.f, 0.f));     // slot(0) = 0
        sync_sample();

        // ---------------- forward sweep ----------------
        int s = 0;
        for (int k = 0; k < p.K; ++k) {
            const int d = __ldg(p.depth + k);
            const float* xk = xrow + (int64_t)k * n;
#pragma unroll 1
            for (int j = 0; j < d; ++j, ++s) {
                auto fwd_pass = [&](auto Gc, int ps, bool ring_store) {
                    constexpr int G = decltype(Gc)::value;
                    const int lo = geo.lo[ps];
                    State st;
                    smem_load(st, psi_base ^ group_base(lo), c_smem_tbl.off[n][ps]);
                    smem_fwd_gates<G>(st, p, n, s, lo, j == 0, xk, lane);
                    if (ring_store) {
                        sync_sample();       // all loads of this pass are done before ring-permuted stores land
                        smem_store(st, psi_base ^ group_base_ring(lo), c_smem_tbl.poff[n]);
                    } else {
                        smem_store(st, psi_base ^ group_base(lo), c_smem_tbl.off[n][ps]);
                    }
                    sync_sample();
                };
                constexpr bool kPartial = GL != kSmemW;
                const int PF = kPartial ? P - 1 : P;          // passes over full windows
#pragma unroll 1
                for (int ps = 0; ps < PF; ++ps) fwd_pass(IntC<kSmemW>{}, ps, !kPartial && ps == P - 1);
                if constexpr (kPartial) fwd_pass(IntC<GL>{}, P - 1, true);
            }
        }

        // ---------------- expectation value: window 0, lam = H psi ----------------
        float e;
        {
            const unsigned gb = group_base(0);
            State st, hm;
            smem_load(st, psi_base ^ gb, c_smem_tbl.off[n][0]);
            if (p.pauli == 0) {
#pragma unroll
                for (int i = 0; i < 32; ++i) hm.a[i] = mul2<0>(__ldg(p.hdiag + ((t_in << kSmemW) | i)), st.a[i]);
            } else {
                const bool isY = p.pauli == 2;
#pragma unroll
                for (int i = 0; i < 32; ++i) hm.a[i] = mul2<0>(p.offset, st.a[i]);
                for (int q = 0; q < n; ++q) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int kk = (t_in << kSmemW) | i;
                        const u64 f = lds64(psi_base ^ (unsigned)(8 * smem_swz(kk ^ (1 << q))));
                        if (!isY) hm.a[i] = fma2<0>(p.coeff, f, hm.a[i]);
                        else hm.a[i] = fma2<2>(((kk >> q) & 1) ? p.coeff : -p.coeff, f, hm.a[i]);
                    }
                }
            }
            e = real_dot(st, hm);
            if constexpr (GRAD) smem_store(hm, lam_base ^ gb, c_smem_tbl.off[n][0]);
        }
        // sum over the threads of the sample: lanes first, then warps
        for (int m = 1; m < tps && m < 32; m <<= 1) e += shfl_xor_(e, m);
        if (tps > 32) {
            if (lane == 0) s_red[warp][0] = e;
            __syncthreads();
            float t = 0.f;
            const int w0 = (s_local << geo.tps_log2) >> 5, nw = tps >> 5;
            for (int w = 0; w < nw; ++w) t += s_red[w0 + w][0];
            e = t;
            __syncthreads();
        }
        if (valid && t_in == 0) p.out[b] = e;

        if constexpr (GRAD) {
            float g = 0.f;
            if (valid) {
                if (p.target) {
                    g = p.gscale * (e + (p.bias ? __ldg(p.bias) : 0.f) - __ldg(p.target + b));
                    if (t_in == 0 && p.gbuf) p.gbuf[b] = g;
                } else {
                    g = __ldg(p.gout + b);
                }
            }
            {   // lam <- g * H psi (each thread rescales the 32 values it just wrote)
                const unsigned gb = lam_base ^ group_base(0);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const unsigned a = gb ^ (unsigned)c_smem_tbl.off[n][0][i];
                    sts64(a, mul2<0>(g, lds64(a)));
                }
            }
            sync_sample();

            // ---------------- reverse (adjoint) sweep ----------------
            float* gxrow = NEED_GX ? p.gx + (valid ? b : 0) * p.ldgx : nullptr;
            s = p.S;
            for (int k = p.K - 1; k >= 0; --k) {
                const int d = __ldg(p.depth + k);
                const float* xk = xrow + (int64_t)k * n;
#pragma unroll 1
                for (int j = d - 1; j >= 0; --j) {
                    --s;
                    auto bwd_pass = [&](auto Gc, int ps, bool ring_load) {
                        constexpr int G = decltype(Gc)::value;
                        const int lo = geo.lo[ps];
                        const int q0 = lo + kSmemW - G;                           // first gated qubit of this pass
                        State st, lm;
                        const unsigned gbl = ring_load ? group_base_ring(lo) : group_base(lo);   // undo the CNOT ring
                        const int* offl = ring_load ? c_smem_tbl.poff[n] : c_smem_tbl.off[n][ps];
                        smem_load(st, psi_base ^ gbl, offl);
                        smem_load(lm, lam_base ^ gbl, offl);
                        if (ring_load) sync_sample();               // everyone has loaded before plain stores land
                        float mv[3 * G];
                        smem_bwd_gates<G>(st, lm, p, n, s, lo, j == 0, xk, lane, mv);
                        const unsigned gbs = group_base(lo);
                        smem_store(st, psi_base ^ gbs, c_smem_tbl.off[n][ps]);
                        smem_store(lm, lam_base ^ gbs, c_smem_tbl.off[n][ps]);
                        // per-sample dL/dx of the folded RX gates of this window
                        if (NEED_GX && j == 0) {
#pragma unroll
                            for (int I = 0; I < G; ++I) {
                                float mx = mv[3 * I], my = mv[3 * I + 1], mz = mv[3 * I + 2];
                                for (int m = 1; m < tps && m < 32; m <<= 1) {
                                    mx += shfl_xor_(mx, m); my += shfl_xor_(my, m); mz += shfl_xor_(mz, m);
                                }
                                if (tps > 32) {
                                    if (lane == 0) {
                                        s_red[warp][3 * I] = mx; s_red[warp][3 * I + 1] = my; s_red[warp][3 * I + 2] = mz;
                                    }
                                } else {
                                    const Vec4<float> rc = ldg4(p.rcoef + (int64_t)s * n + q0 + I);
                                    if (valid && t_in == 0)
                                        gxrow[(int64_t)k * n + q0 + I] = fmaf(rc.z, mz, fmaf(rc.y, my, rc.x * mx));
                                }
                            }
                        }
                        reduce_moments(mv, G, s, q0);
                        sync_sample();
                        if (NEED_GX && j == 0 && tps > 32) {
                            // cross-warp part of the per-sample reduction (n >= 11: one to four samples per CTA)
                            if (t_in < G) {
                                const int I = t_in;
                                const int w0 = (s_local << geo.tps_log2) >> 5, nw = tps >> 5;
                                float mx = 0.f, my = 0.f, mz = 0.f;
                                for (int w = 0; w < nw; ++w) {
                                    mx += s_red[w0 + w][3 * I]; my += s_red[w0 + w][3 * I + 1]; mz += s_red[w0 + w][3 * I + 2];
                                }
                                const Vec4<float> rc = ldg4(p.rcoef + (int64_t)s * n + q0 + I);
                                if (valid) gxrow[(int64_t)k * n + q0 + I] = fmaf(rc.z, mz, fmaf(rc.y, my, rc.x * mx));
                            }
                            __syncthreads();
                        }
                    };
                    constexpr bool kPartial = GL != kSmemW;
                    const int PF = kPartial ? P - 1 : P;
                    if constexpr (kPartial) bwd_pass(IntC<GL>{}, P - 1, true);
#pragma unroll 1
                    for (int ps = PF - 1; ps >= 0; --ps) bwd_pass(IntC<kSmemW>{}, ps, !kPartial && ps == P - 1);
                }
            }
        }
        sync_sample();
    }
}

}  // namespace qon
