// Small-batch latency tier for n = 6..10 (fp32; fp64 up to n = 9; angles given): ONE WARP PER SAMPLE, 2^(n-5)
// amplitudes per lane.
//
// Same idea as hea_warp.cuh, one step wider: the low NL = n-5 qubits index the registers of a lane, the top 5
// index the lanes.  In-lane gates are plain FMAs on register pairs, lane gates exchange every register with
// __shfl_xor; in the CNOT ring the in-lane CNOTs are register renames, the control-in-lane-index one is a
// predicated swap, the run of lane-to-lane CNOTs is ONE composed lane permutation, and the closing CNOT
// (control qubit 0, target qubit n-1) exchanges half the registers with the partner lane.  Gate arithmetic and
// the adjoint sweep are the register tier's (hea_reg.cuh: apply_u / bwd_group / cnot on ScalarState); tables
// and the sample's sin/cos live in shared memory as in hea_warp.cuh.  Without this tier a 100-sample batch
// at n = 8 occupies 4 CTAs of the shared-memory tier and takes ~0.9 ms per fwd+grad call.
#pragma once
#include "hea_warp.cuh"

namespace qon {

//   ENC as in hea_warp.cuh: 0 angles given, 1 frequency layers evaluated in-kernel, 2 also their gradients.
template <typename T, int N, bool GRAD, bool NEED_GX, int ENC, int THREADS>
__global__ void __launch_bounds__(THREADS) hea_warp_wide_kernel(const HeaParams<T> p, const DepthPack dp) {
    constexpr bool FREQ_GRAD = GRAD && ENC == 2;
    constexpr bool WANT_GX = NEED_GX || FREQ_GRAD;
    static_assert(!(NEED_GX && ENC != 0), "grad_x is only materialised when x is");
    constexpr int LQ = 5, NL = N - LQ, NA = 1 << NL;
    constexpr int VP = moment_slots(N);          // 32 for n = 6..10: after the butterfly lane l holds slot l
    constexpr int FVP = freq_slots(N);
    constexpr int WARPS = THREADS / 32;
    static_assert(N >= 6 && N <= 10 && VP == 32, "wide latency tier covers n = 6..10");
    using State = ScalarState<T, NL>;
    extern __shared__ __align__(32) unsigned char warp_smem[];

    const int S = p.S, K = p.K, SN = p.S * N, KN = p.K * N;
    Vec4<T>* uc_s = reinterpret_cast<Vec4<T>*>(warp_smem);
    Vec4<T>* rc_s = uc_s + SN;
    Vec2<T>* sc_all = reinterpret_cast<Vec2<T>*>(uc_s + (WANT_GX ? 2 : 1) * SN);
    T* uv_all = reinterpret_cast<T*>(sc_all + (size_t)WARPS * KN);          // FREQ_GRAD only
    for (int i = threadIdx.x; i < SN; i += THREADS) {
        uc_s[i] = ldg4(p.ucoef + i);
        if constexpr (WANT_GX) rc_s[i] = ldg4(p.rcoef + i);
    }
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // lane-to-lane CNOTs of the ring: control lane bit j+1 -> target lane bit j, j = 0..LQ-2, composed
    int pf = lane, pr = lane;
#pragma unroll
    for (int j = LQ - 2; j >= 0; --j) pf ^= ((pf >> (j + 1)) & 1) << j;     // forward: gather source
#pragma unroll
    for (int j = 0; j <= LQ - 2; ++j) pr ^= ((pr >> (j + 1)) & 1) << j;     // reverse: gather source
    Vec2<T>* sc = sc_all + (size_t)warp * KN;
    T* uv = FREQ_GRAD ? uv_all + (size_t)warp * KN : nullptr;    // source value u of every encoding column

    const int64_t gwarp = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t nwarps = (int64_t)gridDim.x * WARPS;
    T* mrow = GRAD ? p.mpart + gwarp * p.rowlen : nullptr;
    T* frow = GRAD ? mrow + (int64_t)S * VP : nullptr;
    T* srow = GRAD ? frow + (int64_t)K * FVP : nullptr;
    __syncthreads();

    // ring = CNOT(control (i+1)%N -> target i), i = 0..N-1, applied in that order (REVERSE: undone in reverse)
    auto ring = [&](auto Rev, State& st) {
        constexpr bool REV = decltype(Rev)::value;
        auto low = [&]() {          // i = 0..NL-1: both in-lane (renames), then control = lane bit 0 (predicated swap)
            static_for<NL>([&](auto I) {
                constexpr int i = REV ? NL - 1 - decltype(I)::value : decltype(I)::value;
                cnot<i + 1, i>(st, lane);
            });
        };
        auto mid = [&]() {          // i = NL..N-2: one composed lane permutation
            const int src = REV ? pr : pf;
#pragma unroll
            for (int r = 0; r < NA; ++r) { st.re[r] = shfl_idx_(st.re[r], src); st.im[r] = shfl_idx_(st.im[r], src); }
        };
        auto top = [&]() { cnot<0, N - 1>(st, lane); };   // i = N-1: control qubit 0 (in-lane), target = top lane bit
        if constexpr (!REV) { low(); mid(); top(); }
        else { top(); mid(); low(); }
    };
    auto coef = [&](int s, int Q, int k, bool fold, T& ar, T& ai, T& br, T& bi) {
        const Vec4<T> u = uc_s[s * N + Q];
        ar = u.x; ai = u.y; br = u.z; bi = u.w;
        if (fold) {
            const Vec2<T> t = sc[k * N + Q];
            ar = fma_(t.x, u.w, u.x * t.y); ai = fma_(t.x, u.z, u.y * t.y);
            br = fma_(-t.x, u.y, u.z * t.y); bi = fma_(-t.x, u.x, u.w * t.y);
        }
    };

    for (int64_t tile0 = (int64_t)blockIdx.x * WARPS; tile0 < p.B; tile0 += nwarps) {
        const int64_t b = tile0 + warp;
        const bool valid = b < p.B;
        const int64_t bc = valid ? b : p.B - 1;
        __syncwarp();
        if constexpr (ENC == 0) {
            const T* xrow = p.x + bc * p.ldx;
#pragma unroll 4
            for (int c = lane; c < KN; c += 32) {
                T sn, cs;
                sincos_half(__ldg(xrow + c), sn, cs);
                sc[c] = Vec2<T>{sn, cs};
            }
        } else {
            const T* u0row = p.u0 ? p.u0 + bc * p.ldu0 : nullptr;
            const T* u1row = p.u1 + bc * p.ldu1;
            const int c0 = p.K0 * N;                       // columns below c0 read source 0
#pragma unroll 4
            for (int c = lane; c < KN; c += 32) {
                const T u = __ldg((c < c0 ? u0row : u1row) + __ldg(p.uidx + c));
                T sn, cs;
                sincos_half(fma_(u, __ldg(p.fw + c), p.fb ? __ldg(p.fb + c) : T(0)), sn, cs);
                sc[c] = Vec2<T>{sn, cs};
                if constexpr (FREQ_GRAD) uv[c] = u;
            }
        }
        __syncwarp();

        State ps;
        init_zero_state(ps, lane == 0);
        {
            int s = 0;
            for (int k = 0; k < K; ++k) {
                const int d = dp.d[k];
#pragma unroll 1
                for (int j = 0; j < d; ++j, ++s) {
                    static_for<N>([&](auto Qc) {
                        constexpr int Q = decltype(Qc)::value;
                        T ar, ai, br, bi;
                        coef(s, Q, k, j == 0, ar, ai, br, bi);
                        apply_u<Q, false>(ps, ar, ai, br, bi, lane);
                    });
                    ring(IntC<0>{}, ps);
                }
            }
        }

        State lm;
        apply_ham<LQ>(p, ps, lm, lane);
        T e = real_dot(ps, lm);
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) e += shfl_xor_(e, m);
        if (valid && lane == 0 && p.out) p.out[b] = e;

        if constexpr (GRAD) {
            T g = T(0);
            if (p.target) {   // fused MSE: g = dL/dout for L = gscale/2 * sum (out + bias - y)^2
                T resid = T(0);
                if (valid) {
                    resid = e + (p.bias ? __ldg(p.bias) : T(0)) - __ldg(p.target + b);
                    g = p.gscale * resid;
                    if (lane == 0 && p.gbuf) p.gbuf[b] = g;
                }
                if (lane == 0) { atomicAdd(srow, g); atomicAdd(srow + 1, resid * resid); }
            } else if (valid) {
                g = __ldg(p.gout + b);
            }
            scale_state(lm, g);
            T* gxrow = NEED_GX ? p.gx + (valid ? b : 0) * p.ldgx : nullptr;
            int s = S;
            for (int k = K - 1; k >= 0; --k) {
                const int d = dp.d[k];
#pragma unroll 1
                for (int j = d - 1; j >= 0; --j) {
                    --s;
                    ring(IntC<1>{}, ps);
                    ring(IntC<1>{}, lm);
                    T mv[VP];
#pragma unroll
                    for (int i = 3 * N; i < VP; ++i) mv[i] = T(0);
                    static_for<N>([&](auto Qc) {
                        constexpr int Q = N - 1 - decltype(Qc)::value;
                        T ar, ai, br, bi;
                        coef(s, Q, k, j == 0, ar, ai, br, bi);
                        bwd_group<Q>(ps, lm, ar, ai, br, bi, lane, mv[3 * Q], mv[3 * Q + 1], mv[3 * Q + 2]);
                    });
                    const T tot = butterfly_reduce<T, VP>(mv, lane);      // lane l holds slot l
                    atomicAdd(mrow + (int64_t)s * VP + lane, tot);
                    if constexpr (WANT_GX) {
                        if (j == 0) {   // the warp is the sample: the totals are its own moments
                            const int q3 = (lane < N ? lane : 0) * 3;
                            const T mx = shfl_idx_(tot, q3), my = shfl_idx_(tot, q3 + 1), mz = shfl_idx_(tot, q3 + 2);
                            if (lane < N) {
                                const Vec4<T> r = rc_s[s * N + lane];
                                const T gxv = fma_(r.z, mz, fma_(r.y, my, r.x * mx));
                                if constexpr (NEED_GX) {
                                    if (valid) gxrow[(int64_t)k * N + lane] = gxv;
                                }
                                if constexpr (FREQ_GRAD) {   // theta = fw*u + fb  =>  d/dfw = gx*u, d/dfb = gx (g = 0 if invalid)
                                    atomicAdd(frow + k * FVP + 2 * lane, gxv * uv[k * N + lane]);
                                    atomicAdd(frow + k * FVP + 2 * lane + 1, gxv);
                                }
                            }
                        }
                    }
                }
            }
        }
    }
}

}  // namespace qon
