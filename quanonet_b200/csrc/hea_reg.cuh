// Register-resident tier: 2^LQ lanes own one sample's state, 2^NL amplitudes per lane (n = NL + LQ).
//
//   LQ == 0 : one THREAD owns a whole sample (n <= 5 in fp32, n <= 4 in fp64).  Every gate is
//             in-lane FMAs; the CNOT ring is a compile-time register renaming (zero instructions).
//   LQ  > 0 : gates on the low NL qubits stay in-lane; the top LQ qubits pair lanes with
//             __shfl_xor; CNOTs become renames / predicated selects / lane permutations.
//
// Work per sample (reference: core/quantum_circuits_tq.py:65-127 forward; backward = what
// loss.backward() at solvers/solver_pt.py:235 produces, computed here by adjoint differentiation):
//   forward : psi <- prod_k [ sublayers_k * RXlayer_k ] |0>,  E = <psi|H|psi>
//   reverse : lam = g * H psi ; walk the sublayers backwards; for every fused single-qubit group
//             measure the Pauli moments m_P = Im<lam|P_q|psi> (P = X,Y,Z), then un-apply the group
//             on both psi and lam.  All parameter gradients of the group are linear combinations
//             of (m_X, m_Y, m_Z): per-sample dL/dx is formed in-kernel, the shared-parameter
//             moments are summed over the batch (warp butterfly -> per-warp partial row) and
//             turned into dL/dw by the finalize kernel.
//
// Gate fusion: RY(c)RZ(b)RY(a) is one SU(2) matrix [[al,-conj(be)],[be,conj(al)]] from the prep
// table; in the first sublayer of a block the per-sample RX(theta) is folded in
// (al' = al*c + i s conj(be), be' = be*c - i s conj(al)), so a block of depth d costs d*n fused
// gates of 16 FP32 instructions per amplitude pair instead of (3d+1)*n gates of 8.
#pragma once
#include <utility>
#include "hea_common.cuh"

namespace qon {

template <int I> struct IntC { static constexpr int value = I; };

template <int... Is, typename F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F&& f) {
    (f(IntC<Is>{}), ...);
}
template <int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<F&&>(f));
}

#define QON_FULL 0xffffffffu

template <typename T> __device__ __forceinline__ T shfl_xor_(T v, int m) { return __shfl_xor_sync(QON_FULL, v, m); }
template <typename T> __device__ __forceinline__ T shfl_idx_(T v, int l) { return __shfl_sync(QON_FULL, v, l); }

// ------------------------------------------------------------------------------------------------
// one fused SU(2) gate on qubit Q;  DAG applies the inverse
// ------------------------------------------------------------------------------------------------
template <typename T, int NL, int Q, bool DAG>
__device__ __forceinline__ void apply_u(T (&re)[1 << NL], T (&im)[1 << NL], T ar, T ai, T br, T bi, int lane) {
    constexpr int NA = 1 << NL;
    if constexpr (Q < NL) {
        constexpr int bit = 1 << Q;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            if (i & bit) continue;
            const int j = i | bit;
            const T x0r = re[i], x0i = im[i], x1r = re[j], x1i = im[j];
            if constexpr (!DAG) {  // a0' = al a0 - conj(be) a1 ; a1' = be a0 + conj(al) a1
                re[i] = fma_(-bi, x1i, fma_(-br, x1r, fma_(-ai, x0i, ar * x0r)));
                im[i] = fma_(bi, x1r, fma_(-br, x1i, fma_(ai, x0r, ar * x0i)));
                re[j] = fma_(ai, x1i, fma_(ar, x1r, fma_(-bi, x0i, br * x0r)));
                im[j] = fma_(-ai, x1r, fma_(ar, x1i, fma_(bi, x0r, br * x0i)));
            } else {               // a0' = conj(al) a0 + conj(be) a1 ; a1' = -be a0 + al a1
                re[i] = fma_(bi, x1i, fma_(br, x1r, fma_(ai, x0i, ar * x0r)));
                im[i] = fma_(-bi, x1r, fma_(br, x1i, fma_(-ai, x0r, ar * x0i)));
                re[j] = fma_(-ai, x1i, fma_(ar, x1r, fma_(bi, x0i, -br * x0r)));
                im[j] = fma_(ai, x1r, fma_(ar, x1i, fma_(-bi, x0r, -br * x0i)));
            }
        }
    } else {
        constexpr int lb = Q - NL;
        const bool hi = (lane >> lb) & 1;
        // mine' = cm * mine + cp * partner
        const T cmr = ar, cmi = (hi != DAG) ? -ai : ai;
        const T cpr = (hi == DAG) ? -br : br, cpi = DAG ? -bi : bi;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const T pr = shfl_xor_(re[i], 1 << lb), pi = shfl_xor_(im[i], 1 << lb);
            const T mr = re[i], mi = im[i];
            re[i] = fma_(-cpi, pi, fma_(cpr, pr, fma_(-cmi, mi, cmr * mr)));
            im[i] = fma_(cpi, pr, fma_(cpr, pi, fma_(cmi, mr, cmr * mi)));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// reverse-sweep group: Pauli moments of (lam, psi) on qubit Q, then un-apply the gate on both
// ------------------------------------------------------------------------------------------------
template <typename T, int NL, int Q>
__device__ __forceinline__ void bwd_group(T (&pr)[1 << NL], T (&pi)[1 << NL], T (&lr)[1 << NL], T (&li)[1 << NL],
                                          T ar, T ai, T br, T bi, int lane, T& mX, T& mY, T& mZ) {
    constexpr int NA = 1 << NL;
    if constexpr (Q < NL) {
        constexpr int bit = 1 << Q;
        T x = 0, y = 0, z = 0;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            if (i & bit) continue;
            const int j = i | bit;
            // m_X += Im(conj(l0) p1) + Im(conj(l1) p0)
            x = fma_(lr[i], pi[j], x); x = fma_(-li[i], pr[j], x);
            x = fma_(lr[j], pi[i], x); x = fma_(-li[j], pr[i], x);
            // m_Y += -Re(conj(l0) p1) + Re(conj(l1) p0)
            y = fma_(-lr[i], pr[j], y); y = fma_(-li[i], pi[j], y);
            y = fma_(lr[j], pr[i], y); y = fma_(li[j], pi[i], y);
            // m_Z += Im(conj(l0) p0) - Im(conj(l1) p1)
            z = fma_(lr[i], pi[i], z); z = fma_(-li[i], pr[i], z);
            z = fma_(-lr[j], pi[j], z); z = fma_(li[j], pr[j], z);
        }
        mX = x; mY = y; mZ = z;
        apply_u<T, NL, Q, true>(pr, pi, ar, ai, br, bi, lane);
        apply_u<T, NL, Q, true>(lr, li, ar, ai, br, bi, lane);
    } else {
        constexpr int lb = Q - NL;
        const bool hi = (lane >> lb) & 1;
        const T cmr = ar, cmi = hi ? ai : -ai;          // DAG coefficients (see apply_u)
        const T cpr = hi ? -br : br, cpi = -bi;
        T x = 0, y = 0, z = 0;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const T qr = shfl_xor_(pr[i], 1 << lb), qi = shfl_xor_(pi[i], 1 << lb);   // partner psi
            const T kr = shfl_xor_(lr[i], 1 << lb), ki = shfl_xor_(li[i], 1 << lb);   // partner lam
            const T mr = pr[i], mi = pi[i], nr = lr[i], ni = li[i];
            x = fma_(nr, qi, x); x = fma_(-ni, qr, x);      // Im(conj(l_mine) p_partner)
            y = fma_(nr, qr, y); y = fma_(ni, qi, y);       // Re(conj(l_mine) p_partner), signed below
            z = fma_(nr, mi, z); z = fma_(-ni, mr, z);      // Im(conj(l_mine) p_mine), signed below
            pr[i] = fma_(-cpi, qi, fma_(cpr, qr, fma_(-cmi, mi, cmr * mr)));
            pi[i] = fma_(cpi, qr, fma_(cpr, qi, fma_(cmi, mr, cmr * mi)));
            lr[i] = fma_(-cpi, ki, fma_(cpr, kr, fma_(-cmi, ni, cmr * nr)));
            li[i] = fma_(cpi, kr, fma_(cpr, ki, fma_(cmi, nr, cmr * ni)));
        }
        mX = x; mY = hi ? y : -y; mZ = hi ? -z : z;
    }
}

// ------------------------------------------------------------------------------------------------
// CNOT(control C, target TG) and the ring
// ------------------------------------------------------------------------------------------------
template <typename T, int NL, int C, int TG>
__device__ __forceinline__ void cnot(T (&re)[1 << NL], T (&im)[1 << NL], int lane) {
    constexpr int NA = 1 << NL;
    if constexpr (C < NL && TG < NL) {          // pure register renaming
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            if (((i >> C) & 1) && !((i >> TG) & 1)) {
                const int j = i | (1 << TG);
                T t = re[i]; re[i] = re[j]; re[j] = t;
                t = im[i]; im[i] = im[j]; im[j] = t;
            }
        }
    } else if constexpr (C >= NL && TG < NL) {  // control in the lane index: predicated swap
        const bool on = (lane >> (C - NL)) & 1;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            if ((i >> TG) & 1) continue;
            const int j = i | (1 << TG);
            const T a = re[i], b = re[j], c = im[i], d = im[j];
            re[i] = on ? b : a; re[j] = on ? a : b;
            im[i] = on ? d : c; im[j] = on ? c : d;
        }
    } else if constexpr (C < NL && TG >= NL) {  // target in the lane index: exchange half the registers
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            if (!((i >> C) & 1)) continue;
            re[i] = shfl_xor_(re[i], 1 << (TG - NL));
            im[i] = shfl_xor_(im[i], 1 << (TG - NL));
        }
    } else {                                     // both in the lane index: lane permutation
        const int src = lane ^ (((lane >> (C - NL)) & 1) << (TG - NL));
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            re[i] = shfl_idx_(re[i], src);
            im[i] = shfl_idx_(im[i], src);
        }
    }
}

template <typename T, int NL, int LQ, bool REVERSE>
__device__ __forceinline__ void cnot_ring(T (&re)[1 << NL], T (&im)[1 << NL], int lane) {
    constexpr int NQ = NL + LQ;
    if constexpr (NQ > 1) {
        static_for<NQ>([&](auto I) {
            constexpr int i = REVERSE ? (NQ - 1 - decltype(I)::value) : decltype(I)::value;
            cnot<T, NL, (i + 1) % NQ, i>(re, im, lane);
        });
    }
}

// ------------------------------------------------------------------------------------------------
// H psi  (pauli 0: diagonal table, 1: offset + coeff sum_q X_q, 2: offset + coeff sum_q Y_q)
// ------------------------------------------------------------------------------------------------
template <typename T, int NL, int LQ>
__device__ __forceinline__ void apply_ham(const HeaParams<T>& p, const T (&pr)[1 << NL], const T (&pi)[1 << NL],
                                          T (&lr)[1 << NL], T (&li)[1 << NL], int lane) {
    constexpr int NA = 1 << NL;
    constexpr int NQ = NL + LQ;
    const int sub = lane & ((1 << LQ) - 1);
    if (p.pauli == 0) {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const T d = __ldg(p.hdiag + ((sub << NL) | i));
            lr[i] = d * pr[i];
            li[i] = d * pi[i];
        }
        return;
    }
    const bool isY = p.pauli == 2;
#pragma unroll
    for (int i = 0; i < NA; ++i) { lr[i] = p.offset * pr[i]; li[i] = p.offset * pi[i]; }
    static_for<NQ>([&](auto Qc) {
        constexpr int Q = decltype(Qc)::value;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            T fr, fi;   // flipped amplitude psi_{k ^ bit}
            bool one;   // bit Q of this amplitude's index
            if constexpr (Q < NL) {
                fr = pr[i ^ (1 << Q)]; fi = pi[i ^ (1 << Q)]; one = (i >> Q) & 1;
            } else {
                fr = shfl_xor_(pr[i], 1 << (Q - NL)); fi = shfl_xor_(pi[i], 1 << (Q - NL));
                one = (lane >> (Q - NL)) & 1;
            }
            if (!isY) {
                lr[i] = fma_(p.coeff, fr, lr[i]); li[i] = fma_(p.coeff, fi, li[i]);
            } else {     // (Y psi)_k = +i psi_flip if bit set else -i psi_flip ; i(a+ib) = -b + ia
                const T sg = one ? p.coeff : -p.coeff;
                lr[i] = fma_(-sg, fi, lr[i]); li[i] = fma_(sg, fr, li[i]);
            }
        }
    });
}

// sum v[0..VP) over the 32 lanes; afterwards lane l with (l & (32/VP - 1)) == 0 holds slot l / (32/VP)
template <typename T, int VP>
__device__ __forceinline__ T butterfly_reduce(T (&v)[VP], int lane) {
    int live = VP;
#pragma unroll
    for (int mask = 16; mask >= 1; mask >>= 1) {
        if (live > 1) {
            const int half = live >> 1;
            const bool hi = lane & mask;
#pragma unroll
            for (int i = 0; i < VP / 2; ++i) {
                if (i < half) {
                    const T send = hi ? v[i] : v[i + half];
                    const T keep = hi ? v[i + half] : v[i];
                    v[i] = keep + shfl_xor_(send, mask);
                }
            }
            live = half;
        } else {
            v[0] += shfl_xor_(v[0], mask);
        }
    }
    return v[0];
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <typename T, int NL, int LQ, bool GRAD, bool NEED_GX, int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) hea_reg_kernel(const HeaParams<T> p) {
    constexpr int NA = 1 << NL;
    constexpr int NQ = NL + LQ;
    constexpr int SPW = 32 >> LQ;                 // samples per warp
    constexpr int VP = moment_slots(NQ);
    constexpr int WARPS = THREADS / 32;
    static_assert(VP <= 32, "moment butterfly needs 3n <= 32");

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane & ((1 << LQ) - 1);
    const int sidx = lane >> LQ;
    const int64_t gwarp = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t nwarps = (int64_t)gridDim.x * WARPS;
    const int64_t ntiles = (p.B + SPW - 1) / SPW;
    T* mrow = GRAD ? p.mpart + gwarp * (int64_t)p.S * VP : nullptr;

    for (int64_t tile = gwarp; tile < ntiles; tile += nwarps) {
        const int64_t b = tile * SPW + sidx;
        const bool valid = b < p.B;
        const T* xrow = p.x + (valid ? b : p.B - 1) * p.ldx;

        T pr[NA], pi[NA];
#pragma unroll
        for (int i = 0; i < NA; ++i) { pr[i] = 0; pi[i] = 0; }
        pr[0] = sub == 0 ? T(1) : T(0);

        // ---------------- forward sweep ----------------
        {
            int s = 0;
            T th[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) th[q] = __ldg(xrow + q);
            for (int k = 0; k < p.K; ++k) {
                T thn[NQ];
                const int kn = k + 1 < p.K ? k + 1 : k;
#pragma unroll
                for (int q = 0; q < NQ; ++q) thn[q] = __ldg(xrow + (int64_t)kn * NQ + q);   // prefetch next block
                const int d = __ldg(p.depth + k);
                {   // first sublayer, RX folded in
                    const Vec4<T>* uc = p.ucoef + (int64_t)s * NQ;
                    static_for<NQ>([&](auto Qc) {
                        constexpr int Q = decltype(Qc)::value;
                        const Vec4<T> u = ldg4(uc + Q);
                        T sn, cs;
                        sincos_half(th[Q], sn, cs);
                        const T ar = fma_(sn, u.w, u.x * cs), ai = fma_(sn, u.z, u.y * cs);
                        const T br = fma_(-sn, u.y, u.z * cs), bi = fma_(-sn, u.x, u.w * cs);
                        apply_u<T, NL, Q, false>(pr, pi, ar, ai, br, bi, lane);
                    });
                    cnot_ring<T, NL, LQ, false>(pr, pi, lane);
                    ++s;
                }
                for (int j = 1; j < d; ++j) {
                    const Vec4<T>* uc = p.ucoef + (int64_t)s * NQ;
                    static_for<NQ>([&](auto Qc) {
                        constexpr int Q = decltype(Qc)::value;
                        const Vec4<T> u = ldg4(uc + Q);
                        apply_u<T, NL, Q, false>(pr, pi, u.x, u.y, u.z, u.w, lane);
                    });
                    cnot_ring<T, NL, LQ, false>(pr, pi, lane);
                    ++s;
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q) th[q] = thn[q];
            }
        }

        // ---------------- expectation value ----------------
        T lr[NA], li[NA];
        apply_ham<T, NL, LQ>(p, pr, pi, lr, li, lane);
        T e = 0;
#pragma unroll
        for (int i = 0; i < NA; ++i) { e = fma_(pr[i], lr[i], e); e = fma_(pi[i], li[i], e); }
#pragma unroll
        for (int m = 1; m < (1 << LQ); m <<= 1) e += shfl_xor_(e, m);
        if (valid && sub == 0) p.out[b] = e;

        if constexpr (GRAD) {
            // ---------------- reverse (adjoint) sweep ----------------
            const T g = valid ? __ldg(p.gout + b) : T(0);
#pragma unroll
            for (int i = 0; i < NA; ++i) { lr[i] *= g; li[i] *= g; }
            T* gxrow = NEED_GX ? p.gx + (valid ? b : 0) * p.ldgx : nullptr;
            int s = p.S;
            T th[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) th[q] = __ldg(xrow + (int64_t)(p.K - 1) * NQ + q);
            for (int k = p.K - 1; k >= 0; --k) {
                T thn[NQ];
                const int kn = k > 0 ? k - 1 : 0;
#pragma unroll
                for (int q = 0; q < NQ; ++q) thn[q] = __ldg(xrow + (int64_t)kn * NQ + q);
                const int d = __ldg(p.depth + k);
                for (int j = d - 1; j >= 1; --j) {
                    --s;
                    const Vec4<T>* uc = p.ucoef + (int64_t)s * NQ;
                    cnot_ring<T, NL, LQ, true>(pr, pi, lane);
                    cnot_ring<T, NL, LQ, true>(lr, li, lane);
                    T mv[VP];
#pragma unroll
                    for (int i = 0; i < VP; ++i) mv[i] = 0;
                    static_for<NQ>([&](auto Qc) {
                        constexpr int Q = NQ - 1 - decltype(Qc)::value;
                        const Vec4<T> u = ldg4(uc + Q);
                        bwd_group<T, NL, Q>(pr, pi, lr, li, u.x, u.y, u.z, u.w, lane, mv[3 * Q], mv[3 * Q + 1], mv[3 * Q + 2]);
                    });
                    const T tot = butterfly_reduce<T, VP>(mv, lane);
                    if ((lane & (32 / VP - 1)) == 0) atomicAdd(mrow + (int64_t)s * VP + lane / (32 / VP), tot);
                }
                {   // first sublayer of the block (RX folded in): also yields dL/dx
                    --s;
                    const Vec4<T>* uc = p.ucoef + (int64_t)s * NQ;
                    const Vec4<T>* rc = p.rcoef + (int64_t)s * NQ;
                    cnot_ring<T, NL, LQ, true>(pr, pi, lane);
                    cnot_ring<T, NL, LQ, true>(lr, li, lane);
                    T mv[VP];
#pragma unroll
                    for (int i = 0; i < VP; ++i) mv[i] = 0;
                    static_for<NQ>([&](auto Qc) {
                        constexpr int Q = NQ - 1 - decltype(Qc)::value;
                        const Vec4<T> u = ldg4(uc + Q);
                        T sn, cs;
                        sincos_half(th[Q], sn, cs);
                        const T ar = fma_(sn, u.w, u.x * cs), ai = fma_(sn, u.z, u.y * cs);
                        const T br = fma_(-sn, u.y, u.z * cs), bi = fma_(-sn, u.x, u.w * cs);
                        bwd_group<T, NL, Q>(pr, pi, lr, li, ar, ai, br, bi, lane, mv[3 * Q], mv[3 * Q + 1], mv[3 * Q + 2]);
                        if constexpr (NEED_GX) {
                            T mx = mv[3 * Q], my = mv[3 * Q + 1], mz = mv[3 * Q + 2];
#pragma unroll
                            for (int m = 1; m < (1 << LQ); m <<= 1) {
                                mx += shfl_xor_(mx, m); my += shfl_xor_(my, m); mz += shfl_xor_(mz, m);
                            }
                            const Vec4<T> r = ldg4(rc + Q);
                            const T gxv = fma_(r.z, mz, fma_(r.y, my, r.x * mx));
                            if (valid && sub == 0) gxrow[(int64_t)k * NQ + Q] = gxv;
                        }
                    });
                    const T tot = butterfly_reduce<T, VP>(mv, lane);
                    if ((lane & (32 / VP - 1)) == 0) atomicAdd(mrow + (int64_t)s * VP + lane / (32 / VP), tot);
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q) th[q] = thn[q];
            }
        }
    }
}

}  // namespace qon
