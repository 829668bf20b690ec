// Register-resident tier: 2^LQ lanes own one sample's state, 2^NL amplitudes per lane (n = NL + LQ).
//
//   LQ == 0 : one THREAD owns a whole sample (n <= 5 in fp32, n <= 4 in fp64).  Every gate is
//             in-lane FMAs; the CNOT ring is a compile-time register renaming (zero instructions).
//             In fp32 each complex amplitude is one aligned 64-bit register pair and every gate is
//             issued as Blackwell packed-FP32 FFMA2 (fma.rn.f32x2): ptxas folds the (re,im) swap,
//             the per-half sign and the scalar-coefficient broadcast into FFMA2 operand modifiers
//             (.LO_HI / .NP / .F32), so a complex multiply-accumulate is 2 instructions, not 4.
//             Half the instruction count halves the SASS footprint (the scalar version ran at a
//             77 % instruction-cache hit rate, see profiles/) and frees issue slots.
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
// gates instead of (3d+1)*n.
#pragma once
#include <type_traits>
#include <utility>
#include "hea_common.cuh"
#include "ffma2.cuh"

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

// =================================================================================================
// Scalar state (fp64, and fp32 when lanes share a sample): separate re / im register arrays
// =================================================================================================
template <typename T, int NL_>
struct ScalarState {
    static constexpr int NL = NL_;
    static constexpr int NA = 1 << NL_;
    T re[NA], im[NA];
};

template <int Q, bool DAG, typename T, int NL>
__device__ __forceinline__ void apply_u(ScalarState<T, NL>& st, T ar, T ai, T br, T bi, int lane) {
    constexpr int NA = 1 << NL;
    T(&re)[NA] = st.re;
    T(&im)[NA] = st.im;
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

// Pauli moments of (lam, psi) on qubit Q (state AFTER the gate), then un-apply the gate on both.
template <int Q, typename T, int NL>
__device__ __forceinline__ void bwd_group(ScalarState<T, NL>& ps, ScalarState<T, NL>& lm, T ar, T ai, T br, T bi,
                                          int lane, T& mX, T& mY, T& mZ) {
    constexpr int NA = 1 << NL;
    T(&pr)[NA] = ps.re;
    T(&pi)[NA] = ps.im;
    T(&lr)[NA] = lm.re;
    T(&li)[NA] = lm.im;
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
        apply_u<Q, true>(ps, ar, ai, br, bi, lane);
        apply_u<Q, true>(lm, ar, ai, br, bi, lane);
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

template <int C, int TG, typename T, int NL>
__device__ __forceinline__ void cnot(ScalarState<T, NL>& st, int lane) {
    constexpr int NA = 1 << NL;
    T(&re)[NA] = st.re;
    T(&im)[NA] = st.im;
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

template <typename T, int NL>
__device__ __forceinline__ void init_zero_state(ScalarState<T, NL>& st, bool owner) {
#pragma unroll
    for (int i = 0; i < (1 << NL); ++i) { st.re[i] = 0; st.im[i] = 0; }
    st.re[0] = owner ? T(1) : T(0);
}

template <typename T, int NL>
__device__ __forceinline__ void scale_state(ScalarState<T, NL>& st, T g) {
#pragma unroll
    for (int i = 0; i < (1 << NL); ++i) { st.re[i] *= g; st.im[i] *= g; }
}

template <typename T, int NL>
__device__ __forceinline__ T real_dot(const ScalarState<T, NL>& a, const ScalarState<T, NL>& b) {
    T e = 0;
#pragma unroll
    for (int i = 0; i < (1 << NL); ++i) { e = fma_(a.re[i], b.re[i], e); e = fma_(a.im[i], b.im[i], e); }
    return e;
}

// lam = H psi  (pauli 0: diagonal table, 1: offset + coeff sum_q X_q, 2: offset + coeff sum_q Y_q)
template <int LQ, typename T, int NL>
__device__ __forceinline__ void apply_ham(const HeaParams<T>& p, const ScalarState<T, NL>& ps, ScalarState<T, NL>& lm,
                                          int lane) {
    constexpr int NA = 1 << NL;
    constexpr int NQ = NL + LQ;
    const int sub = lane & ((1 << LQ) - 1);
    if (p.pauli == 0) {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const T d = __ldg(p.hdiag + ((sub << NL) | i));
            lm.re[i] = d * ps.re[i];
            lm.im[i] = d * ps.im[i];
        }
        return;
    }
    const bool isY = p.pauli == 2;
#pragma unroll
    for (int i = 0; i < NA; ++i) { lm.re[i] = p.offset * ps.re[i]; lm.im[i] = p.offset * ps.im[i]; }
    static_for<NQ>([&](auto Qc) {
        constexpr int Q = decltype(Qc)::value;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            T fr, fi;   // flipped amplitude psi_{k ^ bit}
            bool one;   // bit Q of this amplitude's index
            if constexpr (Q < NL) {
                fr = ps.re[i ^ (1 << Q)]; fi = ps.im[i ^ (1 << Q)]; one = (i >> Q) & 1;
            } else {
                fr = shfl_xor_(ps.re[i], 1 << (Q - NL)); fi = shfl_xor_(ps.im[i], 1 << (Q - NL));
                one = (lane >> (Q - NL)) & 1;
            }
            if (!isY) {
                lm.re[i] = fma_(p.coeff, fr, lm.re[i]); lm.im[i] = fma_(p.coeff, fi, lm.im[i]);
            } else {     // (Y psi)_k = +i psi_flip if bit set else -i psi_flip ; i(a+ib) = -b + ia
                const T sg = one ? p.coeff : -p.coeff;
                lm.re[i] = fma_(-sg, fi, lm.re[i]); lm.im[i] = fma_(sg, fr, lm.im[i]);
            }
        }
    });
}

// =================================================================================================
// Packed state (fp32, one thread per sample): amplitude k = one 64-bit register pair (re, im); all
// gate arithmetic is FFMA2 / FMUL2 through the operand-pattern wrappers of ffma2.cuh.
// =================================================================================================
template <int NL_>
struct PackedState {
    static constexpr int NL = NL_;
    static constexpr int NA = 1 << NL_;
    u64 a[NA];
};

template <int Q, bool DAG, int NL>
__device__ __forceinline__ void apply_u(PackedState<NL>& st, float ar, float ai, float br, float bi, int) {
    constexpr int NA = 1 << NL;
    constexpr int bit = 1 << Q;
    static_assert(Q < NL, "packed state is single-lane");
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        if (i & bit) continue;
        const int j = i | bit;
        const u64 x0 = st.a[i], x1 = st.a[j];
        u64 n0, n1;
        if constexpr (!DAG) {
            // a0' = al a0 - conj(be) a1 = ar x0 + ai (i x0) - br x1 + bi (i x1)
            n0 = mul2<0>(ar, x0);
            n0 = fma2<2>(ai, x0, n0);
            n0 = fma2<6>(br, x1, n0);
            n0 = fma2<2>(bi, x1, n0);
            // a1' = be a0 + conj(al) a1 = br x0 + bi (i x0) + ar x1 + ai (-i x1)
            n1 = mul2<0>(br, x0);
            n1 = fma2<2>(bi, x0, n1);
            n1 = fma2<0>(ar, x1, n1);
            n1 = fma2<3>(ai, x1, n1);
        } else {
            // a0' = conj(al) a0 + conj(be) a1 = ar x0 + ai (-i x0) + br x1 + bi (-i x1)
            n0 = mul2<0>(ar, x0);
            n0 = fma2<3>(ai, x0, n0);
            n0 = fma2<0>(br, x1, n0);
            n0 = fma2<3>(bi, x1, n0);
            // a1' = -be a0 + al a1 = -br x0 + bi (-i x0) + ar x1 + ai (i x1)
            n1 = mul2<6>(br, x0);
            n1 = fma2<3>(bi, x0, n1);
            n1 = fma2<0>(ar, x1, n1);
            n1 = fma2<2>(ai, x1, n1);
        }
        st.a[i] = n0;
        st.a[j] = n1;
    }
}

template <int Q, int NL>
__device__ __forceinline__ void bwd_group(PackedState<NL>& ps, PackedState<NL>& lm, float ar, float ai, float br,
                                          float bi, int lane, float& mX, float& mY, float& mZ) {
    constexpr int NA = 1 << NL;
    constexpr int bit = 1 << Q;
    // yx = (m_Y, m_X) accumulated as one packed value (two chains), z = m_Z scalar (two chains)
    u64 yx0 = 0ull, yx1 = 0ull;
    float z0 = 0.f, z1 = 0.f;
    int t = 0;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        if (i & bit) continue;
        const int j = i | bit;
        const u64 p0 = ps.a[i], p1 = ps.a[j];
        float l0r, l0i, l1r, l1i, p0r, p0i, p1r, p1i;
        unpack2(lm.a[i], l0r, l0i);
        unpack2(lm.a[j], l1r, l1i);
        unpack2(p0, p0r, p0i);
        unpack2(p1, p1r, p1i);
        u64 yx = (t & 1) ? yx1 : yx0;
        float z = (t & 1) ? z1 : z0;
        // (m_Y, m_X) += (-Re, Im) of conj(l0) p1  +  (Re, Im) of conj(l1) p0
        yx = fma2<4>(l0r, p1, yx);     // l0r * (-p1r,  p1i)
        yx = fma2<7>(l0i, p1, yx);     // l0i * (-p1i, -p1r)
        yx = fma2<0>(l1r, p0, yx);     // l1r * ( p0r,  p0i)
        yx = fma2<3>(l1i, p0, yx);     // l1i * ( p0i, -p0r)
        // m_Z += Im(conj(l0) p0) - Im(conj(l1) p1)
        z = fmaf(l0r, p0i, z); z = fmaf(-l0i, p0r, z);
        z = fmaf(-l1r, p1i, z); z = fmaf(l1i, p1r, z);
        if (t & 1) { yx1 = yx; z1 = z; } else { yx0 = yx; z0 = z; }
        ++t;
    }
    mY = lo2(yx0) + lo2(yx1);
    mX = hi2(yx0) + hi2(yx1);
    mZ = z0 + z1;
    apply_u<Q, true>(ps, ar, ai, br, bi, lane);
    apply_u<Q, true>(lm, ar, ai, br, bi, lane);
}

template <int C, int TG, int NL>
__device__ __forceinline__ void cnot(PackedState<NL>& st, int) {
#pragma unroll
    for (int i = 0; i < (1 << NL); ++i) {
        if (((i >> C) & 1) && !((i >> TG) & 1)) {
            const int j = i | (1 << TG);
            const u64 t = st.a[i]; st.a[i] = st.a[j]; st.a[j] = t;
        }
    }
}

template <int NL>
__device__ __forceinline__ void init_zero_state(PackedState<NL>& st, bool owner) {
#pragma unroll
    for (int i = 0; i < (1 << NL); ++i) st.a[i] = 0ull;
    st.a[0] = pack2(owner ? 1.f : 0.f, 0.f);
}

template <int NL>
__device__ __forceinline__ void scale_state(PackedState<NL>& st, float g) {
#pragma unroll
    for (int i = 0; i < (1 << NL); ++i) st.a[i] = mul2<0>(g, st.a[i]);
}

template <int NL>
__device__ __forceinline__ float real_dot(const PackedState<NL>& a, const PackedState<NL>& b) {
    u64 e0 = 0ull, e1 = 0ull;
#pragma unroll
    for (int i = 0; i < (1 << NL); ++i) {
        if (i & 1) e1 = fma2_vv(a.a[i], b.a[i], e1);
        else e0 = fma2_vv(a.a[i], b.a[i], e0);
    }
    return (lo2(e0) + hi2(e0)) + (lo2(e1) + hi2(e1));
}

template <int LQ, int NL>
__device__ __forceinline__ void apply_ham(const HeaParams<float>& p, const PackedState<NL>& ps, PackedState<NL>& lm,
                                          int) {
    static_assert(LQ == 0, "packed state is single-lane");
    constexpr int NA = 1 << NL;
    if (p.pauli == 0) {
#pragma unroll
        for (int i = 0; i < NA; ++i) lm.a[i] = mul2<0>(__ldg(p.hdiag + i), ps.a[i]);
        return;
    }
    const bool isY = p.pauli == 2;
#pragma unroll
    for (int i = 0; i < NA; ++i) lm.a[i] = mul2<0>(p.offset, ps.a[i]);
    static_for<NL>([&](auto Qc) {
        constexpr int Q = decltype(Qc)::value;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const u64 f = ps.a[i ^ (1 << Q)];
            if (!isY) {
                lm.a[i] = fma2<0>(p.coeff, f, lm.a[i]);
            } else {
                const float sg = ((i >> Q) & 1) ? p.coeff : -p.coeff;   // +i f if bit set else -i f
                lm.a[i] = fma2<2>(sg, f, lm.a[i]);
            }
        }
    });
}

// =================================================================================================
// shared pieces
// =================================================================================================
template <int LQ, bool REVERSE, typename State>
__device__ __forceinline__ void cnot_ring(State& st, int lane) {
    constexpr int NQ = State::NL + LQ;
    if constexpr (NQ > 1) {
        static_for<NQ>([&](auto I) {
            constexpr int i = REVERSE ? (NQ - 1 - decltype(I)::value) : decltype(I)::value;
            cnot<(i + 1) % NQ, i>(st, lane);
        });
    }
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

template <typename T>
__device__ __forceinline__ void fold_rx_coef(const Vec4<T>& u, T theta, T& ar, T& ai, T& br, T& bi) {
    T sn, cs;
    sincos_half(theta, sn, cs);
    ar = fma_(sn, u.w, u.x * cs); ai = fma_(sn, u.z, u.y * cs);
    br = fma_(-sn, u.y, u.z * cs); bi = fma_(-sn, u.x, u.w * cs);
}

template <typename T, int NL, int LQ>
struct StateOf {
    using type = ScalarState<T, NL>;
};
template <int NL>
struct StateOf<float, NL, 0> {
    using type = PackedState<NL>;
};

// ------------------------------------------------------------------------------------------------
// the kernel
//   ENC = 0: encoding angles x given;  1: angles formed in-kernel from (u0, u1, fw, fb);
//   ENC = 2: as 1, and the frequency-layer gradients dL/dfw, dL/dfb are reduced over the batch in-kernel.
// ------------------------------------------------------------------------------------------------
template <typename T, int NL, int LQ, bool GRAD, bool NEED_GX, int ENC, int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) hea_reg_kernel(const HeaParams<T> p) {
    using State = typename StateOf<T, NL, LQ>::type;
    constexpr int NQ = NL + LQ;
    constexpr int SPW = 32 >> LQ;                 // samples per warp
    constexpr int VP = moment_slots(NQ);
    constexpr int FVP = freq_slots(NQ);
    constexpr int WARPS = THREADS / 32;
    constexpr bool FREQ_GRAD = GRAD && ENC == 2;
    constexpr bool WANT_GX = NEED_GX || FREQ_GRAD;
    static_assert(VP <= 32 && FVP <= 32, "moment butterfly needs 3n <= 32");
    static_assert(!(NEED_GX && ENC != 0), "grad_x is only materialised when x is");
#ifndef QON_SYNC_WARPS
#define QON_SYNC_WARPS 0
#endif
    constexpr bool SYNC_WARPS = QON_SYNC_WARPS != 0;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane & ((1 << LQ) - 1);
    const int sidx = lane >> LQ;
    const int64_t gwarp = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t nwarps = (int64_t)gridDim.x * WARPS;
    const int64_t ntiles = (p.B + SPW - 1) / SPW;
    T* mrow = GRAD ? p.mpart + gwarp * p.rowlen : nullptr;
    T* frow = GRAD ? mrow + (int64_t)p.S * VP : nullptr;             // frequency-gradient slots
    T* srow = GRAD ? frow + (int64_t)p.K * FVP : nullptr;            // [sum g, sum residual^2]

    // CTA-uniform trip count (tile0 is the CTA's first tile), so an optional per-sublayer barrier is legal.
    for (int64_t tile0 = (int64_t)blockIdx.x * WARPS; tile0 < ntiles; tile0 += nwarps) {
        const int64_t tile = tile0 + warp;
        const int64_t b = tile * SPW + sidx;
        const bool valid = b < p.B;
        const int64_t bc = valid ? b : p.B - 1;
        const T* xrow = ENC == 0 ? p.x + bc * p.ldx : nullptr;
        const T* u0row = ENC != 0 && p.u0 ? p.u0 + bc * p.ldu0 : nullptr;
        const T* u1row = ENC != 0 ? p.u1 + bc * p.ldu1 : nullptr;

        auto load_angles = [&](int k, T(&th)[NQ]) {
            if constexpr (ENC == 0) {
#pragma unroll
                for (int q = 0; q < NQ; ++q) th[q] = __ldg(xrow + (int64_t)k * NQ + q);
            } else {
                const T* ur = k < p.K0 ? u0row : u1row;
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const int col = k * NQ + q;
                    const T u = __ldg(ur + __ldg(p.uidx + col));
                    th[q] = fma_(u, __ldg(p.fw + col), p.fb ? __ldg(p.fb + col) : T(0));
                }
            }
        };

        State ps;
        init_zero_state(ps, sub == 0);

        // ---------------- forward sweep ----------------
        {
            int s = 0;
            T th[NQ];
            load_angles(0, th);
            for (int k = 0; k < p.K; ++k) {
                T thn[NQ];
                load_angles(k + 1 < p.K ? k + 1 : k, thn);                       // prefetch next block
                const int d = __ldg(p.depth + k);
#pragma unroll 1
                for (int j = 0; j < d; ++j, ++s) {
                    if constexpr (SYNC_WARPS) __syncthreads();
                    const Vec4<T>* uc = p.ucoef + (int64_t)s * NQ;
                    static_for<NQ>([&](auto Qc) {
                        constexpr int Q = decltype(Qc)::value;
                        const Vec4<T> u = ldg4(uc + Q);
                        T ar = u.x, ai = u.y, br = u.z, bi = u.w;
                        if (j == 0) fold_rx_coef(u, th[Q], ar, ai, br, bi);   // first sublayer: RX folded in
                        apply_u<Q, false>(ps, ar, ai, br, bi, lane);
                    });
                    cnot_ring<LQ, false>(ps, lane);
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q) th[q] = thn[q];
            }
        }

        // ---------------- expectation value ----------------
        State lm;
        apply_ham<LQ>(p, ps, lm, lane);
        T e = real_dot(ps, lm);
#pragma unroll
        for (int m = 1; m < (1 << LQ); m <<= 1) e += shfl_xor_(e, m);
        if (valid && sub == 0 && p.out) p.out[b] = e;

        if constexpr (GRAD) {
            // ---------------- reverse (adjoint) sweep ----------------
            T g = T(0);
            if (p.target) {   // fused MSE: g = dL/dout for L = gscale/2 * sum (out + bias - y)^2
                T resid = T(0);
                if (valid) {
                    resid = e + (p.bias ? __ldg(p.bias) : T(0)) - __ldg(p.target + b);
                    g = p.gscale * resid;
                    if (sub == 0 && p.gbuf) p.gbuf[b] = g;
                }
                T sg = sub == 0 ? g : T(0), sq = sub == 0 ? resid * resid : T(0);
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) { sg += shfl_xor_(sg, m); sq += shfl_xor_(sq, m); }
                if (lane == 0) { atomicAdd(srow, sg); atomicAdd(srow + 1, sq); }
            } else if (valid) {
                g = __ldg(p.gout + b);
            }
            scale_state(lm, g);
            T* gxrow = NEED_GX ? p.gx + (valid ? b : 0) * p.ldgx : nullptr;
            int s = p.S;
            T th[NQ];
            load_angles(p.K - 1, th);
            for (int k = p.K - 1; k >= 0; --k) {
                T thn[NQ];
                load_angles(k > 0 ? k - 1 : 0, thn);
                const int d = __ldg(p.depth + k);
#pragma unroll 1
                for (int j = d - 1; j >= 0; --j) {
                    if constexpr (SYNC_WARPS) __syncthreads();
                    --s;
                    const Vec4<T>* uc = p.ucoef + (int64_t)s * NQ;
                    const Vec4<T>* rc = p.rcoef + (int64_t)s * NQ;
                    cnot_ring<LQ, true>(ps, lane);
                    cnot_ring<LQ, true>(lm, lane);
                    T mv[VP];
#pragma unroll
                    for (int i = 0; i < VP; ++i) mv[i] = 0;
                    T fv[FREQ_GRAD ? FVP : 1];
                    if constexpr (FREQ_GRAD) {
#pragma unroll
                        for (int i = 0; i < FVP; ++i) fv[i] = 0;
                    }
                    static_for<NQ>([&](auto Qc) {
                        constexpr int Q = NQ - 1 - decltype(Qc)::value;
                        const Vec4<T> u = ldg4(uc + Q);
                        T ar = u.x, ai = u.y, br = u.z, bi = u.w;
                        if (j == 0) fold_rx_coef(u, th[Q], ar, ai, br, bi);
                        bwd_group<Q>(ps, lm, ar, ai, br, bi, lane, mv[3 * Q], mv[3 * Q + 1], mv[3 * Q + 2]);
                        if constexpr (WANT_GX) {
                            if (j == 0) {   // dL/dtheta of the folded RX from the three moments
                                T mx = mv[3 * Q], my = mv[3 * Q + 1], mz = mv[3 * Q + 2];
#pragma unroll
                                for (int m = 1; m < (1 << LQ); m <<= 1) {
                                    mx += shfl_xor_(mx, m); my += shfl_xor_(my, m); mz += shfl_xor_(mz, m);
                                }
                                const Vec4<T> r = ldg4(rc + Q);
                                const T gxv = fma_(r.z, mz, fma_(r.y, my, r.x * mx));
                                if constexpr (NEED_GX) {
                                    if (valid && sub == 0) gxrow[(int64_t)k * NQ + Q] = gxv;
                                }
                                if constexpr (FREQ_GRAD) {   // theta = fw*u + fb  =>  d/dfw = gx*u, d/dfb = gx
                                    const int col = k * NQ + Q;
                                    const T uval = __ldg((k < p.K0 ? u0row : u1row) + __ldg(p.uidx + col));
                                    const T gq = sub == 0 ? gxv : T(0);      // invalid samples carry g = 0
                                    fv[2 * Q] = gq * uval;
                                    fv[2 * Q + 1] = gq;
                                }
                            }
                        }
                    });
                    const T tot = butterfly_reduce<T, VP>(mv, lane);
                    if ((lane & (32 / VP - 1)) == 0) atomicAdd(mrow + (int64_t)s * VP + lane / (32 / VP), tot);
                    if constexpr (FREQ_GRAD) {
                        if (j == 0) {
                            const T ft = butterfly_reduce<T, FVP>(fv, lane);
                            if ((lane & (32 / FVP - 1)) == 0) atomicAdd(frow + (int64_t)k * FVP + lane / (32 / FVP), ft);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q) th[q] = thn[q];
            }
        }
    }
}

}  // namespace qon
