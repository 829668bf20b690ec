// Generic tier: one CTA owns one sample; the state (psi and, in the reverse sweep, lam) lives in
// shared memory when it fits (n <= 13 forward / 12 with gradients in fp32) and otherwise in a
// per-CTA slice of the HBM workspace (the streamed tier).  Same fused-gate algebra and the same
// prep/finalize tables as the register tier (hea_reg.cuh); used for qubit counts the register
// tier does not cover.  Reference semantics: core/quantum_circuits_tq.py:65-127.
#pragma once
#include "hea_common.cuh"

namespace qon {

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red /* >= 33 entries */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    __syncthreads();                 // protect red[] from the previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    T t = 0;
    for (int w = 0; w < nw; ++w) t += red[w];   // fixed order -> deterministic
    return t;
}

// The CNOT ring is never applied to the data: it is a GF(2)-linear relabelling of the basis states, tracked as a
// map between LOGICAL qubits and PHYSICAL index bits (round 1 swept the state n times per ring, 2n more in the
// reverse sweep).  For logical qubit q:  m[q] = physical XOR mask that flips it (column of A^-1),
// r[q] = physical parity mask that reads it (row of A), with logical = A * physical.  CNOT(control c -> target t)
// maps A <- C A:  r[t] ^= r[c],  m[c] ^= m[t]   (core/quantum_circuits_tq.py:98-101: c = (i+1)%n, t = i, i ascending).
struct RingMap { unsigned m[24], r[24]; };

__device__ __forceinline__ void ring_map_apply(RingMap& rm, int n, bool reverse) {
    if (n < 2) return;
    for (int t = 0; t < n; ++t) {
        const int i = reverse ? n - 1 - t : t, c = (i + 1) % n;
        rm.r[i] ^= rm.r[c];
        rm.m[c] ^= rm.m[i];
    }
}

// gate on the logical qubit with flip mask m / read mask r: pairs (i, i ^ m), |0> component = the index of even parity
template <typename T>
__device__ __forceinline__ void g_apply(T* re, T* im, int64_t N, unsigned m, unsigned r, T ar, T ai, T br, T bi) {
    const int64_t half = N >> 1;
    const int lb = __ffs((int)m) - 1;
    const int64_t low = ((int64_t)1 << lb) - 1;
    for (int64_t t = threadIdx.x; t < half; t += blockDim.x) {
        int64_t i = ((t >> lb) << (lb + 1)) | (t & low);
        int64_t j = i ^ (int64_t)m;
        if (__popc((unsigned)i & r) & 1) { const int64_t w = i; i = j; j = w; }
        const T x0r = re[i], x0i = im[i], x1r = re[j], x1i = im[j];
        re[i] = fma_(-bi, x1i, fma_(-br, x1r, fma_(-ai, x0i, ar * x0r)));
        im[i] = fma_(bi, x1r, fma_(-br, x1i, fma_(ai, x0r, ar * x0i)));
        re[j] = fma_(ai, x1i, fma_(ar, x1r, fma_(-bi, x0i, br * x0r)));
        im[j] = fma_(-ai, x1r, fma_(ar, x1i, fma_(bi, x0r, br * x0i)));
    }
    __syncthreads();
}

// Two gates on two different logical qubits in ONE pass over the state (half the shared-memory traffic and half the
// barriers of two g_apply calls): the four states {e, e^m1, e^m2, e^m1^m2} form a closed group; its representative
// has the two pivot bits of (m1, m2) cleared, e00 is the member whose two logical bits are 0.
struct G2 {
    int lo, hi;          // pivot bit positions, lo < hi
    unsigned m1, m2;
};
__device__ __forceinline__ G2 g2_make(unsigned m1, unsigned m2) {
    const int l1 = __ffs((int)m1) - 1;
    const unsigned m2p = ((m2 >> l1) & 1u) ? (m2 ^ m1) : m2;      // same group, pivot of m1 cleared
    const int l2 = __ffs((int)m2p) - 1;
    G2 g;
    g.lo = l1 < l2 ? l1 : l2;
    g.hi = l1 < l2 ? l2 : l1;
    g.m1 = m1;
    g.m2 = m2;
    return g;
}
__device__ __forceinline__ int64_t g2_rep(const G2& g, int64_t t) {
    int64_t i = ((t >> g.lo) << (g.lo + 1)) | (t & (((int64_t)1 << g.lo) - 1));
    return ((i >> g.hi) << (g.hi + 1)) | (i & (((int64_t)1 << g.hi) - 1));
}
template <typename T>
__device__ __forceinline__ void gate2x2(T& x0r, T& x0i, T& x1r, T& x1i, T ar, T ai, T br, T bi) {
    const T n0r = fma_(-bi, x1i, fma_(-br, x1r, fma_(-ai, x0i, ar * x0r)));
    const T n0i = fma_(bi, x1r, fma_(-br, x1i, fma_(ai, x0r, ar * x0i)));
    const T n1r = fma_(ai, x1i, fma_(ar, x1r, fma_(-bi, x0i, br * x0r)));
    const T n1i = fma_(-ai, x1r, fma_(ar, x1i, fma_(bi, x0r, br * x0i)));
    x0r = n0r; x0i = n0i; x1r = n1r; x1i = n1i;
}
// U^+ on a pair
template <typename T>
__device__ __forceinline__ void ungate2x2(T& p0r, T& p0i, T& p1r, T& p1i, T ar, T ai, T br, T bi) {
    const T n0r = fma_(bi, p1i, fma_(br, p1r, fma_(ai, p0i, ar * p0r)));
    const T n0i = fma_(-bi, p1r, fma_(br, p1i, fma_(-ai, p0r, ar * p0i)));
    const T n1r = fma_(-ai, p1i, fma_(ar, p1r, fma_(bi, p0i, -br * p0r)));
    const T n1i = fma_(ai, p1r, fma_(ar, p1i, fma_(-bi, p0r, -br * p0i)));
    p0r = n0r; p0i = n0i; p1r = n1r; p1i = n1i;
}
template <typename T>
__device__ __forceinline__ void g_apply2(T* re, T* im, int64_t N, unsigned m1, unsigned r1, unsigned m2, unsigned r2,
                                         const T (&u1)[4], const T (&u2)[4]) {
    const G2 g = g2_make(m1, m2);
    const int64_t quarter = N >> 2;
    for (int64_t t = threadIdx.x; t < quarter; t += blockDim.x) {
        const int64_t i0 = g2_rep(g, t);
        const int64_t e00 = i0 ^ ((__popc((unsigned)i0 & r1) & 1) ? (int64_t)m1 : 0) ^ ((__popc((unsigned)i0 & r2) & 1) ? (int64_t)m2 : 0);
        const int64_t e10 = e00 ^ (int64_t)m1, e01 = e00 ^ (int64_t)m2, e11 = e10 ^ (int64_t)m2;
        T ar0 = re[e00], ai0 = im[e00], ar1 = re[e10], ai1 = im[e10], ar2 = re[e01], ai2 = im[e01], ar3 = re[e11], ai3 = im[e11];
        gate2x2(ar0, ai0, ar1, ai1, u1[0], u1[1], u1[2], u1[3]);      // qubit 1: (00, 10) and (01, 11)
        gate2x2(ar2, ai2, ar3, ai3, u1[0], u1[1], u1[2], u1[3]);
        gate2x2(ar0, ai0, ar2, ai2, u2[0], u2[1], u2[2], u2[3]);      // qubit 2: (00, 01) and (10, 11)
        gate2x2(ar1, ai1, ar3, ai3, u2[0], u2[1], u2[2], u2[3]);
        re[e00] = ar0; im[e00] = ai0; re[e10] = ar1; im[e10] = ai1;
        re[e01] = ar2; im[e01] = ai2; re[e11] = ar3; im[e11] = ai3;
    }
    __syncthreads();
}

// moments of one pair, accumulated: x, y, z += Im <lam| {X, Y, Z} |psi> restricted to the pair
template <typename T>
__device__ __forceinline__ void pair_moments(T p0r, T p0i, T p1r, T p1i, T l0r, T l0i, T l1r, T l1i, T& x, T& y, T& z) {
    x = fma_(l0r, p1i, x); x = fma_(-l0i, p1r, x); x = fma_(l1r, p0i, x); x = fma_(-l1i, p0r, x);
    y = fma_(-l0r, p1r, y); y = fma_(-l0i, p1i, y); y = fma_(l1r, p0r, y); y = fma_(l1i, p0i, y);
    z = fma_(l0r, p0i, z); z = fma_(-l0i, p0r, z); z = fma_(-l1r, p1i, z); z = fma_(l1i, p1r, z);
}
// reverse sweep for two logical qubits in one pass: moments of qubit 1, un-apply it on (psi, lam), moments of qubit 2 on
// the result, un-apply it.  Per-warp partials: part1 / part2.
template <typename T>
__device__ __forceinline__ void g_bwd_group2(T* pr, T* pi, T* lr, T* li, int64_t N, unsigned m1, unsigned r1, unsigned m2,
                                             unsigned r2, const T (&u1)[4], const T (&u2)[4], T (*part1)[3], T (*part2)[3]) {
    const G2 g = g2_make(m1, m2);
    const int64_t quarter = N >> 2;
    T x1 = 0, y1 = 0, z1 = 0, x2 = 0, y2 = 0, z2 = 0;
    for (int64_t t = threadIdx.x; t < quarter; t += blockDim.x) {
        const int64_t i0 = g2_rep(g, t);
        const int64_t e00 = i0 ^ ((__popc((unsigned)i0 & r1) & 1) ? (int64_t)m1 : 0) ^ ((__popc((unsigned)i0 & r2) & 1) ? (int64_t)m2 : 0);
        const int64_t e[4] = {e00, e00 ^ (int64_t)m1, e00 ^ (int64_t)m2, e00 ^ (int64_t)m1 ^ (int64_t)m2};     // 00, 10, 01, 11
        T P[4][2], L[4][2];
#pragma unroll
        for (int c = 0; c < 4; ++c) { P[c][0] = pr[e[c]]; P[c][1] = pi[e[c]]; L[c][0] = lr[e[c]]; L[c][1] = li[e[c]]; }
        pair_moments(P[0][0], P[0][1], P[1][0], P[1][1], L[0][0], L[0][1], L[1][0], L[1][1], x1, y1, z1);
        pair_moments(P[2][0], P[2][1], P[3][0], P[3][1], L[2][0], L[2][1], L[3][0], L[3][1], x1, y1, z1);
        ungate2x2(P[0][0], P[0][1], P[1][0], P[1][1], u1[0], u1[1], u1[2], u1[3]);
        ungate2x2(P[2][0], P[2][1], P[3][0], P[3][1], u1[0], u1[1], u1[2], u1[3]);
        ungate2x2(L[0][0], L[0][1], L[1][0], L[1][1], u1[0], u1[1], u1[2], u1[3]);
        ungate2x2(L[2][0], L[2][1], L[3][0], L[3][1], u1[0], u1[1], u1[2], u1[3]);
        pair_moments(P[0][0], P[0][1], P[2][0], P[2][1], L[0][0], L[0][1], L[2][0], L[2][1], x2, y2, z2);
        pair_moments(P[1][0], P[1][1], P[3][0], P[3][1], L[1][0], L[1][1], L[3][0], L[3][1], x2, y2, z2);
        ungate2x2(P[0][0], P[0][1], P[2][0], P[2][1], u2[0], u2[1], u2[2], u2[3]);
        ungate2x2(P[1][0], P[1][1], P[3][0], P[3][1], u2[0], u2[1], u2[2], u2[3]);
        ungate2x2(L[0][0], L[0][1], L[2][0], L[2][1], u2[0], u2[1], u2[2], u2[3]);
        ungate2x2(L[1][0], L[1][1], L[3][0], L[3][1], u2[0], u2[1], u2[2], u2[3]);
#pragma unroll
        for (int c = 0; c < 4; ++c) { pr[e[c]] = P[c][0]; pi[e[c]] = P[c][1]; lr[e[c]] = L[c][0]; li[e[c]] = L[c][1]; }
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        x1 += __shfl_xor_sync(0xffffffffu, x1, s); y1 += __shfl_xor_sync(0xffffffffu, y1, s); z1 += __shfl_xor_sync(0xffffffffu, z1, s);
        x2 += __shfl_xor_sync(0xffffffffu, x2, s); y2 += __shfl_xor_sync(0xffffffffu, y2, s); z2 += __shfl_xor_sync(0xffffffffu, z2, s);
    }
    if ((threadIdx.x & 31) == 0) {
        T* a = part1[threadIdx.x >> 5];
        T* b = part2[threadIdx.x >> 5];
        a[0] = x1; a[1] = y1; a[2] = z1;
        b[0] = x2; b[1] = y2; b[2] = z2;
    }
    __syncthreads();
}

template <typename T>
__device__ __forceinline__ void fold_rx(const Vec4<T>& u, T theta, T& ar, T& ai, T& br, T& bi) {
    T sn, cs;
    sincos_half(theta, sn, cs);
    ar = fma_(sn, u.w, u.x * cs); ai = fma_(sn, u.z, u.y * cs);
    br = fma_(-sn, u.y, u.z * cs); bi = fma_(-sn, u.x, u.w * cs);
}

// moments + un-apply on (psi, lam) for one logical qubit; per-warp partial moments go to part[warp][0..2] (summed in
// fixed order by the caller after the barrier that also orders the state writes)
template <typename T>
__device__ __forceinline__ void g_bwd_group(T* pr, T* pi, T* lr, T* li, int64_t N, unsigned m, unsigned r,
                                            T ar, T ai, T br, T bi, T (*part)[3]) {
    const int64_t half = N >> 1;
    const int lb = __ffs((int)m) - 1;
    const int64_t low = ((int64_t)1 << lb) - 1;
    T x = 0, y = 0, z = 0;
    for (int64_t t = threadIdx.x; t < half; t += blockDim.x) {
        int64_t i = ((t >> lb) << (lb + 1)) | (t & low);
        int64_t j = i ^ (int64_t)m;
        if (__popc((unsigned)i & r) & 1) { const int64_t w = i; i = j; j = w; }
        const T p0r = pr[i], p0i = pi[i], p1r = pr[j], p1i = pi[j];
        const T l0r = lr[i], l0i = li[i], l1r = lr[j], l1i = li[j];
        x = fma_(l0r, p1i, x); x = fma_(-l0i, p1r, x); x = fma_(l1r, p0i, x); x = fma_(-l1i, p0r, x);
        y = fma_(-l0r, p1r, y); y = fma_(-l0i, p1i, y); y = fma_(l1r, p0r, y); y = fma_(l1i, p0i, y);
        z = fma_(l0r, p0i, z); z = fma_(-l0i, p0r, z); z = fma_(-l1r, p1i, z); z = fma_(l1i, p1r, z);
        pr[i] = fma_(bi, p1i, fma_(br, p1r, fma_(ai, p0i, ar * p0r)));
        pi[i] = fma_(-bi, p1r, fma_(br, p1i, fma_(-ai, p0r, ar * p0i)));
        pr[j] = fma_(-ai, p1i, fma_(ar, p1r, fma_(bi, p0i, -br * p0r)));
        pi[j] = fma_(ai, p1r, fma_(ar, p1i, fma_(-bi, p0r, -br * p0i)));
        lr[i] = fma_(bi, l1i, fma_(br, l1r, fma_(ai, l0i, ar * l0r)));
        li[i] = fma_(-bi, l1r, fma_(br, l1i, fma_(-ai, l0r, ar * l0i)));
        lr[j] = fma_(-ai, l1i, fma_(ar, l1r, fma_(bi, l0i, -br * l0r)));
        li[j] = fma_(ai, l1r, fma_(ar, l1i, fma_(-bi, l0r, -br * l0i)));
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        x += __shfl_xor_sync(0xffffffffu, x, s);
        y += __shfl_xor_sync(0xffffffffu, y, s);
        z += __shfl_xor_sync(0xffffffffu, z, s);
    }
    if ((threadIdx.x & 31) == 0) {
        T* row = part[threadIdx.x >> 5];
        row[0] = x; row[1] = y; row[2] = z;
    }
    __syncthreads();
}

// STATE_GLOBAL: psi/lam slices in the HBM workspace instead of dynamic shared memory
template <typename T, bool GRAD, bool NEED_GX, bool STATE_GLOBAL>
__global__ void __launch_bounds__(512) hea_generic_kernel(const HeaParams<T> p, const int n, const int VP, T* gstate) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ T red[33];
    __shared__ T part[4][32][3];          // per-warp moment partials: two gates per pass, double-buffered by pass parity
    __shared__ RingMap rm;
    const int64_t N = (int64_t)1 << n;
    const int nwarps = (blockDim.x + 31) >> 5;
    T* base = STATE_GLOBAL ? gstate + (int64_t)blockIdx.x * (GRAD ? 4 : 2) * N : reinterpret_cast<T*>(smem_raw);
    T* pr = base;
    T* pi = base + N;
    T* lr = GRAD ? base + 2 * N : nullptr;
    T* li = GRAD ? base + 3 * N : nullptr;
    T* mrow = GRAD ? p.mpart + (int64_t)blockIdx.x * p.rowlen : nullptr;

    for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
        const T* xrow = p.x + b * p.ldx;
        for (int64_t i = threadIdx.x; i < N; i += blockDim.x) { pr[i] = i == 0 ? T(1) : T(0); pi[i] = 0; }
        if (threadIdx.x < n) { rm.m[threadIdx.x] = 1u << threadIdx.x; rm.r[threadIdx.x] = 1u << threadIdx.x; }
        __syncthreads();
        // ---- forward
        int s = 0;
        for (int k = 0; k < p.K; ++k) {
            const int d = p.depth[k];
            for (int j = 0; j < d; ++j, ++s) {
                auto coef = [&](int q, T(&c)[4]) {
                    const Vec4<T> u = ldg4(p.ucoef + (int64_t)s * n + q);
                    c[0] = u.x; c[1] = u.y; c[2] = u.z; c[3] = u.w;
                    if (j == 0) fold_rx(u, xrow[(int64_t)k * n + q], c[0], c[1], c[2], c[3]);
                };
                int q = 0;
                for (; q + 1 < n; q += 2) {          // two qubits per pass over the state
                    T c1[4], c2[4];
                    coef(q, c1);
                    coef(q + 1, c2);
                    g_apply2(pr, pi, N, rm.m[q], rm.r[q], rm.m[q + 1], rm.r[q + 1], c1, c2);
                }
                if (q < n) {
                    T c1[4];
                    coef(q, c1);
                    g_apply(pr, pi, N, rm.m[q], rm.r[q], c1[0], c1[1], c1[2], c1[3]);
                }
                if (threadIdx.x == 0) ring_map_apply(rm, n, false);      // the ring relabels; the data stays
                __syncthreads();
            }
        }
        // ---- expectation: e = sum_k Re(conj(psi_k) (H psi)_k); lam = H psi (scaled by g below)
        T e = 0;
        for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
            T hr, hi;
            if (p.pauli == 0) {
                int64_t kl = 0;                                   // logical basis index of physical slot i
                for (int q = 0; q < n; ++q) kl |= (int64_t)(__popc((unsigned)i & rm.r[q]) & 1) << q;
                const T d = p.hdiag[kl];
                hr = d * pr[i]; hi = d * pi[i];
            } else {
                hr = p.offset * pr[i]; hi = p.offset * pi[i];
                for (int q = 0; q < n; ++q) {
                    const int64_t f = i ^ (int64_t)rm.m[q];
                    if (p.pauli == 1) { hr = fma_(p.coeff, pr[f], hr); hi = fma_(p.coeff, pi[f], hi); }
                    else {
                        const T sg = (__popc((unsigned)i & rm.r[q]) & 1) ? p.coeff : -p.coeff;
                        hr = fma_(-sg, pi[f], hr); hi = fma_(sg, pr[f], hi);
                    }
                }
            }
            e = fma_(pr[i], hr, e); e = fma_(pi[i], hi, e);
            if (GRAD) { lr[i] = hr; li[i] = hi; }
        }
        e = block_sum(e, red);
        if (threadIdx.x == 0) p.out[b] = e;
        __syncthreads();
        if constexpr (GRAD) {
            T g;
            if (p.target) {
                g = p.gscale * (e + (p.bias ? p.bias[0] : T(0)) - p.target[b]);
                if (threadIdx.x == 0) p.gbuf[b] = g;
            } else {
                g = p.gout[b];
            }
            for (int64_t i = threadIdx.x; i < N; i += blockDim.x) { lr[i] *= g; li[i] *= g; }
            __syncthreads();
            // ---- reverse sweep
            T* gxrow = NEED_GX ? p.gx + b * p.ldgx : nullptr;
            s = p.S;
            for (int k = p.K - 1; k >= 0; --k) {
                const int d = p.depth[k];
                for (int j = d - 1; j >= 0; --j) {
                    --s;
                    if (threadIdx.x == 0) ring_map_apply(rm, n, true);       // un-apply the ring: relabel back
                    __syncthreads();
                    auto coef = [&](int q, T(&c)[4]) {
                        const Vec4<T> u = ldg4(p.ucoef + (int64_t)s * n + q);
                        c[0] = u.x; c[1] = u.y; c[2] = u.z; c[3] = u.w;
                        if (j == 0) fold_rx(u, xrow[(int64_t)k * n + q], c[0], c[1], c[2], c[3]);
                    };
                    auto emit = [&](int q, T (*pp)[3]) {     // thread 0: per-warp partials -> this CTA's row, fixed order
                        T mx = 0, my = 0, mz = 0;
                        for (int w = 0; w < nwarps; ++w) { mx += pp[w][0]; my += pp[w][1]; mz += pp[w][2]; }
                        T* m = mrow + (int64_t)s * VP + 3 * q;
                        m[0] += mx; m[1] += my; m[2] += mz;      // row is private to this CTA
                        if (NEED_GX && j == 0) {
                            const Vec4<T> r = ldg4(p.rcoef + (int64_t)s * n + q);
                            gxrow[(int64_t)k * n + q] = fma_(r.z, mz, fma_(r.y, my, r.x * mx));
                        }
                    };
                    int q = n - 1, pass = 0;
                    for (; q >= 1; q -= 2, ++pass) {      // two qubits per pass over (psi, lam)
                        T c1[4], c2[4];
                        coef(q, c1);
                        coef(q - 1, c2);
                        T (*p1)[3] = part[2 * (pass & 1)];
                        T (*p2)[3] = part[2 * (pass & 1) + 1];
                        g_bwd_group2(pr, pi, lr, li, N, rm.m[q], rm.r[q], rm.m[q - 1], rm.r[q - 1], c1, c2, p1, p2);
                        if (threadIdx.x == 0) { emit(q, p1); emit(q - 1, p2); }
                    }
                    if (q == 0) {
                        T c1[4];
                        coef(0, c1);
                        T (*pp)[3] = part[2 * (pass & 1)];
                        g_bwd_group(pr, pi, lr, li, N, rm.m[0], rm.r[0], c1[0], c1[1], c1[2], c1[3], pp);
                        if (threadIdx.x == 0) emit(0, pp);
                    }
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace qon
