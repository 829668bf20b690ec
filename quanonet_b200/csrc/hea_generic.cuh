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
__global__ void hea_generic_kernel(const HeaParams<T> p, const int n, const int VP, T* gstate) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ T red[33];
    __shared__ T part[2][32][3];          // per-warp moment partials, double-buffered by gate parity
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
                for (int q = 0; q < n; ++q) {
                    const Vec4<T> u = ldg4(p.ucoef + (int64_t)s * n + q);
                    T ar = u.x, ai = u.y, br = u.z, bi = u.w;
                    if (j == 0) fold_rx(u, xrow[(int64_t)k * n + q], ar, ai, br, bi);
                    g_apply(pr, pi, N, rm.m[q], rm.r[q], ar, ai, br, bi);
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
                    for (int q = n - 1; q >= 0; --q) {
                        const Vec4<T> u = ldg4(p.ucoef + (int64_t)s * n + q);
                        T ar = u.x, ai = u.y, br = u.z, bi = u.w;
                        if (j == 0) fold_rx(u, xrow[(int64_t)k * n + q], ar, ai, br, bi);
                        T (*pp)[3] = part[q & 1];
                        g_bwd_group(pr, pi, lr, li, N, rm.m[q], rm.r[q], ar, ai, br, bi, pp);
                        if (threadIdx.x == 0) {
                            T mx = 0, my = 0, mz = 0;
                            for (int w = 0; w < nwarps; ++w) { mx += pp[w][0]; my += pp[w][1]; mz += pp[w][2]; }   // fixed order
                            T* m = mrow + (int64_t)s * VP + 3 * q;
                            m[0] += mx; m[1] += my; m[2] += mz;      // row is private to this CTA
                            if (NEED_GX && j == 0) {
                                const Vec4<T> r = ldg4(p.rcoef + (int64_t)s * n + q);
                                gxrow[(int64_t)k * n + q] = fma_(r.z, mz, fma_(r.y, my, r.x * mx));
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace qon
