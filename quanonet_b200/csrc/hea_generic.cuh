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

template <typename T>
__device__ __forceinline__ void g_apply(T* re, T* im, int64_t N, int q, T ar, T ai, T br, T bi, bool dag) {
    const int64_t half = N >> 1;
    const int64_t bit = (int64_t)1 << q;
    for (int64_t t = threadIdx.x; t < half; t += blockDim.x) {
        const int64_t i = ((t >> q) << (q + 1)) | (t & (bit - 1));
        const int64_t j = i | bit;
        const T x0r = re[i], x0i = im[i], x1r = re[j], x1i = im[j];
        if (!dag) {
            re[i] = fma_(-bi, x1i, fma_(-br, x1r, fma_(-ai, x0i, ar * x0r)));
            im[i] = fma_(bi, x1r, fma_(-br, x1i, fma_(ai, x0r, ar * x0i)));
            re[j] = fma_(ai, x1i, fma_(ar, x1r, fma_(-bi, x0i, br * x0r)));
            im[j] = fma_(-ai, x1r, fma_(ar, x1i, fma_(bi, x0r, br * x0i)));
        } else {
            re[i] = fma_(bi, x1i, fma_(br, x1r, fma_(ai, x0i, ar * x0r)));
            im[i] = fma_(-bi, x1r, fma_(br, x1i, fma_(-ai, x0r, ar * x0i)));
            re[j] = fma_(-ai, x1i, fma_(ar, x1r, fma_(bi, x0i, -br * x0r)));
            im[j] = fma_(ai, x1r, fma_(ar, x1i, fma_(-bi, x0r, -br * x0i)));
        }
    }
    __syncthreads();
}

template <typename T>
__device__ __forceinline__ void g_cnot(T* re, T* im, int64_t N, int c, int tg) {
    const int64_t quarter = N >> 2;
    const int lo = c < tg ? c : tg, hi = c < tg ? tg : c;
    for (int64_t t = threadIdx.x; t < quarter; t += blockDim.x) {
        // insert zero bits at positions lo and hi
        int64_t i = ((t >> lo) << (lo + 1)) | (t & (((int64_t)1 << lo) - 1));
        i = ((i >> hi) << (hi + 1)) | (i & (((int64_t)1 << hi) - 1));
        const int64_t a = i | ((int64_t)1 << c);
        const int64_t b = a | ((int64_t)1 << tg);
        T v = re[a]; re[a] = re[b]; re[b] = v;
        v = im[a]; im[a] = im[b]; im[b] = v;
    }
    __syncthreads();
}

template <typename T>
__device__ __forceinline__ void g_ring(T* re, T* im, int64_t N, int n, bool reverse) {
    if (n < 2) return;
    for (int t = 0; t < n; ++t) {
        const int i = reverse ? n - 1 - t : t;
        g_cnot(re, im, N, (i + 1) % n, i);
    }
}

template <typename T>
__device__ __forceinline__ void fold_rx(const Vec4<T>& u, T theta, T& ar, T& ai, T& br, T& bi) {
    T sn, cs;
    sincos_half(theta, sn, cs);
    ar = fma_(sn, u.w, u.x * cs); ai = fma_(sn, u.z, u.y * cs);
    br = fma_(-sn, u.y, u.z * cs); bi = fma_(-sn, u.x, u.w * cs);
}

// moments + un-apply on (psi, lam) for qubit q; returns block-wide totals in mx,my,mz
template <typename T>
__device__ __forceinline__ void g_bwd_group(T* pr, T* pi, T* lr, T* li, int64_t N, int q,
                                            T ar, T ai, T br, T bi, T* red, T& mx, T& my, T& mz) {
    const int64_t half = N >> 1;
    const int64_t bit = (int64_t)1 << q;
    T x = 0, y = 0, z = 0;
    for (int64_t t = threadIdx.x; t < half; t += blockDim.x) {
        const int64_t i = ((t >> q) << (q + 1)) | (t & (bit - 1));
        const int64_t j = i | bit;
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
    mx = block_sum(x, red);
    my = block_sum(y, red);
    mz = block_sum(z, red);   // block_sum's barriers also order the state writes above
}

// STATE_GLOBAL: psi/lam slices in the HBM workspace instead of dynamic shared memory
template <typename T, bool GRAD, bool NEED_GX, bool STATE_GLOBAL>
__global__ void hea_generic_kernel(const HeaParams<T> p, const int n, const int VP, T* gstate) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ T red[33];
    const int64_t N = (int64_t)1 << n;
    T* base = STATE_GLOBAL ? gstate + (int64_t)blockIdx.x * (GRAD ? 4 : 2) * N : reinterpret_cast<T*>(smem_raw);
    T* pr = base;
    T* pi = base + N;
    T* lr = GRAD ? base + 2 * N : nullptr;
    T* li = GRAD ? base + 3 * N : nullptr;
    T* mrow = GRAD ? p.mpart + (int64_t)blockIdx.x * p.rowlen : nullptr;

    for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
        const T* xrow = p.x + b * p.ldx;
        for (int64_t i = threadIdx.x; i < N; i += blockDim.x) { pr[i] = i == 0 ? T(1) : T(0); pi[i] = 0; }
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
                    g_apply(pr, pi, N, q, ar, ai, br, bi, false);
                }
                g_ring(pr, pi, N, n, false);
            }
        }
        // ---- expectation: e = sum_k Re(conj(psi_k) (H psi)_k); lam = H psi (scaled by g below)
        T e = 0;
        for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
            T hr, hi;
            if (p.pauli == 0) {
                const T d = p.hdiag[i];
                hr = d * pr[i]; hi = d * pi[i];
            } else {
                hr = p.offset * pr[i]; hi = p.offset * pi[i];
                for (int q = 0; q < n; ++q) {
                    const int64_t f = i ^ ((int64_t)1 << q);
                    if (p.pauli == 1) { hr = fma_(p.coeff, pr[f], hr); hi = fma_(p.coeff, pi[f], hi); }
                    else {
                        const T sg = ((i >> q) & 1) ? p.coeff : -p.coeff;
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
                    g_ring(pr, pi, N, n, true);
                    g_ring(lr, li, N, n, true);
                    for (int q = n - 1; q >= 0; --q) {
                        const Vec4<T> u = ldg4(p.ucoef + (int64_t)s * n + q);
                        T ar = u.x, ai = u.y, br = u.z, bi = u.w;
                        if (j == 0) fold_rx(u, xrow[(int64_t)k * n + q], ar, ai, br, bi);
                        T mx, my, mz;
                        g_bwd_group(pr, pi, lr, li, N, q, ar, ai, br, bi, red, mx, my, mz);
                        if (threadIdx.x == 0) {
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
