// Tensor-core tier (n = 5, fp32 results): the sample-independent part of every block — its ansatz sublayers,
// i.e. RY.RZ.RY on every qubit + the CNOT ring, `depth` times — is pre-fused into ONE 32x32 complex unitary
// (reference: the shared sublayers of core/quantum_circuits_tq.py:89-101) and applied to 128 samples at a time as
// a real 128 x 64 x 64 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulator in TMEM).
//
// The per-sample part of a block, the RX data-encoding layer (core/quantum_circuits_tq.py:80-86), is diagonal
// in the Hadamard basis:  RX(t) = H exp(-i t Z / 2) H.  The Hadamards are sample-independent and go into the
// block matrices (M_k = H W_k H, last block: W_k H), so between two GEMMs a thread multiplies its sample's 32
// amplitudes by 32 phases  prod_q exp(-/+ i t_q / 2)  — elementwise, no pairing, no shuffles.
//
// Precision: the tensor cores multiply f16 x f16 -> f32.  Both operands are split hi + lo (11 + 11 significant
// bits, operands pre-scaled into the f16 normal range) and three products are accumulated in one f32
// accumulator: hi.hi + hi.lo + lo.hi  (the dropped lo.lo term is 2^-22 relative).  Measured parity vs the fp64
// oracle is reported in profiles/ (same 1e-5 norm-relative bar as the FFMA2 kernels).
//
// This header: constants, the forward prep kernel (block matrices -> f16 hi/lo operand images) and the phase table;
// the kernels are in hea_tc2.cuh (round 2's first forward kernel — one CTA-wide producer warp, bar.sync hand-offs —
// measured 2.28 ms per 1M samples against 1.87 ms for the per-tile MMA warps of hea_tc2.cuh and was removed;
// profiles/r2_tc_fwd_v1_ncu_summary.md keeps its profile).
#pragma once
#include "hea_common.cuh"
#include "ffma2.cuh"
#include "tc_ptx.cuh"
#include <cuda_fp16.h>

namespace qon {

#ifndef QON_TC_SLOW_INLINE
#define QON_TC_SLOW_INLINE 0      // experiment switch (scripts/build_tc_variant.sh): huge-angle sin/cos inlined instead of called
#endif
#ifndef QON_TC_REV_REGS
#define QON_TC_REV_REGS 232       // experiment switch: registers of the reverse kernel's compute warps (the MMA warpgroup keeps the rest)
#endif
#ifndef QON_TC_REV_STAGES
#define QON_TC_REV_STAGES 2       // experiment switch (scripts/build_tc_variant.sh): B-image ring stages per tile in the reverse kernel
#endif

constexpr int kTcImgBytes = 16384;             // per block: B_hi (8 KB) | B_lo (8 KB)
constexpr float kTcSA = 32768.f;               // state scale  (|amplitude| <= 1 -> f16 normal range)
constexpr float kTcSB = 1.f;                   // matrix scale: 1 keeps a GEMM's output at the operand scale (lo parts of
                                               // small entries are f16 subnormals: absolute error 2^-25, norm-relative 3e-8)

// byte offset of element (nn, kk) of a 64 x 64 f16 operand stored K-major without swizzle:
// core matrices of 8 rows x 16 bytes, K-adjacent core matrices 128 B apart (LBO), row groups 1024 B apart (SBO)
__host__ __device__ constexpr int tc_b_offset(int nn, int kk) {
    return (nn >> 3) * 1024 + (kk >> 3) * 128 + (nn & 7) * 16 + (kk & 7) * 2;
}

// ---------------------------------------------------------------------------------------------------------
// prep: block matrices in fp64 -> f16 hi/lo shared-memory images.  grid = K CTAs of 512 threads: lane j owns basis
// column j, warp pi one of the 16 row pairs of a gate (a barrier between gates).
//   M_k = [H] W_k H,  W_k = prod_{sublayers} Ring * (x)_q U[s,q]   (leading H dropped for the last block)
// Real form consumed by the GEMM (row vector x B):  out[2i + c'] = sum_{j,c} in[2j + c] * B[2j + c][2i + c'],
//   B[2j][2i] = Re M_ij, B[2j+1][2i] = -Im M_ij, B[2j][2i+1] = Im M_ij, B[2j+1][2i+1] = Re M_ij,
// stored as Bt[nn = 2i + c'][kk = 2j + c] (K-major).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_prep_body(const float* __restrict__ w, int K, const DepthPack& dp,
                                             unsigned char* __restrict__ bimg, int k, double (*vr)[33], double (*vi)[33],
                                             double (*uu)[4]) {
    constexpr int n = 5, N = 32;
    const int j = threadIdx.x & 31, pi = threadIdx.x >> 5;      // lane = basis column, warp = one of the 16 row pairs of a gate
    int s0 = 0;
    for (int kk = 0; kk < k; ++kk) s0 += dp.d[kk];
    const int d = dp.d[k];
    const double r = 0.17677669529663688110;   // 1 / sqrt(32)
    for (int z = pi; z < N; z += 16) {
        vr[z][j] = (__popc(z & j) & 1) ? -r : r;
        vi[z][j] = 0.0;
    }
    __syncthreads();
    for (int s = s0; s < s0 + d; ++s) {
        // the sublayer's five fused rotations U = RY(c) RZ(b) RY(a) = [[al, -conj(be)], [be, conj(al)]]: lane q computes
        // U[s, q] once (fp64 sin/cos), every lane reads it
        __syncthreads();
        if (pi == 0 && j < n) {
            const double a = (double)w[((int64_t)s * 3 + 0) * n + j];
            const double b = (double)w[((int64_t)s * 3 + 1) * n + j];
            const double c = (double)w[((int64_t)s * 3 + 2) * n + j];
            double sa, ca, sb, cb, sc, cc;
            sincos(0.5 * a, &sa, &ca);
            sincos(0.5 * b, &sb, &cb);
            sincos(0.5 * c, &sc, &cc);
            uu[j][0] = cb * (cc * ca - sc * sa); uu[j][1] = -sb * (cc * ca + sc * sa);
            uu[j][2] = cb * (sc * ca + cc * sa); uu[j][3] = sb * (cc * sa - sc * ca);
        }
        __syncthreads();
        for (int q = 0; q < n; ++q) {
            const double ar = uu[q][0], ai = uu[q][1], br = uu[q][2], bi = uu[q][3];
            const int z = ((pi >> q) << (q + 1)) | (pi & ((1 << q) - 1)), z1 = z | (1 << q);
            const double x0r = vr[z][j], x0i = vi[z][j], x1r = vr[z1][j], x1i = vi[z1][j];
            vr[z][j] = ar * x0r - ai * x0i - br * x1r - bi * x1i;
            vi[z][j] = ar * x0i + ai * x0r - br * x1i + bi * x1r;
            vr[z1][j] = br * x0r - bi * x0i + ar * x1r + ai * x1i;
            vi[z1][j] = br * x0i + bi * x0r + ar * x1i - ai * x1r;
            __syncthreads();
        }
        for (int i = 0; i < n; ++i) {   // CNOT ring: control (i+1)%n -> target i, i ascending
            const int c = (i + 1) % n;
            const int z = ((pi >> i) << (i + 1)) | (pi & ((1 << i) - 1)), z1 = z | (1 << i);
            if ((z >> c) & 1) {
                double t = vr[z][j]; vr[z][j] = vr[z1][j]; vr[z1][j] = t;
                t = vi[z][j]; vi[z][j] = vi[z1][j]; vi[z1][j] = t;
            }
            __syncthreads();
        }
    }
    if (k < K - 1) {   // back to the Hadamard basis for the next block's diagonal encoding layer
        const double h = 0.70710678118654752440;
        for (int q = 0; q < n; ++q) {
            const int z = ((pi >> q) << (q + 1)) | (pi & ((1 << q) - 1)), z1 = z | (1 << q);
            const double x0r = vr[z][j], x0i = vi[z][j], x1r = vr[z1][j], x1i = vi[z1][j];
            vr[z][j] = h * (x0r + x1r); vi[z][j] = h * (x0i + x1i);
            vr[z1][j] = h * (x0r - x1r); vi[z1][j] = h * (x0i - x1i);
            __syncthreads();
        }
    }
    __half* hi = reinterpret_cast<__half*>(bimg + (size_t)k * kTcImgBytes);
    __half* lo = hi + 4096;
    auto put = [&](int nn, int kk, double v) {
        const double vs = v * (double)kTcSB;
        const __half h = __double2half(vs);
        const __half l = __double2half(vs - (double)__half2float(h));
        const int o = tc_b_offset(nn, kk) >> 1;
        hi[o] = h;
        lo[o] = l;
    };
    for (int i = pi; i < N; i += 16) {
        const double re = vr[i][j], im = vi[i][j];
        put(2 * i, 2 * j, re);
        put(2 * i, 2 * j + 1, -im);
        put(2 * i + 1, 2 * j, im);
        put(2 * i + 1, 2 * j + 1, re);
    }
}
__global__ void __launch_bounds__(512) tc_prep_kernel(const float* __restrict__ w, int K, DepthPack dp,
                                                     unsigned char* __restrict__ bimg) {
    __shared__ double vr[32][33], vi[32][33], uu[5][4];
    tc_prep_body(w, K, dp, bimg, blockIdx.x, vr, vi, uu);
}

// ---------------------------------------------------------------------------------------------------------
// phases of one encoding layer for the 16 basis states with qubit 4 = 0 (the other 16 are conjugates:
// p[31 - z] = conj(p[z])), times `scale`:  p[z] = scale * prod_q (cos(t_q/2) -/+ i sin(t_q/2)),  - for z_q = 0;
// kept as packed (re, im) pairs: 24 complex products of 2 FFMA2 each
// ---------------------------------------------------------------------------------------------------------
// packed complex helpers: a (packed re, im) times the complex scalar (br, bi), or conj(a) times it — 2 FFMA2 each
__device__ __forceinline__ u64 tc_cmul(u64 a, float br, float bi) { return fma2<2>(bi, a, mul2<0>(br, a)); }
__device__ __forceinline__ u64 tc_cmul_conj(u64 a, float br, float bi) { return fma2<1>(bi, a, mul2<5>(br, a)); }

__device__ __forceinline__ void tc_phase_table(const float (&th)[5], float scale, u64 (&p)[16]) {
    float s[5], c[5];
    // five independent branch-free evaluations (one basic block: the scheduler interleaves them); the rare huge
    // angle takes the accurate slow path afterwards
    float big = 0.f;
#pragma unroll
    for (int q = 0; q < 5; ++q) { sincos_half_fast(th[q], s[q], c[q]); big = fmaxf(big, fabsf(th[q])); }
    if (__builtin_expect(big > 65536.0f, 0)) {
#pragma unroll
        for (int q = 0; q < 5; ++q) {
#if QON_TC_SLOW_INLINE
            sincos_half(th[q], s[q], c[q]);
#else
            const float2 sc = sincos_half_slow(th[q]);
            s[q] = sc.x;
            c[q] = sc.y;
#endif
        }
    }
    // qubits 0..2: l[z2 z1 z0] for z2 = 0; l[7 - j] = conj(l[j]).  e_q(z_q) = cos(t_q/2) -/+ i sin(t_q/2)
    u64 l[4];
    {
        const u64 b0 = tc_cmul(pack2(c[0], -s[0]), c[1], -s[1]);      // z0 = 0, z1 = 0
        const u64 b1 = tc_cmul(pack2(c[0], s[0]), c[1], -s[1]);       // z0 = 1, z1 = 0
        // z1 = 1: (z0 = 0) = conj(b1), (z0 = 1) = conj(b0)
        l[0] = tc_cmul(b0, c[2], -s[2]);
        l[1] = tc_cmul(b1, c[2], -s[2]);
        l[2] = tc_cmul_conj(b1, c[2], -s[2]);
        l[3] = tc_cmul_conj(b0, c[2], -s[2]);
    }
    // qubits 3, 4 with z4 = 0: h[z3], carrying the scale
    const float c4 = c[4] * scale, s4 = s[4] * scale;
    float h0r, h0i, h1r, h1i;
    unpack2(tc_cmul(pack2(c[3], -s[3]), c4, -s4), h0r, h0i);
    unpack2(tc_cmul(pack2(c[3], s[3]), c4, -s4), h1r, h1i);
#pragma unroll
    for (int z = 0; z < 16; ++z) {
        const int j = z & 7;
        const float hr = z < 8 ? h0r : h1r, hi = z < 8 ? h0i : h1i;
        p[z] = j < 4 ? tc_cmul(l[j], hr, hi) : tc_cmul_conj(l[7 - j], hr, hi);
    }
}

}  // namespace qon
