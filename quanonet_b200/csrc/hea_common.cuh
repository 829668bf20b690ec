// Shared definitions for the HEA statevector kernels (sm_100a).
//
// Circuit being simulated (reference: core/quantum_circuits_tq.py:65-104), in the
// canonical form the C-ABI takes: K blocks; block k = one RX data-encoding layer (n per-sample angles
// x[b, k*n + q]) followed by depth[k] >= 1 ansatz sublayers; sublayer s = on every qubit q the shared
// single-qubit unitary U[s,q] = RY(w[s,2,q]) RZ(w[s,1,q]) RY(w[s,0,q]), then the CNOT ring
// control=(i+1)%n -> target=i for i = 0..n-1 (skipped when n == 1).
//
// Amplitude index convention: qubit q is bit q of the amplitude index (qubit 0 = LSB).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qon {

template <typename T> struct alignas(4 * sizeof(T)) Vec4 { T x, y, z, w; };

// Per (sublayer s, qubit q) tables written by the prep kernel.
//   ucoef[s*n+q] = (Re a, Im a, Re b, Im b) with U[s,q] = [[a, -conj(b)], [b, conj(a)]]   (SU(2))
//   rcoef[s*n+q] = (rX, rY, rZ, 0): A X A^dagger = rX X + rY Y + rZ Z for A = U[s,q]; turns the
//                  Pauli moments measured after the fused gate U*RX(theta) into dE/dtheta.
template <typename T>
struct HeaParams {
    const T* x;          // (B, n*K) encoding angles, row stride ldx elements
    int64_t ldx;
    int64_t B;
    T* out;              // (B,) expectation values
    const T* gout;       // (B,) upstream gradient dL/dout          [grad only; unused when target != null]
    const T* target;     // (B,) regression target y: the kernel forms the MSE upstream gradient itself,
                         //      g_b = gscale * (out_b + bias - y_b), and writes it to gbuf   [optional]
    const T* bias;       // device scalar added to out inside the residual (may be null = 0)
    T* gbuf;             // (B,) where g_b is written when target != null
    T gscale;
    T* gx;               // (B, n*K) dL/dx, row stride ldgx; may be null [grad only]
    int64_t ldgx;
    const Vec4<T>* ucoef;
    const Vec4<T>* rcoef;
    const T* hdiag;      // (2^n,) diagonal of H in LSB0 order (pauli == 0)
    const int* depth;    // (K,) device copy of depth_per_block
    T* mpart;            // (rows, rowlen) per-warp partial sums [grad only]; row = [S*VP Pauli moments |
                         //   K*FVP frequency-layer gradient slots | sum g | sum residual^2 | pad]
    int64_t rowlen;
    int K;
    int S;
    int pauli;           // 0: diagonal table; 1: offset + coeff*sum X_i; 2: offset + coeff*sum Y_i
    T offset, coeff;
    // Fused encoding ("ENC" kernels): the angle of column c is  fw[c] * u_src(c)[b, uidx[c]] + fb[c]  with
    // src(c) = 0 for blocks k < K0 and 1 otherwise — the frequency layers of core/models_pt.py:14-68 and the
    // trunk-first concatenation of :163-164, evaluated in-kernel so x and grad_x are never materialised.
    const T* u0;         // (B, in0) rows, stride ldu0  (QuanONet: trunk input; HEAQNN: unused, K0 = 0)
    const T* u1;         // (B, in1) rows, stride ldu1  (QuanONet: branch input; HEAQNN: the input)
    int64_t ldu0, ldu1;
    int K0;
    const int* uidx;     // (n*K,) source-row index of every column (prep kernel: local column % in)
    const T* fw;         // (n*K,) frequency weights (fixed-scale mode: the constant scale)
    const T* fb;         // (n*K,) frequency bias, or null
    int in0, in1;        // input widths of the two sources (uidx[c] = local column % in): kernels that want the index
                         // without a dependent load recompute it
};

// slots per block for the frequency-layer gradients: (d/dfw, d/dfb) per qubit, padded for the butterfly
__host__ __device__ constexpr int freq_slots(int n) { return 2 * n <= 2 ? 2 : (2 * n <= 4 ? 4 : (2 * n <= 8 ? 8 : (2 * n <= 16 ? 16 : 32))); }

__host__ __device__ constexpr int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// number of moment slots per sublayer (3 per qubit, padded to a power of two for the butterfly)
__host__ __device__ constexpr int moment_slots(int n) { return next_pow2(3 * n) < 4 ? 4 : next_pow2(3 * n); }

// sin and cos of t/2 in fp32, ~0.6 ulp and unbiased.  CUDA's sincosf() is not used: on this path its
// (<= 2 ulp) error is systematically signed, so c^2 + s^2 - 1 drifts the state norm LINEARLY over the
// ~300 RX gates of a circuit and costs ~2.5x in end-to-end accuracy (measured: 7.6e-6 -> 2.9e-6
// rel-L2 vs the fp64 oracle at Q5 Net40-2-20-2).  Cody-Waite reduction by pi/2 with FMAs, minimax
// polynomials on [-pi/4, pi/4] (sin: rel 3.6e-9, cos: abs 9.6e-11), quadrant fix-up by bit tricks.
// branch-free core, valid for |t/2| <= 32768 (three-part Cody-Waite reduction stays exact up to there)
__device__ __forceinline__ void sincos_half_fast(float t, float& s, float& c) {
    const float h = 0.5f * t;
    const float kf = rintf(h * 0.636619772367581343f);
    const int k = __float2int_rn(kf);
    float r = fmaf(kf, -1.57079601287841796875f, h);          // pi/2 split in three parts
    r = fmaf(kf, -3.1391647326017846e-07f, r);
    r = fmaf(kf, -5.3903029534742384e-15f, r);
    const float r2 = r * r;
    float p = fmaf(-0.00019516654723058025f, r2, 0.008332173533202262f);
    p = fmaf(p, r2, -0.16666654873265133f);
    const float sr = fmaf(r2 * r, p, r);
    float q = fmaf(2.4437728180492216e-05f, r2, -0.0013887361431077323f);
    q = fmaf(q, r2, 0.041666646747316696f);
    q = fmaf(q, r2, -0.5f);
    const float cr = fmaf(q, r2, 1.0f);
    const float ss = (k & 1) ? cr : sr;
    const float cc = (k & 1) ? sr : cr;
    s = __int_as_float(__float_as_int(ss) ^ ((k & 2) << 30));
    c = __int_as_float(__float_as_int(cc) ^ (((k + 1) & 2) << 30));
}
__device__ __forceinline__ void sincos_half(float t, float& s, float& c) {
    if (__builtin_expect(fabsf(t) > 65536.0f, 0)) { sincosf(0.5f * t, &s, &c); return; }   // Payne-Hanek territory
    sincos_half_fast(t, s, c);
}
// the rare huge angle, out of line: one copy of the Payne-Hanek path (with its local-memory table) per kernel
static __device__ __noinline__ float2 sincos_half_slow(float t) {
    float s, c;
    sincosf(0.5f * t, &s, &c);
    return make_float2(s, c);
}
__device__ __forceinline__ void sincos_half(double t, double& s, double& c) { sincos(0.5 * t, &s, &c); }

// depth_per_block as a kernel parameter (constant bank: loop bounds read from it are provably warp-uniform)
constexpr int kMaxBlocks = 1024;
struct DepthPack { unsigned char d[kMaxBlocks]; };

// geometry of the fp32 shared-memory tier (hea_smem.cuh)
constexpr int kSmemMinN = 6, kSmemMaxN = 13, kSmemW = 5, kSmemMaxP = 3;
struct SmemGeom {
    int n, P;
    int lo[kSmemMaxP], gm[kSmemMaxP];
    int tps_log2;      // log2(threads per sample) = n - 5
    int spc;           // samples per CTA = THREADS >> tps_log2
    int region_bytes;  // 8 << n
    int vp;            // moment slots per sublayer in a partial row (>= 3n)
};

__device__ __forceinline__ Vec4<float> ldg4(const Vec4<float>* p) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    return Vec4<float>{v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ Vec4<double> ldg4(const Vec4<double>* p) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    const double2 b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    return Vec4<double>{a.x, a.y, b.x, b.y};
}

__device__ __forceinline__ float fma_(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return fma(a, b, c); }

}  // namespace qon
