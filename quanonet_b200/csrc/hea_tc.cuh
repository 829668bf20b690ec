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
// Data flow per CTA (one per SM): 4 compute warpgroups, each owning one 128-sample tile whose state lives in
// TMEM (lane = sample): D (64 f32 columns: re/im interleaved) and the A operand (64 columns: 32 of f16x2 hi,
// 32 of f16x2 lo).  Thread = sample: tcgen05.ld its D row -> phases -> split -> tcgen05.st its A row; one elected
// thread per warpgroup then issues the 12 MMAs of the block against the block's B image in shared memory and
// commits to the warpgroup's mbarrier.  A 17th warp streams the B images (16 KB per block: [hi | lo], laid out
// by the prep kernel exactly as the no-swizzle K-major shared-memory descriptor expects) through a 4-stage ring
// with 1-D bulk async copies.  While one warpgroup waits for its MMAs the other three run their CUDA-core part.
#pragma once
#include "hea_common.cuh"
#include "ffma2.cuh"
#include "tc_ptx.cuh"
#include <cuda_fp16.h>

namespace qon {

constexpr int kTcStages = 4;
constexpr int kTcImgBytes = 16384;             // per block: B_hi (8 KB) | B_lo (8 KB)
constexpr int kTcThreads = 17 * 32;
constexpr float kTcSA = 32768.f;               // state scale  (|amplitude| <= 1 -> f16 normal range)
constexpr float kTcSB = 1.f;                   // matrix scale: 1 keeps a GEMM's output at the operand scale (lo parts of
                                               // small entries are f16 subnormals: absolute error 2^-25, norm-relative 3e-8)

// byte offset of element (nn, kk) of a 64 x 64 f16 operand stored K-major without swizzle:
// core matrices of 8 rows x 16 bytes, K-adjacent core matrices 128 B apart (LBO), row groups 1024 B apart (SBO)
__host__ __device__ constexpr int tc_b_offset(int nn, int kk) {
    return (nn >> 3) * 1024 + (kk >> 3) * 128 + (nn & 7) * 16 + (kk & 7) * 2;
}

// ---------------------------------------------------------------------------------------------------------
// prep: block matrices in fp64 -> f16 hi/lo shared-memory images.  grid = K CTAs of 32 threads; thread j
// pushes basis column j through the block.
//   M_k = [H] W_k H,  W_k = prod_{sublayers} Ring * (x)_q U[s,q]   (leading H dropped for the last block)
// Real form consumed by the GEMM (row vector x B):  out[2i + c'] = sum_{j,c} in[2j + c] * B[2j + c][2i + c'],
//   B[2j][2i] = Re M_ij, B[2j+1][2i] = -Im M_ij, B[2j][2i+1] = Im M_ij, B[2j+1][2i+1] = Re M_ij,
// stored as Bt[nn = 2i + c'][kk = 2j + c] (K-major).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) tc_prep_kernel(const float* __restrict__ w, int K, DepthPack dp,
                                                     unsigned char* __restrict__ bimg) {
    constexpr int n = 5, N = 32;
    __shared__ double vr[N][N + 1], vi[N][N + 1];
    const int k = blockIdx.x, j = threadIdx.x;
    int s0 = 0;
    for (int kk = 0; kk < k; ++kk) s0 += dp.d[kk];
    const int d = dp.d[k];
    const double r = 0.17677669529663688110;   // 1 / sqrt(32)
    for (int z = 0; z < N; ++z) {
        vr[z][j] = (__popc(z & j) & 1) ? -r : r;
        vi[z][j] = 0.0;
    }
    for (int s = s0; s < s0 + d; ++s) {
        for (int q = 0; q < n; ++q) {
            const double a = (double)w[((int64_t)s * 3 + 0) * n + q];
            const double b = (double)w[((int64_t)s * 3 + 1) * n + q];
            const double c = (double)w[((int64_t)s * 3 + 2) * n + q];
            double sa, ca, sb, cb, sc, cc;
            sincos(0.5 * a, &sa, &ca);
            sincos(0.5 * b, &sb, &cb);
            sincos(0.5 * c, &sc, &cc);
            // U = RY(c) RZ(b) RY(a) = [[al, -conj(be)], [be, conj(al)]]
            const double ar = cb * (cc * ca - sc * sa), ai = -sb * (cc * ca + sc * sa);
            const double br = cb * (sc * ca + cc * sa), bi = sb * (cc * sa - sc * ca);
            for (int z = 0; z < N; ++z) {
                if (z & (1 << q)) continue;
                const int z1 = z | (1 << q);
                const double x0r = vr[z][j], x0i = vi[z][j], x1r = vr[z1][j], x1i = vi[z1][j];
                vr[z][j] = ar * x0r - ai * x0i - br * x1r - bi * x1i;
                vi[z][j] = ar * x0i + ai * x0r - br * x1i + bi * x1r;
                vr[z1][j] = br * x0r - bi * x0i + ar * x1r + ai * x1i;
                vi[z1][j] = br * x0i + bi * x0r + ar * x1i - ai * x1r;
            }
        }
        for (int i = 0; i < n; ++i) {   // CNOT ring: control (i+1)%n -> target i, i ascending
            const int c = (i + 1) % n;
            for (int z = 0; z < N; ++z) {
                if (((z >> c) & 1) && !((z >> i) & 1)) {
                    const int z1 = z | (1 << i);
                    double t = vr[z][j]; vr[z][j] = vr[z1][j]; vr[z1][j] = t;
                    t = vi[z][j]; vi[z][j] = vi[z1][j]; vi[z1][j] = t;
                }
            }
        }
    }
    if (k < K - 1) {   // back to the Hadamard basis for the next block's diagonal encoding layer
        const double h = 0.70710678118654752440;
        for (int q = 0; q < n; ++q)
            for (int z = 0; z < N; ++z) {
                if (z & (1 << q)) continue;
                const int z1 = z | (1 << q);
                const double x0r = vr[z][j], x0i = vi[z][j], x1r = vr[z1][j], x1i = vi[z1][j];
                vr[z][j] = h * (x0r + x1r); vi[z][j] = h * (x0i + x1i);
                vr[z1][j] = h * (x0r - x1r); vi[z1][j] = h * (x0i - x1i);
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
    for (int i = 0; i < N; ++i) {
        const double re = vr[i][j], im = vi[i][j];
        put(2 * i, 2 * j, re);
        put(2 * i, 2 * j + 1, -im);
        put(2 * i + 1, 2 * j, im);
        put(2 * i + 1, 2 * j + 1, re);
    }
}

// ---------------------------------------------------------------------------------------------------------
// phases of one encoding layer for the 16 basis states with qubit 4 = 0 (the other 16 are conjugates:
// p[31 - z] = conj(p[z])), times `scale`:  p[z] = scale * prod_q (cos(t_q/2) -/+ i sin(t_q/2)),  - for z_q = 0
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_cmul(float ar, float ai, float br, float bi, float& cr, float& ci) {
    cr = fmaf(-ai, bi, ar * br);
    ci = fmaf(ai, br, ar * bi);
}

__device__ __forceinline__ void tc_phase_table(const float (&th)[5], float scale, float (&pr)[16], float (&pi)[16]) {
    float s[5], c[5];
    // five independent branch-free evaluations (one basic block: the scheduler interleaves them); the rare huge
    // angle takes the accurate slow path afterwards
    float big = 0.f;
#pragma unroll
    for (int q = 0; q < 5; ++q) { sincos_half_fast(th[q], s[q], c[q]); big = fmaxf(big, fabsf(th[q])); }
    if (__builtin_expect(big > 65536.0f, 0)) {
#pragma unroll
        for (int q = 0; q < 5; ++q) sincos_half(th[q], s[q], c[q]);
    }
    // qubits 0..2: l[z2 z1 z0]; l[7 - j] = conj(l[j])
    float lr[4], li[4];
    {
        float b0r, b0i, b1r, b1i;
        tc_cmul(c[0], -s[0], c[1], -s[1], b0r, b0i);     // z0 = 0, z1 = 0
        tc_cmul(c[0], s[0], c[1], -s[1], b1r, b1i);      // z0 = 1, z1 = 0
        // z1 = 1: (z0 = 0) = conj(b1), (z0 = 1) = conj(b0)
        tc_cmul(b0r, b0i, c[2], -s[2], lr[0], li[0]);
        tc_cmul(b1r, b1i, c[2], -s[2], lr[1], li[1]);
        tc_cmul(b1r, -b1i, c[2], -s[2], lr[2], li[2]);
        tc_cmul(b0r, -b0i, c[2], -s[2], lr[3], li[3]);
    }
    // qubits 3, 4 with z4 = 0: h[z3]
    float h0r, h0i, h1r, h1i;
    const float c4 = c[4] * scale, s4 = s[4] * scale;
    tc_cmul(c[3], -s[3], c4, -s4, h0r, h0i);
    tc_cmul(c[3], s[3], c4, -s4, h1r, h1i);
#pragma unroll
    for (int z = 0; z < 16; ++z) {
        const int j = z & 7;
        const float xr = j < 4 ? lr[j] : lr[7 - j];
        const float xi = j < 4 ? li[j] : -li[7 - j];
        if (z < 8) tc_cmul(xr, xi, h0r, h0i, pr[z], pi[z]);
        else tc_cmul(xr, xi, h1r, h1i, pr[z], pi[z]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// forward kernel.  ENC = 0: angles x given; 1: angles formed in-kernel from (u0, u1, fw, fb) as in hea_reg.cuh
// ---------------------------------------------------------------------------------------------------------
template <int ENC, bool DBG>
__global__ void __launch_bounds__(kTcThreads, 1)
hea_tc_fwd_kernel(const HeaParams<float> p, const unsigned char* __restrict__ bimg, float* dbg, int* err) {
    constexpr int NQ = 5;
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ __align__(8) uint64_t bar_full[kTcStages], bar_empty[kTcStages], bar_d[4];
    __shared__ uint32_t tmem_base_s;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ntiles = (p.B + 127) / 128;
    const int64_t rounds = (ntiles + (int64_t)gridDim.x * 4 - 1) / ((int64_t)gridDim.x * 4);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kTcStages; ++i) {
            tc::mbar_init(tc::smem_u32(&bar_full[i]), 1);
            tc::mbar_init(tc::smem_u32(&bar_empty[i]), 4);
        }
        for (int i = 0; i < 4; ++i) tc::mbar_init(tc::smem_u32(&bar_d[i]), 1);
        tc::mbar_fence_init();
    }
    if (warp == 1) tc::tmem_alloc512(tc::smem_u32(&tmem_base_s));
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 16) {
        // ------------------------------------------------ B-image producer
        if (lane == 0) {
            const int64_t total = rounds * p.K;
            int kblk = 0;
            for (int64_t g = 0; g < total; ++g) {
                const int stage = (int)(g % kTcStages);
                if (g >= kTcStages && !tc::mbar_wait(tc::smem_u32(&bar_empty[stage]), (uint32_t)((g / kTcStages - 1) & 1), err))
                    break;
                tc::mbar_expect_tx(tc::smem_u32(&bar_full[stage]), kTcImgBytes);
                tc::bulk_g2s(tc::smem_u32(tc_smem + (size_t)stage * kTcImgBytes), bimg + (size_t)kblk * kTcImgBytes,
                             kTcImgBytes, tc::smem_u32(&bar_full[stage]));
                if (++kblk == p.K) kblk = 0;
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------ compute warpgroups
        const int wg = warp >> 2, quarter = warp & 3;
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        const uint32_t tD = tmem_base + lane_sel + (uint32_t)wg * 128u;   // this warp's 32 lanes of the tile
        const uint32_t tA = tD + 64u;
        const uint32_t mD = tmem_base + (uint32_t)wg * 128u, mA = mD + 64u;   // MMA view: all 128 lanes
        const uint32_t bar_mine = tc::smem_u32(&bar_d[wg]);
        const bool issuer = quarter == 0 && lane == 0;
        constexpr uint32_t idesc = tc::idesc_f16(128, 64);
        uint32_t dpar = 0;
        int64_t g = 0;
        bool dead = false;   // a wait timed out: keep the barrier protocol, skip the work

        for (int64_t round = 0; round < rounds; ++round) {
            const int64_t tile = (round * gridDim.x + blockIdx.x) * 4 + wg;
            const int64_t b = tile * 128 + quarter * 32 + lane;
            const bool valid = b < p.B;
            const int64_t bc = valid ? b : p.B - 1;
            const float* xrow = ENC == 0 ? p.x + bc * p.ldx : nullptr;
            const float* u0row = ENC != 0 && p.u0 ? p.u0 + bc * p.ldu0 : nullptr;
            const float* u1row = ENC != 0 ? p.u1 + bc * p.ldu1 : nullptr;
            auto load_angles = [&](int k, float(&th)[NQ]) {
                if constexpr (ENC == 0) {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) th[q] = __ldg(xrow + (int64_t)k * NQ + q);
                } else {
                    const float* ur = k < p.K0 ? u0row : u1row;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const int col = k * NQ + q;
                        const float u = __ldg(ur + __ldg(p.uidx + col));
                        th[q] = fmaf(u, __ldg(p.fw + col), p.fb ? __ldg(p.fb + col) : 0.f);
                    }
                }
            };
            float th[NQ];
            load_angles(0, th);
            for (int k = 0; k < p.K; ++k, ++g) {
                float thn[NQ];
                load_angles(k + 1 < p.K ? k + 1 : k, thn);
                // phases of this block's encoding layer; the GEMM output carries sA*sB, the operand wants sA
                float pr[16], pi[16];
                tc_phase_table(th, k == 0 ? 1.f : 1.f / kTcSB, pr, pi);
                if (k > 0) {
                    if (!dead && !tc::mbar_wait(bar_mine, dpar, err)) dead = true;
                    dpar ^= 1u;
                    tc::tc_fence_after();
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[16];
                    if (k > 0) {
                        tc::tmem_ld16(tD + 16u * c, r);
                        tc::tmem_wait_ld();
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            r[2 * i] = __float_as_uint(kTcSA * 0.17677669529663688110f);
                            r[2 * i + 1] = 0u;
                        }
                    }
                    if constexpr (DBG) {
                        if (dbg && k > 0 && blockIdx.x == 0 && wg == 0 && round == 0)
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                dbg[((size_t)(k - 1) * 128 + quarter * 32 + lane) * 64 + 16 * c + i] = __uint_as_float(r[i]);
                    }
                    uint32_t ahi[8], alo[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int z = 8 * c + i;
                        const u64 v = pack2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
                        u64 nv;
                        if (z < 16) {
                            nv = mul2<0>(pr[z], v);
                            nv = fma2<2>(pi[z], v, nv);
                        } else {
                            nv = mul2<0>(pr[31 - z], v);
                            nv = fma2<3>(pi[31 - z], v, nv);
                        }
                        float xr, xi;
                        unpack2(nv, xr, xi);
                        const float hr = __uint_as_float(__float_as_uint(xr) & 0xFFFFE000u);
                        const float hi_ = __uint_as_float(__float_as_uint(xi) & 0xFFFFE000u);
                        ahi[i] = tc::cvt_f16x2(hr, hi_);
                        alo[i] = tc::cvt_f16x2(xr - hr, xi - hi_);
                    }
                    tc::tmem_st8(tA + 8u * c, ahi);
                    tc::tmem_st8(tA + 32u + 8u * c, alo);
                }
                tc::tmem_wait_st();
                tc::tc_fence_before();
                asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");
                if (issuer) {
                    tc::tc_fence_after();
                    const int stage = (int)(g % kTcStages);
                    if (!dead && !tc::mbar_wait(tc::smem_u32(&bar_full[stage]), (uint32_t)((g / kTcStages) & 1), err)) dead = true;
                    const uint32_t sb = tc::smem_u32(tc_smem + (size_t)stage * kTcImgBytes);
                    if (!dead) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            tc::mma_f16_ts(mD, mA + 8u * j, tc::smem_desc_kmajor(sb + 256u * j, 128u, 1024u), idesc, j > 0);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            tc::mma_f16_ts(mD, mA + 8u * j, tc::smem_desc_kmajor(sb + 8192u + 256u * j, 128u, 1024u), idesc, 1u);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            tc::mma_f16_ts(mD, mA + 32u + 8u * j, tc::smem_desc_kmajor(sb + 256u * j, 128u, 1024u), idesc, 1u);
                    }
                    tc::mma_commit(bar_mine);
                    tc::mma_commit(tc::smem_u32(&bar_empty[stage]));
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q) th[q] = thn[q];
            }
            // ---------------- expectation value of the tile's final state
            if (!dead && !tc::mbar_wait(bar_mine, dpar, err)) dead = true;
            dpar ^= 1u;
            tc::tc_fence_after();
            float e = 0.f, nrm = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t r[16];
                tc::tmem_ld16(tD + 16u * c, r);
                tc::tmem_wait_ld();
                if constexpr (DBG) {
                    if (dbg && blockIdx.x == 0 && wg == 0 && round == 0)
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            dbg[((size_t)(p.K - 1) * 128 + quarter * 32 + lane) * 64 + 16 * c + i] = __uint_as_float(r[i]);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float re = __uint_as_float(r[2 * i]), im = __uint_as_float(r[2 * i + 1]);
                    const float pz = fmaf(re, re, im * im);
                    e = fmaf(__ldg(p.hdiag + 8 * c + i), pz, e);
                    nrm += pz;
                }
            }
            // every thread of the warpgroup has drained D before the next tile's first MMA may overwrite it:
            // that MMA is issued after the bar.sync of the next block 0, which every thread reaches after this point
            // The exact state has unit norm.  The tensor cores accumulate with truncation, which shrinks every
            // amplitude by the same ~4.7e-7 per GEMM (measured: -4.76e-7 +- 0.7e-7 over 128 samples); dividing by the
            // computed norm removes that coherent drift (5e-5 over 60 blocks -> 2e-6) and the operand scales.
            float res = e / nrm;
            if (__ldcg(err) != 0) res = __int_as_float(0x7fc00000);   // a barrier wait timed out: poison, never guess
            if (valid && p.out) p.out[b] = res;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc512(tmem_base);
}

}  // namespace qon
