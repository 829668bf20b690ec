// Tensor-core tier, second kernel: forward AND the one-kernel adjoint backward (n = 5, fp32 results).
//
// Forward as in hea_tc.cuh (block unitaries M_k = [H] W_k H as f16x3 GEMMs, RX layers as diagonal phases in
// the Hadamard basis).  Reverse sweep (what loss.backward() at solvers/solver_pt.py:235 produces, by adjoint
// differentiation): one GEMM per SUBLAYER un-applies it on psi and lam at once,
//     G_s = [H if s is the first sublayer of its block] R_s^+ Ring^+ [H if the cut is held in the Hadamard basis],
// so the registers always hold the pair (psi, lam) AFTER a sublayer's CNOT ring.  The three Pauli moments per qubit
// that the finalize kernel turns into angle gradients are defined BEFORE the ring; CNOTs are Clifford, so they are
// measured as Pauli strings T = Ring P_q Ring^+ (csrc/tc_strings.cuh, generated and verified by
// scripts/gen_tc_strings.py; the whole sweep is emulated against the fp64 oracle in tests/harness/tc_emulate_bwd.py).
// The gradient of an encoding angle is a Z-type (diagonal) moment in the Hadamard basis, taken right after the
// GEMM of a block's first sublayer, before the conjugate phases are applied.
//
// Roles in a CTA (1 per SM): NT tiles of 128 samples (NT = 4 forward-only, 2 with gradients); tile t is owned by
// compute warps 4t..4t+3 (thread = sample = TMEM lane) and by MMA warp 4*NT + t, whose elected thread streams the
// step's B image through a private ring (1-D bulk async copies) and issues the tcgen05.mma's.  Hand-offs are
// mbarriers only: a_ready[t] (compute -> MMA: "A operand written"), d_ready[t] (tcgen05.commit: "D complete").
#pragma once
#include "hea_reg.cuh"
#include "hea_tc.cuh"
#include "tc_strings.cuh"

namespace qon {

template <bool GRAD> struct TcGeom {
    static constexpr int NT = GRAD ? 2 : 4;                 // tiles per CTA
    static constexpr int NS = GRAD ? 4 : 3;                 // B-image ring stages per tile
    static constexpr int COMPUTE_WARPS = 4 * NT;
    static constexpr int WARPS = COMPUTE_WARPS + 4;         // + one warpgroup hosting the NT MMA warps
    static constexpr int THREADS = WARPS * 32;
    static constexpr int TILE_COLS = GRAD ? 256 : 128;      // TMEM columns per tile
    static constexpr int SMEM = NT * NS * kTcImgBytes;
    // register hand-over (setmaxnreg) from the MMA warpgroup to the compute warps: 168 * 384 = 232 * 256 + 40 * 128
    static constexpr int REGS_COMPUTE = 232, REGS_MMA = 40;
};

// a (.) P(z) + c with the operand patterns of ffma2.cuh applied to the packed operand z
#define QON_P2V_DECL ".reg .b64 pb; .reg .f32 zx, zy, tx, ty; mov.b64 {zx, zy}, %2; "
#define QON_FMA2VP_CASE(N)                                                                             \
    if constexpr (PAT == N)                                                                            \
        asm("{ " QON_P2V_DECL QON_P2_##N "fma.rn.f32x2 %0, pb, %1, %3; }" : "=l"(d) : "l"(a), "l"(z), "l"(c));
template <int PAT>
__device__ __forceinline__ u64 fma2_vp(u64 a, u64 z, u64 c) {
    u64 d;
    QON_FMA2VP_CASE(0) QON_FMA2VP_CASE(2) QON_FMA2VP_CASE(3) QON_FMA2VP_CASE(6)
    return d;
}

__host__ __device__ constexpr bool tc_parity(int v) { return ((v ^ (v >> 1) ^ (v >> 2) ^ (v >> 3) ^ (v >> 4)) & 1) != 0; }

__device__ __forceinline__ u64 tc_pair(const uint32_t (&r)[64], int z) {
    return pack2(__uint_as_float(r[2 * z]), __uint_as_float(r[2 * z + 1]));
}

// ---------------------------------------------------------------------------------------------------------
// prep of the reverse images: grid = S CTAs of 32 threads, image s' = S-1-s (execution order of the sweep).
// Image of sublayer s (block k, sublayers s0 .. last): the COMPOSITE matrix taking the block's output cut to the cut
// after sublayer s-1, in the Hadamard basis:  C_s = H (R_s^+ Ring^+) ... (R_last^+ Ring^+) [H if k < K-1],
// so one operand split per block serves all its GEMMs.
// ---------------------------------------------------------------------------------------------------------
// per_block (hea_tc3.cuh): grid = K CTAs, image K-1-k = the whole-block un-apply matrix C_{s0(k)} = M_k^+.
__device__ __forceinline__ void tc_prep_rev_body(const float* __restrict__ w, int K, int S, const DepthPack& dp,
                                                 unsigned char* __restrict__ rimg, int per_block, int cta,
                                                 double (*vr)[33], double (*vi)[33], double (*uu)[4]) {
    constexpr int n = 5, N = 32;
    const int j = threadIdx.x & 31, pi = threadIdx.x >> 5;      // lane = basis column, warp = one of the 16 row pairs of a gate
    int s = cta, k = 0, s0 = 0;
    if (per_block) {
        k = cta;
        for (int kk = 0; kk < k; ++kk) s0 += dp.d[kk];
        s = s0;
    } else {
        while (s0 + dp.d[k] <= s) { s0 += dp.d[k]; ++k; }
    }
    const bool input_had = k < K - 1;   // the operand is the block's OUTPUT cut, held in the Hadamard basis except for the last block
    const double h = 0.70710678118654752440;
    auto fwht = [&]() {
        for (int q = 0; q < n; ++q) {
            const int z = ((pi >> q) << (q + 1)) | (pi & ((1 << q) - 1)), z1 = z | (1 << q);
            const double x0r = vr[z][j], x0i = vi[z][j], x1r = vr[z1][j], x1i = vi[z1][j];
            vr[z][j] = h * (x0r + x1r); vi[z][j] = h * (x0i + x1i);
            vr[z1][j] = h * (x0r - x1r); vi[z1][j] = h * (x0i - x1i);
            __syncthreads();
        }
    };
    for (int z = pi; z < N; z += 16) { vr[z][j] = z == j ? 1.0 : 0.0; vi[z][j] = 0.0; }
    __syncthreads();
    if (input_had) fwht();
    for (int ss = s0 + dp.d[k] - 1; ss >= s; --ss) {   // un-apply the block's sublayers last .. s
        for (int i = n - 1; i >= 0; --i) {   // Ring^+ : the CNOTs in reverse order
            const int c = (i + 1) % n;
            const int z = ((pi >> i) << (i + 1)) | (pi & ((1 << i) - 1)), z1 = z | (1 << i);
            if ((z >> c) & 1) {
                double t = vr[z][j]; vr[z][j] = vr[z1][j]; vr[z1][j] = t;
                t = vi[z][j]; vi[z][j] = vi[z1][j]; vi[z1][j] = t;
            }
            __syncthreads();
        }
        if (pi == 0 && j < n) {      // lane q computes U[ss, q] once
            const double a = (double)w[((int64_t)ss * 3 + 0) * n + j];
            const double b = (double)w[((int64_t)ss * 3 + 1) * n + j];
            const double c = (double)w[((int64_t)ss * 3 + 2) * n + j];
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
            // U^+ = [[conj(al), conj(be)], [-be, al]]
            const int z = ((pi >> q) << (q + 1)) | (pi & ((1 << q) - 1)), z1 = z | (1 << q);
            const double x0r = vr[z][j], x0i = vi[z][j], x1r = vr[z1][j], x1i = vi[z1][j];
            vr[z][j] = ar * x0r + ai * x0i + br * x1r + bi * x1i;
            vi[z][j] = ar * x0i - ai * x0r + br * x1i - bi * x1r;
            vr[z1][j] = -br * x0r + bi * x0i + ar * x1r - ai * x1i;
            vi[z1][j] = -br * x0i - bi * x0r + ar * x1i + ai * x1r;
            __syncthreads();
        }
    }
    fwht();     // EVERY cut of the reverse sweep is held in the Hadamard basis: one moment routine in the hot loop
    __half* hi = reinterpret_cast<__half*>(rimg + (size_t)(per_block ? K - 1 - k : S - 1 - s) * kTcImgBytes);
    __half* lo = hi + 4096;
    auto put = [&](int nn, int kk, double v) {
        const double vs = v * (double)kTcSB;
        const __half hh = __double2half(vs);
        const __half ll = __double2half(vs - (double)__half2float(hh));
        const int o = tc_b_offset(nn, kk) >> 1;
        hi[o] = hh;
        lo[o] = ll;
    };
    for (int i = pi; i < N; i += 16) {
        const double re = vr[i][j], im = vi[i][j];
        put(2 * i, 2 * j, re);
        put(2 * i, 2 * j + 1, -im);
        put(2 * i + 1, 2 * j, im);
        put(2 * i + 1, 2 * j + 1, re);
    }
}
// forward images (CTAs 0 .. K-1) and reverse images (the rest) in ONE launch: the two sets are independent
__global__ void __launch_bounds__(512) tc_prep_all_kernel(const float* __restrict__ w, int K, int S, DepthPack dp,
                                                         unsigned char* __restrict__ bimg, unsigned char* __restrict__ rimg,
                                                         int per_block) {
    __shared__ double vr[32][33], vi[32][33], uu[5][4];
    if ((int)blockIdx.x < K) tc_prep_body(w, K, dp, bimg, blockIdx.x, vr, vi, uu);
    else tc_prep_rev_body(w, K, S, dp, rimg, per_block, (int)blockIdx.x - K, vr, vi, uu);
}

// ---------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------
// whole D row (64 f32 columns) -> registers
__device__ __forceinline__ void tc_load_state(uint32_t taddr, uint32_t (&r)[64]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t t[16];
        tc::tmem_ld16(taddr + 16u * c, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) r[16 * c + i] = t[i];
    }
    tc::tmem_wait_ld();
}

// one amplitude (re, im) -> packed f16 hi and lo parts: hi = the 11 leading significant bits (mask), lo = x - hi (exact in
// f32, one packed FFMA2), both converted with round-to-nearest
__device__ __forceinline__ void tc_split(u64 x, uint32_t& ahi, uint32_t& alo) {
    float xr, xi;
    unpack2(x, xr, xi);
    const float hr = __uint_as_float(__float_as_uint(xr) & 0xFFFFE000u);
    const float hi_ = __uint_as_float(__float_as_uint(xi) & 0xFFFFE000u);
    float lr, li;
    unpack2(fma2<6>(1.f, pack2(hr, hi_), x), lr, li);      // x - h
    ahi = tc::cvt_f16x2(hr, hi_);
    alo = tc::cvt_f16x2(lr, li);
}

// registers (scaled state) -> f16 hi | lo operand rows at taddr (+0: hi, +32: lo)
__device__ __forceinline__ void tc_store_operand(uint32_t taddr, const uint32_t (&r)[64]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t ahi[8], alo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) tc_split(tc_pair(r, 8 * c + i), ahi[i], alo[i]);
        tc::tmem_st8(taddr + 8u * c, ahi);
        tc::tmem_st8(taddr + 32u + 8u * c, alo);
    }
}

// r <- scale-folded phases (.) r   (CONJ: conjugate phases).  v p = v_re (p_re, p_im) + v_im (-p_im, p_re): the state's
// two components are the broadcast scalars, the packed phase is the vector operand
template <bool CONJ>
__device__ __forceinline__ void tc_apply_phases(uint32_t (&r)[64], const u64 (&p)[16]) {
#pragma unroll
    for (int z = 0; z < 32; ++z) {
        const float vr = __uint_as_float(r[2 * z]), vi = __uint_as_float(r[2 * z + 1]);
        const int t = z < 16 ? z : 31 - z;
        const bool cj = (z >= 16) != CONJ;     // p[31 - z] = conj(p[z])
        const u64 nv = cj ? tc_cmul_conj(p[t], vr, vi) : tc_cmul(p[t], vr, vi);
        float xr, xi;
        unpack2(nv, xr, xi);
        r[2 * z] = __float_as_uint(xr);
        r[2 * z + 1] = __float_as_uint(xi);
    }
}

// Im <lam| T |psi> for the 15 strings of one cut type: mv[3q + {0,1,2}] = {X,Y,Z}_q moments
template <bool HAD>
__device__ __forceinline__ void tc_moments(const uint32_t (&ps)[64], const uint32_t (&lm)[64], float (&mv)[16]) {
    // 15 independent accumulation chains (one per string), the strings innermost: a chain's FFMA2s are 15
    // instructions apart, so none waits on its predecessor
    u64 acc[15];
#pragma unroll
    for (int T = 0; T < 15; ++T) acc[T] = 0ull;
    static_for<32>([&](auto Zc) {
        constexpr int zp = decltype(Zc)::value;
        const u64 l = tc_pair(lm, zp);
        static_for<15>([&](auto Tc) {
            constexpr int T = decltype(Tc)::value;
            constexpr TcString st = HAD ? kTcStrHad[T] : kTcStrComp[T];
            constexpr bool neg = tc_parity((zp ^ st.mx) & st.mz);
            // Im(i^k s c), c = conj(lam_zp) psi_{zp^mx}:  k=0: s Im c | 1: s Re c | 2: -s Im c | 3: -s Re c
            constexpr bool want_re = (st.k & 1) != 0;
            constexpr bool minus = neg != (st.k >= 2);
            constexpr int PAT = want_re ? (minus ? 6 : 0) : (minus ? 2 : 3);
            acc[T] = fma2_vp<PAT>(l, tc_pair(ps, zp ^ st.mx), acc[T]);
        });
    });
#pragma unroll
    for (int T = 0; T < 15; ++T) mv[T] = lo2(acc[T]) + hi2(acc[T]);
    mv[15] = 0.f;
}

// encoding-angle gradients: sum_z (1 - 2 z_q) Im(conj(mu_z) phi_z), q = 0..4
__device__ __forceinline__ void tc_xgrad(const uint32_t (&ps)[64], const uint32_t (&lm)[64], float (&gq)[5]) {
    u64 acc[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) acc[q] = 0ull;
    static_for<32>([&](auto Zc) {
        constexpr int z = decltype(Zc)::value;
        const u64 l = tc_pair(lm, z), pp = tc_pair(ps, z);
        static_for<5>([&](auto Qc) {
            constexpr int q = decltype(Qc)::value;
            if constexpr (((z >> q) & 1) != 0) acc[q] = fma2_vp<2>(l, pp, acc[q]);
            else acc[q] = fma2_vp<3>(l, pp, acc[q]);
        });
    });
#pragma unroll
    for (int q = 0; q < 5; ++q) gq[q] = lo2(acc[q]) + hi2(acc[q]);
}

// dL/dout of one sample: the fused MSE residual (g = gscale (out + bias - y)) or the upstream gradient
__device__ __forceinline__ float tc_sample_g(const HeaParams<float>& p, float e, int64_t b, bool valid, float& resid) {
    resid = 0.f;
    if (!valid) return 0.f;
    if (p.target) {
        resid = e + (p.bias ? __ldg(p.bias) : 0.f) - __ldg(p.target + b);
        return p.gscale * resid;
    }
    return p.gout ? __ldg(p.gout + b) : 0.f;
}

// the same five sums from ONE packed product per amplitude and a pairwise tree: with t_z = Im(conj(mu_z) phi_z),
//   S_q = sum_z (1 - 2 z_q) t_z = T - 2 O_q,   T = sum_z t_z,   O_q = sum over z with z_q = 1,
// level q pairs the partial sums that differ in bit q: 32 products + 57 packed adds instead of 160 FFMA2
__device__ __forceinline__ void tc_xgrad_tree(const uint32_t (&ps)[64], const uint32_t (&lm)[64], float (&gq)[5]) {
    u64 u[32];
#pragma unroll
    for (int z = 0; z < 32; ++z) u[z] = fma2_vp<3>(tc_pair(lm, z), tc_pair(ps, z), 0ull);    // (mu_r phi_i, -mu_i phi_r)
    u64 odd[5];
    static_for<5>([&](auto Qc) {
        constexpr int q = decltype(Qc)::value;
        constexpr int n = 32 >> q;                 // live partial sums at this level
        odd[q] = u[1];
#pragma unroll
        for (int j = 1; j < n / 2; ++j) odd[q] = add2(odd[q], u[2 * j + 1]);
#pragma unroll
        for (int j = 0; j < n / 2; ++j) u[j] = add2(u[2 * j], u[2 * j + 1]);
    });
    const float T = lo2(u[0]) + hi2(u[0]);
#pragma unroll
    for (int q = 0; q < 5; ++q) gq[q] = fmaf(-2.f, lo2(odd[q]) + hi2(odd[q]), T);
}

// bounded mbarrier wait without busy work: try_wait suspends in hardware for up to ~20 us per probe
__device__ __forceinline__ bool tc_wait(uint32_t bar, uint32_t parity, int* err) {
    for (int it = 0; it < 100000; ++it) {
        uint32_t ok;
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(bar), "r"(parity), "r"(20000u)
            : "memory");
        if (ok) return true;
        if ((it & 63) == 63 && *reinterpret_cast<volatile int*>(err) != 0) break;
    }
    atomicExch(err, 1);
    return false;
}

// ---------------------------------------------------------------------------------------------------------
// the kernel.  ENC as in hea_reg.cuh (0: x given, 1: fused encoding, 2: + frequency-layer gradients)
// images: [K forward block images | S reverse sublayer images in sweep order]
// ---------------------------------------------------------------------------------------------------------
// SPLIT: the training step as TWO kernels — the forward-only kernel (4 tiles, 16 compute warps) leaves every sample's
// final state row in `state` (256 B per sample), the gradient kernel (SPLIT = true) starts its reverse sweep from
// there instead of running the forward sweep itself on its 2 tiles / 8 warps.
template <bool GRAD, bool NEED_GX, int ENC, bool DBG, bool SPLIT = false>
__global__ void __launch_bounds__(TcGeom<GRAD>::THREADS, 1)
hea_tc_kernel(const HeaParams<float> p, const unsigned char* __restrict__ images, float* dbg, int* err, float* state,
              int flags, unsigned* gmax = nullptr) {
    using G = TcGeom<GRAD>;
    static_assert(GRAD || !SPLIT, "SPLIT selects the reverse-only gradient kernel");
    constexpr int NQ = 5, NT = G::NT, NS = G::NS;
    constexpr bool FREQ_GRAD = GRAD && ENC == 2;
    constexpr bool WANT_GX = NEED_GX || FREQ_GRAD;
    static_assert(!(NEED_GX && ENC != 0), "grad_x is only materialised when x is");
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ __align__(8) uint64_t bar_full[NT][NS], bar_a[NT], bar_d[NT];
    __shared__ uint32_t tmem_base_s;

    const int lane = threadIdx.x & 31, warp = tc::warp_uniform(threadIdx.x >> 5);
    const int64_t ntiles = (p.B + 127) / 128;
    // Tile slot (CTA, t) takes tiles t * grid + CTA, + grid * NT, ...: a partial last round spreads over the CTAs one tile
    // each (a lone tile has the SM to itself and runs faster than one of a pair), and a slot simply stops after its last
    // live tile instead of idling through dead ones.
    auto live_rounds = [&](int t) -> int64_t {
        const int64_t first = (int64_t)t * gridDim.x + blockIdx.x, step = (int64_t)gridDim.x * NT;
        return first < ntiles ? (ntiles - first + step - 1) / step : 0;
    };
    const int nsteps = (GRAD && SPLIT ? 0 : p.K) + (GRAD ? p.S : 0);

    if (threadIdx.x == 0) {
        for (int t = 0; t < NT; ++t) {
            for (int i = 0; i < NS; ++i) tc::mbar_init(tc::smem_u32(&bar_full[t][i]), 1);
            tc::mbar_init(tc::smem_u32(&bar_a[t]), 4);
            tc::mbar_init(tc::smem_u32(&bar_d[t]), 1);
        }
        tc::mbar_fence_init();
    }
    if (warp == G::COMPUTE_WARPS) tc::tmem_alloc512(tc::smem_u32(&tmem_base_s));
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = (uint32_t)tc::warp_uniform((int)tmem_base_s);

    if (warp >= G::COMPUTE_WARPS) {
        // =================================================== MMA warps: one elected thread per tile (warp-uniform warp
        // index + elect.sync: the descriptors and TMEM addresses stay in uniform registers)
        if constexpr (GRAD) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(G::REGS_MMA));   // hand registers to the compute warps
        const int t = warp - G::COMPUTE_WARPS;
        if (t < NT && tc::elect_one()) {
            const uint32_t mD = tmem_base + (uint32_t)t * G::TILE_COLS;
            const uint32_t mA = mD + (GRAD ? 128u : 64u);
            const uint32_t ring = tc::smem_u32(tc_smem + (size_t)t * NS * kTcImgBytes);
            const uint32_t bar_a_t = tc::smem_u32(&bar_a[t]), bar_d_t = tc::smem_u32(&bar_d[t]);
            constexpr uint32_t idesc = tc::idesc_f16(128, 64);
            const int64_t total = live_rounds(t) * nsteps;
            auto fetch = [&](int64_t g) {
                const int stage = (int)(g % NS);
                const uint32_t fb = tc::smem_u32(&bar_full[t][stage]);
                tc::mbar_expect_tx(fb, kTcImgBytes);
                tc::bulk_g2s(ring + (uint32_t)stage * kTcImgBytes, images + (size_t)((GRAD && SPLIT ? p.K : 0) + g % nsteps) * kTcImgBytes, kTcImgBytes, fb);
            };
            for (int64_t g = 0; g < NS - 1 && g < total; ++g) fetch(g);
            bool dead = false;
            uint32_t apar = 0;
            auto gemm = [&](uint32_t d, uint32_t a, uint32_t sb) {      // D = A_hi B_hi + A_hi B_lo + A_lo B_hi
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    tc::mma_f16_ts(d, a + 8u * j, tc::smem_desc_kmajor(sb + 256u * j, 128u, 1024u), idesc, j > 0);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    tc::mma_f16_ts(d, a + 8u * j, tc::smem_desc_kmajor(sb + 8192u + 256u * j, 128u, 1024u), idesc, 1u);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    tc::mma_f16_ts(d, a + 32u + 8u * j, tc::smem_desc_kmajor(sb + 256u * j, 128u, 1024u), idesc, 1u);
            };
            for (int64_t g = 0; g < total; ++g) {
                const int stage = (int)(g % NS);
                const bool rev = GRAD && (SPLIT || (int)(g % nsteps) >= p.K);
                const uint32_t sb = ring + (uint32_t)stage * kTcImgBytes;
                if (!dead && !tc_wait(bar_a_t, apar, err)) dead = true;
                apar ^= 1u;
                tc::tc_fence_after();
                if (!dead && !tc_wait(tc::smem_u32(&bar_full[t][stage]), (uint32_t)((g / NS) & 1), err)) dead = true;
                if (!dead) {
                    gemm(mD, mA, sb);
                    if (rev) gemm(mD + 64u, mA + 64u, sb);
                }
                tc::mma_commit(bar_d_t);
                // the stage of step g-1 is free: its MMAs completed before a_ready(g) could be signalled
                if (g + NS - 1 < total) fetch(g + NS - 1);
            }
        }
        __syncwarp();
    } else {
        // =================================================== compute warps
        if constexpr (GRAD) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(G::REGS_COMPUTE));
        const int t = warp >> 2, quarter = warp & 3;
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        const uint32_t tDp = tmem_base + lane_sel + (uint32_t)t * G::TILE_COLS;     // D psi
        const uint32_t tDl = tDp + 64u;                                              // D lam      (GRAD)
        const uint32_t tAp = tDp + (GRAD ? 128u : 64u);                              // A psi: hi 32 | lo 32
        const uint32_t tAl = tAp + 64u;                                              // A lam      (GRAD)
        const uint32_t bar_a_t = tc::smem_u32(&bar_a[t]), bar_d_t = tc::smem_u32(&bar_d[t]);
        uint32_t dpar = 0;
        bool dead = false;
        auto wait_d = [&]() {
            if (!dead && !tc_wait(bar_d_t, dpar, err)) dead = true;
            dpar ^= 1u;
            tc::tc_fence_after();
        };
        auto signal_a = [&]() {      // "operand written" / "accumulator drained": the MMA warp may issue the next step
            tc::tmem_wait_st();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(bar_a_t);
        };
        float hmax = 1.f;
        if constexpr (GRAD) {
            float m = 0.f;
            for (int z = 0; z < 32; ++z) m = fmaxf(m, fabsf(__ldg(p.hdiag + z)));
            hmax = m > 0.f ? m : 1.f;
        }
        const int64_t gwarp = (int64_t)blockIdx.x * G::COMPUTE_WARPS + warp;
        float* mrow = GRAD ? p.mpart + gwarp * p.rowlen : nullptr;
        float* frow = GRAD ? mrow + (int64_t)p.S * 16 : nullptr;
        float* srow = GRAD ? frow + (int64_t)p.K * 16 : nullptr;

        int qm0[NQ], qm1[NQ];        // q % in of the two input sources
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            qm0[q] = ENC != 0 ? q % p.in0 : 0;
            qm1[q] = ENC != 0 ? q % p.in1 : 0;
        }

        const int64_t rounds = live_rounds(t);
        for (int64_t round = 0; round < rounds; ++round) {
            const int64_t tile = round * gridDim.x * NT + (int64_t)t * gridDim.x + blockIdx.x;
            const int64_t b = tile * 128 + quarter * 32 + lane;
            const bool valid = b < p.B;
            const int64_t bc = valid ? b : p.B - 1;
            const float* xrow = ENC == 0 ? p.x + bc * p.ldx : nullptr;
            const float* u0row = ENC != 0 && p.u0 ? p.u0 + bc * p.ldu0 : nullptr;
            const float* u1row = ENC != 0 ? p.u1 + bc * p.ldu1 : nullptr;
            // The angles of a block are loaded ONE BLOCK AHEAD as raw inputs and turned into angles when the block starts
            // (load_angles_late): an in-order warp stalls at the first use of a load.  The input column of angle c is
            // uidx[c] = local column % in (prep kernel, qon_capi.cu), recomputed here from the block index so that the
            // gather does not hang on a load of the index table.
            auto load_inputs = [&](int k, float(&un)[NQ]) {
                if constexpr (ENC == 0) {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) un[q] = __ldg(xrow + (int64_t)k * NQ + q);
                } else {
                    const bool s0 = k < p.K0;
                    const float* ur = s0 ? u0row : u1row;
                    const int in = s0 ? p.in0 : p.in1;
                    const int base = ((s0 ? k : k - p.K0) * NQ) % in;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        int idx = base + (s0 ? qm0[q] : qm1[q]);
                        if (idx >= in) idx -= in;
                        un[q] = __ldg(ur + idx);
                    }
                }
            };
            auto angles_from = [&](int k, const float(&un)[NQ], float(&th)[NQ]) {
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    if constexpr (ENC == 0) th[q] = un[q];
                    else th[q] = fmaf(un[q], __ldg(p.fw + k * NQ + q), p.fb ? __ldg(p.fb + k * NQ + q) : 0.f);
                }
            };
            auto load_angles = [&](int k, float(&th)[NQ]) {      // both steps at once (the reverse sweep of this kernel)
                float un[NQ];
                load_inputs(k, un);
                angles_from(k, un, th);
            };
            const bool dump = DBG && dbg && blockIdx.x == 0 && t == 0 && round == 0;
            const int drow = quarter * 32 + lane;

            // ---------------------------------------------------------------- forward sweep
            float th[NQ];
            if constexpr (!SPLIT) {
            load_angles(0, th);
            for (int k = 0; k < p.K; ++k) {
                float un[NQ];
                load_inputs(k + 1 < p.K ? k + 1 : k, un);
                u64 ph[16];
                tc_phase_table(th, 1.f, ph);
                if (k > 0) wait_d();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[16];
                    if (k > 0) {
                        tc::tmem_ld16(tDp + 16u * c, r);
                        tc::tmem_wait_ld();
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            r[2 * i] = __float_as_uint(kTcSA * 0.17677669529663688110f);
                            r[2 * i + 1] = 0u;
                        }
                    }
                    if (DBG && dump && k > 0)
#pragma unroll
                        for (int i = 0; i < 16; ++i) dbg[((size_t)(k - 1) * 128 + drow) * 128 + 16 * c + i] = __uint_as_float(r[i]);
                    uint32_t ahi[8], alo[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int z = 8 * c + i;
                        const float vr = __uint_as_float(r[2 * i]), vi = __uint_as_float(r[2 * i + 1]);
                        const u64 nv = z < 16 ? tc_cmul(ph[z], vr, vi) : tc_cmul_conj(ph[31 - z], vr, vi);
                        tc_split(nv, ahi[i], alo[i]);
                    }
                    tc::tmem_st8(tAp + 8u * c, ahi);
                    tc::tmem_st8(tAp + 32u + 8u * c, alo);
                }
                signal_a();
                angles_from(k + 1 < p.K ? k + 1 : k, un, th);
            }
            wait_d();
            }

            if constexpr (!GRAD) {
                // ------------------------------------------------------------ expectation value only
                float e = 0.f, nrm = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[16];
                    tc::tmem_ld16(tDp + 16u * c, r);
                    tc::tmem_wait_ld();
                    if (DBG && dump)
#pragma unroll
                        for (int i = 0; i < 16; ++i) dbg[((size_t)(p.K - 1) * 128 + drow) * 128 + 16 * c + i] = __uint_as_float(r[i]);
                    if (state && valid) {      // final state row for the split gradient kernel (it renormalises)
                        uint4* dst = reinterpret_cast<uint4*>(state + b * 64 + 16 * c);
#pragma unroll
                        for (int v = 0; v < 4; ++v) dst[v] = make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float re = __uint_as_float(r[2 * i]), im = __uint_as_float(r[2 * i + 1]);
                        const float pz = fmaf(re, re, im * im);
                        e = fmaf(__ldg(p.hdiag + 8 * c + i), pz, e);
                        nrm += pz;
                    }
                }
                // unit norm is exact for the true state; dividing by the computed norm removes the coherent shrink of
                // the truncating tensor-core accumulation (-4.8e-7 +- 0.7e-7 per GEMM, measured) and the operand scale
                float res = e / nrm;
                if (__ldcg(err) != 0) res = __int_as_float(0x7fc00000);
                if (valid && p.out) p.out[b] = res;
                if (gmax) {
                    // training step with GEMM-form weight gradients (hea_tc3.cuh): max |dL/dout| over the batch, as
                    // ordered bits (non-negative floats; a NaN sorts above every number and poisons the step)
                    float resid;
                    const float g = tc_sample_g(p, res, b, valid, resid);
                    const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(fabsf(g)));
                    if (lane == 0) atomicMax(gmax, m);
                }
            } else {
                // ------------------------------------------------------------ expectation, lam = g H psi
                uint32_t ps[64], lm[64];
                if constexpr (SPLIT) {
                    const uint4* src = reinterpret_cast<const uint4*>(state + bc * 64);
#pragma unroll
                    for (int v = 0; v < 16; ++v) {
                        const uint4 q4 = __ldcs(src + v);
                        ps[4 * v] = q4.x; ps[4 * v + 1] = q4.y; ps[4 * v + 2] = q4.z; ps[4 * v + 3] = q4.w;
                    }
                } else {
                    tc_load_state(tDp, ps);
                }
                if (DBG && dump)
#pragma unroll
                    for (int i = 0; i < 64; ++i) dbg[((size_t)(p.K - 1) * 128 + drow) * 128 + i] = __uint_as_float(ps[i]);
                float e = 0.f, nrm = 0.f;
#pragma unroll
                for (int z = 0; z < 32; ++z) {
                    const float re = __uint_as_float(ps[2 * z]), im = __uint_as_float(ps[2 * z + 1]);
                    const float pz = fmaf(re, re, im * im);
                    e = fmaf(__ldg(p.hdiag + z), pz, e);
                    nrm += pz;
                }
                e = e / nrm;
                if (__ldcg(err) != 0) e = __int_as_float(0x7fc00000);
                if (valid && p.out) p.out[b] = e;
                float g = 0.f;
                if (p.target) {   // fused MSE: g = dL/dout for L = gscale/2 * sum (out + bias - y)^2
                    float resid = 0.f;
                    if (valid) {
                        resid = e + (p.bias ? __ldg(p.bias) : 0.f) - __ldg(p.target + b);
                        g = p.gscale * resid;
                        if (p.gbuf) p.gbuf[b] = g;
                    }
                    float sg = g, sq = resid * resid;
#pragma unroll
                    for (int m = 16; m >= 1; m >>= 1) { sg += shfl_xor_(sg, m); sq += shfl_xor_(sq, m); }
                    if (lane == 0) { atomicAdd(srow, sg); atomicAdd(srow + 1, sq); }
                } else if (valid) {
                    g = __ldg(p.gout + b);
                }
                // psi to norm sA; lam_hat = (h / hmax) psi (norm <= sA); the scalar g * hmax / sA^2 goes onto the moments
                const float rn = kTcSA * rsqrtf(nrm);
                const float glam = g * hmax * (1.f / (kTcSA * kTcSA));
                const float ihm = 1.f / hmax;
#pragma unroll
                for (int z = 0; z < 32; ++z) {
                    const float re = __uint_as_float(ps[2 * z]) * rn, im = __uint_as_float(ps[2 * z + 1]) * rn;
                    const float hz = __ldg(p.hdiag + z) * ihm;
                    ps[2 * z] = __float_as_uint(re);
                    ps[2 * z + 1] = __float_as_uint(im);
                    lm[2 * z] = __float_as_uint(re * hz);
                    lm[2 * z + 1] = __float_as_uint(im * hz);
                }

                // ------------------------------------------------------------ reverse (adjoint) sweep
                float* gxrow = NEED_GX ? p.gx + (valid ? b : 0) * p.ldgx : nullptr;
                int s = p.S;
                int step = p.K;
                load_angles(p.K - 1, th);
                for (int k = p.K - 1; k >= 0; --k) {
                    // flags bit 0: the two tiles of the CTA start every block together (named barrier over the 256 compute
                    // threads), so the two warps of a scheduler run the same ~28 KB loop body at the same time and share
                    // its instruction fetches
                    if (flags & 1) asm volatile("bar.sync 1, %0;" ::"n"(G::COMPUTE_WARPS * 32) : "memory");
                    float thn[NQ];
                    load_angles(k > 0 ? k - 1 : 0, thn);
                    const int d = __ldg(p.depth + k);
                    // (ps, lm) = the block's output cut.  ONE operand split per block: every cut inside the block is a
                    // GEMM of this operand with a composite matrix (prep), so the GEMM producing cut s-1 runs while the
                    // CUDA cores measure the moments of cut s.
                    tc_store_operand(tAp, ps);
                    tc_store_operand(tAl, lm);
                    signal_a();
                    for (int j = d - 1; j >= 0; --j, ++step) {
                        --s;
                        // Pauli moments of sublayer s on the cut held in registers, scaled by this sample's g
                        float mv[16];
                        // every cut is held in the Hadamard basis except the very first one (the final state, in the
                        // computational basis): one moment routine in the steady-state loop keeps its body inside
                        // the instruction cache (two variants: GPC instruction requests at 95 % of peak, ncu)
                        if (k == p.K - 1 && j == d - 1) tc_moments<false>(ps, lm, mv);
                        else tc_moments<true>(ps, lm, mv);
#pragma unroll
                        for (int i = 0; i < 15; ++i) mv[i] *= glam;
                        const float tot = butterfly_reduce<float, 16>(mv, lane);
                        if ((lane & 1) == 0) atomicAdd(mrow + (int64_t)s * 16 + (lane >> 1), tot);
                        u64 ph[16];
                        if (j == 0) tc_phase_table(th, 1.f, ph);
                        wait_d();
                        tc_load_state(tDp, ps);
                        tc_load_state(tDl, lm);
                        if (j > 0) signal_a();        // accumulators drained: next composite of this block
                        if (DBG && dump) {
#pragma unroll
                            for (int i = 0; i < 64; ++i) {
                                dbg[((size_t)step * 128 + drow) * 128 + i] = __uint_as_float(ps[i]);
                                dbg[((size_t)step * 128 + drow) * 128 + 64 + i] = __uint_as_float(lm[i]);
                            }
                        }
                        if (j == 0) {
                            // Hadamard basis, right after the block's encoding layer
                            if constexpr (WANT_GX) {
                                float gq[5];
                                tc_xgrad(ps, lm, gq);
                                float fv[FREQ_GRAD ? 16 : 1];
                                if constexpr (FREQ_GRAD) {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) fv[i] = 0.f;
                                }
#pragma unroll
                                for (int q = 0; q < NQ; ++q) {
                                    const float gxv = gq[q] * glam;
                                    if constexpr (NEED_GX) {
                                        if (valid) gxrow[(int64_t)k * NQ + q] = gxv;
                                    }
                                    if constexpr (FREQ_GRAD) {
                                        const int col = k * NQ + q;
                                        const float uval = __ldg((k < p.K0 ? u0row : u1row) + __ldg(p.uidx + col));
                                        fv[2 * q] = gxv * uval;
                                        fv[2 * q + 1] = gxv;
                                    }
                                }
                                if constexpr (FREQ_GRAD) {
                                    const float ft = butterfly_reduce<float, 16>(fv, lane);
                                    if ((lane & 1) == 0) atomicAdd(frow + (int64_t)k * 16 + (lane >> 1), ft);
                                }
                            }
                            if (s > 0) {
                                // conjugate phases; the scale restores |psi| = sA (the truncating accumulation shrinks
                                // both states by the same factor per GEMM), applied to lam as well
                                float nr = 0.f;
#pragma unroll
                                for (int z = 0; z < 32; ++z) {
                                    const float re = __uint_as_float(ps[2 * z]), im = __uint_as_float(ps[2 * z + 1]);
                                    nr += fmaf(re, re, im * im);
                                }
                                const float corr = kTcSA * rsqrtf(nr);
#pragma unroll
                                for (int i = 0; i < 16; ++i) ph[i] = mul2<0>(corr, ph[i]);
                                tc_apply_phases<true>(ps, ph);
                                tc_apply_phases<true>(lm, ph);
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < NQ; ++q) th[q] = thn[q];
                }
                // a barrier wait that timed out leaves garbage: poison this warp's partial sums so that the loss and
                // every gradient of the step read NaN instead of a plausible number
                if (lane == 0 && __ldcg(err) != 0) {
                    atomicAdd(mrow, __int_as_float(0x7fc00000));
                    atomicAdd(srow + 1, __int_as_float(0x7fc00000));
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == G::COMPUTE_WARPS) tc::tmem_dealloc512(tmem_base);
}

}  // namespace qon
