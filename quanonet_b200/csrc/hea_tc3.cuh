// Tensor-core tier, reverse sweep with GEMM-form weight gradients (n = 5, fp32 results).
//
// The trainable sublayers of a block are sample-independent (reference: core/quantum_circuits_tq.py:89-101), so every
// Pauli moment the finalize kernel needs inside block k follows from ONE batch-summed outer product taken at the
// block's output cut:
//     Y_k = sum_b g_b |psi_b><lam_b|            a (64 x B) . (B x 64) real product = tcgen05.mma with the SAMPLES as
//                                               the reduction dimension (both operands MN-major in shared memory)
//     cut after the rotations of sublayer s:    Y <- T Y T^+   (T: the sample-independent gates behind the cut)
//     moment of P_q there:                      sum_b g_b Im <lam_b| P_q |psi_b>  =  Im tr(P_q Y)
// (tc_moment_kernel below; emulated against the fp64 oracle in tests/harness/tc_emulate_outer.py).  The per-sample work
// of the sweep shrinks from 15 Pauli-string moments per SUBLAYER on the CUDA cores (hea_tc2.cuh: 480 FFMA2) to one
// operand split per BLOCK, which feeds both the block's un-apply GEMMs (A operand in TMEM) and the outer product
// (the same f16 hi/lo values, stored once more to shared memory; as (sample, amplitude) rows they are the K-major A
// operand of the first and the MN-major operands of the second).
//
// Batch accumulation: a 128-sample tile's 64 x 64 real accumulator is folded in registers to the 32 x 32 complex Y
// (2,048 floats; the lam operand is stored component-planar so that a thread of the 16x256b TMEM load owns the four
// real products of one entry) and added to the PRIVATE accumulator of its (CTA, tile slot) with plain vector
// loads / stores — no atomics (a 64-bit red.add per entry measured 10.6 ms per 1M samples: the LSU retires about one
// atomic lane per clock per SM).  Each slot adds its tiles in a fixed order and tc_moment_kernel sums the slots in a
// fixed order in fp64, so the gradients stay bitwise deterministic.  The f16 operands need a bounded range: g_b is
// divided by E, the power of two above max_b |g_b| that the forward kernel of the step leaves in `gmax` (hea_tc2.cuh).
//
// Roles per CTA as in hea_tc2.cuh (2 tiles of 128 samples, 4 compute warps + 1 MMA warp each).  TMEM per tile, 256
// columns: D psi | D lam | A psi (hi, lo) | A lam (hi, lo); the outer product's accumulator (M = 64: lanes 0..15 of each
// subpartition) ALIASES the A psi columns — the MMA warp waits for the psi un-apply GEMM (bar_x) before issuing it, the
// compute warps drain it (bar_g) before they write the next block's operand.
//
// One block of the sweep, per tile (every hand-off an mbarrier; every wait bounded, a time-out poisons the outputs):
//   compute warps                                              MMA warp (one elected thread, uniform registers)
//   split psi -> TMEM A psi + smem            -- bar_a  -->    psi un-apply GEMM (12 MMAs)              -- commit bar_x
//   split lam -> TMEM A lam + smem (planar)   -- bar_a2 -->    lam un-apply GEMM (12 MMAs)              -- commit bar_d
//   load this slot's running sums (global)                     wait bar_x; outer product (24 MMAs)      -- commit bar_g
//   wait bar_x: tcgen05.ld psi, norm, sin/cos + phase table,   fetch the next block's image (cp.async.bulk)
//               conjugate phases on psi        (under the lam GEMM)
//   wait bar_d: tcgen05.ld lam, conjugate phases, encoding-angle gradients (tree), frequency-layer partial sums
//   wait bar_g: tcgen05.ld.16x256b the outer product, fold re/im, add the running sums, store
//   inputs of block k-1 (gathered a block ahead) -> angles
#pragma once
#include "hea_tc2.cuh"

namespace qon {

struct TcRev {
    static constexpr int NT = 2, NS = QON_TC_REV_STAGES;
    static constexpr int COMPUTE_WARPS = 4 * NT, WARPS = COMPUTE_WARPS + 4, THREADS = WARPS * 32;
    static constexpr int TILE_COLS = 256;
    static constexpr int OPER_BYTES = 65536;                 // per tile: psi hi | psi lo | lam hi | lam lo, 16 KB each
    static constexpr int TILE_SMEM = OPER_BYTES + NS * kTcImgBytes;
    static constexpr int SMEM = NT * TILE_SMEM;
    static constexpr int REGS_COMPUTE = QON_TC_REV_REGS, REGS_MMA = (168 * 384 - 256 * QON_TC_REV_REGS) / 128;      // 232 / 40
};
constexpr int kTcOuterMaxBlocks = 256;
constexpr int kTcAccLen = 2048;                              // floats per (slot, block): the 32 x 32 complex Y in fragment order

// power of two above max |g| (bits from the forward kernel); NaN / inf poison the step
__device__ __forceinline__ float tc_gscale_pow2(unsigned gb) {
    const float gm = __uint_as_float(gb);
    if (!(gm < __int_as_float(0x7f800000))) return __int_as_float(0x7fc00000);
    return __uint_as_float((gb & 0x7f800000u) + 0x00800000u);
}

// registers (scaled state) -> f16 hi | lo: A operand rows in TMEM (+0: hi, +32: lo) and the sample's row of the
// shared-memory operand, offset(sample r, row m) = (r >> 3) * 1024 + (m >> 3) * 128 + (r & 7) * 16 + (m & 7) * 2.
// Row order m of the shared-memory copy: interleaved m = 2 z + c (psi), or PLANAR m = 16 (z >> 3) + 8 c + (z & 7)
// (lam: the real and imaginary rows of an amplitude end up 8 TMEM lanes apart, i.e. in one thread of the 16x256b load).
template <bool PLANAR>
__device__ __forceinline__ void tc_store_operand2(uint32_t taddr, unsigned char* sm_hi, unsigned char* sm_lo,
                                                  const uint32_t (&r)[64]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t ahi[8], alo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) tc_split(tc_pair(r, 8 * c + i), ahi[i], alo[i]);
        tc::tmem_st8(taddr + 8u * c, ahi);
        tc::tmem_st8(taddr + 32u + 8u * c, alo);
        if constexpr (PLANAR) {
            uint32_t re[4], im[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { re[i] = __byte_perm(ahi[2 * i], ahi[2 * i + 1], 0x5410); im[i] = __byte_perm(ahi[2 * i], ahi[2 * i + 1], 0x7632); }
            *reinterpret_cast<uint4*>(sm_hi + (2 * c) * 128) = make_uint4(re[0], re[1], re[2], re[3]);
            *reinterpret_cast<uint4*>(sm_hi + (2 * c + 1) * 128) = make_uint4(im[0], im[1], im[2], im[3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) { re[i] = __byte_perm(alo[2 * i], alo[2 * i + 1], 0x5410); im[i] = __byte_perm(alo[2 * i], alo[2 * i + 1], 0x7632); }
            *reinterpret_cast<uint4*>(sm_lo + (2 * c) * 128) = make_uint4(re[0], re[1], re[2], re[3]);
            *reinterpret_cast<uint4*>(sm_lo + (2 * c + 1) * 128) = make_uint4(im[0], im[1], im[2], im[3]);
        } else {
            *reinterpret_cast<uint4*>(sm_hi + (2 * c) * 128) = make_uint4(ahi[0], ahi[1], ahi[2], ahi[3]);
            *reinterpret_cast<uint4*>(sm_hi + (2 * c + 1) * 128) = make_uint4(ahi[4], ahi[5], ahi[6], ahi[7]);
            *reinterpret_cast<uint4*>(sm_lo + (2 * c) * 128) = make_uint4(alo[0], alo[1], alo[2], alo[3]);
            *reinterpret_cast<uint4*>(sm_lo + (2 * c + 1) * 128) = make_uint4(alo[4], alo[5], alo[6], alo[7]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// the reverse-sweep kernel.  images: K whole-block un-apply images in sweep order (tc_prep_all_kernel, per_block);
// state: the final state rows left by the forward-only kernel; acc: [2 gridDim.x slots][K][2048] floats, slot =
// t * gridDim.x + CTA (written, not accumulated, in a slot's first round: no clearing needed; the slots that have a tile
// are the first min(tiles, 2 gridDim.x))
// ---------------------------------------------------------------------------------------------------------
template <bool NEED_GX, int ENC>
__global__ void __launch_bounds__(TcRev::THREADS, 1)
hea_tc_rev_kernel(const HeaParams<float> p, const unsigned char* __restrict__ images, int* err,
                  const float* __restrict__ state, const unsigned* __restrict__ gmax, float* __restrict__ acc,
                  float* dbg, int flags) {
    using G = TcRev;
    constexpr int NQ = 5, NT = G::NT, NS = G::NS;
    constexpr bool FREQ_GRAD = ENC == 2;
    constexpr bool WANT_GX = NEED_GX || FREQ_GRAD;
    static_assert(!(NEED_GX && ENC != 0), "grad_x is only materialised when x is");
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    __shared__ __align__(8) uint64_t bar_full[NT][NS], bar_a[NT], bar_a2[NT], bar_x[NT], bar_d[NT], bar_g[NT];
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

    if (threadIdx.x == 0) {
        for (int t = 0; t < NT; ++t) {
            for (int i = 0; i < NS; ++i) tc::mbar_init(tc::smem_u32(&bar_full[t][i]), 1);
            tc::mbar_init(tc::smem_u32(&bar_a[t]), 4);
            tc::mbar_init(tc::smem_u32(&bar_a2[t]), 4);
            tc::mbar_init(tc::smem_u32(&bar_x[t]), 1);
            tc::mbar_init(tc::smem_u32(&bar_d[t]), 1);
            tc::mbar_init(tc::smem_u32(&bar_g[t]), 1);
        }
        tc::mbar_fence_init();
    }
    if (warp == G::COMPUTE_WARPS) tc::tmem_alloc512(tc::smem_u32(&tmem_base_s));
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = (uint32_t)tc::warp_uniform((int)tmem_base_s);

    if (warp >= G::COMPUTE_WARPS) {
        // =================================================== MMA warps: one elected thread per tile
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(G::REGS_MMA));
        const int t = warp - G::COMPUTE_WARPS;
        if (t < NT && tc::elect_one()) {
            const uint32_t mD = tmem_base + (uint32_t)t * G::TILE_COLS;
            const uint32_t mA = mD + 128u;
            const uint32_t mG = mA;                                        // outer-product accumulator: aliases A psi
            const uint32_t oper = tc::smem_u32(tc_smem + (size_t)t * G::TILE_SMEM);
            const uint32_t ring = oper + G::OPER_BYTES;
            const uint32_t bar_a_t = tc::smem_u32(&bar_a[t]), bar_a2_t = tc::smem_u32(&bar_a2[t]), bar_x_t = tc::smem_u32(&bar_x[t]);
            const uint32_t bar_d_t = tc::smem_u32(&bar_d[t]), bar_g_t = tc::smem_u32(&bar_g[t]);
            constexpr uint32_t idesc = tc::idesc_f16(128, 64);
            constexpr uint32_t idesc_o = tc::idesc_f16(64, 64) | (1u << 15) | (1u << 16);      // A, B MN-major
            // operand rows: groups of 8 samples 1024 B apart (the K direction of the outer product), groups of 8
            // components 128 B apart (its M / N direction)
            constexpr uint32_t o_lbo = 1024u, o_sbo = 128u;      // (confirmed against the exact accumulator of a tile: tests/harness/tc_check_outer.py)
            const int64_t total = live_rounds(t) * p.K;
            auto fetch = [&](int64_t g) {
                const int stage = (int)(g % NS);
                const uint32_t fb = tc::smem_u32(&bar_full[t][stage]);
                tc::mbar_expect_tx(fb, kTcImgBytes);
                tc::bulk_g2s(ring + (uint32_t)stage * kTcImgBytes, images + (size_t)(g % p.K) * kTcImgBytes, kTcImgBytes, fb);
            };
            for (int64_t g = 0; g < (NS > 1 ? NS - 1 : 1) && g < total; ++g) fetch(g);
            bool dead = false;
            uint32_t apar = 0, xpar = 0;
            [[maybe_unused]] uint32_t dpar_m = 0;
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
            // D_G[lam component][psi component] = sum over the tile's 128 samples, 16 per instruction:
            // lam_hi psi_hi + lam_hi psi_lo + lam_lo psi_hi
            auto outer = [&]() {
                const uint64_t dlh = tc::smem_desc_kmajor(oper + 32768u, o_lbo, o_sbo), dll = tc::smem_desc_kmajor(oper + 49152u, o_lbo, o_sbo);
                const uint64_t dph = tc::smem_desc_kmajor(oper, o_lbo, o_sbo), dpl = tc::smem_desc_kmajor(oper + 16384u, o_lbo, o_sbo);
#pragma unroll 1
                for (int j = 0; j < 8; ++j) {
                    const uint64_t o = (uint64_t)(128 * j);          // 16 samples = 2,048 bytes further (address field: >> 4)
                    tc::mma_f16_ss(mG, dlh + o, dph + o, idesc_o, j > 0);
                    tc::mma_f16_ss(mG, dlh + o, dpl + o, idesc_o, 1u);
                    tc::mma_f16_ss(mG, dll + o, dph + o, idesc_o, 1u);
                }
            };
            for (int64_t g = 0; g < total; ++g) {
                const int stage = (int)(g % NS);
                const uint32_t sb = ring + (uint32_t)stage * kTcImgBytes;
                if (!dead && !tc_wait(bar_a_t, apar, err)) dead = true;
                apar ^= 1u;
                tc::tc_fence_after();
                if (!dead && !tc_wait(tc::smem_u32(&bar_full[t][stage]), (uint32_t)((g / NS) & 1), err)) dead = true;
                if (!dead) gemm(mD, mA, sb);                       // psi: runs while the compute warps still split lam
                tc::mma_commit(bar_x_t);
                if (!dead && !tc_wait(bar_a2_t, apar ^ 1u, err)) dead = true;
                tc::tc_fence_after();
                if (!dead) gemm(mD + 64u, mA + 64u, sb);
                tc::mma_commit(bar_d_t);
                // the A psi columns are free once the psi GEMM has completed (the lam GEMM keeps the pipe busy meanwhile)
                if (!dead && !tc_wait(bar_x_t, xpar, err)) dead = true;
                xpar ^= 1u;
                tc::tc_fence_after();
                if (!dead) outer();
                tc::mma_commit(bar_g_t);
                if constexpr (NS > 1) {
                    if (g + NS - 1 < total) fetch(g + NS - 1);
                } else {
                    // one stage: the next image may land once both un-apply GEMMs have read this one (the compute warps
                    // work on their results for far longer than the copy takes)
                    if (!dead && !tc_wait(bar_d_t, dpar_m, err)) dead = true;
                    dpar_m ^= 1u;
                    if (g + 1 < total) fetch(g + 1);
                }
            }
        }
        __syncwarp();
    } else {
        // =================================================== compute warps
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(G::REGS_COMPUTE));
        const int t = warp >> 2, quarter = warp & 3;
        const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
        const uint32_t tDp = tmem_base + lane_sel + (uint32_t)t * G::TILE_COLS;     // D psi
        const uint32_t tDl = tDp + 64u;                                              // D lam
        const uint32_t tAp = tDp + 128u;                                             // A psi: hi 32 | lo 32; D_G
        const uint32_t tAl = tAp + 64u;                                              // A lam
        const int srow_t = quarter * 32 + lane;                                      // sample row inside the tile
        unsigned char* op = tc_smem + (size_t)t * G::TILE_SMEM + (srow_t >> 3) * 1024 + (srow_t & 7) * 16;
        const uint32_t bar_a_t = tc::smem_u32(&bar_a[t]), bar_a2_t = tc::smem_u32(&bar_a2[t]), bar_x_t = tc::smem_u32(&bar_x[t]);
        const uint32_t bar_d_t = tc::smem_u32(&bar_d[t]), bar_g_t = tc::smem_u32(&bar_g[t]);
        uint32_t dpar = 0, gpar = 0, xpar = 0;
        bool dead = false;
        auto wait_on = [&](uint32_t bar, uint32_t& par) {
            if (!dead && !tc_wait(bar, par, err)) dead = true;
            par ^= 1u;
            tc::tc_fence_after();
        };
        auto signal = [&](uint32_t bar) {      // "operand written": the MMA warp may issue the GEMMs that read it
            tc::fence_proxy_async_smem();
            tc::tmem_wait_st();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(bar);
        };
        float hmax = 1.f;
        {
            float m = 0.f;
            for (int z = 0; z < 32; ++z) m = fmaxf(m, fabsf(__ldg(p.hdiag + z)));
            hmax = m > 0.f ? m : 1.f;
        }
        const float E = tc_gscale_pow2(__ldcg(gmax));
        const float invE = 1.f / E;
        const float xscale = E * hmax * (1.f / (kTcSA * kTcSA));     // lam carries g / E and h / hmax
        const int64_t gwarp = (int64_t)blockIdx.x * G::COMPUTE_WARPS + warp;
        float* mrow = p.mpart + gwarp * p.rowlen;
        float* frow = mrow + (int64_t)p.S * 16;
        float* srow = frow + (int64_t)p.K * 16;

        const int step0 = ENC != 0 ? NQ % p.in0 : 0, step1 = ENC != 0 ? NQ % p.in1 : 0;
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
            const bool tile_live = tile * 128 < p.B;
            const int64_t bc = valid ? b : p.B - 1;
            const float* xrow = ENC == 0 ? p.x + bc * p.ldx : nullptr;
            const float* u0row = ENC != 0 && p.u0 ? p.u0 + bc * p.ldu0 : nullptr;
            const float* u1row = ENC != 0 ? p.u1 + bc * p.ldu1 : nullptr;
            // The angles of block k are loaded ONE BLOCK AHEAD as raw inputs (un) and turned into angles at the bottom of
            // the iteration, when the gather has long arrived: an in-order warp stalls at the first use of a load, and with
            // two warps per scheduler nobody covers for it.  The input column of angle c is uidx[c] = local column % in
            // (prep kernel, qon_capi.cu); it is recomputed from the block index here, so the gather does not hang on a
            // load of the index table.
            int cbase = 0, cblock = -2;      // (5 k') mod in of the block loaded last: stepped down instead of divided
            auto load_inputs = [&](int k, float(&un)[NQ]) {
                if constexpr (ENC == 0) {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) un[q] = __ldg(xrow + (int64_t)k * NQ + q);
                } else {
                    const bool s0 = k < p.K0;
                    const float* ur = s0 ? u0row : u1row;
                    const int in = s0 ? p.in0 : p.in1;
                    if (k == cblock - 1 && (k + 1 < p.K0) == s0) {      // one block down inside the same source
                        cbase -= s0 ? step0 : step1;
                        if (cbase < 0) cbase += in;
                    } else if (k != cblock) {
                        cbase = ((s0 ? k : k - p.K0) * NQ) % in;
                    }
                    cblock = k;
                    const int base = cbase;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        int idx = base + (s0 ? qm0[q] : qm1[q]);
                        if (idx >= in) idx -= in;
                        un[q] = __ldg(ur + idx);      // through L1: bypassing it (ld.global.nc.L1::no_allocate) measured 9.1 vs 7.9 ms per step
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

            // ------------------------------------------------------------ expectation, lam = (g / E) (h / hmax) psi
            uint32_t ps[64], lm[64];
            {
                const uint4* src = reinterpret_cast<const uint4*>(state + bc * 64);
#pragma unroll
                for (int v = 0; v < 16; ++v) {
                    const uint4 q4 = __ldcs(src + v);
                    ps[4 * v] = q4.x; ps[4 * v + 1] = q4.y; ps[4 * v + 2] = q4.z; ps[4 * v + 3] = q4.w;
                }
            }
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
            float resid;
            const float g = tc_sample_g(p, e, b, valid, resid);
            if (p.target) {
                if (valid && p.gbuf) p.gbuf[b] = g;
                float sg = g, sq = resid * resid;
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) { sg += shfl_xor_(sg, m); sq += shfl_xor_(sq, m); }
                if (lane == 0) { atomicAdd(srow, sg); atomicAdd(srow + 1, sq); }
            }
            {
                const float rn = kTcSA * rsqrtf(nrm);
                const float gl = g * invE * (1.f / hmax);
#pragma unroll
                for (int z = 0; z < 32; ++z) {
                    const float re = __uint_as_float(ps[2 * z]) * rn, im = __uint_as_float(ps[2 * z + 1]) * rn;
                    const float hz = __ldg(p.hdiag + z) * gl;
                    ps[2 * z] = __float_as_uint(re);
                    ps[2 * z + 1] = __float_as_uint(im);
                    lm[2 * z] = __float_as_uint(re * hz);
                    lm[2 * z + 1] = __float_as_uint(im * hz);
                }
            }

            // ------------------------------------------------------------ reverse (adjoint) sweep, one step per block
            float* gxrow = NEED_GX ? p.gx + (valid ? b : 0) * p.ldgx : nullptr;
            // dL/dx of four consecutive blocks is collected in registers and stored as 80 contiguous bytes (five float4
            // when the row is 16-byte aligned): a thread's 20-byte pieces, 1,200 bytes apart from its neighbours', were
            // partial-sector writes (dL/dx on top of the weight gradients: 2.2 -> 2.1 ms per 1M-sample step, most of which is the
            // moments themselves)
            float gxq[NEED_GX ? 20 : 1];
            const bool gx_vec = NEED_GX && ((reinterpret_cast<uintptr_t>(gxrow) & 15) == 0);
            float th[NQ], uv[NQ];        // uv: the inputs behind th (frequency-layer gradients)
            load_inputs(p.K - 1, uv);
            angles_from(p.K - 1, uv, th);
            for (int k = p.K - 1; k >= 0; --k) {
                float un[NQ];
                load_inputs(k > 0 ? k - 1 : 0, un);
                // (ps, lm) = the block's output cut: one split feeds the un-apply GEMMs and the outer product
                tc_store_operand2<false>(tAp, op, op + 16384, ps);
                signal(bar_a_t);
                tc_store_operand2<true>(tAl, op + 32768, op + 49152, lm);
                signal(bar_a2_t);
                // this slot's running sums of block k (written a whole round ago): loaded now, added when the outer product
                // is drained.  The slots' accumulators (2 x 148 x K x 8 KB = 142 MB at K = 60) cycle through L2 once per round
                // and are re-fetched from HBM (ncu: 4.5 GB read + 4.0 GB written per 1M-sample step, 16 % of the HBM
                // bandwidth over the kernel).  Tried and dropped: an L2 prefetch of block k-1's lines from here (4 % slower),
                // evict_last / evict_first fractional L2 policies on these accesses (no change in traffic or time), issuing these
                // loads before the operand split (no change).
                float4* ak = reinterpret_cast<float4*>(acc + ((size_t)(t * gridDim.x + blockIdx.x) * p.K + k) * kTcAccLen) + quarter * 128 + lane;
                float4 old4[4];
                if (round > 0) {
#pragma unroll
                    for (int v = 0; v < 4; ++v) old4[v] = __ldcg(ak + 32 * v);
                }
                // psi first: its un-apply GEMM completes (bar_x) while the lam GEMM still runs, so the norm, the phase table
                // and psi's conjugate phases are computed under the lam GEMM
                wait_on(bar_x_t, xpar);
                tc_load_state(tDp, ps);
                u64 ph[16];
                float inv_c2 = 1.f;
                if (k > 0) {
                    tc_phase_table(th, 1.f, ph);
                    // conjugate phases; the scale restores |psi| = sA (the truncating accumulation shrinks both states
                    // by the same factor per GEMM), applied to lam as well
                    float nr4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int z = 0; z < 32; ++z) {
                        const float re = __uint_as_float(ps[2 * z]), im = __uint_as_float(ps[2 * z + 1]);
                        nr4[z & 3] = fmaf(re, re, fmaf(im, im, nr4[z & 3]));
                    }
                    const float nr = (nr4[0] + nr4[1]) + (nr4[2] + nr4[3]);
                    const float corr = kTcSA * rsqrtf(nr);
                    inv_c2 = nr * (1.f / (kTcSA * kTcSA));
#pragma unroll
                    for (int i = 0; i < 16; ++i) ph[i] = mul2<0>(corr, ph[i]);
                    tc_apply_phases<true>(ps, ph);
                }
                wait_on(bar_d_t, dpar);
                tc_load_state(tDl, lm);
                if (k > 0) tc_apply_phases<true>(lm, ph);
                // encoding-angle gradients of block k (Hadamard basis, at the block's encoding layer):
                // sum_z (1 - 2 z_q) Im(conj(mu_z) phi_z) is unchanged by the common conjugate phases up to their scale corr^2
                if constexpr (WANT_GX) {
                    float gq[5];
                    tc_xgrad_tree(ps, lm, gq);
                    float fv[FREQ_GRAD ? 16 : 1];
                    if constexpr (FREQ_GRAD) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) fv[i] = 0.f;
                    }
                    const float xs = xscale * inv_c2;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const float gxv = gq[q] * xs;
                        if constexpr (NEED_GX) gxq[q] = gxv;
                        if constexpr (FREQ_GRAD) {
                            fv[2 * q] = gxv * uv[q];
                            fv[2 * q + 1] = gxv;
                        }
                    }
                    if constexpr (FREQ_GRAD) {
                        const float ft = butterfly_reduce<float, 16>(fv, lane);
                        if ((lane & 1) == 0) atomicAdd(frow + (int64_t)k * 16 + (lane >> 1), ft);
                    }
                    if constexpr (NEED_GX) {
                        // gxq[0..4] = block k, [5..9] = block k + 1, ...: store when k reaches a multiple of 4
                        if ((k & 3) == 0 && valid) {
                            const int nb = p.K - k < 4 ? p.K - k : 4;
                            float* dst = gxrow + (int64_t)k * NQ;
                            if (gx_vec && nb == 4) {
#pragma unroll
                                for (int v = 0; v < 5; ++v)
                                    __stcs(reinterpret_cast<float4*>(dst) + v, make_float4(gxq[4 * v], gxq[4 * v + 1], gxq[4 * v + 2], gxq[4 * v + 3]));
                            } else {
#pragma unroll
                                for (int i = 0; i < 20; ++i)
                                    if (i < nb * NQ) dst[i] = gxq[i];
                            }
                        }
#pragma unroll
                        for (int i = 19; i >= NQ; --i) gxq[i] = gxq[i - NQ];      // make room for block k - 1
                    }
                }
                // drain the outer product of block k.  Warp `quarter` owns rows 16 quarter .. + 15 of D_G (lanes 0..15 of its
                // subpartition) = amplitudes i = 8 quarter .. + 7 of lam, real rows then imaginary rows; thread T of the
                // 16x256b fragment holds, for j = 4 g + (T % 4):  D[(i, re)][(j, re)], D[(i, re)][(j, im)], D[(i, im)][(j, re)],
                // D[(i, im)][(j, im)] with i = 8 quarter + T / 4, so  Y[j][i] = sum_b psi_j conj(lam_i)  folds in registers:
                //   Re = D[re][re] + D[im][im],   Im = D[re][im] - D[im][re]
                wait_on(bar_g_t, gpar);
                {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {      // 32 columns at a time
                        uint32_t r[16];
                        tc::tmem_ld_16x256b_x4(tAp + 32u * h, r);
                        tc::tmem_wait_ld();
                        if (dbg && blockIdx.x == 0 && t == 0 && round == 0 && k == p.K - 1)
#pragma unroll
                            for (int i = 0; i < 16; ++i) dbg[(size_t)srow_t * 32 + 16 * h + i] = __uint_as_float(r[i]);
#pragma unroll
                        for (int v = 0; v < 2; ++v) {
                            float4 y;
                            y.x = __uint_as_float(r[8 * v]) + __uint_as_float(r[8 * v + 3]);
                            y.y = __uint_as_float(r[8 * v + 1]) - __uint_as_float(r[8 * v + 2]);
                            y.z = __uint_as_float(r[8 * v + 4]) + __uint_as_float(r[8 * v + 7]);
                            y.w = __uint_as_float(r[8 * v + 5]) - __uint_as_float(r[8 * v + 6]);
                            if (!tile_live) y = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (round > 0) { y.x += old4[2 * h + v].x; y.y += old4[2 * h + v].y; y.z += old4[2 * h + v].z; y.w += old4[2 * h + v].w; }
                            __stcg(ak + 32 * (2 * h + v), y);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < NQ; ++q) uv[q] = un[q];
                angles_from(k > 0 ? k - 1 : 0, uv, th);
            }
            // a barrier wait that timed out leaves garbage: poison this warp's partial sums so that the loss and
            // every gradient of the step read NaN instead of a plausible number
            if (lane == 0 && __ldcg(err) != 0) {
                atomicAdd(p.mpart, __int_as_float(0x7fc00000));      // row 0: the only row finalize reads for the moments
                atomicAdd(srow + 1, __int_as_float(0x7fc00000));
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == G::COMPUTE_WARPS) tc::tmem_dealloc512(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// slot sums: ysum[k][f] = sum over the slots of acc[slot][k][f], in slot order, in fp64.  grid = K * 8 CTAs of 256
// threads (thread = one float of a block's 2,048; a slot's floats are contiguous, so a warp reads 128-byte lines) —
// the 142 MB of accumulators stream through all SMs instead of through the moment kernel's K CTAs.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tc_slot_reduce_kernel(const float* __restrict__ acc, int nslots, int K,
                                                             double* __restrict__ ysum) {
    const int k = blockIdx.x >> 3, f = ((blockIdx.x & 7) << 8) + threadIdx.x;
    const float* a = acc + (size_t)k * kTcAccLen + f;
    const size_t stride = (size_t)K * kTcAccLen;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;      // four chains, combined in a fixed order
    int sl = 0;
    for (; sl + 4 <= nslots; sl += 4) {
        s0 += (double)__ldcs(a + (size_t)sl * stride);
        s1 += (double)__ldcs(a + (size_t)(sl + 1) * stride);
        s2 += (double)__ldcs(a + (size_t)(sl + 2) * stride);
        s3 += (double)__ldcs(a + (size_t)(sl + 3) * stride);
    }
    for (; sl < nslots; ++sl) s0 += (double)__ldcs(a + (size_t)sl * stride);
    ysum[(size_t)k * kTcAccLen + f] = (s0 + s1) + (s2 + s3);
}

// ---------------------------------------------------------------------------------------------------------
// moments from the outer products: grid = K CTAs of 1,024 threads (thread = one entry of the 32 x 32 complex Y_k,
// fp64 in shared memory).  Adds the 15 moments of every sublayer of block k to partial row 0 (the finalize kernels
// sum the rows): mrow0[s * 16 + 3 q + {0, 1, 2}] += Im tr({X, Y, Z}_q Y) at the cut after the rotations of sublayer s.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) tc_moment_kernel(const double* __restrict__ ysum, const unsigned* __restrict__ gmax,
                                                         const float* __restrict__ hdiag, const float* __restrict__ w,
                                                         int K, DepthPack dp, float* __restrict__ mrow0) {
    constexpr int n = 5, N = 32;
    __shared__ double yr[N][N + 1], yi[N][N + 1];
    __shared__ double us[n][8];      // per qubit: U^+ = [[u00, u01], [u10, u11]] as (re, im) pairs
    const int k = blockIdx.x, t = threadIdx.x, r = t >> 5, c = t & 31;
    int s0 = 0;
    for (int kk = 0; kk < k; ++kk) s0 += dp.d[kk];
    const int d = dp.d[k];
    float hm = 0.f;
    for (int z = 0; z < N; ++z) hm = fmaxf(hm, fabsf(__ldg(hdiag + z)));
    if (!(hm > 0.f)) hm = 1.f;
    const double scale = (double)tc_gscale_pow2(__ldg(gmax)) * (double)hm / ((double)kTcSA * (double)kTcSA);
    {
        // Y[j][i] = sum_b g_b psi_j conj(lam_i), stored by the reverse kernel in fragment order: warp i >> 3, thread
        // T = 4 (i & 7) + (j & 3), float4 number (j >> 3), components 2 ((j >> 2) & 1) + {0: Re, 1: Im}
        const int j = r, i = c;
        const size_t idx = ((size_t)((i >> 3) * 4 + (j >> 3)) * 32 + 4 * (i & 7) + (j & 3)) * 4 + 2 * ((j >> 2) & 1);
        const double re = ysum[(size_t)k * kTcAccLen + idx], im = ysum[(size_t)k * kTcAccLen + idx + 1];
        yr[j][i] = re * scale;
        yi[j][i] = im * scale;
    }
    // Y <- U Y U^+ for the 2 x 2 matrix in us[q] on qubit q: one thread per 2 x 2 sub-block
    auto conj_gate = [&](int q) {
        __syncthreads();
        const int bq = 1 << q;
        if (!(r & bq) && !(c & bq)) {
            const int r1 = r | bq, c1 = c | bq;
            const double u00r = us[q][0], u00i = us[q][1], u01r = us[q][2], u01i = us[q][3];
            const double u10r = us[q][4], u10i = us[q][5], u11r = us[q][6], u11i = us[q][7];
            const double a00r = yr[r][c], a00i = yi[r][c], a01r = yr[r][c1], a01i = yi[r][c1];
            const double a10r = yr[r1][c], a10i = yi[r1][c], a11r = yr[r1][c1], a11i = yi[r1][c1];
            // T = U A
            const double t00r = u00r * a00r - u00i * a00i + u01r * a10r - u01i * a10i;
            const double t00i = u00r * a00i + u00i * a00r + u01r * a10i + u01i * a10r;
            const double t01r = u00r * a01r - u00i * a01i + u01r * a11r - u01i * a11i;
            const double t01i = u00r * a01i + u00i * a01r + u01r * a11i + u01i * a11r;
            const double t10r = u10r * a00r - u10i * a00i + u11r * a10r - u11i * a10i;
            const double t10i = u10r * a00i + u10i * a00r + u11r * a10i + u11i * a10r;
            const double t11r = u10r * a01r - u10i * a01i + u11r * a11r - u11i * a11i;
            const double t11i = u10r * a01i + u10i * a01r + u11r * a11i + u11i * a11r;
            // Y' = T U^+ :  Y'[a][b] = T[a][0] conj(U[b][0]) + T[a][1] conj(U[b][1])
            yr[r][c] = t00r * u00r + t00i * u00i + t01r * u01r + t01i * u01i;
            yi[r][c] = t00i * u00r - t00r * u00i + t01i * u01r - t01r * u01i;
            yr[r][c1] = t00r * u10r + t00i * u10i + t01r * u11r + t01i * u11i;
            yi[r][c1] = t00i * u10r - t00r * u10i + t01i * u11r - t01r * u11i;
            yr[r1][c] = t10r * u00r + t10i * u00i + t11r * u01r + t11i * u01i;
            yi[r1][c] = t10i * u00r - t10r * u00i + t11i * u01r - t11r * u01i;
            yr[r1][c1] = t10r * u10r + t10i * u10i + t11r * u11r + t11i * u11i;
            yi[r1][c1] = t10i * u10r - t10r * u10i + t11i * u11r - t11r * u11i;
        }
    };
    if (k < K - 1) {     // the output cut of every block but the last is held in the Hadamard basis
        if (t < n) {
            const double h = 0.70710678118654752440;
            us[t][0] = h; us[t][1] = 0.0; us[t][2] = h; us[t][3] = 0.0;
            us[t][4] = h; us[t][5] = 0.0; us[t][6] = -h; us[t][7] = 0.0;
        }
        for (int q = 0; q < n; ++q) conj_gate(q);
    }
    auto unring = [](int z) {        // Ring^+ on a basis index: the CNOTs (control (i+1)%n -> target i) in reverse order
        for (int i = n - 1; i >= 0; --i)
            if ((z >> ((i + 1) % n)) & 1) z ^= 1 << i;
        return z;
    };
    const int pr = unring(r), pc = unring(c);
    for (int s = s0 + d - 1; s >= s0; --s) {
        // Y <- Ring^+ Y Ring
        __syncthreads();
        const double vr = yr[r][c], vi = yi[r][c];
        if (t < n) {     // U^+ of the sublayer's fused rotations RY(c) RZ(b) RY(a) = [[al, -conj(be)], [be, conj(al)]]
            const int q = t;
            const double a = (double)w[((int64_t)s * 3 + 0) * n + q];
            const double b = (double)w[((int64_t)s * 3 + 1) * n + q];
            const double cc_ = (double)w[((int64_t)s * 3 + 2) * n + q];
            double sa, ca, sb, cb, sc, cc;
            sincos(0.5 * a, &sa, &ca);
            sincos(0.5 * b, &sb, &cb);
            sincos(0.5 * cc_, &sc, &cc);
            const double ar = cb * (cc * ca - sc * sa), ai = -sb * (cc * ca + sc * sa);
            const double br = cb * (sc * ca + cc * sa), bi = sb * (cc * sa - sc * ca);
            // U^+ = [[conj(al), conj(be)], [-be, al]]
            us[q][0] = ar; us[q][1] = -ai; us[q][2] = br; us[q][3] = -bi;
            us[q][4] = -br; us[q][5] = -bi; us[q][6] = ar; us[q][7] = ai;
        }
        __syncthreads();
        yr[pr][pc] = vr;
        yi[pr][pc] = vi;
        __syncthreads();
        if (t < 3 * n) {
            const int q = t / 3, v = t % 3, bq = 1 << q;
            double acc = 0.0;
            for (int z = 0; z < N; ++z) {
                const bool one = (z >> q) & 1;
                if (v == 0) acc += yi[z ^ bq][z];                               // Im tr(X_q Y)
                else if (v == 1) acc += one ? yr[z ^ bq][z] : -yr[z ^ bq][z];   // Im tr(Y_q Y)
                else acc += one ? -yi[z][z] : yi[z][z];                         // Im tr(Z_q Y)
            }
            mrow0[(int64_t)s * 16 + t] += (float)acc;
        }
        for (int q = 0; q < n; ++q) conj_gate(q);
    }
}

}  // namespace qon
