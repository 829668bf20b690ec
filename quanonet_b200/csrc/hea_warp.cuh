// Small-batch latency tier (n <= 5): ONE AMPLITUDE PER LANE, 2^n lanes per sample, 32 >> n samples per warp.
//
// The reference trains with batch_size = 100 (utils/common.py:128) — far too few samples to fill a B200
// with one thread per sample (100 threads = 4 warps, each walking ~340k instructions).  Here a sample's
// critical path is ~600 shuffle latencies: a gate on qubit q is one __shfl_xor pair plus 2 dependent FMAs
// per lane, the CNOT ring of a sublayer is a lane permutation that is folded into the shuffles of the
// adjacent gate (the composed ring map is computed once per lane), and nothing on the path waits on memory:
//   * the prep tables (gate coefficients, dE/dtheta coefficients) are staged into shared memory once per
//     CTA — a sublayer touches 80 B of table, so reading it from L1/L2 as the sweep goes costs a cold miss
//     every other sublayer on every SM (measured: 1650 cycles / sublayer before, ~400 after);
//   * sin / cos of every encoding half-angle of the sample are computed once, lane-parallel, into shared
//     memory (K*n values spread over 2^n lanes) and reused by the forward and the reverse sweep;
//   * the loops are software-pipelined by hand: the coefficients of the next sublayer (shared-memory loads,
//     the RX fold, the per-lane signs) are formed while the state chain of the current one runs, and in the
//     reverse sweep the moment reduction of sublayer s+1 is issued alongside the chain of sublayer s.  All
//     loop bounds come from kernel parameters (DepthPack in the constant bank), so the compiler can prove
//     the shuffles run converged and keeps each sublayer one basic block.
// Arithmetic is the same adjoint differentiation as hea_reg.cuh (Pauli moments per fused gate, warp
// butterfly, per-warp partial rows, fp64 fixed-order finalize) — results agree to rounding.
#pragma once
#include "hea_reg.cuh"

namespace qon {

constexpr int kWarpThreads = 128;

template <typename T> struct alignas(2 * sizeof(T)) Vec2 { T x, y; };

template <typename T>
__host__ __device__ inline size_t warp_smem_bytes(int n, int K, int S, bool want_gx, bool freq_grad, int threads) {
    const size_t tbl = (size_t)S * n * sizeof(Vec4<T>) * (want_gx ? 2 : 1);
    const size_t per_sample = (size_t)K * n * (freq_grad ? 3 : 2) * sizeof(T);   // (sin, cos) [+ source value u]
    const int spw = n >= 5 ? 1 : (32 >> n);                                          // samples per warp
    return tbl + (size_t)(threads / 32) * spw * per_sample;
}

//   ENC = 0: encoding angles x given;  1: angles formed in-kernel from (u0, u1, fw, fb) — the frequency layers of
//   core/models_pt.py:14-68;  2: as 1, and dL/dfw, dL/dfb are reduced over the batch in-kernel (hea_reg.cuh).
template <typename T, int N, bool GRAD, bool NEED_GX, int ENC, int THREADS>
__global__ void __launch_bounds__(THREADS) hea_warp_kernel(const HeaParams<T> p, const DepthPack dp) {
    constexpr bool FREQ_GRAD = GRAD && ENC == 2;
    constexpr bool WANT_GX = NEED_GX || FREQ_GRAD;
    static_assert(!(NEED_GX && ENC != 0), "grad_x is only materialised when x is");
    constexpr int NA = 1 << N;
    constexpr int SPW = 32 >> N;
    constexpr int VP = moment_slots(N);
    constexpr int FVP = freq_slots(N);
    constexpr int WARPS = THREADS / 32;
    constexpr int STR = 32 / VP;
    constexpr int TOP = N - 1;             // the qubit whose gate carries the ring permutation
    extern __shared__ __align__(32) unsigned char warp_smem[];

    const int S = p.S, K = p.K, SN = p.S * N, KN = p.K * N;
    Vec4<T>* uc_s = reinterpret_cast<Vec4<T>*>(warp_smem);
    Vec4<T>* rc_s = uc_s + SN;
    Vec2<T>* sc_all = reinterpret_cast<Vec2<T>*>(uc_s + (WANT_GX ? 2 : 1) * SN);
    T* uv_all = reinterpret_cast<T*>(sc_all + (size_t)WARPS * SPW * KN);          // FREQ_GRAD only

    for (int i = threadIdx.x; i < SN; i += THREADS) {
        uc_s[i] = ldg4(p.ucoef + i);
        if constexpr (WANT_GX) rc_s[i] = ldg4(p.rcoef + i);
    }

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int amp = lane & (NA - 1);
    const int sidx = lane >> N;
    const int base = lane & ~(NA - 1);
    // composed CNOT ring (control (i+1)%N -> target i, i = 0..N-1) as a source-index map, both directions
    auto ring_fwd_src = [](int a) {
#pragma unroll
        for (int i = N - 1; i >= 0; --i) a ^= ((a >> ((i + 1) % N)) & 1) << i;
        return a;
    };
    auto ring_rev_src = [](int a) {
#pragma unroll
        for (int i = 0; i < N; ++i) a ^= ((a >> ((i + 1) % N)) & 1) << i;
        return a;
    };
    // forward: the ring follows the gate on TOP, so lane l takes the gate output of lane rf = ring_fwd_src(l);
    // reverse: the ring is undone by its own permutation (fusing it costs 8 indexed shuffles: measured slower)
    const int rf = N > 1 ? ring_fwd_src(amp) : amp;
    const int f_mine = base | rf, f_part = base | (rf ^ (1 << TOP));
    const int r_mine = base | (N > 1 ? ring_rev_src(amp) : amp);
    T sgn[N], sgf[N];   // -1 where the qubit's bit is set in this lane's amplitude index (sgf: in rf's, for TOP)
#pragma unroll
    for (int q = 0; q < N; ++q) {
        sgn[q] = ((amp >> q) & 1) ? T(-1) : T(1);
        sgf[q] = (((q == TOP ? rf : amp) >> q) & 1) ? T(-1) : T(1);
    }
    const T hd = p.pauli == 0 ? __ldg(p.hdiag + amp) : T(0);
    Vec2<T>* sc = sc_all + (size_t)(warp * SPW + sidx) * KN;   // (sin, cos) of theta/2 per encoding column
    T* uv = FREQ_GRAD ? uv_all + (size_t)(warp * SPW + sidx) * KN : nullptr;   // source value u of the column

    const int64_t gwarp = (int64_t)blockIdx.x * WARPS + warp;
    const int64_t nwarps = (int64_t)gridDim.x * WARPS;
    const int64_t ntiles = (p.B + SPW - 1) / SPW;
    T* mrow = GRAD ? p.mpart + gwarp * p.rowlen : nullptr;
    T* frow = GRAD ? mrow + (int64_t)S * VP : nullptr;               // frequency-gradient slots
    T* srow = GRAD ? frow + (int64_t)K * FVP : nullptr;              // [sum g, sum residual^2]
    __syncthreads();

    T re, im, lr, li;

    // per-lane coefficients of the n fused gates of one sublayer: mine' = cm * mine + cp * partner
    struct Coef { T cmr[N], cmi[N], cpr[N], cpi[N]; };
    auto set_coef = [&](auto Dag, Coef& c, int q, T ar, T ai, T br, T bi) {
        constexpr bool DAG = decltype(Dag)::value;   // false: U (forward sweep); true: U^dagger (reverse sweep)
        c.cmr[q] = ar;
        c.cmi[q] = DAG ? -sgn[q] * ai : sgf[q] * ai;
        c.cpr[q] = DAG ? sgn[q] * br : -sgf[q] * br;
        c.cpi[q] = DAG ? -bi : bi;
    };
    // sublayer s inside a block
    auto prepare_plain = [&](auto Dag, int s, Coef& c) {
        const Vec4<T>* uc = uc_s + s * N;
#pragma unroll
        for (int q = 0; q < N; ++q) {
            const Vec4<T> u = uc[q];
            set_coef(Dag, c, q, u.x, u.y, u.z, u.w);
        }
    };
    // sublayer s opening block k: RX(theta) folded in.  fold == false evaluates the same expressions with
    // (sin, cos) = (0, 1), which is exact — a branch-free way to serve "maybe a block opener".
    auto prepare_fold = [&](auto Dag, int s, int k, bool fold, Coef& c) {
        const Vec4<T>* uc = uc_s + s * N;
        const Vec2<T>* sk = sc + k * N;
#pragma unroll
        for (int q = 0; q < N; ++q) {
            const Vec4<T> u = uc[q];
            Vec2<T> t = sk[q];
            if (!fold) t = Vec2<T>{T(0), T(1)};
            set_coef(Dag, c, q, fma_(t.x, u.w, u.x * t.y), fma_(t.x, u.z, u.y * t.y), fma_(-t.x, u.y, u.z * t.y),
                     fma_(-t.x, u.x, u.w * t.y));
        }
    };
    auto gate = [&](const Coef& c, int q, T& xr, T& xi) {
        const T qr = shfl_xor_(xr, 1 << q), qi = shfl_xor_(xi, 1 << q);
        const T tr = fma_(-c.cmi[q], xi, c.cmr[q] * xr), ti = fma_(c.cmi[q], xr, c.cmr[q] * xi);
        xr = fma_(-c.cpi[q], qi, fma_(c.cpr[q], qr, tr));
        xi = fma_(c.cpi[q], qr, fma_(c.cpr[q], qi, ti));
    };
    // gate on TOP with "mine" and "partner" both arriving through a lane permutation (the ring)
    auto gate_perm = [&](const Coef& c, int src_mine, int src_part, T& xr, T& xi) {
        const T mr = shfl_idx_(xr, src_mine), mi = shfl_idx_(xi, src_mine);
        const T qr = shfl_idx_(xr, src_part), qi = shfl_idx_(xi, src_part);
        xr = fma_(-c.cmi[TOP], mi, c.cmr[TOP] * mr) + fma_(-c.cpi[TOP], qi, c.cpr[TOP] * qr);
        xi = fma_(c.cmi[TOP], mr, c.cmr[TOP] * mi) + fma_(c.cpi[TOP], qr, c.cpr[TOP] * qi);
    };
    auto fwd_chain = [&](const Coef& c) {
#pragma unroll
        for (int q = 0; q < TOP; ++q) gate(c, q, re, im);
        if constexpr (N > 1) gate_perm(c, f_mine, f_part, re, im);
        else gate(c, 0, re, im);
    };

    // CTA-uniform trip count: with it every shuffle below is provably converged
    for (int64_t tile0 = (int64_t)blockIdx.x * WARPS; tile0 < ntiles; tile0 += nwarps) {
        const int64_t b = (tile0 + warp) * SPW + sidx;
        const bool valid = b < p.B;
        const int64_t bc = valid ? b : p.B - 1;
        __syncwarp();
        if constexpr (ENC == 0) {
            const T* xrow = p.x + bc * p.ldx;
#pragma unroll 4
            for (int c = amp; c < KN; c += NA) {
                T sn, cs;
                sincos_half(__ldg(xrow + c), sn, cs);
                sc[c] = Vec2<T>{sn, cs};
            }
        } else {
            const T* u0row = p.u0 ? p.u0 + bc * p.ldu0 : nullptr;
            const T* u1row = p.u1 + bc * p.ldu1;
            const int c0 = p.K0 * N;                       // columns below c0 read source 0
#pragma unroll 4
            for (int c = amp; c < KN; c += NA) {
                const T u = __ldg((c < c0 ? u0row : u1row) + __ldg(p.uidx + c));
                T sn, cs;
                sincos_half(fma_(u, __ldg(p.fw + c), p.fb ? __ldg(p.fb + c) : T(0)), sn, cs);
                sc[c] = Vec2<T>{sn, cs};
                if constexpr (FREQ_GRAD) uv[c] = u;
            }
        }
        __syncwarp();

        // ---------------- forward sweep ----------------
        re = amp == 0 ? T(1) : T(0);
        im = T(0);
        {
            Coef cur, nxt;
            prepare_fold(IntC<0>{}, 0, 0, true, cur);
            int s = 0;
            for (int k = 0; k < K; ++k) {
                const int d = dp.d[k];
#pragma unroll 1
                for (int j = 0; j + 1 < d; ++j, ++s) {
                    prepare_plain(IntC<0>{}, s + 1, nxt);
                    fwd_chain(cur);
                    cur = nxt;
                }
                const bool more = k + 1 < K;
                prepare_fold(IntC<0>{}, more ? s + 1 : s, more ? k + 1 : k, true, nxt);
                fwd_chain(cur);
                cur = nxt;
                ++s;
            }
        }

        // ---------------- expectation value ----------------
        if (p.pauli == 0) {
            lr = hd * re;
            li = hd * im;
        } else {
            const bool isY = p.pauli == 2;
            lr = p.offset * re;
            li = p.offset * im;
#pragma unroll
            for (int q = 0; q < N; ++q) {
                const T fr = shfl_xor_(re, 1 << q), fi = shfl_xor_(im, 1 << q);
                if (!isY) {
                    lr = fma_(p.coeff, fr, lr); li = fma_(p.coeff, fi, li);
                } else {    // (Y psi)_k = +i psi_flip if bit set else -i psi_flip
                    const T c = -sgn[q] * p.coeff;
                    lr = fma_(-c, fi, lr); li = fma_(c, fr, li);
                }
            }
        }
        T e = fma_(im, li, re * lr);
#pragma unroll
        for (int m = 1; m < NA; m <<= 1) e += shfl_xor_(e, m);
        if (valid && amp == 0 && p.out) p.out[b] = e;

        if constexpr (GRAD) {
            T g = T(0);
            if (p.target) {   // fused MSE: g = dL/dout for L = gscale/2 * sum (out + bias - y)^2
                T resid = T(0);
                if (valid) {
                    resid = e + (p.bias ? __ldg(p.bias) : T(0)) - __ldg(p.target + b);
                    g = p.gscale * resid;
                    if (amp == 0 && p.gbuf) p.gbuf[b] = g;
                }
                T sg = amp == 0 ? g : T(0), sq = amp == 0 ? resid * resid : T(0);
#pragma unroll
                for (int m = 16; m >= NA; m >>= 1) { sg += shfl_xor_(sg, m); sq += shfl_xor_(sq, m); }
                if (lane == 0) { atomicAdd(srow, sg); atomicAdd(srow + 1, sq); }
            } else if (valid) {
                g = __ldg(p.gout + b);
            }
            lr *= g;
            li *= g;
            T* gxrow = NEED_GX ? p.gx + (valid ? b : 0) * p.ldgx : nullptr;

            // Pauli moments of (lam, psi) on qubit q from this lane's and its partner's amplitudes
            auto moments = [&](int q, T(&mv)[VP], T mr, T mi, T nr, T ni, T qr, T qi) {
                mv[3 * q] = fma_(-ni, qr, nr * qi);                       // Im(conj(l_mine) p_partner)
                mv[3 * q + 1] = -sgn[q] * fma_(ni, qi, nr * qr);          // -/+ Re(conj(l_mine) p_partner)
                mv[3 * q + 2] = sgn[q] * fma_(-ni, mr, nr * mi);          // +/- Im(conj(l_mine) p_mine)
            };
            // reverse chain of one sublayer: undo the ring, then per gate, high qubit first: moments of the
            // state after the gate, then the un-application on psi and lam
            auto rev_chain = [&](const Coef& c, T(&mv)[VP]) {
#pragma unroll
                for (int i = 3 * N; i < VP; ++i) mv[i] = T(0);
                if constexpr (N > 1) {   // undo the ring
                    re = shfl_idx_(re, r_mine); im = shfl_idx_(im, r_mine);
                    lr = shfl_idx_(lr, r_mine); li = shfl_idx_(li, r_mine);
                }
#pragma unroll
                for (int q = TOP; q >= 0; --q) {
                    const T qr = shfl_xor_(re, 1 << q), qi = shfl_xor_(im, 1 << q);
                    const T kr = shfl_xor_(lr, 1 << q), ki = shfl_xor_(li, 1 << q);
                    const T tr = fma_(-c.cmi[q], im, c.cmr[q] * re), ti = fma_(c.cmi[q], re, c.cmr[q] * im);
                    const T ur = fma_(-c.cmi[q], li, c.cmr[q] * lr), ui = fma_(c.cmi[q], lr, c.cmr[q] * li);
                    moments(q, mv, re, im, lr, li, qr, qi);
                    re = fma_(-c.cpi[q], qi, fma_(c.cpr[q], qr, tr));
                    im = fma_(c.cpi[q], qr, fma_(c.cpr[q], qi, ti));
                    lr = fma_(-c.cpi[q], ki, fma_(c.cpr[q], kr, ur));
                    li = fma_(c.cpi[q], kr, fma_(c.cpr[q], ki, ui));
                }
            };
            // batch reduction of one sublayer's moments (+ the per-sample dL/dtheta when it opens block k >= 0)
            auto flush = [&](T(&mv)[VP], int s, int k) {
                if constexpr (WANT_GX && SPW > 1) {
                    if (k >= 0) {   // per-sample: reduce over the sample's own lanes
#pragma unroll
                        for (int q = 0; q < N; ++q) {
                            T mx = mv[3 * q], my = mv[3 * q + 1], mz = mv[3 * q + 2];
#pragma unroll
                            for (int m = 1; m < NA; m <<= 1) {
                                mx += shfl_xor_(mx, m); my += shfl_xor_(my, m); mz += shfl_xor_(mz, m);
                            }
                            const Vec4<T> r = rc_s[s * N + q];
                            const T gxv = fma_(r.z, mz, fma_(r.y, my, r.x * mx));
                            if constexpr (NEED_GX) {
                                if (amp == q && valid) gxrow[(int64_t)k * N + q] = gxv;
                            }
                            if constexpr (FREQ_GRAD) {   // theta = fw*u + fb  =>  d/dfw = gx*u, d/dfb = gx
                                T fw_ = gxv * uv[k * N + q], fb_ = gxv;      // invalid samples carry g = 0
#pragma unroll
                                for (int m = NA; m < 32; m <<= 1) { fw_ += shfl_xor_(fw_, m); fb_ += shfl_xor_(fb_, m); }
                                if (lane == q) { atomicAdd(frow + k * FVP + 2 * q, fw_); atomicAdd(frow + k * FVP + 2 * q + 1, fb_); }
                            }
                        }
                    }
                }
                const T tot = butterfly_reduce<T, VP>(mv, lane);
                if ((lane & (STR - 1)) == 0) atomicAdd(mrow + (int64_t)s * VP + lane / STR, tot);
                if constexpr (WANT_GX && SPW == 1) {   // the warp IS the sample: the batch totals are its moments
                    if (k >= 0) {
                        const int q3 = (amp < N ? amp : 0) * 3;
                        const T mx = shfl_idx_(tot, q3 * STR), my = shfl_idx_(tot, (q3 + 1) * STR),
                                mz = shfl_idx_(tot, (q3 + 2) * STR);
                        if (amp < N) {
                            const Vec4<T> r = rc_s[s * N + amp];
                            const T gxv = fma_(r.z, mz, fma_(r.y, my, r.x * mx));
                            if constexpr (NEED_GX) {
                                if (valid) gxrow[(int64_t)k * N + amp] = gxv;
                            }
                            if constexpr (FREQ_GRAD) {
                                atomicAdd(frow + k * FVP + 2 * amp, gxv * uv[k * N + amp]);
                                atomicAdd(frow + k * FVP + 2 * amp + 1, gxv);
                            }
                        }
                    }
                }
            };

            // The reverse sweep is NOT software-pipelined: measured on B200, prefetching the next sublayer's
            // coefficients and deferring the moment reduction costs more in register moves and issue slots
            // than it hides (36 us -> 40 us per 120 sublayers at B = 100, 164 us -> 208 us at B = 4096).
            Coef c;
            T mv[VP];
            int s = S;
            for (int k = K - 1; k >= 0; --k) {
                const int d = dp.d[k];
#pragma unroll 1
                for (int j = d - 1; j >= 1; --j) {
                    --s;
                    prepare_plain(IntC<1>{}, s, c);
                    rev_chain(c, mv);
                    flush(mv, s, -1);
                }
                --s;
                prepare_fold(IntC<1>{}, s, k, true, c);
                rev_chain(c, mv);
                flush(mv, s, k);
            }
        }
    }
}

}  // namespace qon
