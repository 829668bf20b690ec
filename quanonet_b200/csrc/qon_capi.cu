// C-ABI of quanonet_b200 (see include/quanonet_b200.h): argument checking, launch planning,
// the prep / finalize kernels and tier dispatch.  No torch types, no allocation, no host sync.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/quanonet_b200.h"
#include "hea_dispatch.cuh"
#include "qon_peer.cuh"

namespace qon {
namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

constexpr int kMaxQubits = 24;

// Tensor-core tier switch (hea_tc.cuh, hea_tc2.cuh): QON_TC=0/1 in the environment, or qon_tensor_tier() at run time
// (enable == 3 runs the training step as one kernel instead of forward-only + reverse-only, for A/B runs).
struct TcConfig { int enable; float* dbg; int* err; int64_t min_batch; };
TcConfig& tc_config() {
    static TcConfig c = [] {
        const char* e = getenv("QON_TC");
        const char* m = getenv("QON_TC_MIN_B");
        return TcConfig{e ? atoi(e) : 1, nullptr, nullptr, m ? (int64_t)atoll(m) : (int64_t)5121};   // measured crossover with the latency tier: ~4,100-5,000 samples in every mode (scripts/tc_modes.py, tc_ab.py)
    }();
    return c;
}

// ---------------------------------------------------------------------------------------------
// prep: per-(sublayer, qubit) gate tables, Hamiltonian diagonal, depth array, column -> source-row
// index table of the fused encoding, zeroed partial sums
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void prep_kernel(const T* __restrict__ w, int n, int S, int K, DepthPack dp,
                            Vec4<T>* ucoef, Vec4<T>* rcoef, int* depth,
                            T* hdiag, const T* ham_diag, int diag_order, double offset, double coeff,
                            T* mpart, int64_t mpart_len, int* uidx, int K0, int in0, int in1) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = tid; t < (int64_t)S * n; t += nth) {
        const int s = (int)(t / n), q = (int)(t % n);
        const double a = (double)w[((int64_t)s * 3 + 0) * n + q];
        const double b = (double)w[((int64_t)s * 3 + 1) * n + q];
        const double c = (double)w[((int64_t)s * 3 + 2) * n + q];
        double sa, ca, sb, cb, sc, cc;
        sincos(0.5 * a, &sa, &ca);
        sincos(0.5 * b, &sb, &cb);
        sincos(0.5 * c, &sc, &cc);
        // U = RY(c) RZ(b) RY(a) = [[al, -conj(be)], [be, conj(al)]]
        Vec4<T> u;
        u.x = (T)(cb * (cc * ca - sc * sa));
        u.y = (T)(-sb * (cc * ca + sc * sa));
        u.z = (T)(cb * (sc * ca + cc * sa));
        u.w = (T)(sb * (cc * sa - sc * ca));
        ucoef[t] = u;
        // U X U^dagger = rX X + rY Y + rZ Z
        double Sa, Ca, Sb, Cb, Sc, Cc;
        sincos(a, &Sa, &Ca);
        sincos(b, &Sb, &Cb);
        sincos(c, &Sc, &Cc);
        Vec4<T> r;
        r.x = (T)(Ca * Cb * Cc - Sa * Sc);
        r.y = (T)(Ca * Sb);
        r.z = (T)(-Ca * Cb * Sc - Sa * Cc);
        r.w = (T)0;
        rcoef[t] = r;
    }
    for (int64_t t = tid; t < K; t += nth) depth[t] = dp.d[t];
    if (uidx) {
        for (int64_t c = tid; c < (int64_t)n * K; c += nth) {
            const int k = (int)(c / n);
            const int64_t local = k < K0 ? c : c - (int64_t)K0 * n;
            uidx[c] = (int)(local % (k < K0 ? in0 : in1));
        }
    }
    const int64_t N = (int64_t)1 << n;
    for (int64_t k = tid; k < N; k += nth) {
        if (ham_diag) {
            int64_t src = k;
            if (diag_order == QON_DIAG_MSB0) {   // our bit q  <->  their bit n-1-q
                src = 0;
                for (int q = 0; q < n; ++q) src |= ((k >> q) & 1) << (n - 1 - q);
            }
            hdiag[k] = ham_diag[src];
        } else {
            hdiag[k] = (T)(offset + coeff * (double)(n - 2 * __popcll((unsigned long long)k)));
        }
    }
    for (int64_t t = tid; t < mpart_len; t += nth) mpart[t] = (T)0;
}

// ---------------------------------------------------------------------------------------------
// finalize: sum the per-warp / per-CTA moment rows (fixed order, fp64) and turn the three Pauli
// moments of each fused gate into the gradients of its three angles:
//   U = RY(c) RZ(b) RY(a):  dc = mY ; db = cos(c) mZ + sin(c) mX ;
//                           da = cos(b) mY - sin(b) (cos(c) mX - sin(c) mZ)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void finalize_kernel(const T* __restrict__ mpart, int rows, int64_t rowlen, int VP, int n,
                                const T* __restrict__ w, T* __restrict__ grad_w) {
    extern __shared__ double sm[];   // [RG][VP]
    const int s = blockIdx.x;
    const int RG = blockDim.x / VP;
    const int slot = threadIdx.x % VP, rg = threadIdx.x / VP;
    if (rg < RG) {
        double acc = 0.0;
        for (int r = rg; r < rows; r += RG) acc += (double)mpart[(int64_t)r * rowlen + (int64_t)s * VP + slot];
        sm[rg * VP + slot] = acc;
    }
    __syncthreads();
    if (threadIdx.x < n) {
        const int q = threadIdx.x;
        double m[3];
        for (int v = 0; v < 3; ++v) {
            double acc = 0.0;
            for (int g = 0; g < RG; ++g) acc += sm[g * VP + 3 * q + v];
            m[v] = acc;
        }
        const double b = (double)w[((int64_t)s * 3 + 1) * n + q];
        const double c = (double)w[((int64_t)s * 3 + 2) * n + q];
        double Sb, Cb, Sc, Cc;
        sincos(b, &Sb, &Cb);
        sincos(c, &Sc, &Cc);
        grad_w[((int64_t)s * 3 + 0) * n + q] = (T)(Cb * m[1] - Sb * (Cc * m[0] - Sc * m[2]));
        grad_w[((int64_t)s * 3 + 1) * n + q] = (T)(Cc * m[2] + Sc * m[0]);
        grad_w[((int64_t)s * 3 + 2) * n + q] = (T)m[1];
    }
}

// frequency-layer gradients and the two scalar sums of the fused-encoding training kernels:
// block k < K: slots (2q, 2q+1) of block k -> grad_fw / grad_fb[k*n+q];  block K: [sum g, sum resid^2]
template <typename T>
__global__ void finalize_enc_kernel(const T* __restrict__ mpart, int rows, int64_t rowlen, int64_t off, int FVP,
                                    int n, int K, T* __restrict__ grad_fw, T* __restrict__ grad_fb,
                                    T* __restrict__ sums) {
    __shared__ double sm[256];
    const int k = blockIdx.x;
    const int width = k < K ? FVP : 2;
    const int64_t base = off + (k < K ? (int64_t)k * FVP : (int64_t)K * FVP);
    const int RG = blockDim.x / width;
    const int slot = threadIdx.x % width, rg = threadIdx.x / width;
    double acc = 0.0;
    if (rg < RG)
        for (int r = rg; r < rows; r += RG) acc += (double)mpart[(int64_t)r * rowlen + base + slot];
    sm[threadIdx.x] = rg < RG ? acc : 0.0;
    __syncthreads();
    if (threadIdx.x < width) {
        double tot = 0.0;
        for (int g = 0; g < RG; ++g) tot += sm[g * width + threadIdx.x];
        if (k < K) {
            const int q = threadIdx.x >> 1;
            if (q < n) {
                if ((threadIdx.x & 1) == 0) { if (grad_fw) grad_fw[(int64_t)k * n + q] = (T)tot; }
                else if (grad_fb) grad_fb[(int64_t)k * n + q] = (T)tot;
            }
        } else if (sums) {
            sums[threadIdx.x] = (T)tot;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// finalize + exchange in ONE kernel (data-parallel training, fp32): the work of finalize_kernel and
// finalize_enc_kernel, with every gradient pushed straight into the peers' slots of the symmetric buffers as
// it is produced (qon_peer.cuh protocol), then the last CTA to finish signals the peers, waits for theirs and
// writes the rank-ordered sum into the flat gradient buffer.  grid = S + K + 1 CTAs:
//   CTA s < S      : the 3n angle gradients of sublayer s      -> flat[w_off + s*3n ...]
//   CTA S + k      : dL/dfw, dL/dfb of encoding block k         -> flat[fw_off + k*n ...], flat[fb_off + k*n ...]
//   CTA S + K      : [sum g, sum residual^2] -> flat[sums_off ...]; zeros for every other index of the buffer
// ---------------------------------------------------------------------------------------------
struct FlatLayout { int64_t w_off, fw_off, fb_off, sums_off, len; };

__global__ void __launch_bounds__(256) finalize_exchange_kernel(const float* __restrict__ mpart, int rows_m, int rows, int64_t rowlen,
                                                                int VP, int FVP, int n, int S, int K,
                                                                const float* __restrict__ w, float* flat, FlatLayout fl,
                                                                PeerPtrs pp, int world, int rank, int64_t max_len,
                                                                long long timeout_cycles) {
    __shared__ double sm[256];
    __shared__ int s_last, s_timeout;
    const int t = threadIdx.x, b = blockIdx.x;
    char* mine = pp.p[rank];
    unsigned* my_flags = reinterpret_cast<unsigned*>(mine);
    unsigned* epoch_ctr = reinterpret_cast<unsigned*>(mine + 128);
    unsigned* err_word = reinterpret_cast<unsigned*>(mine + 132);
    unsigned* done_ctr = reinterpret_cast<unsigned*>(mine + 136);
    const unsigned e = *epoch_ctr + 1u;              // only the last CTA of a launch advances it
    const size_t set_off = (size_t)(e & 1u) * world * (size_t)max_len;
    auto push = [&](int64_t idx, float v) {
        for (int p = 0; p < world; ++p)
            (reinterpret_cast<float*>(pp.p[p] + kPeerHeaderBytes) + set_off + (size_t)rank * max_len)[idx] = v;
    };

    if (b < S) {
        const int RG = 256 / VP, slot = t % VP, rg = t / VP;
        double acc = 0.0;
        if (rg < RG)
            for (int r = rg; r < rows_m; r += RG) acc += (double)mpart[(int64_t)r * rowlen + (int64_t)b * VP + slot];
        sm[t] = rg < RG ? acc : 0.0;
        __syncthreads();
        if (t < n) {
            double m[3];
            for (int v = 0; v < 3; ++v) {
                double a2 = 0.0;
                for (int g = 0; g < RG; ++g) a2 += sm[g * VP + 3 * t + v];
                m[v] = a2;
            }
            const double bb = (double)w[((int64_t)b * 3 + 1) * n + t];
            const double cc = (double)w[((int64_t)b * 3 + 2) * n + t];
            double Sb, Cb, Sc, Cc;
            sincos(bb, &Sb, &Cb);
            sincos(cc, &Sc, &Cc);
            const int64_t base = fl.w_off + (int64_t)b * 3 * n + t;
            push(base, (float)(Cb * m[1] - Sb * (Cc * m[0] - Sc * m[2])));
            push(base + n, (float)(Cc * m[2] + Sc * m[0]));
            push(base + 2 * n, (float)m[1]);
        }
    } else {
        const int k = b - S;
        const int width = k < K ? FVP : 2;
        const int64_t base = (int64_t)S * VP + (k < K ? (int64_t)k * FVP : (int64_t)K * FVP);
        const int RG = 256 / width, slot = t % width, rg = t / width;
        double acc = 0.0;
        if (rg < RG)
            for (int r = rg; r < rows; r += RG) acc += (double)mpart[(int64_t)r * rowlen + base + slot];
        sm[t] = rg < RG ? acc : 0.0;
        __syncthreads();
        if (t < width) {
            double tot = 0.0;
            for (int g = 0; g < RG; ++g) tot += sm[g * width + t];
            if (k < K) {
                const int q = t >> 1;
                if (q < n && fl.fw_off >= 0) push(((t & 1) ? fl.fb_off : fl.fw_off) + (int64_t)k * n + q, (float)tot);
            } else {
                push(fl.sums_off + t, (float)tot);
            }
        }
        if (k == K) {      // indices no CTA produces (other parameters, padding) must not carry stale slot contents
            const int64_t nw = (int64_t)3 * n * S, ne = (int64_t)n * K;
            for (int64_t i = t; i < fl.len; i += 256) {
                const bool covered = (i >= fl.w_off && i < fl.w_off + nw) || (i >= fl.sums_off && i < fl.sums_off + 2) ||
                                     (fl.fw_off >= 0 && ((i >= fl.fw_off && i < fl.fw_off + ne) ||
                                                         (i >= fl.fb_off && i < fl.fb_off + ne)));
                if (!covered) push(i, 0.f);
            }
        }
    }
    // last CTA of this rank: exchange
    __threadfence_system();
    __syncthreads();
    if (t == 0) {
        s_timeout = 0;
        s_last = atomicAdd(done_ctr, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (t == 0) *done_ctr = 0u;
    if (t < world) st_release_sys(reinterpret_cast<unsigned*>(pp.p[t]) + rank, e);
    if (t < world) {
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(my_flags + t) - e) < 0) {
            if (clock64() - t0 > timeout_cycles) { s_timeout = 1; break; }
        }
    }
    __syncthreads();
    const bool bad = s_timeout != 0;
    const float* slots = reinterpret_cast<const float*>(mine + kPeerHeaderBytes) + set_off;
    for (int64_t i = t; i < fl.len; i += 256) {
        float acc = 0.f;
        for (int p = 0; p < world; ++p) acc += __ldcg(slots + (size_t)p * max_len + i);
        flat[i] = bad ? __int_as_float(0x7fc00000) : acc;
    }
    if (t == 0) {
        *epoch_ctr = e;
        if (bad) *err_word = e;
    }
}

// ---------------------------------------------------------------------------------------------
// FP32 FFMA peak probe
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float y, float z) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], y, z);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456f) out[0] = s;   // never true; keeps the chain alive
}

// ---------------------------------------------------------------------------------------------
// planning
// ---------------------------------------------------------------------------------------------
struct DeviceInfo { int sms = 0; bool ok = false; };

bool device_info(int* dev_out, DeviceInfo* info) {
    static std::mutex mu;
    static DeviceInfo cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return false; }
    std::lock_guard<std::mutex> lk(mu);
    if (!cache[dev].ok) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
            cudaGetLastError();
            return false;
        }
        cache[dev].sms = sms;
        cache[dev].ok = true;
    }
    *dev_out = dev;
    *info = cache[dev];
    return true;
}

constexpr int kModes = 6;   // see hea_reg_inst.cuh

RegLaunchInfo reg_info_cached(int dev, int dtype, int nl, int lq, int mode) {
    static std::mutex mu;
    static RegLaunchInfo cache[64][2][6][6][kModes];
    static bool have[64][2][6][6][kModes];
    std::lock_guard<std::mutex> lk(mu);
    if (!have[dev][dtype][nl][lq][mode]) {
        cache[dev][dtype][nl][lq][mode] = dtype == 0 ? reg_info_f32(nl, lq, mode) : reg_info_f64(nl, lq, mode);
        have[dev][dtype][nl][lq][mode] = true;
    }
    return cache[dev][dtype][nl][lq][mode];
}

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// First qubit count served by the fp32 shared-memory tier; below it the lane-distributed register tier
// is used.  QON_SMEM_FIRST_N overrides the default for experiments (e.g. 11 = lanes up to n = 10).
// Largest batch served by the one-amplitude-per-lane latency layout (n <= 5, fp32).  Above it the
// one-thread-per-sample FFMA2 kernel has the higher throughput.  QON_LANES_MAX_B overrides (0 disables).
int64_t lanes_max_batch() {
    static const int64_t v = [] {
        const char* e = getenv("QON_LANES_MAX_B");
        // measured (Q5, 120 sublayers, fwd+grad): 69 us at B = 100, 229 us at 4096, 449 us at 8192, 890 us at
        // 16384 — where the one-thread-per-sample kernel (one ~900 us wave) draws level
        return e ? (int64_t)atoll(e) : (int64_t)12288;
    }();
    return v;
}

// Largest batch the wide latency tier (n = 6..10, hea_warp_wide.cuh) serves; above it the shared-memory tier's
// throughput wins.  QON_WIDE_MAX_B overrides every entry (0 disables).
int64_t wide_max_batch(int n, int dtype) {
    static const int64_t ov = [] { const char* e = getenv("QON_WIDE_MAX_B"); return e ? (int64_t)atoll(e) : (int64_t)-1; }();
    if (ov >= 0) return ov;
    // measured against the shared-memory tier (fwd+grad, 60 sublayers; scripts/small_batch_widths.py): n = 6: 103 us vs
    // 785 us at B = 100, level at ~8,000; n = 8: 171 vs 934 us, level at ~4,000; n = 9: 480 vs 1,033 us, level at
    // ~1,500; n = 10 (32 amplitudes per lane, 255 registers): 1,215 vs 1,110 us — not used
    static const int64_t tbl[5] = {8192, 4096, 2048, 1024, 0};      // n = 6..10
    // fp64 (vs the lane-distributed register kernels; B = 100 / 2,000): n = 6: 110 / 270 us vs 642 / 654; n = 7:
    // 155 / 447 vs 815 / 839; n = 8: 277 / 876 vs 966 / 1,041; n = 9: 533 / 2,216 vs 1,182 / 2,473
    static const int64_t tbl64[5] = {2048, 2048, 2048, 1024, 0};     // fp64, n = 6..9
    if (n < 6 || n > 10) return 0;
    return dtype == QON_F32 ? tbl[n - 6] : tbl64[n - 6];
}

// QON_HBM_TIER=generic falls back to the one-CTA-per-sample kernel for n >= 14 (experiments / A-B tests)
bool hbm_disabled() {
    static const bool v = [] {
        const char* e = getenv("QON_HBM_TIER");
        return e && strcmp(e, "generic") == 0;
    }();
    return v;
}

// Measured on B200 (profiles/r1_sweep_f32.md): the shared-memory tier beats the lane-distributed register
// layout at every n >= 6 with gradients (n = 10: 51.9 vs 24.6 TFLOP/s) and ties or wins forward-only.
int smem_first_n(bool /*grad*/) {
    static const int v = [] {
        const char* e = getenv("QON_SMEM_FIRST_N");
        const int d = e ? atoi(e) : 0;
        return d > 0 && d < kSmemMinN ? kSmemMinN : d;
    }();
    return v > 0 ? v : kSmemMinN;
}
inline bool mode_is_grad(int mode) { return mode == 1 || mode == 2 || mode == 4 || mode == 5; }
inline bool mode_is_enc(int mode) { return mode >= 3; }

struct Plan {
    int tier = -1;            // 0 register, 1 shared-memory, 2 HBM-streamed
    int nl = 0, lq = 0;
    int grid = 0, rows = 0, vp = 0, fvp = 0, S = 0;
    GenericPlan gp{};
    SmemPlan sp{};
    bool fast_smem = false;   // tier 1 served by hea_smem.cuh (fp32) instead of the generic kernel
    HbmPlan hp{};
    bool fast_hbm = false;    // tier 2 served by hea_hbm.cuh (fp32) instead of the generic kernel
    WarpPlan wp{};
    bool fast_warp = false;   // tier 0 served by hea_warp.cuh (one amplitude per lane, small batches)
    size_t off_u = 0, off_r = 0, off_h = 0, off_d = 0, off_i = 0, off_m = 0, off_state = 0, off_tc = 0, total = 0;
    int64_t rowlen = 0, mpart_len = 0;
};

int make_plan(int64_t B, int n, int K, const int* depth, int dtype, int mode, Plan* pl) {
    if (B < 0) return fail(QON_ERR_BAD_ARG, "B must be >= 0 (got %lld)", (long long)B);
    if (n < 1 || n > kMaxQubits) return fail(QON_ERR_UNSUPPORTED, "n must be in [1, %d] (got %d)", kMaxQubits, n);
    if (K < 1 || K > kMaxBlocks) return fail(QON_ERR_UNSUPPORTED, "K must be in [1, %d] (got %d)", kMaxBlocks, K);
    if (!depth) return fail(QON_ERR_BAD_ARG, "depth_per_block is NULL");
    if (dtype != QON_F32 && dtype != QON_F64) return fail(QON_ERR_BAD_ARG, "dtype must be QON_F32 or QON_F64");
    int64_t S = 0;
    for (int k = 0; k < K; ++k) {
        if (depth[k] < 1 || depth[k] > 255)
            return fail(QON_ERR_UNSUPPORTED, "depth_per_block[%d] = %d outside [1, 255]", k, depth[k]);
        S += depth[k];
    }
    int dev;
    DeviceInfo di;
    if (!device_info(&dev, &di)) return fail(QON_ERR_NO_DEVICE, "no usable CUDA device");
    const size_t es = dtype == QON_F32 ? 4 : 8;
    const bool grad = mode_is_grad(mode);
    pl->S = (int)S;
    const int max_local = dtype == QON_F32 ? 5 : 4;
    pl->fast_smem = false;
    pl->fast_hbm = false;
    pl->fast_warp = false;
    if ((!mode_is_enc(mode) || (dtype == QON_F32 && n <= 9)) && n >= 6 && n <= (dtype == QON_F32 ? 10 : 9) &&
        B <= wide_max_batch(n, dtype)) {
        pl->wp = warp_plan(n, K, (int)S, (int)es, mode);
        pl->fast_warp = pl->wp.ok;
    }
    if (pl->fast_warp) {
        // wide latency tier: one warp per sample, 2^(n-5) amplitudes per lane
        pl->tier = 0;
        pl->nl = n - 5;
        pl->lq = 5;
        const int warps = pl->wp.threads / 32;
        int64_t grid = (B + warps - 1) / warps;
        const int64_t cap = (int64_t)di.sms * pl->wp.blocks_per_sm;
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        pl->grid = (int)grid;
        pl->rows = (int)grid * warps;
        pl->vp = moment_slots(n);
        pl->fvp = freq_slots(n);
    } else if (dtype == QON_F32 && !mode_is_enc(mode) && n >= smem_first_n(grad) && n <= kSmemMaxN) {
        // fp32 shared-memory tier: register-blocked FFMA2 passes over a state held in shared memory
        pl->sp = smem_plan(n, mode);
        if (!pl->sp.ok) return fail(QON_ERR_UNSUPPORTED, "shared-memory tier kernel does not fit for n=%d", n);
        pl->fast_smem = true;
        pl->tier = 1;
        const int64_t rounds = (B + pl->sp.geo.spc - 1) / pl->sp.geo.spc;
        int64_t grid = rounds;
        const int64_t cap = (int64_t)di.sms * pl->sp.blocks_per_sm;
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        pl->grid = (int)grid;
        pl->rows = (int)grid * (pl->sp.threads / 32);
        pl->vp = pl->sp.geo.vp;
        pl->fvp = 0;
    } else if (dtype == QON_F32 && !mode_is_enc(mode) && n >= kHbmMinN && n <= kHbmMaxN && !hbm_disabled()) {
        // fp32 HBM-streamed tier: 13-qubit tiles through shared memory, two passes over HBM per sublayer
        pl->hp = hbm_plan(B, n, K, mode);
        if (!pl->hp.ok) return fail(QON_ERR_UNSUPPORTED, "HBM-streamed tier kernel does not fit for n=%d", n);
        pl->fast_hbm = true;
        pl->tier = 2;
        pl->grid = pl->hp.grid_rev;
        pl->rows = pl->hp.rows;
        pl->vp = (3 * n + 3) / 4 * 4;
        pl->fvp = 0;
    } else if (n <= max_local + 5) {
        pl->tier = 0;
        pl->nl = n <= max_local ? n : max_local;
        pl->lq = n - pl->nl;
        // small batches (n <= 5): one amplitude per lane, 2^n lanes per sample — the latency tier
        if (n <= 5 && B <= lanes_max_batch()) {
            pl->wp = warp_plan(n, K, (int)S, (int)es, mode);
            pl->fast_warp = pl->wp.ok;
        }
        int warps, blocks_per_sm;
        if (pl->fast_warp) {
            pl->nl = 0;
            pl->lq = n;
            warps = pl->wp.threads / 32;
            blocks_per_sm = pl->wp.blocks_per_sm;
            // (capping the resident CTAs to shrink the partial-row table was measured: slower at every batch size)
        } else {
            RegLaunchInfo ri = reg_info_cached(dev, dtype, pl->nl, pl->lq, mode);
            if (!ri.ok)
                return fail(QON_ERR_UNSUPPORTED, "register-tier kernel (n=%d, lanes 2^%d, mode %d) is not built%s", n,
                            pl->lq, mode, mode_is_enc(mode) ? " (fused encoding needs n <= 5 in fp32, n <= 4 in fp64)" : "");
            warps = ri.threads / 32;
            blocks_per_sm = ri.blocks_per_sm;
        }
        const int64_t spw = 32 >> pl->lq;
        const int64_t tiles = (B + spw - 1) / spw;
        int64_t grid = (tiles + warps - 1) / warps;
        const int64_t cap = (int64_t)di.sms * blocks_per_sm;
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        pl->grid = (int)grid;
        pl->rows = (int)grid * warps;
        pl->vp = moment_slots(n);
        pl->fvp = freq_slots(n);
    } else {
        if (mode_is_enc(mode))
            return fail(QON_ERR_UNSUPPORTED, "fused encoding is built for the register tier only (n=%d)", n);
        pl->gp = generic_plan(n, dtype, mode);
        if (!pl->gp.state_global && pl->gp.blocks_per_sm < 1)
            return fail(QON_ERR_UNSUPPORTED, "generic kernel does not fit for n=%d", n);
        pl->tier = pl->gp.state_global ? 2 : 1;
        int64_t cap = (int64_t)di.sms * (pl->gp.blocks_per_sm > 0 ? pl->gp.blocks_per_sm : 1);
        int64_t grid = B < cap ? B : cap;
        if (grid < 1) grid = 1;
        pl->grid = (int)grid;
        pl->rows = (int)grid;
        pl->vp = (3 * n + 3) / 4 * 4;
        pl->fvp = 0;
    }
    if (dtype == QON_F32 && n == 5) {
        // the tensor-core tier (hea_tc2.cuh, hea_tc3.cuh) writes one partial row per compute warp: 8 warps x min(SMs, tiles) CTAs
        int64_t tcg = (B + 127) / 128;          // CTAs of the gradient kernels: one per SM once there are that many tiles
        if (tcg > di.sms) tcg = di.sms;
        if (tcg < 1) tcg = 1;
        if (pl->rows < (int)tcg * 8) pl->rows = (int)tcg * 8;
    }
    size_t off = 0;
    pl->off_u = off; off = align_up(off + (size_t)S * n * 4 * es);
    pl->off_r = off; off = align_up(off + (size_t)S * n * 4 * es);
    pl->off_h = off; off = align_up(off + ((size_t)1 << n) * es);
    pl->off_d = off; off = align_up(off + (size_t)K * sizeof(int));
    pl->off_i = off; off = align_up(off + (size_t)K * n * sizeof(int));
    pl->off_m = off;
    pl->rowlen = (int64_t)S * pl->vp + (int64_t)K * pl->fvp + 4;
    pl->mpart_len = grad ? (int64_t)pl->rows * pl->rowlen : 0;
    off = align_up(off + (size_t)pl->mpart_len * es);
    pl->off_state = off;
    if (pl->fast_hbm) off = align_up(off + pl->hp.bytes);
    else if (pl->tier == 2) off = align_up(off + (size_t)pl->grid * (grad ? 4 : 2) * ((size_t)1 << n) * es);
    pl->off_tc = off;
    if (dtype == QON_F32 && n == 5) off = align_up(off + tc_workspace_bytes(K, (int)S, grad ? B : 0, di.sms));   // tensor-core tier: operand images
    pl->total = off;
    return 0;
}

struct Job {
    // circuit
    const void* w = nullptr; void* out = nullptr;
    int64_t B = 0; int n = 0, K = 0; const int* depth = nullptr;
    const void* ham_diag = nullptr; int diag_order = 0; double offset = 0, coeff = 0; int ham_kind = 0; int dtype = 0;
    void* ws = nullptr; size_t ws_bytes = 0; void* stream = nullptr;
    // angles given
    const void* x = nullptr; int64_t ldx = 0;
    // fused encoding
    bool enc = false;
    const void *u0 = nullptr, *u1 = nullptr; int64_t ldu0 = 0, ldu1 = 0; int in0 = 0, in1 = 0, K0 = 0;
    const void *fw = nullptr, *fb = nullptr;
    void *grad_fw = nullptr, *grad_fb = nullptr, *sums = nullptr;
    // gradients
    bool grad = false;
    const void *grad_out = nullptr, *target = nullptr, *bias = nullptr; double gscale = 0; void* gbuf = nullptr;
    void* grad_x = nullptr; int64_t ldgx = 0; void* grad_w = nullptr;
    // data-parallel exchange fused into finalize (fp32, fused-encoding training step)
    float* flat = nullptr; FlatLayout fl{}; PeerPtrs peers{}; int world = 0, rank = 0; int64_t peer_max_len = 0;
};

int job_mode(const Job& j) {
    if (j.enc) return !j.grad ? 3 : ((j.grad_fw || j.grad_fb) ? 5 : 4);
    return !j.grad ? 0 : (j.grad_x ? 1 : 2);
}

template <typename T>
int run(const Job& j) {
    const int mode = job_mode(j);
    const int n = j.n, K = j.K;
    Plan pl;
    if (int rc = make_plan(j.B, n, K, j.depth, j.dtype, mode, &pl)) return rc;
    if (!j.w) return fail(QON_ERR_BAD_ARG, "w must be non-NULL");
    if (!j.enc) {
        if (!j.x || !j.out) return fail(QON_ERR_BAD_ARG, "x and out must be non-NULL");
        if (j.ldx < (int64_t)n * K) return fail(QON_ERR_BAD_ARG, "ldx (%lld) < n*K (%d)", (long long)j.ldx, n * K);
    } else {
        if (j.K0 < 0 || j.K0 > K) return fail(QON_ERR_BAD_ARG, "K0 (%d) outside [0, K=%d]", j.K0, K);
        if (!j.fw) return fail(QON_ERR_BAD_ARG, "fw must be non-NULL");
        if (j.K0 > 0 && (!j.u0 || j.in0 < 1 || j.ldu0 < j.in0)) return fail(QON_ERR_BAD_ARG, "bad source 0 (u0/in0/ldu0)");
        if (j.K0 < K && (!j.u1 || j.in1 < 1 || j.ldu1 < j.in1)) return fail(QON_ERR_BAD_ARG, "bad source 1 (u1/in1/ldu1)");
        if (!j.grad && !j.out) return fail(QON_ERR_BAD_ARG, "out must be non-NULL");
    }
    if (j.grad && !j.grad_w && !j.flat) return fail(QON_ERR_BAD_ARG, "grad_w must be non-NULL");
    if (j.grad && !j.grad_out && !j.target) return fail(QON_ERR_BAD_ARG, "grad_out (or target) must be non-NULL");
    if (j.grad_x && j.ldgx < (int64_t)n * K)
        return fail(QON_ERR_BAD_ARG, "ldgx (%lld) < n*K (%d)", (long long)j.ldgx, n * K);
    if (j.ham_kind < QON_HAM_DIAG || j.ham_kind > QON_HAM_PAULI_Y) return fail(QON_ERR_BAD_ARG, "bad ham_kind %d", j.ham_kind);
    if (j.ham_diag && j.ham_kind != QON_HAM_DIAG) return fail(QON_ERR_BAD_ARG, "ham_diag given with a Pauli-X/Y observable");
    if (j.diag_order != QON_DIAG_LSB0 && j.diag_order != QON_DIAG_MSB0) return fail(QON_ERR_BAD_ARG, "bad diag_order %d", j.diag_order);
    if (!j.ws || j.ws_bytes < pl.total)
        return fail(QON_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", pl.total, j.ws_bytes);
    if ((uintptr_t)j.ws % 256) return fail(QON_ERR_BAD_ARG, "workspace must be 256-byte aligned");
    const uintptr_t am = sizeof(T) - 1;
    if (((uintptr_t)j.x & am) || ((uintptr_t)j.w & am) || ((uintptr_t)j.out & am) || ((uintptr_t)j.u0 & am) ||
        ((uintptr_t)j.u1 & am))
        return fail(QON_ERR_BAD_ARG, "device pointers must be aligned to the element size");
    cudaStream_t st = (cudaStream_t)j.stream;
    char* base = (char*)j.ws;
    HeaParams<T> p{};
    p.x = (const T*)j.x; p.ldx = j.ldx; p.B = j.B; p.out = (T*)j.out;
    p.gout = (const T*)j.grad_out; p.gx = (T*)j.grad_x; p.ldgx = j.ldgx;
    p.target = (const T*)j.target; p.bias = (const T*)j.bias; p.gbuf = (T*)j.gbuf; p.gscale = (T)j.gscale;
    p.ucoef = (const Vec4<T>*)(base + pl.off_u);
    p.rcoef = (const Vec4<T>*)(base + pl.off_r);
    p.hdiag = (const T*)(base + pl.off_h);
    p.depth = (const int*)(base + pl.off_d);
    p.mpart = (T*)(base + pl.off_m);
    p.rowlen = pl.rowlen;
    p.K = K; p.S = pl.S; p.pauli = j.ham_kind; p.offset = (T)j.offset; p.coeff = (T)j.coeff;
    p.u0 = (const T*)j.u0; p.u1 = (const T*)j.u1; p.ldu0 = j.ldu0; p.ldu1 = j.ldu1; p.K0 = j.K0;
    p.in0 = j.in0 > 0 ? j.in0 : 1; p.in1 = j.in1 > 0 ? j.in1 : 1;
    p.uidx = j.enc ? (const int*)(base + pl.off_i) : nullptr;
    p.fw = (const T*)j.fw; p.fb = (const T*)j.fb;

    DepthPack dp;
    memset(&dp, 0, sizeof dp);
    for (int k = 0; k < K; ++k) dp.d[k] = (unsigned char)j.depth[k];
    {
        const int64_t work = pl.mpart_len > ((int64_t)1 << n) ? pl.mpart_len : ((int64_t)1 << n);
        int64_t blocks = (work + 255) / 256;
        if (blocks > 1184) blocks = 1184;
        if (blocks < 1) blocks = 1;
        prep_kernel<T><<<(int)blocks, 256, 0, st>>>(
            (const T*)j.w, n, pl.S, K, dp, (Vec4<T>*)(base + pl.off_u), (Vec4<T>*)(base + pl.off_r),
            (int*)(base + pl.off_d), (T*)(base + pl.off_h), (const T*)j.ham_diag, j.diag_order, j.offset, j.coeff,
            p.mpart, pl.mpart_len, j.enc ? (int*)(base + pl.off_i) : nullptr, j.K0, j.in0 > 0 ? j.in0 : 1,
            j.in1 > 0 ? j.in1 : 1);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail((int)e, "prep launch failed: %s", cudaGetErrorString(e));
    }
    // partial rows the finalize kernels have to sum: every row of the plan, unless the tensor-core tier ran — it fills
    // one row per compute warp, and with the GEMM-form weight gradients (hea_tc3.cuh) the moments sit in row 0 alone
    int rows_mom = j.B > 0 ? pl.rows : 0, rows_enc = rows_mom;
    bool use_tc = false;
    if (j.B > 0) {
        cudaError_t e;
        const TcConfig& tcc = tc_config();
        if constexpr (sizeof(T) == 4)
            use_tc = tcc.enable && n == 5 && j.ham_kind == QON_HAM_DIAG && j.B >= tcc.min_batch;
        if (use_tc) {
            if constexpr (sizeof(T) == 4) {
                int dev; DeviceInfo di;
                if (!device_info(&dev, &di)) return fail(QON_ERR_NO_DEVICE, "no usable CUDA device");
                const int version = tcc.enable == 3 ? 3 : (tcc.enable == 2 || !tc_outer_supported(K)) ? 2 : 4;
                e = tc_launch(mode, version, di.sms, (const HeaParams<float>&)p, (const float*)j.w, dp,
                              base + pl.off_tc, tcc.dbg, tcc.err, st);
                if (j.grad) {
                    int64_t tcg = (j.B + 127) / 128;
                    if (tcg > di.sms) tcg = di.sms;
                    rows_enc = (int)tcg * 8 < pl.rows ? (int)tcg * 8 : pl.rows;
                    rows_mom = version == 4 ? 1 : rows_enc;
                }
            } else e = cudaErrorInvalidValue;
        } else if (pl.fast_warp) {
            if constexpr (sizeof(T) == 4) e = warp_launch_f32(n, mode, pl.grid, pl.wp, (const HeaParams<float>&)p, dp, st);
            else e = warp_launch_f64(n, mode, pl.grid, pl.wp, (const HeaParams<double>&)p, dp, st);
        } else if (pl.tier == 0) {
            if constexpr (sizeof(T) == 4) e = reg_launch_f32(pl.nl, pl.lq, mode, pl.grid, (const HeaParams<float>&)p, st);
            else e = reg_launch_f64(pl.nl, pl.lq, mode, pl.grid, (const HeaParams<double>&)p, st);
        } else if (pl.fast_smem) {
            if constexpr (sizeof(T) == 4) e = smem_launch(mode, pl.grid, pl.sp, (const HeaParams<float>&)p, st);
            else e = cudaErrorInvalidValue;
        } else if (pl.fast_hbm) {
            if constexpr (sizeof(T) == 4)
                e = hbm_run((const HeaParams<float>&)p, j.depth, n, K, mode, pl.hp, base + pl.off_state, st);
            else e = cudaErrorInvalidValue;
        } else {
            T* gstate = pl.tier == 2 ? (T*)(base + pl.off_state) : nullptr;
            if constexpr (sizeof(T) == 4)
                e = generic_launch_f32(n, mode, pl.grid, pl.gp, (const HeaParams<float>&)p, pl.vp, (float*)gstate, st);
            else
                e = generic_launch_f64(n, mode, pl.grid, pl.gp, (const HeaParams<double>&)p, pl.vp, (double*)gstate, st);
        }
        if (e != cudaSuccess) return fail((int)e, "kernel launch failed: %s", cudaGetErrorString(e));
    }
    if (j.grad && j.flat) {
        if constexpr (sizeof(T) == 4) {
            finalize_exchange_kernel<<<pl.S + K + 1, 256, 0, st>>>(
                (const float*)p.mpart, rows_mom, rows_enc, pl.rowlen, pl.vp, pl.fvp, n, pl.S, K, (const float*)j.w, j.flat, j.fl, j.peers,
                j.world, j.rank, j.peer_max_len, kPeerTimeoutCycles);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return fail((int)e, "finalize+exchange launch failed: %s", cudaGetErrorString(e));
        }
        return 0;
    }
    if (j.grad) {
        int threads = 256;
        while (threads < pl.vp) threads <<= 1;
        const int RG = threads / pl.vp;
        finalize_kernel<T><<<pl.S, threads, (size_t)RG * pl.vp * sizeof(double), st>>>(
            p.mpart, rows_mom, pl.rowlen, pl.vp, n, (const T*)j.w, (T*)j.grad_w);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail((int)e, "finalize launch failed: %s", cudaGetErrorString(e));
        if (j.enc && (j.grad_fw || j.grad_fb || j.sums)) {
            finalize_enc_kernel<T><<<K + 1, 256, 0, st>>>(p.mpart, rows_enc, pl.rowlen, (int64_t)pl.S * pl.vp, pl.fvp, n, K,
                                                          (T*)j.grad_fw, (T*)j.grad_fb, (T*)j.sums);
            e = cudaGetLastError();
            if (e != cudaSuccess) return fail((int)e, "finalize_enc launch failed: %s", cudaGetErrorString(e));
        }
    }
    return 0;
}

int dispatch(const Job& j) {
    if (j.dtype == QON_F32) return run<float>(j);
    if (j.dtype == QON_F64) return run<double>(j);
    return fail(QON_ERR_BAD_ARG, "dtype must be QON_F32 or QON_F64");
}

void fill_common(Job& j, const void* w, void* out, int64_t B, int n, int K, const int* depth, const void* ham_diag,
                 int diag_order, double off, double coeff, int ham_kind, int dtype, void* ws, size_t ws_bytes,
                 void* stream) {
    j.w = w; j.out = out; j.B = B; j.n = n; j.K = K; j.depth = depth; j.ham_diag = ham_diag; j.diag_order = diag_order;
    j.offset = off; j.coeff = coeff; j.ham_kind = ham_kind; j.dtype = dtype; j.ws = ws; j.ws_bytes = ws_bytes;
    j.stream = stream;
}

}  // namespace
}  // namespace qon

using namespace qon;

extern "C" {

int qon_abi_version(void) { return QON_ABI_VERSION; }

const char* qon_last_error(void) { return g_err.c_str(); }

size_t qon_workspace_bytes(int64_t B, int n, int K, const int* depth_per_block, int dtype, int need_grad) {
    // the variants of one entry-point family are different kernels (different occupancy -> different number
    // of partial rows); report the maximum so one workspace serves them all
    const int fwd_modes[] = {0, 3}, grad_modes[] = {1, 2, 4, 5};
    const int* modes = need_grad ? grad_modes : fwd_modes;
    const int count = need_grad ? 4 : 2;
    size_t total = 0;
    for (int i = 0; i < count; ++i) {
        Plan pl;
        const int rc = make_plan(B, n, K, depth_per_block, dtype, modes[i], &pl);
        if (rc == 0) { if (pl.total > total) total = pl.total; }
        else if (modes[i] < 3) return 0;          // the plain variants must plan; fused encoding is optional
    }
    return total;
}

int64_t qon_latency_tier_max_batch(void) { return lanes_max_batch(); }

int qon_encoded_supported(int64_t B, int n, int dtype, int need_grad) {
    Plan pl;
    int one = 1;
    const int rc = make_plan(B, n, 1, &one, dtype, need_grad ? 5 : 3, &pl);
    return rc == 0 ? 1 : 0;
}

int qon_encoded_supported_for(int64_t B, int n, int K, const int* depth_per_block, int dtype, int need_grad) {
    // planned with the REAL circuit: the wide latency tier's shared-memory footprint grows with K and S, so a
    // deep n = 6..9 circuit can fit the one-block probe of qon_encoded_supported() and still have no kernel
    if (K < 1 || !depth_per_block) return 0;
    Plan pl;
    const int rc = make_plan(B, n, K, depth_per_block, dtype, need_grad ? 5 : 3, &pl);
    return rc == 0 ? 1 : 0;
}

size_t qon_peer_buffer_bytes(int64_t max_len, int world) {
    if (max_len < 1 || world < 1 || world > kPeerMaxWorld) return 0;
    return peer_buffer_bytes(max_len, world);
}

int qon_peer_allreduce_f32(const float* src, float* dst, int64_t len, void* const* peer_bufs, int world, int rank,
                           int64_t max_len, void* stream) {
    if (world < 1 || world > kPeerMaxWorld) return fail(QON_ERR_UNSUPPORTED, "world must be in [1, %d] (got %d)", kPeerMaxWorld, world);
    if (rank < 0 || rank >= world) return fail(QON_ERR_BAD_ARG, "rank %d outside [0, %d)", rank, world);
    if (len < 0 || len > max_len || max_len > (1 << 20)) return fail(QON_ERR_BAD_ARG, "len (%lld) must be in [0, max_len = %lld <= 2^20]", (long long)len, (long long)max_len);
    if (!src || !dst || !peer_bufs) return fail(QON_ERR_BAD_ARG, "src, dst and peer_bufs must be non-NULL");
    PeerPtrs pp{};
    for (int p = 0; p < world; ++p) {
        if (!peer_bufs[p] || ((uintptr_t)peer_bufs[p] & 255)) return fail(QON_ERR_BAD_ARG, "peer buffer %d is NULL or not 256-byte aligned", p);
        pp.p[p] = (char*)peer_bufs[p];
    }
    const long long timeout = kPeerTimeoutCycles;   // (querying cudaDevAttrClockRate costs milliseconds per call)
    peer_allreduce_kernel<<<1, kPeerThreads, 0, (cudaStream_t)stream>>>(src, dst, (int)len, pp, world, rank, max_len, timeout);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "peer all-reduce launch failed: %s", cudaGetErrorString(e));
    return 0;
}

int qon_plan_tier(int64_t B, int n, int dtype, int need_grad, int* lanes_log2) {
    int one = 1;
    Plan pl;
    if (make_plan(B, n, 1, &one, dtype, need_grad ? 1 : 0, &pl)) return -1;
    if (lanes_log2) *lanes_log2 = pl.tier == 0 ? pl.lq : -1;
    return pl.tier;
}

int qon_hea_forward(const void* x, int64_t ldx, const void* w, void* out, int64_t B, int n, int K,
                    const int* depth_per_block, const void* ham_diag, int diag_order, double ham_offset,
                    double ham_coeff, int ham_kind, int dtype, void* workspace, size_t workspace_bytes, void* stream) {
    Job j;
    fill_common(j, w, out, B, n, K, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff, ham_kind, dtype,
                workspace, workspace_bytes, stream);
    j.x = x; j.ldx = ldx;
    return dispatch(j);
}

int qon_hea_forward_backward(const void* x, int64_t ldx, const void* w, const void* grad_out, void* out, void* grad_x,
                             int64_t ldgx, void* grad_w, int64_t B, int n, int K, const int* depth_per_block,
                             const void* ham_diag, int diag_order, double ham_offset, double ham_coeff, int ham_kind,
                             int dtype, void* workspace, size_t workspace_bytes, void* stream) {
    if (!grad_out) return fail(QON_ERR_BAD_ARG, "grad_out must be non-NULL");
    Job j;
    fill_common(j, w, out, B, n, K, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff, ham_kind, dtype,
                workspace, workspace_bytes, stream);
    j.x = x; j.ldx = ldx; j.grad = true; j.grad_out = grad_out; j.grad_x = grad_x; j.ldgx = ldgx; j.grad_w = grad_w;
    return dispatch(j);
}

int qon_hea_mse_forward_backward(const void* x, int64_t ldx, const void* w, const void* target, const void* bias,
                                 double grad_scale, void* out, void* grad_out_written, void* grad_x, int64_t ldgx,
                                 void* grad_w, int64_t B, int n, int K, const int* depth_per_block,
                                 const void* ham_diag, int diag_order, double ham_offset, double ham_coeff, int ham_kind,
                                 int dtype, void* workspace, size_t workspace_bytes, void* stream) {
    if (!target) return fail(QON_ERR_BAD_ARG, "target must be non-NULL");
    if (!grad_out_written) return fail(QON_ERR_BAD_ARG, "grad_out_written must be non-NULL");
    Job j;
    fill_common(j, w, out, B, n, K, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff, ham_kind, dtype,
                workspace, workspace_bytes, stream);
    j.x = x; j.ldx = ldx; j.grad = true; j.target = target; j.bias = bias; j.gscale = grad_scale;
    j.gbuf = grad_out_written; j.grad_x = grad_x; j.ldgx = ldgx; j.grad_w = grad_w;
    return dispatch(j);
}

int qon_encoded_forward(const void* u0, int64_t ldu0, int in0, int K0, const void* u1, int64_t ldu1, int in1,
                        const void* fw, const void* fb, const void* w, void* out, int64_t B, int n, int K,
                        const int* depth_per_block, const void* ham_diag, int diag_order, double ham_offset,
                        double ham_coeff, int ham_kind, int dtype, void* workspace, size_t workspace_bytes,
                        void* stream) {
    Job j;
    fill_common(j, w, out, B, n, K, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff, ham_kind, dtype,
                workspace, workspace_bytes, stream);
    j.enc = true; j.u0 = u0; j.ldu0 = ldu0; j.in0 = in0; j.K0 = K0; j.u1 = u1; j.ldu1 = ldu1; j.in1 = in1;
    j.fw = fw; j.fb = fb;
    return dispatch(j);
}

int qon_encoded_mse_step(const void* u0, int64_t ldu0, int in0, int K0, const void* u1, int64_t ldu1, int in1,
                         const void* fw, const void* fb, const void* w, const void* target, const void* bias,
                         double grad_scale, void* out, void* grad_w, void* grad_fw, void* grad_fb, void* sums,
                         int64_t B, int n, int K, const int* depth_per_block, const void* ham_diag, int diag_order,
                         double ham_offset, double ham_coeff, int ham_kind, int dtype, void* workspace,
                         size_t workspace_bytes, void* stream) {
    if (!target) return fail(QON_ERR_BAD_ARG, "target must be non-NULL");
    if (!sums) return fail(QON_ERR_BAD_ARG, "sums must be non-NULL");
    Job j;
    fill_common(j, w, out, B, n, K, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff, ham_kind, dtype,
                workspace, workspace_bytes, stream);
    j.enc = true; j.u0 = u0; j.ldu0 = ldu0; j.in0 = in0; j.K0 = K0; j.u1 = u1; j.ldu1 = ldu1; j.in1 = in1;
    j.fw = fw; j.fb = fb; j.grad = true; j.target = target; j.bias = bias; j.gscale = grad_scale;
    j.grad_w = grad_w; j.grad_fw = grad_fw; j.grad_fb = grad_fb; j.sums = sums;
    return dispatch(j);
}

int qon_encoded_mse_step_dp(const void* u0, int64_t ldu0, int in0, int K0, const void* u1, int64_t ldu1, int in1,
                            const void* fw, const void* fb, const void* w, const void* target, const void* bias,
                            double grad_scale, void* out, float* flat, int64_t flat_len, int64_t w_off, int64_t fw_off,
                            int64_t fb_off, int64_t sums_off, void* const* peer_bufs, int world, int rank,
                            int64_t max_len, int64_t B, int n, int K, const int* depth_per_block, const void* ham_diag,
                            int diag_order, double ham_offset, double ham_coeff, int ham_kind, void* workspace,
                            size_t workspace_bytes, void* stream) {
    if (!target) return fail(QON_ERR_BAD_ARG, "target must be non-NULL");
    if (!flat || !peer_bufs) return fail(QON_ERR_BAD_ARG, "flat and peer_bufs must be non-NULL");
    if (world < 1 || world > kPeerMaxWorld) return fail(QON_ERR_UNSUPPORTED, "world must be in [1, %d] (got %d)", kPeerMaxWorld, world);
    if (rank < 0 || rank >= world) return fail(QON_ERR_BAD_ARG, "rank %d outside [0, %d)", rank, world);
    if (flat_len < 1 || flat_len > max_len) return fail(QON_ERR_BAD_ARG, "flat_len (%lld) must be in [1, max_len = %lld]", (long long)flat_len, (long long)max_len);
    if (n < 1 || K < 1 || !depth_per_block) return fail(QON_ERR_BAD_ARG, "bad circuit description");
    int64_t S = 0;
    for (int k = 0; k < K; ++k) S += depth_per_block[k];
    const int64_t nw = 3 * (int64_t)n * S, ne = (int64_t)n * K;
    const bool freq = fw_off >= 0 || fb_off >= 0;
    if (w_off < 0 || w_off + nw > flat_len || sums_off < 0 || sums_off + 2 > flat_len ||
        (freq && (fw_off < 0 || fb_off < 0 || fw_off + ne > flat_len || fb_off + ne > flat_len)))
        return fail(QON_ERR_BAD_ARG, "gradient offsets fall outside the flat buffer");
    Job j;
    fill_common(j, w, out, B, n, K, depth_per_block, ham_diag, diag_order, ham_offset, ham_coeff, ham_kind, QON_F32,
                workspace, workspace_bytes, stream);
    j.enc = true; j.u0 = u0; j.ldu0 = ldu0; j.in0 = in0; j.K0 = K0; j.u1 = u1; j.ldu1 = ldu1; j.in1 = in1;
    j.fw = fw; j.fb = fb; j.grad = true; j.target = target; j.bias = bias; j.gscale = grad_scale;
    // grad_fw / grad_fb only select the kernel variant (frequency gradients reduced in-kernel); nothing is written
    // through them: the fused finalize pushes every gradient into the peers' slots and then into `flat`
    j.grad_fw = freq ? (void*)(flat + fw_off) : nullptr;
    j.grad_fb = freq ? (void*)(flat + fb_off) : nullptr;
    j.flat = flat;
    j.fl = FlatLayout{w_off, freq ? fw_off : -1, freq ? fb_off : -1, sums_off, flat_len};
    for (int p = 0; p < world; ++p) {
        if (!peer_bufs[p] || ((uintptr_t)peer_bufs[p] & 255)) return fail(QON_ERR_BAD_ARG, "peer buffer %d is NULL or not 256-byte aligned", p);
        j.peers.p[p] = (char*)peer_bufs[p];
    }
    j.world = world; j.rank = rank; j.peer_max_len = max_len;
    return dispatch(j);
}

int qon_tensor_tier(int enable, int64_t min_batch, void* debug_state, void* error_flag) {
    TcConfig& c = tc_config();
    const int prev = c.enable;
    if (enable >= 0) c.enable = enable;
    if (min_batch >= 0) c.min_batch = min_batch;
    c.dbg = (float*)debug_state;
    c.err = (int*)error_flag;
    return prev;
}

double qon_measure_fp32_peak_tflops(int iters, void* stream) {
    int dev;
    DeviceInfo di;
    if (!device_info(&dev, &di)) { fail(QON_ERR_NO_DEVICE, "no usable CUDA device"); return -1.0; }
    if (iters < 1) iters = 1;
    cudaStream_t st = (cudaStream_t)stream;
    float* dummy = nullptr;
    if (cudaMalloc(&dummy, 4) != cudaSuccess) { fail(QON_ERR_NO_DEVICE, "cudaMalloc failed"); return -1.0; }
    const int grid = di.sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    ffma_peak_kernel<<<grid, threads, 0, st>>>(dummy, iters / 4 + 1, 0.999f, 1e-4f);   // warm-up
    cudaEventRecord(e0, st);
    ffma_peak_kernel<<<grid, threads, 0, st>>>(dummy, iters, 0.999f, 1e-4f);
    cudaEventRecord(e1, st);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(dummy);
    if (e != cudaSuccess || ms <= 0.f) { fail((int)e, "peak probe failed: %s", cudaGetErrorString(e)); return -1.0; }
    const double flops = 2.0 * 256.0 * (double)iters * (double)grid * threads;
    return flops / (ms * 1e-3) / 1e12;
}

}  // extern "C"
