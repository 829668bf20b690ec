// One-shot sum all-reduce of a small fp32 vector over NVLink peer memory: the exchange step of the
// data-parallel training step (SURVEY §8e — 2,402 floats per step at Q5 Net40-2-20-2).
//
// NCCL needs ~25-40 us for a 10 KB all-reduce (launch + protocol); at the reference's batch size the whole
// training step is ~90 us, so the collective is written here as ONE single-CTA kernel per rank over
// symmetric peer-mapped buffers (allocated and exchanged by torch's symmetric memory — plumbing only):
//   1. push   : every rank stores its vector into slot [set][rank] of EVERY peer's buffer (coalesced P2P stores
//               through NVSwitch; the local copy is an ordinary store);
//   2. signal : __threadfence_system, CTA barrier, then lane p release-stores the epoch into peer p's flag
//               word [rank];
//   3. wait   : lane p acquire-spins on the local flag word [p] until it carries this epoch (bounded, ~30 s:
//               on timeout the result is NaN-poisoned and an error word is set — never a hang);
//   4. reduce : out[i] = sum_p slot[set][p][i] in rank order — every rank adds the same numbers in the same
//               order, so replicas stay bit-identical.
// Slots are double-buffered by epoch parity: a rank can start epoch e+2 (same set as e) only after every peer
// has signalled e+1, i.e. after every peer's epoch-e kernel has retired, so no slot is overwritten while
// it is still being read.  The epoch counter lives in device memory (local, not shared), which keeps the
// call CUDA-graph-replayable.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qon {

constexpr int kPeerMaxWorld = 8;
constexpr int kPeerHeaderBytes = 256;    // [0,128): flag words [world]; 128: epoch counter; 132: error word
constexpr int kPeerThreads = 1024;
// how long a rank waits for its peers before giving up (~30 s of SM clocks): long enough for a rank that is busy
// on the host (checkpoint I/O, evaluation) between steps, short enough that a dead peer never hangs the GPU
constexpr long long kPeerTimeoutCycles = 60000000000LL;

struct PeerPtrs { char* p[kPeerMaxWorld]; };

__host__ __device__ inline size_t peer_buffer_bytes(int64_t max_len, int world) {
    return (size_t)kPeerHeaderBytes + (size_t)2 * world * (size_t)max_len * sizeof(float);
}

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(const float* __restrict__ src, float* dst, int len,
                                                                      PeerPtrs pp, int world, int rank, int64_t max_len,
                                                                      long long timeout_cycles) {
    __shared__ unsigned s_epoch;
    __shared__ int s_timeout;
    const int t = threadIdx.x;
    char* mine = pp.p[rank];
    unsigned* my_flags = reinterpret_cast<unsigned*>(mine);
    unsigned* epoch_ctr = reinterpret_cast<unsigned*>(mine + 128);
    unsigned* err_word = reinterpret_cast<unsigned*>(mine + 132);
    if (t == 0) { s_epoch = *epoch_ctr + 1u; s_timeout = 0; }
    __syncthreads();
    const unsigned e = s_epoch;
    const size_t set_off = (size_t)(e & 1u) * world * (size_t)max_len;

    // 1. push (src is fully read here, before any write to dst: in-place calls are fine)
    for (int p = 0; p < world; ++p) {
        float* slot = reinterpret_cast<float*>(pp.p[p] + kPeerHeaderBytes) + set_off + (size_t)rank * max_len;
        for (int i = t; i < len; i += kPeerThreads) slot[i] = src[i];
    }
    // 2. signal
    __threadfence_system();
    __syncthreads();
    if (t < world) st_release_sys(reinterpret_cast<unsigned*>(pp.p[t]) + rank, e);
    // 3. wait
    if (t < world) {
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(my_flags + t) - e) < 0) {
            if (clock64() - t0 > timeout_cycles) { s_timeout = 1; break; }
        }
    }
    __syncthreads();
    const bool bad = s_timeout != 0;
    // 4. reduce in rank order
    const float* slots = reinterpret_cast<const float*>(mine + kPeerHeaderBytes) + set_off;
    for (int i = t; i < len; i += kPeerThreads) {
        float acc = 0.f;
        for (int p = 0; p < world; ++p) acc += __ldcg(slots + (size_t)p * max_len + i);
        dst[i] = bad ? __int_as_float(0x7fc00000) : acc;
    }
    if (t == 0) {
        *epoch_ctr = e;
        if (bad) *err_word = e;
    }
}

}  // namespace qon
