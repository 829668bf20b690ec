/* quanonet_b200 — C-ABI of the B200-native HEA statevector simulator.
 *
 * This is the drop-in boundary for the ONE hot path of Wang-Ruocheng/QuanONet: the batched
 * statevector evaluation (+ adjoint gradients) of the hardware-efficient-ansatz circuit that the
 * reference runs through TorchQuantum in core/quantum_circuits_tq.py (_TQHEACircuit.forward :65-104,
 * ._measure :106-127; backward = loss.backward() at solvers/solver_pt.py:235).  The reference is pure
 * Python and has no FFI of its own; these entry points are what a binding for that path would call
 * (see INTEGRATION.md for the ctypes stub and the torch custom op built on it).
 *
 * Circuit in canonical form (host code canonicalises the reference's block_configs into it):
 *   K blocks; block k = RX(x[b, k*n + q]) on every qubit q, then depth_per_block[k] >= 1 sublayers;
 *   sublayer s = RY(w[s,2,q]) RZ(w[s,1,q]) RY(w[s,0,q]) on every qubit q (RY(w[s,0,q]) first), then the
 *   CNOT ring control=(i+1)%n -> target=i, i = 0..n-1 (no ring when n == 1);  S = sum_k depth[k].
 *   out[b] = <psi_b| H |psi_b>.   Qubit q is bit q of the amplitude index (qubit 0 = LSB).
 *
 * Ownership: the caller allocates every buffer (inputs, outputs, workspace); the library never
 *   allocates, frees or keeps device pointers past return.
 * Errors: 0 = success; negative = invalid argument (checked on the host before any launch);
 *   positive = cudaError_t of a failed launch.  Never throws, never exits.  qon_last_error() returns
 *   a thread-local message for the last non-zero return on this thread.
 * Threading / streams: stateless and re-entrant.  All work is enqueued on `stream` (a cudaStream_t
 *   passed as void*; NULL = legacy default stream) of the CURRENT device; no host synchronisation,
 *   CUDA-graph capturable.  Two calls may share a workspace only if stream-ordered.
 * Layout: x/gx rows are contiguous (row stride given in elements); w is (S,3,n) contiguous;
 *   out/grad_out are (B,) contiguous.  Device pointers must be aligned to the element size.
 */
#ifndef QUANONET_B200_H
#define QUANONET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QON_ABI_VERSION 3   /* 2: + qon_latency_tier_max_batch, qon_peer_*, qon_encoded_mse_step_dp; 3: + qon_tensor_tier */

/* dtype: arithmetic type of x, w, out, gradients and the state */
#define QON_F32 0 /* float32 / complex64  — parity with the TorchQuantum path            */
#define QON_F64 1 /* float64 / complex128 — parity with MindQuantum's double-precision path */

/* observable */
#define QON_HAM_DIAG 0    /* ham_diag given: H = diag(ham_diag); else H = offset + coeff * sum_q Z_q
                             (core/quantum_circuits_tq.py:112-126, _ham_params :141-146)             */
#define QON_HAM_PAULI_X 1 /* H = offset + coeff * sum_q X_q  (core/quantum_circuits_ms.py:28-39)     */
#define QON_HAM_PAULI_Y 2 /* H = offset + coeff * sum_q Y_q                                          */

/* index order of ham_diag */
#define QON_DIAG_LSB0 0 /* entry k: qubit q = bit q of k — MindQuantum/Qiskit builders
                           (core/quantum_circuits_ms.py:56-60)                                     */
#define QON_DIAG_MSB0 1 /* entry k: wire 0 = most-significant bit — TorchQuantum get_states_1d order,
                           what core/quantum_circuits_tq.py:112-114 multiplies against              */

/* error codes (negative) */
#define QON_ERR_BAD_ARG (-1)
#define QON_ERR_UNSUPPORTED (-2)
#define QON_ERR_WORKSPACE (-3)
#define QON_ERR_NO_DEVICE (-4)

int qon_abi_version(void);
const char* qon_last_error(void);

/* Bytes of device workspace needed by qon_hea_forward (need_grad = 0) or
 * qon_hea_forward_backward (need_grad = 1) for this problem on the current device.
 * Returns 0 and sets qon_last_error() on invalid arguments. */
size_t qon_workspace_bytes(int64_t B, int n, int K, const int* depth_per_block, int dtype, int need_grad);

/* Forward: out[b] = <psi_b|H|psi_b>.  Replaces _TQHEACircuit.forward + _measure
 * (core/quantum_circuits_tq.py:65-127).
 *   x               device, (B, n*K), row stride ldx elements
 *   w               device, (S, 3, n)
 *   out             device, (B,)
 *   depth_per_block HOST, (K,) ints, each >= 1
 *   ham_diag        device, (2^n,) or NULL
 */
int qon_hea_forward(const void* x, int64_t ldx, const void* w, void* out,
                    int64_t B, int n, int K, const int* depth_per_block,
                    const void* ham_diag, int diag_order, double ham_offset, double ham_coeff, int ham_kind,
                    int dtype, void* workspace, size_t workspace_bytes, void* stream);

/* Forward + adjoint backward in one pass.  Replaces forward + autograd backward of the same module
 * (what compare_backends.py:188-199 and solvers/solver_pt.py:233-235 exercise).
 *   grad_out  device, (B,)          dL/dout
 *   out       device, (B,)          written
 *   grad_x    device, (B, n*K) row stride ldgx, or NULL to skip it (fixed-scale encoders)
 *   grad_w    device, (S, 3, n)     OVERWRITTEN with sum_b grad_out[b] * dout[b]/dw
 * The batch reduction of grad_w is deterministic (fixed summation order for a given B and device).
 */
int qon_hea_forward_backward(const void* x, int64_t ldx, const void* w, const void* grad_out,
                             void* out, void* grad_x, int64_t ldgx, void* grad_w,
                             int64_t B, int n, int K, const int* depth_per_block,
                             const void* ham_diag, int diag_order, double ham_offset, double ham_coeff,
                             int ham_kind, int dtype, void* workspace, size_t workspace_bytes, void* stream);

/* Training-step variant: the upstream gradient of the reference's MSE loss
 * (loss_fn = nn.MSELoss(), solvers/solver_pt.py:66,233-235) is formed inside the kernel,
 *     g[b] = grad_scale * (out[b] + bias - target[b])        (grad_scale = 2 / B_global for a mean),
 * so forward, loss gradient and adjoint backward are ONE pass over the batch.
 *   target           device, (B,)
 *   bias             device scalar (the model's `bias` parameter, core/models_pt.py:151,166) or NULL (= 0)
 *   out              device, (B,)   expectation values WITHOUT bias
 *   grad_out_written device, (B,)   receives g[b]  (sum_b g[b] = dL/dbias; loss = sum (g/grad_scale)^2 / B)
 * Other arguments as qon_hea_forward_backward. */
int qon_hea_mse_forward_backward(const void* x, int64_t ldx, const void* w, const void* target, const void* bias,
                                 double grad_scale, void* out, void* grad_out_written, void* grad_x, int64_t ldgx,
                                 void* grad_w, int64_t B, int n, int K, const int* depth_per_block,
                                 const void* ham_diag, int diag_order, double ham_offset, double ham_coeff,
                                 int ham_kind, int dtype, void* workspace, size_t workspace_bytes, void* stream);

/* ---- Fused encoding ("the model in one kernel") ------------------------------------------------------
 * The callers of the circuit, QuanONetPT.forward / HEAQNNPT.forward (core/models_pt.py:153-166,205-213),
 * build the encoding-angle matrix x = cat([trunk_enc, branch_enc]) with the frequency layers
 * _TiledElementWise / _ScaleRepeat (core/models_pt.py:14-68):  x[b,c] = fw[c] * u_src(c)[b, local(c) % in] + fb[c],
 * src(c) = source 0 (trunk) for the first K0 blocks, source 1 (branch) afterwards; local(c) = column index
 * within its source.  These entry points evaluate that inside the kernel, so x and grad_x (1.2 KB per sample
 * each at Q5 Net40-2-20-2) are never materialised.  Built for n <= 5 (fp32) / n <= 4 (fp64) at any batch size and
 * for n = 6..9 (fp32) at latency-tier batch sizes — qon_encoded_supported() tells.
 *   u0  device (B, in0) row stride ldu0 — may be NULL when K0 == 0 (HEAQNN);   u1  device (B, in1), stride ldu1
 *   fw  device (n*K,) frequency weights (fixed-scale mode: the constant scale);  fb  device (n*K,) or NULL
 */
int qon_encoded_forward(const void* u0, int64_t ldu0, int in0, int K0, const void* u1, int64_t ldu1, int in1,
                        const void* fw, const void* fb, const void* w, void* out, int64_t B, int n, int K,
                        const int* depth_per_block, const void* ham_diag, int diag_order, double ham_offset,
                        double ham_coeff, int ham_kind, int dtype, void* workspace, size_t workspace_bytes,
                        void* stream);

/* One training step's gradient in one pass: frequency layers, forward, MSE upstream gradient
 * g[b] = grad_scale*(out[b] + bias - target[b]), adjoint backward, and the batch reductions of ALL parameter
 * gradients (replaces solvers/solver_pt.py:231-235 for QuanONetPT / HEAQNNPT).
 *   out      device (B,) or NULL       expectation values without bias
 *   grad_w   device (S,3,n)            overwritten
 *   grad_fw, grad_fb  device (n*K,) or both NULL (fixed-scale encoders)   overwritten
 *   sums     device (2,)               [sum_b g[b] (= dL/dbias),  sum_b (out+bias-target)^2] */
int qon_encoded_mse_step(const void* u0, int64_t ldu0, int in0, int K0, const void* u1, int64_t ldu1, int in1,
                         const void* fw, const void* fb, const void* w, const void* target, const void* bias,
                         double grad_scale, void* out, void* grad_w, void* grad_fw, void* grad_fb, void* sums,
                         int64_t B, int n, int K, const int* depth_per_block, const void* ham_diag, int diag_order,
                         double ham_offset, double ham_coeff, int ham_kind, int dtype, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Which kernel tier a problem maps to on the current device: 0 = register tier (state in registers,
 * 2^lanes_log2 lanes per sample), 1 = shared-memory tier, 2 = HBM-streamed tier; -1 = unsupported. */
int qon_plan_tier(int64_t B, int n, int dtype, int need_grad, int* lanes_log2);

/* Largest batch (per call) that the small-batch latency tier serves (n <= 5, angles given: one amplitude per
 * lane, see csrc/hea_warp.cuh); larger batches use the one-thread-per-sample throughput layout.  Callers that
 * can choose between the x-given and the fused-encoding entry points use it to pick the faster one. */
int64_t qon_latency_tier_max_batch(void);

/* 1 if the fused-encoding entry points (qon_encoded_forward / qon_encoded_mse_step[_dp]) have a kernel for a batch of
 * B samples at n qubits in `dtype` on the current device, else 0: n <= 5 (fp32) / n <= 4 (fp64) at any batch size,
 * and n = 6..9 in fp32 for batches small enough for the wide latency tier. */
int qon_encoded_supported(int64_t B, int n, int dtype, int need_grad);
/* The same question for a concrete circuit (K blocks, depth_per_block): for n = 6..9 the answer depends on the
 * circuit's size (the wide latency tier stages the whole gate table and the sample's angles in shared memory), so
 * callers that know the circuit ask this one and fall back to the x-given entry points on 0. */
int qon_encoded_supported_for(int64_t B, int n, int K, const int* depth_per_block, int dtype, int need_grad);

/* Exchange step of the data-parallel training step (the reference has no multi-GPU path — main.py:52
 * pins one GPU — so this replaces the torch.distributed all_reduce a DDP port would add at
 * solvers/solver_pt.py:235-236): one-shot sum all-reduce of a small fp32 vector over NVLink peer memory, ONE
 * single-CTA kernel per rank (push to every peer's slot, release/acquire flags, fixed-order sum; see
 * csrc/qon_peer.cuh).  `peer_bufs` is a HOST array of `world` device pointers to symmetric buffers of
 * qon_peer_buffer_bytes(max_len, world) bytes each, peer_bufs[rank] being this rank's own; the buffers must be
 * zero-filled once before first use and every rank must make the same sequence of calls.  src may equal dst.
 * A peer that does not show up within ~30 s poisons dst with NaN instead of hanging. */
size_t qon_peer_buffer_bytes(int64_t max_len, int world);
int qon_peer_allreduce_f32(const float* src, float* dst, int64_t len, void* const* peer_bufs, int world, int rank,
                           int64_t max_len, void* stream);

/* qon_encoded_mse_step for data-parallel training (fp32): compute step and exchange step in one pass.  The finalize
 * kernel pushes every gradient into the peers' slots of the symmetric buffers as it is produced, the last CTA
 * signals / waits / sums in rank order, and the all-reduced gradient of every parameter lands in `flat`:
 *   flat[w_off ...]    (S,3,n) ansatz gradient          flat[fw_off ...], flat[fb_off ...]  (n*K,) each, or both -1
 *   flat[sums_off ..]  [sum g (= dL/dbias), sum of squared residuals]        every other index of flat: 0
 * peer_bufs / world / rank / max_len as in qon_peer_allreduce_f32 (same buffers, same protocol: the two calls may be
 * mixed on one buffer as long as every rank makes the same sequence). */
int qon_encoded_mse_step_dp(const void* u0, int64_t ldu0, int in0, int K0, const void* u1, int64_t ldu1, int in1,
                            const void* fw, const void* fb, const void* w, const void* target, const void* bias,
                            double grad_scale, void* out, float* flat, int64_t flat_len, int64_t w_off, int64_t fw_off,
                            int64_t fb_off, int64_t sums_off, void* const* peer_bufs, int world, int rank,
                            int64_t max_len, int64_t B, int n, int K, const int* depth_per_block, const void* ham_diag,
                            int diag_order, double ham_offset, double ham_coeff, int ham_kind, void* workspace,
                            size_t workspace_bytes, void* stream);

/* Tensor-core tier (n = 5, fp32, diagonal observable — csrc/hea_tc.cuh, hea_tc2.cuh, hea_tc3.cuh): every entry point above
 * routes batches of at least `min_batch` samples to it: the ansatz sublayers of a block (sample-independent in the
 * reference, core/quantum_circuits_tq.py:89-101) pre-fused into one 32x32 unitary and applied as split-f16 GEMMs on
 * tcgen05 tensor cores, the RX encoding layers as diagonal phases in the Hadamard basis; in gradient calls the weight
 * gradients come from one batch-summed outer product per block on the tensor cores as well; same results to the 1e-5
 * norm-relative bar.  On by default (QON_TC=0 in the environment disables it; QON_TC_MIN_B sets the threshold).
 *   enable     1 = on, 0 = off (the FFMA2 register kernels serve every batch), -1 = leave unchanged;
 *              2 / 3 = on, with the earlier formulations of the gradient step (per-sublayer Pauli-string moments in
 *              two kernels / in one kernel), kept for A/B measurements
 *   min_batch  smallest batch routed to the tier (default 5121); < 0 = leave unchanged
 *   debug_state / error_flag: device pointers for kernel bring-up (state dump of the first tile; protocol
 *   time-out flag), NULL in production — a time-out poisons the outputs with NaN, it never hangs.
 * Process-wide; returns the previous `enable`. */
int qon_tensor_tier(int enable, int64_t min_batch, void* debug_state, void* error_flag);

/* FP32 FFMA-saturating micro-benchmark (the metric is "% of FP32 peak" and MEASURED_PEAKS.json has
 * no FP32 entry): runs `iters` dependent-chain FFMA rounds on every SM and returns achieved
 * TFLOP/s measured with CUDA events on `stream`; negative on error.  Synchronises the stream. */
double qon_measure_fp32_peak_tflops(int iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QUANONET_B200_H */
