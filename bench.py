#!/usr/bin/env python
"""bench.py — QuanONet Q5 fwd+adjoint-grad training throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], SURVEY §8d "C2"): Advection-shaped QuanONet, n = 5 qubits,
net (40,2,20,2), branch_in 100, trunk_in 2, trainable frequency layers; synthetic batch of 1M
(function, query-point) samples PER GPU (weak scaling: global batch = N x 1M); one step = one full
training step over the batch: frequency layers -> fused forward + MSE + adjoint-backward kernel ->
chain rule to the frequency parameters -> all-reduce of the flat 2,401-float gradient (N > 1) ->
Adam update.  Prints ONE JSON line (rank 0).

`value`   : whole-job samples/s with the batch resident in HBM (CUDA events, max over ranks).
`e2e`     : the same step fed from pinned HOST memory every step (H2D of branch/trunk/target inside
            the timed region, double-buffered on a copy stream) plus a D2H read of the loss.
`roofline`: the dominant kernel (hea_reg_kernel, fwd+grad) against the FP32 FFMA peak measured on this
            GPU in the same run (MEASURED_PEAKS.json has no FP32 entry); algorithmic flops per sample
            = 22*N*G + 7*N = 1,478,624.
`cpu_baseline` / `--impl reference`: the TorchQuantum-faithful complex64 restatement of the
            reference path (oracle/tq_faithful.py; torchquantum itself is not installable here)
            timed on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

N_QUBITS = 5
NET = (40, 2, 20, 2)          # branch_depth, branch_linear_depth, trunk_depth, trunk_linear_depth
BRANCH_IN, TRUNK_IN = 100, 2
METRIC = "quanonet_q5_fwd_adjoint_grad_samples_per_s"
WORKLOAD = "Advection-shaped QuanONet Net40-2-20-2 Q5 (TF), MSE training step, synthetic"


def alg_flops(n=N_QUBITS, net=NET):
    N = 1 << n
    K = net[0] + net[2]
    S = net[0] * net[1] + net[2] * net[3]
    G = n * K + 3 * n * S
    return 6 * N * G + 5 * N, 22 * N * G + 7 * N


def make_model(device, seed=0):
    from quanonet_b200.core.models_pt import QuanONetPT
    torch.manual_seed(seed)
    m = QuanONetPT(N_QUBITS, BRANCH_IN, TRUNK_IN, NET, scale_coeff=0.1, if_trainable_freq=True,
                   ham_bound=(-5.0, 5.0))
    with torch.no_grad():   # MindSpore initialises the frequency bias U(-pi, pi) (core/layers.py:24-27)
        m.branch_freq.bias.uniform_(-np.pi, np.pi)
        m.trunk_freq.bias.uniform_(-np.pi, np.pi)
    return m.to(device)


def synth_batch(B, seed, device=None, pin=False):
    g = torch.Generator().manual_seed(seed)
    branch = torch.randn(B, BRANCH_IN, generator=g)
    trunk = torch.rand(B, TRUNK_IN, generator=g)
    y = torch.randn(B, 1, generator=g)
    if device is not None:
        return branch.to(device), trunk.to(device), y.to(device)
    if pin:
        return branch.pin_memory(), trunk.pin_memory(), y.pin_memory()
    return branch, trunk, y


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU every 200 ms (NVML) while active."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, device_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = torch.cuda.get_device_properties(device_index).uuid
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                try:
                    bits = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    bits = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.025)      # the timed region is ~0.25 s at the default K: sample every 25 ms

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------- shared config
def bench_config(world, B, fused=True, scaling="weak"):
    """`config` of the JSON line — identical for the B200 arm and the reference arm."""
    return {"workload": WORKLOAD, "num_qubits": N_QUBITS, "net_size": list(NET),
            "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
            "l2_policy": "inputs larger than L2: 412 MB of (branch, trunk, target) per step vs 126 MB L2",
            "step": ("ONE kernel: frequency layers + forward + MSE + adjoint-grad + batch reduction of all "
                     "2,401 gradients; then all-reduce + Adam" if fused else
                     "freq layers (torch) + fused fwd/MSE/adjoint-grad kernel + chain rule (torch) + "
                     "all-reduce + Adam")}


# ---------------------------------------------------------------------------------- CPU reference arm
REF_BATCH = 1000          # bounded sample of the 1M-sample batch; the reference's own default is 100 (utils/common.py:128)
REF_BATCH_SMALL = 100


def cpu_reference_step_fn(B, threads=None):
    """One training step of the reference path on the host: frequency layers, TorchQuantum-faithful
    complex64 circuit with autograd (oracle/tq_faithful.py), MSELoss, backward, Adam."""
    from oracle.tq_faithful import tq_forward
    if threads:
        torch.set_num_threads(threads)
    model = make_model("cpu")
    blocks = model.quantum_layer.block_configs
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    branch, trunk, y = synth_batch(B, seed=1)

    def step():
        opt.zero_grad()
        x = torch.cat([model.trunk_freq(trunk), model.branch_freq(branch)], dim=1)
        pred = tq_forward(x, model.quantum_layer.ansatz_weights, N_QUBITS, blocks, 0.0, 1.0) + model.bias
        loss = torch.nn.functional.mse_loss(pred, y)
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step


def time_cpu_reference(B, steps, warmup):
    """(samples/s, ms/step) of the CPU reference: `steps` training steps of a FIXED batch of B samples."""
    step = cpu_reference_step_fn(B)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    sps, ms = time_cpu_reference(REF_BATCH, steps, warmup)
    sps_small, ms_small = time_cpu_reference(REF_BATCH_SMALL, max(3, min(steps, 10)), 1)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(max(1, args.gpus), args.batch),
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} training steps of a fixed {REF_BATCH}-sample slice of the batch "
                                   f"(TorchQuantum-faithful complex64 restatement with autograd; torchquantum is "
                                   f"not installable offline)",
                         "batch_per_step": REF_BATCH,
                         "reference_default_batch": {"batch_per_step": REF_BATCH_SMALL, "value": sps_small,
                                                     "ms_per_step": ms_small}},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch.distributed as dist
    from quanonet_b200 import _lib
    from quanonet_b200.ops import fp32_peak_tflops, hea_expval
    from quanonet_b200.train import DataParallelTrainer

    _lib.load()                                   # fail loudly if the CUDA library is missing
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback for the B200 arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    K, W = args.steps, max(args.warmup, 3)
    f_fwd, f_all = alg_flops()
    sampler = ClockSampler(local)                 # NVML init + handle lookup happen HERE, before any timed region
    align = torch.zeros(1, device=dev)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def aligned_start():
        """barrier + synchronize, then a device-side rendezvous (a 1-float all-reduce the timed stream waits on)
        so every rank's first timed kernel starts within microseconds of the others, whatever the host skew."""
        sync_all()
        if world > 1:
            dist.all_reduce(align)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(trainer, data, n_steps, gB=None):
        """n_steps training steps with an event after every step.  Returns (total ms max over ranks, total ms of
        the fastest rank, per-step trace dict, last loss)."""
        (branch, trunk), y = data
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps + 1)]
        aligned_start()
        evs[0].record()
        for i in range(n_steps):
            loss = trainer.step((branch, trunk), y, gB)
            evs[i + 1].record()
        sync_all()
        total = evs[0].elapsed_time(evs[-1])
        per = torch.tensor([evs[i].elapsed_time(evs[i + 1]) for i in range(n_steps)], device=dev, dtype=torch.float64)
        if world > 1:
            allper = [torch.empty_like(per) for _ in range(world)]
            dist.all_gather(allper, per)
            allper = torch.stack(allper)            # (world, steps)
        else:
            allper = per[None]
        step_max, step_min = allper.max(0).values, allper.min(0).values
        trace = {"first_step_ms_max": float(step_max[0]), "first_step_ms_min": float(step_min[0]),
                 "later_steps_ms_max_mean": float(step_max[1:].mean()) if n_steps > 1 else None,
                 "later_steps_ms_min_mean": float(step_min[1:].mean()) if n_steps > 1 else None,
                 "rank_total_ms": [float(v) for v in allper.sum(1)]}
        return max_over_ranks(total), -max_over_ranks(-total), trace, loss

    # ---- device-resident timing (value): weak scaling, B samples per GPU
    model = make_model(dev, seed=0)
    trainer = DataParallelTrainer(model, lr=1e-3, optimizer="adam", use_fused_encoding=not args.unfused)
    kernel_events = trainer.kernel_events = []
    branch, trunk, y = synth_batch(B, seed=100 + rank, device=dev)
    peak = fp32_peak_tflops(4000) if rank == 0 else None
    for _ in range(W):
        trainer.step((branch, trunk), y)
    kernel_events.clear()
    with sampler:
        ms_total, ms_min, trace, loss = timed_steps(trainer, ((branch, trunk), y), K)
    ms_step = ms_total / K
    value = world * B * K / (ms_total * 1e-3)
    final_loss = float(loss)

    # ---- replicas stay identical; the fused finalize+exchange kernel agrees with NCCL (N > 1)
    replica_check = None
    if world > 1:
        flat = trainer.flat_param.detach().clone()
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(bool(torch.equal(g.view(torch.int32), gathered[0].view(torch.int32))) for g in gathered)
        replica_check = {"params_bit_identical_across_ranks": same, "n_params": int(flat.numel())}
        saved_ar, saved_fx = trainer._all_reduce, trainer._fused_exchange
        trainer.compute_grads((branch, trunk), y)
        g_lib = trainer.flat_grad.detach().clone()
        trainer._fused_exchange = False
        trainer._all_reduce = lambda t: dist.all_reduce(t)
        trainer.compute_grads((branch, trunk), y)
        g_nccl = trainer.flat_grad.detach().clone()
        trainer._all_reduce, trainer._fused_exchange = saved_ar, saved_fx
        rel = float((g_lib - g_nccl).norm() / g_nccl.norm())
        replica_check.update({"exchange": "fused finalize+exchange kernel" if saved_fx else type(saved_ar).__name__,
                              "grad_rel_diff_vs_nccl_allreduce": max_over_ranks(rel)})
        torch.cuda.synchronize()

    if getattr(trainer, "_fused_exchange", False):
        # N > 1: the finalize kernel of the timed steps also waits for the peers (fused exchange), so its events
        # include rank skew.  For the roofline, time the compute kernels alone on a few extra, untimed steps.
        trainer._fused_exchange = False
        kernel_events.clear()
        for _ in range(3):
            trainer.step((branch, trunk), y)
        sync_all()
        trainer._fused_exchange = True
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    # the same kernel call on the FFMA2 register kernels (tensor-core tier switched off), for the A/B in the record
    from quanonet_b200.ops import tensor_tier
    tc_on = bool(tensor_tier(None)) and B >= 5121 and not args.unfused
    ffma2_ms = None
    if tc_on and rank == 0 or (tc_on and world > 1):
        saved_fx = getattr(trainer, "_fused_exchange", False)
        trainer._fused_exchange = False
        tensor_tier(False)
        kernel_events.clear()
        saved_ar = trainer._all_reduce
        trainer._all_reduce = (lambda t: t) if world > 1 else saved_ar      # kernels only; replicas resynchronised below
        flat_saved = trainer.flat_param.detach().clone()
        for _ in range(3):
            trainer.compute_grads((branch, trunk), y)
        sync_all()
        ffma2_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events][1:]))
        tensor_tier(True)
        trainer._all_reduce, trainer._fused_exchange = saved_ar, saved_fx
        trainer.flat_param.copy_(flat_saved)
    trainer.kernel_events = None

    # ---- strong scaling (BASELINE config 3): the SAME global batch of `B` samples sharded B/N per GPU
    strong = None
    if world > 1:
        Bs = B // world
        sdata = ((branch[:Bs], trunk[:Bs]), y[:Bs])
        for _ in range(W):
            trainer.step(*sdata, Bs * world)
        s_total, s_min, s_trace, _ = timed_steps(trainer, sdata, K, Bs * world)
        strong = {"scaling": "strong", "global_batch": Bs * world, "batch_per_gpu": Bs, "ms_per_step": s_total / K,
                  "ms_per_step_fastest_rank": s_min / K, "value": Bs * world * K / (s_total * 1e-3),
                  "unit": "samples/s", "per_step": s_trace}

    # ---- forward-only throughput (inference path), same batch
    with torch.no_grad():
        x_enc = torch.cat([model.trunk_freq(trunk), model.branch_freq(branch)], dim=1)
        q = model.quantum_layer
        depths = [d for _, d in q.block_configs]
        for _ in range(2):
            hea_expval(x_enc, q.ansatz_weights, N_QUBITS, depths, None, 0, q.ham_offset, q.ham_coeff, 0)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(5):
            hea_expval(x_enc, q.ansatz_weights, N_QUBITS, depths, None, 0, q.ham_offset, q.ham_coeff, 0)
        f1.record()
        torch.cuda.synchronize()
        fwd_ms = f0.elapsed_time(f1) / 5
        del x_enc

    # ---- end to end from pinned host memory (e2e), through the product API: DataParallelTrainer.run_host_fed
    #      double-buffers the H2D copy of batch i+1 under the step of batch i and reads every step's loss back
    host = [synth_batch(B, seed=200 + rank + 7 * i, pin=True) for i in range(2)]

    def host_batches(n_steps):
        for i in range(n_steps):
            b_, t_, y_ = host[i % 2]
            yield (b_, t_), y_

    trainer.run_host_fed(host_batches(3))
    aligned_start()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    g0.record()
    e2e_losses = trainer.run_host_fed(host_batches(K))
    g1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        dist.barrier()
    # the slower of the device clock and the host clock around the same K steps, max over ranks
    e2e_ms = max_over_ranks(max(g0.elapsed_time(g1), wall_ms))
    e2e_value = world * B * K / (e2e_ms * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    assert len(e2e_losses) == K and all(v == v for v in e2e_losses)

    # ---- the reference's default batch (utils/common.py:128: 100 samples per step): the latency-bound point
    #      of SURVEY §8(e).  Eager steps at every N (all-reduce included); CUDA-graph replay of the step at N=1.
    sB = 100
    small_model = make_model(dev, seed=0)
    small_tr = DataParallelTrainer(small_model, lr=1e-3, optimizer="adam", optimizer_kwargs={"capturable": True})
    sb, st_, sy = synth_batch(sB, seed=300 + rank, device=dev)
    for _ in range(5):
        small_tr.step((sb, st_), sy)
    aligned_start()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(50):
        small_tr.step((sb, st_), sy)
    s1.record()
    sync_all()
    small = {"batch_per_gpu": sB, "eager_us_per_step": max_over_ranks(s0.elapsed_time(s1)) / 50 * 1e3}
    if world == 1:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            small_tr.step((sb, st_), sy)
        for _ in range(5):
            graph.replay()
        torch.cuda.synchronize()
        s0.record()
        for _ in range(200):
            graph.replay()
        s1.record()
        torch.cuda.synchronize()
        small["graph_us_per_step"] = s0.elapsed_time(s1) / 200 * 1e3
        small["graph_samples_per_s"] = sB / (small["graph_us_per_step"] * 1e-6)
        del graph

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sps, ms = time_cpu_reference(REF_BATCH, 5, 1)
        cpu_base = {"value": sps, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                    "sample": f"5 training steps of a fixed {REF_BATCH}-sample slice of the same workload "
                              f"(TorchQuantum-faithful complex64 restatement, autograd backward)",
                    "batch_per_step": REF_BATCH, "ms_per_step": ms}

    if rank == 0:
        achieved = f_all * B / (kern_ms * 1e-3) / 1e12
        nominal = 148 * 128 * 2 * 1.965e9 / 1e12
        traffic = None
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):   # dram bytes per SAMPLE from the committed `ncu --set full` capture
            try:
                key = ("tensor_tier_bytes_per_sample" if tc_on else
                       "fused_encoding_bytes_per_sample" if trainer.fused_encoding else "x_given_bytes_per_sample")
                traffic = json.load(open(tf))[key] * B
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "ms_per_step_fastest_rank": ms_min / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": bench_config(world, B, trainer.fused_encoding),
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / K, "api": "DataParallelTrainer.run_host_fed",
                    "timer": "max(CUDA events, host wall clock) over the K steps, max over ranks"},
            # this library's kernels per step: prep, circuit kernel, finalize (+ finalize_enc) — with N > 1 the
            # finalize kernel is also the all-reduce (finalize_exchange_kernel), else one peer all-reduce kernel more;
            # the tensor-core tier adds its operand-image prep kernel (tc_prep_all_kernel), runs the step as forward-only +
            # reverse kernels and turns the batch-summed outer products into Pauli moments in two more
            # (tc_slot_reduce_kernel, tc_moment_kernel): prep, tc_prep_all, forward, reverse, slot sums, moments, finalize x2
            "gpu_launches": ((3 if getattr(trainer, "_fused_exchange", False) else
                              (4 if trainer.fused_encoding else 3) + (1 if world > 1 else 0)) + (4 if tc_on else 0)) * K,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "kernel": ("hea_tc_kernel<forward> + hea_tc_rev_kernel<grad%s> (tcgen05: block-unitary GEMMs, batch-summed outer products for the weight gradients; FFMA2 phases; +prep, tc_moment_kernel, finalize)"
                                    if tc_on else "hea_reg_kernel<float,5,0,grad%s> (+prep, finalize)") % (
                             ",fused-encoding" if trainer.fused_encoding else ""), "kernel_ms": kern_ms,
                         "flops_per_sample": f_all, "peak_source": "FFMA probe measured on this GPU in this run "
                         "(MEASURED_PEAKS.json has no FP32 entry)", "peak_nominal": nominal,
                         "frac_of_nominal": achieved / nominal,
                         "note": ("ALGORITHMIC gate-by-gate flops (SURVEY §8d) over the FP32 CUDA-core peak; with the tensor-core "
                                  "tier on, the sample-independent sublayers run as split-f16 GEMMs on tcgen05 (executed tensor "
                                  "flops are not credited), so the fraction can approach or exceed 1" if tc_on else
                                  "ALGORITHMIC gate-by-gate flops (SURVEY §8d) over the FP32 CUDA-core peak")},
            "tensor_tier": {"enabled": tc_on, "min_batch": 5121,
                            "ffma2_kernel_ms": ffma2_ms, "ffma2_samples_per_s": (B / (ffma2_ms * 1e-3)) if ffma2_ms else None},
            "per_step": trace,
            "forward_only": {"value": B / (fwd_ms * 1e-3), "unit": "samples/s",
                             "tflops": f_fwd * B / (fwd_ms * 1e-3) / 1e12},
            "small_batch": small,
            "final_loss": final_loss,
        }
        if strong is not None:
            line["strong_scaling"] = strong
        if replica_check is not None:
            line["replica_check"] = replica_check
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1_000_000, help="samples per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--unfused", action="store_true",
                    help="materialise the encoding matrix with torch ops instead of the fused-encoding kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
