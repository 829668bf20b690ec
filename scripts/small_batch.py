"""Latency of one training step at the reference's default batch sizes (utils/common.py:128: batch_size = 100)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from quanonet_b200.train import DataParallelTrainer, autograd_step
dev = torch.device("cuda:0")
for B in (100, 1000, 10000, 100000):
    model = bench.make_model(dev)
    tr = DataParallelTrainer(model, lr=1e-3)
    branch, trunk, y = bench.synth_batch(B, 1, device=dev)
    for _ in range(5): tr.step((branch, trunk), y)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 50
    for _ in range(n): tr.step((branch, trunk), y)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    for p in model.parameters(): p.grad = None
    for _ in range(3): autograd_step(model, opt, (branch, trunk), y)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): autograd_step(model, opt, (branch, trunk), y)
    torch.cuda.synchronize(); dta = (time.perf_counter() - t0) / 20
    print(f"B={B:7d}: fused step {dt*1e3:8.3f} ms ({B/dt:.3e} samples/s) | reference-style autograd step {dta*1e3:8.3f} ms ({B/dta:.3e} samples/s)")

# kernel-only latency and CUDA-graph replay of the whole step
from quanonet_b200.ops import hea_mse_backward
print("--- kernel only (x given) and graph-replayed training step ---")
for B in (100, 1000, 4096):
    model = bench.make_model(dev)
    tr = DataParallelTrainer(model, lr=1e-3, optimizer_kwargs={'capturable': True})
    branch, trunk, y = bench.synth_batch(B, 1, device=dev)
    x = torch.cat([model.trunk_freq(trunk), model.branch_freq(branch)], 1).detach()
    q = model.quantum_layer; depths = [d for _, d in q.block_configs]
    f = lambda: hea_mse_backward(x, q.ansatz_weights.detach(), y.reshape(-1), model.bias.detach(), 2.0 / B, 5, depths, None, 0, 0.0, 1.0, 0, True)
    for _ in range(3): f()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize(); kms = e0.elapsed_time(e1) / 20
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): tr.step((branch, trunk), y)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        tr.step((branch, trunk), y)
    for _ in range(5): g.replay()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): g.replay()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 200
    print(f"B={B:6d}: op call (prep+kernel+finalize, incl. host) {kms*1e3:7.1f} us | CUDA-graph step {dt*1e6:7.1f} us ({B/dt:.3e} samples/s)")
