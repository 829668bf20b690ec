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
