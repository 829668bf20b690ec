"""fwd+grad kernel-call latency at small batches for n = 2..12 (net (20,2,10,2)-like: 30 blocks of depth 2).
usage: small_batch_widths.py [n ...]   env SB_B="100,1000" batch sizes, SB_DTYPE=f32|f64"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quanonet_b200.ops import hea_mse_backward, plan_tier
dev = torch.device("cuda:0")
Bs = [int(b) for b in os.environ.get("SB_B", "100").split(",")]
dt = torch.float64 if os.environ.get("SB_DTYPE", "f32") == "f64" else torch.float32
K = 30
for n in [int(a) for a in sys.argv[1:]] or range(2, 13):
    for B in Bs:
        g = torch.Generator().manual_seed(n)
        depths = [2] * K
        x = ((torch.rand(B, n * K, generator=g) * 2 - 1) * np.pi).to(dev, dt)
        w = ((torch.rand(2 * K, 3, n, generator=g) * 2 - 1) * np.pi).to(dev, dt)
        y = torch.randn(B, generator=g).to(dev, dt); bias = torch.zeros(1, device=dev, dtype=dt)
        f = lambda: hea_mse_backward(x, w, y, bias, 2.0 / B, n, depths, None, 0, 0.0, 1.0, 0, True)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        print(f"n={n:2d} B={B:6d}: fwd+grad call {e0.elapsed_time(e1) / 10 * 1e3:9.1f} us  tier={plan_tier(B, n, dt)} {dt}", flush=True)
