"""ncu target: the small-batch latency tier (Q5, depth-2 blocks), forward and fwd+grad, at several batch sizes.
Usage: profile_target_small.py B [B ...]   (env PT_K = number of blocks, default 60 = Net40-2-20-2)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quanonet_b200.ops import hea_expval, hea_mse_backward
Bs = [int(a) for a in sys.argv[1:]] or [100]
K = int(os.environ.get("PT_K", "60"))
n, depths = 5, [2] * K
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
w = ((torch.rand(2 * K, 3, 5, generator=g) * 2 - 1) * np.pi).to(dev)
bias = torch.zeros(1, device=dev)
for B in Bs:
    x = ((torch.rand(B, n * K, generator=g) * 2 - 1) * np.pi).to(dev)
    y = torch.randn(B, generator=g).to(dev)
    for _ in range(3):
        hea_expval(x, w, n, depths, None, 0, 0.0, 1.0, 0)
        hea_mse_backward(x, w, y, bias, 2.0 / B, n, depths, None, 0, 0.0, 1.0, 0, True)
    torch.cuda.synchronize()
print("ok")
