"""compute-sanitizer target: every latency-tier kernel variant on tiny problems (ragged batches, mixed depths)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quanonet_b200.ops import hea_expval, hea_expval_backward, hea_mse_backward, encoded_expval, encoded_mse_step
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
depths = [1, 2, 3]
K, S = len(depths), sum(depths)
for dtype in (torch.float32, torch.float64):
    for n in range(1, 10):
        if n > 5 and dtype == torch.float64:
            continue
        B = 37
        x = ((torch.rand(B, n * K, generator=g) * 2 - 1) * np.pi).to(dev, dtype)
        w = ((torch.rand(S, 3, n, generator=g) * 2 - 1) * np.pi).to(dev, dtype)
        y = torch.randn(B, generator=g).to(dev, dtype); bias = torch.zeros(1, device=dev, dtype=dtype)
        hea_expval(x, w, n, depths, None, 0, 0.0, 1.0, 0)
        hea_expval_backward(y, x, w, n, depths, None, 0, 0.3, 0.7, 1, True)
        hea_expval_backward(y, x, w, n, depths, None, 0, 0.3, 0.7, 2, False)
        hea_mse_backward(x, w, y, bias, 2.0 / B, n, depths, None, 0, 0.0, 1.0, 0, True)
        if n <= (5 if dtype == torch.float32 else 4):
            u0 = torch.rand(B, 2, generator=g).to(dev, dtype); u1 = torch.randn(B, 7, generator=g).to(dev, dtype)
            fw = torch.rand(n * K, generator=g).to(dev, dtype); fb = torch.rand(n * K, generator=g).to(dev, dtype)
            encoded_expval(u0, u1, fw, fb, 1, w, n, depths, None, 0, 0.0, 1.0, 0)
            encoded_mse_step(u0, u1, fw, fb, 1, w, y, bias, 2.0 / B, n, depths, None, 0, 0.0, 1.0, 0, True)
            encoded_mse_step(u0, u1, fw, None, 1, w, y, bias, 2.0 / B, n, depths, None, 0, 0.0, 1.0, 0, False)
torch.cuda.synchronize()
print("sanitize target ok")
