"""Short ncu target: a few launches of the forward and the fwd+grad kernels at the C2 shape
(x given, and the fused-encoding training kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quanonet_b200.ops import hea_expval, hea_mse_backward, encoded_mse_step
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 256 * 4
n, depths = 5, [2] * 60
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = ((torch.rand(B, 300, generator=g) * 2 - 1) * np.pi).to(dev)
w = ((torch.rand(120, 3, 5, generator=g) * 2 - 1) * np.pi).to(dev)
y = torch.randn(B, generator=g).to(dev)
branch = torch.randn(B, 100, generator=g).to(dev); trunk = torch.rand(B, 2, generator=g).to(dev)
fw = torch.full((300,), 0.1, device=dev); fb = ((torch.rand(300, generator=g) * 2 - 1) * np.pi).to(dev)
bias = torch.zeros(1, device=dev)
for _ in range(3):
    hea_expval(x, w, n, depths, None, 0, 0.0, 1.0, 0)
    hea_mse_backward(x, w, y, bias, 2.0 / B, n, depths, None, 0, 0.0, 1.0, 0, True)
    encoded_mse_step(trunk, branch, fw, fb, 20, w, y, bias, 2.0 / B, n, depths, None, 0, 0.0, 1.0, 0, True)
torch.cuda.synchronize()
print("ok")
