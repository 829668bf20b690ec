"""Short ncu target for the HBM-streamed tier: forward and fwd+grad at n = 16, one block of depth 2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quanonet_b200.ops import hea_expval, hea_expval_backward
dev = torch.device("cuda:0")
n, B = 16, 296
depths = [2, 2]
g = torch.Generator().manual_seed(0)
x = ((torch.rand(B, n * 2, generator=g) * 2 - 1) * np.pi).to(dev)
w = ((torch.rand(4, 3, n, generator=g) * 2 - 1) * np.pi).to(dev)
go = torch.randn(B, generator=g).to(dev)
for _ in range(2):
    hea_expval(x, w, n, depths, None, 0, 0.0, 0.5, 0)
    hea_expval_backward(go, x, w, n, depths, None, 0, 0.0, 0.5, 0, True)
torch.cuda.synchronize()
print("ok")
