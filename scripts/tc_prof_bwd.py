"""Short run of the fused-encoding MSE training step for ncu: python scripts/tc_prof_bwd.py <tc 0|1> [B]"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quanonet_b200 import _lib
from quanonet_b200.ops import encoded_mse_step
lib = _lib.load()
tc = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 256 * 2
lib.qon_tensor_tier(tc, 0, None, None)
dev = torch.device("cuda:0")
n, td, bd = 5, 20, 40
depths = [2] * 60
g = torch.Generator().manual_seed(0)
branch = torch.randn(B, 100, generator=g).to(dev); trunk = torch.rand(B, 2, generator=g).to(dev)
y = torch.randn(B, generator=g).to(dev)
fw = (torch.randn(300, generator=g) * 0.3).to(dev); fb = ((torch.rand(300, generator=g) * 2 - 1) * np.pi).to(dev)
w = ((torch.rand(120, 3, 5, generator=g) * 2 - 1) * np.pi).to(dev)
bias = torch.tensor([0.05], device=dev)
for _ in range(3):
    out = encoded_mse_step(trunk, branch, fw, fb, td, w, y, bias, 2.0 / B, n, depths, None, 0, 0.0, 1.0, 0, True)
torch.cuda.synchronize()
print("ok", float(out[3][1]))
