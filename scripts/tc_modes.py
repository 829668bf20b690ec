"""Timing of the tensor-core training-step kernels by mode at K = 60, depth 2 (B samples):
angles given (mode 2: weight gradients only, mode 1: + dL/dx) vs fused encoding (mode 5).
    python scripts/tc_modes.py [B] [tier]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quanonet_b200 import _lib
from quanonet_b200.ops import _backward_impl, encoded_mse_step, hea_expval
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
tier = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lib.qon_tensor_tier(tier, 0, None, None)
dev = torch.device("cuda:0")
n, td = 5, 20
depths = [2] * 60
g = torch.Generator().manual_seed(0)
x = ((torch.rand(B, 300, generator=g) * 2 - 1) * np.pi).to(dev)
gout = torch.randn(B, generator=g).to(dev)
w = ((torch.rand(120, 3, 5, generator=g) * 2 - 1) * np.pi).to(dev)
branch = torch.randn(B, 100, generator=g).to(dev); trunk = torch.rand(B, 2, generator=g).to(dev)
y = torch.randn(B, generator=g).to(dev)
fw = (torch.randn(300, generator=g) * 0.3).to(dev); fb = ((torch.rand(300, generator=g) * 2 - 1) * np.pi).to(dev)
bias = torch.tensor([0.05], device=dev)


def timed(name, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"tier {tier} {name}: {ms:.3f} ms  {B / ms * 1e3:.3e} samples/s", flush=True)


timed("forward, angles given", lambda: hea_expval(x, w, n, depths, None, 0, 0.0, 1.0, 0))
timed("grad w (mode 2), angles given", lambda: _backward_impl(gout, x, w, n, depths, None, 0, 0.0, 1.0, 0, False))
timed("grad w + dL/dx (mode 1), angles given", lambda: _backward_impl(gout, x, w, n, depths, None, 0, 0.0, 1.0, 0, True))
timed("fused encoding + MSE + frequency-layer gradients (mode 5)",
      lambda: encoded_mse_step(trunk, branch, fw, fb, td, w, y, bias, 2.0 / B, n, depths, None, 0, 0.0, 1.0, 0, True))
lib.qon_tensor_tier(1, 5121, None, None)
