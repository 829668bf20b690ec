"""A/B of the training-step kernels at the bench config (fused encoding + MSE + adjoint gradients, B = 1M):
FFMA2 register kernel (tier 0) vs tensor-core layouts (1 = 2 tiles, 3 = 3 tiles sequential).
    python scripts/tc_ab.py [B] [tiers...]"""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quanonet_b200 import _lib
from quanonet_b200.ops import encoded_mse_step, hea_expval
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
tiers = [int(v) for v in sys.argv[2:]] or [0, 1, 3]
dev = torch.device("cuda:0")
n, td = 5, 20
depths = [2] * 60
g = torch.Generator().manual_seed(0)
branch = torch.randn(B, 100, generator=g).to(dev); trunk = torch.rand(B, 2, generator=g).to(dev)
y = torch.randn(B, generator=g).to(dev)
fw = (torch.randn(300, generator=g) * 0.3).to(dev); fb = ((torch.rand(300, generator=g) * 2 - 1) * np.pi).to(dev)
w = ((torch.rand(120, 3, 5, generator=g) * 2 - 1) * np.pi).to(dev)
bias = torch.tensor([0.05], device=dev)
ref = None
out = {}
for tier in tiers:
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    lib.qon_tensor_tier(tier, 0, None, err.data_ptr())
    for _ in range(2):
        r = encoded_mse_step(trunk, branch, fw, fb, td, w, y, bias, 2.0 / B, n, depths, None, 0, 0.0, 1.0, 0, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        r = encoded_mse_step(trunk, branch, fw, fb, td, w, y, bias, 2.0 / B, n, depths, None, 0, 0.0, 1.0, 0, True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    vec = torch.cat([t.reshape(-1).double() for t in r]).cpu().numpy()
    if ref is None:
        ref = vec
    rel = float(np.linalg.norm(vec - ref) / np.linalg.norm(ref))
    out[tier] = dict(ms=ms, samples_per_s=B / ms * 1e3, rel_vs_first=rel, err=int(err.item()))
    print(f"tier {tier}: {ms:.3f} ms  {B / ms * 1e3:.3e} samples/s  rel vs tier {tiers[0]}: {rel:.2e}  err {int(err.item())}", flush=True)
lib.qon_tensor_tier(1, 5121, None, None)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tc_ab.json"), "w"))
