"""Short forward run of the tensor-core tier for ncu (B = 4 rounds of 148 x 512 samples)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quanonet_b200 import _lib
from quanonet_b200.ops import hea_expval
lib = _lib.load()
tc = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B = int(sys.argv[2]) if len(sys.argv) > 2 else 148 * 512 * 4
lib.qon_tensor_tier(tc, 0, None, None)
dev = torch.device("cuda:0")
n, depths = 5, [2] * 60
g = torch.Generator().manual_seed(0)
x = ((torch.rand(B, 300, generator=g) * 2 - 1) * np.pi).to(dev)
w = ((torch.rand(120, 3, 5, generator=g) * 2 - 1) * np.pi).to(dev)
for _ in range(3):
    o = hea_expval(x, w, n, depths, None, 0, 0.0, 1.0, 0)
torch.cuda.synchronize()
print("ok", float(o.sum()))
