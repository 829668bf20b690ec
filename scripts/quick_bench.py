"""Scratch timing of the raw ops (not the contract bench): python scripts/quick_bench.py [B] [n]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quanonet_b200.ops import hea_expval, hea_expval_backward, fp32_peak_tflops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5
net = (40, 2, 20, 2)
K = net[0] + net[2]; depths = [net[3]] * net[2] + [net[1]] * net[0]; S = sum(depths)
dev = torch.device("cuda:0")
x = (torch.rand(B, n * K, device=dev) * 2 - 1) * np.pi
w = (torch.rand(S, 3, n, device=dev) * 2 - 1) * np.pi
g = torch.randn(B, device=dev)
off, co = 0.0, 5.0 / n
N = 1 << n; G = n * K + 3 * n * S
f_fwd = 6 * N * G + 5 * N; f_all = 22 * N * G + 7 * N
print("fp32 peak probe TFLOP/s:", fp32_peak_tflops(4000))
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t = timeit(lambda: hea_expval(x, w, n, depths, None, 0, off, co, 0))
print(f"forward        : {t:8.3f} ms  {B / t * 1e3:.3e} samples/s  {f_fwd * B / t / 1e9:.2f} TFLOP/s(alg)")
t = timeit(lambda: hea_expval_backward(g, x, w, n, depths, None, 0, off, co, 0, True))
print(f"fwd+grad (gx)  : {t:8.3f} ms  {B / t * 1e3:.3e} samples/s  {f_all * B / t / 1e9:.2f} TFLOP/s(alg)")
t = timeit(lambda: hea_expval_backward(g, x, w, n, depths, None, 0, off, co, 0, False))
print(f"fwd+grad (nogx): {t:8.3f} ms  {B / t * 1e3:.3e} samples/s  {f_all * B / t / 1e9:.2f} TFLOP/s(alg)")
