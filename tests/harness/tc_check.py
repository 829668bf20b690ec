"""GPU check of the tensor-core tier (csrc/hea_tc.cuh): parity vs the fp64 oracle and vs the FFMA2 register
kernel, intermediate-state dump on mismatch, and forward throughput of both at B = 1M.
    python tests/harness/tc_check.py [--quick]
"""
import ctypes, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hea_oracle as orc
from quanonet_b200 import _lib
from quanonet_b200.ops import hea_expval
sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))
import tc_emulate as emu

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
lib = _lib.load()
dev = torch.device("cuda:0")
n = 5
results = {}


def run(x, w, depths, tc, dbg=None, err=None, min_b=0):
    lib.qon_tensor_tier(int(tc), min_b, None if dbg is None else dbg.data_ptr(), None if err is None else err.data_ptr())
    out = hea_expval(x, w, n, list(depths), None, 0, 0.0, 1.0, 0)
    torch.cuda.synchronize()
    return out[:, 0].double().cpu().numpy()


def case(depths, B, seed, dump=False):
    rng = np.random.default_rng(seed)
    K, S = len(depths), sum(depths)
    x = rng.uniform(-np.pi, np.pi, (B, n * K))
    w = rng.uniform(-np.pi, np.pi, (S, 3, n))
    blocks = [(n, d) for d in depths]
    nref = min(B, 512)
    ref = orc.hea_forward(x[:nref], w, n, blocks, orc.ham_from_bound(n))
    xt = torch.tensor(x, dtype=torch.float32, device=dev)
    wt = torch.tensor(w, dtype=torch.float32, device=dev)
    dbg = torch.zeros(K * 128 * 128, dtype=torch.float32, device=dev) if dump else None
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    o_tc = run(xt, wt, depths, True, dbg, err)
    o_reg = run(xt, wt, depths, False)
    e_tc = float(np.linalg.norm(o_tc[:nref] - ref) / np.linalg.norm(ref))
    e_reg = float(np.linalg.norm(o_reg[:nref] - ref) / np.linalg.norm(ref))
    e_x = float(np.linalg.norm(o_tc - o_reg) / np.linalg.norm(o_reg))
    flag = int(err.item())
    print(f"K={K} S={S} B={B}: tc vs oracle {e_tc:.2e} | ffma2 vs oracle {e_reg:.2e} | tc vs ffma2 (all rows) {e_x:.2e} | err flag {flag}",
          flush=True)
    results[f"K{K}_S{S}_B{B}"] = dict(tc_vs_oracle=e_tc, ffma2_vs_oracle=e_reg, tc_vs_ffma2=e_x, err=flag)
    if dump:
        d = dbg.cpu().numpy().reshape(K, 128, 128)[:, :, :64]
        # expected scaled state after block k's GEMM, exact arithmetic
        s0 = 0
        exp = np.zeros((K, min(B, 128), 64))
        for b in range(min(B, 128)):
            amp = np.full(32, emu.SA / np.sqrt(32), complex)
            s0 = 0
            for k, dd in enumerate(depths):
                th = x[b, k * n:(k + 1) * n]
                ph = np.array([np.prod([np.exp((-1j if not (z >> q) & 1 else 1j) * th[q] / 2) for q in range(n)]) for z in range(32)])
                M = emu.block_matrix(w, s0, dd, k == K - 1)
                s0 += dd
                amp = M @ (amp * ph * (1.0 if k == 0 else 1.0 / emu.SB)) * emu.SB
                exp[k, b, 0::2], exp[k, b, 1::2] = amp.real, amp.imag
        nb = min(B, 128)
        for k in range(K):
            den = np.linalg.norm(exp[k])
            print(f"   block {k}: D dump vs expected rel {np.linalg.norm(d[k, :nb] - exp[k]) / den:.3e}", flush=True)
        np.savez(os.path.join(OUT, f"tc_dbg_K{K}.npz"), dbg=d, exp=exp, x=x, w=w)
    return e_tc


def bench(B=1_000_000, depths=(2,) * 60, iters=5):
    K, S = len(depths), sum(depths)
    g = torch.Generator(device="cpu").manual_seed(0)
    x = (torch.rand(B, n * K, generator=g) * 2 - 1).mul_(np.pi).to(dev)
    w = (torch.rand(S, 3, n, generator=g) * 2 - 1).mul_(np.pi).to(dev)
    res = {}
    for name, tc in (("ffma2", False), ("tc", True)):
        lib.qon_tensor_tier(int(tc), 0, None, None)
        for _ in range(2):
            hea_expval(x, w, n, list(depths), None, 0, 0.0, 1.0, 0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            o = hea_expval(x, w, n, list(depths), None, 0, 0.0, 1.0, 0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res[name] = dict(ms=ms, samples_per_s=B / ms * 1e3, checksum=float(o.double().sum()))
        print(f"forward B={B} K={K}: {name}: {ms:.3f} ms  {B / ms * 1e3:.3e} samples/s", flush=True)
    results["bench_fwd"] = res


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    ok = True
    ok &= case([1], 128, 1, dump=True) < 1e-5
    ok &= case([2, 1], 128, 2, dump=True) < 1e-5
    if ok or not quick:
        case([2] * 6, 1000, 3)
        case([2] * 60, 1000, 4)
        case([2] * 60, 70001, 5)
        case([1, 3, 2] * 7, 5000, 6)
        bench()
    json.dump(results, open(os.path.join(OUT, "tc_check.json"), "w"), indent=1)
