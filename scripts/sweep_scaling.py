"""Qubit-scaling sweep (BASELINE config 4; shape of the reference's scripts/reproduce_scaling.sh:28,56-80 and
reproduce_circuit.sh:33,55-68, extended to Q16): forward and forward+adjoint-grad throughput of the raw
circuit op per qubit count, with the tier the library picks and the roofline that bounds it.

    python scripts/sweep_scaling.py [--qubits 2 3 ...] [--dtype f32|f64] [--out profiles/sweep.jsonl]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quanonet_b200.ops import hea_expval, hea_expval_backward, plan_tier, fp32_peak_tflops

TIERS = {0: "register", 1: "shared", 2: "hbm"}

def timeit(fn, min_reps=3, budget_s=2.0):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); one = time.perf_counter() - t0
    reps = int(max(min_reps, min(50, budget_s / max(one, 1e-4))))
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, nargs="*", default=list(range(2, 17)))
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--net", type=int, nargs=4, default=[20, 2, 10, 2])
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dt = torch.float32 if args.dtype == "f32" else torch.float64
    dev = torch.device("cuda:0")
    peak = fp32_peak_tflops(4000)
    hbm = 6545.9
    rows = []
    for n in args.qubits:
        b_d, b_l, t_d, t_l = args.net
        K = b_d + t_d; depths = [t_l] * t_d + [b_l] * b_d; S = sum(depths)
        N = 1 << n; G = n * K + 3 * n * S
        f_fwd, f_all = 6 * N * G + 5 * N, 22 * N * G + 7 * N
        tier, lq = plan_tier(1 << 20, n, dt, True)
        # batch: ~2^25 amplitudes in flight for the register tier, fewer for the CTA-per-sample tiers
        if tier == 0: B = max(4096, min(1 << 20, (1 << 25) >> n))
        elif tier == 1: B = max(2048, min(1 << 20, (1 << 25) >> n))
        else: B = max(64, 1184 >> max(0, n - 16))
        g = torch.Generator().manual_seed(n)
        x = ((torch.rand(B, n * K, generator=g) * 2 - 1) * np.pi).to(dev, dt)
        w = ((torch.rand(S, 3, n, generator=g) * 2 - 1) * np.pi).to(dev, dt)
        go = torch.randn(B, generator=g).to(dev, dt)
        off, co = 0.0, 5.0 / n
        t_f = timeit(lambda: hea_expval(x, w, n, depths, None, 0, off, co, 0))
        t_g = timeit(lambda: hea_expval_backward(go, x, w, n, depths, None, 0, off, co, 0, True))
        es = 8 if dt == torch.float32 else 16
        stream_bytes_fwd = es * 2 * N * (K + S)          # one read+write of the state per layer
        row = {"n": n, "dtype": args.dtype, "net": args.net, "tier": TIERS[tier], "lanes_per_sample": (1 << lq) if tier == 0 else None,
               "B": B, "fwd_ms": t_f, "grad_ms": t_g, "fwd_samples_per_s": B / t_f * 1e3, "grad_samples_per_s": B / t_g * 1e3,
               "fwd_tflops_alg": f_fwd * B / t_f / 1e9, "grad_tflops_alg": f_all * B / t_g / 1e9,
               "fwd_frac_fp32_peak": f_fwd * B / t_f / 1e9 / peak, "grad_frac_fp32_peak": f_all * B / t_g / 1e9 / peak,
               "fwd_stream_gbs": stream_bytes_fwd * B / t_f / 1e6, "fwd_stream_frac_hbm": stream_bytes_fwd * B / t_f / 1e6 / hbm}
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            for r in rows: f.write(json.dumps(r) + "\n")

if __name__ == "__main__":
    main()
