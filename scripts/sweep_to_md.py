"""jsonl written by scripts/sweep_scaling.py -> markdown table (profiles/)."""
import json, sys
rows = [json.loads(l) for f in sys.argv[1:] for l in open(f) if l.strip().startswith("{")]
print("| n | dtype | tier | B | fwd samples/s | fwd alg TFLOP/s | % FP32 peak | fwd+grad samples/s | alg TFLOP/s | % FP32 peak | fwd stream GB/s | % HBM |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    tier = r["tier"] + (f" ({r['lanes_per_sample']} lane/sample)" if r.get("lanes_per_sample") else "")
    print(f"| {r['n']} | {r['dtype']} | {tier} | {r['B']} | {r['fwd_samples_per_s']:.3e} | {r['fwd_tflops_alg']:.2f} | "
          f"{100 * r['fwd_frac_fp32_peak']:.1f} | {r['grad_samples_per_s']:.3e} | {r['grad_tflops_alg']:.2f} | "
          f"{100 * r['grad_frac_fp32_peak']:.1f} | {r['fwd_stream_gbs']:.0f} | {100 * r['fwd_stream_frac_hbm']:.1f} |")
