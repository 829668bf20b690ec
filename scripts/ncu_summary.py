#!/usr/bin/env python
"""Turn an ncu report (+ optional launch-list csv) into the markdown summary kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [gpurun_out/launches.csv] > profiles/NAME.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import Counter, defaultdict

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__icc_request_hit_rate.pct", "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active",
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio")


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    print(f"# ncu summary of `{rep.split('/')[-1]}`\n")
    print("Captured with `ncu --set full --clock-control none --import-source on` on one B200 "
          "(numbers under the profiler are not bench values; they explain them).\n")
    names = [re.sub(r"\(.*", "", r[kn]) for r in data]
    print("| metric | unit | " + " | ".join(f"`{n[-48:]}`" for n in names) + " |")
    print("|---|---|" + "---|" * len(names))
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print(f"| {m} | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
    print("\nWarp stall reasons (warps per issue-active cycle):\n")
    print("| stall | " + " | ".join(f"`{n[-48:]}`" for n in names) + " |")
    print("|---|" + "---|" * len(names))
    for i, h in enumerate(hdr):
        mm = STALL.match(h)
        if mm and any(float(r[i] or 0) > 0.005 for r in data):
            print(f"| {mm.group(1)} | " + " | ".join(f"{float(r[i] or 0):.3f}" for r in data) + " |")
    # source page: opcode mix and sample share per stall reason
    src = ncu_csv(rep, "source")
    kern, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            kern.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    seen = set()
    for k in kern:
        if k["name"] in seen or not k["rows"]:
            continue
        seen.add(k["name"])
        h, d = k["rows"][0], k["rows"][1:]
        col = h.index
        ops, samples = Counter(), Counter()
        stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
        tot = Counter()
        for r in d:
            s = r[col("Source")].split()
            if not s:
                continue
            op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
            ops[op] += 1
            samples[op] += int(r[col("# Samples")] or 0)
            for st in stalls:
                tot[st] += int(r[col(st)] or 0)
        T = sum(tot.values()) or 1
        print(f"\n### `{k['name'][:110]}`\n")
        print(f"SASS instructions: {len(d)} ({len(d) * 16 / 1024:.0f} KB). Opcode mix (static count / sampled share): "
              + ", ".join(f"{o} {c} / {100 * samples[o] / max(sum(samples.values()), 1):.0f}%" for o, c in ops.most_common(8)))
        print("\nSample share by stall reason: "
              + ", ".join(f"{s[6:]} {100 * v / T:.1f}%" for s, v in tot.most_common() if v * 200 > T))
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2])))
        h = None
        agg = defaultdict(lambda: [0, 0.0])
        for r in rows:
            if "Kernel Name" in r:
                h = r
                continue
            if h and len(r) == len(h):
                d = dict(zip(h, r))
                k = re.sub(r"\(.*", "", d["Kernel Name"])[-70:]
                agg[k][0] += 1
                agg[k][1] += float(d["Metric Value"].replace(",", ""))
        tot = sum(v[1] for v in agg.values()) or 1
        print("\n## Launch list (`ncu --metrics gpu__time_duration.sum`, cold-cache, serialised — compare shares)\n")
        print("| kernel | launches | total ms | share |\n|---|---|---|---|")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / tot:.1f}% |")


if __name__ == "__main__":
    main()
