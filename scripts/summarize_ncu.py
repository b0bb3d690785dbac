#!/usr/bin/env python
"""Turn gpurun_out/{step_full.ncu-rep,launches.csv,bench.json} into tracked files under profiles/.

  python scripts/summarize_ncu.py r01          # writes profiles/r01_*.{md,csv,json}
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    rep = os.path.join(OUT, "step_full.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    md = ["# %s — ncu capture of the step kernel (B200, sm_100a)" % tag, "",
          "Source: `gpurun_out/step_full.ncu-rep` (`ncu --set full --clock-control none --import-source on "
          "-k regex:g2048_step_kernel -s 60 -c 3 python bench.py --steps 100 --warmup 20 --e2e-steps 3 --fused-steps 0`), summarised by "
          "`scripts/summarize_ncu.py`.  Kernel: `%s`, %d captured launches of 1,048,576 boards each.  ncu "
          "serialises launches and flushes caches between replays, so durations are cold-cache and exclude the "
          "programmatic-dependent-launch overlap of the real run; compare shares, not absolutes." %
          (data[0][ki], len(data)), "", "| metric | unit | per launch |", "|---|---|---|"]
    vals = {}
    for k in RAW_KEYS:
        if k in hdr:
            i = hdr.index(k)
            v = [r[i] for r in data]
            vals[k] = v
            md.append("| `%s` | %s | %s |" % (k, units[i], " / ".join(v)))
    stalls = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                stalls.append((float(data[-1][i]), h))
            except ValueError:
                pass
    md += ["", "Top warp stall reasons (warps per issue-active cycle, last launch):", ""]
    for v, h in sorted(stalls, reverse=True)[:8]:
        md.append("- `%s`: %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    n = 1 << 20
    rd = float(vals["dram__bytes_read.sum"][-1]) * (1e6 if "Mbyte" in units[hdr.index("dram__bytes_read.sum")] else 1)
    wu = units[hdr.index("dram__bytes_write.sum")]
    wr = float(vals["dram__bytes_write.sum"][-1]) * (1e6 if "Mbyte" in wu else 1e3 if "Kbyte" in wu else 1)
    inst = float(vals["smsp__inst_executed.sum"][-1])
    md += ["", "Derived:", "",
           "- warp instructions per board-step: %.1f (= %d / 32768 warps)" % (inst / (n / 32), inst),
           "- DRAM traffic per launch: %.2f MB read + %.2f MB written = %.1f B per board (algorithmic 38 B: 17 B "
           "in, 21 B out; the 21 MB of results are still dirty in the 126 MB L2 when the isolated launch ends and "
           "are written back later, so ncu attributes no DRAM writes to it)" % (rd / 1e6, wr / 1e6, (rd + wr) / n),
           "- tensor pipe: 0 % (pure integer path)"]
    with open(os.path.join(PROF, "%s_step_kernel_ncu.md" % tag), "w") as f:
        f.write("\n".join(md) + "\n")
    with open(os.path.join(PROF, "step_kernel_traffic.json"), "w") as f:
        json.dump({"source": "%s_step_kernel_ncu.md" % tag, "boards_per_launch": n,
                   "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr,
                   "dram_bytes_per_board": (rd + wr) / n}, f, indent=1)
    # launch list
    lrows = list(csv.reader(open(os.path.join(OUT, "launches.csv"))))
    hi = [i for i, r in enumerate(lrows) if r and r[0] == "ID"][0]
    lh, ld = lrows[hi], lrows[hi + 1:]
    kn, mv = lh.index("Kernel Name"), lh.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in ld:
        if len(r) > mv:
            agg[r[kn]].append(float(r[mv].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(PROF, "%s_launches_summary.md" % tag), "w") as f:
        f.write("# " + tag + " — ncu launch list of `python bench.py --steps 100 --warmup 20 --e2e-steps 3 --fused-steps 0` (gpu__time_duration.sum, ns)\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200`; per-launch times are "
                "cold-cache and serialised.  The timed region of bench.py launches only `g2048_step_kernel<0, 0>` (lean outputs, host-side step index); "
                "`<1, 0>` and the reset kernel belong to the e2e leg and set-up.\n\n"
                "| kernel | launches | mean ns | share of listed GPU time |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %.0f | %.3f |\n" % (k, len(v), sum(v) / len(v), sum(v) / tot))
    with open(os.path.join(PROF, "%s_launches.csv" % tag), "w") as f:
        f.write(open(os.path.join(OUT, "launches.csv")).read())
    bj = os.path.join(OUT, "bench.json")
    if os.path.exists(bj):
        with open(os.path.join(PROF, "%s_bench_line.json" % tag), "w") as f:
            f.write(open(bj).read())
    print("\n".join(md[-12:]))


if __name__ == "__main__":
    main()
