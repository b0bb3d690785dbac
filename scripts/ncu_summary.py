#!/usr/bin/env python
"""Summarise an ncu report of a step kernel into a tracked markdown file under profiles/.

  python scripts/ncu_summary.py gpurun_out/r2b_step_lean.ncu-rep profiles/r02_step_lean_ncu.md "title / command line"

Writes the launch-level counters (duration, registers, instructions, pipe utilisation, DRAM bytes, warp-stall
reasons per issue-active cycle) and the 20 SASS instructions with the most warp-stall samples (ncu source page)."""
import csv
import io
import re
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    rep, out, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    n_boards = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 20
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, data = rows[0], rows[1], rows[2:]

    def col(name):
        for i, h in enumerate(hdr):
            if h == name or h.endswith("." + name) or h.endswith(name):
                return i
        return None
    ki = hdr.index("Kernel Name")
    md = ["# ncu: %s" % (title or rep), "",
          "Report `%s` (`ncu --set full --clock-control none --import-source on`), %d captured launch(es) of `%s`.  ncu "
          "serialises launches, flushes caches between replays and disables programmatic dependent launch overlap: "
          "durations are cold-cache; compare shares and counters, not absolutes." % (rep, len(data), data[0][ki]), "",
          "| metric | unit | per launch |", "|---|---|---|"]
    for k in RAW:
        i = col(k)
        if i is not None:
            md.append("| `%s` | %s | %s |" % (k, units[i], " / ".join(r[i][:14] for r in data)))
    st = [(h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), num(data[-1][i]))
          for i, h in enumerate(hdr) if re.search(r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$", h)]
    st.sort(key=lambda x: -x[1])
    md += ["", "Warp stall reasons (warps per issue-active cycle, last launch): " +
           ", ".join("`%s` %.2f" % s for s in st[:10]), ""]
    ie, ai = col("smsp__inst_executed.sum"), None
    warps = None
    gi, bi = col("launch__grid_size"), col("launch__block_size")
    di, wi = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    md.append("Derived (%d boards per launch): %.1f warp-instructions per board-warp (= inst_executed / (boards / 32)); "
              "DRAM read %.1f B + write %.1f B per board (results of an isolated launch stay dirty in the 126 MB L2, so "
              "ncu attributes almost no DRAM writes to it)."
              % (n_boards, num(data[-1][ie]) / (n_boards / 32.0),
                 num(data[-1][di]) * scale.get(units[di], 1.0) / n_boards, num(data[-1][wi]) * scale.get(units[wi], 1.0) / n_boards))
    # source page: top stall instructions
    srows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    secs = [i for i, r in enumerate(srows) if r and r[0] == "Kernel Name"]
    if secs:
        s = secs[-1]
        h2 = srows[s + 1]
        d2 = [r for r in srows[s + 2:] if len(r) == len(h2)]
        ix = {h: i for i, h in enumerate(h2)}
        tot = sum(num(r[ix["# Samples"]]) for r in d2) or 1.0
        keys = ["stall_math", "stall_not_selected", "stall_dispatch", "stall_wait", "stall_long_sb", "stall_short_sb",
                "stall_barrier", "stall_no_inst", "stall_mio", "stall_branch_resolving", "stall_selected"]
        md += ["", "## Warp-stall sampling by SASS instruction (last captured launch, %d samples)" % tot, "",
               "Totals: " + ", ".join("`%s` %d" % (k[6:], sum(num(r[ix[k]]) for r in d2)) for k in keys if k in ix), "",
               "| addr | instruction | samples | % | math | not_sel | dispatch | wait | long_sb | short_sb |", "|---|---|---|---|---|---|---|---|---|---|"]
        for r in sorted(d2, key=lambda r: -num(r[ix["# Samples"]]))[:20]:
            md.append("| %s | `%s` | %d | %.1f | %d | %d | %d | %d | %d | %d |" % (
                r[ix["Address"]][-4:], r[ix["Source"]][:70], num(r[ix["# Samples"]]), 100 * num(r[ix["# Samples"]]) / tot,
                num(r[ix["stall_math"]]), num(r[ix["stall_not_selected"]]), num(r[ix["stall_dispatch"]]),
                num(r[ix["stall_wait"]]), num(r[ix["stall_long_sb"]]), num(r[ix["stall_short_sb"]])))
    open(out, "w").write("\n".join(md) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
