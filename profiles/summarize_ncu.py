"""profiles/summarize_ncu.py -- turn ncu artefacts brought back in gpurun_out/ into the small,
committed summaries under profiles/.

    python profiles/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches_summary.md
    python profiles/summarize_ncu.py full gpurun_out/prof_step_r1.ncu-rep profiles/r1_step_kernel_ncu.md [traffic.json]
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            agg[r[ki]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src}; gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| share | launches | mean us | kernel |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| {100 * sum(v) / tot:.2f}% | {len(v)} | {sum(v) / len(v) / 1000:.2f} | `{k[:110]}` |\n")


def full(src, dst, traffic_json=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for n, r in enumerate(rows[2:]):
            f.write(f"## launch {n}: `{r[idx['Kernel Name']][:100]}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
            f.write("\n")
    if traffic_json:
        def num(r, k):
            v = float(r[idx[k]].replace(",", ""))
            u = units[idx[k]].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        rs = rows[2:]
        rd = sum(num(r, "dram__bytes_read.sum") for r in rs) / len(rs)
        wr = sum(num(r, "dram__bytes_write.sum") for r in rs) / len(rs)
        json.dump({"kernel": rs[0][idx["Kernel Name"]][:60], "launches_averaged": len(rs),
                   "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr,
                   "dram_bytes_per_launch": rd + wr, "source": src,
                   "note": "ncu flushes caches before each replay: reads are the cold 320 B state + 32 B actions per "
                           "arena; the ~529 B/arena of writes stay in the 126 MB L2 within the profiled window"},
                  open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
