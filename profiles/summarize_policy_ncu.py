"""profiles/summarize_policy_ncu.py REP OUT.md [traffic.json] -- key counters of an `ncu --set full` capture of the tcgen05 policy
forward (tensor pipe, issue slots, DRAM / L2 traffic, stall mix)."""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
STALLS = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]


def val(r, k):
    v, u = r[idx[k]].replace(",", ""), units[idx[k]]
    try:
        f = float(v)
    except ValueError:
        return v
    return f * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Kinst": 1e3, "Minst": 1e6}.get(u, 1)


with open(out, "w") as f:
    f.write(f"# ncu --set full: {data[0][idx['Kernel Name']][:80]} ({rep.split('/')[-1]}, {len(data)} launches, --clock-control none)\n\n")
    f.write("| metric | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n|---|" + "---|" * len(data) + "\n")
    for k in KEYS:
        if k in idx:
            f.write(f"| `{k}` [{units[idx[k]]}] | " + " | ".join(r[idx[k]] for r in data) + " |\n")
    f.write("\nStalled warps per issue-active cycle (largest first, launch 0):\n\n")
    st = sorted(((float(data[0][idx[h]]), h.split("stalled_")[1].split("_per_issue")[0]) for h in STALLS if data[0][idx[h]] not in ("", "n/a")), reverse=True)
    for v, n in st[:8]:
        f.write(f"* {n}: {v:.2f}\n")
if len(sys.argv) > 3:
    rd = sum(val(r, "dram__bytes_read.sum") for r in data) / len(data)
    wr = sum(val(r, "dram__bytes_write.sum") for r in data) / len(data)
    json.dump({"kernel": data[0][idx["Kernel Name"]][:60], "launches_averaged": len(data), "dram_bytes_read_per_launch": rd,
               "dram_bytes_write_per_launch": wr, "dram_bytes_per_launch": rd + wr,
               "tensor_pipe_active_pct": sum(val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") for r in data) / len(data),
               "issue_active_pct": sum(val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") for r in data) / len(data),
               "source": rep.split("/")[-1],
               "note": "8 192 rows x 4 chains; reads = the 5 MB of packed weight images + inputs (everything else hits the 126 MB L2); "
                       "outputs stay in L2 within the profiled window"}, open(sys.argv[3], "w"), indent=1)
