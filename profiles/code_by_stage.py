"""profiles/code_by_stage.py <lib.so> [report.ncu-rep] -- static code size (and, with an ncu capture, executed warp instructions
and stall samples) of step_kernel_v4<3,0> per call site in v4::step_body, i.e. per stage / role, and per out-of-line routine.
Uses `cuobjdump -xelf` + `nvdisasm -gi` (the library is built with -lineinfo) and `ncu --page source --csv`.
The library must be the build the capture was taken from (instructions are joined by their offset in the kernel); the
sanity check is that the `executed` column sums to the capture's warp-instruction count."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

KERNEL = "_ZN2hh14step_kernel_v4ILi3ELi0EEEvNS_9StatePtrsENS_6ParamsEPKiPfS5_S5_Ph"


def disassemble(lib):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.startswith("hh_api.") and f.endswith(".cubin")][0]
    return subprocess.run(["nvdisasm", "-gi", os.path.join(d, cubin)], capture_output=True, text=True).stdout.splitlines()


def stage_of_line(src_lines, ln):
    """The statement of step_body at line `ln` (HH_ROLE(...) / function call), shortened."""
    s = src_lines[ln - 1].strip() if 0 < ln <= len(src_lines) else "?"
    return f"{ln}: {s[:86]}"


def main(lib, rep=None):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    v4 = open(os.path.join(root, "hhmarl_2d_b200", "csrc", "hh_v4.cuh")).read().splitlines()
    lines = disassemble(lib)
    start = next(i for i, l in enumerate(lines) if l.strip().startswith(".section") and ".text." + KERNEL in l)
    per_addr = {}
    chain, sub, last_main = [], "(main body)", "(main body)"
    pat = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
    ins = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);")
    for l in lines[start + 1:]:
        if l.strip().startswith(".section"):
            break
        m = pat.search(l)
        if m:
            chain.append((os.path.basename(m.group(1)), int(m.group(2)), os.path.basename(m.group(3) or ""), int(m.group(4) or 0)))
            continue
        if l.startswith("$") or (l.strip().endswith(":") and "$" in l and "_ZN" in l):
            name = l.strip().rstrip(":").split("$")[-1]
            sub = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0][-60:]
            continue
        m = ins.match(l)
        if m:
            off = int(m.group(1), 16)
            key = sub
            if sub == "(main body)":
                site = [c for c in chain if c[0] == "hh_v4.cuh" and c[2] == "hh_api.cu"]
                if site:
                    key = stage_of_line(v4, site[0][1])
                elif chain and chain[-1][0] == "hh_api.cu":
                    key = "(kernel prologue / epilogue, hh_api.cu)"
                elif not chain:
                    key = last_main            # nvdisasm repeats the location only when it changes
                last_main = key
            per_addr[off] = key
            chain = []
    stat = collections.defaultdict(lambda: [0, 0, 0, 0])   # static instrs, executed, no_inst, all stalls
    for off, key in per_addr.items():
        stat[key][0] += 1
    if rep:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
        h = rows[heads[0]]
        col = {n: i for i, n in enumerate(h)}
        data = [r for r in rows[heads[0] + 1:(heads[1] - 1 if len(heads) > 1 else len(rows))] if len(r) >= len(h)]
        base = int(data[0][col["Address"]], 16)
        stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        for r in data:
            key = per_addr.get(int(r[col["Address"]], 16) - base)
            if key is None:
                continue
            stat[key][1] += int(r[col["Instructions Executed"]] or 0)
            stat[key][2] += int(r[col["stall_no_inst"]] or 0)
            stat[key][3] += sum(int(r[col[s]] or 0) for s in stalls)
    tot = [sum(v[k] for v in stat.values()) for k in range(4)]
    print(f"{'call site in v4::step_body / out-of-line routine':100s} {'instr':>6s} {'KB':>6s} {'executed':>9s} {'no_inst':>7s} {'stalls':>6s}")
    for key, v in sorted(stat.items(), key=lambda kv: -kv[1][0]):
        print(f"{key:100s} {v[0]:6d} {v[0] * 16 / 1024:6.1f} {v[1]:9d} {v[2]:7d} {v[3]:6d}")
    print(f"{'total':100s} {tot[0]:6d} {tot[0] * 16 / 1024:6.1f} {tot[1]:9d} {tot[2]:7d} {tot[3]:6d}")


if __name__ == "__main__":
    main(*sys.argv[1:3])
