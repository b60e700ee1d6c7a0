"""profiles/stall_by_address.py <report.ncu-rep> -- where the `no_instruction` samples of the step kernel sit.
Reads the source page of an `ncu --set full --import-source on` capture (`ncu -i REP --page source --csv --print-source sass`)
and reports, for the first kernel of the capture: the stall mix, and the share of stall_no_inst samples that fall on the first
instruction of a 128-byte instruction-cache line, on the instruction after a taken-branch / call / return, and elsewhere."""
import collections
import csv
import subprocess
import sys


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    name = rows[heads[0] - 1][1] if heads[0] > 0 else "?"
    h = rows[heads[0]]
    data = [r for r in rows[heads[0] + 1:(heads[1] - 1 if len(heads) > 1 else len(rows))] if len(r) >= len(h)]
    col = {n: i for i, n in enumerate(h)}
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = collections.Counter({s: sum(int(r[col[s]] or 0) for r in data) for s in stalls})
    n_all = sum(tot.values())
    print(f"kernel: {name[:90]}")
    print(f"{len(data)} SASS instructions ({len(data) * 16 // 1024} KB), {sum(int(r[col['Instructions Executed']] or 0) for r in data)} warp instructions executed, "
          f"{n_all} stall samples")
    print("stall mix: " + ", ".join(f"{s[6:]} {100 * v / n_all:.1f}%" for s, v in tot.most_common(8)))
    where = collections.Counter()
    prev = ""
    executed_lines = set()
    for r in data:
        addr = int(r[col["Address"]], 16)
        src = r[col["Source"]].split()
        op = (src[1] if src and src[0].startswith("@") else (src[0] if src else "")).split(".")[0]
        n = int(r[col["stall_no_inst"]] or 0)
        if int(r[col["Instructions Executed"]] or 0) > 0:
            executed_lines.add(addr // 128)
        if addr % 128 == 0:
            where["first instruction of a 128 B line"] += n
        elif prev in ("BRA", "CALL", "RET", "BSYNC", "EXIT", "BRX", "JMP") or op == "BSYNC":
            where["after a branch / call / return or at a reconvergence point"] += n
        else:
            where["inside a line, straight-line code"] += n
        prev = op
    n_ni = sum(where.values())
    for k, v in where.most_common():
        print(f"  no_instruction: {k}: {v} samples = {100 * v / max(n_ni, 1):.0f}%")
    print(f"128 B instruction lines with executed instructions: {len(executed_lines)} = {len(executed_lines) * 128 // 1024} KB "
          f"(L1.5 instruction cache: 32 KB, L0: ~6 KB per SM sub-partition)")


if __name__ == "__main__":
    main(sys.argv[1])
