"""Digest of `ncu -i X.ncu-rep --page source --csv --print-source sass`: per kernel, the SASS lines with the most warp-stall
samples and a per-opcode total (debug aid for profiles/).   python tools/ncu_source_digest.py report.ncu-rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kern, hdr, rows = None, None, []


def flush():
    if not rows:
        return
    si, ii = hdr.index("# Samples"), hdr.index("Source")
    tot = sum(int(r[si] or 0) for r in rows)
    print(f"== {kern}: {tot} samples")
    by_op = {}
    for r in rows:
        op = r[ii].split()[0] if r[ii].split() else "?"
        if op.startswith("@"):
            op = r[ii].split()[1]
        by_op[op] = by_op.get(op, 0) + int(r[si] or 0)
    print("   per opcode:", ", ".join(f"{k} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(by_op.items(), key=lambda kv: -kv[1])[:14]))
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i][si] or 0))[:top]
    for i in sorted(order):
        print(f"   {100 * int(rows[i][si] or 0) / max(tot, 1):5.1f}%  [{i:5d}] {rows[i][ii].strip()}")


for rec in csv.reader(io.StringIO(txt)):
    if not rec:
        continue
    if rec[0] == "Kernel Name":
        flush()
        kern, hdr, rows = rec[1], None, []
    elif rec[0] == "Address":
        hdr = rec
    elif hdr is not None and len(rec) >= len(hdr) - 2:
        rows.append(rec)
flush()
