"""Per-source-line stall REASONS from an ncu report (the source page's stall_* columns) and the SASS size of the
kernel: which lines wait on what.

    python tools/ncu_stalls.py report.ncu-rep [top=30] [kernel-substring]
"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
want = sys.argv[3] if len(sys.argv) > 3 else ""


def page(src):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", src], capture_output=True, text=True).stdout
    return list(csv.reader(raw.splitlines()))


# SASS size per kernel
sass_count = {}
rows = page("sass")
k = 0
while k < len(rows):
    if rows[k] and rows[k][0] == "Kernel Name":
        name = rows[k][1]
        k += 2
        cnt = 0
        while k < len(rows) and not (rows[k] and rows[k][0] == "Kernel Name"):
            if rows[k] and rows[k][0].startswith("0x"):
                cnt += 1
            k += 1
        sass_count.setdefault(name, cnt)
    else:
        k += 1

rows = page("cuda,sass")
k, seen, name = 0, set(), ""
while k < len(rows):
    r = rows[k]
    if r and r[0] == "Function Name":
        name = r[1]
    if r and r[0] == "Line No":
        hdr = r
        body = []
        k += 1
        while k < len(rows) and rows[k] and rows[k][0] not in ("File Path", "Function Name", "Line No"):
            body.append(rows[k])
            k += 1
        if (want and want not in name):
            continue
        cols = {}
        for i, h in enumerate(hdr):
            cols.setdefault(h, i)
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        total = defaultdict(int)
        lines = []
        for b in body:
            if len(b) < len(hdr) or b[0] == "":
                continue  # SASS rows (no line number): their samples are already summed into the CUDA row above
            d = {}
            for s in stalls:
                try:
                    d[s] = int(b[cols[s]])
                except ValueError:
                    d[s] = 0
                total[s] += d[s]
            lines.append((b[0], b[1], d))
        allsamp = sum(total.values()) or 1
        if allsamp < 50:
            continue
        ns = sass_count.get(name, 0)
        print(f"== {name[:130]}\n   SASS instructions {ns} ({ns * 16 / 1024:.1f} KB)  samples {allsamp}")
        print("   totals: " + ", ".join(f"{s[6:]}={total[s] / allsamp * 100:.1f}%" for s in sorted(stalls, key=lambda s: -total[s])[:9]))
        for ln, src, d in sorted(lines, key=lambda t: -sum(t[2].values()))[:top]:
            tot = sum(d.values())
            topst = sorted(((v, s) for s, v in d.items() if v), reverse=True)[:4]
            print(f"   {tot / allsamp * 100:5.1f}%  " + " ".join(f"{s[6:]}:{v / allsamp * 100:.1f}" for v, s in topst) + f"  | L{ln}: {src.strip()[:80]}")
    else:
        k += 1
