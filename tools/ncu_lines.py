"""Per-source-line stall samples / executed instructions from an ncu report.

    python tools/ncu_lines.py report.ncu-rep [top=40] [kernel-substring]
(uses `ncu -i report --page source --csv --print-source cuda,sass`; needs -lineinfo builds)
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# split per kernel ("Function Name" rows precede each header)
blocks, name = [], ""
for k, r in enumerate(rows):
    if r and r[0] == "Function Name":
        name = r[1]
    if r and r[0] == "Line No":
        blocks.append((name, k))
for bi, (name, k) in enumerate(blocks):
    if want and want not in name:
        continue
    end = blocks[bi + 1][1] - 3 if bi + 1 < len(blocks) else len(rows)
    hdr = rows[k]
    isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")

    def iv(r, i):
        try:
            return int(r[i])
        except (ValueError, IndexError):
            return 0

    lines = [r for r in rows[k + 1:end] if r and r[0] != "" and len(r) > iex]
    tot = sum(iv(r, isamp) for r in lines) or 1
    totex = sum(iv(r, iex) for r in lines) or 1
    print(f"== {name[:110]}  samples {tot} warp-instr {totex}")
    for r in sorted(lines, key=lambda r: -iv(r, isamp))[:top]:
        print(f"L{r[0]:>4s} samp {iv(r, isamp) / tot * 100:5.1f}% exec {iv(r, iex) / totex * 100:5.1f}%  {r[1].strip()[:100]}")
