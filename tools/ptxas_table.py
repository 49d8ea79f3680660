"""Registers / stack / spills per kernel from an `nvcc -Xptxas -v` log.   python tools/ptxas_table.py log [filter]"""
import re
import subprocess
import sys

log = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for b in re.split(r"ptxas info\s+: Compiling entry function '", log)[1:]:
    name = b.split("'")[0]
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = dem.replace("nngp_fused::fused_loglik_kernel", "").replace("(EvalArgs)", "").replace("void ", "")
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", b)
    u = re.search(r"Used (\d+) registers", b)
    if flt and flt not in dem:
        continue
    print(f"{dem:60s} regs {u.group(1):>3s}  stack {m.group(1):>4s}  spill st/ld {m.group(2):>4s}/{m.group(3):>4s}")
