"""Instruction mix of one kernel's SASS (static count over the whole function).
    python tools/sass_mix.py obj-or-so 'substring of demangled template args, e.g. <double, 4, 4, 1, false, 2, 2, 1, false, true>'"""
import collections
import re
import subprocess
import sys

obj, want = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, counts = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        counts[cur][m.group(2)] += 1
for name, c in counts.items():
    if want in name:
        tot = sum(c.values())
        fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
        print(f"{name[:140]}\n  total {tot}  fp64 {fp64}  " + " ".join(f"{k}:{v}" for k, v in c.most_common(18)))
