"""Summarise an `ncu --page source --csv` export: stall samples per block of SASS instructions."""
import collections
import csv
import sys

path = sys.argv[1]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 250
rows = list(csv.reader(open(path)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = blocks[-1]
hdr = b["rows"][0]
data = [r for r in b["rows"][1:] if len(r) == len(hdr)]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp]) for r in data)
print(b["name"][:90], "instrs", len(data), "samples", tot)
allst = collections.Counter()
for start in range(0, len(data), step):
    seg = data[start:start + step]
    s = sum(int(r[isamp]) for r in seg)
    if s == 0:
        continue
    e = sum(int(r[iex]) for r in seg) / len(seg)
    st = collections.Counter()
    for r in seg:
        for h in stalls:
            v = r[hdr.index(h)]
            if v:
                st[h] += int(v)
    allst.update(st)
    top = ", ".join(f"{k[6:]}={v}" for k, v in st.most_common(5))
    print(f"{start:5d} samples={s:6d} ({100*s/tot:4.1f}%) avgexec={e:9.0f} {top}")
print("total:", ", ".join(f"{k[6:]}={v} ({100*v/tot:.1f}%)" for k, v in allst.most_common(10)))
if len(sys.argv) > 3:  # dump the hottest instructions
    top = sorted(data, key=lambda r: -int(r[isamp]))[: int(sys.argv[3])]
    for r in top:
        print(r[isamp], r[ia][-5:], r[isrc][:90])
