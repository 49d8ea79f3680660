"""Shared-memory wavefronts per instruction and per source line from an ncu report (the source page's
"L1 Wavefronts Shared" / "... Ideal" columns): which accesses load the shared-memory data stage.

    python tools/ncu_smem.py report.ncu-rep [kernel-substring] [top=40]
Per warp-iteration figures divide by `warp-iterations` = executions of the hottest shared-memory instruction.
"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
k, done = 0, set()
while k < len(rows):
    r = rows[k]
    if r and r[0] in ("Kernel Name", "Function Name"):
        name = r[1]
    if r and ("Address" in r[:2] or r[0] == "Line No") and "L1 Wavefronts Shared" in r:
        hdr = r
        body = []
        k += 1
        while k < len(rows) and rows[k] and rows[k][0] not in ("Kernel Name", "Function Name", "File Path", "Line No", "Address"):
            body.append(rows[k])
            k += 1
        if (want and want not in name) or name in done:
            continue
        col = {h: i for i, h in reversed(list(enumerate(hdr)))}
        iw, ii, ie = col["L1 Wavefronts Shared"], col["L1 Wavefronts Shared Ideal"], col["Instructions Executed"]
        isrc = col["Source"]
        isass = 3 if hdr[0] == "Line No" else isrc
        ops, cur_line = defaultdict(lambda: [0, 0, 0, 0]), ""
        per_line = defaultdict(lambda: [0, 0, 0, ""])
        seen_addr = set()
        for b in body:
            if len(b) <= iw:
                continue
            is_cuda = hdr[0] == "Line No" and b[0] != ""
            if is_cuda:
                cur_line = f"L{b[0]}: {b[isrc].strip()[:90]}"
                continue
            try:
                w, idl, ex = int(b[iw]), int(b[ii]), int(b[ie])
            except ValueError:
                continue
            if w == 0:
                continue
            txt = b[isass].split()
            op = txt[1] if txt[0].startswith("@") else txt[0]
            addr = b[2] if hdr[0] == "Line No" else b[0]
            if addr not in seen_addr:  # an inlined instruction is listed under every line of its call stack
                seen_addr.add(addr)
                o = ops[op]
                o[0] += w; o[1] += idl; o[2] += ex; o[3] += 1
            pl = per_line[cur_line]
            pl[0] += w; pl[1] += idl; pl[2] += ex
        if not ops:
            continue
        done.add(name)
        iters = max(v[2] / max(v[3], 1) for v in ops.values())  # executions of one instruction ~ warp-iterations
        # (per-line figures below are inclusive: an instruction inlined from a helper counts for the helper's line AND the call site)
        tot = sum(v[0] for v in ops.values())
        print(f"== {name[:120]}\n   shared wavefronts {tot}  (~{tot / iters:.0f} per warp-iteration, {iters:.0f} iterations)")
        for op, v in sorted(ops.items(), key=lambda t: -t[1][0]):
            print(f"   {op:16s} static {v[3]:4d}  wavefronts/iter {v[0] / iters:7.1f}  ideal/iter {v[1] / iters:7.1f}  per executed instr {v[0] / max(v[2], 1):5.2f}")
        if hdr[0] == "Line No":
            for ln, v in sorted(per_line.items(), key=lambda t: -t[1][0])[:top]:
                print(f"   {v[0] / iters:7.1f} (ideal {v[1] / iters:6.1f}, {v[0] / max(v[2], 1):5.2f}/instr)  {ln}")
    else:
        k += 1
