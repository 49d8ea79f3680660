"""Turn an ncu report (gpurun_out/*.ncu-rep) into the small text summaries kept under profiles/."""
import csv
import json
import subprocess
import sys

rep, out_prefix = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
summ = []
with open(out_prefix + "_summary.txt", "w") as f:
    f.write(f"# from {rep} (ncu --set full --clock-control none); one block per profiled launch\n")
    for r in rows[2:]:
        d = {}
        for k in keys:
            if k in hdr:
                d[k] = (r[hdr.index(k)], units[hdr.index(k)])
        f.write("\n")
        for k, (v, u) in d.items():
            f.write(f"{k} = {v} {u}\n")
        st = sorted(((float(r[hdr.index(h)] or 0), h) for h in stall), reverse=True)[:8]
        f.write("stall cycles per issued instruction: " + ", ".join(
            f"{h.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for v, h in st) + "\n")
        summ.append(d)
print(open(out_prefix + "_summary.txt").read())
# traffic of the fused kernel for bench.py's roofline.traffic
for d in summ:
    if "fused_loglik" in d["Kernel Name"][0] and d["dram__bytes_read.sum"][0] not in ("", "-nan"):
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
        rd = float(d["dram__bytes_read.sum"][0]) * scale[d["dram__bytes_read.sum"][1]]
        wr = float(d["dram__bytes_write.sum"][0]) * scale[d["dram__bytes_write.sum"][1]]
        smem = d.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", ("", ""))[0]
        smem_pct = d.get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", ("", ""))[0]
        json.dump({"fused_dram_bytes_per_launch": rd + wr, "read": rd, "write": wr, "source": rep,
                   "smem_wavefronts_per_launch": float(smem) if smem else None,
                   "smem_data_stage_pct_of_peak": float(smem_pct) if smem_pct else None,
                   "fp64_pipe_pct_active": float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0]),
                   "issue_slots_pct_active": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0])},
                  open(out_prefix + "_traffic.json", "w"))
