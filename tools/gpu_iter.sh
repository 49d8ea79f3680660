#!/bin/bash
# Kernel-iteration session: microbenchmark, GPU parity tests, short bench, stage-1 timing, full ncu captures.
# Usage: gpurun -- bash tools/gpu_iter.sh TAG [noncu]
TAG=${1:-it}
mkdir -p gpurun_out

timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 30 --warmup 5 --cpu-seconds 2 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench.json"))
    print("cfg3 ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "parity", d["parity"]["rel_err_vs_oracle"])
    print("sweep", d["detail"]["sweep_cfg5"]["ms_per_eval"], "cold", d["detail"]["cold_e2e"]["ms"], d["detail"]["cold_e2e"]["knn_ms"])
    c4 = d["detail"]["cfg4"]; print("cfg4 ms", c4["ms_per_eval"], "frac", c4["roofline_frac"], "knn_s", c4["knn_build_s"], "parity", c4["parity"]["rel_err_vs_oracle"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python tools/knn_bench.py cfg3 --lams 0.5,1,2,4 > gpurun_out/${TAG}_knn_cfg3.txt 2>&1; cat gpurun_out/${TAG}_knn_cfg3.txt
timeout 300 python tools/knn_bench.py cfg4 --lams 0.5,1,2,4 > gpurun_out/${TAG}_knn_cfg4.txt 2>&1; cat gpurun_out/${TAG}_knn_cfg4.txt
if [ "$2" != "noncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_loglik -c 2 -f -o gpurun_out/${TAG}_prof_fused \
    python tools/prof_driver.py cfg3 float64 2 > gpurun_out/${TAG}_ncu_fused.log 2>&1; echo "ncu fused rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_loglik -c 1 -f -o gpurun_out/${TAG}_prof_fused_m30 \
    python tools/prof_driver.py cfg4 float64 1 2000000 > gpurun_out/${TAG}_ncu_fused_m30.log 2>&1; echo "ncu fused m30 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_grid_query -c 3 -f -o gpurun_out/${TAG}_prof_knn \
    python tools/prof_driver.py cfg3 float64 1 > gpurun_out/${TAG}_ncu_knn.log 2>&1; echo "ncu knn rc=$?"
fi
ls gpurun_out | grep ${TAG} | head -30
