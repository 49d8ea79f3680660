#!/bin/bash
# One GPU-box session: GPU parity tests, the bench line (both arms), the ncu launch list and one full capture
# of the two fused kernels the metric depends on.  Usage (from the repo root): gpurun -- bash tools/gpu_check.sh [tag]
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/${TAG}_bench.json
tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
if [ "$2" != "noncu" ]; then
timeout 600 python bench.py --steps 3 --warmup 3 --no-cfg4 --cpu-seconds 0.5 > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cfg4 --cpu-seconds 0.5 > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 300 python tools/prof_driver.py cfg3 float64 3 > gpurun_out/${TAG}_prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_loglik -c 3 -f -o gpurun_out/${TAG}_prof_fused \
    python tools/prof_driver.py cfg3 float64 3 > gpurun_out/${TAG}_ncu_fused.log 2>&1; echo "ncu fused rc=$?"
timeout 300 python tools/prof_driver.py cfg4 float64 2 2000000 > gpurun_out/${TAG}_prof_plain4.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_loglik -c 2 -f -o gpurun_out/${TAG}_prof_fused_m30 \
    python tools/prof_driver.py cfg4 float64 2 2000000 > gpurun_out/${TAG}_ncu_fused_m30.log 2>&1; echo "ncu fused m30 rc=$?"
fi
ls -la gpurun_out | tail -20
if [ "$2" != "noncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_grid_query -c 8 -f -o gpurun_out/${TAG}_prof_knn \
    python tools/prof_driver.py cfg3 float64 1 > gpurun_out/${TAG}_ncu_knn.log 2>&1; echo "ncu knn rc=$?"
fi
