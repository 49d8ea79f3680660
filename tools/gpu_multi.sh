#!/bin/bash
# Multi-GPU session: gpurun --gpus N -- bash tools/gpu_multi.sh N [tag]
N=${1:-2}; TAG=${2:-r2_${N}gpu}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.csv 2>&1
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 tools/check_multi_gpu.py cfg3 > gpurun_out/${TAG}_check.log 2>&1; echo "check rc=$?"; tail -2 gpurun_out/${TAG}_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29643 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench torchrun rc=$?"; cut -c1-330 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 900 python bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/${TAG}_bench_sp.json 2> gpurun_out/${TAG}_bench_sp.err; echo "bench single-process rc=$?"; cut -c1-330 gpurun_out/${TAG}_bench_sp.json; tail -3 gpurun_out/${TAG}_bench_sp.err
