#!/bin/bash
# 8-GPU session (charged 8x): the consistency check and the bench line in both launch modes.
# gpurun --gpus 8 -- bash tools/gpu_multi8.sh [tag]
N=8; TAG=${1:-r2_8gpu}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 tools/check_multi_gpu.py cfg3 > gpurun_out/${TAG}_check.log 2>&1; echo "check rc=$?"; tail -1 gpurun_out/${TAG}_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29643 bench.py --gpus $N --steps 50 --warmup 5 --cpu-seconds 2 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench torchrun rc=$?"; cut -c1-330 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --gpus $N --steps 50 --warmup 5 --cpu-seconds 2 > gpurun_out/${TAG}_bench_sp.json 2> gpurun_out/${TAG}_bench_sp.err; echo "bench single-process rc=$?"; cut -c1-330 gpurun_out/${TAG}_bench_sp.json; tail -3 gpurun_out/${TAG}_bench_sp.err
