#!/bin/bash
# Kernel-shape tuning session (development library, NNGP_TUNE_SHAPE knobs): gpurun -- bash tools/gpu_tune.sh
mkdir -p gpurun_out
KNOBS=${KNOBS15:-default,44,44f,44f3} timeout 600 python tools/tune.py cfg3 float64 > gpurun_out/tune_m15.log 2>&1; echo "tune m15 rc=$?"; cat gpurun_out/tune_m15.log
KNOBS=${KNOBS15:-default,44,44f,44f3} TUNE_D=3 timeout 600 python tools/tune.py cfg3 float64 > gpurun_out/tune_m15_3d.log 2>&1; echo "tune m15 3d rc=$?"; cat gpurun_out/tune_m15_3d.log
KNOBS=${KNOBS30:-default,84,84f,162f2,162f3b} timeout 600 python tools/tune.py cfg4 float64 2000000 > gpurun_out/tune_m30.log 2>&1; echo "tune m30 rc=$?"; cat gpurun_out/tune_m30.log
KNOBS=${KNOBS30:-default,84,84f,84r,162f2,162f3b} TUNE_D=2 timeout 600 python tools/tune.py cfg4 float64 2000000 > gpurun_out/tune_m30_2d.log 2>&1; echo "tune m30 2d rc=$?"; cat gpurun_out/tune_m30_2d.log
