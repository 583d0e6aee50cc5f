#!/bin/bash
# One gpurun call for a change to the backward kernels: backward tests, the config-4 training step, launch list of a
# config-2 training step.  Usage: gpurun --timeout 900 -- 'bash tools/gpu_bwd_check.sh [tag]'  -> gpurun_out/bwd_check_<tag>.log
TAG=${1:-r2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/bwd_check_$TAG.log) 2>&1
echo "=== backward tests"; timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_dropin.py -x -q --timeout 300 2>&1 | tail -8
echo "=== bench --workload c4 --train"
timeout 300 python bench.py --workload c4 --train --steps 3 --warmup 3 --no-cpu-baseline --no-torch-gpu-baseline 2>&1 | tail -1
echo "=== launch list of a c2 training step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_${TAG}_train_c2.csv \
  python bench.py --workload c2 --train --steps 1 --warmup 1 --no-cpu-baseline --no-torch-gpu-baseline 2>&1 | tail -1
python tools/launch_table.py gpurun_out/launches_${TAG}_train_c2.csv | head -30
echo "=== done"
