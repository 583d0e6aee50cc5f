#!/bin/bash
# ncu launch list of the bench command + one full capture of the decoder kernel (1 GPU).
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh [tag]'
TAG=${1:-r1}
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_profile_$TAG.log) 2>&1
BENCH="python bench.py --workload c2 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
echo "=== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $BENCH | tail -2
echo "=== full capture of k_mlp_tc"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 3 -c 1 -f -o gpurun_out/prof_mlp_tc_$TAG $BENCH | tail -2
ls -la gpurun_out
echo "=== done"
