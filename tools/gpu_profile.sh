#!/bin/bash
# ncu launch list of the bench command + full captures of the decoder kernel and of the two row-prep kernels (1 GPU).
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh [tag] [workload]'
TAG=${1:-r1}
WL=${2:-c2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_profile_$TAG.log) 2>&1
BENCH="python bench.py --workload $WL --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
echo "=== launch list ($BENCH)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $BENCH | tail -2
echo "=== full capture of k_mlp_tc"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_mlp_tc -s 3 -c 1 -f -o gpurun_out/prof_mlp_tc_$TAG $BENCH | tail -2
echo "=== full capture of k_roi_align_rays, k_rowprep"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_roi_align_rays|k_rowprep' -s 6 -c 3 -f -o gpurun_out/prof_prep_$TAG $BENCH | tail -2
ls -la gpurun_out
echo "=== done"
