#!/bin/bash
# compute-sanitizer over the backward kernels on small cases.  Usage: gpurun --timeout 900 -- 'bash tools/gpu_sanitize_bwd.sh [tag]'
TAG=${1:-r2}
mkdir -p gpurun_out
exec > >(tee gpurun_out/sanitizer_bwd_$TAG.log) 2>&1
echo "== memcheck: packed / fp32-row wgrad self tests (small), backward vs reference autograd goldens (both modes)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_backward.py -x -q --timeout 500 \
  -k "(wgrad and (64-64 or 1000 or 333)) or reproduces" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|Error" | head -20
echo "== racecheck: packed wgrad self test, backward golden (full mode)"
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_backward.py -x -q --timeout 500 \
  -k "(packed_wgrad and 1000) or (reproduces and ragged and False)" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | head -20
echo "== done"
