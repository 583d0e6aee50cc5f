#!/bin/bash
# One gpurun call mirroring the driver's end-of-round sequence: GPU tests, smoke, bench (ours + reference arm).
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [tag]'      -> gpurun_out/check_<tag>.log
TAG=${1:-r1}
mkdir -p gpurun_out
exec > >(tee gpurun_out/check_$TAG.log) 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total,power.limit --format=csv
python -c "import os; print('cpus', os.cpu_count())"; free -g | head -2
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -6
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
echo "=== bench (default: c3, N = 1)"; timeout 600 python bench.py 2>&1 | tail -1
echo "=== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1
echo "=== done"
