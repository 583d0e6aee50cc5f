#!/bin/bash
# One gpurun call: primitives, parity (fp32 engine, then tcgen05 engine in its own process), smoke, short benches.
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
mkdir -p gpurun_out
exec > >(tee gpurun_out/gpu_check.log) 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
python -c "import os; print('cpus', os.cpu_count())"; free -g | head -2
echo "=== tc primitives"; timeout 300 python -m pytest tests/test_gpu_tc_primitives.py -q --timeout 120 2>&1 | tail -15
echo "=== parity: fp32 engine + engine-independent tests"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -k "not tc_bf16x3 and not primitives and not module_surface" 2>&1 | tail -25
echo "=== parity: tcgen05 engine"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -k "tc_bf16x3 or module_surface" 2>&1 | tail -40
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
echo "=== bench c2 simt"; timeout 600 python bench.py --workload c2 --engine simt_fp32 --steps 3 --warmup 3 --torch-gpu-baseline --cpu-sample-pairs 262144 2>&1 | tail -3
echo "=== bench c2 tc"; timeout 600 python bench.py --workload c2 --engine auto --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -3
echo "=== bench c3 tc (default)"; timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -3
echo "=== done"
