#!/bin/bash
# One-rep A/B of the energy-experiment builds (build/dbg/lib_<tag>.so from tools/build_variants.sh): decoder-kernel time,
# step time, SM clock.  Usage: bash tools/ab_exp.sh c3 base e1 e2 ...   ("new" = the product library in csrc/)
WL=$1; shift
mkdir -p gpurun_out
for tag in "$@"; do
  if [ "$tag" = new ]; then unset LIDF_QUERY_LIB; else export LIDF_QUERY_LIB=$PWD/build/dbg/lib_$tag.so; fi
  echo -n "$tag: "
  timeout 300 python bench.py --workload $WL --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-torch-gpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('kernel_ms', round(d['roofline']['kernel_ms'], 3), 'step_ms', round(d['ms_per_step'], 3), 'sm_mhz', d['clocks']['sm_mhz'], 'frac', round(d['roofline']['frac'], 4))
except Exception as e:
    print('failed', e)
"
done 2>&1 | tee -a gpurun_out/ab_exp.log
