#!/bin/bash
# A/B timing of alternative builds of the library (build/dbg/lib_<tag>.so from tools/build_variants.sh, selected with
# LIDF_QUERY_LIB): decoder-kernel
# time, step time and SM clock of the bench workload for every tag given.  Usage: bash tools/ab_libs.sh c3 base h1000 ...
WL=$1; shift
for tag in "$@"; do
  export LIDF_QUERY_LIB=$PWD/build/dbg/lib_$tag.so
  for rep in 1 2; do
    echo -n "$tag run $rep: "
    timeout 300 python bench.py --workload $WL --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('kernel_ms', round(d['roofline']['kernel_ms'], 3), 'step_ms', round(d['ms_per_step'], 3), 'sm_mhz', d['clocks']['sm_mhz'])
except Exception as e:
    print('failed', e)
"
  done
done
