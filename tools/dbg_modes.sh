#!/bin/bash
# Timing experiment: decoder-kernel time with parts of k_mlp_tc compiled out (build/dbg/lib_m<mode>.so, made beforehand by
# `bash tools/build_variants.sh modes`; TC_DEBUG_MODE bits:
# 1 no MMA issue, 2 no TMEM loads in the epilogues, 4 no epilogue math/stores, 8 no sincos in the operand build).
WL=${1:-c2}
for m in 0 1 2 4 6 7 8 14; do
  if [ $m -eq 0 ]; then unset LIDF_QUERY_LIB; else export LIDF_QUERY_LIB=$PWD/build/dbg/lib_m$m.so; fi
  echo -n "mode $m: "
  timeout 300 python bench.py --workload $WL --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('kernel_ms', round(d['roofline']['kernel_ms'], 3), 'step_ms', round(d['ms_per_step'], 3), 'sm_mhz', d['clocks']['sm_mhz'])
except Exception as e:
    print('failed', e)
"
done
