#!/bin/bash
# One bench line per BASELINE.json config (1 GPU): c1 64x64x16 IMNet+IMNet, c2 4x320x240x64, c3 8x640x480x64 (the bench
# default), c4's per-GPU share 4x640x480x64, c5 8x640x480x128 + second stage.  Output: gpurun_out/configs.jsonl
mkdir -p gpurun_out
: > gpurun_out/configs.jsonl
run() { echo "--- $*" >&2; timeout 900 python bench.py "$@" --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/configs.jsonl; }
run --workload c1 --offdec IMNET --steps 20
run --workload c2 --steps 10
run --workload c3 --steps 5
run --workload c4 --steps 5
run --workload c5 --stage2 --no-e2e --steps 3
python - <<'PY'
import json
for line in open("gpurun_out/configs.jsonl"):
    d = json.loads(line)
    e = d.get("e2e") or {}
    s2 = d.get("stage2") or {}
    print(d["config"]["workload"][:2], f"value {d['value']:.3e} pts/s  step {d['ms_per_step']:.2f} ms  kernel {d['roofline']['kernel_ms']:.2f} ms  frac {d['roofline']['frac']:.3f}"
          f"  e2e {e.get('value', float('nan')):.3e} ({e.get('ms_per_step', float('nan')):.1f} ms, sync {e.get('sync_ms_per_step', float('nan')):.1f})  launches {d['gpu_launches']}"
          f"  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}  stage2 {s2.get('ms')}")
PY
