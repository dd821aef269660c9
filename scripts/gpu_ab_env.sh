#!/bin/bash
# A/B of one environment knob over bench workloads (run under gpurun):
#   scripts/gpu_ab_env.sh <workload> "<bench flags>" VAR v1 v2 ...     ('-' = variable unset)
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
wl=$1; flags=$2; var=$3; shift 3
for v in "$@"; do
  if [ "$v" = "-" ]; then unset $var; else export $var=$v; fi
  python bench.py --workload $wl $flags --no-e2e --no-cpu > gpurun_out/ab_${wl}_${var}_$v.json 2> gpurun_out/ab_${wl}_${var}_$v.err || tail -3 gpurun_out/ab_${wl}_${var}_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_${wl}_${var}_$v.json"))
    print("$wl $var=$v", "ms %.4f" % d["ms_per_step"], "frac %.4f" % d["roofline"]["frac"], d["roofline"]["kernels_ms_event_bracketed"])
except Exception as e:
    print("$wl $var=$v failed:", e)
PY
done
