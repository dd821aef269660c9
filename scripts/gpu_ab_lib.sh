#!/bin/bash
# A/B of library variants built by scripts/build_variant.sh (run under gpurun):
#   scripts/gpu_ab_lib.sh <workload> "<bench flags>" <variant> [<variant> ...]      ('-' = the stock library)
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
wl=$1; flags=$2; shift 2
for v in "$@"; do
  if [ "$v" = "-" ]; then unset MOIHGP_B200_LIB; else export MOIHGP_B200_LIB=$PWD/multioutputihgp_b200/lib/ab/libmoihgp_$v.so; fi
  python bench.py --workload $wl $flags --no-e2e --no-cpu --no-also > gpurun_out/ablib_${wl}_$v.json 2> gpurun_out/ablib_${wl}_$v.err || tail -3 gpurun_out/ablib_${wl}_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ablib_${wl}_$v.json"))
    print("$wl lib=$v", "ms %.4f" % d["ms_per_step"], "frac %.4f" % d["roofline"]["frac"], d["roofline"]["kernels_ms_event_bracketed"])
except Exception as e:
    print("$wl lib=$v failed:", e)
PY
done
