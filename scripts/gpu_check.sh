#!/bin/bash
# One GPU-box check, parameterised (run under gpurun):
#   scripts/gpu_check.sh [tests[:<pytest -k expr>]] [smoke] [bench:<workload>[:<extra bench.py flags>]] [ref] ...
# Outputs go to gpurun_out/ (merged back by gpurun).  Examples:
#   gpurun -- 'bash scripts/gpu_check.sh tests smoke bench:c3 bench:c4:--no-e2e bench:c5:--no-e2e'
#   gpurun -- 'bash scripts/gpu_check.sh tests:config5 bench:c4:"--no-e2e --no-cpu --tlen 2000000"'
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
TAG=${TAG:-run}
for what in "$@"; do
  kind=${what%%:*}; rest=${what#*:}; [ "$rest" = "$what" ] && rest=""
  case $kind in
    tests)
      if [ -n "$rest" ]; then K=(-k "$rest"); else K=(); fi
      python -m pytest tests -m gpu -x -q --durations=8 "${K[@]}" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -14 gpurun_out/${TAG}_tests.log ;;
    smoke)
      python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log | cut -c1-400 ;;
    bench)
      wl=${rest%%:*}; flags=${rest#*:}; [ "$flags" = "$rest" ] && flags=""
      python bench.py --workload $wl $flags > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/${TAG}_bench_$wl.err | cut -c1-300
      python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_bench_$wl.json"))
    r = d["roofline"]
    print("$wl", "ms %.4f" % d["ms_per_step"], "value %.4g" % d["value"], "frac %.4f" % r["frac"], r["kernels_ms_event_bracketed"])
    print("   e2e", d.get("e2e"), "cpu", d.get("cpu_baseline"), "clocks", d.get("clocks"))
    for k, v in (d.get("also") or {}).items():
        print("   also", k, v)
except Exception as e:
    print("no bench line:", e)
PY
      ;;
    ref)
      python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/${TAG}_bench_ref.json ;;
    *) echo "unknown item $what" ;;
  esac
done
