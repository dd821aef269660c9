# config 4 with the thread-per-sub-chunk final kernel vs the warp-per-chunk one (MOIHGP_SCAN_FINAL_WARP=1), after the GPU tests
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
for v in new old; do
if [ $v = old ]; then export MOIHGP_SCAN_FINAL_WARP=1; else unset MOIHGP_SCAN_FINAL_WARP; fi
python bench.py --workload c4 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bq_c4_$v.json 2> gpurun_out/bq_c4_$v.err; tail -2 gpurun_out/bq_c4_$v.err
python - <<PY
import json
d = json.load(open("gpurun_out/bq_c4_$v.json"))
print("c4 $v", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernels_ms_event_bracketed"])
PY
done
