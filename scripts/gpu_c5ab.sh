# config 5 with the thread-per-sub-chunk objective kernels vs the warp-scan ones (MOIHGP_OBJ_WARP=1), after the GPU tests
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
for v in new old; do
if [ $v = old ]; then export MOIHGP_OBJ_WARP=1; else unset MOIHGP_OBJ_WARP; fi
python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bq_c5_$v.json 2> gpurun_out/bq_c5_$v.err; tail -2 gpurun_out/bq_c5_$v.err
python - <<PY
import json
d = json.load(open("gpurun_out/bq_c5_$v.json"))
print("c5 $v", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernels_ms_event_bracketed"])
PY
done
