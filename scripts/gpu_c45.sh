cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
for w in c4 c5; do
python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bq_$w.json 2> gpurun_out/bq_$w.err; tail -2 gpurun_out/bq_$w.err
python - <<PY
import json
d = json.load(open("gpurun_out/bq_$w.json"))
print("$w", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernels_ms_event_bracketed"])
PY
done
