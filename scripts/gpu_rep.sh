cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
for rep in 1 2 3 4; do
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/rep.json 2>/dev/null
python - <<PY
import json
d = json.load(open("gpurun_out/rep.json"))
print("rep$rep", round(d["ms_per_step"], 3), round(d["roofline"]["frac"], 4), d["roofline"]["kernels_ms_event_bracketed"], d["clocks"])
PY
done
