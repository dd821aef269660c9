# A/B of an environment switch in one GPU session: usage  gpu_ab_env.sh VAR  (runs bench with VAR unset and VAR=0, 3 reps)
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
VAR=$1
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for rep in 1 2 3; do
for v in default off; do
if [ $v = off ]; then export $VAR=0; else unset $VAR; fi
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/ab_$v.json 2>/dev/null
python - <<PY
import json
d = json.load(open("gpurun_out/ab_$v.json"))
print("$VAR $v rep$rep", round(d["ms_per_step"], 3), d["roofline"]["kernels_ms_event_bracketed"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done; done
