# quick GPU check: parity tests + C3 bench without the e2e / CPU legs
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tq.log
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bq_c3.json 2> gpurun_out/bq_c3.err
tail -5 gpurun_out/tq.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/bq_c3.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_share"], d["clocks"])
PY
