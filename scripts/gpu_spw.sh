cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for spw in 4 2 1; do
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --spw $spw > gpurun_out/bq_spw$spw.json 2> gpurun_out/bq_spw$spw.err
python - <<PY
import json
d = json.load(open("gpurun_out/bq_spw$spw.json"))
print("spw=$spw", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_share"], d["clocks"])
PY
done
