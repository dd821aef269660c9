cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bq_c5.json 2> gpurun_out/bq_c5.err; tail -2 gpurun_out/bq_c5.err
python - <<PY
import json
d = json.load(open("gpurun_out/bq_c5.json"))
print("c5", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernels_ms_event_bracketed"])
PY
CMD="python bench.py --workload c5 --steps 1 --warmup 3 --no-e2e --no-cpu --tlen 200000"
$CMD > gpurun_out/plain_ncu_c5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_project_mma|k_gradU_mma|k_obj_scan" -s 12 -c 4 -f -o gpurun_out/prof_c5 $CMD > gpurun_out/ncu_c5.log 2>&1
tail -2 gpurun_out/ncu_c5.log | cut -c1-200
