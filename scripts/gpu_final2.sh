# round-end check on one GPU: tests, smoke, default bench, the same workload on the chunked-scan path, launch list (ncu)
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.err | cut -c1-200
python bench.py --path scan --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c3_scanpath.json 2> gpurun_out/bench_c3_scanpath.err; tail -1 gpurun_out/bench_c3_scanpath.err | cut -c1-200
python - <<'PY'
import json
for f in ("bench_default", "bench_c3_scanpath"):
    d = json.load(open("gpurun_out/%s.json" % f))
    print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernels_ms_event_bracketed"], (d.get("e2e") or {}).get("value"), d.get("gpu_launches"), d.get("clocks"))
PY
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_c3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_c3_final.csv $CMD > gpurun_out/ncu_launch.log 2>&1
grep -c "k_filter_chain\|k_smooth_chain" gpurun_out/launches_c3_final.csv
