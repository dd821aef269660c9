# round-end evidence: GPU tests, smoke, default bench (e2e + cpu legs), configs 4 and 5, ncu summaries of the new kernels
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/tfull.log; tail -3 gpurun_out/tfull.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
for w in c4 c5; do
python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err
done
python - <<'PY'
import json
for f in ("bench_default", "bench_c4", "bench_c5"):
    d = json.load(open("gpurun_out/%s.json" % f))
    print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernels_ms_event_bracketed"], d.get("e2e"), d.get("clocks"))
PY
CMD="python bench.py --workload c4 --steps 1 --warmup 3 --no-e2e --no-cpu --tlen 2000000"
$CMD > gpurun_out/plain_ncu_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_scan_lanes|k_project_mma" -s 2 -c 2 -f -o gpurun_out/prof_c4_scan $CMD > gpurun_out/ncu_c4.log 2>&1
tail -1 gpurun_out/ncu_c4.log | cut -c1-200
CMD="python bench.py --workload c5 --steps 1 --warmup 3 --no-e2e --no-cpu --tlen 400000"
$CMD > gpurun_out/plain_ncu_c5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_project_mma|k_gradU_mma|k_obj_lanes|k_obj_carry" -s 15 -c 5 -f -o gpurun_out/prof_c5 $CMD > gpurun_out/ncu_c5.log 2>&1
tail -1 gpurun_out/ncu_c5.log | cut -c1-200
