cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
NG=${1:-4}
for mode in with without; do
if [ $mode = without ]; then export BENCH_SKIP_ALLREDUCE=1; else unset BENCH_SKIP_ALLREDUCE; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $NG --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_g${NG}_$mode.json 2> gpurun_out/bench_g${NG}_$mode.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_g${NG}_$mode.json"))
print("$mode all-reduce: gpus", d["n_gpus"], "value %.4g" % d["value"], "ms", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 4), d["roofline"]["kernels_ms_event_bracketed"], d["clocks"])
PY
done
