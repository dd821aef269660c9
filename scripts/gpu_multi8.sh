cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
NG=${1:-8}
nvidia-smi -L | wc -l; free -g | head -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/bench_g$NG.json 2> gpurun_out/bench_g$NG.err
echo "rc=$? bytes=$(wc -c < gpurun_out/bench_g$NG.json)"; tail -3 gpurun_out/bench_g$NG.err | cut -c1-300
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_g$NG.json"))
    print("gpus", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"] if d["e2e"] else None, d["clocks"])
except Exception as e:
    print("no json:", e)
PY
TS_P=256 TS_L=64 TS_T=8000000 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 scripts/gpu_time_shard.py 2>&1 | grep "time-sharded" | tee -a gpurun_out/time_shard_g$NG.txt
