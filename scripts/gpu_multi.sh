cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
NG=${1:-2}
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/bench_g$NG.json 2> gpurun_out/bench_g$NG.err
echo "rc=$? bytes=$(wc -c < gpurun_out/bench_g$NG.json)"
tail -5 gpurun_out/bench_g$NG.err | cut -c1-400
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_g$NG.json"))
    print("gpus", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"] if d["e2e"] else None, d["clocks"])
except Exception as e:
    print("no json:", e)
PY
