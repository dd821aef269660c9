# N-GPU evidence: GPU tests (incl. the 2-GPU time-sharded test), weak-scaling bench under torchrun, time-sharded objective
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
NG=${1:-2}
nvidia-smi -L | head -8
python -m pytest tests -m gpu -x -q -k "time_sharded or begin_finish" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $NG --steps 20 --warmup 3 > gpurun_out/bench_g$NG.json 2> gpurun_out/bench_g$NG.err
echo "rc=$? bytes=$(wc -c < gpurun_out/bench_g$NG.json)"
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_g$NG.json"))
    print("gpus", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"] if d["e2e"] else None, d["clocks"])
except Exception as e:
    print("no json:", e)
PY
for T in 1000000 4000000; do
TS_P=256 TS_L=64 TS_T=$T python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 scripts/gpu_time_shard.py 2>&1 | grep "time-sharded" | tee -a gpurun_out/time_shard_g$NG.txt
done
for T in 2000000 8000000; do
TS_T=$T python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29535 scripts/gpu_time_shard_fsn.py 2>&1 | grep -E "time-sharded|last rank|Error|error" | tee -a gpurun_out/time_shard_fsn_g$NG.txt
done
