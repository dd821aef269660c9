set -x
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t2.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -3 gpurun_out/t2.log
