cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
python scripts/probe_online_latency.py 2>&1 | tee gpurun_out/online_latency.txt
ONLINE_WINDOW=1 python scripts/bench_online.py 2>&1 | tail -1 | tee -a gpurun_out/online_latency.txt
ONLINE_WINDOW=64 python scripts/bench_online.py 2>&1 | tail -1 | tee -a gpurun_out/online_latency.txt
