cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
bash scripts/gpu_probe_online.sh 2>&1 | tail -14
