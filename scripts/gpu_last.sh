cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/gpu_fuzz.py 2>&1 | tail -30 | tee gpurun_out/fuzz.txt
