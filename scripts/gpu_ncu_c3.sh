# ncu --set full of the two many-chains kernels on a short-T version of the C3 workload (same shapes, T = 2048)
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --tlen 2048"
$CMD > gpurun_out/plain_ncu.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:chain -s 6 -c 2 -f -o gpurun_out/prof_v3 $CMD > gpurun_out/ncu_v3.log 2>&1
tail -3 gpurun_out/ncu_v3.log
