cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
CMD="python bench.py --workload c4 --steps 1 --warmup 3 --no-e2e --no-cpu --tlen 2000000"
$CMD > gpurun_out/plain_ncu_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_scan_lanes|k_project_mma" -s 2 -c 2 -f -o gpurun_out/prof_c4_scan $CMD > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log | cut -c1-300
