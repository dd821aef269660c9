cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
CMD="python bench.py --workload c5 --steps 1 --warmup 3 --no-e2e --no-cpu --tlen 400000"
$CMD > gpurun_out/plain_ncu_c5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_project_mma|k_gradU_mma|k_obj_lanes|k_obj_carry" -s 15 -c 5 -f -o gpurun_out/prof_c5 $CMD > gpurun_out/ncu_c5.log 2>&1
tail -2 gpurun_out/ncu_c5.log | cut -c1-200
