# profiles for the default bench command (BASELINE configs[2]): launch list + --set full of both kernels at full size
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_c3.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_c3_v3.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain_c3b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:chain -s 6 -c 2 -f -o gpurun_out/prof_c3_full $CMD > gpurun_out/ncu_c3_full.log 2>&1
tail -2 gpurun_out/ncu_c3_full.log | cut -c1-200
