#!/bin/bash
# DRAM bytes of one fused pass of every single-GPU bench workload at FULL size (run under gpurun, ONE GPU):
#   scripts/gpu_traffic.sh [c3 c4 c5]
# ncu collects dram__bytes_read/write + duration for every kernel of `bench.py --steps 1 --warmup 3` (4 passes); the CSVs land in
# gpurun_out/traffic_<wl>.csv and scripts/make_traffic.py turns them into profiles/traffic.json (stamped with the source hash).
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
for wl in ${@:-c3 c4 c5}; do
  CMD="python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu --no-also"
  $CMD > gpurun_out/traffic_${wl}_plain.log 2>&1 || { echo "plain run of $wl failed"; continue; }
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'k_' --csv \
      --log-file gpurun_out/traffic_$wl.csv $CMD > gpurun_out/traffic_${wl}_ncu.log 2>&1
  echo "$wl: $(grep -c dram__bytes_read gpurun_out/traffic_$wl.csv) kernels captured"
done
python scripts/make_traffic.py gpurun_out --dry
