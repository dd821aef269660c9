#!/bin/bash
# ncu --set full capture of selected kernels of one bench workload (run under gpurun, ONE GPU):
#   scripts/gpu_ncu.sh <name> <workload> "<kernel regex>" "<bench flags>" [skip] [count]
# The plain command runs first (must exit 0 without ncu); the report lands in gpurun_out/<name>.ncu-rep and its summary
# (scripts/ncu_summary.py) in gpurun_out/<name>.txt.
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
name=$1; wl=$2; rx=$3; flags=$4; skip=${5:-2}; cnt=${6:-2}
CMD="python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu --no-also $flags"
$CMD > gpurun_out/${name}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${name}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o gpurun_out/$name $CMD > gpurun_out/${name}_ncu.log 2>&1
tail -2 gpurun_out/${name}_ncu.log | cut -c1-300
python scripts/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/$name.txt 2>&1; head -60 gpurun_out/$name.txt
