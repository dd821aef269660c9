# A/B of library variants built side by side (MOIHGP_B200_LIB): config 5, kernel times
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
for v in base varA varB; do
if [ $v = base ]; then unset MOIHGP_B200_LIB; else export MOIHGP_B200_LIB=$PWD/multioutputihgp_b200/lib/libmoihgp_$v.so; fi
python bench.py --workload c5 --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; tail -1 gpurun_out/ab_$v.err | cut -c1-200
python - <<PY
import json
d = json.load(open("gpurun_out/ab_$v.json"))
print("$v", d["ms_per_step"], d["roofline"]["kernels_ms_event_bracketed"])
PY
done
