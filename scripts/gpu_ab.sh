# A/B of two library builds in one GPU session (A = in-tree libmoihgp.so, B = lib/libmoihgp_varB.so)
cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
for rep in 1 2 3; do
for v in A B; do
if [ $v = B ]; then export MOIHGP_B200_LIB=$PWD/multioutputihgp_b200/lib/libmoihgp_varB.so; else unset MOIHGP_B200_LIB; fi
python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu > gpurun_out/ab_$v.json 2>/dev/null
python - <<PY
import json
d = json.load(open("gpurun_out/ab_$v.json"))
print("$v rep$rep", round(d["ms_per_step"], 3), d["roofline"]["kernels_ms_event_bracketed"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done; done
