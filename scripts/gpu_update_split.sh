cd ${GRAFT_REPO_ROOT:-.}; mkdir -p gpurun_out
python scripts/probe_update.py 2>&1 | tail -1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_polar|k_setup" -c 9 --csv --log-file gpurun_out/update_split.csv python scripts/probe_update.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/update_split.csv")))
hdr = [r for r in rows if "Kernel Name" in r][0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
for r in rows:
    if len(r) > iv and r is not hdr and r[ik]:
        print(r[ik].split("(")[0][-28:], r[iv], "us")
PY
