#!/usr/bin/env python
"""profiles/traffic.json from the ncu CSVs of scripts/gpu_traffic.sh.
usage: python scripts/make_traffic.py <dir with traffic_c3.csv ...> [--dry]
Per workload: DRAM bytes read + written by the library's kernels of ONE fused pass (the run holds `passes` identical passes:
every kernel's bytes are summed over its launches and divided by the number of passes; k_setup / k_polar* belong to
update(params), not to the pass), the per-kernel breakdown, and the sha256 of the CUDA sources (bench.source_hash) the
capture was taken with - bench.py refuses to quote it for any other sources."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import source_hash  # noqa: E402

PASSES = 4          # --steps 1 --warmup 3
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
WORK = {"c3": "BASELINE configs[2], full size (N=4096, T=16384, p=16, L=8, d=3)",
        "c4": "BASELINE configs[3], full size (N=1, T=1e7, p=64, L=32, d=2)",
        "c5": "BASELINE configs[4], full size (N=1, T=1e6, p=256, L=64, d=2)"}
ALG = {"c3": 64.0 * 4096 * 16384 * 8, "c4": 48.0 * 1e7 * 32, "c5": 32.0 * 1e6 * 64}


def parse(path):
    rows = [l for l in open(path) if l.startswith('"')]
    per = {}
    for r in csv.DictReader(rows):
        m_ = re.search(r"\b(k_[A-Za-z0-9_]+)", re.sub(r"\(.*", "", r["Kernel Name"]))     # e.g. "void unnamed>::k_filter_chain<16, 8, 3, 4>(...)"
        if not m_:
            continue
        short = m_.group(1)
        if short.startswith("k_setup") or short.startswith("k_polar"):
            continue
        k = per.setdefault(short, {"launches": 0, "dram_bytes_read": 0.0, "dram_bytes_write": 0.0, "duration_ms_under_ncu": 0.0})
        val = float(r["Metric Value"].replace(",", ""))
        m = r["Metric Name"]
        if m == "dram__bytes_read.sum":
            k["dram_bytes_read"] += val * UNIT[r["Metric Unit"]]
            k["launches"] += 1
        elif m == "dram__bytes_write.sum":
            k["dram_bytes_write"] += val * UNIT[r["Metric Unit"]]
        elif m == "gpu__time_duration.sum":
            k["duration_ms_under_ncu"] += val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r["Metric Unit"], 1e-6)
    for k in per.values():
        for f in ("dram_bytes_read", "dram_bytes_write", "duration_ms_under_ncu"):
            k[f] = k[f] / PASSES
        k["launches_per_pass"] = k.pop("launches") / PASSES
    return per


def main():
    src = sys.argv[1]
    dry = "--dry" in sys.argv
    out = {}
    for wl in ("c3", "c4", "c5"):
        p = os.path.join(src, "traffic_%s.csv" % wl)
        if not os.path.exists(p):
            continue
        # a capture older than the newest CUDA source does not describe the kernels the hash below stands for
        import glob
        newest = max(os.path.getmtime(x) for x in glob.glob(os.path.join(ROOT, "multioutputihgp_b200", "csrc", "*")) if x.endswith((".cu", ".cuh", ".h")))
        if os.path.getmtime(p) < newest and "--force" not in sys.argv:
            print("%s: capture is older than the CUDA sources - re-run scripts/gpu_traffic.sh (or pass --force)" % wl)
            continue
        per = parse(p)
        total = sum(k["dram_bytes_read"] + k["dram_bytes_write"] for k in per.values())
        out[wl] = {"workload": WORK[wl], "source_hash": source_hash(),
                   "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none of `python bench.py --workload %s "
                             "--steps 1 --warmup 3 --no-e2e --no-cpu --no-also` (scripts/gpu_traffic.sh), per pass" % wl,
                   "fused_pass_dram_bytes_per_step": total, "algorithmic_bytes_per_step": ALG[wl], "kernels": per}
        print("%s: %.2f GB per pass (algorithmic %.2f GB, x%.2f), kernels: %s" % (
            wl, total / 1e9, ALG[wl] / 1e9, total / ALG[wl],
            ", ".join("%s %.2f GB / %.3f ms" % (n, (k["dram_bytes_read"] + k["dram_bytes_write"]) / 1e9, k["duration_ms_under_ncu"]) for n, k in per.items())))
    if not dry and out:
        with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
            json.dump(out, f, indent=1)
        print("wrote profiles/traffic.json (source hash %s)" % source_hash())


if __name__ == "__main__":
    main()
