"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X <cmd>`) per kernel.

    python tools/launch_summary.py gpurun_out/launches.csv "<command line>" > profiles/launches_rNN_summary.csv
"""
import csv
import re
import sys
from collections import defaultdict


def main(path, cmd):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    i_name, i_metric, i_unit, i_val = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows[1:]:
        if len(r) <= i_val or r[i_metric] != "gpu__time_duration.sum":
            continue
        v = float(r[i_val].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[i_unit], 1e-6)
        name = re.sub(r"\(.*", "", r[i_name]).replace("void ", "").replace("<unnamed>::", "")
        tot[name] += v * scale
        cnt[name] += 1
    total = sum(tot.values())
    print("# ncu launch list summary (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare shares)")
    print(f"# command: {cmd}")
    print(f"# total kernel time {total:.1f} ms over {sum(cnt.values())} launches")
    print("kernel,launches,total_ms,share")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"\"{k}\",{cnt[k]},{tot[k]:.3f},{tot[k] / total:.4f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
