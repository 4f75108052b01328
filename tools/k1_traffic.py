"""Extract the DRAM traffic of the K1 launch (the largest cov_mma_kernel launch) from an ncu --set full report.

    python tools/k1_traffic.py gpurun_out/prof_k1.ncu-rep > profiles/k1_traffic.json
"""
import csv
import io
import json
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
best = None
for r in rows[2:]:
    if "cov_mma_kernel" not in r[idx["Kernel Name"]]:
        continue
    rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
    wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
    if best is None or rd + wr > best["dram_bytes_per_launch"]:
        best = {"kernel": r[idx["Kernel Name"]].split("(")[0].replace("void <unnamed>::", ""),
                "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
                "duration_ms_under_ncu": float(r[idx["gpu__time_duration.sum"]]),
                "source": sys.argv[1].split("/")[-1], "how": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum"}
print(json.dumps(best, indent=1))
