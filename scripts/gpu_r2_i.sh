#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_c5_launches.csv python scripts/one_c5.py 4096 256 > gpurun_out/r2_c5_ncu.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2_c5_launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
kn, val = H.index("Kernel Name"), H.index("Metric Value")
data = [(r[kn][:70], float(r[val].replace(",", ""))) for r in rows[hdr + 1:] if len(r) > val]
n = len(data) // 3
for name, v in data[-n:]:
    print(f"{v/1e3:10.1f} us  {name}")
PY
