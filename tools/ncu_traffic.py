#!/usr/bin/env python
"""DRAM traffic of one kernel from an `ncu --page raw --csv` dump -> the JSON bench.py reads for `roofline.traffic`.

    python tools/ncu_traffic.py gpurun_out/x_raw.csv wigner_bwd_dg <samples per launch> profiles/r02_wigner_bwd_traffic.json
"""
import csv
import json
import sys

path, filt, samples, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
rows = list(csv.reader(open(path)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for d in data:
    name = d[idx["Kernel Name"]]
    if filt not in name:
        continue
    rd = float(d[idx["dram__bytes_read.sum"]]) * SCALE[units[idx["dram__bytes_read.sum"]]]
    wr = float(d[idx["dram__bytes_write.sum"]]) * SCALE[units[idx["dram__bytes_write.sum"]]]
    us = float(d[idx["gpu__time_duration.sum"]]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units[idx["gpu__time_duration.sum"]], 1.0)
    json.dump({"kernel": name.split("(")[0], "samples_per_launch": samples, "dram_bytes_read": rd, "dram_bytes_write": wr,
               "duration_us_under_ncu": us, "source": path, "how": "ncu --set full --clock-control none, one launch"}, open(out, "w"), indent=1)
    print(open(out).read())
    break
else:
    sys.exit("no kernel matching %r in %s" % (filt, path))
