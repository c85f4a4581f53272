#!/usr/bin/env python
"""Record the DRAM traffic of one kernel launch from an ncu --set full report into profiles/traffic.json
(bench.py's roofline.traffic reads it when the workload size matches).

    python profiles/traffic_from_ncu.py gpurun_out/x.ncu-rep gmm_score_sv_kernel 10000 1000 1024 profiles/x.txt

The report must hold exactly the launches of ONE scoring step (ncu -s / -c).
"""
import csv
import io
import json
import os
import subprocess
import sys

rep, kernel, utts, speakers, comps, source = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
total, n_launches = 0.0, 0
for r in rows[2:]:  # every captured launch of the kernel: one scoring step may be several launches (model groups)
    d = dict(zip(hdr, r))
    if kernel not in d.get("Kernel Name", ""):
        continue
    n_launches += 1
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        total += float(r[i].replace(",", "")) * scale[units[i]]
frames = utts * 298
# algorithmic bytes of the scoring launch: features in + model images once + scores out
# (FP16 images of 64 components x 48: the speakers, the UBM as one more mean set, and the 4 images of the common part)
ks, kp = 48, (comps + 63) // 64 * 64
algo = frames * 39 * 4 + (kp // 64) * (speakers + 1 + 4) * 64 * ks * 2 + utts * (speakers + 1) * 8
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[kernel] = {"launches_per_step": n_launches, "utts": utts, "speakers": speakers, "components": comps, "dram_bytes": int(total), "algorithmic_bytes": int(algo),
                "source": source}
json.dump(data, open(path, "w"), indent=1)
print(json.dumps(data[kernel]))
