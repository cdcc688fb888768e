#!/usr/bin/env python3
"""profiles/dominant_kernel_traffic.json from an `ncu --set full` capture of ONE half-batch launch sequence:
dram__bytes_read.sum + dram__bytes_write.sum of every stage kernel, per frame (feeds bench.py's roofline.traffic).
usage: traffic_json.py <file.ncu-rep> <frames in the captured launch sequence> [config name]
source_hash = sha256 of csrc/ at the time of writing: run it on the tree the capture was taken from; bench.py reports
roofline.traffic = null when the hash on file is not the hash of the tree it runs."""
import csv, json, os, subprocess, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import source_hash
rep, frames = sys.argv[1], int(sys.argv[2])
name = sys.argv[3] if len(sys.argv) > 3 else "rgbd_1080p"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units, data = rows[0], rows[1], rows[2:]
ki, ir, iw = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
def to_bytes(v, u):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
per = collections.OrderedDict()
first = data[0][ki] if data else None
for n_used, r in enumerate(data):
    if n_used > 0 and r[ki] == first:            # the capture window ran into the next launch sequence: one sequence only
        data = data[:n_used]
        break
for r in data:
    k = r[ki].split("(")[0].replace("void ", "").replace("orbx::", "")
    per[k] = per.get(k, 0.0) + to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
total = sum(per.values())
doc = {name: {"bytes_per_frame": total / frames, "source_hash": source_hash(),
              "per_kernel_MB_per_%d_frames" % frames: {k: v / 1e6 for k, v in per.items()},
              "source": "%s: dram__bytes_read.sum + dram__bytes_write.sum of the %d launches of one %d-frame half-batch (ncu --set full)" % (
                  rep.split("/")[-1], len(data), frames)}}
print(json.dumps(doc, indent=1))
