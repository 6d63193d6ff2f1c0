"""Write profiles/k1_traffic.json (DRAM bytes of one K1 launch per precision) from `ncu --set full` captures of
tools/k1_only.py.  usage: python tools/k1_traffic.py bf16x3=rep1.ncu-rep bf16=rep2.ncu-rep"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for arg in sys.argv[1:]:
    prec, rep = arg.split("=")
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, r = rows[0], rows[1], rows[-1]
    get = lambda m: float(r[hdr.index(m)]) * UNIT[units[hdr.index(m)]]  # noqa: E731
    out[prec] = dict(dram_bytes_read=get("dram__bytes_read.sum"), dram_bytes_write=get("dram__bytes_write.sum"),
                     kernel=r[hdr.index("Kernel Name")].split("(")[0], source="profiles/" + os.path.basename(rep).replace(".ncu-rep", ".txt"))
with open(os.path.join(ROOT, "profiles", "k1_traffic.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out, indent=1))
