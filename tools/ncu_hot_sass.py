"""Top warp-stall-sampled SASS instructions of one kernel from an `ncu --set full --import-source on` report.
usage: python tools/ncu_hot_sass.py report.ncu-rep kernel_regex [top_n] [launch_index]"""
import csv
import io
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], text=True,
                              stderr=subprocess.DEVNULL)
# the export concatenates one table per launch, each starting with a "Kernel Name" line
blocks = out.split('"Kernel Name",')[1:]
blk = blocks[which]
rows = list(csv.reader(io.StringIO(blk)))
print("kernel:", rows[0][0][:100], "(launch %d of %d)" % (which, len(blocks)))
hdr = rows[1]
si, so, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
data = []
for i, r in enumerate(rows[2:]):
    if len(r) > max(si, ie) and r[si].isdigit():
        data.append((int(r[si]), i, int(r[ie] or 0), r[so].strip()))
tot = sum(d[0] for d in data) or 1
print("total samples", tot)
for n, i, e, s in sorted(data, reverse=True)[:top]:
    print("%6d %5.1f%%  sass#%4d  exec %9d  %s" % (n, 100.0 * n / tot, i, e, s[:100]))
