"""Summarise an `ncu --set full` report (.ncu-rep) into a small text table for profiles/ (read here, without a GPU).
usage: python tools/ncu_summary.py report.ncu-rep [title] > profiles/rNN_xxx.txt"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%act"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_%elapsed"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2->sm"),
    ("l1tex__m_l1tex2xbar_write_bytes.sum", "sm->l2"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"),
    ("launch__registers_per_thread", "regs"),
    ("sm__cycles_elapsed.max", "cycles"),
]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("# %s" % title)
    print("# ncu --set full --clock-control none; one line per captured launch; values as printed by ncu (unit in [])")
    cols = [(hdr.index(m), n, units[hdr.index(m)]) for m, n in METRICS if m in hdr]
    ki, gi = hdr.index("Kernel Name"), hdr.index("Grid Size")
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("hm::", "").replace("<unnamed>::", "")
        print("%-28s grid=%-14s " % (name[:28], r[gi]) + "  ".join("%s=%.4g[%s]" % (n, float(r[i]), u) for i, n, u in cols if r[i]))


if __name__ == "__main__":
    main()
