#!/usr/bin/env python
"""Markdown table of the roofline-relevant counters of every kernel in an `ncu --set full` report.
usage: ncu_full_summary.py X.ncu-rep [...]   (runs `ncu -i X --page raw --csv` here; no GPU needed)"""
import csv
import io
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]


def main():
    for rep in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        print("### %s\n" % rep.split("/")[-1])
        print("| kernel | " + " | ".join(c[1] for c in COLS) + " |")
        print("|---|" + "---|" * len(COLS))
        kn = hdr.index("Kernel Name")
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("fcb::", "").replace("tc::", "")
            t = re.search(r"k_aggregate<\(int\)(\d), \(bool\)(\d)>", r[kn])
            if t:
                name = "k_aggregate<B=%s,T=%s>" % t.groups()
            cells = []
            for key, _ in COLS:
                if key not in hdr:
                    cells.append("-")
                    continue
                i = hdr.index(key)
                v, u = r[i].replace(",", ""), units[i]
                try:
                    f = float(v)
                    if u in ("Mbyte", "Gbyte", "Kbyte", "byte"):
                        f *= {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u]
                        cells.append("%.1f MB" % (f / 1e6))
                    elif u in ("ns", "us", "ms"):
                        cells.append("%.1f" % (f * {"ns": 1e-3, "us": 1, "ms": 1e3}[u]))
                    else:
                        cells.append(("%.1f" % f) if f != int(f) else "%d" % f)
                except ValueError:
                    cells.append(v)
            print("| %s | " % name + " | ".join(cells) + " |")
        print()


if __name__ == "__main__":
    main()
