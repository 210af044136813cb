#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list: keeps (id, short kernel name, grid, block, ns)
per launch and appends per-kernel totals with their share of the captured time.
usage: condense_launch_list.py in.csv out.csv [first_id]   (first_id: skip launches before the timed steps)"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(.*", "", name)              # drop the argument list
    name = re.sub(r"<.*", "", name)               # and template arguments
    return name.replace("void ", "").replace("fcb::", "").replace("tc::", "").strip()[:60]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
    tot = collections.Counter()
    cnt = collections.Counter()
    with open(dst, "w") as f:
        f.write("id,kernel,grid,block,ns\n")
        for r in rows:
            if int(r[0]) < first:
                continue
            full = r[4]
            tmpl = re.search(r"k_aggregate<\(int\)(\d), \(bool\)(\d)>", full)
            k = short(full) + ("<B=%s,T=%s>" % tmpl.groups() if tmpl else "")
            ns = int(float(r[14]))
            f.write("%s,%s,%s,%s,%d\n" % (r[0], k, r[8].replace(",", " "), r[7].replace(",", " "), ns))
            tot[k] += ns
            cnt[k] += 1
        all_ns = sum(tot.values()) or 1
        f.write("# per-kernel totals (cold-cache, serialised: compare shares, not absolutes)\n")
        f.write("# kernel,launches,total_us,share\n")
        for k, v in tot.most_common():
            f.write("# %s,%d,%.1f,%.4f\n" % (k, cnt[k], v / 1e3, v / all_ns))


if __name__ == "__main__":
    main()
