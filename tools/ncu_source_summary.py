#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` output: per-opcode executed warp-instructions and
warp-stall samples for each kernel section.  usage: ncu_source_summary.py file.csv [units_per_kernel]"""
import collections
import csv
import sys


def sections(path):
    rows = list(csv.reader(open(path)))
    sec, name = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            if sec:
                yield name, sec
            name, sec = r[1], []
        else:
            sec.append(r)
    if sec:
        yield name, sec


def main():
    path = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    seen = set()
    for name, sec in sections(path):
        if name in seen:
            continue
        seen.add(name)
        hdr, data = sec[0], [r for r in sec[1:] if len(r) == len(sec[0])]
        isrc, ie = hdr.index("Source"), hdr.index("Instructions Executed")
        tot = sum(int(r[ie]) for r in data)
        print("=== %s\n  warp instructions executed: %d, SASS lines: %d" % (name[:100], tot, len(data)))
        stall = collections.Counter()
        for r in data:
            for i, h in enumerate(hdr):
                if h.startswith("stall_") and "Not Issued" not in h:
                    stall[h] += int(r[i] or 0)
        ssum = sum(stall.values()) or 1
        print("  stall samples: " + ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / ssum) for k, v in stall.most_common(8)))
        ops = collections.Counter()
        for r in data:
            t = r[isrc].split()
            op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            ops[op.split(".")[0]] += int(r[ie])
        for k, v in ops.most_common(18):
            print("  %-10s %12d %5.1f%%%s" % (k, v, 100.0 * v / tot, ("  per unit %.1f" % (v / units)) if units else ""))
        isamp = hdr.index("# Samples")
        print("  hottest SASS lines by stall samples:")
        for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:14]:
            why = max(((h[6:], int(r[i] or 0)) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h),
                      key=lambda kv: kv[1])
            print("   %7s samples  %-14s %s" % (r[isamp], why[0], r[isrc][:80]))


if __name__ == "__main__":
    main()
