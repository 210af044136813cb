#!/usr/bin/env python
"""Opcode histogram of the largest backward-branch loop of one SASS function (static count).
usage: sass_loop_hist.py obj mangled_name"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", "-fun", sys.argv[2], sys.argv[1]], capture_output=True, text=True).stdout
ins = []
for l in out.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for a, t in ins:
    m = re.search(r"BRA\S*\s+(?:\S+,\s+)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        loops.append((int(m.group(1), 16), a))
print("instructions:", len(ins), "loops:", [(hex(a), hex(b)) for a, b in loops])
lo, hi = max(loops, key=lambda ab: ab[1] - ab[0])
inner = [(a, b) for a, b in loops if lo < a and b < hi]
def hist(lo, hi, excl):
    c = collections.Counter()
    for a, t in ins:
        if lo <= a <= hi and not any(x <= a <= y for x, y in excl):
            tt = t.split()
            op = tt[1] if tt[0].startswith("@") else tt[0]
            c["IMAD.MOV" if "IMAD.MOV" in t else op.split(".")[0]] += 1
    return c
c = hist(lo, hi, inner)
print("outer loop %s-%s minus inner loops: %d instrs" % (hex(lo), hex(hi), sum(c.values())))
print(c.most_common())
for a, b in inner:
    ci = hist(a, b, [])
    print("inner loop %s-%s: %d instrs" % (hex(a), hex(b), sum(ci.values())), ci.most_common(6))
