#!/usr/bin/env python
"""Attribute the warp-stall samples of an `ncu --page source --csv --print-source cuda,sass`
export to regions of a kernel: SASS instructions in address order, each tagged with the last
line of `anchor file` seen (inlined helpers inherit the region of their call site)."""
import csv, sys, collections
path, anchor = sys.argv[1], sys.argv[2]
bounds = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else []
rows = list(csv.reader(open(path)))
cur_file = None; hdr = None; insts = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if r[2] == "-":
        line = int(r[0]); continue
    d = dict(zip(hdr[2:], r[2:]))
    if not r[2].startswith("0x"): continue
    insts[int(r[2], 16)] = (cur_file, line, r[3], d)
addrs = sorted(insts)
region = collections.OrderedDict()
last = 0
tot = 0
stall_keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
for a in addrs:
    f, line, sass, d = insts[a]
    if f.endswith(anchor): last = line
    key = last
    for b in bounds:
        pass
    reg = region.setdefault(key, collections.Counter())
    n = int(d["# Samples"] or 0)
    reg["samples"] += n; tot += n
    reg["inst"] += int(d["Instructions Executed"] or 0)
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    reg["op_" + op.split(".")[0]] += int(d["Instructions Executed"] or 0)
    for k in stall_keys:
        reg[k] += int(d.get(k) or 0)
# merge into coarse buckets by bounds
def bucket(line):
    for i, b in enumerate(bounds):
        if line < b: return i
    return len(bounds)
agg = collections.OrderedDict()
for line, c in region.items():
    agg.setdefault(bucket(line) if bounds else line, collections.Counter()).update(c)
for k, c in sorted(agg.items()):
    top = sorted(((v, s) for s, v in c.items() if s.startswith("stall_")), reverse=True)[:5]
    ops = sorted(((v, s) for s, v in c.items() if s.startswith("op_")), reverse=True)[:6]
    print("%s samples %6d (%4.1f%%) inst %10d  %s | %s" % (k, c["samples"], 100.0 * c["samples"] / tot, c["inst"],
          " ".join("%s=%d" % (s[6:], v) for v, s in top), " ".join("%s=%d" % (s[3:], v) for v, s in ops)))
print("total samples", tot)
