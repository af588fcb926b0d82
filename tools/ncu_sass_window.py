#!/usr/bin/env python
"""Print a window of SASS (offsets relative to the kernel start) with stall samples from an
`ncu --page source --csv --print-source cuda,sass` export: ncu_sass_window.py file.csv lo hi"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
hdr = None; cur = None; insts = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name": continue
    if r[2] == "-": line = int(r[0]); continue
    if not r[2].startswith("0x"): continue
    d = dict(zip(hdr[2:], r[2:]))
    insts[int(r[2], 16)] = (cur.split('/')[-1], line, r[3], d)
A = sorted(insts); base = A[0]
for a in A:
    if lo <= a - base <= hi:
        f, l, s, d = insts[a]
        st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not" not in k and v and int(v) > 0}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print("%6x %-16s:%4d %-72s ex=%-7s smp=%-5s %s" % (a - base, f[:16], l, s[:72], d['Instructions Executed'], d['# Samples'], top))
