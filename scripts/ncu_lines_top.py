#!/usr/bin/env python
"""Source lines of one kernel of a .ncu-rep (captured with --import-source on) ranked by warp instructions executed.
usage: ncu_lines_top.py <rep> <kernel-name substring> [min share, default 0.0035]"""
import csv, collections, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0035
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Line No")
I, S = hdr.index("Instructions Executed"), hdr.index("# Samples")
def num(x):
    try: return int(x)
    except ValueError: return 0
lines, f, fn = [], "", ""
for r in rows:
    if r and r[0] == "File Path": f = r[1]; continue
    if r and r[0] == "Function Name": fn = r[1]; continue
    if len(r) > 10 and r[0].isdigit() and pat in fn:
        lines.append((f.split('/')[-1], int(r[0]), r[1].strip(), num(r[I]), num(r[S])))
tot = sum(l[3] for l in lines) or 1; tots = sum(l[4] for l in lines) or 1
print('total warp instructions %.3e' % tot)
bf = collections.Counter()
for l in lines: bf[l[0]] += l[3]
print({k: '%.1f%%' % (100 * v / tot) for k, v in bf.items()})
for l in sorted(lines, key=lambda l: (l[0], l[1])):
    if l[3] / tot > thr:
        print('%-12s %4d %5.2f%% %5.2f%%s | %s' % (l[0][:12], l[1], 100 * l[3] / tot, 100 * l[4] / tots, l[2][:120]))
