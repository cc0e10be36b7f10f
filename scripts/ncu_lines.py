#!/usr/bin/env python
"""Top CUDA source lines by warp-stall samples from a .ncu-rep captured with --import-source on."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[2]
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
S, I, T = ix['# Samples'], ix['Instructions Executed'], ix['Thread Instructions Executed']
stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
lines = []
for r in rows[3:]:
    if len(r) < len(hdr) or not r[0]:
        continue
    try:
        lines.append((int(r[S] or 0), int(r[I] or 0), int(r[T] or 0), int(r[0]), r[1].strip()[:100], r))
    except ValueError:
        pass
tot = sum(l[0] for l in lines); toti = sum(l[1] for l in lines)
print('total samples', tot, 'warp instructions', toti)
agg = {}
for h, i in stall_cols:
    agg[h] = sum(int(l[5][i] or 0) for l in lines)
print('stall mix:', ', '.join(f"{h[6:]} {100*v/tot:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.005 * tot))
for s, i, t, ln, src, r in sorted(lines, reverse=True)[:top]:
    st = sorted(((int(r[c] or 0), h[6:]) for h, c in stall_cols), reverse=True)[:3]
    print(f"{100*s/tot:5.1f}% smp {100*i/toti:5.1f}% ins thr/ins {t/max(i,1):4.1f} L{ln}: {src}   [{', '.join(f'{n}:{v}' for v, n in st)}]")
