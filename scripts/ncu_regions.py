#!/usr/bin/env python
"""Aggregate the per-source-line instruction and stall-sample counts of an evolve-kernel .ncu-rep by code region of
camb_b200/csrc/cuda/evolve.cu (line ranges) and by file for inlined library code (thermo.cuh, libdevice math)."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
REG = [(323, 409, 'derivs: preamble (background, constraints, tight coupling, vbdot)'),
       (410, 509, 'derivs: photon T/E hierarchy'), (510, 574, 'derivs: neutrino hierarchies'), (575, 731, 'derivs: source/transfer output branch'),
       (732, 758, 'group_max / Ctl'), (759, 1010, 'initial, plan_switches, apply_switch, dp45_begin'), (1011, 1132, 'advance_schedule'),
       (1133, 1183, 'attempt_start (step control, thermo look-ups)'), (1184, 1253, 'attempt_finish (error norm, accept)'),
       (1254, 1310, 'main loop: trip control'), (1311, 1348, 'main loop: stage combinations'), (1349, 1378, 'main loop: post-derivs bookkeeping')]
agg, tot_s, tot_i, tot_w, tot_x = {}, 0, 0, 0, 0
cur_file, hdr, ix = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur_file, hdr = r[1], None
        continue
    if r and r[0] == 'Line No':
        hdr, ix = r, {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        continue
    if hdr is None or len(r) < len(hdr) or not r[0]:
        continue
    try:
        ln, s_, i_ = int(r[0]), int(r[ix['# Samples']] or 0), int(r[ix['Instructions Executed']] or 0)
        w_ = int(r[ix['L1 Wavefronts Shared']] or 0); x_ = int(r[ix['L1 Wavefronts Shared Excessive']] or 0)
    except (ValueError, KeyError):
        continue
    base = (cur_file or '?').split('/')[-1]
    name = base
    if base == 'evolve.cu':
        name = 'evolve.cu: other'
        for a_, b_, n in REG:
            if a_ <= ln <= b_:
                name = 'evolve.cu ' + n
                break
    k = agg.setdefault(name, [0, 0, 0, 0])
    k[0] += s_; k[1] += i_; k[2] += w_; k[3] += x_
    tot_s += s_; tot_i += i_; tot_w += w_; tot_x += x_
print(f'total samples {tot_s}, warp instructions {tot_i}, shared wavefronts {tot_w} (excessive = bank conflicts {tot_x})')
for n, (s_, i_, w_, x_) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{100*i_/max(tot_i,1):5.1f}% instr {100*s_/max(tot_s,1):5.1f}% samples {100*w_/max(tot_w,1):5.1f}% smem wavefronts ({100*x_/max(w_,1):4.1f}% excessive)  {n}')
