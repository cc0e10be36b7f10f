#!/usr/bin/env python
"""Instruction / sample shares of evolve_reg.cu by region, from a .ncu-rep captured with --import-source on.
usage: ncu_regions_reg.py <rep> [trips]  (trips: warp-level derivs evaluations, to print instructions per trip)"""
import csv, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[2]
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
S, I = ix['# Samples'], ix['Instructions Executed']
src = open(sys.argv[3] if len(sys.argv) > 3 else "camb_b200/csrc/cuda/evolve_reg.cu").read().splitlines()
# region boundaries from the function starts of evolve_reg.cu
marks = []
for n, line in enumerate(src, 1):
    m = re.match(r"(?:__device__|__global__|template|static|cudaError_t).*?\b(\w+)\(", line)
    if line.startswith("__device__") or line.startswith("__global__"):
        mm = re.search(r"(\w+)\(", line)
        if mm: marks.append((n, mm.group(1)))
    elif "// ---- the uniform sweep" in line: marks.append((n, "derivs_reg: sweep"))
    elif "// ---- rows with metric" in line: marks.append((n, "derivs_reg: special rows"))
    elif "// ---- output branch" in line: marks.append((n, "derivs_reg: output branch"))
    elif "switch (t) {" in line: marks.append((n, "main loop: stage combinations"))
    elif "tev = C.tau + r_dp_c[t] * h;" in line: marks.append((n, "main loop: after stage"))
marks.sort()
def region(ln):
    name = "?"
    for n, nm in marks:
        if n <= ln: name = nm
        else: break
    return name
agg = {}
cur_file = None
tot_i = tot_s = 0
for r in rows[3:]:
    if len(r) < len(hdr) or not r[0].isdigit():
        continue
    ln = int(r[0]); text = r[1]
    try:
        s, i = int(r[S] or 0), int(r[I] or 0)
    except ValueError:
        continue
    # lines of other files (headers) are recognised by not matching the evolve_reg.cu text
    if ln <= len(src) and src[ln - 1].strip()[:60] == text.strip()[:60]:
        key = region(ln)
    else:
        key = "headers (thermo.cuh, intrinsics, math)"
    a = agg.setdefault(key, [0, 0]); a[0] += i; a[1] += s
    tot_i += i; tot_s += s
trips = float(sys.argv[2]) if len(sys.argv) > 2 else None
print("total warp instructions %.4e, samples %d" % (tot_i, tot_s))
for k, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    extra = "  %.1f / trip" % (i / trips) if trips else ""
    print("%5.1f%% instr %5.1f%% samples  %s%s" % (100 * i / tot_i, 100 * s / tot_s, k, extra))
