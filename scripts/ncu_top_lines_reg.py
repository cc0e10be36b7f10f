#!/usr/bin/env python
"""Top source lines of evolve_reg.cu (and of the headers it inlines) by warp instructions executed, from a .ncu-rep captured with
--import-source on; plus the opcode mix.  usage: ncu_top_lines_reg.py <rep> [source file as captured]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
src = open(sys.argv[2] if len(sys.argv) > 2 else "camb_b200/csrc/cuda/evolve_reg.cu").read().splitlines()
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[2]
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
S, I = ix['# Samples'], ix['Instructions Executed']
own, hdrs, ops = collections.defaultdict(lambda: [0, 0]), collections.defaultdict(lambda: [0, 0]), collections.Counter()
tot = tot_s = 0
for r in rows[3:]:
    if len(r) < len(hdr):
        continue
    try:
        s, i = int(r[S] or 0), int(r[I] or 0)
    except ValueError:
        continue
    if r[0].isdigit():
        ln, text = int(r[0]), r[1].strip()
        tot += i; tot_s += s
        tgt = own[(ln, text[:100])] if ln <= len(src) and src[ln - 1].strip()[:60] == text[:60] else hdrs[text[:100]]
        tgt[0] += i; tgt[1] += s
    else:
        t = r[3].strip().split()
        if t:
            ops[(t[1] if t[0].startswith('@') and len(t) > 1 else t[0]).split('.')[0]] += i
print("all three launches: %.4e warp instructions, %d samples" % (tot, tot_s))
print("--- evolve_reg.cu lines")
for k, v in sorted(own.items(), key=lambda kv: -kv[1][0])[:30]:
    print("%5.2f%% instr %5.2f%% samples  %d: %s" % (100 * v[0] / tot, 100 * v[1] / tot_s, k[0], k[1]))
print("--- inlined headers (line attribution of out-of-line math subroutines is the compiler's)")
for k, v in sorted(hdrs.items(), key=lambda kv: -kv[1][0])[:10]:
    print("%5.2f%% instr %5.2f%% samples  %s" % (100 * v[0] / tot, 100 * v[1] / tot_s, k))
print("--- opcode mix")
n = sum(ops.values())
print(", ".join("%s %.1f%%" % (k, 100 * v / n) for k, v in ops.most_common(18)))
