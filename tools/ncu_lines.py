#!/usr/bin/env python3
"""Summarise an ncu report per CUDA source line: samples (stall sampling) and instructions executed.

usage: tools/ncu_lines.py report.ncu-rep kernel_regex [top_n]
"""
import csv, subprocess, sys, collections

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
lines = collections.OrderedDict()
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 8:
        continue
    if r[0] != "":   # a CUDA line summary row
        try:
            samples = int(r[6]); inst = int(r[7])
        except ValueError:
            continue
        key = (cur_file, int(r[0]))
        s = lines.setdefault(key, [0, 0, r[1].strip()[:100]])
        s[0] += samples; s[1] += inst
tot_s = sum(v[0] for v in lines.values()) or 1
tot_i = sum(v[1] for v in lines.values()) or 1
print("total samples %d, total warp-instructions %d" % (tot_s, tot_i))
for (f, ln), (s, i, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% ins  %-14s:%-4d %s" % (100.0 * s / tot_s, 100.0 * i / tot_i, f, ln, src))
