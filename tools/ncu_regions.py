#!/usr/bin/env python3
"""Aggregate executed warp-instructions and stall samples of a kernel by source REGION (function),
using function start lines parsed from the .cuh files.  usage: ncu_regions.py report kernel_regex"""
import csv, subprocess, sys, re, collections, pathlib
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# function boundaries per file
root = pathlib.Path(__file__).resolve().parent.parent / "sela_b200" / "csrc"
bounds = {}
for f in root.glob("*.cu*"):
    marks = []
    for i, line in enumerate(f.read_text().splitlines(), 1):
        m = re.match(r"^(?:template.*\n)?(?:__device__|__global__|inline|static)?.*?\b(warp_\w+|lane_\w+|k_\w+|sample_to_x\w*|dequantise|ring_\w+|zigzag|unzigzag|choose_unit|desc_ok|unpack8|mad_wide_u32|dadd|dmul|ddiv|dsub|dsqrt|shfl_\w+)\s*\(", line)
        if m and not line.strip().startswith("//") and ("__device__" in line or "__global__" in line or "void" in line.split("(")[0] or "int" in line.split("(")[0] or "double" in line.split("(")[0] or "bool" in line.split("(")[0]):
            marks.append((i, m.group(1)))
    bounds[f.name] = marks
def region(fname, ln):
    marks = bounds.get(fname, [])
    name = fname
    for start, n in marks:
        if start <= ln:
            name = n
        else:
            break
    return name
cur = None
agg = collections.defaultdict(lambda: [0, 0])
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No") or len(r) < 8 or r[0] == "": continue
    try: s, i = int(r[6]), int(r[7])
    except ValueError: continue
    a = agg[region(cur, int(r[0]))]; a[0] += s; a[1] += i
ts = sum(v[0] for v in agg.values()) or 1; ti = sum(v[1] for v in agg.values()) or 1
print("total samples %d, warp-instructions %d" % (ts, ti))
for k, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%5.1f%% ins %5.1f%% smp  %s" % (100.0 * i / ti, 100.0 * s / ts, k))
