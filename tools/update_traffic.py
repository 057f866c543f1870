#!/usr/bin/env python3
"""profiles/traffic.json from an ncu --set full capture of bench.py: DRAM bytes per launch, FP64-pipe and issue
utilisation of k_encode_units<stereo>, keyed by a hash of the kernel sources (bench.py prints traffic_stale when
the sources have changed since).   usage: tools/update_traffic.py gpurun_out/prof_TAG.ncu-rep profiles/TAG_ncu_full_summary.txt"""
import csv, hashlib, json, pathlib, subprocess, sys
ROOT = pathlib.Path(__file__).resolve().parent.parent
rep, where = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
def get(r, k):
    return float(r[h.index(k)]) * mul.get(units[h.index(k)], 1)
hs = hashlib.sha256()
for f in sorted((ROOT / "sela_b200" / "csrc").glob("*")):
    hs.update(f.name.encode())
    hs.update(f.read_bytes())
doc = {"source_hash": hs.hexdigest()[:16], "capture": "%s (ncu --set full, one B200, bench.py --steps 2 --warmup 3)" % where}
for r in rows[2:]:
    if "k_encode_units" in r[h.index("Kernel Name")] and "k_encode_units<stereo>" not in doc:
        rd, wr = get(r, "dram__bytes_read.sum"), get(r, "dram__bytes_write.sum")
        doc["k_encode_units<stereo>"] = {
            "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "traffic": int(rd + wr),
            "fp64_busy_frac": get(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") / 100,
            "issue_active_frac": get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") / 100,
            "duration_ms": float(r[h.index("gpu__time_duration.sum")]),
            "workload": "BASELINE config 2 (12919 stereo frames), one launch"}
json.dump(doc, open(ROOT / "profiles" / "traffic.json", "w"), indent=1)
print(json.dumps(doc, indent=1))
