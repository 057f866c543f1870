"""A/B an environment switch on the end-to-end leg of bench.py:  python tools/e2e_ab.py VAR a b [reps]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
var, vals, reps = sys.argv[1], sys.argv[2:4], int(sys.argv[4]) if len(sys.argv) > 4 else 2
for rep in range(reps):
    for v in vals:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "40", "--warmup", "3", "--no-cpu", "--no-sharded"],
                           env=dict(os.environ, **{var: v}), capture_output=True, text=True, timeout=900)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:
            print(var, v, "FAILED", r.stderr[-300:])
            continue
        e = d["e2e"]
        print("%s=%s  device %.3f ms (encode %.3f + decode %.3f)  e2e %.3f ms (encode %.3f + decode %.3f)  exact %s" % (
            var, v, d["ms_per_step"], d["encode_ms"], d["decode_ms"], e["ms_per_step"], e["encode_ms"], e["decode_ms"],
            d["round_trip_bit_exact"]), flush=True)
