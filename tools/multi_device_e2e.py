"""Host-buffer timings of the C ABI with one device and with several (selab200_init_devices):
the same pinned buffers, the same calls.  Usage: python tools/multi_device_e2e.py [n_devices] [minutes]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import numpy as np

import sela_b200
from sela_b200 import _lib, synth

n_dev = int(sys.argv[1]) if len(sys.argv) > 1 else 2
minutes = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
channels = 2
pcm = synth.sine_noise(44100, channels, seconds=60.0 * minutes, seed=1)
n_frames = pcm.shape[0] // 2048
pcm = pcm[: n_frames * 2048]
out = {"n_frames": n_frames, "channels": channels}
ref = {}
for devs in ([0], list(range(n_dev))):
    _lib.init(devs if len(devs) > 1 else devs[0])
    t = {}
    for name, fn in (
        ("encode_container", lambda: sela_b200.encode_container(pcm, channels, 44100, device=devs if len(devs) > 1 else 0)),
        ("decode_container", None),
    ):
        if fn is None:
            blob = ref["blob"]
            fn = lambda: sela_b200.decode_container(blob, device=devs if len(devs) > 1 else 0)
        best = 1e9
        for _ in range(6):
            t0 = time.perf_counter()
            r = fn()
            best = min(best, time.perf_counter() - t0)
        t[name] = best * 1e3
        if name == "encode_container":
            if "blob" in ref:
                assert ref["blob"].tobytes() == r.tobytes()
            ref["blob"] = r
        else:
            if "pcm" in ref:
                assert np.array_equal(ref["pcm"], r[1])
            ref["pcm"] = r[1]
    out["devices_%d_ms" % len(devs)] = t
print(json.dumps(out))
