"""Host-buffer timings of the C ABI with one device and with several (selab200_init_devices): the calls and the
pinned buffers bench.py's end-to-end leg uses (selab200_encode_frames / selab200_decode_frames, plus the container
pair), on the BASELINE file.  Outputs of the two configurations are compared byte for byte.
Usage: python tools/multi_device_e2e.py [n_devices] [minutes]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sela_b200 import _lib, synth  # noqa: E402

n_dev = int(sys.argv[1]) if len(sys.argv) > 1 else 2
minutes = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
CH = 2
pcm = synth.sine_noise(44100, CH, seconds=60.0 * minutes, seed=1)
n_frames = pcm.shape[0] // 2048
n_samples = n_frames * 2048 * CH
L = _lib.lib()


def pinned(nbytes, dtype):
    p = L.selab200_host_alloc(nbytes)
    return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p)).view(dtype)


def best(fn, reps=8):
    fn(), fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    return min(t) * 1e3


out = {"n_frames": n_frames, "channels": CH, "minutes": minutes}
keep = {}
for devs in ([0], list(range(n_dev))):
    _lib.init(devs if len(devs) > 1 else devs[0])
    h_pcm = pinned(n_samples * 2, np.int16)
    h_pcm[:] = pcm[: n_frames * 2048].reshape(-1)
    h_out = pinned(n_samples * 2, np.int16)
    cap = L.selab200_encode_words_bound(n_frames, CH)
    h_words = pinned(cap * 4, np.uint32)
    h_descs = pinned(n_frames * CH * 32, np.uint8)
    ccap = L.selab200_container_bound(n_frames, CH)
    h_cont = pinned(ccap, np.uint8)
    used, cused = C.c_size_t(0), C.c_size_t(0)
    t = {}
    t["encode_frames_ms"] = best(lambda: _lib.check(L.selab200_encode_frames(
        h_pcm.ctypes.data, n_frames, CH, h_descs.ctypes.data, h_words.ctypes.data, cap, C.addressof(used))))
    t["decode_frames_ms"] = best(lambda: _lib.check(L.selab200_decode_frames(
        h_descs.ctypes.data, n_frames, CH, h_words.ctypes.data, used.value, h_out.ctypes.data)))
    assert np.array_equal(h_out, h_pcm)
    t["encode_container_ms"] = best(lambda: _lib.check(L.selab200_encode_container(
        h_pcm.ctypes.data, n_frames, CH, 44100, 16, h_cont.ctypes.data, ccap, C.addressof(cused))))

    def decode_container():
        h = C.c_void_p(0)
        info = np.zeros(64, np.uint8)
        _lib.check(L.selab200_container_open(h_cont.ctypes.data, cused.value, C.addressof(h), info.ctypes.data))
        rc = L.selab200_container_decode(h, h_out.ctypes.data)
        L.selab200_container_close(h)
        _lib.check(rc)

    h_out[:] = 0
    t["open_decode_container_ms"] = best(decode_container)
    assert np.array_equal(h_out, h_pcm)
    got = (h_descs.tobytes(), h_words[: used.value].tobytes(), h_cont[: cused.value].tobytes())
    if keep:
        assert got == keep["got"], "several devices must produce the bytes one device produces"
    keep["got"] = got
    t["gsamples_s_encode_frames"] = n_samples / t["encode_frames_ms"] / 1e6
    t["gsamples_s_decode_frames"] = n_samples / t["decode_frames_ms"] / 1e6
    out["devices_%d" % len(devs)] = t
    for a in (h_pcm, h_out, h_words, h_descs, h_cont):
        L.selab200_host_free(a.ctypes.data)
print(json.dumps(out))
