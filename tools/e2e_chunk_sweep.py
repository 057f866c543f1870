"""Time the host-buffer encode / decode calls separately for several chunk sizes (GPU box)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, ".")
from sela_b200 import _lib, synth
pcm = synth.sine_noise(44100, 2, seconds=600, seed=1)
n_frames = pcm.shape[0] // 2048
L = _lib.lib(); _lib.init(0)
def pinned(nbytes, dtype):
    p = L.selab200_host_alloc(nbytes)
    return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p)).view(dtype)
h_pcm = pinned(pcm.size * 2, np.int16); h_pcm[:] = pcm.reshape(-1)
h_out = pinned(pcm.size * 2, np.int16)
cap = L.selab200_encode_words_bound(n_frames, 2)
h_words = pinned(cap * 4, np.uint32); h_descs = pinned(n_frames * 2 * 32, np.uint8)
used = C.c_size_t(0)
for cf in [int(a) for a in sys.argv[1:]] or [808, 1615, 3230, 6460, 12919]:
    os.environ["SELAB200_CHUNK_FRAMES"] = str(cf)
    te, td = [], []
    for it in range(8):
        t0 = time.perf_counter()
        _lib.check(L.selab200_encode_frames(h_pcm.ctypes.data, n_frames, 2, h_descs.ctypes.data, h_words.ctypes.data, cap, C.addressof(used)))
        t1 = time.perf_counter()
        _lib.check(L.selab200_decode_frames(h_descs.ctypes.data, n_frames, 2, h_words.ctypes.data, used.value, h_out.ctypes.data))
        t2 = time.perf_counter()
        if it >= 2: te.append(t1 - t0); td.append(t2 - t1)
    print("chunk %6d frames: encode %.2f ms  decode %.2f ms  total %.2f ms  ok=%s" % (
        cf, 1e3 * np.median(te), 1e3 * np.median(td), 1e3 * (np.median(te) + np.median(td)), np.array_equal(h_out, h_pcm)))
