"""Minimal WAV/SELA file helpers for tests and tools (host-side Python, no codec logic).

write_wav produces the canonical 44-byte-header layout the reference re-writes
(src/file/wav_file.cpp:7-37, 222-243); extra chunks can be injected to exercise the parser
(src/file/wav_file.cpp:153-167)."""
import struct

import numpy as np


def write_wav(path, pcm, sample_rate, extra_chunks=()):
    pcm = np.ascontiguousarray(pcm, dtype="<i2")
    channels = pcm.shape[1] if pcm.ndim == 2 else 1
    data = pcm.tobytes()
    extras = b"".join(cid + struct.pack("<I", len(body)) + body for cid, body in extra_chunks)
    fmt = struct.pack("<HHIIHH", 1, channels, sample_rate, sample_rate * channels * 2, channels * 2, 16)
    body = b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt + extras + b"data" + struct.pack("<I", len(data)) + data
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)


def read_wav_pcm(path):
    """Returns (sample_rate, channels, int16 array [n, channels]) of a canonical 44-byte-header WAV."""
    raw = open(path, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE" and raw[12:16] == b"fmt " and raw[36:40] == b"data"
    channels, rate = struct.unpack("<H", raw[22:24])[0], struct.unpack("<I", raw[24:28])[0]
    n = struct.unpack("<I", raw[40:44])[0]
    return rate, channels, np.frombuffer(raw[44:44 + n], dtype="<i2").reshape(-1, channels)


SELA_SYNC = b"\x00\xff\x55\xaa"  # 0xAA55FF00 little endian (src/include/data/sela_frame.hpp)


def pack_container(descs, words, sample_rate, channels, bits_per_sample=16):
    """(descs, words) of whole frames -> the bytes file::SelaFile::writeToFile emits
    (src/file/sela_file.cpp:105-137).  Plain byte shuffling for tests and tools."""
    words = np.ascontiguousarray(words, dtype="<u4")
    n_frames = len(descs) // channels
    out = [b"SeLa" + struct.pack("<IHBI", sample_rate, bits_per_sample, channels, n_frames)]
    for f in range(n_frames):
        out.append(SELA_SYNC)
        for d in descs[f * channels:(f + 1) * channels]:
            a, b = int(d["refl_offset"]), int(d["res_offset"])
            out.append(struct.pack("<BBBBHB", d["channel"], d["subframe_type"], d["parent_channel"],
                                   d["refl_rice_param"], d["refl_words"], d["lpc_order"]))
            out.append(words[a:a + int(d["refl_words"])].tobytes())
            out.append(struct.pack("<BHH", d["res_rice_param"], d["res_words"], d["samples"]))
            out.append(words[b:b + int(d["res_words"])].tobytes())
    return b"".join(out)
