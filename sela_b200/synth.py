"""Synthetic PCM of the shapes BASELINE.json names.

s[n,c] = clip(round(0.8*32767*(0.5*sin(2*pi*440*(c+1)*n/fs + 0.1*c) + 0.05*N(0,1)))),
NumPy default_rng(seed); seeds: config 2/3 -> 1, config 4 -> 2, config 5 -> file index.
Returned interleaved int16 [samples_per_channel, channels], truncated to whole
2048-sample frames (the reference drops the tail, src/file/wav_file.cpp:184).
"""
import numpy as np

FRAME = 2048


def sine_noise(sample_rate, channels, seconds=None, seed=1, n_frames=None, chunk=1 << 20):
    if n_frames is None:
        n_frames = int(sample_rate * seconds) // FRAME
    n = n_frames * FRAME
    rng = np.random.default_rng(seed)
    out = np.empty((n, channels), np.int16)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        t = np.arange(lo, hi, dtype=np.float64)[:, None]
        c = np.arange(channels, dtype=np.float64)[None, :]
        v = 0.5 * np.sin(2 * np.pi * 440.0 * (c + 1) * t / sample_rate + 0.1 * c)
        v = v + 0.05 * rng.standard_normal((hi - lo, channels))
        out[lo:hi] = np.clip(np.rint(0.8 * 32767 * v), -32768, 32767).astype(np.int16)
    return out


def config1_frame():
    """The reference's own test input (test/lpctests.cpp:16-18): (int32)(32767*sin(i*pi/180))."""
    i = np.arange(FRAME, dtype=np.float64)
    return (32767 * np.sin(i * (np.pi / 180))).astype(np.int32)
