"""Seeded test-signal families shared by the CPU and GPU parity tests."""
import numpy as np

from sela_b200 import synth

N = 2048


def families(seed=0):
    """name -> int32[2048] in int16 range.  Deterministic."""
    rng = np.random.default_rng(seed)
    i = np.arange(N, dtype=np.float64)
    out = {}
    out["sine_deg"] = synth.config1_frame()                                  # test/lpctests.cpp:16-18
    out["cosine_deg"] = (32767 * np.cos(i * (np.pi / 180))).astype(np.int32)
    out["zeros"] = np.zeros(N, np.int32)
    out["dc_1234"] = np.full(N, 1234, np.int32)
    out["dc_neg"] = np.full(N, -20000, np.int32)
    imp = np.zeros(N, np.int32); imp[0] = 32767
    out["impulse0"] = imp
    imp2 = np.zeros(N, np.int32); imp2[1000] = -32768
    out["impulse_mid"] = imp2
    out["alt_fullscale"] = np.where(np.arange(N) % 2 == 0, 32767, -32768).astype(np.int32)
    out["ramp"] = (np.arange(N) * 16 - 16384).astype(np.int32)
    out["white_full"] = rng.integers(-32768, 32768, N).astype(np.int32)
    out["white_small"] = rng.integers(-3, 4, N).astype(np.int32)
    out["sine_noise"] = synth.sine_noise(44100, 1, n_frames=1, seed=seed + 7)[:, 0].astype(np.int32)
    out["two_tone"] = np.rint(12000 * np.sin(i * 0.05) + 9000 * np.sin(i * 1.3 + 1)).astype(np.int32)
    out["ill_cond"] = np.rint(30000 * np.sin(i * 0.001)).astype(np.int32)    # nearly DC sine, ill-conditioned
    out["chirp"] = np.rint(20000 * np.sin(i * i * 2e-4)).astype(np.int32)
    out["ar1"] = None
    e = rng.standard_normal(N) * 800
    a = np.zeros(N)
    for j in range(1, N):
        a[j] = 0.97 * a[j - 1] + e[j]
    out["ar1"] = np.clip(np.rint(a), -32768, 32767).astype(np.int32)
    out["step"] = np.where(np.arange(N) < 700, -15000, 15000).astype(np.int32)
    out["square"] = np.where((np.arange(N) // 37) % 2 == 0, 9000, -9000).astype(np.int32)
    out["last_sample_only"] = np.zeros(N, np.int32); out["last_sample_only"][-1] = 5
    return out


def random_frames(count, seed):
    """count random subframes drawn from several signal models (int32[count, 2048])."""
    rng = np.random.default_rng(seed)
    i = np.arange(N, dtype=np.float64)
    out = np.zeros((count, N), np.int32)
    for f in range(count):
        kind = f % 6
        if kind == 0:
            v = rng.uniform(500, 30000) * np.sin(i * rng.uniform(0.001, 3.0) + rng.uniform(0, 6.28)) \
                + rng.standard_normal(N) * rng.uniform(0, 2000)
        elif kind == 1:
            v = rng.standard_normal(N) * rng.uniform(1, 9000)
        elif kind == 2:
            e = rng.standard_normal(N) * rng.uniform(10, 1500)
            pole = rng.uniform(-0.99, 0.99)
            v = np.zeros(N)
            for j in range(1, N):
                v[j] = pole * v[j - 1] + e[j]
        elif kind == 3:
            v = np.full(N, float(rng.integers(-32768, 32768)))
        elif kind == 4:
            v = sum(rng.uniform(100, 6000) * np.sin(i * rng.uniform(0.01, 3.1) + rng.uniform(0, 6.28))
                    for _ in range(5))
        else:
            v = rng.integers(-32768, 32768, N).astype(np.float64) * (rng.random(N) < rng.uniform(0.001, 0.2))
        out[f] = np.clip(np.rint(v), -32768, 32767).astype(np.int32)
    return out
