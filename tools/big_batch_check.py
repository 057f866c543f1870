"""Scale sanity check on the GPU box: a batch well beyond BASELINE config 2 (device-resident API),
8-channel and mono shapes; verifies the exact round trip and prints throughput."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from sela_b200 import synth
from sela_b200.device import DeviceCodec

def run(channels, n_frames, rate, tile=1):
    base = synth.sine_noise(rate, channels, n_frames=n_frames // tile, seed=3)
    pcm_np = np.tile(base.reshape(-1), tile)
    n_frames = pcm_np.size // (2048 * channels)
    pcm = torch.from_numpy(pcm_np).cuda()
    out = torch.empty_like(pcm)
    codec = DeviceCodec(n_frames, channels)
    codec.encode(pcm); torch.cuda.synchronize(); codec.check_status()
    n_words = int(codec.words_used.item())
    codec.decode(out, n_words); torch.cuda.synchronize(); codec.check_status()
    ok = bool(torch.equal(out, pcm))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record(); codec.encode(pcm); ev[1].record(); codec.decode(out, n_words); ev[2].record()
    torch.cuda.synchronize()
    te, td = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    n = pcm_np.size
    print("ch=%d frames=%d samples=%.1fM  encode %.2f ms (%.1f GS/s)  decode %.2f ms (%.1f GS/s)  bits/sample %.2f  round-trip %s" % (
        channels, n_frames, n / 1e6, te, n / te / 1e6, td, n / td / 1e6, n_words * 32 / n, ok))
    assert ok

run(2, 12919 * 12, 44100, tile=12)     # 155k stereo frames (0.63 G samples)
run(8, 10547, 48000)                   # BASELINE config 4, one GPU's eighth
run(1, 40000, 44100, tile=4)           # mono
