"""The Rice-decode kernel (K5) in isolation against the HBM roofline, as a function of batch size.

north_star asks for ">= 70 % of HBM roofline on the Rice decode kernel".  The kernel is one lane per
stream; at BASELINE's batch (25 838 streams = 808 warps) it is starved for parallelism.  This sweeps
the number of streams by tiling the coded 10-minute stereo file, and both ring geometries.
Algorithmic bytes: residue words read + 4 B per decoded sample written.
Run on the GPU box:  python tools/rice_decode_roofline.py [max_tile]
"""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from sela_b200 import _lib, synth
from sela_b200.device import DeviceCodec

PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
max_tile = int(sys.argv[1]) if len(sys.argv) > 1 else 48
pcm = synth.sine_noise(44100, 2, seconds=600, seed=1)
n_frames = pcm.shape[0] // 2048
codec = DeviceCodec(n_frames, 2)
codec.encode(torch.from_numpy(pcm.reshape(-1)).cuda()); torch.cuda.synchronize(); codec.check_status()
n_words = int(codec.words_used.item())
descs = codec.descs.cpu().numpy().view(_lib.DESC_DTYPE).copy()
words = codec.words[:n_words].clone()
res_words = int(descs["res_words"].astype(np.int64).sum())
L = _lib.lib()
rows = []
for tile in [t for t in (1, 4, 16, 48, 96) if t <= max_tile]:
    d = np.tile(descs, tile)
    for r in range(tile):
        sl = slice(r * descs.size, (r + 1) * descs.size)
        d["refl_offset"][sl] += r * n_words
        d["res_offset"][sl] += r * n_words
    d_descs = torch.from_numpy(d.view(np.uint8).reshape(-1)).cuda()
    d_words = words.repeat(tile)
    n_sub = d.size
    out = torch.empty(n_sub * 2048, dtype=torch.int32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for ring in (64,):
        os.environ["SELAB200_RICE_RING"] = str(ring)
        def run():
            _lib.check(L.selab200_rice_decode_frames_device(d_descs.data_ptr(), n_frames * tile, 2, d_words.data_ptr(),
                                                            n_words * tile, out.data_ptr(), status.data_ptr(), C.c_void_p(stream)))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        assert int(status.item()) == 0
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = 5
        ev[0].record()
        for _ in range(reps):
            run()
        ev[1].record(); torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        samples = n_sub * 2048
        alg = res_words * tile * 4 + samples * 4
        gbs = alg / ms / 1e6
        rows.append(dict(streams=n_sub, ring=ring, ms=ms, gsamples_s=samples / ms / 1e6, gb_s=gbs, frac_of_hbm_peak=gbs / PEAK))
        print("streams %8d  ring %3d  %8.3f ms  %7.1f GSamples/s  %7.1f GB/s  = %.1f %% of measured HBM peak (%.0f GB/s)" % (
            n_sub, ring, ms, samples / ms / 1e6, gbs, 100 * gbs / PEAK, PEAK))
    if tile > 1:  # every tile must decode to the same residues as the first
        a = out[: descs.size * 2048]
        assert torch.equal(out[-descs.size * 2048:], a)
    del out, d_words, d_descs
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/rice_decode_roofline.json", "w"), indent=1)
