"""N>1 host logic on CPU: world_size-2 (and 3) gloo process groups run the scatter / local-code /
gather path of sela_b200.distributed with the CPU oracle standing in for the per-rank device coder;
the gathered result must equal coding the whole batch at once."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, channels, n_frames, q):
    import torch.distributed as dist
    import oracle_lib as ol
    from sela_b200 import distributed as sd, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O = ol.load("port")
    pcm = synth.sine_noise(44100, channels, n_frames=n_frames, seed=17) if rank == 0 else None
    enc = lambda block, ch: O.encode_frames(block, ch, threads=1)
    dec = lambda d, w, ch: O.decode_frames(d, w, ch, threads=1)
    descs, words = sd.encode_sharded(pcm, n_frames, channels, enc, root=0)
    out = sd.decode_sharded(descs, words, n_frames, channels, dec, root=0)
    if rank == 0:
        d_ref, w_ref = O.encode_frames(pcm, channels, threads=1)
        q.put((descs.tobytes() == d_ref.tobytes(), bool(np.array_equal(words, w_ref)),
               bool(np.array_equal(out, pcm.reshape(-1)))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,channels,n_frames", [(2, 2, 7), (3, 1, 4), (2, 8, 1)])
def test_sharded_encode_decode_equals_single(world, channels, n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, channels, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == (True, True, True)


def test_frame_block_mirrors_reference_thread_split():
    from sela_b200.distributed import frame_block
    for n, w in [(12919, 8), (7, 2), (3, 8), (0, 4), (84375, 8)]:
        blocks = [frame_block(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        assert all(b - a == n // w for a, b in blocks[:-1])


def _file_worker(rank, world, port, channels, n_frames, tmp, q):
    import torch.distributed as dist
    import oracle_lib as ol
    from sela_b200 import codec, distributed as sd, synth, wavio
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O = ol.load("port")
    wav, sela, back = os.path.join(tmp, "in.wav"), os.path.join(tmp, "out.sela"), os.path.join(tmp, "back.wav")
    pcm = synth.sine_noise(48000, channels, n_frames=n_frames, seed=23)
    if rank == 0:   # a tail that is not a whole frame, and an extra chunk in front of the data
        wavio.write_wav(wav, np.concatenate([pcm, pcm[:300]]), 48000, extra_chunks=[(b"LIST", b"INFOxyz0")])
    dist.barrier()

    def container_fn(block, ch, rate):          # the CPU oracle stands in for selab200_encode_container
        d, w = O.encode_frames(block, ch, threads=1)
        return wavio.pack_container(d, w, rate, ch)

    def decode_fn(blob):                        # ... and for selab200_container_open/_decode
        info, offsets = codec.container_frame_offsets(blob)
        assert info["n_frames"] == info["header_frames"]
        d, w = _unpack(blob, info)
        return O.decode_frames(d, w, info["channels"], threads=1)

    n, total = sd.encode_file_sharded(wav, sela, container_fn)
    m = sd.decode_file_sharded(sela, back, codec.container_frame_offsets, decode_fn)
    if rank == 0:
        d_ref, w_ref = O.encode_frames(pcm, channels, threads=1)
        whole = wavio.pack_container(d_ref, w_ref, 48000, channels)
        wavio.write_wav(os.path.join(tmp, "canon.wav"), pcm, 48000)
        q.put((n, m, total == len(whole), open(sela, "rb").read() == whole,
               open(back, "rb").read() == open(os.path.join(tmp, "canon.wav"), "rb").read()))
    dist.barrier()
    dist.destroy_process_group()


def _unpack(blob, info):
    """.sela bytes -> (descs, words), plain byte shuffling (inverse of wavio.pack_container)."""
    import struct
    from sela_b200._lib import DESC_DTYPE
    ch, at, words, descs = info["channels"], 15, [], []
    n_words = 0
    for _ in range(info["n_frames"]):
        at += 4
        for _ in range(ch):
            c, t, p, rk, rn, order = struct.unpack_from("<BBBBHB", blob, at)
            at += 7
            words.append(np.frombuffer(blob, "<u4", rn, at))
            at += 4 * rn
            sk, sn, samples = struct.unpack_from("<BHH", blob, at)
            at += 5
            words.append(np.frombuffer(blob, "<u4", sn, at))
            at += 4 * sn
            descs.append((c, t, p, rk, rn, order, sk, sn, samples, 0, n_words, n_words + rn))
            n_words += rn + sn
    return np.array(descs, DESC_DTYPE), np.concatenate(words) if words else np.zeros(0, np.uint32)


@pytest.mark.parametrize("world,channels,n_frames", [(2, 8, 5), (3, 2, 7), (2, 1, 1)])
def test_sharded_files_equal_single_process(world, channels, n_frames, tmp_path):
    """encode_file_sharded / decode_file_sharded: each rank reads, codes and writes its own block of the
    files; the only traffic is one all_gather of sizes / one broadcast of offsets."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_file_worker, args=(r, world, port, channels, n_frames, str(tmp_path), q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == (n_frames, n_frames, True, True, True)
