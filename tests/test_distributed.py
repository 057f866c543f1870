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
