"""Multi-GPU sharding of the frame batch: one process per GPU, torch.distributed plumbing.

Frames are independent, so the path shards with NO data-path collective: every
rank codes a contiguous block of frames.  The only exchange is the one the north star names --
a scatter of PCM blocks from the rank that holds the file and a variable-length gather of the
coded subframes back to it, so that concatenating the ranks' outputs in rank order reproduces
the single-device (and the reference's, src/sela/encoder.cpp:58-84) frame order bit for bit.
Over NCCL these are grouped point-to-point transfers on NVLink (NCCL has no gatherv); the same
code runs over gloo on CPU tensors, which is how tests/test_distributed.py covers it.

The local coder is passed in (`encode_fn(pcm, channels) -> (descs, words)`), so the host logic
here is independent of the device code.
"""
import numpy as np
import torch
import torch.distributed as dist

from ._lib import DESC_DTYPE, FRAME


def frame_block(n_frames, rank, world):
    """Contiguous block of frames for `rank`: n//world each, the last rank takes the rest
    (the reference's thread split, src/sela/encoder.cpp:58-73)."""
    per = n_frames // world
    lo = per * rank
    hi = n_frames if rank == world - 1 else per * (rank + 1)
    return lo, hi


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def _to_t(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(_dev())


def _p2p(ops):
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def scatter_frames(pcm, n_frames, channels, src=0):
    """Rank `src` holds the interleaved int16 PCM of n_frames frames; every rank gets its block."""
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = frame_block(n_frames, rank, world)
    stride = FRAME * channels * 2
    if rank == src:
        full = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1)
        ops = []
        for r in range(world):
            if r == src:
                continue
            a, b = frame_block(n_frames, r, world)
            if b > a:
                ops.append(dist.P2POp(dist.isend, _to_t(full[a * FRAME * channels:b * FRAME * channels]), r))
        _p2p(ops)
        return full[lo * FRAME * channels:hi * FRAME * channels].copy()
    buf = torch.empty((hi - lo) * stride, dtype=torch.uint8, device=_dev())
    if hi > lo:
        _p2p([dist.P2POp(dist.irecv, buf, src)])
    return buf.cpu().numpy().view(np.int16)


def gather_encoded(descs, words, dst=0):
    """Concatenate every rank's (descs, words) on `dst` in rank order, re-basing word offsets."""
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = torch.tensor([descs.size, words.size], dtype=torch.int64, device=_dev())
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [t.cpu().tolist() for t in all_sizes]
    if rank != dst:
        ops = []
        if descs.size:
            ops.append(dist.P2POp(dist.isend, _to_t(descs), dst))
        if words.size:
            ops.append(dist.P2POp(dist.isend, _to_t(words), dst))
        _p2p(ops)
        return None, None
    parts_d, parts_w, ops, bufs = [], [], [], {}
    for r in range(world):
        nd, nw = all_sizes[r]
        if r == dst:
            continue
        bufs[r] = (torch.empty(nd * DESC_DTYPE.itemsize, dtype=torch.uint8, device=_dev()),
                   torch.empty(nw * 4, dtype=torch.uint8, device=_dev()))
        if nd:
            ops.append(dist.P2POp(dist.irecv, bufs[r][0], r))
        if nw:
            ops.append(dist.P2POp(dist.irecv, bufs[r][1], r))
    _p2p(ops)
    base = 0
    for r in range(world):
        if r == dst:
            d, w = descs.copy(), np.ascontiguousarray(words, dtype=np.uint32)
        else:
            d = bufs[r][0].cpu().numpy().view(DESC_DTYPE).copy()
            w = bufs[r][1].cpu().numpy().view(np.uint32)
        d["refl_offset"] += base
        d["res_offset"] += base
        base += w.size
        parts_d.append(d)
        parts_w.append(w)
    return np.concatenate(parts_d), np.concatenate(parts_w)


def encode_sharded(pcm, n_frames, channels, encode_fn, root=0):
    """Scatter -> local encode on every rank -> gather.  Returns (descs, words) on `root`, (None, None) elsewhere."""
    block = scatter_frames(pcm, n_frames, channels, src=root)
    if block.size:
        descs, words = encode_fn(block, channels)
    else:
        descs, words = np.zeros(0, DESC_DTYPE), np.zeros(0, np.uint32)
    return gather_encoded(descs, words, dst=root)


def decode_sharded(descs, words, n_frames, channels, decode_fn, root=0):
    """Root holds (descs, words); each rank decodes its block of frames; PCM is gathered on root."""
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = frame_block(n_frames, rank, world)
    if rank == root:
        descs = np.ascontiguousarray(descs, dtype=DESC_DTYPE)
        words = np.ascontiguousarray(words, dtype=np.uint32)
        ops, mine = [], None
        for r in range(world):
            a, b = frame_block(n_frames, r, world)
            d = descs[a * channels:b * channels].copy()
            if d.size:
                w_lo = int(min(d["refl_offset"].min(), d["res_offset"].min()))
                w_hi = int(max((d["refl_offset"] + d["refl_words"]).max(), (d["res_offset"] + d["res_words"]).max()))
            else:
                w_lo = w_hi = 0
            d["refl_offset"] -= w_lo
            d["res_offset"] -= w_lo
            w = words[w_lo:w_hi]
            if r == root:
                mine = (d, w)
                continue
            hdr = torch.tensor([d.size, w.size], dtype=torch.int64, device=_dev())
            ops.append(dist.P2POp(dist.isend, hdr, r))
            if d.size:
                ops.append(dist.P2POp(dist.isend, _to_t(d), r))
            if w.size:
                ops.append(dist.P2POp(dist.isend, _to_t(w), r))
        _p2p(ops)
        d, w = mine
    else:
        hdr = torch.zeros(2, dtype=torch.int64, device=_dev())
        _p2p([dist.P2POp(dist.irecv, hdr, root)])
        nd, nw = hdr.cpu().tolist()
        bd = torch.empty(nd * DESC_DTYPE.itemsize, dtype=torch.uint8, device=_dev())
        bw = torch.empty(nw * 4, dtype=torch.uint8, device=_dev())
        ops = []
        if nd:
            ops.append(dist.P2POp(dist.irecv, bd, root))
        if nw:
            ops.append(dist.P2POp(dist.irecv, bw, root))
        _p2p(ops)
        d, w = bd.cpu().numpy().view(DESC_DTYPE), bw.cpu().numpy().view(np.uint32)
    local = decode_fn(d, w, channels) if d.size else np.zeros(0, np.int16)
    # gather PCM blocks (fixed size per rank, known from the frame split)
    if rank != root:
        if local.size:
            _p2p([dist.P2POp(dist.isend, _to_t(local), root)])
        return None
    out = np.zeros(n_frames * FRAME * channels, np.int16)
    ops, bufs = [], {}
    for r in range(world):
        a, b = frame_block(n_frames, r, world)
        if r == root:
            out[a * FRAME * channels:b * FRAME * channels] = local
        elif b > a:
            bufs[r] = torch.empty((b - a) * FRAME * channels * 2, dtype=torch.uint8, device=_dev())
            ops.append(dist.P2POp(dist.irecv, bufs[r], r))
    _p2p(ops)
    for r, t in bufs.items():
        a, b = frame_block(n_frames, r, world)
        out[a * FRAME * channels:b * FRAME * channels] = t.cpu().numpy().view(np.int16)
    return out
