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


def _pwrite_all(path, data, offset):
    """pwrite until everything is out (a single call may stop short on large buffers)."""
    import os
    view = memoryview(data)
    fd = os.open(path, os.O_WRONLY)
    try:
        done = 0
        while done < len(view):
            done += os.pwrite(fd, view[done:], offset + done)
    finally:
        os.close(fd)


# ---------------------------------------------------------------- whole files --
#
# When every rank can open the files (one box, BASELINE config 4), nothing has to travel between
# GPUs at all: WAV frames sit at fixed byte positions, so each rank reads its own block of the input;
# the only thing ranks must agree on is WHERE in the output each block lands, i.e. one all_gather of
# body lengths (encode) or one broadcast of frame byte offsets from the rank that walked the headers
# (decode).  Every subframe of a .sela stream starts at a byte position = 3 (mod 4) whatever precedes
# it (15 + 4(f+1) + 12g + 4W), so a block coded as a stand-alone container is, minus its 15-byte
# header, exactly the byte range it occupies in the whole file.

_SELA_HEADER = 15


def _wav_layout(path):
    """(sample_rate, channels, bits, data_offset, data_bytes) of a RIFF/WAVE file (any chunk order)."""
    import struct
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            raise ValueError("%s: not a RIFF/WAVE file" % path)
        fmt = None
        while True:
            hdr = f.read(8)
            if len(hdr) < 8:
                raise ValueError("%s: no data chunk" % path)
            cid, size = hdr[:4], struct.unpack("<I", hdr[4:])[0]
            if cid == b"fmt ":
                body = f.read(size)
                _, channels, rate, _, _, bits = struct.unpack("<HHIIHH", body[:16])
                fmt = (rate, channels, bits)
            elif cid == b"data":
                if fmt is None:
                    raise ValueError("%s: data chunk before fmt chunk" % path)
                return fmt + (f.tell(), size)
            else:
                f.seek(size, 1)


def encode_file_sharded(wav_path, sela_path, container_fn):
    """Every rank encodes its block of `wav_path`'s frames with `container_fn(pcm, channels, rate) ->
    bytes of a stand-alone .sela container` and writes it in place into `sela_path`.  The result is
    byte-identical to a single-process encode.  Returns (n_frames, total_bytes)."""
    import struct
    rank, world = dist.get_rank(), dist.get_world_size()
    rate, channels, bits, data_off, data_bytes = _wav_layout(wav_path)
    if bits != 16:
        raise ValueError("Only 16bits per sample wav is supported.")
    stride = FRAME * channels * 2
    n_frames = data_bytes // stride                       # whole frames only (src/file/wav_file.cpp:184)
    lo, hi = frame_block(n_frames, rank, world)
    pcm = np.fromfile(wav_path, dtype="<i2", offset=data_off + lo * stride, count=(hi - lo) * FRAME * channels)
    body = memoryview(np.ascontiguousarray(np.frombuffer(container_fn(pcm, channels, rate), np.uint8)))[_SELA_HEADER:] \
        if hi > lo else b""
    mine = torch.tensor([len(body)], dtype=torch.int64, device=_dev())
    sizes = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(sizes, mine)
    sizes = [int(t.item()) for t in sizes]
    total = _SELA_HEADER + sum(sizes)
    if rank == 0:
        with open(sela_path, "wb") as f:
            f.write(b"SeLa" + struct.pack("<IHBI", rate, bits, channels, n_frames))
            f.truncate(total)
    dist.barrier()
    if body:
        _pwrite_all(sela_path, body, _SELA_HEADER + sum(sizes[:rank]))
    dist.barrier()
    return n_frames, total


def decode_file_sharded(sela_path, wav_path, frame_offsets_fn, decode_fn):
    """Rank 0 walks the frame headers (`frame_offsets_fn(bytes) -> (info, offsets)`, host only) and
    broadcasts the block boundaries; every rank reads its byte range, decodes it as a stand-alone
    container (`decode_fn(bytes) -> int16 PCM`) and writes its slice of the WAV data chunk.
    Returns n_frames."""
    import struct
    rank, world = dist.get_rank(), dist.get_world_size()
    meta = torch.zeros(4 + world + 1, dtype=torch.int64, device=_dev())
    if rank == 0:
        blob = np.memmap(sela_path, dtype=np.uint8, mode="r")
        info, offsets = frame_offsets_fn(blob)
        n_frames = info["n_frames"]
        cuts = [int(offsets[frame_block(n_frames, r, world)[0]]) for r in range(world)] + [int(offsets[n_frames])]
        meta[:] = torch.tensor([n_frames, info["channels"], info["sample_rate"], info["bits_per_sample"]] + cuts)
        del blob
    dist.broadcast(meta, 0)
    meta = meta.cpu().tolist()
    n_frames, channels, rate, bits = meta[:4]
    cuts = meta[4:]
    lo, hi = frame_block(n_frames, rank, world)
    stride = FRAME * channels * 2
    payload = n_frames * channels * FRAME * (bits // 8)
    if rank == 0:
        with open(wav_path, "wb") as f:  # header of file::WavFile::WavFile (src/file/wav_file.cpp:7-37)
            f.write(b"RIFF" + struct.pack("<I", (payload + 36) & 0xffffffff) + b"WAVEfmt " +
                    struct.pack("<IHHIIHH", 16, 1, channels, rate, (rate * channels * bits) // 8 & 0xffffffff,
                                (channels * bits) // 8, bits) +
                    b"data" + struct.pack("<I", payload & 0xffffffff))
            f.truncate(44 + n_frames * stride)
    dist.barrier()
    if hi > lo:
        with open(sela_path, "rb") as f:
            f.seek(cuts[rank])
            body = f.read(cuts[rank + 1] - cuts[rank])
        block = b"SeLa" + struct.pack("<IHBI", rate, bits, channels, hi - lo) + body
        pcm = np.ascontiguousarray(decode_fn(block), dtype="<i2")
        if pcm.size != (hi - lo) * FRAME * channels:
            raise ValueError("block decoded to %d samples, expected %d" % (pcm.size, (hi - lo) * FRAME * channels))
        _pwrite_all(wav_path, pcm.view(np.uint8).reshape(-1), 44 + lo * stride)
    dist.barrier()
    return n_frames


# ------------------------------------------------------- device-resident sharding --
#
# The same scatter / code / gather with every buffer in HBM (NCCL moves device memory; the only
# host<->device traffic left is the root's own upload of the file and download of the result).
# `DeviceCodec` (sela_b200/device.py) does the local coding through the *_device C-ABI calls.

def _timed(fn, dev):
    """Run fn() between two CUDA events on the current stream; returns (result, milliseconds)."""
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    torch.cuda.synchronize(dev)
    return out, a.elapsed_time(b)


def scatter_frames_device(pcm_dev, n_frames, channels, src=0):
    """pcm_dev: on `src` an int16 CUDA tensor with the whole interleaved file; returns this rank's block (a
    CUDA int16 tensor; on `src` a view of pcm_dev)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = frame_block(n_frames, rank, world)
    per = FRAME * channels
    if rank == src:
        ops = []
        for r in range(world):
            a, b = frame_block(n_frames, r, world)
            if r != src and b > a:
                ops.append(dist.P2POp(dist.isend, pcm_dev[a * per:b * per].view(torch.uint8), r))  # (NCCL has no int16)
        _p2p(ops)
        return pcm_dev[lo * per:hi * per]
    buf = torch.empty((hi - lo) * per, dtype=torch.int16, device=_dev())
    if hi > lo:
        _p2p([dist.P2POp(dist.irecv, buf.view(torch.uint8), src)])
    return buf


def gather_encoded_device(descs_dev, words_dev, n_words, n_frames, channels, dst=0):
    """descs_dev: uint8 CUDA tensor (32 bytes per subframe of this rank's block), words_dev: int32 CUDA tensor,
    n_words: words used.  On `dst` returns (descs, words) for the whole file, offsets re-based, both on the
    device; elsewhere (None, None)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = torch.tensor([n_words], dtype=torch.int64, device=_dev())
    sizes = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(sizes, mine)
    sizes = [int(t.item()) for t in sizes]
    if rank != dst:
        ops = []
        if descs_dev.numel():
            ops.append(dist.P2POp(dist.isend, descs_dev, dst))
        if n_words:
            ops.append(dist.P2POp(dist.isend, words_dev[:n_words], dst))
        _p2p(ops)
        return None, None
    all_descs = torch.empty(n_frames * channels * 32, dtype=torch.uint8, device=_dev())
    all_words = torch.empty(sum(sizes) + 8, dtype=torch.int32, device=_dev())
    ops, base = [], 0
    bases = []
    for r in range(world):
        a, b = frame_block(n_frames, r, world)
        bases.append(base)
        d_view = all_descs[a * channels * 32:b * channels * 32]
        w_view = all_words[base:base + sizes[r]]
        if r == dst:
            d_view.copy_(descs_dev)
            w_view.copy_(words_dev[:sizes[r]])
        else:
            if b > a:
                ops.append(dist.P2POp(dist.irecv, d_view, r))
            if sizes[r]:
                ops.append(dist.P2POp(dist.irecv, w_view, r))
        base += sizes[r]
    _p2p(ops)
    # re-base the two word offsets of every descriptor (int64 fields at bytes 16 and 24 of the 32-byte record)
    table = all_descs.view(torch.int64).view(-1, 4)
    for r in range(world):
        a, b = frame_block(n_frames, r, world)
        if b > a and bases[r]:
            table[a * channels:b * channels, 2:4] += bases[r]
    return all_descs, all_words[:base]


def encode_sharded_device(pcm_dev, n_frames, channels, root=0):
    """Scatter (NCCL) -> device-resident encode on every rank -> gather (NCCL).  Returns
    ((descs, words) on root / (None, None) elsewhere, {'scatter_ms', 'encode_ms', 'gather_ms'})."""
    from .device import DeviceCodec
    dev = _dev()
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = frame_block(n_frames, rank, world)
    block, t_scatter = _timed(lambda: scatter_frames_device(pcm_dev, n_frames, channels, src=root), dev)
    codec = DeviceCodec(max(hi - lo, 1), channels, device=dev.index)

    def run():
        if hi > lo:
            codec.encode(block.contiguous())
    _, t_enc = _timed(run, dev)
    codec.check_status()
    n_words = int(codec.words_used.item()) if hi > lo else 0
    descs = codec.descs if hi > lo else codec.descs[:0]
    out, t_gather = _timed(lambda: gather_encoded_device(descs, codec.words, n_words, n_frames, channels, dst=root), dev)
    return out, {"scatter_ms": t_scatter, "encode_ms": t_enc, "gather_ms": t_gather}


def decode_sharded_device(descs_dev, words_dev, n_frames, channels, root=0):
    """Root holds (descs, words) on the device; each rank decodes its block of frames on its GPU; the PCM is
    gathered on root (int16 CUDA tensor; None elsewhere).  Returns (pcm, timings)."""
    from .device import DeviceCodec
    dev = _dev()
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = frame_block(n_frames, rank, world)
    per = FRAME * channels

    def scatter():
        if rank == root:
            table = descs_dev.view(torch.int64).view(-1, 4)
            ops, mine = [], None
            for r in range(world):
                a, b = frame_block(n_frames, r, world)
                if b <= a:
                    continue
                d = descs_dev[a * channels * 32:b * channels * 32].clone()
                t = d.view(torch.int64).view(-1, 4)
                w_lo = int(torch.minimum(t[:, 2], t[:, 3]).min().item())
                rec = descs_dev[(b * channels - 1) * 32:(b * channels) * 32].cpu().numpy().view(DESC_DTYPE)[0]
                w_hi = max(int(rec["refl_offset"]) + int(rec["refl_words"]), int(rec["res_offset"]) + int(rec["res_words"]))
                t[:, 2:4] -= w_lo
                w = words_dev[w_lo:w_hi]
                if r == root:
                    mine = (d, w.clone())
                    continue
                hdr = torch.tensor([w_hi - w_lo], dtype=torch.int64, device=dev)
                ops += [dist.P2POp(dist.isend, hdr, r), dist.P2POp(dist.isend, d, r), dist.P2POp(dist.isend, w, r)]
            _p2p(ops)
            return mine if mine is not None else (descs_dev[:0], words_dev[:0])
        if hi <= lo:
            return torch.empty(0, dtype=torch.uint8, device=dev), torch.empty(0, dtype=torch.int32, device=dev)
        hdr = torch.zeros(1, dtype=torch.int64, device=dev)
        d = torch.empty((hi - lo) * channels * 32, dtype=torch.uint8, device=dev)
        _p2p([dist.P2POp(dist.irecv, hdr, root), dist.P2POp(dist.irecv, d, root)])
        w = torch.empty(int(hdr.item()) + 8, dtype=torch.int32, device=dev)
        _p2p([dist.P2POp(dist.irecv, w[:int(hdr.item())], root)])
        return d, w[:int(hdr.item())]

    (d, w), t_scatter = _timed(scatter, dev)
    codec = DeviceCodec(max(hi - lo, 1), channels, device=dev.index, words_capacity=max(int(w.numel()), 1) + 8)
    local = torch.empty((hi - lo) * per, dtype=torch.int16, device=dev)

    def run():
        if hi > lo:
            codec.descs.copy_(d)
            codec.words[:w.numel()].copy_(w)
            codec.decode(local, int(w.numel()))
    _, t_dec = _timed(run, dev)
    codec.check_status()

    def gather():
        if rank != root:
            if hi > lo:
                _p2p([dist.P2POp(dist.isend, local.view(torch.uint8), root)])
            return None
        out = torch.empty(n_frames * per, dtype=torch.int16, device=dev)
        ops = []
        for r in range(world):
            a, b = frame_block(n_frames, r, world)
            if r == root:
                out[a * per:b * per].copy_(local)
            elif b > a:
                ops.append(dist.P2POp(dist.irecv, out[a * per:b * per].view(torch.uint8), r))
        _p2p(ops)
        return out
    out, t_gather = _timed(gather, dev)
    return out, {"scatter_ms": t_scatter, "decode_ms": t_dec, "gather_ms": t_gather}
