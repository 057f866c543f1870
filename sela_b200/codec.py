"""Host-side Python mirror of the reference's operator interface for the hot path.

Function names follow the reference classes they stand for (argument meaning and
error behaviour as in include/sela_b200.h):

    encode_frames / decode_frames   sela::Encoder/Decoder::processFrames
                                    (= frame::FrameEncoder/FrameDecoder::process per frame)
    lpc_residues / lpc_samples      lpc::ResidueGenerator / lpc::SampleGenerator ::process
    rice_encode / rice_decode       rice::RiceEncoder / rice::RiceDecoder ::process
    encode_container / decode_container / container_info
                                    sela::Encoder::process + file::SelaFile::writeToFile, and
                                    file::SelaFile::readFromFile + sela::Decoder::processFrames,
                                    on the byte-packed .sela stream

Everything computes on the GPU through the C ABI; NumPy only carries host buffers.
The C++ mirror of the same interface (data::, frame::, file::, sela:: classes and
the `sela` CLI) is in sela_b200/host/.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import DESC_DTYPE, FRAME, INFO_DTYPE, MAX_ORDER, SelaB200Error, check, init, lib  # noqa: F401


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def encode_frames(pcm, channels, words_capacity=None, device=0):
    """pcm: int16, n_frames*2048*channels interleaved samples -> (descs, words)."""
    init(device)
    pcm = _c(pcm, np.int16).reshape(-1)
    n_frames = pcm.size // (FRAME * channels)
    if n_frames * FRAME * channels != pcm.size:
        raise ValueError("pcm must hold whole 2048-sample frames")
    L = lib()
    cap = words_capacity if words_capacity is not None else L.selab200_encode_words_bound(n_frames, channels)
    descs = np.zeros(n_frames * channels, DESC_DTYPE)
    words = np.empty(max(cap, 1), np.uint32)
    used = C.c_size_t(0)
    check(L.selab200_encode_frames(pcm.ctypes.data, n_frames, channels, descs.ctypes.data, words.ctypes.data,
                                   cap, C.addressof(used)))
    return descs, words[:used.value].copy()


def decode_frames(descs, words, channels, device=0):
    """(descs, words) -> int16 interleaved PCM, n_frames*2048*channels samples."""
    init(device)
    descs = _c(descs, DESC_DTYPE)
    words = _c(words, np.uint32)
    n_frames = descs.size // channels
    if n_frames * channels != descs.size:
        raise ValueError("descs must hold `channels` subframes per frame")
    pcm = np.empty(n_frames * FRAME * channels, np.int16)
    check(lib().selab200_decode_frames(descs.ctypes.data, n_frames, channels, words.ctypes.data, words.size,
                                       pcm.ctypes.data))
    return pcm


def lpc_residues(samples, device=0):
    """int32 [n_sub, 2048] -> (order u8[n_sub], q int32[n_sub,100], residues int32[n_sub,2048])."""
    init(device)
    samples = _c(samples, np.int32).reshape(-1, FRAME)
    n = samples.shape[0]
    order = np.zeros(n, np.uint8)
    q = np.zeros((n, MAX_ORDER), np.int32)
    res = np.zeros((n, FRAME), np.int32)
    check(lib().selab200_lpc_residues(samples.ctypes.data, n, order.ctypes.data, q.ctypes.data, res.ctypes.data))
    return order, q, res


def lpc_samples(residues, order, q, device=0):
    init(device)
    residues = _c(residues, np.int32).reshape(-1, FRAME)
    n = residues.shape[0]
    order = _c(order, np.uint8).reshape(n)
    qq = np.zeros((n, MAX_ORDER), np.int32)
    q = np.asarray(q)
    qq[:, :q.shape[1]] = q
    out = np.zeros((n, FRAME), np.int32)
    check(lib().selab200_lpc_samples(residues.ctypes.data, n, order.ctypes.data, qq.ctypes.data, out.ctypes.data))
    return out


def rice_encode(values, counts=None, words_stride=None, device=0):
    """values int32 [n_streams, stride] (stride <= 2048) -> (k u32[n], n_words u32[n], words u32[n, words_stride])."""
    init(device)
    values = _c(values, np.int32)
    if values.ndim == 1:
        values = values.reshape(1, -1)
    n, stride = values.shape
    counts = np.full(n, stride, np.uint32) if counts is None else _c(counts, np.uint32)
    if words_stride is None:
        words_stride = stride * 2 + 8
    k = np.zeros(n, np.uint32)
    nw = np.zeros(n, np.uint32)
    words = np.zeros((n, words_stride), np.uint32)
    check(lib().selab200_rice_encode(values.ctypes.data, counts.ctypes.data, n, stride, k.ctypes.data,
                                     nw.ctypes.data, words.ctypes.data, words_stride))
    return k, nw, words


def rice_decode(words, n_words, k, counts, out_stride=None, device=0):
    init(device)
    words = _c(words, np.uint32)
    if words.ndim == 1:
        words = words.reshape(1, -1)
    n, words_stride = words.shape
    n_words = _c(n_words, np.uint32).reshape(n)
    k = _c(k, np.uint32).reshape(n)
    counts = _c(counts, np.uint32).reshape(n)
    if out_stride is None:
        out_stride = int(counts.max()) if n else 0
    out = np.zeros((n, max(out_stride, 1)), np.int32)
    check(lib().selab200_rice_decode(words.ctypes.data, n_words.ctypes.data, words_stride, k.ctypes.data,
                                     counts.ctypes.data, n, out.ctypes.data, out_stride))
    return out


def encode_container(pcm, channels, sample_rate, bits_per_sample=16, capacity=None, device=0):
    """int16 interleaved PCM (whole frames) -> the bytes of the .sela file (uint8 array)."""
    init(device)
    pcm = _c(pcm, np.int16).reshape(-1)
    n_frames = pcm.size // (FRAME * channels)
    if n_frames * FRAME * channels != pcm.size:
        raise ValueError("pcm must hold whole 2048-sample frames")
    L = lib()
    cap = capacity if capacity is not None else L.selab200_container_bound(n_frames, channels)
    out = np.empty(max(cap, 1), np.uint8)
    used = C.c_size_t(0)
    check(L.selab200_encode_container(pcm.ctypes.data, n_frames, channels, sample_rate, bits_per_sample,
                                      out.ctypes.data, cap, C.addressof(used)))
    return out[:used.value]


def container_info(container):
    """Header fields and frame walk of a .sela byte stream (host only, no device needed)."""
    buf = _c(np.frombuffer(container, np.uint8) if isinstance(container, (bytes, bytearray)) else container, np.uint8)
    info = np.zeros(1, INFO_DTYPE)
    check(lib().selab200_container_info_get(buf.ctypes.data, buf.size, info.ctypes.data))
    return {k: int(info[0][k]) for k in INFO_DTYPE.names if k != "reserved"}


def container_frame_offsets(container):
    """(info dict, uint64 array of n_frames+1 byte offsets: frame starts and the end) -- host only."""
    buf = _c(np.frombuffer(container, np.uint8) if isinstance(container, (bytes, bytearray)) else container, np.uint8)
    info = np.zeros(1, INFO_DTYPE)
    L = lib()
    check(L.selab200_container_info_get(buf.ctypes.data, buf.size, info.ctypes.data))
    offsets = np.zeros(int(info[0]["n_frames"]) + 1, np.uint64)
    check(L.selab200_container_frame_offsets(buf.ctypes.data, buf.size, offsets.ctypes.data, offsets.size,
                                             info.ctypes.data))
    return {k: int(info[0][k]) for k in INFO_DTYPE.names if k != "reserved"}, offsets


def decode_container(container, device=0):
    """.sela byte stream -> (info dict, int16 interleaved PCM of info['n_frames'] frames)."""
    init(device)
    buf = _c(np.frombuffer(container, np.uint8) if isinstance(container, (bytes, bytearray)) else container, np.uint8)
    L = lib()
    info = np.zeros(1, INFO_DTYPE)
    handle = C.c_void_p(0)
    check(L.selab200_container_open(buf.ctypes.data, buf.size, C.addressof(handle), info.ctypes.data))
    try:
        n = int(info[0]["n_frames"]) * int(info[0]["channels"]) * FRAME
        pcm = np.empty(n, np.int16)
        check(L.selab200_container_decode(handle, pcm.ctypes.data))
    finally:
        L.selab200_container_close(handle)
    return {k: int(info[0][k]) for k in INFO_DTYPE.names if k != "reserved"}, pcm
