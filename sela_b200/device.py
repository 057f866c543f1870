"""Device-resident entry points for callers that already hold the data in HBM.

PyTorch is used only for what it is good at here: device memory (tensors), streams
and events.  The work itself is the C ABI's *_device calls, enqueued on the
current torch stream without synchronising.
"""
import ctypes as C

import torch

from ._lib import FRAME, check, init, lib


class DeviceCodec:
    """Pre-allocated descriptor table, word arena and workspaces for a fixed batch shape."""

    def __init__(self, n_frames, channels, device=None, words_capacity=None):
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        init(self.device.index)
        L = lib()
        self.n_frames, self.channels = n_frames, channels
        self.n_sub = n_frames * channels
        self.capacity = words_capacity or L.selab200_encode_words_bound(n_frames, channels)
        dev = self.device
        self.descs = torch.zeros(self.n_sub * 32, dtype=torch.uint8, device=dev)
        self.words = torch.zeros(self.capacity, dtype=torch.int32, device=dev)
        self.words_used = torch.zeros(1, dtype=torch.int64, device=dev)
        self.status = torch.zeros(2, dtype=torch.int32, device=dev)
        self.enc_ws_bytes = L.selab200_encode_workspace_bytes(n_frames, channels)
        self.dec_ws_bytes = L.selab200_decode_workspace_bytes(n_frames, channels)
        self.enc_ws = torch.zeros(self.enc_ws_bytes, dtype=torch.uint8, device=dev)
        self.dec_ws = torch.zeros(self.dec_ws_bytes, dtype=torch.uint8, device=dev)

    def encode(self, pcm):
        """pcm: int16 cuda tensor, n_frames*2048*channels interleaved. Asynchronous."""
        assert pcm.dtype == torch.int16 and pcm.is_cuda and pcm.numel() == self.n_sub * FRAME
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(lib().selab200_encode_frames_device(
            pcm.data_ptr(), self.n_frames, self.channels, self.descs.data_ptr(), self.words.data_ptr(),
            self.capacity, self.words_used.data_ptr(), self.status.data_ptr(), self.enc_ws.data_ptr(),
            self.enc_ws_bytes, C.c_void_p(stream)))

    def decode(self, pcm_out, n_words):
        """Decode self.descs / self.words[:n_words] into pcm_out (int16 cuda tensor). Asynchronous."""
        assert pcm_out.dtype == torch.int16 and pcm_out.is_cuda and pcm_out.numel() == self.n_sub * FRAME
        stream = torch.cuda.current_stream(self.device).cuda_stream
        check(lib().selab200_decode_frames_device(
            self.descs.data_ptr(), self.n_frames, self.channels, self.words.data_ptr(), int(n_words),
            pcm_out.data_ptr(), self.status.data_ptr() + 4, self.dec_ws.data_ptr(), self.dec_ws_bytes,
            C.c_void_p(stream)))

    def check_status(self):
        """Synchronises; raises if either the last encode or decode reported an error."""
        from ._lib import SelaB200Error, STATUS_NAMES
        st = self.status.cpu().tolist()
        for what, s in zip(("encode", "decode"), st):
            if s != 0:
                raise SelaB200Error(s, "%s kernel reported %s" % (what, STATUS_NAMES.get(s, s)))


def rice_decode_frames(descs, words, channels, device=0):
    """Residue streams of every subframe -> int32 [n_sub, 2048] (the Rice-decode kernel on its own,
    selab200_rice_decode_frames_device).  descs: numpy array of DESC_DTYPE, words: uint32 arena.
    Returns (residues, n_flagged)."""
    import numpy as np
    from ._lib import DESC_DTYPE
    init(device)
    dev = torch.device("cuda", device)
    L = lib()
    n_sub = descs.size
    assert n_sub % channels == 0
    d_descs = torch.from_numpy(np.ascontiguousarray(descs).view(np.uint8).reshape(-1).copy()).to(dev)
    w = np.ascontiguousarray(words, dtype=np.uint32)
    d_words = torch.zeros(w.size + 8, dtype=torch.int32, device=dev)
    d_words[:w.size] = torch.from_numpy(w.view(np.int32)).to(dev)
    out = torch.empty(n_sub * FRAME, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    check(L.selab200_rice_decode_frames_device(d_descs.data_ptr(), n_sub // channels, channels, d_words.data_ptr(),
                                               w.size, out.data_ptr(), status.data_ptr(), C.c_void_p(stream)))
    torch.cuda.synchronize(dev)
    st = int(status.item())
    if st != 0:
        from ._lib import SelaB200Error, STATUS_NAMES
        raise SelaB200Error(st, "rice decode kernel reported %s" % STATUS_NAMES.get(st, st))
    flagged = C.c_uint32(0)
    check(L.selab200_rice_decode_flagged(C.addressof(flagged)))
    return out.cpu().numpy().reshape(n_sub, FRAME), flagged.value
