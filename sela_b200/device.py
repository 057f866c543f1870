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
