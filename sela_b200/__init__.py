"""sela_b200 -- B200-native implementation of SELA's per-frame encode/decode hot path.

CUDA kernels (sm_100a) behind the C ABI of include/sela_b200.h; this package holds
the kernels (csrc/), the C++ mirror of the reference interface (host/) and a thin
Python mirror used by the tests and bench.py.  No CPU fallback.
"""
from .codec import (DESC_DTYPE, FRAME, SelaB200Error, container_info, decode_container,  # noqa: F401
                    decode_frames, encode_container, encode_frames, init, lpc_residues, lpc_samples,
                    rice_decode, rice_encode)
