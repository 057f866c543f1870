"""Golden vectors (tests/golden/golden_frames.npz, generated from the unmodified reference by
tests/golden/make_golden.py): the plain-C oracle on the CPU, the CUDA path on the GPU."""
import pathlib

import numpy as np
import pytest

import oracle_lib as ol

GOLD = np.load(pathlib.Path(__file__).parent / "golden" / "golden_frames.npz")
CASES = sorted(k[4:] for k in GOLD.files if k.startswith("pcm_"))


@pytest.mark.parametrize("case", CASES)
def test_oracle_port_reproduces_golden(case):
    P = ol.load("port")
    pcm = GOLD["pcm_" + case]
    ch = pcm.shape[1]
    descs, words = P.encode_frames(pcm, ch, threads=2)
    assert descs.tobytes() == GOLD["descs_" + case].tobytes()
    assert np.array_equal(words, GOLD["words_" + case])
    assert np.array_equal(P.decode_frames(descs, words, ch), GOLD["decoded_" + case])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_cuda_path_reproduces_golden(case):
    import sela_b200
    pcm = GOLD["pcm_" + case]
    ch = pcm.shape[1]
    descs, words = sela_b200.encode_frames(pcm, ch)
    assert descs.tobytes() == GOLD["descs_" + case].tobytes()
    assert np.array_equal(words, GOLD["words_" + case])
    out = sela_b200.decode_frames(GOLD["descs_" + case].view(sela_b200.DESC_DTYPE), GOLD["words_" + case], ch)
    assert np.array_equal(out, GOLD["decoded_" + case])
