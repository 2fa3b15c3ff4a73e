"""CPU: the MITH head oracle (oracle/mith_port.py) against outputs of the reference HashLayer (tests/golden/mith_golden.npz)."""
import os

import numpy as np
import pytest
import torch

from clip_based_cross_modal_hash_b200 import synth
from oracle import mith_port as port

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mith_golden.npz"))


def inputs(B, L, seed, padded):   # same generator calls as tests/golden/make_mith_golden.py
    g = torch.Generator().manual_seed(seed)
    cls = torch.randn((B, 512), generator=g)
    tokens = torch.randn((L, B, 512), generator=g)
    mask = None
    if padded:
        lens = torch.randint(2, L + 1, (B,), generator=g)
        mask = torch.arange(L)[None, :] >= lens[:, None]
    return cls, tokens, mask


@pytest.mark.parametrize("nbits", [16, 64])
@pytest.mark.parametrize("modality", ["img", "txt"])
def test_mith_head_matches_reference(nbits, modality):
    hsd = synth.mith_head_state_dict(512, nbits, seed=51)
    cls, tokens, mask = inputs(5, 49, 61, False) if modality == "img" else inputs(6, 32, 62, True)
    with torch.no_grad():
        got = port.encode(hsd, modality, cls, tokens, mask)
    for name, v in zip(("res", "cls_hash", "tok_hash", "trans"), got):
        want = Z["mith%d/%s_%s" % (nbits, modality, name)]
        assert v.shape == want.shape
        assert np.abs(v.numpy() - want).max() <= 2e-5, (name, np.abs(v.numpy() - want).max())
    code = port.generate_hash(got[1], got[2]).numpy()
    want = Z["mith%d/%s_code" % (nbits, modality)]
    margin = np.abs((got[1] + got[2]).numpy())
    assert np.array_equal(code[margin > 1e-4], want[margin > 1e-4])
