"""CPU: the encoder oracle (oracle/clip_port.py) against outputs of the reference itself
(tests/golden/encoder_golden.npz, written by tests/golden/make_encoder_golden.py in the build container)."""
import os

import numpy as np
import pytest
import torch

from clip_based_cross_modal_hash_b200 import synth
from oracle import clip_port as port

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encoder_golden.npz"))
CASES = {"tiny": (synth.TINY, 3, 5), "vitb32": (synth.VIT_B32, 2, 4)}


def fingerprint(sd):
    rows = []
    for k in sorted(sd):
        t = sd[k].double().flatten()
        rows.append([float(t.sum()), float((t * t).sum())])
    return np.asarray(rows, dtype=np.float64)


def close(got, want, tol=2e-5):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    err = np.abs(got - want).max()
    assert err <= tol * max(1.0, np.abs(want).max()), err


@pytest.fixture(scope="module", params=list(CASES))
def case(request):
    cfg, nimg, ntxt = CASES[request.param]
    sd = synth.clip_state_dict(cfg, seed=11)
    # the seeded weights must be the ones the golden file was made with (same torch CPU generator stream)
    np.testing.assert_allclose(fingerprint(sd), Z[request.param + "/fingerprint"], rtol=1e-12, atol=1e-12)
    text, pad = synth.random_captions(ntxt, seed=22, vocab=cfg["vocab_size"])
    return request.param, sd, synth.random_images(nimg, seed=21), text, pad


def test_encode_image_matches_reference(case):
    tag, sd, image, _, _ = case
    with torch.no_grad():
        close(port.encode_image(sd, image), Z[tag + "/img_cls"])
        cls, seq, attn = port.encode_image(sd, image, return_patches=True)
    close(cls, Z[tag + "/img_cls_rp"])
    close(seq, Z[tag + "/img_seq"])
    close(attn, Z[tag + "/img_attn"], 1e-5)


def test_encode_text_matches_reference(case):
    tag, sd, _, text, pad = case
    with torch.no_grad():
        close(port.encode_text(sd, text), Z[tag + "/txt_eos"])
        close(port.encode_text(sd, text, key_padding_mask=pad), Z[tag + "/txt_eos_masked"])
        eos, seq, attn, newmask = port.encode_text(sd, text, key_padding_mask=pad, return_patches=True)
    close(eos, Z[tag + "/txt_eos_rp"])
    # rows of padded queries are compared too: the reference computes them (they only see unpadded keys)
    close(seq, Z[tag + "/txt_seq"])
    close(attn, Z[tag + "/txt_attn"], 1e-5)
    assert np.array_equal(newmask.numpy(), Z[tag + "/txt_newmask"])


@pytest.mark.parametrize("nbits", [16, 64])
def test_heads_match_reference(nbits):
    feat = torch.from_numpy(Z["head_feat"])
    hsd = synth.dsph_head_state_dict(512, nbits, seed=41)
    for m in ("img", "txt"):
        h = port.dsph_head(hsd, feat, m)
        close(h, Z["dsph%d/%s" % (nbits, m)], 1e-6)
        assert np.array_equal(port.make_hash_code_sign(h).numpy(), Z["dsph%d/%s_code" % (nbits, m)])
    hsd = synth.dcmht_head_state_dict(512, nbits, seed=42)
    for m in ("img", "txt"):
        h = port.dcmht_head(hsd, feat, m)
        close(h, Z["dcmht%d/%s" % (nbits, m)], 1e-6)
        assert np.array_equal(port.make_hash_code_dcmht(h).numpy(), Z["dcmht%d/%s_code" % (nbits, m)])
