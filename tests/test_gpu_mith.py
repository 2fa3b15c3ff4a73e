"""GPU: the MITH hash head (through the C ABI) against outputs of the reference HashLayer (tests/golden/mith_golden.npz) and the
fp32 oracle (oracle/mith_port.py).

Tolerance: the residual MLPs, the concept transformer and the concept projection run as bf16 tcgen05 GEMMs with fp32
accumulation (same class as the CLIP towers).  cls_hash is a smooth function of its input: max abs error
<= 6e-2.  tokens_hash goes through the top-k concept selection and the softmax over the selected tokens, which are
DISCONTINUOUS: a similarity within bf16 noise of a token's k-th value switches that token on or off for a concept and moves
the merged token by O(1/#selected tokens).  With K concepts spread over (0, 1) the gap between a token's 8th and 9th
similarity is ~1/K, so at bf16 noise (~1e-2 relative) a few per cent of the (token, concept) decisions differ from an fp32
evaluation, on random-normal token inputs more than on trained features.  Therefore the head runs the token-path MLPs in
SPLIT precision by default (operands and weights as bf16 hi + lo parts, K-concatenated GEMM: ~2^-16 relative): against the
reference golden tokens_hash then agrees to max 3e-2 / mean 5e-3 with no selection flip.  With split_precision=False (plain
bf16, 3x fewer FLOPs on those GEMMs) and for the end-to-end model (where the bf16 towers already perturb the token features)
the stated tolerance is: mean abs error <= 2e-2, at most 5 % of the entries off by more than 6e-2 ("selection flips"; 2.8 %
observed at K = 64, 0 at K = 16).  Code bits: at most 3 % differ from the reference, and every differing bit either has a
reference margin |cls_hash + tokens_hash| < 0.1 or sits on a selection flip."""
import os

import numpy as np
import pytest
import torch

from clip_based_cross_modal_hash_b200 import models, synth
from oracle import clip_port, mith_port

pytestmark = pytest.mark.gpu
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


def unpack(packed, nbits):
    w = packed.cpu().numpy().view(np.uint32)
    return ((w[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(w.shape[0], -1)[:, :nbits].astype(np.float32) * 2 - 1


def check_hash(got, want, what):
    err = np.abs(got - want)
    if what == "cls_hash":
        assert err.max() <= 6e-2 and err.mean() <= 1e-2, (what, err.max(), err.mean())
    elif what == "tokens_hash_split":
        assert err.max() <= 3e-2 and err.mean() <= 5e-3, (what, err.max(), err.mean())
    else:
        assert err.mean() <= 2e-2 and (err > 6e-2).mean() <= 0.05, (what, err.mean(), (err > 6e-2).mean())
    return err > 6e-2


def check_code(code, ref_sum, flips=None):
    want = np.sign(ref_sum)
    diff = code != want
    assert diff.mean() <= 0.03, diff.mean()
    unexplained = diff & (np.abs(ref_sum) >= 0.1)
    if flips is not None:
        unexplained &= ~flips
    assert not unexplained.any(), np.abs(ref_sum[unexplained])


@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("nbits", [16, 64])
def test_mith_head_matches_reference_golden(nbits, split):
    head = models.MithHashLayer(synth.mith_head_state_dict(512, nbits, seed=51), "cuda", split_precision=split)
    for m in ("img", "txt"):
        cls, tokens, mask = inputs(5, 49, 61, False) if m == "img" else inputs(6, 32, 62, True)
        r = head.encode_img(cls.cuda(), tokens.cuda()) if m == "img" else head.encode_txt(cls.cuda(), tokens.cuda(), mask.cuda())
        res, ch, th, trans = (t.cpu().numpy() for t in r)
        p = "mith%d/%s_" % (nbits, m)
        assert res.shape == Z[p + "res"].shape and trans.shape == Z[p + "trans"].shape
        cos = (res * Z[p + "res"]).sum(-1)
        assert cos.min() >= 0.9995, cos.min()
        check_hash(ch, Z[p + "cls_hash"], "cls_hash")
        flips = check_hash(th, Z[p + "tok_hash"], "tokens_hash_split" if split else "tokens_hash")
        tcos = (trans * Z[p + "trans"]).sum(-1)          # both sides are unit vectors, [K, B]
        assert np.median(tcos) >= 0.999 and (tcos < 0.98).mean() <= 0.05, (np.median(tcos), (tcos < 0.98).mean())
        check_code(np.sign(ch + th), Z[p + "cls_hash"] + Z[p + "tok_hash"], flips)


def test_mith_model_end_to_end_vs_oracle():
    """images/captions -> ViT-B/32 towers (return_patches) -> MITH head -> packed codes, against the fp32 oracle chain."""
    nbits, B = 32, 4
    sd = synth.clip_state_dict(synth.VIT_B32, seed=11)
    hsd = synth.mith_head_state_dict(512, nbits, seed=7)
    model = models.MITH(sd, hsd)
    image = synth.random_images(B, seed=81)
    text, pad = synth.random_captions(B, seed=82)
    with torch.no_grad():
        cls, seq, _ = clip_port.encode_image(sd, image, return_patches=True)
        wi = mith_port.encode(hsd, "img", cls, seq, None)
        eos, tseq, _, newmask = clip_port.encode_text(sd, text, pad, return_patches=True)
        wt = mith_port.encode(hsd, "txt", eos, tseq, newmask)
    gi, gt = model.encode_image(image), model.encode_text(text, pad)
    flips = []
    for got, want in ((gi, wi), (gt, wt)):
        assert tuple(got[3].shape) == tuple(want[3].shape)          # trans_tokens [K, B, D]
        check_hash(got[1].cpu().numpy(), want[1].numpy(), "cls_hash")
        flips.append(check_hash(got[2].cpu().numpy(), want[2].numpy(), "tokens_hash"))
    ih, th = model.generate_hash(image, text, pad)
    assert torch.allclose(ih, gi[1] + gi[2]) and torch.allclose(th, gt[1] + gt[2])
    code_i, code_t = unpack(model.encode_image_packed(image), nbits), unpack(model.encode_text_packed(text, pad), nbits)
    assert np.array_equal(code_i, np.sign(ih.cpu().numpy())) and np.array_equal(code_t, np.sign(th.cpu().numpy()))
    check_code(code_i, (wi[1] + wi[2]).numpy(), flips[0])
    check_code(code_t, (wt[1] + wt[2]).numpy(), flips[1])
    # get_code drives the same path (key_padding_mask forwarded for MITH)
    loader = [(image, text, pad, None, torch.arange(B))]
    ci, ct = models.get_code(model, loader, B)
    assert np.array_equal(unpack(ci, nbits), code_i) and np.array_equal(unpack(ct, nbits), code_t)
