"""GPU: the CLIP encoder path (through the C ABI) against the fp32 oracle (oracle/clip_port.py) and against outputs of
the reference itself (tests/golden/encoder_golden.npz).

Stated tolerance (calibrated on the reference under bf16 autocast: rel-L2 1.0e-2, cosine 0.99994):
final 512-d features rel-L2 <= 2e-2 and cosine >= 0.9995 per row; per-kernel tests are tighter and say so.
Hash bits: every bit that differs from the oracle must have a pre-activation smaller than the observed feature error
allows (|tanh| / |p1-p0| below a small margin)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from clip_based_cross_modal_hash_b200 import _lib, encoder, models, synth
from oracle import clip_port as port

pytestmark = pytest.mark.gpu
Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encoder_golden.npz"))
REL_L2, COSINE = 2e-2, 0.9995


def st():
    return torch.cuda.current_stream().cuda_stream


def check_features(got, want, rel=REL_L2, cos=COSINE):
    got, want = got.detach().float().cpu().reshape(-1, got.shape[-1]), torch.as_tensor(want).float().reshape(-1, got.shape[-1])
    assert got.shape == want.shape
    assert torch.isfinite(got).all()
    r = ((got - want).norm(dim=-1) / want.norm(dim=-1).clamp_min(1e-6)).max().item()
    c = torch.nn.functional.cosine_similarity(got, want, dim=-1).min().item()
    assert r <= rel and c >= cos, (r, c)
    return r, c


# ---- kernels -------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,D", [(1, 128), (50, 768), (12800, 768), (8192, 512), (77, 1024), (333, 384), (9, 256)])
@pytest.mark.parametrize("out_f32", [0, 1])
def test_layernorm_matches_torch(rows, D, out_f32):
    g = torch.Generator(device="cuda").manual_seed(rows + D)
    x = torch.randn(rows, D, device="cuda", generator=g) * 3 + 0.5
    gain, bias = torch.randn(D, device="cuda", generator=g), torch.randn(D, device="cuda", generator=g)
    out = torch.empty(rows, D, device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    _lib.check(_lib.lib().cmh_layernorm(x.data_ptr(), rows, D, gain.data_ptr(), bias.data_ptr(), 1e-5, out.data_ptr(), out_f32, st()))
    want = torch.nn.functional.layer_norm(x, (D,), gain, bias, 1e-5)
    if out_f32:
        assert torch.allclose(out, want, rtol=1e-5, atol=2e-5)
    else:  # one bf16 rounding of the fp32 result
        assert torch.allclose(out.float(), want, rtol=2 ** -8, atol=1e-3)


def attention_reference(qkv, B, L, H, pad, causal):
    D = H * 64
    q, k, v = (t.reshape(B, L, H, 64).permute(0, 2, 1, 3) for t in qkv.float().reshape(B, L, 3 * D).split(D, dim=-1))
    s = (q @ k.transpose(-1, -2)) * 0.125
    if causal:
        s = s + torch.full((L, L), float("-inf"), device=qkv.device).triu_(1)
    if pad is not None:
        s = s.masked_fill(pad.bool()[:, None, None, :], float("-inf"))
    return (torch.softmax(s, dim=-1) @ v).permute(0, 2, 1, 3).reshape(B * L, D)


@pytest.mark.parametrize("B,L,H,causal,padded", [
    (3, 50, 12, 0, False), (256, 50, 12, 0, False), (5, 32, 8, 1, True), (64, 32, 8, 1, False), (2, 77, 8, 1, True),
    (4, 16, 2, 0, True), (3, 128, 2, 1, False), (2, 1, 2, 0, False), (7, 33, 4, 1, True), (2, 64, 2, 0, False)])
def test_attention_matches_fp32_reference(B, L, H, causal, padded):
    g = torch.Generator(device="cuda").manual_seed(B * 131 + L)
    qkv = (torch.randn(B * L, 3 * H * 64, device="cuda", generator=g) * 1.5).to(torch.bfloat16)
    pad = None
    if padded:
        lens = torch.randint(1, L + 1, (B,), device="cuda", generator=g)
        pad = (torch.arange(L, device="cuda")[None, :] >= lens[:, None]).to(torch.uint8).contiguous()
    out = torch.empty(B * L, H * 64, device="cuda", dtype=torch.bfloat16)
    _lib.check(_lib.lib().cmh_attention_bf16(qkv.data_ptr(), B, L, H, None if pad is None else pad.data_ptr(), causal,
                                             out.data_ptr(), st()))
    want = attention_reference(qkv, B, L, H, pad, causal)
    # P is rounded to bf16 before P.V and the output once more: 2 roundings of 2^-9 relative on O(1) values
    assert torch.allclose(out.float(), want, rtol=2e-2, atol=2e-2), float((out.float() - want).abs().max())
    assert float((out.float() - want).abs().mean()) < 3e-3


# ---- full towers ---------------------------------------------------------------------------------------------------
CASES = {"tiny": (synth.TINY, 3, 5), "vitb32": (synth.VIT_B32, 2, 4)}


@pytest.fixture(scope="module", params=list(CASES))
def case(request):
    cfg, nimg, ntxt = CASES[request.param]
    sd = synth.clip_state_dict(cfg, seed=11)
    text, pad = synth.random_captions(ntxt, seed=22, vocab=cfg["vocab_size"])
    return request.param, sd, synth.random_images(nimg, seed=21), text, pad


def test_encode_image_matches_reference_golden(case):
    tag, sd, image, _, _ = case
    bb = encoder.ClipBackbone(sd, return_patches=False)
    cls = bb.encode_image(image.cuda())
    assert cls.dtype == torch.float32 and cls.is_cuda and tuple(cls.shape) == Z[tag + "/img_cls"].shape
    check_features(cls, Z[tag + "/img_cls"])
    bb = encoder.ClipBackbone(sd, return_patches=True)
    cls2, seq, attn = bb.encode_image(image)          # host tensor in: copied by the shim
    assert torch.equal(cls2, cls)
    assert tuple(seq.shape) == Z[tag + "/img_seq"].shape and tuple(attn.shape) == Z[tag + "/img_attn"].shape
    check_features(seq, Z[tag + "/img_seq"])
    assert np.allclose(attn.cpu().numpy(), Z[tag + "/img_attn"], rtol=0.1, atol=2e-3)
    assert np.allclose(attn.sum(-1).cpu().numpy(), Z[tag + "/img_attn"].sum(-1), atol=2e-3)


def test_encode_text_matches_reference_golden(case):
    tag, sd, _, text, pad = case
    bb = encoder.ClipBackbone(sd, return_patches=False)
    check_features(bb.encode_text(text.cuda()), Z[tag + "/txt_eos"])
    check_features(bb.encode_text(text, key_padding_mask=pad), Z[tag + "/txt_eos_masked"])
    bb = encoder.ClipBackbone(sd, return_patches=True)
    eos, seq, attn, newmask = bb.encode_text(text.cuda(), key_padding_mask=pad.cuda())
    check_features(eos, Z[tag + "/txt_eos_rp"])
    assert tuple(seq.shape) == Z[tag + "/txt_seq"].shape
    check_features(seq, Z[tag + "/txt_seq"])
    assert np.allclose(attn.cpu().numpy(), Z[tag + "/txt_attn"], rtol=0.1, atol=2e-3)
    assert np.array_equal(newmask.cpu().numpy(), Z[tag + "/txt_newmask"])


@pytest.mark.parametrize("batch", [1, 7, 64])
def test_towers_match_oracle_on_fresh_batches(batch):
    """Other batch sizes than the golden file's, tiny config (the CPU oracle finishes in seconds)."""
    sd = synth.clip_state_dict(synth.TINY, seed=5)
    bb = encoder.ClipBackbone(sd)
    image = synth.random_images(batch, seed=batch)
    text, pad = synth.random_captions(batch, seed=batch + 1, vocab=synth.TINY["vocab_size"])
    with torch.no_grad():
        check_features(bb.encode_image(image), port.encode_image(sd, image))
        check_features(bb.encode_text(text, pad), port.encode_text(sd, text, pad))
    # determinism: same input, same bits
    assert torch.equal(bb.encode_image(image), bb.encode_image(image))


def test_full_batch_properties_vit_b32():
    """BASELINE size (batch 256, ViT-B/32), size-independent properties instead of a CPU oracle run:
    batch invariance (a sample's features do not depend on what else is in the batch or where it sits: every output
    element has a fixed K-order accumulation), permutation equivariance, and run-to-run determinism — all bit-exact."""
    sd = synth.clip_state_dict(synth.VIT_B32, seed=11)
    bb = encoder.ClipBackbone(sd)
    image = synth.random_images(256, seed=123).cuda()
    text, pad = synth.random_captions(256, seed=124)
    text = text.cuda()
    fi, ft = bb.encode_image(image), bb.encode_text(text)
    assert torch.isfinite(fi).all() and torch.isfinite(ft).all()
    assert torch.equal(fi, bb.encode_image(image)) and torch.equal(ft, bb.encode_text(text))
    assert torch.equal(bb.encode_image(image[:7]), fi[:7]) and torch.equal(bb.encode_text(text[:7]), ft[:7])
    assert torch.equal(bb.encode_image(image[100:133]), fi[100:133])
    perm = torch.randperm(256, generator=torch.Generator().manual_seed(5)).cuda()
    assert torch.equal(bb.encode_image(image[perm]), fi[perm]) and torch.equal(bb.encode_text(text[perm]), ft[perm])
    # and the first rows against the fp32 oracle (the CPU finishes 4 samples in seconds)
    with torch.no_grad():
        check_features(fi[:4], port.encode_image(sd, image[:4].cpu()))
        check_features(ft[:4], port.encode_text(sd, text[:4].cpu()))


def test_residual_stream_error_stays_bounded_per_block():
    """Per-block check of the ViT-B/32 image tower against the oracle trace: the bf16 error must not compound."""
    sd = synth.clip_state_dict(synth.VIT_B32, seed=11)
    image = synth.random_images(2, seed=77)
    trace = []
    with torch.no_grad():
        port.encode_image(sd, image, trace=trace)
    for layers in (1, 4, 12):
        sub = {k: v for k, v in sd.items() if not (k.startswith("visual.transformer.resblocks.") and int(k.split(".")[3]) >= layers)}
        want = trace[layers]                                            # residual stream after `layers` blocks
        # compare through ln_post + proj of all tokens (the ABI does not expose the raw stream)
        want = port.layer_norm(want, sd["visual.ln_post.weight"], sd["visual.ln_post.bias"]) @ sd["visual.proj"].float()
        bb = encoder.ClipBackbone(sub, return_patches=True)
        cls, seq, _ = bb.encode_image(image)
        check_features(seq.permute(1, 0, 2), want[:, 1:])
        check_features(cls, want[:, 0])


# ---- heads ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nbits", [16, 64])
def test_heads_match_reference_golden(nbits):
    feat = torch.from_numpy(Z["head_feat"]).cuda()
    for name, layer, sdfn, seed in (("dsph", models.DsphHashLayer, synth.dsph_head_state_dict, 41),
                                    ("dcmht", models.DcmhtHashLayer, synth.dcmht_head_state_dict, 42)):
        head = layer(sdfn(512, nbits, seed=seed), "cuda")
        for m, fn in (("img", head.encode_img), ("txt", head.encode_txt)):
            want = Z["%s%d/%s" % (name, nbits, m)]
            got = fn(feat)
            assert np.allclose(got.cpu().numpy(), want, rtol=1e-5, atol=2e-6), np.abs(got.cpu().numpy() - want).max()
            _, packed = head._run(feat, m, True)
            code = Z["%s%d/%s_code" % (name, nbits, m)]                 # +-1 from the reference's make_hash_code
            bits = ((packed.cpu().numpy().view(np.uint32)[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(len(code), -1)[:, :nbits]
            assert np.array_equal(bits.astype(np.float32) * 2 - 1, code)


@pytest.mark.parametrize("method", ["DSPH", "DCMHT"])
def test_model_codes_agree_with_oracle(method):
    """encode -> head -> packed bits on ViT-B/32; a flipped bit is only allowed where the oracle's margin is tiny."""
    nbits, B = 64, 6
    sd = synth.clip_state_dict(synth.VIT_B32, seed=11)
    hsd = (synth.dsph_head_state_dict if method == "DSPH" else synth.dcmht_head_state_dict)(512, nbits, seed=3)
    model = getattr(models, method)(sd, hsd)
    image = synth.random_images(B, seed=91)
    text, _ = synth.random_captions(B, seed=92)
    with torch.no_grad():
        fi, ft = port.encode_image(sd, image), port.encode_text(sd, text)
        if method == "DSPH":
            oi, ot = port.dsph_head(hsd, fi, "img"), port.dsph_head(hsd, ft, "txt")
            mi, mt = oi, ot                                              # margin = tanh value
        else:
            oi, ot = port.dcmht_head(hsd, fi, "img"), port.dcmht_head(hsd, ft, "txt")
            mi = oi.reshape(B, -1, 2)[..., 1] - oi.reshape(B, -1, 2)[..., 0]
            mt = ot.reshape(B, -1, 2)[..., 1] - ot.reshape(B, -1, 2)[..., 0]
    hi, ht = model(image, text)
    assert torch.allclose(hi.cpu(), oi, atol=5e-2) and torch.allclose(ht.cpu(), ot, atol=5e-2)
    code_i = model.make_hash_code(hi.clone()).cpu()
    for packed, margin, code in ((model.encode_image_packed(image), mi, code_i), (model.encode_text_packed(text), mt, None)):
        bits = ((packed.cpu().numpy().view(np.uint32)[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(B, -1)[:, :nbits]
        want = (margin > 0).numpy()
        flipped = bits.astype(bool) != want
        assert np.abs(margin.numpy()[flipped]).max(initial=0.0) < 5e-2, "a bit with a clear margin flipped"
        assert flipped.mean() < 0.05
        if code is not None:  # packed bits == make_hash_code of the float output of the same path
            assert np.array_equal(bits.astype(np.float32) * 2 - 1, code.numpy())


def test_get_code_matches_batchwise_encoding():
    """models.get_code (pipelined H2D, scatter by dataset index) == per-batch packed encoding, rows placed by `index`."""
    nbits, B, nb = 32, 5, 3
    sd = synth.clip_state_dict(synth.TINY, seed=8)
    model = models.DSPH(sd, synth.dsph_head_state_dict(synth.TINY["embed_dim"], nbits, seed=9))
    perm = torch.randperm(B * nb, generator=torch.Generator().manual_seed(0))
    loader = []
    for i in range(nb):
        text, pad = synth.random_captions(B, seed=40 + i, vocab=synth.TINY["vocab_size"])
        loader.append((synth.random_images(B, seed=30 + i).pin_memory(), text.pin_memory(), pad, None, perm[i * B:(i + 1) * B]))
    img_codes, txt_codes = models.get_code(model, loader, B * nb)
    assert img_codes.shape == (B * nb, 1) and img_codes.dtype == torch.int32
    for image, text, _, _, index in loader:
        assert torch.equal(img_codes[index.cuda()], model.encode_image_packed(image))
        assert torch.equal(txt_codes[index.cuda()], model.encode_text_packed(text))


def test_encode_to_map_pipeline_matches_float_path():
    """get_code (packed) -> calc_map_k_packed == reference-shaped path: float head outputs -> make_hash_code -> calc_map_k."""
    from clip_based_cross_modal_hash_b200 import calc_utils

    nbits, B, nb, C = 32, 16, 4, 10
    sd = synth.clip_state_dict(synth.TINY, seed=12)
    model = models.DSPH(sd, synth.dsph_head_state_dict(synth.TINY["embed_dim"], nbits, seed=13))
    loader, img_f, txt_f = [], [], []
    for i in range(nb):
        text, pad = synth.random_captions(B, seed=60 + i, vocab=synth.TINY["vocab_size"])
        image = synth.random_images(B, seed=50 + i)
        loader.append((image, text, pad, None, torch.arange(B) + i * B))
        hi, ht = model(image, text)
        img_f.append(model.make_hash_code(hi.clone()))
        txt_f.append(model.make_hash_code(ht.clone()))
    img_codes, txt_codes = models.get_code(model, loader, B * nb)
    labels = synth.random_labels(B * nb, C, seed=70)
    img_f, txt_f = torch.cat(img_f), torch.cat(txt_f)
    assert bool((img_f.abs() == 1).all()), "a tanh output was exactly 0"
    q = slice(0, B)   # first batch queries the whole set
    want = calc_utils.calc_map_k(img_f[q], txt_f, labels[q], labels, 20)
    got = calc_utils.calc_map_k_packed(img_codes[q], txt_codes, labels[q], labels, nbits, 20)
    assert got.dtype == torch.float32 and got.device.type == "cpu"
    assert float(got) == float(want)


def test_registry_surface_from_config_and_checkpoint_roundtrip(tmp_path):
    """BaseTrainer.build_model's calls (runners/base.py:98-107): from_config(cfg, output_dim, train_num) reads the CLIP
    checkpoint from cfg['clip_path']; load_state_dict takes a trained checkpoint with backbone.* / hash.* keys."""
    sd = synth.clip_state_dict(synth.TINY, seed=21)
    ckpt = tmp_path / "clip_tiny.pt"
    torch.save(sd, ckpt)                                  # load_backbone falls back to torch.load (models/base.py:23-24)
    model = models.DSPH.from_config({"clip_path": str(ckpt)}, output_dim=32, train_num=100).float().to("cuda")
    assert model.output_dim == 32 and hasattr(model, "backbone") and hasattr(model, "hash")
    trained = models.DSPH(sd, synth.dsph_head_state_dict(synth.TINY["embed_dim"], 32, seed=99))
    full = trained.state_dict()
    assert any(k.startswith("backbone.visual.") for k in full) and "hash.img_hash.fc.weight" in full
    full["hyp.proxies"] = torch.zeros(3, 32)              # loss parameters of a real checkpoint are ignored
    model.load_state_dict(full)
    image = synth.random_images(3, seed=1)
    assert torch.equal(model.encode_image(image), trained.encode_image(image))


def test_encoder_rejects_bad_arguments():
    sd = synth.clip_state_dict(synth.TINY, seed=1)
    bb = encoder.ClipBackbone(sd)
    with pytest.raises(ValueError):
        bb.encode_image(torch.zeros(2, 3, 64, 64))
    with pytest.raises(_lib.CmhError):   # longer than the 128-token kernel limit / context
        bb.encode_text(torch.zeros(2, 200, dtype=torch.int64))
    with pytest.raises(_lib.CmhError):   # workspace too small
        c = bb.visual.c
        out = torch.empty(2, c.out_dim, device="cuda")
        ws = torch.empty(1024, dtype=torch.uint8, device="cuda")
        _lib.check(_lib.lib().cmh_encode_image(ctypes.byref(c), torch.zeros(2, 3, 224, 224, device="cuda").data_ptr(), 2,
                                               ws.data_ptr(), ws.numel(), out.data_ptr(), None, None, st()))


def test_uint8_images_are_normalised_on_the_gpu():
    """uint8 pixels + GPU-side ToTensor/Normalize (dataset/transformer_dataset.py:41-45) == feeding the normalised fp32 tensor."""
    sd = synth.clip_state_dict(synth.TINY, seed=3)
    bb = encoder.ClipBackbone(sd)
    u8 = synth.random_images_u8(6, seed=4)
    want = bb.encode_image(synth.normalize_u8(u8))
    got = bb.encode_image(u8)
    # same bf16 patches up to the last fp32 ulp of the affine map before rounding: features agree far inside the bf16 tolerance
    assert (got - want).norm() / want.norm() < 2e-3
    oracle = port.encode_image(sd, synth.normalize_u8(u8))
    check_features(got.cpu(), oracle)
    model = models.DSPH(sd, synth.dsph_head_state_dict(synth.TINY["embed_dim"], 32, seed=9))
    text, pad = synth.random_captions(6, seed=40, vocab=synth.TINY["vocab_size"])
    loader = [(u8.pin_memory(), text.pin_memory(), pad, None, torch.arange(6))]
    codes_u8, _ = models.get_code(model, loader, 6)
    assert torch.equal(codes_u8, model.encode_image_packed(u8))
