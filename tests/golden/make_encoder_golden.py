#!/usr/bin/env python
"""Generate tests/golden/encoder_golden.npz by EXECUTING THE REFERENCE's CLIP and hash heads in this container.

Run (build container only — /root/reference does not exist on the GPU box):

    python tests/golden/make_encoder_golden.py

The reference classes are imported unmodified from /root/reference (``models/CLIP/model.py`` CLIP,
``models/DSPH/hash/hash.py`` and ``models/DCMHT/hash/hash.py`` HashLayer; ``ftfy``/``termcolor``/``xlrd`` are absent
from the image and stubbed in ``sys.modules`` — none of them is on the numeric path).  Weights come from
``synth.clip_state_dict`` (seeded CPU generator; identical on the GPU box, same torch build) and are loaded with
``load_state_dict(strict=True)`` so the key set is checked against the reference; a fingerprint of the weights is
stored so that a different RNG stream would be detected rather than silently compared.

Stored per case: inputs are regenerated from seeds (``synth.random_images`` / ``random_captions``); outputs are the
reference's fp32 CPU results of encode_image / encode_text (with and without ``return_patches``) and of the heads.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

for name, attrs in (("ftfy", {"fix_text": lambda s: s}), ("termcolor", {"colored": lambda s, *a, **k: s}),
                    ("xlrd", {})):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

from clip_based_cross_modal_hash_b200 import synth  # noqa: E402

from models.CLIP.model import CLIP  # noqa: E402  (the unmodified reference)
from models.DCMHT.hash.hash import HashLayer as DcmhtHash  # noqa: E402
from models.DSPH.hash.hash import HashLayer as DsphHash  # noqa: E402
from runners.base import BaseTrainer  # noqa: E402
from runners.DCMHT.runner import DCMHTTrainer  # noqa: E402


def fingerprint(sd: dict) -> np.ndarray:
    """Order-independent digest of a state dict: per key, sum and sum of squares in fp64."""
    rows = []
    for k in sorted(sd):
        t = sd[k].double().flatten()
        rows.append([float(t.sum()), float((t * t).sum())])
    return np.asarray(rows, dtype=np.float64)


def build_ref(cfg: dict, sd: dict, return_patches: bool) -> CLIP:
    model = CLIP(cfg["embed_dim"], cfg["image_resolution"], cfg["vision_layers"], cfg["vision_width"],
                 cfg["vision_patch_size"], cfg["context_length"], cfg["vocab_size"], cfg["transformer_width"],
                 cfg["transformer_heads"], cfg["transformer_layers"], return_patches=return_patches)
    model.load_state_dict(sd, strict=True)
    return model.float().eval()


def main() -> None:
    torch.manual_seed(0)
    out = {}
    for tag, cfg, nimg, ntxt in (("tiny", synth.TINY, 3, 5), ("vitb32", synth.VIT_B32, 2, 4)):
        sd = synth.clip_state_dict(cfg, seed=11)
        out[tag + "/fingerprint"] = fingerprint(sd)
        image = synth.random_images(nimg, seed=21)
        text, pad = synth.random_captions(ntxt, seed=22, vocab=cfg["vocab_size"])
        with torch.no_grad():
            plain = build_ref(cfg, sd, False)
            out[tag + "/img_cls"] = plain.encode_image(image).numpy()
            out[tag + "/txt_eos"] = plain.encode_text(text).numpy()
            out[tag + "/txt_eos_masked"] = plain.encode_text(text, key_padding_mask=pad).numpy()
            full = build_ref(cfg, sd, True)
            cls, seq, attn = full.encode_image(image)
            out[tag + "/img_cls_rp"], out[tag + "/img_seq"], out[tag + "/img_attn"] = cls.numpy(), seq.numpy(), attn.numpy()
            eos, tseq, tattn, newmask = full.encode_text(text, key_padding_mask=pad)
            out[tag + "/txt_eos_rp"], out[tag + "/txt_seq"] = eos.numpy(), tseq.numpy()
            out[tag + "/txt_attn"], out[tag + "/txt_newmask"] = tattn.numpy(), newmask.numpy()
        print(tag, "encoders done", flush=True)

    # hash heads on seeded feature batches (eval mode, as in get_code: runners/base.py:242-257, change_state "valid")
    g = torch.Generator().manual_seed(31)
    feat = torch.randn((6, 512), generator=g)
    for nbits in (16, 64):
        hsd = synth.dsph_head_state_dict(512, nbits, seed=41)
        head = DsphHash(inputDim=512, outputDim=nbits)
        head.load_state_dict(hsd, strict=True)
        head.eval()
        with torch.no_grad():
            hi, ht = head.encode_img(feat), head.encode_txt(feat)
            out["dsph%d/img" % nbits], out["dsph%d/txt" % nbits] = hi.numpy(), ht.numpy()
            out["dsph%d/img_code" % nbits] = BaseTrainer.make_hash_code(hi.clone()).numpy()
            out["dsph%d/txt_code" % nbits] = BaseTrainer.make_hash_code(ht.clone()).numpy()
        hsd = synth.dcmht_head_state_dict(512, nbits, seed=42)
        head = DcmhtHash(feature_size=512, outputDim=nbits, num_heads=8, batch_first=True, hash_func_="softmax")
        head.load_state_dict(hsd, strict=True)
        head.eval()
        with torch.no_grad():
            hi, ht = head.encode_img(feat), head.encode_txt(feat)
            out["dcmht%d/img" % nbits], out["dcmht%d/txt" % nbits] = hi.numpy(), ht.numpy()
            out["dcmht%d/img_code" % nbits] = DCMHTTrainer.make_hash_code(hi).numpy()
            out["dcmht%d/txt_code" % nbits] = DCMHTTrainer.make_hash_code(ht).numpy()
    out["head_feat"] = feat.numpy()
    path = os.path.join(HERE, "encoder_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
