#!/usr/bin/env python
"""Generate tests/golden/mith_golden.npz by EXECUTING THE REFERENCE's MITH HashLayer (models/MITH/hash/hash.py) in this
container (build container only; /root/reference does not exist on the GPU box):

    python tests/golden/make_mith_golden.py

Weights: ``synth.mith_head_state_dict`` loaded with ``load_state_dict(strict=True)``; inputs: seeded token tensors in the
reference's layouts (cls [B, 512], tokens [L, B, 512], key_padding_mask [B, L]).  Outputs: the four tensors of
``encode_img`` / ``encode_txt`` and the sign code of ``MITHTrainer.generate_hash``.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference")
for name, attrs in (("ftfy", {"fix_text": lambda s: s}), ("termcolor", {"colored": lambda s, *a, **k: s}), ("xlrd", {})):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

from clip_based_cross_modal_hash_b200 import synth  # noqa: E402
from models.MITH.hash.hash import HashLayer  # noqa: E402  (the unmodified reference)


def inputs(B, L, seed, padded):
    g = torch.Generator().manual_seed(seed)
    cls = torch.randn((B, 512), generator=g)
    tokens = torch.randn((L, B, 512), generator=g)
    mask = None
    if padded:
        lens = torch.randint(2, L + 1, (B,), generator=g)
        mask = torch.arange(L)[None, :] >= lens[:, None]
    return cls, tokens, mask


def main():
    out = {}
    for nbits in (16, 64):
        hsd = synth.mith_head_state_dict(512, nbits, seed=51)
        head = HashLayer(clip_embed_dim=512, k_bits=nbits, dropout=0.0, transformer_layers=2, activation="gelu", top_k_label=8,
                         res_mlp_layers=2)
        head.load_state_dict(hsd, strict=True)
        head.eval()
        with torch.no_grad():
            cls, tokens, _ = inputs(5, 49, 61, False)
            r = head.encode_img(img_cls=cls, img_tokens=tokens)
            for name, v in zip(("res", "cls_hash", "tok_hash", "trans"), r):
                out["mith%d/img_%s" % (nbits, name)] = v.numpy()
            out["mith%d/img_code" % nbits] = (r[1] + r[2]).sign_().numpy()
            cls, tokens, mask = inputs(6, 32, 62, True)
            r = head.encode_txt(cls, tokens, mask)
            for name, v in zip(("res", "cls_hash", "tok_hash", "trans"), r):
                out["mith%d/txt_%s" % (nbits, name)] = v.numpy()
            out["mith%d/txt_code" % nbits] = (r[1] + r[2]).sign_().numpy()
    path = os.path.join(HERE, "mith_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
