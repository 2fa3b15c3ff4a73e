"""Seeded synthetic inputs of the shapes BASELINE.json names.

Host-side generators shared by tests and bench.py.  They produce tensors in the *reference's* formats:
codes as +-1 fp32 ``[n, K]`` (what ``runners/base.py:242-266 get_code`` hands to ``calc_map_k``) and
labels as int64 multi-hot ``[n, C]`` (``dataset/transformer_dataset.py:95-100``).
"""
from __future__ import annotations

import torch

# name -> (Q, N, K bits, C classes, k)  — BASELINE.json configs C1..C4
CONFIGS = {
    "C1": dict(Q=1_000, N=5_000, K=16, C=24, k=50),
    "C2": dict(Q=5_000, N=117_000, K=64, C=80, k=None),
    "C3": dict(Q=2_100, N=190_000, K=128, C=21, k=None),
    "C4-16": dict(Q=10_000, N=1_000_000, K=16, C=80, k=1_000),
    "C4-32": dict(Q=10_000, N=1_000_000, K=32, C=80, k=1_000),
    "C4-64": dict(Q=10_000, N=1_000_000, K=64, C=80, k=1_000),
    "C4-128": dict(Q=10_000, N=1_000_000, K=128, C=80, k=1_000),
}


def random_codes(n: int, nbits: int, seed: int, device="cpu") -> torch.Tensor:
    """i.i.d. Bernoulli(1/2) bits as +-1 fp32: distances ~ Binomial(K, 1/2), maximal tie pressure."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    codes = torch.randint(0, 2, (n, nbits), generator=g, dtype=torch.int8).to(torch.float32) * 2 - 1
    return codes.to(device)


def random_labels(n: int, ncls: int, seed: int, p: float = 0.07, device="cpu") -> torch.Tensor:
    """int64 multi-hot, each class Bernoulli(p) plus one forced class per row (no empty rows => no nan)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    lab = (torch.rand(n, ncls, generator=g) < p).to(torch.int64)
    forced = torch.randint(0, ncls, (n,), generator=g)
    lab[torch.arange(n), forced] = 1
    return lab.to(device)


def clustered_codes(n: int, nbits: int, seed: int, centers: int = 32, flip: float = 0.08) -> torch.Tensor:
    """Codes drawn around a few centres with bit-flip noise: what a trained hash head produces
    (many gallery items at tiny distances, long empty tail) — stresses the select thresholds."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    cen = torch.randint(0, 2, (centers, nbits), generator=g, dtype=torch.int8)
    pick = torch.randint(0, centers, (n,), generator=g)
    noise = (torch.rand(n, nbits, generator=g) < flip).to(torch.int8)
    return ((cen[pick] ^ noise).to(torch.float32)) * 2 - 1


# ---- encoder side: seeded weights, images and captions ----------------------------------------
VIT_B32 = dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768, vision_patch_size=32,
               context_length=77, vocab_size=49408, transformer_width=512, transformer_heads=8, transformer_layers=12)
# a 2-layer miniature with the same structure (head dim 64): fast CPU parity cases
TINY = dict(embed_dim=128, image_resolution=224, vision_layers=2, vision_width=128, vision_patch_size=32,
            context_length=77, vocab_size=1000, transformer_width=128, transformer_heads=2, transformer_layers=2)
SOT_ID, EOT_ID = 49406, 49407


def _normal(g, shape, std):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std


def clip_state_dict(cfg: dict = VIT_B32, seed: int = 0) -> dict:
    """A random ``state_dict`` with exactly the keys/shapes of the reference's ``CLIP`` (models/CLIP/model.py:270-340)
    for a ViT backbone; scales follow ``initialize_parameters`` (:342-371).  LayerNorm gains/biases and the Linear
    biases are randomised too (the reference initialises them to 1/0) so that parity tests exercise them."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    vw, tw, E = cfg["vision_width"], cfg["transformer_width"], cfg["embed_dim"]
    P, grid = cfg["vision_patch_size"], cfg["image_resolution"] // cfg["vision_patch_size"]

    def tower(prefix, width, layers):
        proj_std, attn_std, fc_std = (width ** -0.5) * ((2 * layers) ** -0.5), width ** -0.5, (2 * width) ** -0.5
        for i in range(layers):
            p = "%sresblocks.%d." % (prefix, i)
            sd[p + "attn.in_proj_weight"] = _normal(g, (3 * width, width), attn_std)
            sd[p + "attn.in_proj_bias"] = _normal(g, (3 * width,), 0.02)
            sd[p + "attn.out_proj.weight"] = _normal(g, (width, width), proj_std)
            sd[p + "attn.out_proj.bias"] = _normal(g, (width,), 0.02)
            sd[p + "ln_1.weight"] = 1 + _normal(g, (width,), 0.05)
            sd[p + "ln_1.bias"] = _normal(g, (width,), 0.05)
            sd[p + "mlp.c_fc.weight"] = _normal(g, (4 * width, width), fc_std)
            sd[p + "mlp.c_fc.bias"] = _normal(g, (4 * width,), 0.02)
            sd[p + "mlp.c_proj.weight"] = _normal(g, (width, 4 * width), proj_std)
            sd[p + "mlp.c_proj.bias"] = _normal(g, (width,), 0.02)
            sd[p + "ln_2.weight"] = 1 + _normal(g, (width,), 0.05)
            sd[p + "ln_2.bias"] = _normal(g, (width,), 0.05)

    sd["visual.class_embedding"] = _normal(g, (vw,), vw ** -0.5)
    sd["visual.positional_embedding"] = _normal(g, (grid * grid + 1, vw), vw ** -0.5)
    sd["visual.proj"] = _normal(g, (vw, E), vw ** -0.5)
    sd["visual.conv1.weight"] = _normal(g, (vw, 3, P, P), (3 * P * P) ** -0.5)
    sd["visual.ln_pre.weight"] = 1 + _normal(g, (vw,), 0.05)
    sd["visual.ln_pre.bias"] = _normal(g, (vw,), 0.05)
    tower("visual.transformer.", vw, cfg["vision_layers"])
    sd["visual.ln_post.weight"] = 1 + _normal(g, (vw,), 0.05)
    sd["visual.ln_post.bias"] = _normal(g, (vw,), 0.05)
    sd["positional_embedding"] = _normal(g, (cfg["context_length"], tw), 0.01)
    sd["text_projection"] = _normal(g, (tw, E), tw ** -0.5)
    sd["logit_scale"] = torch.tensor(2.6593)
    tower("transformer.", tw, cfg["transformer_layers"])
    sd["token_embedding.weight"] = _normal(g, (cfg["vocab_size"], tw), 0.02)
    sd["ln_final.weight"] = 1 + _normal(g, (tw,), 0.05)
    sd["ln_final.bias"] = _normal(g, (tw,), 0.05)
    return sd


def dsph_head_state_dict(in_dim: int, nbits: int, seed: int = 0) -> dict:
    """Keys of models/DSPH/hash/hash.py HashLayer (img_hash.fc / txt_hash.fc Linear(in_dim, nbits))."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    for m in ("img", "txt"):
        sd["%s_hash.fc.weight" % m] = _normal(g, (nbits, in_dim), in_dim ** -0.5)
        sd["%s_hash.fc.bias" % m] = _normal(g, (nbits,), 0.02)
    return sd


def dcmht_head_state_dict(in_dim: int, nbits: int, seed: int = 0) -> dict:
    """Keys of models/DCMHT/hash/hash.py HashLayer: per modality an nn.MultiheadAttention, a norm
    (BatchNorm1d for img, LayerNorm for txt, :58-59) and fc2 Linear(in_dim, 2*nbits)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    for m in ("img", "txt"):
        p = "%s_hash." % m
        sd[p + "atten.in_proj_weight"] = _normal(g, (3 * in_dim, in_dim), in_dim ** -0.5)
        sd[p + "atten.in_proj_bias"] = _normal(g, (3 * in_dim,), 0.02)
        sd[p + "atten.out_proj.weight"] = _normal(g, (in_dim, in_dim), in_dim ** -0.5)
        sd[p + "atten.out_proj.bias"] = _normal(g, (in_dim,), 0.02)
        sd[p + "norm.weight"] = 1 + _normal(g, (in_dim,), 0.05)
        sd[p + "norm.bias"] = _normal(g, (in_dim,), 0.05)
        if m == "img":
            sd[p + "norm.running_mean"] = _normal(g, (in_dim,), 0.1)
            sd[p + "norm.running_var"] = 1 + 0.2 * torch.rand((in_dim,), generator=g)
            sd[p + "norm.num_batches_tracked"] = torch.tensor(7)
        sd[p + "fc2.weight"] = _normal(g, (2 * nbits, in_dim), in_dim ** -0.5)
        sd[p + "fc2.bias"] = _normal(g, (2 * nbits,), 0.05)
    return sd


def random_images(batch: int, seed: int, resolution: int = 224) -> torch.Tensor:
    """fp32 NCHW with post-``Normalize`` statistics (dataset/transformer_dataset.py:41-45)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn((batch, 3, resolution, resolution), generator=g, dtype=torch.float32)


def random_images_u8(batch: int, seed: int, resolution: int = 224) -> torch.Tensor:
    """uint8 NCHW pixels: what the dataloader holds after Resize / CenterCrop, before ToTensor + Normalize
    (dataset/transformer_dataset.py:41-45)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randint(0, 256, (batch, 3, resolution, resolution), generator=g, dtype=torch.uint8)


def normalize_u8(images: torch.Tensor) -> torch.Tensor:
    """ToTensor + Normalize of dataset/transformer_dataset.py:40,44 on uint8 NCHW -> the fp32 tensor the reference feeds."""
    mean = torch.tensor((0.48145466, 0.4578275, 0.40821073)).view(1, 3, 1, 1)
    std = torch.tensor((0.26862954, 0.26130258, 0.27577711)).view(1, 3, 1, 1)
    return (images.to(torch.float32) / 255.0 - mean) / std


def random_captions(batch: int, seed: int, max_words: int = 32, vocab: int = 49408):
    """int64 [B, max_words]: SOT, uniform word ids, EOT at position in [3, max_words-2], zero padding; and the
    key_padding_mask (text == 0) of dataset/transformer_dataset.py:82-86.  For a reduced vocabulary the SOT/EOT ids
    are the two largest ids, so that argmax still finds EOT (models/CLIP/model.py:379)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sot, eot = vocab - 2, vocab - 1
    text = torch.zeros((batch, max_words), dtype=torch.int64)
    ends = torch.randint(3, max_words - 1, (batch,), generator=g)
    words = torch.randint(1, sot, (batch, max_words), generator=g)
    for b in range(batch):
        e = int(ends[b])
        text[b, 0] = sot
        text[b, 1:e] = words[b, 1:e]
        text[b, e] = eot
    return text, text == 0


def mith_head_state_dict(dim: int, nbits: int, seed: int = 0, layers: int = 2, res_mlp_layers: int = 2) -> dict:
    """Keys of models/MITH/hash/hash.py HashLayer (gcl_i == gcl_t shared, lct_i / lct_t with a `layers`-block Transformer,
    per-bit Linear(dim, 1) hashing, sin-cos `position.pe` buffer [nbits, 1, dim], concept projections)."""
    import math

    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    for i in range(res_mlp_layers):
        sd["gcl_i.mlp.mlps.%d.0.weight" % i] = _normal(g, (4 * dim, dim), dim ** -0.5)
        sd["gcl_i.mlp.mlps.%d.0.bias" % i] = _normal(g, (4 * dim,), 0.02)
        sd["gcl_i.mlp.mlps.%d.3.weight" % i] = _normal(g, (dim, 4 * dim), (4 * dim) ** -0.5)
        sd["gcl_i.mlp.mlps.%d.3.bias" % i] = _normal(g, (dim,), 0.02)
    for i in range(res_mlp_layers):
        sd["gcl_i.mlp.lns.%d.weight" % i] = 1 + _normal(g, (dim,), 0.05)
        sd["gcl_i.mlp.lns.%d.bias" % i] = _normal(g, (dim,), 0.05)
    sd["gcl_i.common_concept_embedding.weight"] = _normal(g, (nbits, dim), 2 * dim ** -0.5)
    for k in [k for k in sd if k.startswith("gcl_i.")]:
        sd["gcl_t." + k[6:]] = sd[k].clone()                       # one shared module in the reference (hash.py:217-218)
    pe = torch.zeros(nbits, dim)
    position = torch.arange(0, nbits, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2).float() * (-math.log(10000.0) / dim))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    pe = (pe / dim ** 0.5).unsqueeze(1)
    for t in ("lct_i.", "lct_t."):
        sd[t + "position.pe"] = pe.clone()
        proj_std, attn_std, fc_std = (dim ** -0.5) * ((2 * layers) ** -0.5), dim ** -0.5, (2 * dim) ** -0.5
        for i in range(layers):
            p = "%stransformer.resblocks.%d." % (t, i)
            sd[p + "attn.in_proj_weight"] = _normal(g, (3 * dim, dim), attn_std)
            sd[p + "attn.in_proj_bias"] = _normal(g, (3 * dim,), 0.02)
            sd[p + "attn.out_proj.weight"] = _normal(g, (dim, dim), proj_std)
            sd[p + "attn.out_proj.bias"] = _normal(g, (dim,), 0.02)
            sd[p + "ln_1.weight"] = 1 + _normal(g, (dim,), 0.05)
            sd[p + "ln_1.bias"] = _normal(g, (dim,), 0.05)
            sd[p + "mlp.c_fc.weight"] = _normal(g, (4 * dim, dim), fc_std)
            sd[p + "mlp.c_fc.bias"] = _normal(g, (4 * dim,), 0.02)
            sd[p + "mlp.c_proj.weight"] = _normal(g, (dim, 4 * dim), proj_std)
            sd[p + "mlp.c_proj.bias"] = _normal(g, (dim,), 0.02)
            sd[p + "ln_2.weight"] = 1 + _normal(g, (dim,), 0.05)
            sd[p + "ln_2.bias"] = _normal(g, (dim,), 0.05)
        for k in range(nbits):
            sd["%shashing.fc_list.%d.weight" % (t, k)] = _normal(g, (1, dim), dim ** -0.5)
            sd["%shashing.fc_list.%d.bias" % (t, k)] = _normal(g, (1,), 0.05)
    for m in ("img", "txt"):
        sd["%s_concept_proj.weight" % m] = _normal(g, (dim, dim), dim ** -0.5)
        sd["%s_concept_proj.bias" % m] = _normal(g, (dim,), 0.02)
    return sd
