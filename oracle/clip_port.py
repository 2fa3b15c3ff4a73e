"""ORACLE (test infrastructure, never on the product path): fp32 CPU restatement of the reference's CLIP
ViT-B/32 encoders and the DSPH / DCMHT hash heads, written as plain functions over a ``state_dict``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s reference / cpu_baseline legs may import this.

Every function cites the reference lines it restates (paths relative to the reference root).  The layout here
is batch-first ``[B, L, D]`` with explicit matmuls (the reference is ``[L, B, D]`` through
``nn.MultiheadAttention``); results are compared with the reference itself in
``tests/golden/make_encoder_golden.py`` -> ``tests/golden/encoder_golden.npz`` (parity pinned on outputs of the
reference executed in the build container; the reference ships no tests of its own).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

SD = Dict[str, torch.Tensor]
EOT_ID = 49407  # models/CLIP/model.py:384


def layer_norm(x: torch.Tensor, g: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """models/CLIP/model.py:153-159 — nn.LayerNorm evaluated in fp32 (biased variance, eps inside the sqrt)."""
    x = x.float()
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g.float() + b.float()


def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    """models/CLIP/model.py:162-164"""
    return x * torch.sigmoid(1.702 * x)


def attention(x: torch.Tensor, sd: SD, prefix: str, heads: int, causal: bool,
              key_padding_mask: Optional[torch.Tensor]):
    """nn.MultiheadAttention(x, x, x, need_weights=True, attn_mask, key_padding_mask) as called at
    models/CLIP/model.py:181-189.  x: [B, L, D].  Returns (out [B, L, D], head-averaged probabilities [B, L, L])."""
    B, L, D = x.shape
    dh = D // heads
    qkv = x @ sd[prefix + "in_proj_weight"].float().t() + sd[prefix + "in_proj_bias"].float()
    q, k, v = (t.reshape(B, L, heads, dh).permute(0, 2, 1, 3) for t in qkv.split(D, dim=-1))  # [B, H, L, dh]
    s = (q * (dh ** -0.5)) @ k.transpose(-1, -2)                                               # [B, H, L, L]
    if causal:  # build_attention_mask, models/CLIP/model.py:358-364: -inf strictly above the diagonal
        s = s + torch.full((L, L), float("-inf")).triu_(1)
    if key_padding_mask is not None:  # True = ignore that key
        s = s.masked_fill(key_padding_mask.bool()[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B, L, D)
    o = o @ sd[prefix + "out_proj.weight"].float().t() + sd[prefix + "out_proj.bias"].float()
    return o, p.mean(dim=1)


def resblock(x: torch.Tensor, sd: SD, prefix: str, heads: int, causal: bool, key_padding_mask):
    """ResidualAttentionBlock.forward, models/CLIP/model.py:191-197."""
    a, w = attention(layer_norm(x, sd[prefix + "ln_1.weight"], sd[prefix + "ln_1.bias"]), sd, prefix + "attn.",
                     heads, causal, key_padding_mask)
    x = x + a
    h = layer_norm(x, sd[prefix + "ln_2.weight"], sd[prefix + "ln_2.bias"])
    h = quick_gelu(h @ sd[prefix + "mlp.c_fc.weight"].float().t() + sd[prefix + "mlp.c_fc.bias"].float())
    x = x + (h @ sd[prefix + "mlp.c_proj.weight"].float().t() + sd[prefix + "mlp.c_proj.bias"].float())
    return x, w


def _num_layers(sd: SD, prefix: str) -> int:
    return len({k[len(prefix):].split(".")[0] for k in sd if k.startswith(prefix)})


def encode_image(sd: SD, image: torch.Tensor, return_patches: bool = False, trace: Optional[list] = None):
    """CLIP.encode_image -> VisionTransformer.forward, models/CLIP/model.py:232-268, 370-371.

    image [B, 3, R, R] fp32 -> cls [B, E]   (return_patches: (cls [B, E], seq [L-1, B, E], attn [B, L-1]))
    ``trace`` (optional list) receives the residual stream after ln_pre and after every block ([B, L, D])."""
    w = sd["visual.conv1.weight"].float()                       # [D, 3, P, P], stride P, no bias (:219)
    D, _, P, _ = w.shape
    B, C, R, _ = image.shape
    g = R // P
    # non-overlapping conv == one matmul over flattened patches; k = (c, ky, kx)
    patches = image.float().reshape(B, C, g, P, g, P).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, C * P * P)
    x = patches @ w.reshape(D, -1).t()                                                     # :235-238
    cls = sd["visual.class_embedding"].float().expand(B, 1, D)
    x = torch.cat([cls, x], dim=1) + sd["visual.positional_embedding"].float()             # :241-242
    x = layer_norm(x, sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"])                # :243
    if trace is not None:
        trace.append(x.clone())
    heads = D // 64                                                                         # :300
    attn = None
    for i in range(_num_layers(sd, "visual.transformer.resblocks.")):
        x, attn = resblock(x, sd, "visual.transformer.resblocks.%d." % i, heads, False, None)
        if trace is not None:
            trace.append(x.clone())
    x = layer_norm(x, sd["visual.ln_post.weight"], sd["visual.ln_post.bias"])              # :257 (all tokens)
    x = x @ sd["visual.proj"].float()                                                       # :259-260
    cls_token = x[:, 0]
    if return_patches:                                                                      # :263-267
        return cls_token, x[:, 1:].permute(1, 0, 2), attn[:, 0, 1:]
    return cls_token


def encode_text(sd: SD, text: torch.Tensor, key_padding_mask: Optional[torch.Tensor] = None,
                return_patches: bool = False, trace: Optional[list] = None):
    """CLIP.encode_text, models/CLIP/model.py:373-396.

    text [B, L] int64 -> eos [B, E]  (return_patches: (eos, seq [L, B, E], attn [B, L], new_mask [B, L] bool))."""
    B, L = text.shape
    x = sd["token_embedding.weight"].float()[text] + sd["positional_embedding"].float()[:L]  # :374-376
    D = x.shape[-1]
    heads = D // 64                                                                          # :465
    if trace is not None:
        trace.append(x.clone())
    attn = None
    for i in range(_num_layers(sd, "transformer.resblocks.")):
        x, attn = resblock(x, sd, "transformer.resblocks.%d." % i, heads, True, key_padding_mask)
        if trace is not None:
            trace.append(x.clone())
    eos = text.argmax(dim=-1)                                                                # :379
    rows = torch.arange(B)
    x = layer_norm(x, sd["ln_final.weight"], sd["ln_final.bias"]) @ sd["text_projection"].float()  # :386-388
    eos_token = x[rows, eos]                                                                 # :392
    if return_patches:
        a = attn[rows, eos].clone()                                                          # :381
        a[rows, eos] = 0                                                                     # :382
        new_mask = None if key_padding_mask is None else (key_padding_mask.bool() | (text == EOT_ID))  # :384
        return eos_token, x.permute(1, 0, 2), a, new_mask
    return eos_token


# ---- hash heads ---------------------------------------------------------------------------------------------
def dsph_head(hsd: SD, feat: torch.Tensor, modality: str) -> torch.Tensor:
    """models/DSPH/hash/hash.py:6-15 in eval mode (dropout = identity): tanh(Linear(feat)).  modality: img|txt."""
    p = "%s_hash.fc." % modality
    return torch.tanh(feat.float() @ hsd[p + "weight"].float().t() + hsd[p + "bias"].float())


def make_hash_code_sign(code: torch.Tensor) -> torch.Tensor:
    """BaseTrainer.make_hash_code, runners/base.py:407-410 (sign; 0 stays 0)."""
    return torch.sign(code)


def dcmht_head(hsd: SD, feat: torch.Tensor, modality: str, eps: float = 1e-5) -> torch.Tensor:
    """models/DCMHT/hash/hash.py:35-46 in eval mode -> [B, 2K] pairwise-softmax probabilities.

    Self-attention over a length-1 sequence has softmax == 1, so it reduces to out_proj(v_proj(x)); the image
    branch normalises with BatchNorm1d running statistics (eval), the text branch with LayerNorm (:58-59)."""
    p = "%s_hash." % modality
    x = feat.float()
    D = x.shape[-1]
    wv = hsd[p + "atten.in_proj_weight"].float()[2 * D:]
    bv = hsd[p + "atten.in_proj_bias"].float()[2 * D:]
    e = (x @ wv.t() + bv) @ hsd[p + "atten.out_proj.weight"].float().t() + hsd[p + "atten.out_proj.bias"].float()
    if (p + "norm.running_mean") in hsd:
        e = (e - hsd[p + "norm.running_mean"].float()) / torch.sqrt(hsd[p + "norm.running_var"].float() + eps)
        e = e * hsd[p + "norm.weight"].float() + hsd[p + "norm.bias"].float()
    else:
        e = layer_norm(e, hsd[p + "norm.weight"], hsd[p + "norm.bias"], eps)
    e = torch.relu(e @ hsd[p + "fc2.weight"].float().t() + hsd[p + "fc2.bias"].float())
    return torch.softmax(e.reshape(e.shape[0], -1, 2), dim=-1).reshape(e.shape[0], -1)   # models/common/hash.py:20-31


def make_hash_code_dcmht(code: torch.Tensor) -> torch.Tensor:
    """DCMHTTrainer.make_hash_code, runners/DCMHT/runner.py:83-95: argmax over each (2j, 2j+1) pair, 0 -> -1."""
    pairs = code.reshape(code.shape[0], -1, 2)
    return torch.where(pairs[..., 1] > pairs[..., 0], 1.0, -1.0)


def flops_image(width=768, layers=12, L=50, patch=32, out=512) -> float:
    """Algorithmic FLOPs per image: CLS-only final projection."""
    per = 2 * L * width * 3 * width + 2 * 2 * L * L * width + 2 * L * width * width + 2 * 2 * L * width * 4 * width
    return 2 * (L - 1) * 3 * patch * patch * width + layers * per + 2 * width * out


def flops_text(width=512, layers=12, L=32, out=512) -> float:
    per = 2 * L * width * 3 * width + 2 * 2 * L * L * width + 2 * L * width * width + 2 * 2 * L * width * 4 * width
    return layers * per + 2 * width * out
