"""ORACLE (test infrastructure, never on the product path): fp32 CPU restatement of the MITH hash head
(``models/MITH/hash/hash.py``) in evaluation mode, as plain functions over the head's ``state_dict``.

Only ``tests/`` and ``__graft_entry__.smoke()`` may import this.  Layout is batch-first (the reference is [L, N, D]);
results are compared with the reference ``HashLayer`` itself in ``tests/golden/make_mith_golden.py`` ->
``tests/golden/mith_golden.npz`` (parity pinned on outputs of the reference executed in the build container).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import clip_port

SD = Dict[str, torch.Tensor]


def res_mlps(sd: SD, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """ResidualMLPs.forward, hash.py:9-38: x + Linear(4D->D)(GELU(Linear(D->4D)(LayerNorm(x)))) per layer (exact erf GELU)."""
    x = x.float()
    i = 0
    while "%smlps.%d.0.weight" % (prefix, i) in sd:
        h = clip_port.layer_norm(x, sd["%slns.%d.weight" % (prefix, i)], sd["%slns.%d.bias" % (prefix, i)])
        h = F.gelu(h @ sd["%smlps.%d.0.weight" % (prefix, i)].float().t() + sd["%smlps.%d.0.bias" % (prefix, i)].float())
        x = x + (h @ sd["%smlps.%d.3.weight" % (prefix, i)].float().t() + sd["%smlps.%d.3.bias" % (prefix, i)].float())
        i += 1
    return x


def global_concept(sd: SD, g: str, x: torch.Tensor):
    """GlobalConceptLearning.forward, hash.py:86-106 -> (mlp(x), tanh(concept embedding))."""
    x = res_mlps(sd, g + "mlp.", x)
    return x, torch.tanh(x @ sd[g + "common_concept_embedding.weight"].float().t())


def token_aggregation(x: torch.Tensor, sim: torch.Tensor, key_padding_mask: Optional[torch.Tensor], top_k: int) -> torch.Tensor:
    """LocalizedTokenAggregation.forward, hash.py:109-170.  x [B, L, D], sim [B, L, K] -> merged tokens [B, K, D].

    padded tokens and non-positive similarities are dropped (-inf), each token keeps its top_k concepts (ties at the
    k-th value kept, :114-124), softmax over the TOKENS of every (sample, concept) with an all -inf column giving 0
    (:159-160), then the weighted sum of the tokens (:164-168)."""
    sim = sim.float().clone()
    if key_padding_mask is not None:
        sim = sim.masked_fill(key_padding_mask.bool()[:, :, None], float("-inf"))
    neg = torch.full_like(sim, float("-inf"))
    sim = torch.where(sim > 0, sim, neg)
    kth = torch.topk(sim, k=top_k, dim=-1).values.min(dim=-1, keepdim=True).values
    sim = torch.where(sim >= kth, sim, neg)
    p = torch.softmax(sim, dim=1)
    p = torch.where(torch.isnan(p), torch.zeros_like(p), p)
    return p.transpose(1, 2) @ x.float()


def encode(sd: SD, modality: str, cls: torch.Tensor, tokens: torch.Tensor, key_padding_mask: Optional[torch.Tensor] = None,
           top_k: int = 8):
    """HashLayer.encode_img / encode_txt, hash.py:231-247.  cls [B, D]; tokens [L, B, D] (the reference's layout);
    returns (res_cls [B, D] normalised, cls_hash [B, K], tokens_hash [B, K], trans_tokens [K, B, D] normalised)."""
    g, t = ("gcl_i.", "lct_i.") if modality == "img" else ("gcl_t.", "lct_t.")
    res, cls_hash = global_concept(sd, g, cls)
    res = F.normalize(res, dim=-1)
    x = tokens.float().permute(1, 0, 2)                                         # [B, L, D]
    concept = global_concept(sd, g, x)[1]                                       # [B, L, K]
    merged = token_aggregation(x, concept, key_padding_mask, top_k)             # [B, K, D]
    K = merged.shape[1]
    y = merged + sd[t + "position.pe"].float()[:K, 0][None]                     # PositionalEncoding, hash.py:41-64
    i = 0
    while "%stransformer.resblocks.%d.ln_1.weight" % (t, i) in sd:               # Transformer(width, layers, width // 64)
        y, _ = clip_port.resblock(y, sd, "%stransformer.resblocks.%d." % (t, i), y.shape[-1] // 64, False, None)
        i += 1
    w = torch.stack([sd["%shashing.fc_list.%d.weight" % (t, k)].float()[0] for k in range(K)])   # [K, D]
    b = torch.stack([sd["%shashing.fc_list.%d.bias" % (t, k)].float()[0] for k in range(K)])     # [K]
    tokens_hash = torch.tanh((y * w[None]).sum(-1) + b[None])                   # BitwiseHashing, hash.py:67-83
    proj = "img_concept_proj." if modality == "img" else "txt_concept_proj."
    trans = F.normalize(y @ sd[proj + "weight"].float().t() + sd[proj + "bias"].float(), dim=-1)
    return res, cls_hash, tokens_hash, trans.permute(1, 0, 2)


def generate_hash(cls_hash: torch.Tensor, tokens_hash: torch.Tensor) -> torch.Tensor:
    """MITHTrainer.generate_hash + make_hash_code, runners/MITH/runner.py:125-131, runners/base.py:407-410."""
    return torch.sign(cls_hash + tokens_hash)
