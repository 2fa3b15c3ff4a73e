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
