"""Seeded inputs of the HyP-loss parity cases, shared by tests/golden/make_hyp_golden.py and tests/test_hyp_loss.py."""
import torch

CASES = [  # name, B, K, C, threshold, alpha, label density
    ("c5", 128, 64, 80, 0.0, 0.8, 0.05), ("thr", 96, 32, 24, 0.15, 0.8, 0.08), ("noreg", 64, 16, 21, 0.1, 0.0, 0.1),
    ("single", 50, 64, 10, 0.05, 0.8, 0.0),
]


def inputs(B, K, C, density, seed):
    g = torch.Generator().manual_seed(seed)
    x, y = torch.tanh(torch.randn((B, K), generator=g)), torch.tanh(torch.randn((B, K), generator=g))
    label = (torch.rand((B, C), generator=g) < density).long()
    label[torch.arange(B), torch.randint(0, C, (B,), generator=g)] = 1
    proxies = torch.randn((C, K), generator=g)
    return x, y, label, proxies
