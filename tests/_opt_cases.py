"""Seeded cases shared by tests/golden/make_optimizer_golden.py and the optimiser tests."""
import torch

# name, tensor shapes, steps, BertAdam keyword arguments (the DSPH config: configs/DSPH/config.yaml optimizer section)
OPT_CASES = [
    ("dsph_cfg", [(64, 512), (64,), (64, 512), (64,)], 4,
     dict(lr=1e-3, warmup=0.1, t_total=20, schedule="warmup_cosine", b1=0.9, b2=0.98, e=1e-6, weight_decay=0.2, max_grad_norm=1.0)),
    ("no_clip_linear", [(33,), (7, 5, 3), (20000,), (1,)], 3,
     dict(lr=5e-4, warmup=0.25, t_total=8, schedule="warmup_linear", b1=0.8, b2=0.999, e=1e-8, weight_decay=0.0, max_grad_norm=-1)),
    ("constant_lr", [(300, 70), (70,)], 2,
     dict(lr=2e-3, warmup=-1, t_total=-1, schedule="warmup_constant", b1=0.9, b2=0.999, e=1e-6, weight_decay=0.01, max_grad_norm=0.5)),
]


def opt_inputs(shapes, steps, seed):
    g = torch.Generator().manual_seed(seed)
    params = [torch.randn(s, generator=g) for s in shapes]
    grads = [[torch.randn(s, generator=g) * (3.0 if i % 2 == 0 else 0.05) for i, s in enumerate(shapes)] for _ in range(steps)]
    return params, grads
