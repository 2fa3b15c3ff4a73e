"""ORACLE (test infrastructure, never on the product path): plain-torch restatement of ``BertAdam.step``
(models/common/optimizer.py:102-165) and of ``torch.optim.SGD`` with momentum as the reference builds it for the HyP proxies
(runners/DSPH/runner.py:86-89).  Pinned by tests/golden/optimizer_golden.npz (the reference's optimiser classes executed in the
build container, tests/golden/make_optimizer_golden.py)."""
import math

import torch


def _schedule(name, x, warmup):
    if x < warmup:
        return x / warmup
    if name == "warmup_cosine":                                   # optimizer.py:25-28
        return 0.5 * (1.0 + math.cos(math.pi * x))
    if name == "warmup_constant":                                 # :30-35
        return 1.0
    return max((x - 1.0) / (warmup - 1.0), 0)                     # warmup_linear :37-42


def bert_adam_step(params, grads, state, step, lr, warmup, t_total, schedule, b1, b2, e, weight_decay, max_grad_norm):
    """One optimiser step over lists of fp32 tensors, in place; ``state`` = list of dicts with ``m`` and ``v``."""
    for p, g, s in zip(params, grads, state):
        if max_grad_norm > 0:                                     # :138-139 torch.nn.utils.clip_grad_norm_ on ONE tensor
            coef = max_grad_norm / (float(g.norm(2)) + 1e-6)
            if coef < 1:
                g.mul_(coef)
        s["m"].mul_(b1).add_(g, alpha=1 - b1)                      # :144
        s["v"].mul_(b2).addcmul_(g, g, value=1 - b2)               # :146
        update = s["m"] / (s["v"].sqrt() + e)                      # :147
        if weight_decay > 0.0:
            update += weight_decay * p                             # :155-156
        lr_t = lr * _schedule(schedule, step / t_total, warmup) if t_total != -1 else lr   # :158-163
        p.add_(-lr_t * update)                                     # :165-166


def sgd_momentum_step(params, grads, bufs, first, lr, momentum, weight_decay):
    for p, g, b in zip(params, grads, bufs):
        d = g + weight_decay * p if weight_decay != 0 else g.clone()
        if momentum != 0:
            if first:
                b.copy_(d)
            else:
                b.mul_(momentum).add_(d)
            d = b
        p.add_(d, alpha=-lr)
