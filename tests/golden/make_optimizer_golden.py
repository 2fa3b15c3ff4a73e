#!/usr/bin/env python
"""tests/golden/optimizer_golden.npz: the reference's BertAdam (models/common/optimizer.py) and the SGD it builds for the HyP
proxies, run for a few steps on seeded tensors and gradients (build container only); also HyP gradients by autograd of the
reference's HyP module."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference")
from common.register import registry  # noqa: E402,F401  (optimizer.py registers itself)

spec = importlib.util.spec_from_file_location("ref_opt", "/root/reference/models/common/optimizer.py")
ref_opt = importlib.util.module_from_spec(spec)
try:
    spec.loader.exec_module(ref_opt)
except KeyError:           # "already registered" when the package imported it first
    ref_opt = sys.modules.get("models.common.optimizer") or ref_opt
spec = importlib.util.spec_from_file_location("ref_hyp", "/root/reference/models/DSPH/loss/HyP.py")
ref_hyp = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_hyp)

from tests._hyp_cases import CASES, inputs  # noqa: E402
from tests._opt_cases import OPT_CASES, opt_inputs  # noqa: E402


def main():
    out = {}
    for name, shapes, steps, kw in OPT_CASES:
        params, grads = opt_inputs(shapes, steps, seed=7)
        ps = [torch.nn.Parameter(p.clone()) for p in params]
        groups = [{"params": ps[: len(ps) // 2], "lr": kw["lr"] * 0.01}, {"params": ps[len(ps) // 2:], "lr": kw["lr"]}]
        opt = ref_opt.BertAdam(groups, lr=kw["lr"], warmup=kw["warmup"], t_total=kw["t_total"], schedule=kw["schedule"], b1=kw["b1"],
                               b2=kw["b2"], e=kw["e"], weight_decay=kw["weight_decay"], max_grad_norm=kw["max_grad_norm"])
        for s in range(steps):
            for p, g in zip(ps, grads[s]):
                p.grad = g.clone()
            opt.step()
        for i, p in enumerate(ps):
            out["%s/p%d" % (name, i)] = p.detach().numpy()
            out["%s/m%d" % (name, i)] = opt.state[p]["next_m"].numpy()
            out["%s/v%d" % (name, i)] = opt.state[p]["next_v"].numpy()
    # SGD with momentum as built for the proxies
    params, grads = opt_inputs([(80, 64)], 3, seed=9)
    p = torch.nn.Parameter(params[0].clone())
    sgd = torch.optim.SGD([p], lr=0.02, momentum=0.9, weight_decay=0.0005)
    for s in range(3):
        p.grad = grads[s][0].clone()
        sgd.step()
    out["sgd/p"] = p.detach().numpy()
    # HyP gradients (autograd of the reference module)
    for i, (name, B, K, C, thr, alpha, dens) in enumerate(CASES):
        x, y, label, proxies = inputs(B, K, C, dens, 100 + i)
        mod = ref_hyp.HyP(numclass=C, output_dim=K, hypseed=0, alpha=alpha, threshold=thr)
        with torch.no_grad():
            mod.proxies.copy_(proxies)
        x.requires_grad_(), y.requires_grad_()
        mod(x, y, label).backward()
        out["hyp/%s/dx" % name], out["hyp/%s/dy" % name] = x.grad.numpy(), y.grad.numpy()
        out["hyp/%s/dp" % name] = mod.proxies.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "optimizer_golden.npz"), **out)
    print(len(out), "arrays")


if __name__ == "__main__":
    main()
