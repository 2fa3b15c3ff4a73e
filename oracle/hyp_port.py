"""ORACLE (test infrastructure, never on the product path): fp32 CPU restatement of ``HyP.forward``
(models/DSPH/loss/HyP.py:18-69).  Pinned by tests/golden/hyp_golden.npz (the reference module executed in the build container)."""
import torch
import torch.nn.functional as F


def hyp_loss(x, y, label, proxies, threshold: float, alpha: float = 0.8) -> torch.Tensor:
    x, y, proxies = x.float(), y.float(), proxies.float()
    on = label != 0
    pn = F.normalize(proxies, dim=1)
    total = x.new_zeros(())
    for feat in (x, y):                                                   # HyP.py:23-39: image branch, then text branch
        cos = F.normalize(feat, dim=1) @ pn.t()
        total = total + (1 - cos)[on].sum() / on.sum() + torch.relu(cos - threshold)[~on].sum() / (~on).sum()
    if alpha > 0:                                                         # HyP.py:41-64
        multi = label.sum(dim=1) > 1
        lab = label[multi].float()
        disjoint = (lab @ lab.t()) == 0
        if disjoint.any():
            xn, yn = F.normalize(x[multi], dim=1), F.normalize(y[multi], dim=1)
            for sim in (xn @ xn.t(), yn @ yn.t(), xn @ yn.t()):
                total = total + (alpha * torch.relu(sim - threshold))[disjoint].sum() / disjoint.sum()
    return total
