#!/usr/bin/env python
"""tests/golden/hyp_golden.npz: HyP.forward of the reference (models/DSPH/loss/HyP.py) on seeded inputs (build container only)."""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_hyp", "/root/reference/models/DSPH/loss/HyP.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests._hyp_cases import CASES, inputs  # noqa: E402


def main():
    out = {}
    for i, (name, B, K, C, thr, alpha, dens) in enumerate(CASES):
        x, y, label, proxies = inputs(B, K, C, dens, 100 + i)
        mod = ref.HyP(numclass=C, output_dim=K, hypseed=0, alpha=alpha, threshold=thr)
        with torch.no_grad():
            mod.proxies.copy_(proxies)
            out[name] = np.float32(mod(x, y, label).item())
    np.savez(os.path.join(HERE, "hyp_golden.npz"), **out)
    print({k: float(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
