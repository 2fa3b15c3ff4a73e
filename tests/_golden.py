"""Loader for tests/golden/retrieval_golden.npz (written by tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "retrieval_golden.npz")
_npz = None

CASE_NAMES = ["tiny16", "tiny16_full", "mid64", "mid64_full", "wide128", "odd32", "odd48",
              "clustered64", "sparse_rel", "C1"]


def npz():
    global _npz
    if _npz is None:
        _npz = np.load(_PATH)
    return _npz


class Case:
    def __init__(self, name):
        z = npz()
        p = name + "/"
        q, n, nbits, ncls, k = (int(v) for v in z[p + "shape"])
        self.name, self.Q, self.N, self.K, self.C = name, q, n, nbits, ncls
        self.k = None if k < 0 else k
        unb = lambda a, w: np.unpackbits(a, axis=1)[:, :w]
        self.qB = torch.from_numpy(unb(z[p + "q_bits"], nbits).astype(np.float32) * 2 - 1)
        self.rB = torch.from_numpy(unb(z[p + "r_bits"], nbits).astype(np.float32) * 2 - 1)
        self.qL = torch.from_numpy(unb(z[p + "q_lab"], ncls).astype(np.int64))
        self.rL = torch.from_numpy(unb(z[p + "r_lab"], ncls).astype(np.int64))
        self.map_stable = z[p + "map_stable"]
        self.map_shipped = z[p + "map_shipped"]
        self.totals = z[p + "totals"]
        self.tsums = z[p + "tsums"]
        self.order_head = z[p + "order_head"]
        self.hamm = z[p + "hamm"] if (p + "hamm") in z.files else None
        self.hamm_rowsum = z[p + "hamm_rowsum"]
        offs = z[p + "tindex_offs"]
        flat = z[p + "tindex_flat"]
        self.tindex = [flat[offs[i]:offs[i + 1]] for i in range(q)]
