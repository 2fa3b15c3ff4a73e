#!/usr/bin/env python
"""Generate the committed golden vectors by EXECUTING THE REFERENCE in this container.

Run (build container only — /root/reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Imports ``/root/reference/common/calc_utils.py`` unmodified and records, for a handful of small seeded
cases, the inputs (bit-packed with numpy so they do not depend on any RNG implementation) and the
reference outputs:

  hamm          calc_hammingDist(qB, rB)                                    (calc_utils.py:51-56)
  map_stable    calc_map_k(...) with torch.sort patched to stable=True       (calc_utils.py:58-92)
  map_shipped   calc_map_k(...) exactly as shipped (unstable sort)           — reported, not a target
  order_head    first 64 columns of torch.sort(hamms, stable=True) indices   (calc_utils.py:77)
  tindex/totals the 1-based ranks the loop at calc_utils.py:84-89 forms (recomputed here from the
                reference's own hamms/gnds, since the reference does not return them)
  label_sim, cosine, euclid, weight_sim  — a4..a7 on small float inputs

The reference has no tests or fixtures of its own; these files are the pin.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from common import calc_utils as ref  # noqa: E402  (the unmodified reference)

from clip_based_cross_modal_hash_b200 import synth  # noqa: E402

CASES = [
    # name,          Q,    N,  K,  C,   k,   kind
    ("tiny16",      37,  301, 16, 24,   50, "random"),
    ("tiny16_full", 37,  301, 16, 24, None, "random"),
    ("mid64",       16, 1000, 64, 80,  100, "random"),
    ("mid64_full",  16, 1000, 64, 80, None, "random"),
    ("wide128",      8, 2000, 128, 21, 5000, "random"),
    ("odd32",        5,  777, 32, 24,    7, "random"),
    ("odd48",        9,  513, 48, 33,   20, "random"),
    ("clustered64", 12, 1500, 64, 80,  200, "clustered"),
    ("sparse_rel",  10,  400, 32, 100,  50, "sparse"),
    ("C1",        1000, 5000, 16, 24,   50, "random"),
]


class _StableSort:
    """Context manager: torch.sort -> stable=True inside the reference (the one-keyword canonicalisation)."""

    def __enter__(self):
        self._orig = torch.sort

        def stable_sort(*a, **kw):
            kw["stable"] = True
            return self._orig(*a, **kw)

        torch.sort = stable_sort
        return self

    def __exit__(self, *exc):
        torch.sort = self._orig


def make_inputs(q, n, nbits, ncls, kind, seed):
    if kind == "clustered":
        allc = synth.clustered_codes(q + n, nbits, seed)
        qB, rB = allc[:q].clone(), allc[q:].clone()
    else:
        qB = synth.random_codes(q, nbits, seed)
        rB = synth.random_codes(n, nbits, seed + 1)
    p = 0.004 if kind == "sparse" else 0.07
    qL = synth.random_labels(q, ncls, seed + 2, p=p)
    rL = synth.random_labels(n, ncls, seed + 3, p=p)
    return qB, rB, qL, rL


def tindex_from_reference(qB, rB, qL, rL, k):
    """Recompute lines 72-88 with the reference's own helpers to expose the integer intermediates."""
    if k is None:
        k = rL.shape[0]
    gnds = (qL.mm(rL.t()) > 0).float()
    hamms = ref.calc_hammingDist(qB, rB)
    order = torch.sort(hamms, dim=-1, stable=True)[1]
    tsums = gnds.sum(dim=-1).to(torch.int64)
    flat, offs, totals = [], [0], []
    for i in range(qB.shape[0]):
        total = int(min(int(tsums[i]), k))
        t = (torch.nonzero(gnds[i][order[i]])[:total].reshape(-1) + 1).to(torch.int32)
        flat.append(t.numpy())
        offs.append(offs[-1] + t.numel())
        totals.append(total)
    return (np.concatenate(flat) if flat else np.zeros(0, np.int32), np.asarray(offs, np.int64),
            np.asarray(totals, np.int32), tsums.numpy().astype(np.int32), order[:, :64].numpy().astype(np.int32),
            hamms)


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    out = {}
    for ci, (name, q, n, nbits, ncls, k, kind) in enumerate(CASES):
        qB, rB, qL, rL = make_inputs(q, n, nbits, ncls, kind, seed=1000 + 10 * ci)
        map_shipped = ref.calc_map_k(qB, rB, qL, rL, k)
        with _StableSort():
            map_stable = ref.calc_map_k(qB, rB, qL, rL, k)
        tflat, toffs, totals, tsums, order_head, hamms = tindex_from_reference(qB, rB, qL, rL, k)
        pre = name + "/"
        out[pre + "shape"] = np.asarray([q, n, nbits, ncls, -1 if k is None else k], np.int64)
        out[pre + "q_bits"] = np.packbits((qB > 0).numpy(), axis=1)
        out[pre + "r_bits"] = np.packbits((rB > 0).numpy(), axis=1)
        out[pre + "q_lab"] = np.packbits(qL.numpy().astype(bool), axis=1)
        out[pre + "r_lab"] = np.packbits(rL.numpy().astype(bool), axis=1)
        out[pre + "map_stable"] = np.asarray(map_stable.item(), np.float32)
        out[pre + "map_shipped"] = np.asarray(map_shipped.item(), np.float32)
        out[pre + "tindex_flat"] = tflat
        out[pre + "tindex_offs"] = toffs
        out[pre + "totals"] = totals
        out[pre + "tsums"] = tsums
        out[pre + "order_head"] = order_head
        if q * n <= 40_000:
            assert torch.equal(hamms, hamms.round())
            out[pre + "hamm"] = hamms.numpy().astype(np.uint8)
        out[pre + "hamm_rowsum"] = hamms.sum(dim=1).numpy().astype(np.int64)
        print(f"{name:12s} Q={q} N={n} K={nbits} C={ncls} k={k}: map_stable={map_stable.item():.9f} "
              f"map_shipped={map_shipped.item():.9f} min_total={totals.min()}")

    # a4..a7 similarity helpers on small float inputs
    g = torch.Generator().manual_seed(77)
    a = torch.randn(19, 64, generator=g)
    b = torch.randn(23, 64, generator=g)
    la = (torch.rand(19, 24, generator=g) < 0.15).float()
    lb = (torch.rand(23, 24, generator=g) < 0.15).float()
    out["sim/a"] = a.numpy()
    out["sim/b"] = b.numpy()
    out["sim/la"] = la.numpy()
    out["sim/lb"] = lb.numpy()
    out["sim/label_sim"] = ref.calc_label_sim(la, lb).numpy()
    out["sim/label_sim_i64"] = ref.calc_label_sim(la.long(), lb.long()).numpy()
    out["sim/cosine"] = ref.cosine_similarity(a, b).numpy()
    out["sim/cosine_np"] = ref.cosine_similarity(a.numpy(), b.numpy())
    out["sim/euclid"] = ref.euclidean_similarity(a, b).numpy()
    out["sim/euclid_np"] = ref.euclidean_similarity(a.numpy(), b.numpy())
    ls, ws = ref.generate_weight_sim(la, la)
    out["sim/weight_label"] = ls.numpy()
    out["sim/weight_sim"] = ws.numpy()
    # calc_hammingDist on a 1-D query and on codes containing 0 (sign_() of an exact zero)
    c = synth.random_codes(6, 32, 5)
    c[0, 3] = 0.0
    c[2, 7] = 0.0
    r = synth.random_codes(50, 32, 6)
    out["hd/q"] = c.numpy()
    out["hd/r"] = r.numpy()
    out["hd/full"] = ref.calc_hammingDist(c, r).numpy()
    out["hd/one_d"] = ref.calc_hammingDist(c[1], r).numpy()

    path = os.path.join(HERE, "retrieval_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
