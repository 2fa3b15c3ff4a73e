"""TEST INFRASTRUCTURE ONLY — torch-CPU restatement ("port") of the reference evaluator.

Restates ``/root/reference/common/calc_utils.py`` with the same ATen CPU operators the reference
calls (int64 ``mm`` for label similarity, fp32 ``mm`` for Hamming, ``torch.sort``, fp32
``mean``/accumulate), so that (1) its results are the parity target for the CUDA path and (2) its
wall-clock on the GPU box's host cores is a fair stand-in for "the reference's own CPU path"
(``bench.py``: ``cpu_baseline.kind == "port"``; the Python reference itself cannot travel to the GPU
box).  Differences from the reference, all deliberate and result-neutral:

* ``stable`` keyword on the ranking sort.  The reference calls ``torch.sort`` without ``stable=True``
  (calc_utils.py:77); with K-bit codes there are only K+1 distinct distances so ties are the norm and
  the unstable order is unspecified.  ``stable=True`` (ascending gallery index among
  equal distances) is the canonical order every parity claim in this repo refers to;
  ``stable=False`` reproduces the as-shipped call.
* the query axis may be processed in chunks (``query_chunk``) so 10k x 1M fits in host memory; rows
  are independent and the cross-query accumulation stays sequential fp32, so results are unchanged.
* ``return_parts`` additionally hands back the integer intermediates (ranks of the relevant items).

Pinned against the reference by tests/test_oracle_golden.py (fixtures from tests/golden/make_golden.py).
"""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np
import torch

ArrayLike = Union[torch.Tensor, np.ndarray]


# --------------------------------------------------------------------------------------------------
# a4  calc_label_sim            (reference: common/calc_utils.py:8-10)
# --------------------------------------------------------------------------------------------------
def calc_label_sim(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """1.0 where two multi-hot label rows share a class: ``(a @ b.T > 0).float()``."""
    gram = torch.matmul(a, b.t())
    return gram.gt(0).to(torch.float32)


# --------------------------------------------------------------------------------------------------
# a7  generate_weight_sim       (reference: common/calc_utils.py:12-26)
# --------------------------------------------------------------------------------------------------
def generate_weight_sim(a: torch.Tensor, b: torch.Tensor):
    """Binary label similarity plus the NDCG-normalised graded similarity.

    graded = (2**s - 1) / Z_row with s = a @ b.T and
    Z_row = sum_j (2**sorted_desc(s)[row, j] - 1) / log2(j + 2)     (calc_utils.py:17-24).
    The reference builds the log2 table from ``a.shape[0]`` (so it needs a square gram matrix).
    """
    gram = torch.matmul(a, b.t())
    n = a.shape[0]
    binary = gram.gt(0).to(torch.float32)
    ideal, _ = torch.sort(gram, dim=1, descending=True)
    discount = torch.log2(torch.arange(0.0, n) + 2).repeat(1, n).reshape(n, n).to(a.device)
    z = ((2 ** ideal - 1) / discount).sum(dim=1).reshape(-1, 1)
    graded = (2 ** gram - 1) / z
    return binary, graded


# --------------------------------------------------------------------------------------------------
# a6  euclidean_similarity      (reference: common/calc_utils.py:28-36)
# --------------------------------------------------------------------------------------------------
def euclidean_similarity(a: ArrayLike, b: ArrayLike) -> ArrayLike:
    """Pairwise L2 distance; torch -> ``torch.cdist``; numpy -> sklearn ``euclidean_distances``."""
    if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
        return torch.cdist(a, b, p=2.0)
    if isinstance(a, np.ndarray) and isinstance(b, np.ndarray):
        from sklearn.metrics.pairwise import euclidean_distances  # same call as calc_utils.py:33

        return euclidean_distances(a, b)
    raise ValueError(
        "input value must in [torch.Tensor, numpy.ndarray], but it is %s, %s" % (type(a), type(b))
    )


# --------------------------------------------------------------------------------------------------
# a5  cosine_similarity         (reference: common/calc_utils.py:38-49)
# --------------------------------------------------------------------------------------------------
def cosine_similarity(a: ArrayLike, b: ArrayLike) -> ArrayLike:
    """Row-normalise both operands (no epsilon: a zero row gives nan) and multiply."""
    if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
        an = a / a.norm(dim=-1, keepdim=True)
        bn = b / b.norm(dim=-1, keepdim=True)
        return torch.matmul(an, bn.t())
    if isinstance(a, np.ndarray) and isinstance(b, np.ndarray):
        an = a / np.linalg.norm(a, axis=-1, keepdims=True)
        bn = b / np.linalg.norm(b, axis=-1, keepdims=True)
        return np.matmul(an, bn.T)
    raise ValueError(
        "input value must in [torch.Tensor, numpy.ndarray], but it is %s, %s" % (type(a), type(b))
    )


# --------------------------------------------------------------------------------------------------
# a1  calc_hammingDist          (reference: common/calc_utils.py:51-56)
# --------------------------------------------------------------------------------------------------
def calc_hammingDist(B1: torch.Tensor, B2: torch.Tensor) -> torch.Tensor:
    """``0.5 * (K - B1 @ B2.T)`` in fp32; a 1-D ``B1`` is treated as a single query."""
    nbits = B2.shape[1]
    if B1.dim() < 2:
        B1 = B1.unsqueeze(0)
    return 0.5 * (nbits - B1.mm(B2.t()))


# --------------------------------------------------------------------------------------------------
# a2  calc_map_k                (reference: common/calc_utils.py:58-92)
# --------------------------------------------------------------------------------------------------
def calc_map_k(
    qB: torch.Tensor,
    rB: torch.Tensor,
    query_L: torch.Tensor,
    retrieval_L: torch.Tensor,
    k: Optional[int] = None,
    *,
    stable: bool = True,
    query_chunk: Optional[int] = None,
    return_parts: bool = False,
):
    """mAP over the first ``min(R, k)`` relevant items of the full Hamming ranking.

    Follows calc_utils.py:58-92 statement by statement: label gram (int64 mm, :72), row sums (:75),
    Hamming (:76), full sort (:77), ``totals = min(tsums, k)`` (:81), then per query gather the
    relevance along the ranking, take the 1-based ranks of the first ``total`` relevant items and add
    ``mean(arange(1..total) / ranks)`` to a running fp32 sum (:84-89); finally divide by Q (:90).

    ``return_parts`` -> ``(map, tindex_list, totals)`` where ``tindex_list[i]`` is the int64 vector of
    1-based ranks for query i (the integer stage that must be bit-exact on the GPU).
    """
    num_query = query_L.shape[0]
    qB = qB.detach().cpu()
    rB = rB.detach().cpu()
    query_L = query_L.detach().cpu()
    retrieval_L = retrieval_L.detach().cpu()
    if k is None:
        k = retrieval_L.shape[0]
    step = num_query if not query_chunk else int(query_chunk)

    running = 0  # becomes a 0-dim fp32 tensor after the first add, exactly like the reference's `map`
    tindex_list: List[torch.Tensor] = []
    totals_all: List[int] = []
    for lo in range(0, num_query, step):
        hi = min(lo + step, num_query)
        gnds = (query_L[lo:hi].mm(retrieval_L.t()) > 0).to(torch.float32)  # [q, N]; no .squeeze()
        tsums = gnds.sum(dim=-1, keepdim=True, dtype=torch.int32)
        hamms = calc_hammingDist(qB[lo:hi], rB)
        order = torch.sort(hamms, dim=-1, stable=stable)[1]
        totals = torch.min(tsums, torch.tensor([k], dtype=torch.int32).expand_as(tsums))
        for row in range(hi - lo):
            ranked_rel = gnds[row][order[row]]
            total = totals[row].squeeze()
            count = torch.arange(1, total + 1).to(torch.float32)
            tindex = torch.nonzero(ranked_rel)[:total].squeeze().to(torch.float32) + 1.0
            running = running + torch.mean(count / tindex)
            if return_parts:
                tindex_list.append((torch.nonzero(ranked_rel)[:total].reshape(-1) + 1).to(torch.int64))
                totals_all.append(int(total))
    result = running / num_query
    if return_parts:
        return result, tindex_list, totals_all
    return result


def hamming_rank_topk(qB: torch.Tensor, rB: torch.Tensor, k: int, *, stable: bool = True):
    """First ``k`` entries of the ranking calc_map_k builds at calc_utils.py:76-77.

    Returns ``(dist [Q,k] fp32, index [Q,k] int64)`` — the "per-query top-k" of north_star.
    """
    hamms = calc_hammingDist(qB.detach().cpu(), rB.detach().cpu())
    vals, idx = torch.sort(hamms, dim=-1, stable=stable)
    return vals[:, :k].contiguous(), idx[:, :k].contiguous()
