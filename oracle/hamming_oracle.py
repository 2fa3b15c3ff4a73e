"""TEST INFRASTRUCTURE ONLY — integer (numpy) oracle for the bit-packed retrieval stages.

An *independent* formulation of what ``common/calc_utils.py:51-92`` computes, on bit-packed codes:
Hamming distance as XOR + popcount, the stable (distance, gallery-index) ranking obtained by
counting instead of sorting, the 1-based ranks ("tindex") of the first ``min(R, k)`` relevant items,
per-query top-k, and the shard/merge arithmetic of SURVEY.md §8(e).  It shares no code with
``calc_utils_port.py``; tests require both to agree with each other and with the golden vectors.

Bit layout (the one the CUDA kernels use; DESIGN.md "data layout"):
  code word w, bit b  <->  column 32*w + b of the +-1 matrix, bit = 1 iff value > 0
  (reference codes are +-1 floats: runners/base.py:407-410 ``sign_()``, DCMHT argmax
  runners/DCMHT/runner.py:83-95); label word w, bit b <-> class 32*w + b, bit = 1 iff label != 0.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

LABEL_WORDS = 4  # 128 classes max, fixed-width label mask per item (4 x u32)


def pack_codes(codes: np.ndarray) -> np.ndarray:
    """[n, K] +-1 (any float/int dtype) -> [n, ceil(K/32)] uint32, little-endian bit order."""
    codes = np.asarray(codes)
    n, nbits = codes.shape
    words = (nbits + 31) // 32
    bits = np.zeros((n, words * 32), dtype=np.uint8)
    bits[:, :nbits] = codes > 0
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64)).astype(np.uint64)
    out = (bits.reshape(n, words, 32).astype(np.uint64) * weights).sum(axis=2)
    return out.astype(np.uint32)


def pack_labels(labels: np.ndarray) -> np.ndarray:
    """[n, C] multi-hot (0/1) -> [n, 4] uint32; C <= 128."""
    labels = np.asarray(labels)
    n, ncls = labels.shape
    if ncls > 32 * LABEL_WORDS:
        raise ValueError("at most %d classes" % (32 * LABEL_WORDS))
    bits = np.zeros((n, 32 * LABEL_WORDS), dtype=np.uint8)
    bits[:, :ncls] = labels != 0
    weights = (np.uint64(1) << np.arange(32, dtype=np.uint64)).astype(np.uint64)
    out = (bits.reshape(n, LABEL_WORDS, 32).astype(np.uint64) * weights).sum(axis=2)
    return out.astype(np.uint32)


def hamming_matrix(qp: np.ndarray, gp: np.ndarray) -> np.ndarray:
    """popcount(q XOR g) summed over words -> [Q, N] uint16 (== calc_hammingDist on +-1 inputs)."""
    x = qp[:, None, :] ^ gp[None, :, :]
    return np.bitwise_count(x).sum(axis=2, dtype=np.uint16)


def relevance_matrix(qlp: np.ndarray, glp: np.ndarray) -> np.ndarray:
    """[Q, N] bool: share at least one class  (== ``query_L.mm(retrieval_L.T) > 0``, calc_utils.py:72)."""
    return ((qlp[:, None, :] & glp[None, :, :]) != 0).any(axis=2)


def stable_ranks(dist_row: np.ndarray, nbins: int) -> np.ndarray:
    """0-based rank of every gallery item under the (distance, index) order, by counting.

    rank(j) = #{d' < d_j} + #{j' < j : d_j' == d_j}   (SURVEY.md §7 "counting formulation").
    """
    hist = np.bincount(dist_row, minlength=nbins)
    below = np.concatenate(([0], np.cumsum(hist)[:-1]))
    ranks = np.empty(dist_row.shape[0], dtype=np.int64)
    for d in np.nonzero(hist)[0]:
        members = np.nonzero(dist_row == d)[0]  # ascending gallery index
        ranks[members] = below[d] + np.arange(members.shape[0])
    return ranks


def map_parts(
    qp: np.ndarray,
    gp: np.ndarray,
    qlp: np.ndarray,
    glp: np.ndarray,
    nbits: int,
    k: Optional[int] = None,
) -> Tuple[List[np.ndarray], np.ndarray, np.ndarray]:
    """Integer stage of calc_map_k.

    Returns ``(tindex_list, totals [Q] int64, tsums [Q] int64)`` with ``tindex_list[i]`` the ascending
    1-based ranks (int64) of the first ``totals[i] = min(R_i, k)`` relevant gallery items of query i.
    """
    n = gp.shape[0]
    if k is None:
        k = n
    tindex_list: List[np.ndarray] = []
    totals = np.zeros(qp.shape[0], dtype=np.int64)
    tsums = np.zeros(qp.shape[0], dtype=np.int64)
    for i in range(qp.shape[0]):
        dist = hamming_matrix(qp[i : i + 1], gp)[0]
        rel = relevance_matrix(qlp[i : i + 1], glp)[0]
        ranks = stable_ranks(dist, nbits + 1)
        rel_ranks = np.sort(ranks[rel]) + 1
        total = min(int(rel.sum()), int(k))
        tsums[i] = int(rel.sum())
        totals[i] = total
        tindex_list.append(rel_ranks[:total].astype(np.int64))
    return tindex_list, totals, tsums


def topk(qp: np.ndarray, gp: np.ndarray, nbits: int, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Per-query k smallest (distance, index) pairs -> (dist [Q,k] int32, idx [Q,k] int64).

    Slots beyond the gallery size (k > N) are filled with dist = -1, idx = -1.
    """
    nq, n = qp.shape[0], gp.shape[0]
    out_d = np.full((nq, k), -1, dtype=np.int32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    for i in range(nq):
        dist = hamming_matrix(qp[i : i + 1], gp)[0].astype(np.int64)
        key = dist * (n + 1) + np.arange(n)  # total order (dist, idx)
        kk = min(k, n)
        sel = np.sort(np.partition(key, kk - 1)[:kk]) if kk < n else np.sort(key)
        out_d[i, :kk] = sel // (n + 1)
        out_i[i, :kk] = sel % (n + 1)
    return out_d, out_i


def ap_terms_float32(tindex: np.ndarray) -> np.ndarray:
    """fp32 ``count / tindex`` exactly as calc_utils.py:87-89 forms them (IEEE fp32 divide)."""
    count = np.arange(1, tindex.shape[0] + 1, dtype=np.float32)
    return count / tindex.astype(np.float32)


def map_float64(tindex_list: List[np.ndarray]) -> float:
    """Exactly-rounded reference value: every AP is the fp64 sum of the *fp32* terms / total.

    This is what the CUDA "device" mode computes: same fp32 terms as the reference, summed
    without the reference's fp32 rounding noise.  nan if any query has no relevant item, like the
    reference (mean of an empty tensor).
    """
    acc = 0.0
    for t in tindex_list:
        if t.shape[0] == 0:
            return float("nan")
        acc += float(ap_terms_float32(t).astype(np.float64).sum()) / t.shape[0]
    return acc / len(tindex_list)


# --------------------------------------------------------------------------------------------------
# sharded evaluation: contiguous gallery shards, per-shard partials, exact merge
# --------------------------------------------------------------------------------------------------
def shard_bounds(n: int, world: int, align: int = 1) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) gallery ranges, sizes differing by at most ``align`` items."""
    per = -(-n // world)
    per = -(-per // align) * align
    return [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)]


def merged_ranks_from_shards(
    dist_rows: List[np.ndarray], nbins: int
) -> List[np.ndarray]:
    """Global 0-based stable ranks recomputed from per-shard histograms + local in-bucket positions.

    global_rank(j in shard s) = sum_{d'<d} H[d'] + sum_{s'<s} h_{s'}[d] + local_pos_s(j)
    with h_s the shard histogram and H their sum.  Used to validate the merge kernel's arithmetic.
    """
    hists = [np.bincount(d, minlength=nbins) for d in dist_rows]
    total = np.sum(hists, axis=0)
    below = np.concatenate(([0], np.cumsum(total)[:-1]))
    out = []
    carried = np.zeros(nbins, dtype=np.int64)
    for s, d in enumerate(dist_rows):
        local = stable_ranks(d, nbins)
        local_below = np.concatenate(([0], np.cumsum(hists[s])[:-1]))
        local_pos = local - local_below[d]
        out.append(below[d] + carried[d] + local_pos)
        carried = carried + hists[s]
    return out
