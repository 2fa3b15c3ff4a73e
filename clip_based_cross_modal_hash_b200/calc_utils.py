"""Drop-in for the reference's ``common/calc_utils.py`` — same names, positional signatures, return
conventions and error behaviour; the arithmetic runs in the sm_100a kernels of ``libcmh.so``.

    reference (common/calc_utils.py)            here
    calc_label_sim(a, b)            :8-10       cmh_label_sim_f32
    generate_weight_sim(a, b)       :12-26      label gram by cmh kernel + the reference's own torch ops
    euclidean_similarity(a, b)      :28-36      cmh_euclid_sim_f32
    cosine_similarity(a, b)         :38-49      cmh_cosine_sim_f32
    calc_hammingDist(B1, B2)        :51-56      cmh_pack_codes_f32 + cmh_hamming_f32 (XOR + popcount)
    calc_map_k(qB, rB, qL, rL, k)   :58-92      cmh_pack_* + cmh_map_k (two counting passes, no sort)

Install into a reference checkout with ``install_into_reference()`` (monkey-patches ``common.calc_utils``
so ``runners/base.py:78`` binds this ``calc_map_k``) — see INTEGRATION.md.

There is NO CPU fallback: without a CUDA device (or without libcmh.so) every function raises ``CmhError``.

calc_map_k modes (``CMH_MAP_MODE`` env var or the ``mode=`` keyword):
  "device" (default)  integer ranks and the fp32 quotients ``count / tindex`` are formed on the GPU exactly as
                      the reference forms them; their sum is accumulated in fp64 on the GPU and rounded to fp32
                      once.  Differs from the reference's fp32 running sums by at most a few fp32 ulps.
  "parity"            the GPU hands back the integer ranks (``tindex``); ``mean(count / tindex)`` and the
                      running sum are then evaluated with the very same torch CPU ops as calc_utils.py:84-90,
                      so the returned fp32 value is bit-identical to the reference's (with the stable tie order).
"""
from __future__ import annotations

import os
from typing import Optional, Union

import numpy as np
import torch

from . import retrieval as R
from ._lib import CmhError

ArrayLike = Union[torch.Tensor, np.ndarray]

# rows of tindex fetched per parity-mode slab (bounds device + pinned host memory)
_PARITY_SLAB_BYTES = 1 << 30


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise CmhError("clip_based_cross_modal_hash_b200 needs a CUDA device: there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    return t if t.is_cuda else t.to(dev, non_blocking=True)


def _as_f32_dev(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    t = _to_dev(t.detach(), dev)
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


def _sim(kind: str, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    from . import _lib
    import ctypes  # noqa: F401

    dev = a.device if a.is_cuda else (b.device if b.is_cuda else _device())
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != b.shape[1]:
        raise RuntimeError("expected [n, d] and [m, d] operands, got %s and %s" % (tuple(a.shape), tuple(b.shape)))
    ad, bd = _as_f32_dev(a, dev), _as_f32_dev(b, dev)
    out = torch.empty((ad.shape[0], bd.shape[0]), dtype=torch.float32, device=dev)
    fn = getattr(_lib.lib(), "cmh_%s_sim_f32" % kind)
    with torch.cuda.device(dev):
        _lib.check(fn(ad.data_ptr(), ad.shape[0], bd.data_ptr(), bd.shape[0], ad.shape[1], out.data_ptr(),
                      torch.cuda.current_stream().cuda_stream))
    return out


def _back(out: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    return out if like.is_cuda else out.cpu()


# a4 -------------------------------------------------------------------------------------------------------
def calc_label_sim(a: torch.Tensor, b: torch.Tensor):
    """``(a.matmul(b.T) > 0).float()`` (common/calc_utils.py:8-10); result on ``a``'s device."""
    return _back(_sim("label", a, b), a)


# a7 -------------------------------------------------------------------------------------------------------
def generate_weight_sim(a: torch.Tensor, b: torch.Tensor):
    """common/calc_utils.py:12-26 (no caller in the reference; kept for API completeness).  The label gram
    and its threshold come from the cmh kernel; the NDCG normaliser keeps the reference's torch expressions."""
    dev = a.device if a.is_cuda else _device()
    ad, bd = _as_f32_dev(a, dev), _as_f32_dev(b, dev)
    label_sim = _sim("label", ad, bd)
    sim_origin = ad.matmul(bd.transpose(0, 1))
    batch_size = a.shape[0]
    ideal_list = torch.sort(sim_origin, dim=1, descending=True)[0]
    ph = torch.arange(0., batch_size) + 2
    ph = ph.repeat(1, batch_size).reshape(batch_size, batch_size)
    th = torch.log2(ph).to(dev)
    Z = (((2 ** ideal_list - 1) / th).sum(axis=1)).reshape(-1, 1)
    sim_origin = (2 ** sim_origin - 1) / Z
    return _back(label_sim, a), _back(sim_origin, a)


# a6 -------------------------------------------------------------------------------------------------------
def euclidean_similarity(a: ArrayLike, b: ArrayLike):
    """Pairwise L2 distance (common/calc_utils.py:28-36): torch in -> torch out, numpy in -> numpy out."""
    if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
        return _back(_sim("euclid", a, b), a)
    elif isinstance(a, np.ndarray) and isinstance(b, np.ndarray):
        out = _sim("euclid", torch.from_numpy(np.ascontiguousarray(a)), torch.from_numpy(np.ascontiguousarray(b)))
        return out.cpu().numpy().astype(np.result_type(a.dtype, b.dtype, np.float32), copy=False)
    else:
        raise ValueError("input value must in [torch.Tensor, numpy.ndarray], but it is %s, %s" % (type(a), type(b)))


# a5 -------------------------------------------------------------------------------------------------------
def cosine_similarity(a: ArrayLike, b: ArrayLike):
    """Row-normalised gram (common/calc_utils.py:38-49); no epsilon: a zero row gives nan like the reference."""
    if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
        return _back(_sim("cosine", a, b), a)
    elif isinstance(a, np.ndarray) and isinstance(b, np.ndarray):
        out = _sim("cosine", torch.from_numpy(np.ascontiguousarray(a)), torch.from_numpy(np.ascontiguousarray(b)))
        return out.cpu().numpy().astype(np.result_type(a.dtype, b.dtype, np.float32), copy=False)
    else:
        raise ValueError("input value must in [torch.Tensor, numpy.ndarray], but it is %s, %s" % (type(a), type(b)))


# a1 -------------------------------------------------------------------------------------------------------
def calc_hammingDist(B1: torch.Tensor, B2: torch.Tensor) -> torch.Tensor:
    """``0.5 * (K - B1 @ B2.T)`` (common/calc_utils.py:51-56) as fp32 [Q, N] on the inputs' device.

    +-1 inputs take the bit-packed XOR+popcount kernel; anything else (e.g. codes containing the 0 that
    ``sign_()`` gives for an exact zero) takes the dense fp32 kernel so results still match the reference."""
    if len(B1.shape) < 2:
        B1 = B1.unsqueeze(0)
    dev = B1.device if B1.is_cuda else (B2.device if B2.is_cuda else _device())
    b1, b2 = _as_f32_dev(B1, dev), _as_f32_dev(B2, dev)
    nbits = b2.shape[1]
    if nbits <= 128:
        bad = R.new_bad_counter(dev)
        qp, gp = R.pack_codes(b1, bad), R.pack_codes(b2, bad)
        out = R.hamming_matrix(qp, gp, nbits)
        if int(bad.item()) == 0:
            return _back(out, B1)
    return _back(R.hamming_dense(b1, b2), B1)


# a2 -------------------------------------------------------------------------------------------------------
def _parity_reduce(tindex_rows: torch.Tensor, totals: torch.Tensor, running):
    """calc_utils.py:84-89 on the host for a slab of queries: same torch CPU ops, same order."""
    for row in range(tindex_rows.shape[0]):
        total = totals[row]
        count = torch.arange(1, total + 1).type(torch.float32)
        tindex = (tindex_rows[row, :total] - 1).type(torch.float32) + 1.0
        running = running + torch.mean(count / tindex)
    return running


def calc_map_k(qB, rB, query_L, retrieval_L, k=None, *, mode: Optional[str] = None):
    """mAP over the first ``min(R, k)`` relevant items of the full Hamming ranking
    (common/calc_utils.py:58-92).  Accepts tensors on any device; returns a 0-dim fp32 CPU tensor.

    Ties in distance are ranked by ascending gallery index (``torch.sort(..., stable=True)``), the
    canonical order of this repo; the reference's unstable CPU sort is unspecified there.
    Raises ``ValueError`` if the codes are not +-1 or the labels not 0/1."""
    mode = mode or os.environ.get("CMH_MAP_MODE", "device")
    if mode not in ("device", "parity"):
        raise ValueError("mode must be 'device' or 'parity'")
    dev = next((t.device for t in (qB, rB, query_L, retrieval_L) if isinstance(t, torch.Tensor) and t.is_cuda), None)
    dev = dev or _device()
    num_query = query_L.shape[0]
    nbits, ncls = rB.shape[1], retrieval_L.shape[1]
    if nbits > 128 or ncls > 128:
        raise CmhError("calc_map_k supports up to 128 bits and 128 classes (got %d, %d)" % (nbits, ncls))
    n = retrieval_L.shape[0]
    if k is None:
        k = n
    with torch.cuda.device(dev):
        bad = R.new_bad_counter(dev)
        qp = R.pack_codes(_to_dev(qB.detach(), dev), bad)
        gp = R.pack_codes(_to_dev(rB.detach(), dev), bad)
        qlp = R.pack_labels(_to_dev(query_L.detach(), dev), bad)
        glp = R.pack_labels(_to_dev(retrieval_L.detach(), dev), bad)
        if mode == "device":
            res = R.map_k(qp, qlp, gp, glp, nbits, ncls, k)
            host = torch.stack([res.map, bad[0].to(torch.float64)]).cpu()
            if host[1] != 0:
                raise ValueError("calc_map_k: codes must be +-1 and labels 0/1 (%d offending elements)" % int(host[1]))
            return host[0].to(torch.float32)

        # parity mode: integer ranks from the GPU, fp32 reduction by the reference's own CPU ops
        st = R.CudaStages()
        plan = st.make_plan(num_query, n, nbits, ncls)
        hist = st.hist(plan, qp, qlp, gp, glp)
        sc = st.scan(plan, hist, 1, 0, k)
        totals = sc["total"][:num_query].cpu()
        if int(bad.item()) != 0:
            raise ValueError("calc_map_k: codes must be +-1 and labels 0/1 (%d offending elements)" % int(bad.item()))
        cap = max(int(totals.max().item()), 1)
        rows = max(1, min(num_query, _PARITY_SLAB_BYTES // (4 * cap)))
        running = 0
        if rows >= num_query:
            tindex = torch.zeros((num_query, cap), dtype=torch.int32, device=dev)
            st.rank_map(plan, qp, qlp, gp, glp, sc, tindex)
            running = _parity_reduce(tindex.cpu(), totals, running)
        else:  # slabs of queries: rows are independent, the running fp32 sum stays sequential
            for lo in range(0, num_query, rows):
                hi = min(lo + rows, num_query)
                sub = st.make_plan(hi - lo, n, nbits, ncls)
                h = st.hist(sub, qp[lo:hi], qlp[lo:hi], gp, glp)
                s = st.scan(sub, h, 1, 0, k)
                tindex = torch.zeros((hi - lo, cap), dtype=torch.int32, device=dev)
                st.rank_map(sub, qp[lo:hi], qlp[lo:hi], gp, glp, s, tindex)
                running = _parity_reduce(tindex.cpu(), totals[lo:hi], running)
        result = running / num_query
        return result if isinstance(result, torch.Tensor) else torch.tensor(float(result), dtype=torch.float32)


def calc_map_k_packed(q_codes, r_codes, query_L, retrieval_L, nbits: int, k=None):
    """``calc_map_k`` on codes that are already bit-packed on the GPU (``models.get_code`` / ``encode_*_packed``:
    int32 ``[n, W]``, bit b of word w = code column 32w+b).  Labels as in the reference (multi-hot ``[n, C]``, any
    device).  Returns the same 0-dim fp32 CPU tensor as ``calc_map_k`` — no +-1 fp32 buffers, no device->host hop of
    the codes (the reference forces them to the CPU at common/calc_utils.py:62-64)."""
    dev = q_codes.device
    if not (q_codes.is_cuda and r_codes.is_cuda):
        raise CmhError("packed codes must live on the GPU")
    ncls = retrieval_L.shape[1]
    if nbits > 128 or ncls > 128:
        raise CmhError("calc_map_k supports up to 128 bits and 128 classes (got %d, %d)" % (nbits, ncls))
    if q_codes.shape[1] != R.code_words(nbits) or r_codes.shape[1] != R.code_words(nbits):
        raise CmhError("packed codes must have %d words per row for %d bits" % (R.code_words(nbits), nbits))
    k = retrieval_L.shape[0] if k is None else k
    with torch.cuda.device(dev):
        bad = R.new_bad_counter(dev)
        qlp = R.pack_labels(_to_dev(query_L.detach(), dev), bad)
        glp = R.pack_labels(_to_dev(retrieval_L.detach(), dev), bad)
        res = R.map_k(q_codes.contiguous(), qlp, r_codes.contiguous(), glp, nbits, ncls, k)
        host = torch.stack([res.map, bad[0].to(torch.float64)]).cpu()
    if host[1] != 0:
        raise ValueError("calc_map_k_packed: labels must be 0/1 (%d offending elements)" % int(host[1]))
    return host[0].to(torch.float32)


def hamming_topk(qB, rB, k: int):
    """First ``k`` columns of ``torch.sort(calc_hammingDist(qB, rB), stable=True)`` (calc_utils.py:76-77)
    -> ``(dist [Q,k] fp32, index [Q,k] int64)`` on the inputs' device (north_star's "per-query top-k")."""
    dev = qB.device if qB.is_cuda else (rB.device if rB.is_cuda else _device())
    nbits = rB.shape[1]
    with torch.cuda.device(dev):
        bad = R.new_bad_counter(dev)
        qp = R.pack_codes(_to_dev(qB.detach(), dev), bad)
        gp = R.pack_codes(_to_dev(rB.detach(), dev), bad)
        keys = R.topk(qp, gp, nbits, k)
        dist, idx = R.split_keys(keys)
        if int(bad.item()) != 0:
            raise ValueError("hamming_topk: codes must be +-1")
    return _back(dist.to(torch.float32), qB), _back(idx, qB)


def install_into_reference(module=None):
    """Monkey-patch a loaded reference ``common.calc_utils`` module (or import it) with these functions."""
    if module is None:
        import importlib

        module = importlib.import_module("common.calc_utils")
    for name in ("calc_label_sim", "generate_weight_sim", "euclidean_similarity", "cosine_similarity",
                 "calc_hammingDist", "calc_map_k"):
        setattr(module, name, globals()[name])
    return module
