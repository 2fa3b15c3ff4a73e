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


def _needs_grad(*ts) -> bool:
    """The similarity helpers are also called inside training losses (models/DCMHT/DCMHT.py:78, models/baseline/model.py:128)
    on hash outputs that require grad.  The kernels are forward-only, so such calls keep the reference's own differentiable
    torch expression (on the inputs' device) instead of silently cutting the graph."""
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in ts)


# a4 -------------------------------------------------------------------------------------------------------
def calc_label_sim(a: torch.Tensor, b: torch.Tensor):
    """``(a.matmul(b.T) > 0).float()`` (common/calc_utils.py:8-10); result on ``a``'s device."""
    if _needs_grad(a, b):
        return (a.matmul(b.transpose(0, 1)) > 0).float()
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
        if _needs_grad(a, b):
            return torch.cdist(a, b, p=2.0)
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
        if _needs_grad(a, b):
            return torch.matmul(a / a.norm(dim=-1, keepdim=True), (b / b.norm(dim=-1, keepdim=True)).t())
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
class _PackedCache:
    """Packed labels of the last few distinct label tensors.  ``BaseTrainer.valid`` (runners/base.py:317-321) calls
    calc_map_k four times with the same two label matrices (75 MB of int64 at C2): they are uploaded and packed once.
    An entry is keyed by the tensor object (weak reference), its version counter, storage pointer, shape and device."""

    def __init__(self, capacity: int = 4):
        self.capacity = capacity
        self.entries = []  # (weakref, version, data_ptr, shape, dtype, device, packed, bad_count)

    def get(self, t: torch.Tensor, dev: torch.device):
        import weakref

        for e in self.entries:
            if e[0]() is t and e[1] == t._version and e[2] == t.data_ptr() and e[3] == tuple(t.shape) and e[4] == t.dtype and e[5] == dev:
                return e[6], e[7]
        bad = R.new_bad_counter(dev)
        packed = R.pack_labels(_to_dev(t.detach(), dev), bad)
        nbad = int(bad.item())
        try:
            ref = weakref.ref(t)
        except TypeError:
            return packed, nbad
        self.entries.insert(0, (ref, t._version, t.data_ptr(), tuple(t.shape), t.dtype, dev, packed, nbad))
        del self.entries[self.capacity:]
        return packed, nbad

    def clear(self):
        self.entries.clear()


_LABEL_CACHE = _PackedCache()


def _parity_reduce(tindex_rows: torch.Tensor, totals: torch.Tensor, running):
    """calc_utils.py:84-89 on the host for a slab of queries: same torch CPU ops, same order."""
    for row in range(tindex_rows.shape[0]):
        total = totals[row]
        count = torch.arange(1, total + 1).type(torch.float32)
        tindex = (tindex_rows[row, :total] - 1).type(torch.float32) + 1.0
        running = running + torch.mean(count / tindex)
    return running


def _map_k_dense(qB, rB, query_L, retrieval_L, k, dev):
    """calc_map_k for codes that are NOT +-1 (a 0 from ``sign_()`` of an exact zero, +-2 rows from the DistributedSampler
    padding of runners/base.py:180-190,263-264): the reference still computes a value there (common/calc_utils.py:51-56,76),
    so this path evaluates the same expression — fp32 ``0.5 * (K - q.r)`` (cmh_hamming_dense_f32), stable sort, AP loop —
    on the GPU in query slabs.  Rare path: it materialises slab x N floats, which the packed evaluator never does."""
    qf, rf = _as_f32_dev(qB, dev), _as_f32_dev(rB, dev)
    ql, rl = _as_f32_dev(query_L, dev), _as_f32_dev(retrieval_L, dev)
    Q, N = qf.shape[0], rf.shape[0]
    slab = max(1, min(Q, (1 << 26) // max(N, 1)))
    total_sum = torch.zeros((), dtype=torch.float64, device=dev)
    for lo in range(0, Q, slab):
        hi = min(lo + slab, Q)
        gnd = _sim("label", ql[lo:hi], rl)                      # [s, N] 0/1
        hamm = R.hamming_dense(qf[lo:hi], rf)
        order = torch.sort(hamm, dim=-1, stable=True)[1]
        g = torch.gather(gnd, 1, order)
        tsum = g.sum(dim=1)
        total = torch.clamp(tsum, max=float(k))
        crel = torch.cumsum(g, dim=1)                           # 1-based rank among the relevant items
        pos = torch.arange(1, N + 1, device=dev, dtype=torch.float32)[None, :]
        terms = torch.where((g > 0) & (crel <= total[:, None]), crel / pos, torch.zeros_like(crel))
        total_sum += (terms.sum(dim=1, dtype=torch.float64) / total.to(torch.float64)).sum()   # 0/0 -> nan like the reference
    return (total_sum / Q).to(torch.float32).cpu()


def calc_map_k(qB, rB, query_L, retrieval_L, k=None, *, mode: Optional[str] = None):
    """mAP over the first ``min(R, k)`` relevant items of the full Hamming ranking
    (common/calc_utils.py:58-92).  Accepts tensors on any device; returns a 0-dim fp32 CPU tensor.

    Ties in distance are ranked by ascending gallery index (``torch.sort(..., stable=True)``), the
    canonical order of this repo; the reference's unstable CPU sort is unspecified there.
    Codes that are not exactly +-1 take the dense fp32 path (`_map_k_dense`) like the reference's own arithmetic;
    labels are multi-hot: any non-zero entry counts as "has the class" (the reference tests ``gram > 0``)."""
    mode = mode or os.environ.get("CMH_MAP_MODE", "device")
    if mode not in ("device", "parity"):
        raise ValueError("mode must be 'device' or 'parity'")
    dev = next((t.device for t in (qB, rB, query_L, retrieval_L) if isinstance(t, torch.Tensor) and t.is_cuda), None)
    dev = dev or _device()
    num_query = query_L.shape[0]
    nbits, ncls = rB.shape[1], retrieval_L.shape[1]
    n = retrieval_L.shape[0]
    if k is None:
        k = n
    if k <= 0:
        raise ValueError("k must be positive or None")
    with torch.cuda.device(dev):
        if nbits > 128 or ncls > 128:
            return _map_k_dense(qB, rB, query_L, retrieval_L, k, dev)
        bad = R.new_bad_counter(dev)
        qp = R.pack_codes(_to_dev(qB.detach(), dev), bad)
        gp = R.pack_codes(_to_dev(rB.detach(), dev), bad)
        qlp, _ = _LABEL_CACHE.get(query_L, dev)
        glp, _ = _LABEL_CACHE.get(retrieval_L, dev)
        if mode == "device":
            res = R.map_k(qp, qlp, gp, glp, nbits, ncls, k)
            host = torch.stack([res.map, bad[0].to(torch.float64)]).cpu()
            if host[1] != 0:
                return _map_k_dense(qB, rB, query_L, retrieval_L, k, dev)
            return host[0].to(torch.float32)

        # parity mode: integer ranks from the GPU, fp32 reduction by the reference's own CPU ops
        st = R.CudaStages()
        plan = st.make_plan(num_query, n, nbits, ncls)
        hist = st.hist(plan, qp, qlp, gp, glp)
        sc = st.scan(plan, hist, 1, 0, k)
        totals = sc["total"][:num_query].cpu()
        if int(bad.item()) != 0:
            return _map_k_dense(qB, rB, query_L, retrieval_L, k, dev)
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
    device; packed once and cached across calls).  Returns the same 0-dim fp32 CPU tensor as ``calc_map_k`` — no +-1
    fp32 buffers, no device->host hop of the codes (the reference forces them to the CPU at common/calc_utils.py:62-64)."""
    dev = q_codes.device
    if not (q_codes.is_cuda and r_codes.is_cuda):
        raise CmhError("packed codes must live on the GPU")
    ncls = retrieval_L.shape[1]
    if nbits > 128 or ncls > 128:
        raise CmhError("calc_map_k_packed supports up to 128 bits and 128 classes (got %d, %d)" % (nbits, ncls))
    if q_codes.shape[1] != R.code_words(nbits) or r_codes.shape[1] != R.code_words(nbits):
        raise CmhError("packed codes must have %d words per row for %d bits" % (R.code_words(nbits), nbits))
    k = retrieval_L.shape[0] if k is None else k
    with torch.cuda.device(dev):
        qlp, _ = _LABEL_CACHE.get(query_L, dev)
        glp, _ = _LABEL_CACHE.get(retrieval_L, dev)
        res = R.map_k(q_codes.contiguous(), qlp, r_codes.contiguous(), glp, nbits, ncls, k)
        return res.map.to(torch.float32).cpu()


def valid_packed(query_img, query_txt, retrieval_img, retrieval_txt, query_labels, retrieval_labels, nbits: int, k=None):
    """The four retrieval directions of ``BaseTrainer.valid`` / ``test`` (runners/base.py:317-321, 351-355) on packed codes
    from ``models.get_code``: labels are uploaded and packed ONCE, the four evaluations are queued back to back and their
    scalars come back in one device->host copy.  Returns ``(mAPi2t, mAPt2i, mAPi2i, mAPt2t)`` as 0-dim fp32 CPU tensors, in
    the reference's order."""
    dev = query_img.device
    ncls = retrieval_labels.shape[1]
    k = retrieval_labels.shape[0] if k is None else k
    with torch.cuda.device(dev):
        qlp, _ = _LABEL_CACHE.get(query_labels, dev)
        glp, _ = _LABEL_CACHE.get(retrieval_labels, dev)
        pairs = ((query_img, retrieval_txt), (query_txt, retrieval_img), (query_img, retrieval_img), (query_txt, retrieval_txt))
        maps = [R.map_k(q.contiguous(), qlp, g.contiguous(), glp, nbits, ncls, k).map for q, g in pairs]
        host = torch.stack(maps).to(torch.float32).cpu()
    return tuple(host[i] for i in range(4))


def hamming_topk(qB, rB, k: int, out=None):
    """First ``k`` columns of ``torch.sort(calc_hammingDist(qB, rB), stable=True)`` (calc_utils.py:76-77)
    -> ``(dist [Q,k] fp32, index [Q,k] int64)`` on the inputs' device (north_star's "per-query top-k").
    ``out=(dist, index)``: optional preallocated (e.g. pinned host) result tensors."""
    dev = qB.device if qB.is_cuda else (rB.device if rB.is_cuda else _device())
    nbits = rB.shape[1]
    with torch.cuda.device(dev):
        bad = R.new_bad_counter(dev)
        if not qB.is_cuda and not rB.is_cuda and qB.dtype == torch.float32 and rB.dtype == torch.float32 and nbits <= 128:
            keys = R.topk_from_host(qB.detach(), rB.detach(), k, dev, bad=bad)    # host inputs: transfer overlapped with compute
        else:
            qp = R.pack_codes(_to_dev(qB.detach(), dev), bad)
            gp = R.pack_codes(_to_dev(rB.detach(), dev), bad)
            keys = R.topk(qp, gp, nbits, k)
        dist, idx = R.split_keys(keys, dist_dtype=torch.float32)
        return _topk_result(dist, idx, bad, qB, out)


def _topk_result(dist, idx, bad, like, out):
    if out is not None:
        out[0].copy_(dist, non_blocking=True)
        out[1].copy_(idx, non_blocking=True)
        nbad = int(bad.item())   # synchronises: the copies above have landed
        res = out
    else:
        nbad = int(bad.item())
        res = (_back(dist, like), _back(idx, like))
    if nbad != 0:
        raise ValueError("hamming_topk: codes must be +-1 (%d offending elements)" % nbad)
    return res


def hamming_topk_sharded(qB, rB_shard, k: int, idx_offset: int, n_geom: Optional[int] = None, group=None, evaluator=None,
                         method: str = "auto", out=None, return_keys: bool = False):
    """``hamming_topk`` with the gallery sharded by contiguous index range over the ranks of ``group`` (one process per GPU):
    ``rB_shard`` holds this rank's items ``[idx_offset, idx_offset + len)``; queries are replicated.  Every rank returns the
    same global ``(dist, index)``; ``out`` as in ``hamming_topk``.  ``return_keys=True`` skips the split / host copy and hands
    back the int64 ``(dist << 32) | index`` keys on the GPU (ranks that do not need the result on the host)."""
    dev = qB.device if qB.is_cuda else (rB_shard.device if rB_shard.is_cuda else _device())
    nbits = rB_shard.shape[1]
    with torch.cuda.device(dev):
        ev = evaluator if evaluator is not None else R.ShardedEvaluator(group)
        bad = R.new_bad_counter(dev)
        qp = R.pack_codes(_to_dev(qB.detach(), dev), bad)
        gp = R.pack_codes(_to_dev(rB_shard.detach(), dev), bad)
        keys = ev.topk(qp, gp, nbits, k, idx_offset, n_geom=n_geom, method=method)
        if return_keys:
            if int(bad.item()) != 0:
                raise ValueError("hamming_topk: codes must be +-1")
            return keys
        dist, idx = R.split_keys(keys, dist_dtype=torch.float32)
        return _topk_result(dist, idx, bad, qB, out)


_PATCHED = ("calc_label_sim", "generate_weight_sim", "euclidean_similarity", "cosine_similarity", "calc_hammingDist", "calc_map_k")


def install_into_reference(module=None):
    """Bind this module's functions into a reference checkout.

    The reference binds the evaluator by value at import time (``from common.calc_utils import calc_map_k`` in
    runners/base.py:5, runners/MITH/runner.py:5, models/DCMHT/DCMHT.py:8, ...), so patching ``common.calc_utils`` alone is
    not enough once those modules are loaded: every already-imported ``runners.*`` / ``models.*`` module whose attribute IS
    the original reference function is re-bound too, and call sites imported later pick the patched module attributes.  Works
    both before and after ``import runners``.  The similarity helpers stay differentiable: with an input that
    requires grad they evaluate the reference's own torch expression (`_needs_grad`).  Returns the patched module."""
    import importlib
    import sys

    if module is None:
        module = importlib.import_module("common.calc_utils")
    originals = {name: getattr(module, name, None) for name in _PATCHED}
    for name in _PATCHED:
        setattr(module, name, globals()[name])
    for modname, mod in list(sys.modules.items()):
        if mod is None or mod is module or not (modname.startswith("runners") or modname.startswith("models")):
            continue
        for name, orig in originals.items():
            if orig is not None and getattr(mod, name, None) is orig:
                setattr(mod, name, globals()[name])
    return module
