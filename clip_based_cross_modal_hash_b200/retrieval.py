"""Host side of the bit-packed retrieval evaluator: torch tensors in, libcmh.so kernels underneath.

Everything here works on CUDA tensors and launches on torch's current stream.  torch is plumbing only
(device memory, streams, ``torch.distributed``); all arithmetic happens in the kernels of
``csrc/cmh_retrieval.cu`` / ``csrc/cmh_pack.cu`` through the C ABI of ``include/cmh.h``.

Stage functions (``hist`` -> ``scan`` -> ``rank_map`` / ``rank_topk`` -> ``map_finish`` / ``topk_merge``)
mirror the C entry points one to one; ``map_k`` / ``topk`` chain them for one GPU and
``ShardedEvaluator`` chains them across the ranks of a process group with the gallery sharded by
contiguous index range.

Reference being replaced: ``common/calc_utils.py:51-92`` (calc_hammingDist, calc_map_k) and the fp32 +-1
code buffers / all-reduce of ``runners/base.py:242-266``.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import CmhError, Plan, TcOperands, check

EMPTY_KEY = -1  # 0xFFFF_FFFF_FFFF_FFFF viewed as int64
_LABEL_DT = {torch.int64: 0, torch.float32: 1, torch.uint8: 2, torch.bool: 2, torch.int32: 3}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise CmhError("expected a CUDA tensor (there is no CPU path)")
        if dev is not None and t.device != dev:
            raise CmhError("tensors live on different devices")
        dev = t.device
    return dev


def code_words(nbits: int) -> int:
    w = _lib.lib().cmh_code_words(nbits)
    if w < 0:
        raise CmhError("code length %d not supported (1..128 bits)" % nbits)
    return w


def label_words(ncls: int) -> int:
    w = _lib.lib().cmh_label_words(ncls)
    if w < 0:
        raise CmhError("%d classes not supported (0..128)" % ncls)
    return w


# ---------------------------------------------------------------------------------------------------------
# R0 packing
# ---------------------------------------------------------------------------------------------------------
def new_bad_counter(device) -> torch.Tensor:
    return torch.zeros(1, dtype=torch.int64, device=device)


def pack_codes(codes: torch.Tensor, bad: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[n, K] fp32 (+-1) CUDA -> [n, W] int32 (uint32 bit patterns).  ``bad`` (int64[1], device) counts
    elements that are not exactly +-1 (it is added to, not reset).  ``out``: optional contiguous [n, W] int32 destination."""
    _need_cuda(codes, bad)
    if codes.dim() != 2:
        raise CmhError("codes must be [n, K]")
    if codes.dtype != torch.float32:
        codes = codes.to(torch.float32)
    if codes.stride(1) != 1:
        codes = codes.contiguous()
    n, nbits = codes.shape
    if out is None:
        out = torch.empty((n, code_words(nbits)), dtype=torch.int32, device=codes.device)
    elif out.shape != (n, code_words(nbits)) or out.dtype != torch.int32 or not out.is_contiguous():
        raise CmhError("out must be a contiguous int32 [n, W] tensor")
    with torch.cuda.device(codes.device):
        check(_lib.lib().cmh_pack_codes_f32(codes.data_ptr(), n, nbits, codes.stride(0) if n > 1 else nbits,
                                            out.data_ptr(), _ptr(bad), _stream()))
    return out


def pack_labels(labels: torch.Tensor, bad: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[n, C] multi-hot (int64 / int32 / uint8 / bool / fp32) CUDA -> [n, LW] int32."""
    _need_cuda(labels, bad)
    if labels.dim() != 2:
        raise CmhError("labels must be [n, C]")
    if labels.dtype not in _LABEL_DT:
        labels = labels.to(torch.float32)
    if labels.stride(1) != 1:
        labels = labels.contiguous()
    n, ncls = labels.shape
    lw = label_words(ncls)
    if lw == 0:
        raise CmhError("labels need at least one class")
    out = torch.empty((n, lw), dtype=torch.int32, device=labels.device)
    with torch.cuda.device(labels.device):
        check(_lib.lib().cmh_pack_labels(labels.data_ptr(), _LABEL_DT[labels.dtype], n, ncls,
                                         labels.stride(0) if n > 1 else ncls, out.data_ptr(), _ptr(bad), _stream()))
    return out


def unpack_codes(packed: torch.Tensor, nbits: int) -> torch.Tensor:
    """[n, W] int32 -> [n, K] fp32 +-1 (the reference's code format)."""
    _need_cuda(packed)
    n = packed.shape[0]
    out = torch.empty((n, nbits), dtype=torch.float32, device=packed.device)
    with torch.cuda.device(packed.device):
        check(_lib.lib().cmh_unpack_codes_f32(packed.data_ptr(), n, nbits, out.data_ptr(), nbits, _stream()))
    return out


# ---------------------------------------------------------------------------------------------------------
# R1 materialised Hamming matrix (calc_hammingDist)
# ---------------------------------------------------------------------------------------------------------
def hamming_matrix(qp: torch.Tensor, gp: torch.Tensor, nbits: int) -> torch.Tensor:
    _need_cuda(qp, gp)
    Q, N = qp.shape[0], gp.shape[0]
    out = torch.empty((Q, N), dtype=torch.float32, device=qp.device)
    with torch.cuda.device(qp.device):
        check(_lib.lib().cmh_hamming_f32(qp.data_ptr(), Q, gp.data_ptr(), N, nbits, out.data_ptr(), N, _stream()))
    return out


def hamming_dense(b1: torch.Tensor, b2: torch.Tensor) -> torch.Tensor:
    _need_cuda(b1, b2)
    b1 = b1.to(torch.float32).contiguous()
    b2 = b2.to(torch.float32).contiguous()
    Q, K = b1.shape
    N = b2.shape[0]
    out = torch.empty((Q, N), dtype=torch.float32, device=b1.device)
    with torch.cuda.device(b1.device):
        check(_lib.lib().cmh_hamming_dense_f32(b1.data_ptr(), Q, b2.data_ptr(), N, K, out.data_ptr(), _stream()))
    return out


# ---------------------------------------------------------------------------------------------------------
# evaluator stages (one C entry point each)
# ---------------------------------------------------------------------------------------------------------
TC_BLOCKS_PER_SM = int(os.environ.get("CMH_TC_BLOCKS_PER_SM", "0"))   # 0 = the library default (16); measured flat between 4 and 16
_SM_COUNT = None


def _sm_count() -> int:
    global _SM_COUNT
    if _SM_COUNT is None:
        sm = ctypes.c_int(0)
        rc = _lib.lib().cmh_device_info(ctypes.byref(sm), None, None) if torch.cuda.is_available() else -1
        _SM_COUNT = sm.value if rc == 0 and sm.value > 0 else 148
    return _SM_COUNT


class Operands:
    """int8 operand rows of the tensor-core ranking passes (``cmh_tc_*``), expanded once per evaluation from the packed words and
    shared by the histogram and the rank pass.  Holds the device buffers alive."""

    def __init__(self, q_codes, q_labels, g_codes, g_labels):
        self.q_codes, self.q_labels, self.g_codes, self.g_labels = q_codes, q_labels, g_codes, g_labels
        self.c = TcOperands(q_codes.data_ptr(), _ptr(q_labels), g_codes.data_ptr() if g_codes.numel() else None, 
                            _ptr(g_labels) if (g_labels is not None and g_labels.numel()) else None,
                            q_codes.shape[1], 0 if q_labels is None else q_labels.shape[1])


def _expand(packed: torch.Tensor, rows: int, ncols: int, kind: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    nb = _lib.lib().cmh_tc_operand_bytes(ncols)
    if nb < 0:
        raise CmhError("operand width %d not supported" % ncols)
    if out is None:
        out = torch.empty((rows, nb), dtype=torch.int8, device=packed.device)
    check(_lib.lib().cmh_tc_expand(packed.data_ptr(), packed.shape[0], rows, packed.shape[1], ncols, kind, out.data_ptr(), _stream()))
    return out


class CudaStages:
    """The product's stage kernels.  ``ShardedEvaluator`` takes any object with these methods so that the
    exchange logic can be exercised on CPU (gloo) in tests with an oracle-backed stand-in.

    ``tensor_cores=True`` (default): hist / rank_topk / rank_map run with the distances on tcgen05 (csrc/cmh_tc.cu) when the caller
    passes the ``ops`` made by ``operands``; ``False`` keeps the XOR+POPC kernels of csrc/cmh_retrieval.cu (same results)."""

    def __init__(self, tensor_cores: Optional[bool] = None):
        if tensor_cores is None:
            import os
            tensor_cores = os.environ.get("CMH_NO_TC", "0") in ("", "0")
        self.tensor_cores = bool(tensor_cores)

    def make_plan(self, Q, N, nbits, ncls, N_geom=None, target_blocks=0) -> Plan:
        if target_blocks == 0 and self.tensor_cores and TC_BLOCKS_PER_SM > 0:   # tuning knob (scripts/sweep_blocks.sh)
            target_blocks = TC_BLOCKS_PER_SM * _sm_count()
        return _lib.make_plan(Q, N, nbits, ncls, N_geom, target_blocks)

    def operands(self, plan: Plan, qp, qlp, gp, glp) -> Optional[Operands]:
        """Expand the packed words into the int8 operand rows of the tensor-core passes (None when they are switched off)."""
        if not self.tensor_cores:
            return None
        dev = _need_cuda(qp, qlp, gp, glp)
        with torch.cuda.device(dev):
            with_labels = qlp is not None and glp is not None and plan.ncls > 0
            return Operands(_expand(qp, plan.Qpad, plan.nbits, 0), _expand(qlp, plan.Qpad, plan.ncls, 1) if with_labels else None,
                            _expand(gp, gp.shape[0], plan.nbits, 0), _expand(glp, glp.shape[0], plan.ncls, 2) if with_labels else None)

    def hist(self, plan: Plan, qp, qlp, gp, glp, ops: Optional[Operands] = None) -> torch.Tensor:
        dev = _need_cuda(qp, qlp, gp, glp)
        hist = torch.empty((plan.nchunks, plan.bins, plan.Qpad), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            if ops is not None:
                with_labels = qlp is not None and glp is not None and ops.q_labels is not None
                check(_lib.lib().cmh_tc_hist(ctypes.byref(plan), ctypes.byref(ops.c), int(with_labels), hist.data_ptr(), _stream()))
            else:
                check(_lib.lib().cmh_hist(ctypes.byref(plan), qp.data_ptr(), _ptr(qlp), gp.data_ptr(), _ptr(glp),
                                          hist.data_ptr(), _stream()))
        return hist

    def scan(self, plan: Plan, hist_all: torch.Tensor, world: int, rank: int, k: Optional[int],
             with_rel: bool = True) -> Dict[str, torch.Tensor]:
        dev = _need_cuda(hist_all)
        shape = (plan.nchunks, plan.bins, plan.Qpad)
        o = {
            "within_all": torch.empty(shape, dtype=torch.int32, device=dev),
            "below_all": torch.empty((plan.bins, plan.Qpad), dtype=torch.int32, device=dev),
            "within_rel": torch.empty(shape, dtype=torch.int32, device=dev) if with_rel else None,
            "below_rel": torch.empty((plan.bins, plan.Qpad), dtype=torch.int32, device=dev) if with_rel else None,
            "tsum": torch.empty(plan.Qpad, dtype=torch.int32, device=dev),
            "total": torch.empty(plan.Qpad, dtype=torch.int32, device=dev),
            "thresh": torch.empty(plan.Qpad, dtype=torch.int32, device=dev),
        }
        with torch.cuda.device(dev):
            check(_lib.lib().cmh_scan(ctypes.byref(plan), hist_all.data_ptr(), world, rank, 0 if k is None else int(k),
                                      o["within_all"].data_ptr(), _ptr(o["within_rel"]), o["below_all"].data_ptr(),
                                      _ptr(o["below_rel"]), o["tsum"].data_ptr(), o["total"].data_ptr(),
                                      o["thresh"].data_ptr(), _stream()))
        return o

    def hist_totals(self, plan: Plan, hist: torch.Tensor) -> torch.Tensor:
        """[nchunks, bins, Qpad] packed histogram block -> this rank's bucket totals int32 [2, bins, Qpad] (all, relevant)."""
        dev = _need_cuda(hist)
        totals = torch.empty((2, plan.bins, plan.Qpad), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().cmh_hist_totals(ctypes.byref(plan), hist.data_ptr(), totals.data_ptr(), _stream()))
        return totals

    def scan_sharded(self, plan: Plan, hist_local: torch.Tensor, totals_all: torch.Tensor, world: int, rank: int,
                     k: Optional[int]) -> Dict[str, torch.Tensor]:
        """``scan`` for one rank of a sharded gallery from its own histogram block and everyone's bucket totals."""
        dev = _need_cuda(hist_local, totals_all)
        shape = (plan.nchunks, plan.bins, plan.Qpad)
        o = {
            "within_all": torch.empty(shape, dtype=torch.int32, device=dev),
            "below_all": torch.empty((plan.bins, plan.Qpad), dtype=torch.int32, device=dev),
            "within_rel": torch.empty(shape, dtype=torch.int32, device=dev),
            "below_rel": torch.empty((plan.bins, plan.Qpad), dtype=torch.int32, device=dev),
            "tsum": torch.empty(plan.Qpad, dtype=torch.int32, device=dev),
            "total": torch.empty(plan.Qpad, dtype=torch.int32, device=dev),
            "thresh": torch.empty(plan.Qpad, dtype=torch.int32, device=dev),
        }
        with torch.cuda.device(dev):
            check(_lib.lib().cmh_scan_sharded(ctypes.byref(plan), hist_local.data_ptr(), totals_all.data_ptr(), world, rank,
                                              0 if k is None else int(k), o["within_all"].data_ptr(), o["within_rel"].data_ptr(),
                                              o["below_all"].data_ptr(), o["below_rel"].data_ptr(), o["tsum"].data_ptr(),
                                              o["total"].data_ptr(), o["thresh"].data_ptr(), _stream()))
        return o

    def rank_map(self, plan: Plan, qp, qlp, gp, glp, sc: Dict[str, torch.Tensor],
                 tindex: Optional[torch.Tensor] = None, n_total: Optional[int] = None, ops: Optional[Operands] = None) -> torch.Tensor:
        dev = _need_cuda(qp, qlp, gp, glp, tindex)
        ap_partial = torch.empty((plan.nchunks, plan.Qpad), dtype=torch.float64, device=dev)
        cap = 0
        if tindex is not None:
            if tindex.dtype != torch.int32 or tindex.dim() != 2 or tindex.shape[0] < plan.Q or not tindex.is_contiguous():
                raise CmhError("tindex must be a contiguous int32 [Q, cap] tensor")
            cap = tindex.shape[1]
        nt = plan.N if n_total is None else n_total
        if ops is not None and nt < (1 << 24):
            with torch.cuda.device(dev):
                check(_lib.lib().cmh_tc_rank_map(ctypes.byref(plan), ctypes.byref(ops.c), sc["within_all"].data_ptr(),
                                                 sc["within_rel"].data_ptr(), sc["below_all"].data_ptr(), sc["below_rel"].data_ptr(),
                                                 sc["total"].data_ptr(), nt, ap_partial.data_ptr(), _ptr(tindex), cap, _stream()))
            return ap_partial
        with torch.cuda.device(dev):
            check(_lib.lib().cmh_rank_map(ctypes.byref(plan), qp.data_ptr(), qlp.data_ptr(), gp.data_ptr(),
                                          glp.data_ptr(), sc["within_all"].data_ptr(), sc["within_rel"].data_ptr(),
                                          sc["below_all"].data_ptr(), sc["below_rel"].data_ptr(),
                                          sc["total"].data_ptr(), plan.N if n_total is None else n_total,
                                          ap_partial.data_ptr(), _ptr(tindex), cap, _stream()))
        return ap_partial

    def ap_reduce(self, plan: Plan, ap_partial: torch.Tensor) -> torch.Tensor:
        """[nchunks, Qpad] chunk partials -> [1, Qpad]: what a rank contributes to the AP exchange."""
        dev = _need_cuda(ap_partial)
        out = torch.empty((1, plan.Qpad), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().cmh_ap_reduce(ctypes.byref(plan), ap_partial.data_ptr(), ap_partial.numel() // plan.Qpad,
                                           out.data_ptr(), _stream()))
        return out

    def map_finish(self, plan: Plan, ap_partial_all: torch.Tensor, total: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        dev = _need_cuda(ap_partial_all, total)
        ap = torch.empty(plan.Q, dtype=torch.float64, device=dev)
        out = torch.empty((), dtype=torch.float64, device=dev)
        nparts = ap_partial_all.numel() // plan.Qpad
        with torch.cuda.device(dev):
            check(_lib.lib().cmh_map_finish(ctypes.byref(plan), ap_partial_all.data_ptr(), nparts, total.data_ptr(),
                                            ap.data_ptr(), out.data_ptr(), _stream()))
        return ap, out

    def rank_topk(self, plan: Plan, qp, gp, sc: Dict[str, torch.Tensor], k: int, idx_offset: int,
                  keys: Optional[torch.Tensor] = None, ops: Optional[Operands] = None) -> torch.Tensor:
        dev = _need_cuda(qp, gp)
        if keys is None:
            keys = torch.full((plan.Q, k), EMPTY_KEY, dtype=torch.int64, device=dev)
        if ops is not None:
            with torch.cuda.device(dev):
                check(_lib.lib().cmh_tc_rank_topk(ctypes.byref(plan), ctypes.byref(ops.c), sc["within_all"].data_ptr(),
                                                  sc["below_all"].data_ptr(), sc["thresh"].data_ptr(), k, idx_offset,
                                                  keys.data_ptr(), _stream()))
            return keys
        with torch.cuda.device(dev):
            check(_lib.lib().cmh_rank_topk(ctypes.byref(plan), qp.data_ptr(), gp.data_ptr(),
                                           sc["within_all"].data_ptr(), sc["below_all"].data_ptr(),
                                           sc["thresh"].data_ptr(), k, idx_offset, keys.data_ptr(), _stream()))
        return keys

    # ---- candidate path of the top-k (cmh_tc_topk_*) ----
    def topk_cutoff(self, plan_s: Plan, hist_s: torch.Tensor, n_local: int, k: int) -> torch.Tensor:
        """-> int32 [2, Qpad]: row 0 = cutoff distance T, row 1 = shard-index bound for bucket T."""
        cut = torch.empty((2, plan_s.Qpad), dtype=torch.int32, device=hist_s.device)
        with torch.cuda.device(hist_s.device):
            check(_lib.lib().cmh_tc_topk_cutoff(ctypes.byref(plan_s), hist_s.data_ptr(), n_local, k, cut[0].data_ptr(),
                                                cut[1].data_ptr(), _stream()))
        return cut

    def topk_sample_block(self, plan: Plan, plan_s: Optional[Plan], hist_s: Optional[torch.Tensor], idx_offset: int, rank: int,
                          world: int, device) -> torch.Tensor:
        """This rank's block of the sharded sample exchange, int32 ``[bins + 1, Qpad]``: per-distance sample counts + the header
        row ([0] sample items, [1] shard items, [2 + rank] first gallery index).  ``hist_s`` None: an empty shard."""
        meta = torch.empty((plan.bins + 1, plan.Qpad), dtype=torch.int32, device=device)
        with torch.cuda.device(device):
            check(_lib.lib().cmh_tc_topk_sample_block(ctypes.byref(plan_s) if plan_s is not None else None, _ptr(hist_s), plan.Qpad,
                                                      plan.bins, plan.N, idx_offset, rank, world, meta.data_ptr(), _stream()))
        return meta

    def topk_cutoff_sharded(self, plan: Plan, sample_sum: torch.Tensor, k: int, rank: int, world: int) -> torch.Tensor:
        """Global cutoff from the gathered sample blocks ``[world, bins + 1, Qpad]`` of all ranks (``collect_candidates``), index
        bound translated into this shard."""
        if tuple(sample_sum.shape) != (world, plan.bins + 1, plan.Qpad) or not sample_sum.is_contiguous():
            raise CmhError("topk_cutoff_sharded: sample blocks must be a contiguous [world, bins + 1, Qpad] tensor")
        cut = torch.empty((2, plan.Qpad), dtype=torch.int32, device=sample_sum.device)
        with torch.cuda.device(sample_sum.device):
            check(_lib.lib().cmh_tc_topk_cutoff_sharded(ctypes.byref(plan), sample_sum.data_ptr(), k, rank, world, cut[0].data_ptr(),
                                                        cut[1].data_ptr(), _stream()))
        return cut

    def topk_collect(self, plan: Plan, ops: Operands, cutoff: torch.Tensor, cap: int, out=None, chunks: Optional[Tuple[int, int]] = None
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
        """``chunks=(c0, c1)`` collects only chunks [c0, c1) into the ``out=(cand, cnt)`` buffers of an earlier call (slab pipeline)."""
        dev = cutoff.device
        if out is None:
            cand = torch.empty((plan.nchunks, plan.Qpad, cap), dtype=torch.int32, device=dev)
            cnt = torch.empty((plan.nchunks, plan.Qpad), dtype=torch.int32, device=dev)
        else:
            cand, cnt = out
        c0, c1 = chunks if chunks is not None else (0, 0)
        with torch.cuda.device(dev):
            check(_lib.lib().cmh_tc_topk_collect(ctypes.byref(plan), ctypes.byref(ops.c), cutoff[0].data_ptr(), cutoff[1].data_ptr(),
                                                 cap, cand.data_ptr(), cnt.data_ptr(), c0, c1, _stream()))
        return cand, cnt

    def topk_count(self, plan: Plan, cap: int, cand: torch.Tensor, cnt: torch.Tensor, k: int) -> torch.Tensor:
        """-> int32 [bins + 1, Qpad]: rows 0..bins-1 = this shard's candidates per distance, element [bins, 0] = "take the exact
        path" flag (a list overflowed / a cutoff was too tight)."""
        tot = torch.zeros((plan.bins + 1, plan.Qpad), dtype=torch.int32, device=cand.device)
        with torch.cuda.device(cand.device):
            check(_lib.lib().cmh_tc_topk_count(ctypes.byref(plan), cap, cand.data_ptr(), cnt.data_ptr(), k, tot.data_ptr(),
                                               tot[plan.bins].data_ptr(), _stream()))
        return tot

    def topk_place(self, plan: Plan, cap: int, cand: torch.Tensor, cnt: torch.Tensor, totals_all: Optional[torch.Tensor], world: int,
                   rank: int, k: int, idx_offset: int, keys: Optional[torch.Tensor], peers_dev: Optional[int] = None, npeers: int = 0,
                   multicast: Optional[int] = None, verify: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
                   ) -> Optional[torch.Tensor]:
        """``peers_dev``: device address of a table of ``npeers`` peer-mapped [Q, k] key buffers (one per rank); ``multicast``: the
        NVSwitch multicast address of those buffers.  With either, the keys go straight into every rank's buffer.
        ``verify=(sample_all, status)``: sharded verification inside the same kernel — ``status`` (int32[1]) gets bit 0 set when a
        rank overflowed a list or the candidates of all ranks are too few (``sample_all`` = the gathered sample blocks).
        ``totals_all=None`` (one shard): no ``topk_count`` needed, the kernel sums its own per-chunk counts; ``verify=(None, status)``."""
        sample_all, status = verify if verify is not None else (None, None)
        with torch.cuda.device(cand.device):
            check(_lib.lib().cmh_tc_topk_place(ctypes.byref(plan), cap, cand.data_ptr(), cnt.data_ptr(), _ptr(totals_all),
                                               (plan.bins + 1) * plan.Qpad, world, rank, k, idx_offset, _ptr(keys), peers_dev, npeers,
                                               multicast, _ptr(sample_all), _ptr(status), _stream()))
        return keys

    def topk_merge(self, parts: torch.Tensor) -> torch.Tensor:
        dev = _need_cuda(parts)
        world, Q, k = parts.shape
        out = torch.empty((Q, k), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().cmh_topk_merge(parts.data_ptr(), world, Q, k, out.data_ptr(), _stream()))
        return out


def split_keys(keys: torch.Tensor, dist_dtype=torch.int32) -> Tuple[torch.Tensor, torch.Tensor]:
    """int64 keys -> (dist int32 [or ``dist_dtype``], index int64); empty slots -> -1."""
    dev = _need_cuda(keys)
    dist = torch.empty(keys.shape, dtype=torch.int32, device=dev)
    idx = torch.empty(keys.shape, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().cmh_split_keys(keys.data_ptr(), keys.numel(), dist.data_ptr(), idx.data_ptr(), _stream()))
    return (dist if dist_dtype == torch.int32 else dist.to(dist_dtype)), idx


# ---------------------------------------------------------------------------------------------------------
# single-GPU evaluator
# ---------------------------------------------------------------------------------------------------------
@dataclass
class MapResult:
    map: torch.Tensor            # 0-dim fp64, device
    ap: torch.Tensor             # [Q] fp64, device
    tsum: torch.Tensor           # [Q] int32  R_q  (calc_utils.py:75)
    total: torch.Tensor          # [Q] int32  min(R_q, k)  (calc_utils.py:81)
    tindex: Optional[torch.Tensor] = None  # [Q, cap] int32, 0 beyond total  (calc_utils.py:88)


TOPK_STAGE_NAMES = ("expand", "hist_kernel", "scan", "rank_topk_kernel")                       # exact two-pass path
TOPK_FAST_STAGE_NAMES = ("expand", "sample_hist+cutoff", "collect_kernel", "place+verify")     # candidate path, one shard
TOPK_SHARDED_STAGE_NAMES = ("expand", "sample_hist+cutoff", "collect_kernel", "count", "gather_totals", "place+verify+key_exchange")
CAND_MIN_ITEMS = 65536   # shards smaller than this take the two-pass path directly


SAMPLE_DIV = int(os.environ.get("CMH_SAMPLE_DIV", "16"))   # tuning knob (profiles/README.md): sample = shard / SAMPLE_DIV


def candidate_sample(n_local: int) -> int:
    """Size of the gallery prefix whose exact histogram sets the per-query cutoffs: 1/16 of the shard, 4096..65536 items."""
    return max(4096, min(65536, (n_local // SAMPLE_DIV) // 512 * 512))


def candidate_cap(plan: Plan, k: int) -> int:
    """Capacity of one (chunk, query) candidate list: 8x the mean of a k-candidate query, power of two in [128, 1024]."""
    want = max(128, 8 * k // max(plan.nchunks, 1))
    cap = 128
    while cap < want and cap < 1024:
        cap *= 2
    return cap


def candidate_path_ok(st, plan: Plan, n_local: int, k: int) -> bool:
    return (bool(getattr(st, "tensor_cores", False)) and hasattr(st, "topk_collect") and n_local >= CAND_MIN_ITEMS and 16 * k <= n_local
            and (plan.nchunks + 1) * plan.bins * 4 <= 200 * 1024)   # cand_place_kernel keeps [nchunks][bins] counters per query


def collect_candidates(st, plan: Plan, ops, qp, gp, k: int, stages=None, gather=None, idx_offset: int = 0, rank: int = 0,
                       world: int = 1, count: bool = True):
    """sample histogram -> cutoff -> one tensor-core pass -> per-distance totals (+ fallback flag).
    Returns (cap, cand, cnt, tot, meta).

    ``gather`` (sharded runs): a callable that all-gathers an int32 ``[rows, Qpad]`` tensor into ``[world, rows, Qpad]``.  The
    sample blocks are exchanged first, every rank derives the SAME global cutoff / index bound (so a rank keeps ~k/world
    candidates, not k), and the "enough candidates" check is left to the caller, over all ranks (``meta`` = the gathered sample
    blocks; rank r's row ``bins`` holds [0] = its sample items, [1] = its gallery items, [2 + r] = its first gallery index)."""
    dev = qp.device
    n_s = min(candidate_sample(plan.N), plan.N)
    hist_s, plan_s = None, None
    if n_s > 0:
        plan_s = st.make_plan(plan.Q, n_s, plan.nbits, 0)
        hist_s = st.hist(plan_s, qp, None, gp[:n_s], None, ops=ops)
    meta = None
    if gather is not None:
        meta = gather(st.topk_sample_block(plan, plan_s, hist_s, idx_offset, rank, world, dev))
        cutoff = st.topk_cutoff_sharded(plan, meta, k, rank, world)
    elif hist_s is not None:
        cutoff = st.topk_cutoff(plan_s, hist_s, plan.N, k)
    else:   # an empty shard has no candidates
        cutoff = torch.full((2, plan.Qpad), -1, dtype=torch.int32, device=dev)
    _mark(stages)
    cap = candidate_cap(plan, k)
    cand, cnt = st.topk_collect(plan, ops, cutoff, cap)
    _mark(stages)
    tot = None
    if count:   # (a single shard places straight from the lists: ``topk_place(totals_all=None)``)
        tot = st.topk_count(plan, cap, cand, cnt, 0 if gather is not None else k)
        _mark(stages)
    return cap, cand, cnt, tot, meta


MAP_STAGE_NAMES = ("expand", "hist_kernel", "scan", "rank_map_kernel", "map_finish")


def _mark(stages) -> None:
    """Append a CUDA event recorded on the current stream (bench.py's per-stage timing); no-op when stages is None."""
    if stages is not None:
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        stages.append(e)


def _check_k(k: Optional[int]) -> Optional[int]:
    """k=None means the full ranking (calc_utils.py:70-71); anything else must be a positive count."""
    if k is None:
        return None
    k = int(k)
    if k <= 0:
        raise CmhError("k must be a positive integer or None (got %d)" % k)
    return k


def map_k(qp, qlp, gp, glp, nbits: int, ncls: int, k: Optional[int] = None, want_tindex: bool = False,
          tindex_cap: Optional[int] = None, target_blocks: int = 0, stages: Optional[list] = None,
          tensor_cores: Optional[bool] = None) -> MapResult:
    """calc_map_k on packed inputs, one GPU: hist -> scan -> rank/AP -> mean (MAP_STAGE_NAMES).  Scratch comes from torch's
    stream-aware caching allocator, so evaluations on different streams do not share buffers."""
    dev = _need_cuda(qp, qlp, gp, glp)
    k = _check_k(k)
    Q, N = qp.shape[0], gp.shape[0]
    st = CudaStages(tensor_cores)
    plan = st.make_plan(Q, N, nbits, ncls, None, target_blocks)
    tindex = None
    if want_tindex:
        cap = int(tindex_cap if tindex_cap is not None else (min(k, N) if k else N))
        tindex = torch.zeros((Q, max(cap, 1)), dtype=torch.int32, device=dev)
    ops = st.operands(plan, qp, qlp, gp, glp)
    _mark(stages)
    hist = st.hist(plan, qp, qlp, gp, glp, ops=ops)
    _mark(stages)
    sc = st.scan(plan, hist, 1, 0, k)
    _mark(stages)
    ap_partial = st.rank_map(plan, qp, qlp, gp, glp, sc, tindex, ops=ops)
    _mark(stages)
    ap, m = st.map_finish(plan, ap_partial, sc["total"])
    _mark(stages)
    return MapResult(m, ap, sc["tsum"][:Q], sc["total"][:Q], tindex)


def topk(qp, gp, nbits: int, k: int, idx_offset: int = 0, target_blocks: int = 0, stages: Optional[list] = None,
         out: Optional[torch.Tensor] = None, tensor_cores: Optional[bool] = None, exact: Optional[bool] = None) -> torch.Tensor:
    """First k entries of the stable Hamming ranking as int64 keys ``(dist << 32) | index`` [Q, k].

    Large galleries take the candidate path (TOPK_FAST_STAGE_NAMES): per-query cutoff from the exact histogram of a gallery prefix,
    one tensor-core pass that keeps only items within the cutoff, per-query placement — verified on the device (enough
    candidates, no list overflow); if the check fails, or with ``exact=True``, the two-pass counting path (TOPK_STAGE_NAMES) runs.
    Both give the same keys."""
    dev = _need_cuda(qp, gp, out)
    k = _check_k(k)
    if k is None:
        raise CmhError("top-k needs k")
    Q, N = qp.shape[0], gp.shape[0]
    st = CudaStages(tensor_cores)
    plan = st.make_plan(Q, N, nbits, 0, None, target_blocks)
    keys = out if out is not None else torch.empty((Q, k), dtype=torch.int64, device=dev)
    if keys.shape != (Q, k) or keys.dtype != torch.int64 or not keys.is_contiguous():
        raise CmhError("out must be a contiguous int64 [Q, k] tensor")
    if k > N:
        keys.fill_(EMPTY_KEY)
    ops = st.operands(plan, qp, None, gp, None)
    _mark(stages)
    if exact is not True and candidate_path_ok(st, plan, N, k):
        cap, cand, cnt, _, _ = collect_candidates(st, plan, ops, qp, gp, k, stages, count=False)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        st.topk_place(plan, cap, cand, cnt, None, 1, 0, k, idx_offset, keys, verify=(None, status))
        _mark(stages)
        if int(status.item()) == 0:                # verified: every cutoff was wide enough and no list overflowed
            return keys
        if stages is not None:
            del stages[1:]                          # the exact path below re-times its own stages
            _mark(stages)
    hist = st.hist(plan, qp, None, gp, None, ops=ops)
    _mark(stages)
    sc = st.scan(plan, hist, 1, 0, k, with_rel=False)
    _mark(stages)
    st.rank_topk(plan, qp, gp, sc, k, idx_offset, keys=keys, ops=ops)
    _mark(stages)
    return keys


def topk_from_host(q_host: torch.Tensor, g_host: torch.Tensor, k: int, device, slabs: int = 6, bad: Optional[torch.Tensor] = None
                   ) -> torch.Tensor:
    """``topk`` for +-1 fp32 code matrices that still live in (ideally pinned) HOST memory: the gallery crosses PCIe in slabs of
    whole chunks on a copy stream while the slabs that have landed are packed, expanded and COLLECTED on the compute stream, so
    the transfer (the longer of the two by far: 256 MB for 1 M x 64 bit) hides the compute instead of preceding it.
    Same keys as ``topk(pack_codes(q), pack_codes(g), ...)``; falls back to that call when the candidate path does not apply."""
    dev = torch.device(device)
    Q, nbits = q_host.shape
    N = g_host.shape[0]
    k = _check_k(k)
    st = CudaStages()
    plan = st.make_plan(Q, N, nbits, 0)
    main = torch.cuda.current_stream(dev)
    if bad is None:
        bad = new_bad_counter(dev)
    with torch.cuda.device(dev):
        qp = pack_codes(q_host.to(dev, non_blocking=True), bad)
        per_slab = -(-plan.nchunks // max(1, slabs))
        n_s = min(candidate_sample(N), N)
        # the first slab must hold the sample prefix
        first = max(per_slab, -(-n_s // plan.chunk_items))
        if not candidate_path_ok(st, plan, N, k) or first >= plan.nchunks or g_host.dtype != torch.float32 or g_host.stride(1) != 1:
            return topk(qp, pack_codes(g_host.to(dev, non_blocking=True), bad), nbits, k)
        gp = torch.empty((N, code_words(nbits)), dtype=torch.int32, device=dev)
        g_ops = torch.empty((N, _lib.lib().cmh_tc_operand_bytes(nbits)), dtype=torch.int8, device=dev)
        ops = Operands(_expand(qp, plan.Qpad, nbits, 0), None, g_ops, None)
        bounds, c = [], 0
        while c < plan.nchunks:
            c1 = min(plan.nchunks, c + (first if c == 0 else per_slab))
            bounds.append((c, c1))
            c = c1
        copy = torch.cuda.Stream(dev)
        max_items = max((c1 - c0) for c0, c1 in bounds) * plan.chunk_items
        stage = [torch.empty((max_items, nbits), dtype=torch.float32, device=dev) for _ in range(2)]
        free = [None, None]
        cap = candidate_cap(plan, k)
        cand = cnt = cutoff = None
        for j, (c0, c1) in enumerate(bounds):
            lo, hi = c0 * plan.chunk_items, min(N, c1 * plan.chunk_items)
            buf = stage[j & 1][: hi - lo]
            if free[j & 1] is not None:
                copy.wait_event(free[j & 1])
            else:
                copy.wait_stream(main)               # the staging buffers were allocated on the compute stream
            with torch.cuda.stream(copy):
                buf.copy_(g_host[lo:hi], non_blocking=True)
                landed = torch.cuda.Event()
                landed.record(copy)
            main.wait_event(landed)
            pack_codes(buf, bad, out=gp[lo:hi])
            free[j & 1] = torch.cuda.Event()
            free[j & 1].record(main)
            check(_lib.lib().cmh_tc_expand(gp[lo:hi].data_ptr(), hi - lo, hi - lo, gp.shape[1], nbits, 0, g_ops[lo:hi].data_ptr(), _stream()))
            if j == 0:   # the sample prefix is on the device: cutoffs
                plan_s = st.make_plan(Q, n_s, nbits, 0)
                cutoff = st.topk_cutoff(plan_s, st.hist(plan_s, qp, None, gp[:n_s], None, ops=ops), N, k)
            cand, cnt = st.topk_collect(plan, ops, cutoff, cap, out=None if cand is None else (cand, cnt), chunks=(c0, c1))
        keys = torch.empty((Q, k), dtype=torch.int64, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        st.topk_place(plan, cap, cand, cnt, None, 1, 0, k, 0, keys, verify=(None, status))
        if int(status.item()) == 0:
            return keys
        return topk(qp, gp, nbits, k, exact=True)     # verified-failed: exact two-pass path on the codes that are now resident


# ---------------------------------------------------------------------------------------------------------
# sharded evaluator (gallery split by contiguous index range over the ranks of a process group)
# ---------------------------------------------------------------------------------------------------------
def shard_bounds(n: int, world: int, align: int = 4) -> List[Tuple[int, int]]:
    """Contiguous [lo, hi) gallery ranges; shard sizes are multiples of ``align`` (16-byte bulk copies)."""
    per = -(-n // world)
    per = -(-per // align) * align
    return [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)]


TOPK_EXCHANGES = ("auto", "nvls", "nvls_reduce", "peer_stores", "rank_scatter", "allgather_merge")


class ShardedEvaluator:
    """mAP / top-k with the gallery sharded over ``group``; queries (and their labels) are replicated.

    Exchange steps:
      mAP    : per-shard bucket totals [2, bins, Qpad] int32 (all-gather)  ->  scan  ->  ONE fp64 AP partial per query (all-gather)
      top-k  : candidate path (default, large shards): sample blocks gathered -> ONE global cutoff per query -> collect / count ->
               candidate totals gathered -> place + verify -> every rank pushes the [Q, k] slots it owns to all ranks.  On a box
               with NVSwitch multicast memory all three exchanges are store kernels on the multicast address between device-side
               barriers (`cmh_exchange.cu`), otherwise all_gather_into_tensor / all_reduce(MAX) (NCCL; gloo in the CPU tests)
               exact two-pass path (small shards, failed verification, ``exact=True``): per-shard bucket totals -> every item's
               global rank -> ONE all-reduce(MAX) of the [Q, k] key buffer (or, method="allgather_merge": per-shard partial top-k
               keys, ONE all-gather, merge kernel)
    Every rank ends with the identical result.
    """

    def __init__(self, group=None, stages=None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.stages = stages if stages is not None else CudaStages()
        self._symm = {}          # (tag, shape, dtype, device) -> (tensor, handle): symmetric buffers of the fused exchanges
        self._symm_broken = None  # reason symmetric memory is unusable in this process group, once known

    # ---- fused exchange: buffers of every rank mapped into every other rank and at one multicast address (NVLink / NVSwitch) ----
    def _symmetric(self, tag: str, shape, dtype, device):
        """Symmetric-memory buffer + its rendezvous handle (cached per tag and shape); None when this build / box / group cannot
        map peer memory (then the NCCL exchanges are used).  Collective: every rank must call it with the same arguments."""
        if self._symm_broken is not None or device.type != "cuda":
            return None
        key = (tag, tuple(shape), dtype, device.index)
        if key not in self._symm:
            buf, err = None, None
            try:
                import torch.distributed._symmetric_memory as symm_mem

                buf = symm_mem.empty(tuple(shape), dtype=dtype, device=device)
            except Exception as e:
                err = "%s: %s" % (type(e).__name__, e)
            # the allocation is local, the rendezvous is collective: agree first so that no rank waits for one that gave up
            ok = torch.tensor([0 if buf is None else 1], dtype=torch.int32, device=device)
            self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                self._symm_broken = err or "symmetric memory allocation failed on another rank"
                return None
            try:
                grp = self.group if self.group is not None else self.dist.group.WORLD
                hdl = symm_mem.rendezvous(buf, grp)
                if len(hdl.buffer_ptrs) != self.world:
                    raise RuntimeError("rendezvous returned %d buffers for %d ranks" % (len(hdl.buffer_ptrs), self.world))
                if not (getattr(hdl, "has_multicast_support", False) and int(hdl.multicast_ptr)):
                    raise RuntimeError("no NVSwitch multicast mapping for the buffer (NVLS unavailable)")
                self._symm[key] = (buf, hdl)
            except Exception as e:  # not supported here: remember why, use the NCCL exchange from now on
                self._symm_broken = "%s: %s" % (type(e).__name__, e)
                return None
        return self._symm[key]

    def _symmetric_keys(self, Q: int, k: int, device):
        return self._symmetric("keys", (Q, k), torch.int64, device)

    def _gather_small(self, tag: str, t: torch.Tensor, nvls: bool, pre_barrier: bool = True) -> torch.Tensor:
        """All-gather of a small per-rank block ([rows, Qpad] int32, ~2.7 MB) -> [world, rows, Qpad].  With NVSwitch multicast every
        rank broadcasts its block into slot ``rank`` of a symmetric buffer with one store kernel between two device-side barriers
        (peers are done reading the previous contents / all blocks have landed); the result aliases that buffer and stays valid
        until the next gather with the same tag.  ``pre_barrier=False``: the caller guarantees that every rank passes another
        barrier of the group between its last read of this buffer and the next gather (the sharded top-k does: the barriers of
        the key exchange follow the kernels that read the gathered blocks).  Otherwise (gloo, no multicast): NCCL / gloo all-gather."""
        symm = None
        if nvls and t.is_cuda and (t.numel() * t.element_size()) % 16 == 0:
            symm = self._symmetric(tag, (self.world,) + tuple(t.shape), t.dtype, t.device)
        if symm is None:
            return self._gather(t)
        buf, hdl = symm
        t = t.contiguous()
        nbytes = t.numel() * t.element_size()
        if pre_barrier:
            hdl.barrier(channel=0)
        with torch.cuda.device(t.device):
            check(_lib.lib().cmh_nvls_broadcast(t.data_ptr(), int(hdl.multicast_ptr) + self.rank * nbytes, nbytes, _stream()))
        hdl.barrier(channel=1)
        return buf

    def exchange_info(self) -> Dict[str, object]:
        """What the top-k exchange uses in this process group (for logs / bench.py)."""
        if self._symm_broken is not None:
            return {"peer_memory": False, "reason": self._symm_broken[:200]}
        if not self._symm:
            return {"peer_memory": None}
        hdl = next(iter(self._symm.values()))[1]
        mc = bool(getattr(hdl, "has_multicast_support", False)) and bool(getattr(hdl, "multicast_ptr", 0))
        return {"peer_memory": True, "multicast": mc}

    def _gather(self, t: torch.Tensor) -> torch.Tensor:
        t = t.contiguous()
        flat = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(flat, t, group=self.group)   # rank-major concatenation along dim 0
        return flat.view((self.world,) + tuple(t.shape))

    def _geometry(self, n_local: int, n_geom: Optional[int], device) -> int:
        if n_geom is not None:
            return n_geom
        t = torch.tensor([n_local], dtype=torch.int64, device=device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return int(t.item())

    def map_k(self, qp, qlp, gp_local, glp_local, nbits: int, ncls: int, k: Optional[int] = None,
              n_geom: Optional[int] = None, tindex_cap: Optional[int] = None) -> MapResult:
        st = self.stages
        Q, n_local = qp.shape[0], gp_local.shape[0]
        n_geom = self._geometry(n_local, n_geom, qp.device)
        plan = st.make_plan(Q, n_local, nbits, ncls, n_geom)
        ops = st.operands(plan, qp, qlp, gp_local, glp_local) if hasattr(st, "operands") else None
        kw = {"ops": ops} if ops is not None else {}
        hist = st.hist(plan, qp, qlp, gp_local, glp_local, **kw)
        # a rank's chunks follow all chunks of the lower ranks: only the per-bucket totals travel
        totals_all = self._gather(st.hist_totals(plan, hist))  # [world, 2, bins, Qpad]
        sc = st.scan_sharded(plan, hist, totals_all, self.world, self.rank, k)
        tindex = None
        if tindex_cap:
            tindex = torch.zeros((Q, tindex_cap), dtype=torch.int32, device=qp.device)
        ap_partial = st.rank_map(plan, qp, qlp, gp_local, glp_local, sc, tindex, n_total=n_geom * self.world, **kw)
        ap_all = self._gather(st.ap_reduce(plan, ap_partial))  # [world, 1, Qpad]: one fp64 per query and rank
        ap, m = st.map_finish(plan, ap_all, sc["total"])
        if tindex is not None:                                 # each slot is written by exactly one rank
            self.dist.all_reduce(tindex, op=self.dist.ReduceOp.SUM, group=self.group)
        return MapResult(m, ap, sc["tsum"][:Q], sc["total"][:Q], tindex)

    def candidate_path(self, plan: Plan, ops, n_geom: int, k: int, method: str = "auto") -> bool:
        """Whether the single-pass candidate path applies (decided on the common geometry: same answer on every rank)."""
        st = self.stages
        return (method != "allgather_merge" and ops is not None and candidate_path_ok(st, plan, n_geom, k)
                and self.world + 2 <= plan.Qpad)

    def _topk_candidates(self, plan: Plan, ops, qp, gp_local, k: int, idx_offset: int, n_geom: int, method: str, copy: bool,
                         stages: Optional[list]):
        """The candidate path of ``topk`` without any host synchronisation (it is also what ``TopkGraph`` captures): returns
        ``(keys, bad)``; ``bad`` is a device int32[1], non-zero when a candidate list overflowed or the candidates of all ranks
        together were too few for some query — and the caller must then take the exact path.  Local cutoffs come from ONE global
        sample, so the candidates of all ranks are a prefix of the global (distance, index) order."""
        st = self.stages
        Q = plan.Q
        fused = method in ("auto", "nvls", "nvls_reduce", "peer_stores")
        small_nvls = fused and os.environ.get("CMH_SMALL_EXCHANGE", "nvls") != "nccl"
        symm = self._symmetric_keys(Q, k, qp.device) if fused else None
        if method in ("nvls", "nvls_reduce", "peer_stores") and symm is None:
            raise CmhError("%s exchange is not available: %s" % (method, self._symm_broken))
        # The gathered blocks are read by the cutoff / place kernels; every fused key exchange below puts a barrier of the group
        # after the place kernel, so the next step's broadcasts cannot overtake those reads: no barrier before the broadcasts.
        pre = symm is None
        cap, cand, cnt, tot, meta = collect_candidates(
            st, plan, ops, qp, gp_local, k, stages, idx_offset=idx_offset, rank=self.rank, world=self.world,
            gather=lambda t: self._gather_small("sample", t, small_nvls, pre))
        tot_all = self._gather_small("totals", tot, small_nvls, pre)  # [world, bins + 1, Qpad]: per-distance totals + overflow flag
        # verified inside the place kernel, on every rank from the same gathered data: no list overflowed, and the candidates of ALL
        # ranks together (a prefix of the global order) number at least min(k, gallery size) for every query
        bad = torch.zeros(1, dtype=torch.int32, device=qp.device)
        verify = (meta, bad)
        _mark(stages)
        if symm is not None and method == "peer_stores":
            # fused place + exchange: a key's global slot is known, so the place kernel stores it straight into that slot of
            # every rank's buffer over NVLink.  Correct, but scattered 8-byte remote stores run at a fraction of the link rate
            # (2 GPUs: 0.69 ms against 0.16 ms local placement + 0.1 ms NVLS reduction, profiles/README.md): not the default.
            buf, hdl = symm
            if k > n_geom * self.world:
                buf.fill_(EMPTY_KEY)
            hdl.barrier(channel=0)
            st.topk_place(plan, cap, cand, cnt, tot_all, self.world, self.rank, k, idx_offset, None,
                          peers_dev=int(hdl.buffer_ptrs_dev), npeers=self.world, verify=verify)
            hdl.barrier(channel=1)
            keys = buf.clone() if copy else buf
            _mark(stages)
        elif symm is not None and method == "nvls_reduce":
            # The keys are placed into this rank's symmetric buffer (slots owned by other ranks stay EMPTY); then ONE kernel per
            # rank reduces 1/world of the buffer in the switch (multimem.ld_reduce max) and broadcasts it (multimem.st).
            # Barriers: peers must be done with the previous result before it is overwritten; all fills + placements must be
            # visible before the reduction; all broadcasts must have landed before anyone reads.
            buf, hdl = symm
            hdl.barrier(channel=0)
            buf.fill_(EMPTY_KEY)
            st.topk_place(plan, cap, cand, cnt, tot_all, self.world, self.rank, k, idx_offset, buf, verify=verify)
            hdl.barrier(channel=1)
            with torch.cuda.device(qp.device):
                check(_lib.lib().cmh_nvls_allreduce_max_s64(int(hdl.multicast_ptr), buf.numel(), self.rank, self.world, _stream()))
            hdl.barrier(channel=2)
            keys = buf.clone() if copy else buf
            _mark(stages)
        elif symm is not None:
            # Default.  Every slot has exactly one owner: the keys are placed into this rank's symmetric buffer (other slots
            # EMPTY), then each rank PUSHES the slots it owns to all ranks with coalesced multicast stores — no reduction, no
            # round trip through the switch.  Only this rank reads its buffer, so no barrier is needed before the fill; a slot
            # a peer's push already overwrote holds the final key and is pushed again unchanged.  Barriers: every rank has
            # filled + placed before any push lands; all pushes have landed before anyone reads.
            buf, hdl = symm
            buf.fill_(EMPTY_KEY)
            st.topk_place(plan, cap, cand, cnt, tot_all, self.world, self.rank, k, idx_offset, buf, verify=verify)
            hdl.barrier(channel=0)
            with torch.cuda.device(qp.device):
                check(_lib.lib().cmh_nvls_push_owned_s64(buf.data_ptr(), int(hdl.multicast_ptr), buf.numel(), _stream()))
            hdl.barrier(channel=1)
            keys = buf.clone() if copy else buf
            _mark(stages)
        else:
            keys = torch.full((Q, k), EMPTY_KEY, dtype=torch.int64, device=qp.device)
            st.topk_place(plan, cap, cand, cnt, tot_all, self.world, self.rank, k, idx_offset, keys, verify=verify)
            self._exchange_keys(keys, method)
        return keys, bad

    def topk(self, qp, gp_local, nbits: int, k: int, idx_offset: int, n_geom: Optional[int] = None,
             method: str = "auto", exact: Optional[bool] = None, copy: bool = True, stages: Optional[list] = None) -> torch.Tensor:
        """Global top-k keys [Q, k] (identical on every rank) of a gallery sharded by contiguous index range.

        ``method`` names the exchange of the result keys (TOPK_EXCHANGES):
        ``auto`` / ``nvls``: candidate path with the multicast exchanges (`_topk_candidates`); ``auto`` falls back to the NCCL
        exchanges when the group has no multicast memory, ``nvls`` raises.  ``nvls_reduce`` / ``peer_stores``: alternative fused key
        exchanges (reduction inside the switch / per-key stores into every peer), kept for comparison.
        ``rank_scatter``: NCCL all-reduce(MAX) of the [Q, k] buffer in which every rank filled the slots it owns — the counting
        formulation gives every item its GLOBAL stable rank from this rank's own counts plus the other ranks' per-bucket totals, so
        there are no partial lists and no merge.  ``allgather_merge``: BASELINE.json's literal exchange — per-shard partial top-k,
        one all-gather of [Q, k] keys per rank (8 x 80 MB at C4), merge kernel; always the exact two-pass path.
        All are exact and give the same keys (tests/test_sharded_gloo.py, scripts/check_multi_gpu.py).  ``copy=False`` returns the
        symmetric result buffer itself (valid until the next call).
        """
        st = self.stages
        k = _check_k(k)
        if method not in TOPK_EXCHANGES:
            raise CmhError("unknown top-k exchange %r" % method)
        Q, n_local = qp.shape[0], gp_local.shape[0]
        n_geom = self._geometry(n_local, n_geom, qp.device)
        plan = st.make_plan(Q, n_local, nbits, 0, n_geom)
        ops = st.operands(plan, qp, None, gp_local, None) if hasattr(st, "operands") else None
        _mark(stages)
        kw = {"ops": ops} if ops is not None else {}
        if exact is not True and self.candidate_path(plan, ops, n_geom, k, method):
            keys, bad = self._topk_candidates(plan, ops, qp, gp_local, k, idx_offset, n_geom, method, copy, stages)
            if not bool(bad.item()):   # same answer on every rank
                return keys
        hist = st.hist(plan, qp, None, gp_local, None, **kw)
        if method in ("auto", "nvls", "nvls_reduce", "peer_stores"):
            method = "rank_scatter"                                 # the exact two-pass path exchanges through NCCL
        if method == "allgather_merge":
            sc = st.scan(plan, hist, 1, 0, k, with_rel=False)      # local ranking of this shard
            keys = st.rank_topk(plan, qp, gp_local, sc, k, idx_offset, **kw)
            parts = self._gather(keys)                             # [world, Q, k]
            return st.topk_merge(parts)
        totals_all = self._gather(st.hist_totals(plan, hist))      # [world, 2, bins, Qpad]
        sc = st.scan_sharded(plan, hist, totals_all, self.world, self.rank, k)
        keys = st.rank_topk(plan, qp, gp_local, sc, k, idx_offset, **kw)  # slots of global rank < k owned by this shard
        self._exchange_keys(keys, method)
        return keys

    def _exchange_keys(self, keys: torch.Tensor, method: str) -> None:
        """Every slot of the [Q, k] key buffer is owned by exactly one rank (the others hold EMPTY = -1 < every real key)."""
        self.dist.all_reduce(keys, op=self.dist.ReduceOp.MAX, group=self.group)


# ---------------------------------------------------------------------------------------------------------
# the whole top-k step as ONE CUDA graph
# ---------------------------------------------------------------------------------------------------------
class TopkGraph:
    """pack -> expand -> sample histogram -> cutoff -> collect -> count -> (exchanges) -> place, captured once as a CUDA graph over
    static buffers and replayed per evaluation.

    Why: the step is ~20 launches of 10-900 us.  Queued one by one through Python the host needs 1-2 ms per step — at 8 GPUs that
    is longer than the GPU work (0.3 ms of collect per shard), every rank's launches trail its kernels, and each exchange turns
    into a wait for the slowest HOST.  Replayed as a graph the step costs one launch; with the NVSwitch multicast exchanges
    (`_gather_small`, push-owned keys) it contains no NCCL call at all, only kernels and device-side barriers.

    ``q_codes`` [Q, K] / ``g_codes`` [N_local, K]: +-1 fp32 CUDA tensors that STAY the inputs of every replay — refresh them in
    place (``copy_``) between runs.  ``evaluator``: a ShardedEvaluator when the gallery is sharded over a process group
    (``idx_offset`` = global index of this shard's first item, ``n_geom`` = the largest shard); every rank must build and run the
    graph together.  ``run()`` returns the int64 ``(dist << 32) | index`` keys [Q, k] — a static buffer, overwritten by the next
    run; if the device-side verification fails (cutoff too tight for some query, a candidate list overflowed) it falls back to
    the exact two-pass path, and it raises ValueError for codes that are not +-1, like ``calc_utils.hamming_topk``."""

    def __init__(self, q_codes: torch.Tensor, g_codes: torch.Tensor, k: int, evaluator: Optional["ShardedEvaluator"] = None,
                 idx_offset: int = 0, n_geom: Optional[int] = None, method: str = "auto"):
        dev = _need_cuda(q_codes, g_codes)
        if q_codes.dtype != torch.float32 or g_codes.dtype != torch.float32 or not (q_codes.is_contiguous() and g_codes.is_contiguous()):
            raise CmhError("TopkGraph: static code buffers must be contiguous fp32")
        self.q, self.g, self.k, self.ev = q_codes, g_codes, _check_k(k), evaluator
        self.idx_offset, self.method = int(idx_offset), method
        self.nbits = q_codes.shape[1]
        self.st = evaluator.stages if evaluator is not None else CudaStages()
        Q, n_local = q_codes.shape[0], g_codes.shape[0]
        if evaluator is not None:
            if method not in ("auto", "nvls", "nvls_reduce", "peer_stores", "rank_scatter"):
                raise CmhError("TopkGraph: exchange %r has no single-pass form" % method)
            self.n_geom = evaluator._geometry(n_local, n_geom, dev)
        else:
            self.n_geom = n_local
        self.plan = self.st.make_plan(Q, n_local, self.nbits, 0, self.n_geom if evaluator is not None else None)
        if not candidate_path_ok(self.st, self.plan, self.n_geom, self.k) or (evaluator is not None and evaluator.world + 2 > self.plan.Qpad):
            raise CmhError("TopkGraph: the single-pass candidate path does not apply to this shape (use topk)")
        self.status = torch.zeros(2, dtype=torch.int64, device=dev)     # [0] path failed, [1] elements that are not +-1
        self._bad_codes = new_bad_counter(dev)
        self.keys = None
        with torch.cuda.device(dev):
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):   # eager warm-up on the capture stream: symmetric buffers, kernel attributes, allocator pool
                    self._body()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            l0 = int(_lib.lib().cmh_launch_count())
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self._body()
            self.kernels = int(_lib.lib().cmh_launch_count()) - l0       # launches of this library inside one replay
        self._pin = torch.empty(2, dtype=torch.int64).pin_memory()

    def _body(self) -> None:
        st, plan, k = self.st, self.plan, self.k
        self._bad_codes.zero_()
        qp = pack_codes(self.q, self._bad_codes)
        gp = pack_codes(self.g, self._bad_codes)
        ops = st.operands(plan, qp, None, gp, None)
        if self.ev is None:
            cap, cand, cnt, _, _ = collect_candidates(st, plan, ops, qp, gp, k, count=False)
            if self.keys is None:
                self.keys = torch.empty((plan.Q, k), dtype=torch.int64, device=qp.device)
            failed = torch.zeros(1, dtype=torch.int32, device=qp.device)
            st.topk_place(plan, cap, cand, cnt, None, 1, 0, k, self.idx_offset, self.keys, verify=(None, failed))
        else:
            self.keys, failed = self.ev._topk_candidates(plan, ops, qp, gp, k, self.idx_offset, self.n_geom, self.method, False, None)
        self.status[0:1].copy_(failed.reshape(1))
        self.status[1:2].copy_(self._bad_codes)

    def replay(self) -> torch.Tensor:
        """Queue one step on the current stream; no host synchronisation, no verification (``run`` does both)."""
        self.graph.replay()
        return self.keys

    def run(self) -> torch.Tensor:
        self.graph.replay()
        return self.finish()

    def finish(self) -> torch.Tensor:
        """Verification of the last replay: reads the two status words (the one host synchronisation of a step)."""
        self._pin.copy_(self.status, non_blocking=True)
        torch.cuda.current_stream(self.q.device).synchronize()
        failed, bad = int(self._pin[0]), int(self._pin[1])
        if bad:
            raise ValueError("TopkGraph: codes must be +-1")
        if not failed:
            return self.keys
        qp, gp = pack_codes(self.q), pack_codes(self.g)               # same decision on every rank (gathered totals)
        if self.ev is None:
            return topk(qp, gp, self.nbits, self.k, self.idx_offset, exact=True)
        return self.ev.topk(qp, gp, self.nbits, self.k, self.idx_offset, self.n_geom, method=self.method, exact=True)
