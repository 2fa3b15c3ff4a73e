"""Evaluation half of the reference's ``BaseTrainer`` (runners/base.py:242-266, 307-357, 386-405) on packed codes.

    reference (runners/base.py)                 here
    get_code(data_loader, length)   :242-266    models.get_code            packed int32 [length, W] buffers, one byte-MAX all-reduce
    valid(epoch, k)                 :307-339    PackedEvaluation.valid     4 mAPs with labels packed once, best-epoch bookkeeping,
                                                                           .mat dumps and the save_model callback as in the reference
    test()                          :341-357    PackedEvaluation.test
    save_mat(...)                   :386-405    save_mat                   packed codes -> the reference's +-1 float arrays -> scipy savemat

A trainer that keeps the reference's training loop swaps its ``valid`` / ``test`` for these; nothing here trains.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import torch

from . import calc_utils, models, retrieval as R


def _as_float_codes(codes, nbits: Optional[int]):
    """Packed int32 [n, W] device codes -> the reference's +-1 fp32 [n, K] numpy array; float codes pass through."""
    if isinstance(codes, torch.Tensor) and codes.dtype == torch.int32 and nbits is not None:
        return R.unpack_codes(codes, nbits).cpu().numpy()
    if isinstance(codes, torch.Tensor):
        return codes.detach().cpu().numpy()
    return codes


def save_mat(query_img, query_txt, query_labels, retrieval_img, retrieval_txt, retrieval_labels, save_file="i2t", nbits: Optional[int] = None):
    """``BaseTrainer.save_mat`` (runners/base.py:386-405): same keys (q_img, q_txt, r_img, r_txt, q_l, r_l), same +-1 float code
    matrices — unpacked on the GPU (``cmh_unpack_codes_f32``) when the codes come bit-packed from ``models.get_code``."""
    import scipy.io as scio

    def lab(t):
        return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t

    scio.savemat(os.path.join(save_file), {
        "q_img": _as_float_codes(query_img, nbits), "q_txt": _as_float_codes(query_txt, nbits),
        "r_img": _as_float_codes(retrieval_img, nbits), "r_txt": _as_float_codes(retrieval_txt, nbits),
        "q_l": lab(query_labels), "r_l": lab(retrieval_labels)})


class PackedEvaluation:
    """The state ``BaseTrainer`` keeps across validations (runners/base.py:60-63: max_mapi2t, max_mapt2i, best_epoch_i,
    best_epoch_t) and its ``valid`` / ``test`` bodies, on packed codes."""

    def __init__(self, model, query_loader, retrieval_loader, query_labels, retrieval_labels, query_num: int, retrieval_num: int,
                 save_dir: Optional[str] = None, epochs: int = 0, distributed: bool = False, rank: int = 0, group=None,
                 save_model: Optional[Callable[[str, int], None]] = None, logger=None, top_k: Optional[int] = None):
        self.model, self.query_loader, self.retrieval_loader = model, query_loader, retrieval_loader
        self.query_labels, self.retrieval_labels = query_labels, retrieval_labels
        self.query_num, self.retrieval_num = query_num, retrieval_num
        self.save_dir, self.epochs, self.distributed, self.rank, self.group = save_dir, epochs, distributed, rank, group
        self.save_model, self.logger, self.top_k = save_model, logger, top_k
        self.max_mapi2t = self.max_mapt2i = 0.0
        self.best_epoch_i = self.best_epoch_t = 0

    def _codes(self):
        qi, qt = models.get_code(self.model, self.query_loader, self.query_num, distributed=self.distributed, group=self.group)
        ri, rt = models.get_code(self.model, self.retrieval_loader, self.retrieval_num, distributed=self.distributed, group=self.group)
        return qi, qt, ri, rt

    def _maps(self, codes, k):
        qi, qt, ri, rt = codes
        return calc_utils.valid_packed(qi, qt, ri, rt, self.query_labels, self.retrieval_labels, self.model.output_dim, k)

    def _dump(self, codes, name):
        if self.save_dir is None or (self.distributed and self.rank != 0):
            return
        d = os.path.join(self.save_dir, "mat_files")
        os.makedirs(d, exist_ok=True)
        qi, qt, ri, rt = codes
        save_mat(qi, qt, self.query_labels, ri, rt, self.retrieval_labels, save_file=os.path.join(d, name), nbits=self.model.output_dim)

    def valid(self, epoch: int, k: Optional[int] = None):
        """runners/base.py:307-339 -> (mAPi2t, mAPt2i, mAPi2i, mAPt2t) as 0-dim fp32 CPU tensors."""
        codes = self._codes()
        mAPi2t, mAPt2i, mAPi2i, mAPt2t = self._maps(codes, k)
        if self.max_mapi2t < mAPi2t:
            self.best_epoch_i = epoch
            self._dump(codes, "i2t-best.mat")
            if self.save_model is not None and not (self.distributed and self.rank != 0):
                self.save_model(self.save_dir, epoch)
        self.max_mapi2t = max(self.max_mapi2t, mAPi2t)
        if self.max_mapt2i < mAPt2i:
            self.best_epoch_t = epoch
            self._dump(codes, "t2i-best.mat")
            if self.save_model is not None and not (self.distributed and self.rank != 0):
                self.save_model(self.save_dir, epoch)
        self.max_mapt2i = max(self.max_mapt2i, mAPt2i)
        self._dump(codes, "last.mat")
        if self.logger is not None:
            self.logger.info(f">>>>>> [{epoch}/{self.epochs}], MAP(i->t): {mAPi2t}, MAP(t->i): {mAPt2i}, MAP(t->t): {mAPt2t}, "
                             f"MAP(i->i): {mAPi2i}, MAX MAP(i->t): {self.max_mapi2t}, epoch: {self.best_epoch_i}, "
                             f"MAX MAP(t->i): {self.max_mapt2i}, epoch: {self.best_epoch_t}")
        return mAPi2t, mAPt2i, mAPi2i, mAPt2t

    def test(self):
        """runners/base.py:341-357 (mAP@top_k, test.mat)."""
        codes = self._codes()
        maps = self._maps(codes, self.top_k)
        self._dump(codes, "test.mat")
        if self.logger is not None:
            self.logger.info(f">>>>>> TEST, MAP(i->t): {maps[0]}, MAP(t->i): {maps[1]}, MAP(t->t): {maps[3]}, MAP(i->i): {maps[2]}")
        return maps
