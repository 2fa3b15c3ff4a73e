"""The tail of the DSPH training step on the GPU (BASELINE.json config C5): fused optimisers, the HyP
objective's gradient, the tanh(Linear) hash head's backward, and a frozen-backbone training step built from them.

    reference                                               here
    models/common/optimizer.py:52-165  BertAdam             FusedBertAdam   same constructor / get_lr / step semantics; ONE pair of
                                                                            kernel launches per step for all tensors (csrc/cmh_train.cu)
    runners/DSPH/runner.py:86-89       torch.optim.SGD      FusedSGD        momentum + weight decay (the HyP proxies' optimiser)
    models/DSPH/loss/HyP.py:18-69      HyP.forward+autograd HypLoss / hyp_loss_and_grad
    runners/DSPH/runner.py:104-127     train step           DsphHeadTrainer (backbone frozen: its backward is not built)

``FusedBertAdam`` is a drop-in for the reference's ``BertAdam`` inside an unmodified PyTorch trainer: it only needs fp32 CUDA
parameters with ``.grad``.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional

import torch

from . import _lib
from . import retrieval as R


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# models/common/optimizer.py:25-49
def warmup_cosine(x, warmup=0.002):
    return x / warmup if x < warmup else 0.5 * (1.0 + math.cos(math.pi * x))


def warmup_constant(x, warmup=0.002):
    return x / warmup if x < warmup else 1.0


def warmup_linear(x, warmup=0.002):
    return x / warmup if x < warmup else max((x - 1.0) / (warmup - 1.0), 0)


SCHEDULES = {"warmup_cosine": warmup_cosine, "warmup_constant": warmup_constant, "warmup_linear": warmup_linear}


class OptTensor(ctypes.Structure):
    """Mirror of ``struct cmh_opt_tensor`` (include/cmh.h)."""

    _fields_ = [("param", ctypes.c_void_p), ("grad", ctypes.c_void_p), ("m", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("n", ctypes.c_int64), ("lr", ctypes.c_float), ("weight_decay", ctypes.c_float)]


class _MultiTensor:
    """Device-side tables of one multi-tensor launch: the cmh_opt_tensor array and the block -> (tensor, chunk) map."""

    def __init__(self, entries: List[dict], device):
        self.device = device
        self.n = len(entries)
        self.host = (OptTensor * self.n)()
        chunk = _lib.lib().cmh_opt_chunk_elems()
        bt, bc = [], []
        for t, e in enumerate(entries):
            p = e["param"]
            self.host[t] = OptTensor(p.data_ptr(), e["grad"].data_ptr(), e["m"].data_ptr(), None if e.get("v") is None else e["v"].data_ptr(),
                                     p.numel(), 0.0, float(e["weight_decay"]))
            nb = max(1, -(-p.numel() // chunk))
            bt += [t] * nb
            bc += list(range(nb))
        self.nblocks = len(bt)
        self.block_tensor = torch.tensor(bt, dtype=torch.int32, device=device)
        self.block_chunk = torch.tensor(bc, dtype=torch.int32, device=device)
        self.sumsq = torch.zeros(self.n, dtype=torch.float32, device=device)
        nbytes = ctypes.sizeof(self.host)
        self.pinned = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        self.table = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.key = tuple((e["param"].data_ptr(), e["grad"].data_ptr()) for e in entries)

    def upload(self, lrs: List[float]):
        for t, lr in enumerate(lrs):
            self.host[t].lr = lr
        ctypes.memmove(self.pinned.data_ptr(), ctypes.addressof(self.host), ctypes.sizeof(self.host))
        self.table.copy_(self.pinned, non_blocking=True)


def _check_param(p: torch.Tensor):
    if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
        raise _lib.CmhError("fused optimisers need contiguous fp32 CUDA parameters (there is no CPU path)")
    if p.grad is not None and (p.grad.dtype != torch.float32 or not p.grad.is_contiguous() or p.grad.is_sparse):
        raise _lib.CmhError("fused optimisers need dense contiguous fp32 gradients")


class FusedBertAdam(torch.optim.Optimizer):
    """``BertAdam`` (models/common/optimizer.py:52-165): Adam without bias correction, decoupled weight decay added to the update,
    per-tensor gradient clipping, warm-up schedules.  State keys (``step``, ``next_m``, ``next_v``) match the reference."""

    def __init__(self, params, lr, warmup=-1, t_total=-1, schedule="warmup_linear", b1=0.9, b2=0.999, e=1e-6, weight_decay=0.01,
                 max_grad_norm=1.0):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if schedule not in SCHEDULES:
            raise ValueError("Invalid schedule parameter: {}".format(schedule))
        if not 0.0 <= warmup < 1.0 and not warmup == -1:
            raise ValueError("Invalid warmup: {} - should be in [0.0, 1.0[ or -1".format(warmup))
        if not 0.0 <= b1 < 1.0 or not 0.0 <= b2 < 1.0 or not e >= 0.0:
            raise ValueError("Invalid b1 / b2 / e")
        super().__init__(params, dict(lr=lr, schedule=schedule, warmup=warmup, t_total=t_total, b1=b1, b2=b2, e=e,
                                      weight_decay=weight_decay, max_grad_norm=max_grad_norm))
        self._tables: Dict[tuple, _MultiTensor] = {}

    def _scheduled(self, group, step):
        if group["t_total"] != -1:
            return group["lr"] * SCHEDULES[group["schedule"]](step / group["t_total"], group["warmup"])
        return group["lr"]

    def get_lr(self):
        lr = []
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                state = self.state[p]
                if len(state) == 0:
                    return [0]
                lr.append(self._scheduled(group, state["step"]))
        return lr

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        # tensors that share (b1, b2, e, max_grad_norm) go into one launch; lr and weight decay are per tensor
        buckets: Dict[tuple, List[dict]] = {}
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                _check_param(p)
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["next_m"] = torch.zeros_like(p)
                    state["next_v"] = torch.zeros_like(p)
                key = (p.device, group["b1"], group["b2"], group["e"], group["max_grad_norm"])
                buckets.setdefault(key, []).append(dict(param=p, grad=p.grad, m=state["next_m"], v=state["next_v"],
                                                        weight_decay=group["weight_decay"], lr=self._scheduled(group, state["step"]), state=state))
        for (dev, b1, b2, e, mgn), entries in buckets.items():
            tkey = (dev, b1, b2, e, mgn) + tuple((x["param"].data_ptr(), x["grad"].data_ptr()) for x in entries)
            mt = self._tables.get(tkey)
            if mt is None:
                self._tables = {k: v for k, v in self._tables.items() if k[:5] != tkey[:5]}   # the parameter set changed
                mt = self._tables[tkey] = _MultiTensor(entries, dev)
            with torch.cuda.device(dev):
                mt.upload([x["lr"] for x in entries])
                _lib.check(_lib.lib().cmh_bert_adam_step(mt.table.data_ptr(), mt.n, mt.block_tensor.data_ptr(), mt.block_chunk.data_ptr(),
                                                         mt.nblocks, mt.sumsq.data_ptr(), b1, b2, e, mgn, _stream()))
            for x in entries:
                x["state"]["step"] += 1
        return loss


class FusedSGD(torch.optim.Optimizer):
    """``torch.optim.SGD(params, lr, momentum, weight_decay)`` as built for the HyP proxies (runners/DSPH/runner.py:86-89):
    dampening 0, no Nesterov.  State key ``momentum_buffer`` as in torch."""

    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        self._tables: Dict[tuple, _MultiTensor] = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            entries, first = [], None
            for p in group["params"]:
                if p.grad is None:
                    continue
                _check_param(p)
                state = self.state[p]
                fresh = "momentum_buffer" not in state
                if fresh:
                    state["momentum_buffer"] = torch.zeros_like(p)
                if first is None:
                    first = fresh
                elif first != fresh:
                    raise _lib.CmhError("FusedSGD: parameters of one group must start together")
                entries.append(dict(param=p, grad=p.grad, m=state["momentum_buffer"], v=None, weight_decay=group["weight_decay"]))
            if not entries:
                continue
            dev = entries[0]["param"].device
            tkey = (dev, id(group)) + tuple((x["param"].data_ptr(), x["grad"].data_ptr()) for x in entries)
            mt = self._tables.get(tkey)
            if mt is None:
                mt = self._tables[tkey] = _MultiTensor(entries, dev)
            with torch.cuda.device(dev):
                mt.upload([group["lr"]] * len(entries))
                _lib.check(_lib.lib().cmh_sgd_momentum_step(mt.table.data_ptr(), mt.n, mt.block_tensor.data_ptr(), mt.block_chunk.data_ptr(),
                                                            mt.nblocks, group["momentum"], int(bool(first)), _stream()))
        return loss


# ---- HyP objective with gradient -----------------------------------------------------------------------------------------
def hyp_loss_and_grad(x, y, label, proxies, threshold: float, alpha: float = 0.8):
    """``HyP.forward(x, y, label)`` and the gradients ``loss.backward()`` would leave in x, y and the proxies
    (models/DSPH/loss/HyP.py:18-69) -> (loss 0-dim fp32, dx [B, K], dy [B, K], dproxies [C, K]), all on the GPU."""
    dev = x.device
    if dev.type != "cuda":
        raise _lib.CmhError("hyp_loss_and_grad needs CUDA tensors (there is no CPU path)")
    x, y = x.detach().float().contiguous(), y.detach().to(dev).float().contiguous()
    proxies = proxies.detach().to(dev).float().contiguous()
    B, K = x.shape
    C = proxies.shape[0]
    lab = R.pack_labels(label.detach().to(dev))
    ws = torch.empty(256 + (2 * B + C) * K * 4, dtype=torch.uint8, device=dev)
    out = torch.empty((), dtype=torch.float32, device=dev)
    dx, dy, dp = torch.empty_like(x), torch.empty_like(y), torch.empty_like(proxies)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cmh_hyp_loss_grad_f32(x.data_ptr(), y.data_ptr(), lab.data_ptr(), proxies.data_ptr(), B, K, C, float(threshold),
                                                   float(alpha), ws.data_ptr(), ws.numel(), out.data_ptr(), dx.data_ptr(), dy.data_ptr(),
                                                   dp.data_ptr(), _stream()))
    return out, dx, dy, dp


class _HypLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, proxies, label, threshold, alpha):
        loss, dx, dy, dp = hyp_loss_and_grad(x, y, label, proxies, threshold, alpha)
        ctx.save_for_backward(dx, dy, dp)
        return loss

    @staticmethod
    def backward(ctx, g):
        dx, dy, dp = ctx.saved_tensors
        return g * dx, g * dy, g * dp, None, None, None


class HypLoss(torch.nn.Module):
    """Drop-in for the reference's ``HyP`` module (same constructor and ``forward(x, y, label)``); value and gradient come from the
    kernels of csrc/cmh_loss.cu / cmh_train.cu, so an unmodified PyTorch trainer can call ``loss.backward()`` on it."""

    def __init__(self, numclass=80, output_dim=16, hypseed=0, alpha=0.8, threshold=None, device="cuda"):
        assert threshold is not None, "DSPH must provide the threshold parameter"
        super().__init__()
        torch.manual_seed(hypseed)
        self.threshold, self.alpha = threshold, alpha
        proxies = torch.randn(numclass, output_dim)
        torch.nn.init.kaiming_normal_(proxies, mode="fan_out")
        self.proxies = torch.nn.Parameter(proxies.to(device))

    def forward(self, x=None, y=None, label=None):
        return _HypLossFn.apply(x, y, self.proxies, label, self.threshold, self.alpha)


def linear_tanh_backward(feat, y, dy, weight, want_dfeat: bool = False):
    """Backward of ``y = tanh(feat @ weight.T + bias)`` -> (dW [K, D], db [K], dfeat [B, D] or None)."""
    dev = feat.device
    feat, y, dy, weight = (t.detach().float().contiguous() for t in (feat, y, dy, weight))
    B, D = feat.shape
    K = weight.shape[0]
    dz = torch.empty((B, K), dtype=torch.float32, device=dev)
    dW, db = torch.empty_like(weight), torch.empty(K, dtype=torch.float32, device=dev)
    dfeat = torch.empty_like(feat) if want_dfeat else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cmh_linear_tanh_backward_f32(feat.data_ptr(), y.data_ptr(), dy.data_ptr(), weight.data_ptr(), B, D, K,
                                                          dz.data_ptr(), dW.data_ptr(), db.data_ptr(),
                                                          None if dfeat is None else dfeat.data_ptr(), _stream()))
    return dW, db, dfeat


class DsphHeadTrainer:
    """One DSPH training step with the CLIP backbone frozen (runners/DSPH/runner.py:104-127 with ``model.backbone`` in eval mode):
    images / captions -> towers (tcgen05 GEMMs, no grad) -> tanh(Linear) heads -> HyP loss -> gradients of the heads and the
    proxies -> FusedBertAdam (heads, cfg lr) + FusedSGD (proxies).  Dropout of the head (p = 0.2 in training, hash.py:12) is off:
    the head kernels implement the evaluation-mode module."""

    def __init__(self, model, numclass: int, threshold: float, alpha: float = 0.8, lr: float = 1e-3, t_total: int = -1, warmup: float = 0.1,
                 schedule: str = "warmup_cosine", b1: float = 0.9, b2: float = 0.98, e: float = 1e-6, weight_decay: float = 0.2,
                 max_grad_norm: float = 1.0, hyp_lr: float = 0.02, hyp_momentum: float = 0.9, hyp_weight_decay: float = 5e-4, hypseed: int = 0):
        self.model = model
        dev = model.backbone.device_
        self.hyp = HypLoss(numclass, model.output_dim, hypseed, alpha, threshold, device=dev)
        # the head's device tensors become the trainable parameters (the hash layer keeps pointing at them)
        self.params = {}
        for m in ("img", "txt"):
            w, b = model.hash.w[m]
            self.params[m] = (torch.nn.Parameter(w), torch.nn.Parameter(b))
            model.hash.w[m] = (self.params[m][0].data, self.params[m][1].data)
        flat = [p for m in ("img", "txt") for p in self.params[m]]
        for p in flat + [self.hyp.proxies]:
            p.grad = torch.zeros_like(p)
        self.optimizer = FusedBertAdam(flat, lr=lr, warmup=warmup, t_total=t_total, schedule=schedule, b1=b1, b2=b2, e=e,
                                       weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        self.optimizer_loss = FusedSGD([self.hyp.proxies], lr=hyp_lr, momentum=hyp_momentum, weight_decay=hyp_weight_decay)

    @torch.no_grad()
    def step(self, image, text, label):
        """-> loss (0-dim fp32 device tensor) of the batch BEFORE the update, like ``compute_loss`` in the reference."""
        fi, ft = self.model.backbone.encode_image(image), self.model.backbone.encode_text(text)
        yi, yt = self.model.hash.encode_img(fi), self.model.hash.encode_txt(ft)
        loss, dyi, dyt, dp = hyp_loss_and_grad(yi, yt, label, self.hyp.proxies, self.hyp.threshold, self.hyp.alpha)
        for m, feat, y, dy in (("img", fi, yi, dyi), ("txt", ft, yt, dyt)):
            dW, db, _ = linear_tanh_backward(feat, y, dy, self.params[m][0])
            self.params[m][0].grad.copy_(dW)
            self.params[m][1].grad.copy_(db)
        self.hyp.proxies.grad.copy_(dp)
        self.optimizer.step()
        self.optimizer_loss.step()
        return loss

    def sync_state_dict(self):
        """Write the trained head back into the model's checkpoint view (``model.state_dict()``)."""
        for m in ("img", "txt"):
            self.model.hash._sd["%s_hash.fc.weight" % m] = self.params[m][0].detach().cpu()
            self.model.hash._sd["%s_hash.fc.bias" % m] = self.params[m][1].detach().cpu()
        self.model.proxies = self.hyp.proxies.detach().float()
