"""Host-side mirrors of the in-scope reference models (CLIP backbone + per-method hash head) on the C ABI.

Same attribute and method names as ``models/DSPH/DSPH.py`` / ``models/DCMHT/DCMHT.py`` (``.backbone``, ``.hash``,
``.encode_image``, ``.encode_text``, ``.forward(image, text)``, ``.hash.encode_img/.encode_txt``), evaluation mode
only (what ``BaseTrainer.get_code`` runs under ``change_state("valid")``, ``runners/base.py:242-285``): dropout is the
identity and BatchNorm uses its running statistics.  ``encode_*_packed`` go straight from inputs to bit-packed codes
(no +-1 fp32 buffers), which is what the retrieval evaluator consumes.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _lib
from .encoder import ClipBackbone, _stream


class DcmhtHeadStruct(ctypes.Structure):
    """Mirror of ``struct cmh_dcmht_head``."""

    _fields_ = [(n, ctypes.c_void_p) for n in ("w_v", "b_v", "w_out", "b_out", "bn_scale", "bn_shift", "norm_gain", "norm_bias")] \
        + [("eps", ctypes.c_float)] + [(n, ctypes.c_void_p) for n in ("w_fc2", "b_fc2")]


def _words(nbits: int) -> int:
    w = _lib.lib().cmh_code_words(nbits)
    if w <= 0:
        raise _lib.CmhError("%d-bit codes are not supported" % nbits)
    return w


class _HeadBase(torch.nn.Module):
    def __init__(self, state_dict: Dict[str, torch.Tensor], device):
        super().__init__()
        self.device_ = torch.device(device)
        self._sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        self.refresh()

    def state_dict(self, *a, **k):
        return dict(self._sd)

    def load_state_dict(self, state_dict, strict: bool = True):
        for k in self._sd:
            if k in state_dict:
                self._sd[k] = state_dict[k].detach().cpu()
            elif strict:
                raise KeyError(k)
        self.refresh()

    def _dev(self, t):
        return t.detach().to(device=self.device_, dtype=torch.float32).contiguous()

    def _feat(self, feat):
        return feat.to(device=self.device_, dtype=torch.float32).contiguous()


class DsphHashLayer(_HeadBase):
    """models/DSPH/hash/hash.py:17-46 — ``tanh(Linear(inputDim, outputDim))`` per modality."""

    def refresh(self):
        self.w = {m: (self._dev(self._sd["%s_hash.fc.weight" % m]), self._dev(self._sd["%s_hash.fc.bias" % m])) for m in ("img", "txt")}
        self.nbits, self.in_dim = self.w["img"][0].shape

    def _run(self, feat, m, packed: bool):
        feat = self._feat(feat)
        B = feat.shape[0]
        w, b = self.w[m]
        out = torch.empty((B, self.nbits), dtype=torch.float32, device=self.device_)
        codes = torch.empty((B, _words(self.nbits)), dtype=torch.int32, device=self.device_) if packed else None
        with torch.cuda.device(self.device_):
            _lib.check(_lib.lib().cmh_head_dsph(feat.data_ptr(), B, self.in_dim, w.data_ptr(), b.data_ptr(), self.nbits,
                                                out.data_ptr(), None if codes is None else codes.data_ptr(), _stream()))
        return (out, codes) if packed else out

    def encode_img(self, embeds):
        return self._run(embeds, "img", False)

    def encode_txt(self, embeds):
        return self._run(embeds, "txt", False)

    def quantization(self, code):   # hash.py:34-35
        return torch.tanh(code)

    def forward(self, img_embeds, txt_embeds):
        return self.encode_img(img_embeds), self.encode_txt(txt_embeds)


class DcmhtHashLayer(_HeadBase):
    """models/DCMHT/hash/hash.py:48-82 — MHA(len-1) -> BatchNorm1d (image) | LayerNorm (text) -> fc2 -> ReLU -> pair softmax."""

    def refresh(self):
        self.heads, self.keep = {}, []
        for m in ("img", "txt"):
            p = "%s_hash." % m
            D = self._sd[p + "atten.out_proj.weight"].shape[0]
            h = DcmhtHeadStruct()

            def put(t):
                t = self._dev(t)
                self.keep.append(t)
                return t.data_ptr()

            h.w_v, h.b_v = put(self._sd[p + "atten.in_proj_weight"][2 * D:]), put(self._sd[p + "atten.in_proj_bias"][2 * D:])
            h.w_out, h.b_out = put(self._sd[p + "atten.out_proj.weight"]), put(self._sd[p + "atten.out_proj.bias"])
            h.eps = 1e-5
            if (p + "norm.running_mean") in self._sd:   # BatchNorm1d in eval mode == per-feature affine map
                scale = self._sd[p + "norm.weight"].double() / torch.sqrt(self._sd[p + "norm.running_var"].double() + 1e-5)
                shift = self._sd[p + "norm.bias"].double() - self._sd[p + "norm.running_mean"].double() * scale
                h.bn_scale, h.bn_shift = put(scale.float()), put(shift.float())
            else:
                h.norm_gain, h.norm_bias = put(self._sd[p + "norm.weight"]), put(self._sd[p + "norm.bias"])
            h.w_fc2, h.b_fc2 = put(self._sd[p + "fc2.weight"]), put(self._sd[p + "fc2.bias"])
            self.heads[m] = h
            self.in_dim, self.nbits = D, self._sd[p + "fc2.weight"].shape[0] // 2

    def _run(self, feat, m, packed: bool):
        feat = self._feat(feat)
        B = feat.shape[0]
        scratch = torch.empty((B, 2 * self.in_dim + 2 * self.nbits), dtype=torch.float32, device=self.device_)
        probs = torch.empty((B, 2 * self.nbits), dtype=torch.float32, device=self.device_)
        codes = torch.empty((B, _words(self.nbits)), dtype=torch.int32, device=self.device_) if packed else None
        with torch.cuda.device(self.device_):
            _lib.check(_lib.lib().cmh_head_dcmht(feat.data_ptr(), B, self.in_dim, ctypes.byref(self.heads[m]), self.nbits,
                                                 scratch.data_ptr(), probs.data_ptr(),
                                                 None if codes is None else codes.data_ptr(), _stream()))
        return (probs, codes) if packed else probs

    def encode_img(self, embeds):
        return self._run(embeds, "img", False)

    def encode_txt(self, embeds):
        return self._run(embeds, "txt", False)

    def forward(self, img_embeds, txt_embeds):
        return self.encode_img(img_embeds), self.encode_txt(txt_embeds)


class _Model(torch.nn.Module):
    HASH = None

    def __init__(self, clip_state_dict, hash_state_dict, device="cuda"):
        super().__init__()
        self.backbone = ClipBackbone(clip_state_dict, return_patches=False, device=device)
        self.hash = self.HASH(hash_state_dict, device)
        self.output_dim = self.hash.nbits

    def encode_image(self, image):          # models/DSPH/DSPH.py:37-42, models/DCMHT/DCMHT.py:37-42
        return self.hash.encode_img(self.backbone.encode_image(image))

    def encode_text(self, text):            # :44-48
        return self.hash.encode_txt(self.backbone.encode_text(text))

    def forward(self, image, text, labels=None, indexs=None, return_loss=False):   # models/base.py:46-51
        if return_loss:
            raise NotImplementedError("the training objective is outside the B200 hot path (DESIGN.md: out of scope)")
        return self.encode_image(image), self.encode_text(text)

    # encoder -> head -> bit-packed codes in one stream, no +-1 fp32 buffer (replaces get_code's make_hash_code + buffers,
    # runners/base.py:242-257); returns int32 [B, W] in the evaluator's packed layout (include/cmh.h)
    def encode_image_packed(self, image):
        return self.hash._run(self.backbone.encode_image(image), "img", True)[1]

    def encode_text_packed(self, text):
        return self.hash._run(self.backbone.encode_text(text), "txt", True)[1]


class DSPH(_Model):
    HASH = DsphHashLayer

    @staticmethod
    def make_hash_code(code):               # runners/base.py:407-410 (in place, like the reference)
        return code.sign_()


class DCMHT(_Model):
    HASH = DcmhtHashLayer

    @staticmethod
    def make_hash_code(code):               # runners/DCMHT/runner.py:83-95
        pairs = code.reshape(code.shape[0], -1, 2)   # argmax over (2j, 2j+1); a tie picks index 0 -> -1
        return torch.where(pairs[..., 1] > pairs[..., 0], 1.0, -1.0)


def get_code(model, data_loader, length: int, device=None):
    """Drop-in for ``BaseTrainer.get_code`` (runners/base.py:242-266) that stays bit-packed.

    ``data_loader`` yields the reference's batches ``(image, text, key_padding_mask, label, index)`` (host tensors,
    ideally pinned).  Returns ``(img_codes, txt_codes)``: int32 ``[length, W]`` device tensors in the evaluator's packed
    layout, row ``index[i]`` of each buffer holding sample i's code — what ``calc_utils.calc_map_k`` consumes directly
    (no +-1 fp32 ``[length, K]`` buffers, no device->host hop).  The host->device copy of batch i+1
    runs on a side stream while batch i is encoded.
    """
    dev = torch.device(device) if device is not None else model.backbone.device_
    W = _words(model.output_dim)
    img_buf = torch.zeros((length, W), dtype=torch.int32, device=dev)
    txt_buf = torch.zeros((length, W), dtype=torch.int32, device=dev)
    main = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev)

    def stage(batch):
        image, text, _mask, _label, index = batch
        with torch.cuda.stream(side):
            item = (image.to(dev, non_blocking=True), text.to(dev, non_blocking=True),
                    torch.as_tensor(index).to(dev, non_blocking=True).long())
            done = torch.cuda.Event()
            done.record(side)
        return item, done

    def encode(staged):
        (image, text, index), done = staged
        main.wait_event(done)
        for t in (image, text, index):
            t.record_stream(main)
        img_buf[index] = model.encode_image_packed(image)
        txt_buf[index] = model.encode_text_packed(text)

    pending = None
    for batch in data_loader:
        nxt = stage(batch)
        if pending is not None:
            encode(pending)
        pending = nxt
    if pending is not None:
        encode(pending)
    return img_buf, txt_buf
