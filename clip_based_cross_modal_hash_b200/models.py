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
from .encoder import BlockWeights, ClipBackbone, Tower, _stream


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


def load_clip_state_dict(clip_path: str) -> Dict[str, torch.Tensor]:
    """The checkpoint read of ``BaseModel.load_backbone`` (models/base.py:18-31): a TorchScript archive (the OpenAI
    ``ViT-B-32.pt``) or a plain ``torch.save``d state dict."""
    try:
        return torch.jit.load(clip_path, map_location="cpu").eval().state_dict()
    except RuntimeError:
        return torch.load(clip_path, map_location="cpu")


class _RegistryMixin:
    """What ``BaseTrainer.build_model`` needs from a registered model (runners/base.py:98-107): ``from_config`` and a
    ``state_dict`` / ``load_state_dict`` pair with the reference's key prefixes (``backbone.*``, ``hash.*``; other keys of a
    trained checkpoint — loss parameters, feature buffers — are ignored: this path only evaluates)."""

    HEAD_INIT = None   # synth generator used until a trained checkpoint is loaded

    @classmethod
    def from_config(cls, cfg, output_dim=16, train_num=10000, device="cuda"):
        from . import synth

        clip = load_clip_state_dict(cfg.get("clip_path", "./ViT-B-32.pt"))
        embed_dim = clip["text_projection"].shape[1]
        return cls(clip, getattr(synth, cls.HEAD_INIT)(embed_dim, output_dim, seed=0), device=device)

    def state_dict(self, *a, **k):
        out = {"backbone." + n: v for n, v in self.backbone.state_dict().items()}
        out.update({"hash." + n: v for n, v in self.hash.state_dict().items()})
        return out

    def load_state_dict(self, state_dict, strict: bool = True):
        bb = {k[len("backbone."):]: v for k, v in state_dict.items() if k.startswith("backbone.")}
        hh = {k[len("hash."):]: v for k, v in state_dict.items() if k.startswith("hash.")}
        if strict and (not bb or not hh):
            raise KeyError("expected keys with the prefixes 'backbone.' and 'hash.'")
        if bb:
            self.backbone.load_state_dict(bb, strict=strict)
        if hh:
            self.hash.load_state_dict(hh, strict=strict)

    def float(self):           # runners/base.py:106-107 call .float() / .to(device): nothing to convert here
        return self

    def to(self, *a, **k):
        return self

    # models/base.py:57-63.  This path holds no trainable tensors (evaluation only): nothing to freeze.  BaseTrainer still
    # builds its optimiser from `.backbone.parameters()` / `.hash.parameters()` (runners/base.py:134-136); those iterators
    # are empty here, so a trainer that only evaluates should skip `build_optimizer` (torch refuses empty parameter lists).
    def freezen(self):
        return None

    def unfreezen(self):
        return None


class _Model(_RegistryMixin, torch.nn.Module):
    HASH = None

    def __init__(self, clip_state_dict, hash_state_dict, device="cuda"):
        super().__init__()
        self.backbone = ClipBackbone(clip_state_dict, return_patches=False, device=device)
        self.hash = self.HASH(hash_state_dict, device)

    @property
    def output_dim(self):   # follows the head: a checkpoint with another code length may be loaded later
        return self.hash.nbits

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


def hyp_loss(x, y, label, proxies, threshold: float, alpha: float = 0.8) -> torch.Tensor:
    """Forward value of ``HyP.forward(x, y, label)`` (models/DSPH/loss/HyP.py:18-69) on the GPU -> 0-dim fp32 device tensor.
    Evaluation / monitoring only: no gradient flows (the training step is out of this round's scope)."""
    from . import retrieval as R

    dev = x.device
    if dev.type != "cuda":
        raise _lib.CmhError("hyp_loss needs CUDA tensors (there is no CPU path)")
    x, y = x.detach().float().contiguous(), y.detach().to(dev).float().contiguous()
    proxies = proxies.detach().to(dev).float().contiguous()
    B, K = x.shape
    C = proxies.shape[0]
    lab = R.pack_labels(label.detach().to(dev))
    need = 256 + (2 * B + C) * K * 4
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    out = torch.empty((), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cmh_hyp_loss_f32(x.data_ptr(), y.data_ptr(), lab.data_ptr(), proxies.data_ptr(), B, K, C,
                                              float(threshold), float(alpha), ws.data_ptr(), ws.numel(), out.data_ptr(), _stream()))
    return out


class DSPH(_Model):
    HASH = DsphHashLayer
    HEAD_INIT = "dsph_head_state_dict"

    def __init__(self, clip_state_dict, hash_state_dict, device="cuda", proxies=None, threshold: float = 0.0, alpha: float = 0.8):
        super().__init__(clip_state_dict, hash_state_dict, device)
        self.proxies, self.threshold, self.alpha = proxies, threshold, alpha   # HyP(numclass, output_dim, ..., threshold)

    def load_state_dict(self, state_dict, strict: bool = True):
        if "hyp.proxies" in state_dict:                                       # models/DSPH/DSPH.py:33-35
            self.proxies = state_dict["hyp.proxies"].detach().float()
        return super().load_state_dict(state_dict, strict)

    def state_dict(self, *a, **k):
        out = super().state_dict(*a, **k)
        if self.proxies is not None:                                          # keep a save / load round trip lossless
            out["hyp.proxies"] = self.proxies
        return out

    def object_function(self, img_hash, txt_hash, labels=None, indexs=None, **kwargs):
        """models/DSPH/DSPH.py:78-82 -> (loss, loss_dict); the VALUE of the HyP objective only, no gradient."""
        if self.proxies is None:
            raise _lib.CmhError("DSPH.object_function needs the HyP proxies (a trained checkpoint's 'hyp.proxies')")
        if labels is None:
            labels = torch.eye(img_hash.shape[0], dtype=torch.int64)
        loss = hyp_loss(img_hash, txt_hash, labels, self.proxies, self.threshold, self.alpha)
        return loss, {"All loss": loss.data}

    @staticmethod
    def make_hash_code(code):               # runners/base.py:407-410 (in place, like the reference)
        return code.sign_()


class DCMHT(_Model):
    HASH = DcmhtHashLayer
    HEAD_INIT = "dcmht_head_state_dict"

    @staticmethod
    def make_hash_code(code):               # runners/DCMHT/runner.py:83-95
        pairs = code.reshape(code.shape[0], -1, 2)   # argmax over (2j, 2j+1); a tie picks index 0 -> -1
        return torch.where(pairs[..., 1] > pairs[..., 0], 1.0, -1.0)


class MithHeadStruct(ctypes.Structure):
    """Mirror of ``struct cmh_mith_head``."""

    _P4 = ctypes.c_void_p * 4
    _fields_ = [("dim", ctypes.c_int32), ("nbits", ctypes.c_int32), ("mlp_layers", ctypes.c_int32), ("top_k", ctypes.c_int32),
                ("ln_gain", _P4), ("ln_bias", _P4), ("w1", _P4), ("b1", _P4), ("w2", _P4), ("b2", _P4),
                ("w1_split", _P4), ("w2_split", _P4), ("w_concept", ctypes.c_void_p), ("pos", ctypes.c_void_p), ("transformer", Tower),
                ("w_bits", ctypes.c_void_p), ("b_bits", ctypes.c_void_p), ("w_cproj", ctypes.c_void_p), ("b_cproj", ctypes.c_void_p)]


class MithHashLayer(_HeadBase):
    """models/MITH/hash/hash.py:193-254 ``HashLayer`` (evaluation mode): global concept learning on the CLS/EOS feature,
    localized token aggregation + a small transformer + bitwise hashing on the token features."""

    def __init__(self, state_dict, device, top_k_label: int = 8, split_precision: bool = True):
        self.top_k = top_k_label
        # the token path feeds a discontinuous top-k selection: its MLP GEMMs run on hi/lo-split bf16 operands by default
        self.split_precision = split_precision
        super().__init__(state_dict, device)

    def refresh(self):
        sd, self.keep, self.heads, self.blocks = self._sd, [], {}, {}

        def f32(t):
            t = self._dev(t)
            self.keep.append(t)
            return t.data_ptr()

        def bf16(t):
            t = t.detach().to(device=self.device_, dtype=torch.bfloat16).contiguous()
            self.keep.append(t)
            return t.data_ptr()

        def split3(w):   # [out][in] fp32 -> bf16 [out][3*in] = [hi | lo | hi]
            w = w.detach().to(device=self.device_, dtype=torch.float32)
            hi = w.to(torch.bfloat16)
            lo = (w - hi.float()).to(torch.bfloat16)
            t = torch.cat([hi, lo, hi], dim=1).contiguous()
            self.keep.append(t)
            return t.data_ptr()

        for m, g, t in (("img", "gcl_i.", "lct_i."), ("txt", "gcl_t.", "lct_t.")):
            h = MithHeadStruct()
            h.nbits, h.dim = sd[g + "common_concept_embedding.weight"].shape
            h.top_k = self.top_k
            n = 0
            while (g + "mlp.mlps.%d.0.weight" % n) in sd:
                h.ln_gain[n], h.ln_bias[n] = f32(sd[g + "mlp.lns.%d.weight" % n]), f32(sd[g + "mlp.lns.%d.bias" % n])
                h.w1[n], h.b1[n] = bf16(sd[g + "mlp.mlps.%d.0.weight" % n]), f32(sd[g + "mlp.mlps.%d.0.bias" % n])
                h.w2[n], h.b2[n] = bf16(sd[g + "mlp.mlps.%d.3.weight" % n]), f32(sd[g + "mlp.mlps.%d.3.bias" % n])
                if self.split_precision:
                    h.w1_split[n] = split3(sd[g + "mlp.mlps.%d.0.weight" % n])
                    h.w2_split[n] = split3(sd[g + "mlp.mlps.%d.3.weight" % n])
                n += 1
            h.mlp_layers = n
            h.w_concept = f32(sd[g + "common_concept_embedding.weight"])
            h.pos = f32(sd[t + "position.pe"][: h.nbits, 0])
            bp = t + "transformer.resblocks."
            layers = len({k[len(bp):].split(".")[0] for k in sd if k.startswith(bp)})
            blocks = (BlockWeights * layers)()
            for i in range(layers):
                p, b = "%s%d." % (bp, i), blocks[i]
                b.ln1_gain, b.ln1_bias = f32(sd[p + "ln_1.weight"]), f32(sd[p + "ln_1.bias"])
                b.w_qkv, b.b_qkv = bf16(sd[p + "attn.in_proj_weight"]), f32(sd[p + "attn.in_proj_bias"])
                b.w_out, b.b_out = bf16(sd[p + "attn.out_proj.weight"]), f32(sd[p + "attn.out_proj.bias"])
                b.ln2_gain, b.ln2_bias = f32(sd[p + "ln_2.weight"]), f32(sd[p + "ln_2.bias"])
                b.w_fc, b.b_fc = bf16(sd[p + "mlp.c_fc.weight"]), f32(sd[p + "mlp.c_fc.bias"])
                b.w_proj, b.b_proj = bf16(sd[p + "mlp.c_proj.weight"]), f32(sd[p + "mlp.c_proj.bias"])
            self.blocks[m] = blocks
            h.transformer.width, h.transformer.layers, h.transformer.heads = h.dim, layers, h.dim // 64
            h.transformer.out_dim = h.dim
            h.transformer.blocks = blocks
            h.w_bits = f32(torch.stack([sd["%shashing.fc_list.%d.weight" % (t, k)][0] for k in range(h.nbits)]))
            h.b_bits = f32(torch.stack([sd["%shashing.fc_list.%d.bias" % (t, k)][0] for k in range(h.nbits)]))
            h.w_cproj, h.b_cproj = bf16(sd["%s_concept_proj.weight" % m]), f32(sd["%s_concept_proj.bias" % m])
            self.heads[m] = h
            self.nbits, self.in_dim = int(h.nbits), int(h.dim)
        self._ws = {}   # one workspace per modality: the two branches may run on different streams (get_code)

    def _run(self, m, cls, tokens, per_sample, first, L, pad, want_trans=True, packed=False):
        """cls [B, D]; tokens: fp32 [*, D] buffer whose rows b*per_sample + first .. + L-1 are sample b's tokens."""
        h = self.heads[m]
        B, D, K = cls.shape[0], self.in_dim, self.nbits
        dev = self.device_
        need = _lib.lib().cmh_head_mith_workspace_bytes(ctypes.byref(h), B, L)
        if need <= 0:
            raise _lib.CmhError("cmh_head_mith_workspace_bytes failed")
        ws = self._ws.get(m)
        if ws is None or ws.numel() < need:
            ws = self._ws[m] = torch.empty(need, dtype=torch.uint8, device=dev)
        res = torch.empty((B, D), dtype=torch.float32, device=dev)
        cls_hash = torch.empty((B, K), dtype=torch.float32, device=dev)
        tok_hash = torch.empty((B, K), dtype=torch.float32, device=dev)
        trans = torch.empty((B, K, D), dtype=torch.float32, device=dev) if want_trans else None
        codes = torch.empty((B, _words(K)), dtype=torch.int32, device=dev) if packed else None
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().cmh_head_mith(
                ctypes.byref(h), cls.data_ptr(), tokens.data_ptr(), per_sample, first, L, None if pad is None else pad.data_ptr(), B,
                ws.data_ptr(), ws.numel(), res.data_ptr(), cls_hash.data_ptr(), tok_hash.data_ptr(),
                None if trans is None else trans.data_ptr(), None if codes is None else codes.data_ptr(), _stream()))
        return res, cls_hash, tok_hash, trans, codes

    @staticmethod
    def _lnd_to_rows(tokens):
        """The reference's [L, B, D] token layout -> contiguous sample-major rows [B*L, D]."""
        return tokens.permute(1, 0, 2).contiguous()

    def encode_img(self, img_cls, img_tokens):                                  # hash.py:231-238
        cls, tok = self._feat(img_cls), self._lnd_to_rows(self._feat(img_tokens))
        L = img_tokens.shape[0]
        res, ch, th, trans, _ = self._run("img", cls, tok, L, 0, L, None)
        return res, ch, th, trans.permute(1, 0, 2)

    def encode_txt(self, txt_eos, txt_tokens, key_padding_mask):                # hash.py:240-247
        cls, tok = self._feat(txt_eos), self._lnd_to_rows(self._feat(txt_tokens))
        L = txt_tokens.shape[0]
        pad = None if key_padding_mask is None else key_padding_mask.to(self.device_).to(torch.uint8).contiguous()
        res, ch, th, trans, _ = self._run("txt", cls, tok, L, 0, L, pad)
        return res, ch, th, trans.permute(1, 0, 2)

    def forward(self, img_tokens, txt_tokens, img_cls, txt_eos, key_padding_mask):   # hash.py:249-254
        return self.encode_img(img_cls, img_tokens) + self.encode_txt(txt_eos, txt_tokens, key_padding_mask)


class MITH(_RegistryMixin, torch.nn.Module):
    """models/MITH/MITH.py (evaluation mode): backbone with ``return_patches=True`` + ``MithHashLayer``."""

    HEAD_INIT = "mith_head_state_dict"

    def __init__(self, clip_state_dict, hash_state_dict, device="cuda", top_k_label: int = 8):
        super().__init__()
        self.backbone = ClipBackbone(clip_state_dict, return_patches=True, device=device)
        self.hash = MithHashLayer(hash_state_dict, device, top_k_label)

    @property
    def output_dim(self):
        return self.hash.nbits

    def _image(self, image, want_trans, packed):
        cls, tokens, _ = self.backbone.encode_image_raw(image, True, False)     # tokens [B, 50, E], row 0 = CLS
        L = tokens.shape[1]
        return self.hash._run("img", cls, tokens, L, 1, L - 1, None, want_trans, packed)

    def _text(self, text, key_padding_mask, want_trans, packed):
        eos, tokens, _, newmask = self.backbone.encode_text_raw(text, key_padding_mask, True, False)
        L = tokens.shape[1]
        return self.hash._run("txt", eos, tokens, L, 0, L, newmask, want_trans, packed)   # new mask: MITH.py:61-63

    def encode_image(self, image):                                              # MITH.py:52-57
        res, ch, th, trans, _ = self._image(image, True, False)
        return res, ch, th, trans.permute(1, 0, 2)

    def encode_text(self, text, key_padding_mask=None):                         # MITH.py:59-65
        res, ch, th, trans, _ = self._text(text, key_padding_mask, True, False)
        return res, ch, th, trans.permute(1, 0, 2)

    def forward(self, image, text, key_padding_mask=None, labels=None, indexs=None, return_loss=False):   # MITH.py:67-76
        if return_loss:
            raise NotImplementedError("the training objective is outside the B200 hot path (DESIGN.md: out of scope)")
        return self.encode_image(image) + self.encode_text(text, key_padding_mask)

    def generate_hash(self, image, text, key_padding_mask=None):               # runners/MITH/runner.py:125-131
        _, ich, ith, _, _ = self._image(image, False, False)
        _, tch, tth, _, _ = self._text(text, key_padding_mask, False, False)
        return ich + ith, tch + tth

    @staticmethod
    def make_hash_code(code):                                                   # runners/base.py:407-410
        return code.sign_()

    def encode_image_packed(self, image):
        return self._image(image, False, True)[4]

    def encode_text_packed(self, text, key_padding_mask=None):
        return self._text(text, key_padding_mask, False, True)[4]


def merge_code_buffers(*buffers, group=None):
    """Combine per-rank packed code buffers under the reference's DDP evaluation pattern (runners/base.py:259-264): every rank
    filled only the rows of its own sampler indices, the rest is zero.  One all-reduce of the PACKED buffers (COCO 64-bit
    gallery: 0.94 MB) replaces the reference's barrier + fp32 all-reduce(SUM) of [length, K] floats (30 MB).  The reduction is
    a byte-wise MAX (NCCL has no bitwise OR; against zero bytes MAX == OR), so rows that DistributedSampler duplicated to pad
    the last batch stay correct — the reference sums them to +-2."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for b in buffers:
            dist.all_reduce(b.view(torch.uint8), op=dist.ReduceOp.MAX, group=group)
    return buffers


def get_code(model, data_loader, length: int, device=None, distributed: bool = False, group=None):
    """Drop-in for ``BaseTrainer.get_code`` (runners/base.py:242-266) that stays bit-packed.

    ``data_loader`` yields the reference's batches ``(image, text, key_padding_mask, label, index)`` (host tensors,
    ideally pinned).  Returns ``(img_codes, txt_codes)``: int32 ``[length, W]`` device tensors in the evaluator's packed
    layout, row ``index[i]`` of each buffer holding sample i's code — what ``calc_utils.calc_map_k_packed`` consumes directly
    (no +-1 fp32 ``[length, K]`` buffers, no device->host hop).  ``distributed=True``: every rank encodes
    its sampler's share and the buffers are merged by ``merge_code_buffers``.  The host->device copy of batch i+1
    runs on a side stream while batch i is encoded.
    """
    dev = torch.device(device) if device is not None else model.backbone.device_
    W = _words(model.output_dim)
    img_buf = torch.zeros((length, W), dtype=torch.int32, device=dev)
    txt_buf = torch.zeros((length, W), dtype=torch.int32, device=dev)
    main = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev)          # host -> device copies of the next batch
    txt_stream = torch.cuda.Stream(dev)    # the text tower runs beside the image tower: each fills the other's partial waves

    needs_mask = isinstance(model, MITH)   # MITHTrainer.generate_hash passes key_padding_mask (runners/MITH/runner.py:125-131)
    # two sets of device staging buffers (allocated from the first batch's shapes): no allocator traffic across streams
    slots = [None, None]
    free = [None, None]   # event: the batch that last used this slot has been encoded

    def stage(batch, j):
        image, text, mask, _label, index = batch
        index = torch.as_tensor(index)
        n = image.shape[0]
        want_dt = torch.uint8 if image.dtype == torch.uint8 else torch.float32
        if slots[j] is None or slots[j][0].shape[0] < n or slots[j][1].shape[1] != text.shape[1] or slots[j][0].dtype != want_dt:
            slots[j] = [torch.empty((n,) + tuple(image.shape[1:]), dtype=torch.uint8 if image.dtype == torch.uint8 else torch.float32,
                                    device=dev),   # uint8 pixels stay uint8: normalised on the GPU (encoder.ClipBackbone)
                        torch.empty((n, text.shape[1]), dtype=torch.int64, device=dev),
                        torch.empty((n,), dtype=torch.int64, device=dev),
                        torch.empty((n, text.shape[1]), dtype=torch.bool, device=dev)]
            free[j] = None
            torch.cuda.current_stream(dev).synchronize()   # (re)allocation only: first batches or a larger batch
        if free[j] is not None:
            side.wait_event(free[j])
        with torch.cuda.stream(side):
            bufs = slots[j]
            bufs[0][:n].copy_(image, non_blocking=True)
            bufs[1][:n].copy_(text, non_blocking=True)
            bufs[2][:n].copy_(index, non_blocking=True)
            has_mask = needs_mask and mask is not None
            if has_mask:
                bufs[3][:n].copy_(mask, non_blocking=True)
            done = torch.cuda.Event()
            done.record(side)
        return (j, n, has_mask), done

    def encode(staged):
        (j, n, has_mask), done = staged
        image, text, index, mask = (t[:n] for t in slots[j])
        main.wait_event(done)
        txt_stream.wait_stream(main)       # orders the text tower after everything queued so far (inputs, txt_buf writes)
        with torch.cuda.stream(txt_stream):
            tcodes = model.encode_text_packed(text, mask) if has_mask else model.encode_text_packed(text)
        icodes = model.encode_image_packed(image)
        main.wait_stream(txt_stream)
        tcodes.record_stream(main)
        img_buf[index] = icodes
        txt_buf[index] = tcodes
        free[j] = torch.cuda.Event()
        free[j].record(main)

    pending = None
    for i, batch in enumerate(data_loader):
        nxt = stage(batch, i & 1)
        if pending is not None:
            encode(pending)
        pending = nxt
    if pending is not None:
        encode(pending)
    if distributed:
        merge_code_buffers(img_buf, txt_buf, group=group)
    return img_buf, txt_buf


def register_into_reference(registry=None, names=("DSPH", "DCMHT", "MITH")):
    """Make ``BaseTrainer.build_model`` (runners/base.py:98-102: ``registry.get_model_class(arch).from_config(...)``) return
    these models for the given ``arch`` names.  The reference's ``registry.register_model`` decorator insists on a
    ``BaseModel`` subclass and refuses names that are taken (common/register.py:75-91), so the mapping is overridden
    directly.  Call after ``import models`` of the reference (its own classes register themselves at import time)."""
    if registry is None:
        from common.register import registry as _registry   # the reference's module

        registry = _registry
    table = {"DSPH": DSPH, "DCMHT": DCMHT, "MITH": MITH}
    for n in names:
        registry.mapping["model_name_mapping"][n] = table[n]
    return registry
