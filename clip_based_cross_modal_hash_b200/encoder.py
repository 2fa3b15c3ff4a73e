"""Host-side mirror of the reference's CLIP backbone API on top of the C ABI (include/cmh.h, section E).

``ClipBackbone`` keeps the call shapes of ``models/CLIP/model.py:CLIP`` that the in-scope methods use
(``models/DCMHT/DCMHT.py:37-48``, ``models/DSPH/DSPH.py:37-48``, ``models/MITH/MITH.py:52-65``):

    encode_image(image)                      -> cls [B, E]          | (cls, seq [L-1, B, E], attn [B, L-1]) if return_patches
    encode_text(text, key_padding_mask=None) -> eos [B, E]          | (eos, seq [L, B, E], attn [B, L], new_mask [B, L])

and is built from a reference ``state_dict`` exactly like ``build_model`` does (``models/CLIP/model.py:443-489``: every
hyper-parameter is inferred from tensor shapes).  PyTorch is used for device memory and the current stream only; all
arithmetic happens in libcmh.so.  There is no CPU fallback: without CUDA / the library every call raises.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _lib

class BlockWeights(ctypes.Structure):
    """Mirror of ``struct cmh_block_weights``."""

    _fields_ = [(n, ctypes.c_void_p) for n in (
        "ln1_gain", "ln1_bias", "w_qkv", "b_qkv", "w_out", "b_out", "ln2_gain", "ln2_bias", "w_fc", "b_fc", "w_proj", "b_proj")]


class Tower(ctypes.Structure):
    """Mirror of ``struct cmh_tower``."""

    _fields_ = [
        ("width", ctypes.c_int32), ("layers", ctypes.c_int32), ("heads", ctypes.c_int32), ("out_dim", ctypes.c_int32),
        ("blocks", ctypes.POINTER(BlockWeights)),
        ("ln_out_gain", ctypes.c_void_p), ("ln_out_bias", ctypes.c_void_p), ("w_out_proj", ctypes.c_void_p),
        ("pos_emb", ctypes.c_void_p),
        ("patch", ctypes.c_int32), ("resolution", ctypes.c_int32),
        ("w_patch", ctypes.c_void_p), ("cls_emb", ctypes.c_void_p), ("ln_pre_gain", ctypes.c_void_p),
        ("ln_pre_bias", ctypes.c_void_p),
        ("vocab", ctypes.c_int32), ("context", ctypes.c_int32),
        ("tok_emb", ctypes.c_void_p), ("eot_id", ctypes.c_int64),
    ]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _DeviceTower:
    """Device copies of one tower's weights in the ABI's formats + the ctypes struct pointing at them."""

    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, device: torch.device):
        self.keep = []  # tensors referenced by raw pointers

        def mat(t):  # bf16, torch.nn.Linear layout
            t = t.detach().to(device=device, dtype=torch.bfloat16).contiguous()
            self.keep.append(t)
            return t.data_ptr()

        def vec(t):
            t = t.detach().to(device=device, dtype=torch.float32).contiguous()
            self.keep.append(t)
            return t.data_ptr()

        bp = prefix + "transformer.resblocks."
        layers = len({k[len(bp):].split(".")[0] for k in sd if k.startswith(bp)})
        self.blocks = (BlockWeights * layers)()
        for i in range(layers):
            p, b = "%s%d." % (bp, i), self.blocks[i]
            b.ln1_gain, b.ln1_bias = vec(sd[p + "ln_1.weight"]), vec(sd[p + "ln_1.bias"])
            b.w_qkv, b.b_qkv = mat(sd[p + "attn.in_proj_weight"]), vec(sd[p + "attn.in_proj_bias"])
            b.w_out, b.b_out = mat(sd[p + "attn.out_proj.weight"]), vec(sd[p + "attn.out_proj.bias"])
            b.ln2_gain, b.ln2_bias = vec(sd[p + "ln_2.weight"]), vec(sd[p + "ln_2.bias"])
            b.w_fc, b.b_fc = mat(sd[p + "mlp.c_fc.weight"]), vec(sd[p + "mlp.c_fc.bias"])
            b.w_proj, b.b_proj = mat(sd[p + "mlp.c_proj.weight"]), vec(sd[p + "mlp.c_proj.bias"])
        t = Tower()
        t.layers, t.blocks = layers, self.blocks
        if prefix == "visual.":
            conv = sd["visual.conv1.weight"]                                   # [D, 3, P, P]
            t.width, t.patch = conv.shape[0], conv.shape[-1]
            grid = round((sd["visual.positional_embedding"].shape[0] - 1) ** 0.5)   # model.py:449
            t.resolution = t.patch * grid
            t.out_dim = sd["visual.proj"].shape[1]
            t.w_patch = mat(conv.reshape(conv.shape[0], -1))
            t.cls_emb = vec(sd["visual.class_embedding"])
            t.pos_emb = vec(sd["visual.positional_embedding"])
            t.ln_pre_gain, t.ln_pre_bias = vec(sd["visual.ln_pre.weight"]), vec(sd["visual.ln_pre.bias"])
            t.ln_out_gain, t.ln_out_bias = vec(sd["visual.ln_post.weight"]), vec(sd["visual.ln_post.bias"])
            t.w_out_proj = mat(sd["visual.proj"].t())
            self.seq_len = grid * grid + 1
        else:
            t.width = sd["ln_final.weight"].shape[0]
            t.out_dim = sd["text_projection"].shape[1]
            t.vocab, t.context = sd["token_embedding.weight"].shape[0], sd["positional_embedding"].shape[0]
            t.tok_emb = vec(sd["token_embedding.weight"])
            t.pos_emb = vec(sd["positional_embedding"])
            t.ln_out_gain, t.ln_out_bias = vec(sd["ln_final.weight"]), vec(sd["ln_final.bias"])
            t.w_out_proj = mat(sd["text_projection"].t())
            t.eot_id = 49407                                                     # hard-coded in the reference (model.py:384)
            self.seq_len = None
        t.heads = t.width // 64                                                  # model.py:300,465
        self.c = t
        self._ws: Optional[torch.Tensor] = None

    def workspace(self, batch: int, seq_len: int, device) -> torch.Tensor:
        need = _lib.lib().cmh_encoder_workspace_bytes(ctypes.byref(self.c), batch, seq_len)
        if need <= 0:
            raise _lib.CmhError("cmh_encoder_workspace_bytes failed for batch %d, seq_len %d" % (batch, seq_len))
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws


class ClipBackbone(torch.nn.Module):
    """Drop-in for the ``backbone`` attribute of the reference models (``models/base.py:18-31`` -> ``build_clip``)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], return_patches: bool = False, device="cuda"):
        super().__init__()
        self.return_patches = return_patches
        self.device_ = torch.device(device)
        # Normalize(mean, std) of dataset/transformer_dataset.py:40,44 — applied on the GPU when images arrive as uint8
        self.pixel_mean = (0.48145466, 0.4578275, 0.40821073)
        self.pixel_std = (0.26862954, 0.26130258, 0.27577711)
        if self.device_.type != "cuda":
            raise _lib.CmhError("ClipBackbone needs a CUDA device (there is no CPU path)")
        skip = ("input_resolution", "context_length", "vocab_size")              # model.py:472-474
        self._sd = {k: v.detach().float().cpu() for k, v in state_dict.items() if k not in skip}
        if "visual.proj" not in self._sd:
            raise _lib.CmhError("only ViT backbones are supported (no visual.proj in the state dict)")   # model.py:444
        self.refresh()

    # ---- weights ---------------------------------------------------------------------------------------------------
    def refresh(self) -> None:
        """(Re)build the bf16/fp32 device copies from the fp32 master ``state_dict``."""
        with torch.cuda.device(self.device_):
            self.visual = _DeviceTower(self._sd, "visual.", self.device_)
            self.text = _DeviceTower(self._sd, "", self.device_)
        self.embed_dim = int(self.text.c.out_dim)

    def state_dict(self, *a, **k):  # reference keys (torch.save / load compatible, runners/base.py:379-384)
        return dict(self._sd)

    def load_state_dict(self, state_dict, strict: bool = True):
        missing = [k for k in self._sd if k not in state_dict]
        if strict and missing:
            raise KeyError("missing keys: %s" % missing[:5])
        for k in self._sd:
            if k in state_dict:
                self._sd[k] = state_dict[k].detach().float().cpu()
        self.refresh()

    @property
    def dtype(self):
        return torch.float32   # what callers feed (`image.type(self.dtype)`, model.py:366-371); arithmetic is bf16/fp32

    # ---- the two encoders ------------------------------------------------------------------------------------------
    def encode_image_raw(self, image: torch.Tensor, want_tokens: bool, want_attn: bool):
        """-> (cls [B, E], tokens [B, L, E] or None, attn [B, L-1] or None): the C ABI's own (sample-major) layouts."""
        tw = self.visual
        c = tw.c
        if image.dim() != 4 or image.shape[1] != 3 or image.shape[2] != c.resolution or image.shape[3] != c.resolution:
            raise ValueError("expected images [B, 3, %d, %d], got %s" % (c.resolution, c.resolution, tuple(image.shape)))
        u8 = image.dtype == torch.uint8   # raw pixels: /255 and Normalize(mean, std) happen on the GPU (self.pixel_mean / pixel_std)
        image = image.to(device=self.device_, dtype=torch.uint8 if u8 else torch.float32).contiguous()
        B, L, E = image.shape[0], tw.seq_len, c.out_dim
        with torch.cuda.device(self.device_):
            ws = tw.workspace(B, L, self.device_)
            cls = torch.empty((B, E), dtype=torch.float32, device=self.device_)
            tokens = torch.empty((B, L, E), dtype=torch.float32, device=self.device_) if want_tokens else None
            attn = torch.empty((B, L - 1), dtype=torch.float32, device=self.device_) if want_attn else None
            outs = (cls.data_ptr(), None if tokens is None else tokens.data_ptr(), None if attn is None else attn.data_ptr(), _stream())
            if u8:
                mean, std = (ctypes.c_float * 3)(*self.pixel_mean), (ctypes.c_float * 3)(*self.pixel_std)
                _lib.check(_lib.lib().cmh_encode_image_u8(ctypes.byref(c), image.data_ptr(), mean, std, B, ws.data_ptr(), ws.numel(), *outs))
            else:
                _lib.check(_lib.lib().cmh_encode_image(ctypes.byref(c), image.data_ptr(), B, ws.data_ptr(), ws.numel(), *outs))
        return cls, tokens, attn

    def encode_image(self, image: torch.Tensor):
        cls, tokens, attn = self.encode_image_raw(image, self.return_patches, self.return_patches)
        if self.return_patches:   # model.py:262-267
            return cls, tokens[:, 1:].permute(1, 0, 2), attn
        return cls

    def encode_text_raw(self, text: torch.Tensor, key_padding_mask: Optional[torch.Tensor], want_tokens: bool, want_attn: bool):
        """-> (eos [B, E], tokens [B, L, E] | None, attn [B, L] | None, new_mask uint8 [B, L] | None)."""
        tw = self.text
        c = tw.c
        if text.dim() != 2:
            raise ValueError("expected token ids [B, L], got %s" % (tuple(text.shape),))
        text = text.to(device=self.device_, dtype=torch.int64).contiguous()
        B, L, E = text.shape[0], text.shape[1], c.out_dim
        pad = None
        if key_padding_mask is not None:
            pad = key_padding_mask.to(device=self.device_).to(torch.uint8).contiguous()
            if pad.shape != text.shape:
                raise ValueError("key_padding_mask must have the shape of text")
        with torch.cuda.device(self.device_):
            ws = tw.workspace(B, L, self.device_)
            eos = torch.empty((B, E), dtype=torch.float32, device=self.device_)
            tokens = torch.empty((B, L, E), dtype=torch.float32, device=self.device_) if want_tokens else None
            attn = torch.empty((B, L), dtype=torch.float32, device=self.device_) if want_attn else None
            newmask = torch.empty((B, L), dtype=torch.uint8, device=self.device_) if (want_tokens and pad is not None) else None
            _lib.check(_lib.lib().cmh_encode_text(
                ctypes.byref(c), text.data_ptr(), None if pad is None else pad.data_ptr(), B, L, ws.data_ptr(), ws.numel(),
                eos.data_ptr(), None if tokens is None else tokens.data_ptr(), None if attn is None else attn.data_ptr(),
                None if newmask is None else newmask.data_ptr(), _stream()))
        return eos, tokens, attn, newmask

    def encode_text(self, text: torch.Tensor, key_padding_mask: Optional[torch.Tensor] = None):
        rp = self.return_patches
        eos, tokens, attn, newmask = self.encode_text_raw(text, key_padding_mask, rp, rp)
        if rp:   # model.py:394-395
            return eos, tokens.permute(1, 0, 2), attn, (None if newmask is None else newmask.bool())
        return eos
