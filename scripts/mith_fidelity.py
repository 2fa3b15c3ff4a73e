"""tokens_hash fidelity of the MITH head vs the reference golden, plain bf16 vs split-precision token path."""
import os, sys
import numpy as np, torch
sys.path.insert(0, '/root/repo')
from clip_based_cross_modal_hash_b200 import models, synth
Z = np.load('/root/repo/tests/golden/mith_golden.npz')
def inputs(B, L, seed, padded):
    g = torch.Generator().manual_seed(seed)
    cls = torch.randn((B, 512), generator=g); tokens = torch.randn((L, B, 512), generator=g); mask = None
    if padded:
        lens = torch.randint(2, L + 1, (B,), generator=g); mask = torch.arange(L)[None, :] >= lens[:, None]
    return cls, tokens, mask
for split in (False, True):
    for nbits in (16, 64):
        head = models.MithHashLayer(synth.mith_head_state_dict(512, nbits, seed=51), 'cuda', split_precision=split)
        for m in ('img', 'txt'):
            cls, tokens, mask = inputs(5, 49, 61, False) if m == 'img' else inputs(6, 32, 62, True)
            r = head.encode_img(cls.cuda(), tokens.cuda()) if m == 'img' else head.encode_txt(cls.cuda(), tokens.cuda(), mask.cuda())
            th = r[2].cpu().numpy(); ch = r[1].cpu().numpy()
            want = Z['mith%d/%s_tok_hash' % (nbits, m)]
            err = np.abs(th - want)
            code = np.sign(ch + th); ref = np.sign(Z['mith%d/%s_cls_hash' % (nbits, m)] + want)
            print('split=%s K=%d %s: tokens_hash mean err %.4f max %.3f  frac>6e-2 %.4f  code bits differing %.4f' % (
                split, nbits, m, err.mean(), err.max(), (err > 6e-2).mean(), (code != ref).mean()))
