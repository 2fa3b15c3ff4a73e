"""GPU: the evaluation half of BaseTrainer on packed codes (runner.PackedEvaluation.valid / test, save_mat) on the TINY
configuration with fake loaders — values against four separate calc_map_k calls on the reference-format +-1 float codes."""
import os

import numpy as np
import pytest
import scipy.io as scio
import torch

from clip_based_cross_modal_hash_b200 import calc_utils as cu
from clip_based_cross_modal_hash_b200 import models, retrieval as R, runner, synth

pytestmark = pytest.mark.gpu


def _loader(n, B, seed):
    out = []
    for i, lo in enumerate(range(0, n, B)):
        b = min(B, n - lo)
        text, pad = synth.random_captions(b, seed=seed + 100 + i, vocab=synth.TINY["vocab_size"])
        out.append((synth.random_images(b, seed=seed + i).pin_memory(), text.pin_memory(), pad, None, torch.arange(lo, lo + b)))
    return out


def test_valid_test_and_save_mat_on_packed_codes(tmp_path):
    nbits, C, nq, nr = 32, 12, 24, 90
    model = models.DSPH(synth.clip_state_dict(synth.TINY, seed=8), synth.dsph_head_state_dict(synth.TINY["embed_dim"], nbits, seed=9))
    ql, rl = synth.random_labels(nq, C, 1), synth.random_labels(nr, C, 2)
    saved = []
    ev = runner.PackedEvaluation(model, _loader(nq, 8, 10), _loader(nr, 16, 20), ql, rl, nq, nr, save_dir=str(tmp_path), epochs=3,
                                 save_model=lambda d, e: saved.append(e), top_k=10)
    maps = ev.valid(epoch=0, k=None)
    # reference-format check: unpack to +-1 floats and evaluate with four independent calc_map_k calls
    qi, qt = models.get_code(model, ev.query_loader, nq)
    ri, rt = models.get_code(model, ev.retrieval_loader, nr)
    f = lambda c: R.unpack_codes(c, nbits)
    want = (cu.calc_map_k(f(qi), f(rt), ql, rl), cu.calc_map_k(f(qt), f(ri), ql, rl), cu.calc_map_k(f(qi), f(ri), ql, rl),
            cu.calc_map_k(f(qt), f(rt), ql, rl))
    for g, w in zip(maps, want):
        assert g.dtype == torch.float32 and g.device.type == "cpu" and float(g) == float(w)
    assert ev.best_epoch_i == 0 and ev.best_epoch_t == 0 and float(ev.max_mapi2t) == float(maps[0]) and saved == [0, 0]
    maps2 = ev.valid(epoch=1, k=None)                      # same model: no improvement, bookkeeping untouched
    assert [float(a) for a in maps2] == [float(a) for a in maps] and ev.best_epoch_i == 0 and saved == [0, 0]
    for name in ("i2t-best.mat", "t2i-best.mat", "last.mat"):
        m = scio.loadmat(os.path.join(str(tmp_path), "mat_files", name))
        assert set(("q_img", "q_txt", "r_img", "r_txt", "q_l", "r_l")) <= set(m)
        assert m["q_img"].shape == (nq, nbits) and m["r_txt"].shape == (nr, nbits) and set(np.unique(m["q_img"])) <= {-1.0, 1.0}
        assert np.array_equal(m["r_img"], f(ri).cpu().numpy()) and np.array_equal(m["q_l"], ql.numpy())
    t = ev.test()
    want_k = cu.calc_map_k(f(qi), f(rt), ql, rl, 10)
    assert float(t[0]) == float(want_k) and os.path.exists(os.path.join(str(tmp_path), "mat_files", "test.mat"))


def test_output_dim_follows_a_loaded_checkpoint_and_proxies_survive_a_round_trip():
    clip = synth.clip_state_dict(synth.TINY, seed=8)
    E = synth.TINY["embed_dim"]
    m16 = models.DSPH(clip, synth.dsph_head_state_dict(E, 16, seed=1), proxies=torch.randn(5, 16))
    m64 = models.DSPH(clip, synth.dsph_head_state_dict(E, 64, seed=2))
    sd = m16.state_dict()
    assert "hyp.proxies" in sd
    m64.load_state_dict(sd)
    assert m64.output_dim == 16 and torch.equal(m64.proxies, m16.proxies.float())
    loader = _loader(6, 3, 5)
    a, _ = models.get_code(m64, loader, 6)
    b, _ = models.get_code(m16, loader, 6)
    assert torch.equal(a, b)
