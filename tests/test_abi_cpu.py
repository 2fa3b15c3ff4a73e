"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol include/cmh.h declares;
host-only entry points (geometry) behave; the product refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from clip_based_cross_modal_hash_b200 import _lib, calc_utils as cu, retrieval as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols(header="cmh.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cmh_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libcmh.so lacks %s" % n
    debug = _header_symbols("cmh_debug.h")      # tuning / trace hooks live in their own header, outside the product ABI
    assert debug and not set(debug) & set(names)
    for n in debug:
        assert hasattr(lib, n), "libcmh.so lacks %s" % n
    assert sorted(_lib.PROTOTYPES) == sorted(names + debug)  # the ctypes table covers both headers exactly
    assert lib.cmh_abi_version() == 2


def test_no_torch_or_cuda_driver_link_dependency():
    import subprocess

    out = subprocess.run(["ldd", _lib.SO_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libcuda.so" not in out and "libc10" not in out


def test_word_counts_and_errors():
    lib = _lib.lib()
    assert [lib.cmh_code_words(b) for b in (1, 16, 32, 33, 64, 65, 96, 97, 128)] == [1, 1, 1, 2, 2, 4, 4, 4, 4]
    assert lib.cmh_code_words(129) < 0 and lib.cmh_code_words(0) < 0
    assert [lib.cmh_label_words(c) for c in (0, 1, 24, 32, 33, 80, 128)] == [0, 1, 1, 1, 2, 4, 4]
    assert lib.cmh_label_words(129) < 0
    p = _lib.Plan()
    rc = lib.cmh_make_plan(0, 10, 10, 64, 80, 0, ctypes.byref(p))
    assert rc == -1 and b"Q > 0" in lib.cmh_last_error()
    with pytest.raises(_lib.CmhError):
        _lib.make_plan(10, 100, 256, 10)


@pytest.mark.parametrize("Q,N,K,C", [(1, 1, 16, 1), (1000, 5000, 16, 24), (5000, 117000, 64, 80),
                                     (2100, 190000, 128, 21), (10000, 1000000, 64, 80), (10000, 125000, 32, 0)])
def test_plan_geometry(Q, N, K, C):
    p = _lib.make_plan(Q, N, K, C, target_blocks=148 * 16)
    assert p.Qpad % 128 == 0 and 0 <= p.Qpad - Q < 128 and p.bins == K + 1
    assert p.chunk_items % 512 == 0 and p.chunk_items <= 65024          # 16-bit packed counters cannot overflow
    assert p.nchunks * p.chunk_items >= N and (p.nchunks - 1) * p.chunk_items < max(N, 1)
    assert p.hist_elems == p.nchunks * p.bins * p.Qpad and p.ap_elems == p.nchunks * p.Qpad
    assert p.workspace_bytes >= 4 * (3 * p.hist_elems + 2 * p.below_elems)
    # all ranks of a sharded run share the geometry of the largest shard
    p2 = _lib.make_plan(Q, max(N - 3, 0), K, C, N_geom=N, target_blocks=148 * 16)
    assert (p2.nchunks, p2.chunk_items) == (p.nchunks, p.chunk_items)


def test_shard_bounds():
    for n, w in [(117000, 8), (190000, 8), (1000000, 8), (301, 8), (7, 4), (5, 8)]:
        b = R.shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(lo % 4 == 0 for lo, hi in b if hi > lo)
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_fails_loudly_without_cuda():
    a = torch.ones(4, 16)
    lab = torch.ones(4, 3, dtype=torch.int64)
    with pytest.raises(_lib.CmhError):
        cu.calc_map_k(a, a, lab, lab, 2)
    with pytest.raises(_lib.CmhError):
        cu.calc_hammingDist(a, a)
    with pytest.raises(_lib.CmhError):
        R.pack_codes(a)


def test_encoder_workspace_queries_run_on_the_host():
    """cmh_encoder_workspace_bytes / cmh_head_mith_workspace_bytes are pure host arithmetic (no GPU needed)."""
    from clip_based_cross_modal_hash_b200 import encoder, models

    lib = _lib.lib()
    blocks = (encoder.BlockWeights * 12)()
    t = encoder.Tower()
    t.width, t.layers, t.heads, t.out_dim, t.blocks, t.patch, t.resolution = 768, 12, 12, 512, blocks, 32, 224
    one = lib.cmh_encoder_workspace_bytes(ctypes.byref(t), 1, 50)
    big = lib.cmh_encoder_workspace_bytes(ctypes.byref(t), 256, 50)
    # x fp32 + h + qkv + att + fc (bf16) per token row = 768 * (4 + 2 + 6 + 2 + 8) bytes
    assert big >= 256 * 50 * 768 * 22 and big < 256 * 50 * 768 * 24 and 0 < one < big
    assert lib.cmh_encoder_workspace_bytes(ctypes.byref(t), 0, 50) == 0
    h = models.MithHeadStruct()
    h.dim, h.nbits, h.mlp_layers, h.top_k = 512, 64, 2, 8
    h.transformer.width, h.transformer.layers, h.transformer.heads, h.transformer.out_dim = 512, 2, 8, 512
    h.transformer.blocks = (encoder.BlockWeights * 2)()
    assert lib.cmh_head_mith_workspace_bytes(ctypes.byref(h), 256, 49) > 256 * 49 * 512 * 4


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_encoders_fail_loudly_without_cuda():
    from clip_based_cross_modal_hash_b200 import encoder, models, synth

    sd = synth.clip_state_dict(synth.TINY, seed=1)
    with pytest.raises(_lib.CmhError):
        encoder.ClipBackbone(sd, device="cpu")
    with pytest.raises((_lib.CmhError, RuntimeError, AssertionError)):   # torch refuses the CUDA allocation first
        models.DSPH(sd, synth.dsph_head_state_dict(synth.TINY["embed_dim"], 16))
    with pytest.raises(_lib.CmhError):
        models.hyp_loss(torch.zeros(4, 16), torch.zeros(4, 16), torch.ones(4, 3), torch.zeros(3, 16), 0.0)


def test_synth_generators_are_deterministic_and_shaped_like_the_reference():
    from clip_based_cross_modal_hash_b200 import synth

    a, b = synth.clip_state_dict(synth.TINY, seed=3), synth.clip_state_dict(synth.TINY, seed=3)
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    assert a["visual.conv1.weight"].shape == (128, 3, 32, 32) and a["visual.positional_embedding"].shape == (50, 128)
    assert a["text_projection"].shape == (128, 128) and a["token_embedding.weight"].shape[0] == synth.TINY["vocab_size"]
    text, pad = synth.random_captions(9, seed=2)
    assert text.shape == (9, 32) and bool((text.argmax(dim=-1) >= 3).all()) and bool((text[:, 0] == 49406).all())
    assert torch.equal(pad, text == 0) and bool((text.max(dim=-1).values == 49407).all())
    m = synth.mith_head_state_dict(512, 16, seed=1)
    assert all(torch.equal(m["gcl_i." + k[6:]], v) for k, v in m.items() if k.startswith("gcl_t."))
