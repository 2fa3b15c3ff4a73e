"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol include/cmh.h declares;
host-only entry points (geometry) behave; the product refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from clip_based_cross_modal_hash_b200 import _lib, calc_utils as cu, retrieval as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "cmh.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cmh_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libcmh.so lacks %s" % n
    assert sorted(_lib.PROTOTYPES) == names  # the ctypes table covers the header exactly
    assert lib.cmh_abi_version() == 1


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
