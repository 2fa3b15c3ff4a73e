"""CPU: the drop-in bindings against the REAL reference packages (imported from /root/reference in a subprocess, with the three
absent third-party modules stubbed as SURVEY.md §8(c) lists).  Skipped where the reference tree is not present (the GPU box)."""
import os
import subprocess
import sys
import textwrap

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "runners")), reason="reference tree not present")

PRELUDE = textwrap.dedent("""
    import sys, types, zipfile, xml.etree.ElementTree as ET
    sys.path.insert(0, %r); sys.path.insert(0, %r)
    for name, attrs in (("ftfy", {"fix_text": lambda s: s}), ("termcolor", {"colored": lambda s, *a, **k: s})):
        m = types.ModuleType(name); m.__dict__.update(attrs); sys.modules[name] = m
    xl = types.ModuleType("xlrd")                      # models/DSPH/DSPH.py reads loss/codetable.xlsx with xlrd==1.2.0
    class _Sheet:
        def __init__(self, rows): self.rows = rows
        def row(self, i): return [types.SimpleNamespace(value=v) for v in self.rows[i]]
    class _Book:
        def __init__(self, path):
            z = zipfile.ZipFile(path)
            ns = {"m": "http://schemas.openxmlformats.org/spreadsheetml/2006/main"}
            rows = []
            for r in ET.fromstring(z.read("xl/worksheets/sheet1.xml")).iter("{%%s}row" %% ns["m"]):
                rows.append([float(c.find("m:v", ns).text) if c.find("m:v", ns) is not None else 0.0 for c in r])
            self._s = _Sheet(rows)
        def sheets(self): return [self._s]
        def sheet_by_index(self, i): return self._s
    xl.open_workbook = lambda path: _Book(path)
    sys.modules["xlrd"] = xl
    import torch
""") % (REF, ROOT)


def run(body: str):
    out = subprocess.run([sys.executable, "-c", PRELUDE + textwrap.dedent(body)], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    return out.stdout


def test_install_after_the_reference_imports_rebinds_every_by_value_import():
    """main.py:6-7 imports models and runners BEFORE user code runs; runners/base.py:5 binds calc_map_k by value at import time."""
    out = run("""
        import common.calc_utils as ref_cu
        import runners.base, models
        dcmht_mod = sys.modules["models.DCMHT.DCMHT"]   # (the package re-exports the class under the same name)
        original = ref_cu.calc_map_k
        assert runners.base.calc_map_k is original
        from clip_based_cross_modal_hash_b200 import calc_utils as cmh
        cmh.install_into_reference()
        assert ref_cu.calc_map_k is cmh.calc_map_k and ref_cu.calc_hammingDist is cmh.calc_hammingDist
        assert runners.base.calc_map_k is cmh.calc_map_k            # what BaseTrainer.__init__ binds at runners/base.py:78
        for name in ("cosine_similarity", "euclidean_similarity", "calc_label_sim"):
            if hasattr(dcmht_mod, name):
                assert getattr(dcmht_mod, name) is getattr(cmh, name), name
        mith_runner = sys.modules["runners.MITH.runner"]
        assert mith_runner.calc_label_sim is cmh.calc_label_sim     # runners/MITH/runner.py:5,87
        print("rebinding ok")
    """)
    assert "rebinding ok" in out


def test_install_before_the_reference_imports_also_works():
    out = run("""
        from clip_based_cross_modal_hash_b200 import calc_utils as cmh
        cmh.install_into_reference()
        import runners.base
        assert runners.base.calc_map_k is cmh.calc_map_k
        print("early install ok")
    """)
    assert "early install ok" in out


def test_similarity_helpers_keep_the_training_gradient():
    """models/DCMHT/DCMHT.py:78 calls the similarity helper on hash outputs that require grad: the shim must not cut the graph."""
    out = run("""
        import common.calc_utils as ref_cu
        reference = {n: getattr(ref_cu, n) for n in ("cosine_similarity", "euclidean_similarity")}
        from clip_based_cross_modal_hash_b200 import calc_utils as cmh
        cmh.install_into_reference()
        g = torch.Generator().manual_seed(0)
        for name in ("cosine_similarity", "euclidean_similarity"):
            a = torch.randn(6, 16, generator=g, requires_grad=True)
            b = torch.randn(5, 16, generator=g, requires_grad=True)
            s = getattr(ref_cu, name)(a, b)
            assert s.requires_grad
            s.sum().backward()
            a2, b2 = a.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
            reference[name](a2, b2).sum().backward()
            assert torch.allclose(a.grad, a2.grad) and torch.allclose(b.grad, b2.grad), name
        print("gradients ok")
    """)
    assert "gradients ok" in out


def test_registry_returns_the_b200_models_for_build_model():
    """BaseTrainer.build_model (runners/base.py:98-102) resolves the model class through the registry by `arch`."""
    out = run("""
        import models                                    # the reference package: its classes register themselves
        from common.register import registry
        ref_dsph = registry.get_model_class("DSPH")
        from clip_based_cross_modal_hash_b200 import models as cmh_models
        cmh_models.register_into_reference()
        assert registry.get_model_class("DSPH") is cmh_models.DSPH and registry.get_model_class("DSPH") is not ref_dsph
        assert registry.get_model_class("DCMHT") is cmh_models.DCMHT and registry.get_model_class("MITH") is cmh_models.MITH
        import inspect
        sig = inspect.signature(cmh_models.DSPH.from_config)
        assert list(sig.parameters)[:3] == ["cfg", "output_dim", "train_num"]       # runners/base.py:102 call shape
        for attr in ("encode_image", "encode_text", "forward", "state_dict", "load_state_dict", "float", "to", "freezen", "unfreezen"):
            assert hasattr(cmh_models.DSPH, attr), attr
        print("registry ok")
    """)
    assert "registry ok" in out
