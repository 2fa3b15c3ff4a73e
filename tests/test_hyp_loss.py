"""HyP.forward (models/DSPH/loss/HyP.py): oracle vs the reference's golden values (CPU) and the CUDA kernel vs both (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import hyp_port
from tests._hyp_cases import CASES, inputs

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hyp_golden.npz"))


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_matches_reference_golden(i):
    name, B, K, C, thr, alpha, dens = CASES[i]
    x, y, label, proxies = inputs(B, K, C, dens, 100 + i)
    got = hyp_port.hyp_loss(x, y, label, proxies, thr, alpha)
    assert abs(float(got) - float(Z[name])) <= 2e-6 * max(1.0, abs(float(Z[name])))


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(CASES)))
def test_cuda_matches_reference_golden(i):
    from clip_based_cross_modal_hash_b200 import models

    name, B, K, C, thr, alpha, dens = CASES[i]
    x, y, label, proxies = inputs(B, K, C, dens, 100 + i)
    got = models.hyp_loss(x.cuda(), y.cuda(), label, proxies, thr, alpha)
    assert got.dtype == torch.float32 and got.is_cuda and got.dim() == 0
    # fp32 dot products + fp64 sums on the GPU vs fp32 sums in the reference: 1e-5 relative
    assert abs(float(got) - float(Z[name])) <= 1e-5 * max(1.0, abs(float(Z[name]))), (float(got), float(Z[name]))


@pytest.mark.gpu
def test_dsph_object_function_value():
    """models.DSPH.object_function (models/DSPH/DSPH.py:78-82): the loss value through the model object, proxies from a checkpoint."""
    from clip_based_cross_modal_hash_b200 import models, synth

    name, B, K, C, thr, alpha, dens = CASES[1]
    x, y, label, proxies = inputs(B, K, C, dens, 101)
    model = models.DSPH(synth.clip_state_dict(synth.TINY, seed=1), synth.dsph_head_state_dict(synth.TINY["embed_dim"], K, seed=2),
                        threshold=thr, alpha=alpha)
    with pytest.raises(Exception):
        model.object_function(x.cuda(), y.cuda(), label)
    full = model.state_dict()
    full["hyp.proxies"] = proxies
    model.load_state_dict(full)
    loss, parts = model.object_function(x.cuda(), y.cuda(), label)
    assert abs(float(loss) - float(Z[name])) <= 1e-5 * abs(float(Z[name])) and "All loss" in parts
    assert model.freezen() is None and model.unfreezen() is None
