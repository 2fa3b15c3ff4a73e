"""GPU: the tail of the DSPH training step (csrc/cmh_train.cu, optim.py) — fused BertAdam / SGD against the reference's optimiser
classes (golden), HyP gradients against autograd of the reference module (golden), head backward against torch autograd, and a
frozen-backbone training step against the same step done with torch autograd + the optimiser oracle."""
import os

import numpy as np
import pytest
import torch

from clip_based_cross_modal_hash_b200 import models, optim, synth
from oracle import hyp_port, optimizer_port as op
from tests._hyp_cases import CASES, inputs
from tests._opt_cases import OPT_CASES, opt_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "optimizer_golden.npz"))


@pytest.mark.parametrize("case", OPT_CASES, ids=[c[0] for c in OPT_CASES])
def test_fused_bert_adam_matches_reference(case):
    name, shapes, steps, kw = case
    params, grads = opt_inputs(shapes, steps, seed=7)
    ps = [torch.nn.Parameter(p.to(DEV)) for p in params]
    half = len(ps) // 2
    opt = optim.FusedBertAdam([{"params": ps[:half], "lr": kw["lr"] * 0.01}, {"params": ps[half:], "lr": kw["lr"]}], **kw)
    assert opt.get_lr() == []
    for s in range(steps):
        for p, g in zip(ps, grads[s]):
            p.grad = g.to(DEV).clone()
        opt.step()
    for i, p in enumerate(ps):
        assert np.allclose(p.detach().cpu().numpy(), G["%s/p%d" % (name, i)], rtol=2e-5, atol=1e-6), i
        assert np.allclose(opt.state[p]["next_m"].cpu().numpy(), G["%s/m%d" % (name, i)], rtol=2e-5, atol=1e-7)
        assert np.allclose(opt.state[p]["next_v"].cpu().numpy(), G["%s/v%d" % (name, i)], rtol=2e-5, atol=1e-9)
        assert opt.state[p]["step"] == steps
    assert len(opt.get_lr()) == len(ps)


def test_fused_sgd_matches_torch():
    params, grads = opt_inputs([(80, 64)], 3, seed=9)
    p = torch.nn.Parameter(params[0].to(DEV))
    sgd = optim.FusedSGD([p], lr=0.02, momentum=0.9, weight_decay=0.0005)
    for s in range(3):
        p.grad = grads[s][0].to(DEV).clone()
        sgd.step()
    assert np.allclose(p.detach().cpu().numpy(), G["sgd/p"], rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("i", range(len(CASES)), ids=[c[0] for c in CASES])
def test_hyp_gradients_match_reference_autograd(i):
    name, B, K, C, thr, alpha, dens = CASES[i]
    x, y, label, proxies = inputs(B, K, C, dens, 100 + i)
    loss, dx, dy, dp = optim.hyp_loss_and_grad(x.to(DEV), y.to(DEV), label, proxies.to(DEV), thr, alpha)
    want = hyp_port.hyp_loss(x, y, label, proxies, thr, alpha)
    assert abs(float(loss) - float(want)) <= 1e-5 * max(1.0, abs(float(want)))
    for got, key in ((dx, "dx"), (dy, "dy"), (dp, "dp")):
        ref = G["hyp/%s/%s" % (name, key)]
        assert np.allclose(got.cpu().numpy(), ref, rtol=2e-4, atol=1e-6 + 1e-4 * np.abs(ref).max()), key
    # the autograd wrapper: loss.backward() fills the same gradients, scaled by the upstream gradient
    mod = optim.HypLoss(C, K, 0, alpha, thr, device=DEV)
    with torch.no_grad():
        mod.proxies.copy_(proxies.to(DEV))
    xg, yg = x.to(DEV).requires_grad_(), y.to(DEV).requires_grad_()
    (2.0 * mod(xg, yg, label)).backward()
    assert torch.allclose(xg.grad, 2 * dx) and torch.allclose(yg.grad, 2 * dy) and torch.allclose(mod.proxies.grad, 2 * dp)


def test_head_backward_matches_autograd():
    g = torch.Generator().manual_seed(3)
    B, D, K = 37, 96, 24
    feat, W, b = torch.randn(B, D, generator=g), torch.randn(K, D, generator=g) * 0.1, torch.randn(K, generator=g) * 0.1
    dy = torch.randn(B, K, generator=g)
    f, w, bb = feat.clone().requires_grad_(), W.clone().requires_grad_(), b.clone().requires_grad_()
    y = torch.tanh(f @ w.t() + bb)
    y.backward(dy)
    dW, db, dfeat = optim.linear_tanh_backward(feat.to(DEV), y.detach().to(DEV), dy.to(DEV), W.to(DEV), want_dfeat=True)
    assert torch.allclose(dW.cpu(), w.grad, rtol=1e-4, atol=1e-5) and torch.allclose(db.cpu(), bb.grad, rtol=1e-4, atol=1e-5)
    assert torch.allclose(dfeat.cpu(), f.grad, rtol=1e-4, atol=1e-5)


def test_frozen_backbone_training_step_matches_autograd_plus_oracle_optimiser():
    nbits, C, B = 32, 12, 16
    sd = synth.clip_state_dict(synth.TINY, seed=8)
    model = models.DSPH(sd, synth.dsph_head_state_dict(synth.TINY["embed_dim"], nbits, seed=9))
    kw = dict(lr=1e-3, t_total=10, warmup=0.1, schedule="warmup_cosine", b1=0.9, b2=0.98, e=1e-6, weight_decay=0.2, max_grad_norm=1.0)
    tr = optim.DsphHeadTrainer(model, numclass=C, threshold=0.1, alpha=0.8, hyp_lr=0.02, hyp_momentum=0.9, hyp_weight_decay=5e-4, **kw)
    image = synth.random_images(B, seed=1)
    text, _ = synth.random_captions(B, seed=2, vocab=synth.TINY["vocab_size"])
    label = synth.random_labels(B, C, 3, p=0.25)
    # the reference-shaped step on the same (GPU-computed) features: torch autograd + optimiser oracle
    fi, ft = model.backbone.encode_image(image).cpu(), model.backbone.encode_text(text).cpu()
    P = {m: [t.detach().cpu().clone().requires_grad_() for t in tr.params[m]] for m in ("img", "txt")}
    prox = tr.hyp.proxies.detach().cpu().clone().requires_grad_()
    st = {m: [{"m": torch.zeros_like(t), "v": torch.zeros_like(t)} for t in P[m]] for m in P}
    buf = torch.zeros_like(prox)
    losses_ref, losses = [], []
    for s in range(3):
        yi = torch.tanh(fi @ P["img"][0].t() + P["img"][1])
        yt = torch.tanh(ft @ P["txt"][0].t() + P["txt"][1])
        loss = hyp_port.hyp_loss(yi, yt, label, prox, 0.1, 0.8)
        for t in P["img"] + P["txt"] + [prox]:
            t.grad = None
        loss.backward()
        losses_ref.append(float(loss))
        with torch.no_grad():
            flat = P["img"] + P["txt"]
            op.bert_adam_step(flat, [t.grad for t in flat], st["img"] + st["txt"], s, kw["lr"], kw["warmup"], kw["t_total"], kw["schedule"],
                              kw["b1"], kw["b2"], kw["e"], kw["weight_decay"], kw["max_grad_norm"])
            op.sgd_momentum_step([prox], [prox.grad], [buf], s == 0, 0.02, 0.9, 5e-4)
        losses.append(float(tr.step(image, text, label)))
    assert np.allclose(losses, losses_ref, rtol=2e-4), (losses, losses_ref)
    for m in ("img", "txt"):
        for got, want in zip(tr.params[m], P[m]):
            assert torch.allclose(got.detach().cpu(), want.detach(), rtol=5e-3, atol=2e-5)
    assert torch.allclose(tr.hyp.proxies.detach().cpu(), prox.detach(), rtol=1e-3, atol=1e-5)
    # the trained head is what the model now evaluates with, and it round-trips through the checkpoint view
    tr.sync_state_dict()
    codes = model.encode_image_packed(image)
    model2 = models.DSPH(sd, synth.dsph_head_state_dict(synth.TINY["embed_dim"], nbits, seed=1))
    model2.load_state_dict(model.state_dict())
    assert torch.equal(model2.encode_image_packed(image), codes)
