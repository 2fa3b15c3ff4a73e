"""GPU parity: the CUDA evaluator (through the C ABI) against the golden vectors the reference produced
and against the oracles on seeded inputs.  Integer stages must be bit-exact; the fp32 mAP must be
bit-exact in parity mode (same host, same torch ops as the reference) and within 4e-7 relative in
device mode (fp64 accumulation of the reference's own fp32 quotients).
"""
import numpy as np
import pytest
import torch

from clip_based_cross_modal_hash_b200 import calc_utils as cu
from clip_based_cross_modal_hash_b200 import retrieval as R
from clip_based_cross_modal_hash_b200 import synth
from oracle import c_oracle, calc_utils_port as port, hamming_oracle as ho
from tests._golden import CASE_NAMES, Case, npz

pytestmark = pytest.mark.gpu
DEV = "cuda"
DEVICE_MODE_RTOL = 4e-7  # fp64-accumulated sum of the same fp32 terms vs the reference's fp32 running sums


def _u32(t):
    return t.cpu().numpy().view(np.uint32)


def _packed(c):
    qp = R.pack_codes(c.qB.to(DEV))
    gp = R.pack_codes(c.rB.to(DEV))
    qlp = R.pack_labels(c.qL.to(DEV))
    glp = R.pack_labels(c.rL.to(DEV))
    return qp, gp, qlp, glp


@pytest.mark.parametrize("name", CASE_NAMES)
def test_pack_matches_oracle(name):
    c = Case(name)
    bad = R.new_bad_counter(DEV)
    qp = _u32(R.pack_codes(c.rB.to(DEV), bad))
    want = ho.pack_codes(c.rB.numpy())
    assert np.array_equal(qp[:, : want.shape[1]], want) and not qp[:, want.shape[1]:].any()
    for dt in (torch.int64, torch.float32, torch.uint8, torch.int32, torch.bool):
        lp = _u32(R.pack_labels(c.rL.to(DEV).to(dt), bad))
        wl = ho.pack_labels(c.rL.numpy())
        assert np.array_equal(lp, wl[:, : lp.shape[1]]) and not wl[:, lp.shape[1]:].any()
    assert int(bad.item()) == 0
    back = R.unpack_codes(R.pack_codes(c.rB.to(DEV)), c.K).cpu()
    assert torch.equal(back, c.rB)


def test_pack_counts_bad_elements():
    codes = synth.random_codes(33, 48, 3)
    codes[0, 0] = 0.0
    codes[5, 47] = 0.5
    lab = synth.random_labels(10, 24, 4)
    lab[3, 3] = 2
    bad = R.new_bad_counter(DEV)
    R.pack_codes(codes.to(DEV), bad)
    assert int(bad.item()) == 2
    R.pack_labels(lab.to(DEV), bad)
    assert int(bad.item()) == 3


@pytest.mark.parametrize("name", CASE_NAMES)
def test_hamming_matrix_matches_reference(name):
    c = Case(name)
    got = cu.calc_hammingDist(c.qB.to(DEV), c.rB.to(DEV))
    assert got.is_cuda and got.dtype == torch.float32
    assert np.array_equal(got.sum(dim=1).cpu().numpy().astype(np.int64), c.hamm_rowsum)
    if c.hamm is not None:
        assert np.array_equal(got.cpu().numpy().astype(np.uint8), c.hamm)
    if c.Q * c.N <= 2_000_000:
        assert torch.equal(got.cpu(), port.calc_hammingDist(c.qB, c.rB))


def test_hamming_dist_reference_edge_cases():
    z = npz()
    q, r = torch.from_numpy(z["hd/q"]), torch.from_numpy(z["hd/r"])
    # codes containing 0 (sign_() of an exact zero) -> half-integer distances via the dense kernel
    assert np.array_equal(cu.calc_hammingDist(q, r).numpy(), z["hd/full"])
    # 1-D query is unsqueezed like the reference
    got = cu.calc_hammingDist(q[1].to(DEV), r.to(DEV))
    assert got.shape == (1, 50) and np.array_equal(got.cpu().numpy(), z["hd/one_d"])


@pytest.mark.parametrize("name", CASE_NAMES)
def test_map_integer_stages_bit_exact(name):
    c = Case(name)
    qp, gp, qlp, glp = _packed(c)
    res = R.map_k(qp, qlp, gp, glp, c.K, c.C, c.k, want_tindex=True)
    assert np.array_equal(res.total.cpu().numpy(), c.totals)
    assert np.array_equal(res.tsum.cpu().numpy(), c.tsums)
    tix = res.tindex.cpu().numpy()
    for q in range(c.Q):
        assert np.array_equal(tix[q, : c.totals[q]], c.tindex[q]), (name, q)
        assert not tix[q, c.totals[q]:].any()
    # device-mode mAP: same fp32 terms, fp64 accumulation
    want64 = ho.map_float64([t.astype(np.int64) for t in c.tindex])
    assert abs(res.map.item() - want64) <= 1e-12
    assert abs(res.map.item() - float(c.map_stable)) <= DEVICE_MODE_RTOL * max(1.0, abs(float(c.map_stable)))


@pytest.mark.parametrize("name", CASE_NAMES)
def test_calc_map_k_dropin(name):
    c = Case(name)
    want = port.calc_map_k(c.qB, c.rB, c.qL, c.rL, c.k, stable=True)  # same host, same torch ops
    got = cu.calc_map_k(c.qB, c.rB, c.qL, c.rL, c.k, mode="parity")   # CPU tensors in, like the reference
    assert isinstance(got, torch.Tensor) and got.dim() == 0 and got.dtype == torch.float32 and not got.is_cuda
    assert np.float32(got.item()) == np.float32(want.item())
    # vs the value recorded from the reference in the build container (another CPU may sum in another order)
    assert abs(got.item() - float(c.map_stable)) <= 2e-7
    got_dev = cu.calc_map_k(c.qB.to(DEV), c.rB.to(DEV), c.qL.to(DEV), c.rL.to(DEV), c.k)
    assert got_dev.dtype == torch.float32 and not got_dev.is_cuda
    assert abs(got_dev.item() - float(c.map_stable)) <= DEVICE_MODE_RTOL * max(1.0, abs(float(c.map_stable)))


def test_calc_map_k_parity_slabs(monkeypatch):
    c = Case("mid64_full")
    want = port.calc_map_k(c.qB, c.rB, c.qL, c.rL, c.k, stable=True)
    monkeypatch.setattr(cu, "_PARITY_SLAB_BYTES", 4 * 1000 * 5)  # force 5-query slabs
    got = cu.calc_map_k(c.qB, c.rB, c.qL, c.rL, c.k, mode="parity")
    assert np.float32(got.item()) == np.float32(want.item())


@pytest.mark.parametrize("name", CASE_NAMES)
def test_topk_matches_reference_order(name):
    c = Case(name)
    qp, gp, _, _ = _packed(c)
    kk = min(64, c.N)
    keys = R.topk(qp, gp, c.K, kk)
    dist, idx = R.split_keys(keys)
    assert np.array_equal(idx.cpu().numpy()[:, :kk], c.order_head[:, :kk])
    wd, wi = port.hamming_rank_topk(c.qB, c.rB, kk, stable=True) if c.Q * c.N <= 2_000_000 else (None, None)
    if wd is not None:
        assert torch.equal(dist.cpu().to(torch.float32), wd) and torch.equal(idx.cpu(), wi)
    # larger k through the C oracle, and k > N (empty slots = -1)
    for k in (1, 333, c.N, c.N + 5):
        keys = R.topk(qp, gp, c.K, k)
        d, i = (t.cpu().numpy() for t in R.split_keys(keys))
        od, oi = c_oracle.topk(_u32(qp)[:, : (c.K + 31) // 32], _u32(gp)[:, : (c.K + 31) // 32], c.K, k)
        assert np.array_equal(d, od) and np.array_equal(i, oi), (name, k)


def test_hamming_topk_dropin():
    c = Case("mid64")
    d, i = cu.hamming_topk(c.qB, c.rB, 100)
    wd, wi = port.hamming_rank_topk(c.qB, c.rB, 100, stable=True)
    assert torch.equal(d, wd) and torch.equal(i, wi)


# ---------------------------------------------------------------------------------------------------------
# sharded arithmetic on ONE GPU: run every shard's stages in turn and exchange by concatenation
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("name", ["mid64", "odd48", "wide128", "tiny16_full"])
def test_sharded_stages_match_single(name, world):
    c = Case(name)
    qp, gp, qlp, glp = _packed(c)
    st = R.CudaStages()
    bounds = R.shard_bounds(c.N, world)
    n_geom = max(hi - lo for lo, hi in bounds)
    plans = [st.make_plan(c.Q, hi - lo, c.K, c.C, n_geom) for lo, hi in bounds]
    hists = [st.hist(p, qp, qlp, gp[lo:hi], glp[lo:hi]) for p, (lo, hi) in zip(plans, bounds)]
    hist_all = torch.stack(hists)
    cap = max(int(c.totals.max()), 1)
    tindex = torch.zeros((c.Q, cap), dtype=torch.int32, device=DEV)
    parts = []
    totals_all = torch.stack([st.hist_totals(p, h) for p, h in zip(plans, hists)])    # what the ranks all-gather
    for r, (p, (lo, hi)) in enumerate(zip(plans, bounds)):
        sc = st.scan(p, hist_all, world, r, c.k)
        assert np.array_equal(sc["total"][: c.Q].cpu().numpy(), c.totals)
        # the reduced exchange (per-rank bucket totals only) gives bit-identical rank bases
        sc2 = st.scan_sharded(p, hists[r], totals_all, world, r, c.k)
        for key in ("within_all", "within_rel", "below_all", "below_rel", "tsum", "total", "thresh"):
            assert torch.equal(sc[key], sc2[key]), key
        mine = torch.zeros_like(tindex)
        parts.append(st.rank_map(p, qp, qlp, gp[lo:hi], glp[lo:hi], sc, mine, n_total=c.N))
        assert not ((tindex != 0) & (mine != 0)).any()  # each slot owned by exactly one shard
        tindex += mine
    tix = tindex.cpu().numpy()
    for q in range(c.Q):
        assert np.array_equal(tix[q, : c.totals[q]], c.tindex[q])
    ap, m = st.map_finish(plans[0], torch.stack(parts), sc["total"])
    assert abs(m.item() - float(c.map_stable)) <= DEVICE_MODE_RTOL
    # top-k: per-shard partials -> merge
    k = 100
    keys = []
    for p, (lo, hi) in zip(plans, bounds):
        pl = st.make_plan(c.Q, hi - lo, c.K, 0, n_geom)
        h = st.hist(pl, qp, None, gp[lo:hi], None)
        s = st.scan(pl, h, 1, 0, k, with_rel=False)
        keys.append(st.rank_topk(pl, qp, gp[lo:hi], s, k, lo))
    merged = st.topk_merge(torch.stack(keys))
    assert torch.equal(merged, R.topk(qp, gp, c.K, k))


# ---------------------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Q,N,K,C", [(1, 1, 16, 1), (1, 3, 32, 24), (3, 5, 64, 80), (130, 7, 128, 128),
                                     (2, 513, 16, 33), (129, 2051, 96, 65), (5, 1026, 8, 5)])
def test_small_and_ragged_shapes(Q, N, K, C):
    qB, rB = synth.random_codes(Q, K, 11), synth.random_codes(N, K, 12)
    qL, rL = synth.random_labels(Q, C, 13, p=0.3), synth.random_labels(N, C, 14, p=0.3)
    qp, gp = R.pack_codes(qB.to(DEV)), R.pack_codes(rB.to(DEV))
    qlp, glp = R.pack_labels(qL.to(DEV)), R.pack_labels(rL.to(DEV))
    res = R.map_k(qp, qlp, gp, glp, K, C, None, want_tindex=True)
    oq, og = ho.pack_codes(qB.numpy()), ho.pack_codes(rB.numpy())
    tix, totals, tsums = c_oracle.map_tindex(oq, ho.pack_labels(qL.numpy()), og, ho.pack_labels(rL.numpy()), K, None)
    assert np.array_equal(res.total.cpu().numpy(), totals) and np.array_equal(res.tsum.cpu().numpy(), tsums)
    got = res.tindex.cpu().numpy()
    for q in range(Q):
        assert np.array_equal(got[q, : totals[q]], tix[q, : totals[q]])
    k = min(N, 17)
    d, i = (t.cpu().numpy() for t in R.split_keys(R.topk(qp, gp, K, k)))
    od, oi = c_oracle.topk(oq, og, K, k)
    assert np.array_equal(d, od) and np.array_equal(i, oi)


def test_unaligned_gallery_views_use_fallback_copy():
    c = Case("odd32")  # W = 1: a one-row offset is 4-byte aligned only
    qp, gp, qlp, glp = _packed(c)
    full = R.map_k(qp, qlp, gp[1:], glp[1:], c.K, c.C, 20, want_tindex=True)
    oq, og = ho.pack_codes(c.qB.numpy()), ho.pack_codes(c.rB.numpy()[1:])
    tix, totals, _ = c_oracle.map_tindex(oq, ho.pack_labels(c.qL.numpy()), og, ho.pack_labels(c.rL.numpy()[1:]), c.K, 20)
    got = full.tindex.cpu().numpy()
    for q in range(c.Q):
        assert np.array_equal(got[q, : totals[q]], tix[q, : totals[q]])


def test_query_without_relevant_items_gives_nan_like_reference():
    qB, rB = synth.random_codes(4, 32, 1), synth.random_codes(100, 32, 2)
    qL = torch.zeros(4, 10, dtype=torch.int64)
    qL[:, 0] = 1
    rL = torch.zeros(100, 10, dtype=torch.int64)
    rL[:, 1] = 1
    rL[:50, 0] = 1
    qL[2] = 0
    qL[2, 5] = 1  # query 2 shares no class with anything
    want = port.calc_map_k(qB, rB, qL, rL, 10, stable=True)
    assert torch.isnan(want)
    assert torch.isnan(cu.calc_map_k(qB, rB, qL, rL, 10))
    assert torch.isnan(cu.calc_map_k(qB, rB, qL, rL, 10, mode="parity"))


def test_non_binary_codes_do_not_raise_in_calc_map_k_but_do_in_topk():
    """calc_map_k is a drop-in: the reference computes a value for codes containing 0 (common/calc_utils.py:51-56), so does the
    shim (dense path, both modes).  hamming_topk has no reference counterpart for such codes and rejects them."""
    qB, rB = synth.random_codes(4, 32, 1), synth.random_codes(100, 32, 2)
    qL, rL = synth.random_labels(4, 10, 3), synth.random_labels(100, 10, 4)
    bad = rB.clone()
    bad[7, 7] = 0.0
    want = float(port.calc_map_k(qB, bad, qL, rL, 10, stable=True))
    assert abs(float(cu.calc_map_k(qB, bad, qL, rL, 10)) - want) <= 4e-7
    assert abs(float(cu.calc_map_k(qB, bad, qL, rL, 10, mode="parity")) - want) <= 4e-7
    with pytest.raises(ValueError):
        cu.hamming_topk(qB, bad, 10)
    with pytest.raises(ValueError):
        cu.calc_map_k(qB, rB, qL, rL, 0)


def test_similarity_helpers_match_reference():
    z = npz()
    a, b = torch.from_numpy(z["sim/a"]), torch.from_numpy(z["sim/b"])
    la, lb = torch.from_numpy(z["sim/la"]), torch.from_numpy(z["sim/lb"])
    assert np.array_equal(cu.calc_label_sim(la, lb).numpy(), z["sim/label_sim"])
    assert np.array_equal(cu.calc_label_sim(la.long().to(DEV), lb.long().to(DEV)).cpu().numpy(), z["sim/label_sim_i64"])
    # fp32 tolerance: rtol 1e-5, atol 1e-6
    assert np.allclose(cu.cosine_similarity(a, b).numpy(), z["sim/cosine"], rtol=1e-5, atol=1e-6)
    assert np.allclose(cu.cosine_similarity(a.numpy(), b.numpy()), z["sim/cosine_np"], rtol=1e-5, atol=1e-6)
    assert np.allclose(cu.euclidean_similarity(a.to(DEV), b.to(DEV)).cpu().numpy(), z["sim/euclid"], rtol=1e-5, atol=1e-6)
    assert np.allclose(cu.euclidean_similarity(a.numpy(), b.numpy()), z["sim/euclid_np"], rtol=1e-5, atol=1e-6)
    ls, ws = cu.generate_weight_sim(la, la)
    assert np.array_equal(ls.numpy(), z["sim/weight_label"])
    assert np.allclose(ws.numpy(), z["sim/weight_sim"], rtol=1e-5, atol=1e-7)
    with pytest.raises(ValueError):
        cu.cosine_similarity(a, b.numpy())
    with pytest.raises(ValueError):
        cu.euclidean_similarity(a.numpy(), b)
    zero = a.clone()
    zero[3] = 0
    assert torch.isnan(cu.cosine_similarity(zero, b)[3]).all()  # no epsilon, like the reference


# ---------------------------------------------------------------------------------------------------------
# BASELINE.json sizes: oracle on a query subset + size-independent properties on the full result
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", ["C2", "C3"])
def test_full_size_map(cfg):
    s = synth.CONFIGS[cfg]
    Q, N, K, C = s["Q"], s["N"], s["K"], s["C"]
    qB, rB = synth.random_codes(Q, K, 21), synth.random_codes(N, K, 22)
    qL, rL = synth.random_labels(Q, C, 23), synth.random_labels(N, C, 24)
    qp, gp = R.pack_codes(qB.to(DEV)), R.pack_codes(rB.to(DEV))
    qlp, glp = R.pack_labels(qL.to(DEV)), R.pack_labels(rL.to(DEV))
    res = R.map_k(qp, qlp, gp, glp, K, C, None)
    sub = np.arange(0, Q, max(1, Q // 48))
    cap = 4096
    part = R.map_k(qp[sub], qlp[sub], gp, glp, K, C, None, want_tindex=True, tindex_cap=cap)
    W = (K + 31) // 32
    tix, totals, tsums = c_oracle.map_tindex(_u32(qp)[sub][:, :W], ho.pack_labels(qL.numpy()[sub]), _u32(gp)[:, :W],
                                             ho.pack_labels(rL.numpy()), K, None, cap=cap)
    assert np.array_equal(res.total.cpu().numpy()[sub], totals) and np.array_equal(res.tsum.cpu().numpy()[sub], tsums)
    assert np.array_equal(part.tindex.cpu().numpy(), tix)
    # a subset of queries is an independent problem: per-query AP must agree with the full run (the chunk
    # geometry differs, so the fp64 partial sums are added in another order)
    assert (part.ap - res.ap[torch.from_numpy(sub).to(DEV)]).abs().max().item() < 1e-12
    # properties: 0 < AP <= 1, totals <= N, mean of AP == map
    ap = res.ap.cpu().numpy()
    assert (ap > 0).all() and (ap <= 1).all()
    assert abs(ap.mean() - res.map.item()) < 1e-12


def test_full_size_topk_properties():
    Q, N, K, k = 2048, 1_000_000, 64, 1000  # C4 gallery, a fifth of its queries
    qB, rB = synth.random_codes(Q, K, 31), synth.random_codes(N, K, 32)
    qp, gp = R.pack_codes(qB.to(DEV)), R.pack_codes(rB.to(DEV))
    keys = R.topk(qp, gp, K, k)
    assert (keys[:, 1:] > keys[:, :-1]).all()  # strictly ascending (dist, index): sorted and duplicate-free
    dist, idx = R.split_keys(keys)
    assert int(idx.min()) >= 0 and int(idx.max()) < N
    # recompute every reported distance from the packed words
    a = qp.view(torch.int64)[:, None, 0] ^ gp.view(torch.int64)[idx, 0]
    lo = (a & 0xFFFFFFFF).to(torch.int32)
    hi = (a >> 32).to(torch.int32)
    pc = sum(((lo >> b) & 1) + ((hi >> b) & 1) for b in range(32))
    assert torch.equal(pc.to(torch.int32), dist)
    # nothing outside the list beats the list's last entry: check 16 queries with the C oracle
    od, oi = c_oracle.topk(_u32(qp)[:16, :2], _u32(gp)[:, :2], K, k)
    assert np.array_equal(dist[:16].cpu().numpy(), od) and np.array_equal(idx[:16].cpu().numpy(), oi)


@pytest.mark.parametrize("K", [16, 32, 64, 128])
def test_c4_full_size_topk_every_bit_width(K):
    """BASELINE.json configs[3] at its FULL size (10 000 x 1 000 000, top-1000) for every code length of the sweep:
    sortedness / uniqueness of all 10 000 result lists, and bit-exact (distance, index) lists for 16 queries against the C oracle."""
    s = synth.CONFIGS["C4-%d" % K]
    Q, N, k = s["Q"], s["N"], s["k"]
    qB, rB = synth.random_codes(Q, K, 41 + K), synth.random_codes(N, K, 42 + K)
    qp, gp = R.pack_codes(qB.to(DEV)), R.pack_codes(rB.to(DEV))
    del qB, rB
    keys = R.topk(qp, gp, K, k)
    assert (keys[:, 1:] > keys[:, :-1]).all()
    dist, idx = R.split_keys(keys)
    assert int(idx.min()) >= 0 and int(idx.max()) < N and int(dist.min()) >= 0 and int(dist.max()) <= K
    W = (K + 31) // 32
    sub = np.arange(0, Q, Q // 16)[:16]
    od, oi = c_oracle.topk(_u32(qp)[sub][:, :W], _u32(gp)[:, :W], K, k)
    assert np.array_equal(dist.cpu().numpy()[sub], od) and np.array_equal(idx.cpu().numpy()[sub], oi)


@pytest.mark.parametrize("kind", ["clustered", "all_equal", "two_values"])
def test_topk_degenerate_distance_distributions(kind):
    """Trained hash heads give clustered codes (huge tie buckets at tiny distances); the extreme is a gallery of identical codes.
    The top-k must still be the first k of the stable (distance, index) order."""
    Q, N, K, k = 300, 200_000, 64, 1000
    if kind == "clustered":
        qB, rB = synth.clustered_codes(Q, K, 5), synth.clustered_codes(N, K, 6)
    elif kind == "all_equal":
        qB = synth.random_codes(Q, K, 7)
        rB = synth.random_codes(1, K, 8).expand(N, K).contiguous()
    else:
        qB = synth.random_codes(Q, K, 9)
        two = synth.random_codes(2, K, 10)
        rB = two[(torch.arange(N) % 7 == 0).long()].contiguous()
    qp, gp = R.pack_codes(qB.to(DEV)), R.pack_codes(rB.to(DEV))
    keys = R.topk(qp, gp, K, k)
    dist, idx = R.split_keys(keys)
    od, oi = c_oracle.topk(_u32(qp)[:, :2], _u32(gp)[:, :2], K, k)
    assert np.array_equal(dist.cpu().numpy(), od) and np.array_equal(idx.cpu().numpy(), oi)


def test_calc_map_k_non_binary_codes_take_the_dense_path():
    """Codes with a 0 (sign_() of an exact zero) or +-2 rows (DistributedSampler padding + all_reduce SUM, runners/base.py:180-190,
    263-264): the reference computes a value there, and so does the drop-in (dense fp32 path) — compared with the port."""
    from oracle import calc_utils_port as port

    Q, N, K, C = 40, 900, 32, 12
    qB, rB = synth.random_codes(Q, K, 51), synth.random_codes(N, K, 52)
    qL, rL = synth.random_labels(Q, C, 53), synth.random_labels(N, C, 54)
    rB[5] *= 2.0
    rB[17, 3] = 0.0
    qB[2, 0] = 0.0
    want = port.calc_map_k(qB, rB, qL, rL, 50, stable=True)
    got = cu.calc_map_k(qB, rB, qL, rL, 50)
    assert got.dtype == torch.float32 and got.device.type == "cpu" and got.dim() == 0
    assert abs(float(got) - float(want)) <= 4e-7 * max(1.0, abs(float(want)))
    # labels given as counts (non 0/1): the reference only tests gram > 0
    qL2 = qL * 3
    assert abs(float(cu.calc_map_k(qB.sign() + (qB == 0), rB.sign() + (rB == 0), qL2, rL, 50)) -
               float(port.calc_map_k(qB.sign() + (qB == 0), rB.sign() + (rB == 0), qL2, rL, 50))) <= 4e-7


def test_valid_packed_matches_four_calc_map_k_calls():
    Q, N, K, C = 64, 3000, 64, 20
    codes = [synth.random_codes(n, K, 60 + i) for i, n in enumerate((Q, Q, N, N))]
    qL, rL = synth.random_labels(Q, C, 70), synth.random_labels(N, C, 71)
    packed = [R.pack_codes(c.to(DEV)) for c in codes]
    got = cu.valid_packed(packed[0], packed[1], packed[2], packed[3], qL, rL, K, None)
    want = (cu.calc_map_k(codes[0], codes[3], qL, rL), cu.calc_map_k(codes[1], codes[2], qL, rL),
            cu.calc_map_k(codes[0], codes[2], qL, rL), cu.calc_map_k(codes[1], codes[3], qL, rL))
    for g, w in zip(got, want):
        assert g.dtype == torch.float32 and g.device.type == "cpu" and float(g) == float(w)
