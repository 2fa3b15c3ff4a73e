"""CPU: pin every oracle function against the vectors the reference produced (tests/golden)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, calc_utils_port as port, hamming_oracle as ho
from tests._golden import CASE_NAMES, Case, npz


@pytest.mark.parametrize("name", CASE_NAMES)
def test_port_map_matches_reference(name):
    c = Case(name)
    m, tindex, totals = port.calc_map_k(c.qB, c.rB, c.qL, c.rL, c.k, stable=True, return_parts=True)
    # same torch CPU ops as the reference on the same machine class: bit-exact fp32
    assert np.float32(m.item()) == c.map_stable, (m.item(), c.map_stable)
    assert totals == list(c.totals)
    for a, b in zip(tindex, c.tindex):
        assert np.array_equal(a.numpy(), b)
    # chunking the query axis must not change anything
    m2 = port.calc_map_k(c.qB, c.rB, c.qL, c.rL, c.k, stable=True, query_chunk=7)
    assert np.float32(m2.item()) == c.map_stable


@pytest.mark.parametrize("name", ["tiny16", "mid64", "odd32", "sparse_rel"])
def test_port_unstable_matches_shipped(name):
    c = Case(name)
    m = port.calc_map_k(c.qB, c.rB, c.qL, c.rL, c.k, stable=False)
    assert np.float32(m.item()) == c.map_shipped
    # the canonicalisation moves mAP only slightly (documented in DESIGN.md)
    assert abs(float(c.map_shipped) - float(c.map_stable)) < 5e-3


@pytest.mark.parametrize("name", CASE_NAMES)
def test_integer_oracle_matches_reference(name):
    c = Case(name)
    qp, gp = ho.pack_codes(c.qB.numpy()), ho.pack_codes(c.rB.numpy())
    qlp, glp = ho.pack_labels(c.qL.numpy()), ho.pack_labels(c.rL.numpy())
    if c.Q * c.N <= 400_000:
        hm = ho.hamming_matrix(qp, gp)
        assert np.array_equal(hm.sum(axis=1, dtype=np.int64), c.hamm_rowsum)
        if c.hamm is not None:
            assert np.array_equal(hm.astype(np.uint8), c.hamm)
        tindex, totals, tsums = ho.map_parts(qp, gp, qlp, glp, c.K, c.k)
        assert np.array_equal(totals, c.totals) and np.array_equal(tsums, c.tsums)
        for a, b in zip(tindex, c.tindex):
            assert np.array_equal(a, b)
        kk = min(64, c.N)
        d, i = ho.topk(qp, gp, c.K, kk)
        assert np.array_equal(i[:, :kk], c.order_head[:, :kk])
        if c.hamm is not None:
            assert np.array_equal(d, np.take_along_axis(c.hamm.astype(np.int32), i, axis=1))


@pytest.mark.parametrize("name", CASE_NAMES)
def test_c_oracle_matches_reference(name):
    c = Case(name)
    qp, bad = c_oracle.pack_codes(c.qB.numpy())
    assert bad == 0 and np.array_equal(qp, ho.pack_codes(c.qB.numpy()))
    gp, _ = c_oracle.pack_codes(c.rB.numpy())
    qlp, bad = c_oracle.pack_labels(c.qL.numpy())
    assert bad == 0 and np.array_equal(qlp, ho.pack_labels(c.qL.numpy()))
    glp, _ = c_oracle.pack_labels(c.rL.numpy())
    tindex, totals, tsums = c_oracle.map_tindex(qp, qlp, gp, glp, c.K, c.k)
    assert np.array_equal(totals, c.totals) and np.array_equal(tsums, c.tsums)
    for q in range(c.Q):
        assert np.array_equal(tindex[q, : totals[q]], c.tindex[q])
    kk = min(64, c.N)
    d, i = c_oracle.topk(qp, gp, c.K, kk)
    assert np.array_equal(i, c.order_head[:, :kk])
    assert np.array_equal(c_oracle.hamming_u16(qp, gp).sum(axis=1, dtype=np.int64), c.hamm_rowsum)


def test_map_float64_close_to_reference():
    for name in CASE_NAMES:
        c = Case(name)
        v = ho.map_float64([t.astype(np.int64) for t in c.tindex])
        assert abs(v - float(c.map_stable)) <= 4e-7 * max(1.0, abs(v)), (name, v, c.map_stable)


def test_shard_merge_arithmetic():
    c = Case("mid64")
    qp, gp = ho.pack_codes(c.qB.numpy()), ho.pack_codes(c.rB.numpy())
    for world in (2, 3, 8):
        bounds = ho.shard_bounds(c.N, world, align=4)
        assert bounds[0][0] == 0 and bounds[-1][1] == c.N
        for q in range(3):
            full = ho.hamming_matrix(qp[q:q + 1], gp)[0]
            want = ho.stable_ranks(full, c.K + 1)
            got = np.concatenate(ho.merged_ranks_from_shards([full[lo:hi] for lo, hi in bounds], c.K + 1))
            assert np.array_equal(got, want)


def test_similarity_helpers_match_reference():
    z = npz()
    a, b = torch.from_numpy(z["sim/a"]), torch.from_numpy(z["sim/b"])
    la, lb = torch.from_numpy(z["sim/la"]), torch.from_numpy(z["sim/lb"])
    assert np.array_equal(port.calc_label_sim(la, lb).numpy(), z["sim/label_sim"])
    assert np.array_equal(port.calc_label_sim(la.long(), lb.long()).numpy(), z["sim/label_sim_i64"])
    assert np.array_equal(port.cosine_similarity(a, b).numpy(), z["sim/cosine"])
    assert np.allclose(port.cosine_similarity(a.numpy(), b.numpy()), z["sim/cosine_np"], rtol=0, atol=0)
    assert np.array_equal(port.euclidean_similarity(a, b).numpy(), z["sim/euclid"])
    assert np.allclose(port.euclidean_similarity(a.numpy(), b.numpy()), z["sim/euclid_np"], rtol=1e-6)
    ls, ws = port.generate_weight_sim(la, la)
    assert np.array_equal(ls.numpy(), z["sim/weight_label"])
    assert np.allclose(ws.numpy(), z["sim/weight_sim"], rtol=1e-6, atol=0)
    with pytest.raises(ValueError):
        port.cosine_similarity(a, b.numpy())
    with pytest.raises(ValueError):
        port.euclidean_similarity(a.numpy(), b)
    q, r = torch.from_numpy(z["hd/q"]), torch.from_numpy(z["hd/r"])
    assert np.array_equal(port.calc_hammingDist(q, r).numpy(), z["hd/full"])
    assert np.array_equal(port.calc_hammingDist(q[1], r).numpy(), z["hd/one_d"])
