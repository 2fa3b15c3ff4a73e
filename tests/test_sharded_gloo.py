"""CPU, world_size 2 and 3 over gloo: the host-side exchange logic of ShardedEvaluator (shard geometry,
all-gather layouts, AP-partial reduction, top-k merge) with the numpy stage stand-in, against the golden
vectors of the reference.  The CUDA stages are checked on the GPU by tests/test_gpu_retrieval.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clip_based_cross_modal_hash_b200 import retrieval as R
        from oracle import hamming_oracle as ho
        from tests._golden import Case
        from tests._oracle_stages import OracleStages

        c = Case(name)
        W = (c.K + 31) // 32
        W = 4 if W == 3 else W
        LW = (c.C + 31) // 32
        LW = 4 if LW == 3 else LW

        def pad(a, w):
            out = np.zeros((a.shape[0], w), dtype=np.uint32)
            out[:, : min(w, a.shape[1])] = a[:, :w]
            return torch.from_numpy(out.view(np.int32))

        qp, gp = pad(ho.pack_codes(c.qB.numpy()), W), pad(ho.pack_codes(c.rB.numpy()), W)
        qlp, glp = pad(ho.pack_labels(c.qL.numpy()), LW), pad(ho.pack_labels(c.rL.numpy()), LW)
        lo, hi = R.shard_bounds(c.N, world)[rank]
        ev = R.ShardedEvaluator(stages=OracleStages())
        cap = max(int(c.totals.max()), 1)
        res = ev.map_k(qp, qlp, gp[lo:hi], glp[lo:hi], c.K, c.C, c.k, tindex_cap=cap)
        keys = ev.topk(qp, gp[lo:hi], c.K, 50, lo)                                   # rank_scatter: one all-reduce(MAX)
        keys_ag = ev.topk(qp, gp[lo:hi], c.K, 50, lo, method="allgather_merge")      # BASELINE's all-gather + merge
        assert torch.equal(keys, keys_ag)
        torch.save({"map": res.map, "total": res.total, "tsum": res.tsum, "tindex": res.tindex, "keys": keys},
                   os.path.join(out_dir, "r%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "tiny16"), (3, "odd48"), (2, "mid64_full")])
def test_sharded_evaluator_over_gloo(tmp_path, world, name):
    from tests._golden import Case

    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    c = Case(name)
    outs = [torch.load(os.path.join(str(tmp_path), "r%d.pt" % r)) for r in range(world)]
    for o in outs:
        assert abs(o["map"].item() - float(c.map_stable)) <= 4e-7
        assert np.array_equal(o["total"].numpy(), c.totals) and np.array_equal(o["tsum"].numpy(), c.tsums)
        tix = o["tindex"].numpy()
        for q in range(c.Q):
            assert np.array_equal(tix[q, : c.totals[q]], c.tindex[q])
        idx = (o["keys"].numpy().view(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.int64)
        assert np.array_equal(idx, c.order_head[:, :50])
        assert torch.equal(o["keys"], outs[0]["keys"]) and o["map"] == outs[0]["map"]  # every rank agrees


def _candidate_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clip_based_cross_modal_hash_b200 import retrieval as R
        from oracle import hamming_oracle as ho
        from tests._oracle_stages import OracleStages

        Q, K, k = 6, 32, 300
        N = 70_000 * world + 13                                   # shards above the candidate path's minimum size, ragged tail
        g = torch.Generator().manual_seed(7)
        qB = torch.randint(0, 2, (Q, K), generator=g).float() * 2 - 1
        rB = torch.randint(0, 2, (N, K), generator=g).float() * 2 - 1
        qp = torch.from_numpy(ho.pack_codes(qB.numpy()).view(np.int32))
        gp = torch.from_numpy(ho.pack_codes(rB.numpy()).view(np.int32))
        bounds = R.shard_bounds(N, world)
        lo, hi = bounds[rank]
        st = OracleStages(tensor_cores=True)
        calls = {"collect": 0, "rank_topk": 0}
        collect, rank_topk = st.topk_collect, st.rank_topk
        st.topk_collect = lambda *a, **kw: (calls.__setitem__("collect", calls["collect"] + 1), collect(*a, **kw))[1]
        st.rank_topk = lambda *a, **kw: (calls.__setitem__("rank_topk", calls["rank_topk"] + 1), rank_topk(*a, **kw))[1]
        ev = R.ShardedEvaluator(stages=st)
        keys = ev.topk(qp, gp[lo:hi], K, k, lo)                  # candidate path: gathers, global cutoff, verification, MAX exchange
        took_candidates = calls == {"collect": 1, "rank_topk": 0}
        exact = ev.topk(qp, gp[lo:hi], K, k, lo, exact=True)
        # a gallery whose sample prefixes say nothing about the rest: queries 0/1 have 5 000 exact copies at the END of every shard
        rB2 = rB.clone()
        for a, b in bounds:
            rB2[b - 5000:b] = qB[0]
        gp2 = torch.from_numpy(ho.pack_codes(rB2.numpy()).view(np.int32))
        calls.update(collect=0, rank_topk=0)
        keys2 = ev.topk(qp, gp2[lo:hi], K, k, lo)
        fell_back = calls["collect"] == 1 and calls["rank_topk"] == 1
        torch.save({"keys": keys, "exact": exact, "keys2": keys2, "took_candidates": took_candidates, "fell_back": fell_back},
                   os.path.join(out_dir, "c%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def _stable_topk_keys(qB, rB, k):
    d = ((qB.shape[1] - qB.double() @ rB.double().t()) / 2).round().to(torch.int64)
    idx = torch.argsort(d, dim=1, stable=True)[:, :k]
    return (torch.gather(d, 1, idx) << 32) | idx


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_candidate_topk_over_gloo(tmp_path, world):
    """the single-pass candidate path of the sharded top-k, host logic over gloo with the numpy stage stand-in: sample blocks
    gathered -> one global cutoff -> collect / count -> totals gathered -> verification -> placement -> MAX exchange; and a gallery
    with unrepresentative prefixes is caught by the verification on every rank and re-done by the exact path."""
    mp.spawn(_candidate_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    Q, K, k = 6, 32, 300
    N = 70_000 * world + 13
    g = torch.Generator().manual_seed(7)
    qB = torch.randint(0, 2, (Q, K), generator=g).float() * 2 - 1
    rB = torch.randint(0, 2, (N, K), generator=g).float() * 2 - 1
    want = _stable_topk_keys(qB, rB, k)
    from clip_based_cross_modal_hash_b200 import retrieval as R

    rB2 = rB.clone()
    for a, b in R.shard_bounds(N, world):
        rB2[b - 5000:b] = qB[0]
    want2 = _stable_topk_keys(qB, rB2, k)
    for r in range(world):
        o = torch.load(os.path.join(str(tmp_path), "c%d.pt" % r))
        assert o["took_candidates"] and o["fell_back"]
        assert torch.equal(o["keys"], want) and torch.equal(o["exact"], want)
        assert torch.equal(o["keys2"], want2)


def _merge_worker(rank, world, port, out):
    import torch.distributed as dist

    from clip_based_cross_modal_hash_b200 import models

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        full = torch.randint(-2 ** 31, 2 ** 31 - 1, (11, 2), generator=g, dtype=torch.int64).to(torch.int32)
        # DistributedSampler pads 11 rows to 12 by repeating index 0: rank r owns indices r, r+world, ... of the padded list
        padded = list(range(11)) + [0]
        mine = padded[rank::world]
        buf = torch.zeros_like(full)
        buf[mine] = full[mine]
        models.merge_code_buffers(buf)
        out.put((rank, bool(torch.equal(buf, full))))
    finally:
        dist.destroy_process_group()


def test_merge_code_buffers_survives_sampler_padding():
    """get_code's distributed merge: packed buffers, bitwise OR, duplicated (padded) rows stay exact."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    world, port = 2, 29653
    procs = [ctx.Process(target=_merge_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
