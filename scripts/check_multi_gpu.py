#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun on N GPUs of one box):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
         scripts/check_multi_gpu.py

Every rank holds one contiguous gallery shard; the sharded mAP / tindex / top-k (NCCL all-gathers) must equal
the single-GPU result computed on rank 0 from the whole gallery, and the C oracle on a query subset.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from clip_based_cross_modal_hash_b200 import retrieval as R, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ev = R.ShardedEvaluator()
    ok = True
    for (Q, N, K, C, k) in [(300, 20011, 64, 80, None), (257, 50000, 128, 21, 500), (1000, 200000, 32, 24, 1000),
                            (700, 150000 * world, 64, 80, 1000)]:   # last: every shard large enough for the candidate path
        qB, rB = synth.random_codes(Q, K, 1), synth.random_codes(N, K, 2)
        qL, rL = synth.random_labels(Q, C, 3), synth.random_labels(N, C, 4)
        qp, gp = R.pack_codes(qB.to(dev)), R.pack_codes(rB.to(dev))
        qlp, glp = R.pack_labels(qL.to(dev)), R.pack_labels(rL.to(dev))
        lo, hi = R.shard_bounds(N, world)[rank]
        single = R.map_k(qp, qlp, gp, glp, K, C, k, want_tindex=True, tindex_cap=2048)
        res = ev.map_k(qp, qlp, gp[lo:hi], glp[lo:hi], K, C, k, tindex_cap=2048)
        kk = k or 1000
        keys1 = R.topk(qp, gp, K, kk)
        keys = ev.topk(qp, gp[lo:hi], K, kk, lo)
        keys_ag = ev.topk(qp, gp[lo:hi], K, kk, lo, method="allgather_merge")
        keys_rs = ev.topk(qp, gp[lo:hi], K, kk, lo, method="rank_scatter")
        keys_ex = ev.topk(qp, gp[lo:hi], K, kk, lo, exact=True)
        good = (torch.equal(res.tindex, single.tindex) and torch.equal(res.total, single.total)
                and abs(res.map.item() - single.map.item()) < 1e-12 and torch.equal(keys, keys1) and torch.equal(keys_ag, keys1)
                and torch.equal(keys_rs, keys1) and torch.equal(keys_ex, keys1))
        if rank == 0:
            from oracle import c_oracle, hamming_oracle as ho
            W = (K + 31) // 32
            sub = slice(0, 16)
            tix, totals, _ = c_oracle.map_tindex(qp.cpu().numpy().view(np.uint32)[sub, :W], ho.pack_labels(qL.numpy()[sub]),
                                                 gp.cpu().numpy().view(np.uint32)[:, :W], ho.pack_labels(rL.numpy()), K, k, cap=2048)
            good = good and np.array_equal(res.tindex.cpu().numpy()[sub], tix)
            print("Q=%d N=%d K=%d k=%s world=%d: %s  mAP=%.9f" % (Q, N, K, k, world, "OK" if good else "MISMATCH", res.map.item()), flush=True)
        ok = ok and good
    # the fused exchange of the candidate path: multicast stores (when the box has NVLS) and plain peer stores
    Q, N, K, kk = 900, 120000 * world, 64, 1000
    qp, gp = R.pack_codes(synth.random_codes(Q, K, 5).to(dev)), R.pack_codes(synth.random_codes(N, K, 6).to(dev))
    lo, hi = R.shard_bounds(N, world)[rank]
    want = R.topk(qp, gp, K, kk, exact=True)
    outs = {}
    outs["auto"] = bool(torch.equal(ev.topk(qp, gp[lo:hi], K, kk, lo), want))
    outs["auto, zero-copy result"] = bool(torch.equal(ev.topk(qp, gp[lo:hi], K, kk, lo, copy=False), want))
    if ev.exchange_info().get("peer_memory"):
        outs["peer_stores"] = bool(torch.equal(ev.topk(qp, gp[lo:hi], K, kk, lo, method="peer_stores"), want))
        outs["nvls_reduce"] = bool(torch.equal(ev.topk(qp, gp[lo:hi], K, kk, lo, method="nvls_reduce"), want))
    outs["rank_scatter"] = bool(torch.equal(ev.topk(qp, gp[lo:hi], K, kk, lo, method="rank_scatter"), want))
    # the whole sharded step as one CUDA graph per rank, replayed on refreshed static inputs
    n_geom = max(b - a for a, b in R.shard_bounds(N, world))
    q_codes = synth.random_codes(Q, K, 5).to(dev)
    g_codes = synth.random_codes(N, K, 6)[lo:hi].contiguous().to(dev)
    graph = R.TopkGraph(q_codes, g_codes, kk, evaluator=ev, idx_offset=lo, n_geom=n_geom)
    outs["graph"] = bool(torch.equal(graph.run(), want))
    g2 = synth.random_codes(N, K, 66)
    g_codes.copy_(g2[lo:hi].to(dev))
    want2 = R.topk(qp, R.pack_codes(g2.to(dev)), K, kk, exact=True)
    outs["graph, replay on a new gallery"] = bool(torch.equal(graph.run(), want2))
    outs["eager after graph"] = bool(torch.equal(ev.topk(qp, gp[lo:hi], K, kk, lo), want))
    if rank == 0:
        print("fused exchange world=%d: %s  %s" % (world, outs, ev.exchange_info()), flush=True)
    ok = ok and all(outs.values())
    # get_code's distributed merge of packed code buffers (byte-wise MAX over NCCL), with DistributedSampler-style padding
    from clip_based_cross_modal_hash_b200 import models
    g = torch.Generator().manual_seed(0)
    n = 8 * world + 3
    full = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, 2), generator=g, dtype=torch.int64).to(torch.int32).to(dev)
    padded = list(range(n)) + list(range((-n) % world))
    mine = torch.tensor(padded[rank::world], device=dev)
    buf = torch.zeros_like(full)
    buf[mine] = full[mine]
    models.merge_code_buffers(buf)
    good = bool(torch.equal(buf, full))
    if rank == 0:
        print("merge_code_buffers world=%d: %s" % (world, "OK" if good else "MISMATCH"), flush=True)
    ok = ok and good
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    dist.destroy_process_group()
    sys.exit(int(t.item() != 0))


if __name__ == "__main__":
    main()
