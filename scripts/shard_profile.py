"""One GPU plays rank R of an 8-way sharded C4-64 top-k (gathers simulated with the blocks of all shards computed beforehand), so
that `ncu --metrics gpu__time_duration.sum --profile-from-start off` lists the per-kernel GPU time of ONE shard's step without
launch gaps or waits for peers.  Usage: ncu ... python scripts/shard_profile.py [world] [rank]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_based_cross_modal_hash_b200 import retrieval as R, synth  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
me = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
Q, N, K, k = 10000, 1_000_000, 64, 1000
q = synth.random_codes(Q, K, 1).to(dev)
g = synth.random_codes(N, K, 2).to(dev)
st = R.CudaStages(True)
bounds = R.shard_bounds(N, world)
n_geom = max(hi - lo for lo, hi in bounds)
qp = R.pack_codes(q)
shards, blocks = [], []
for r, (lo, hi) in enumerate(bounds):
    gp = R.pack_codes(g[lo:hi].contiguous())
    plan = st.make_plan(Q, hi - lo, K, 0, n_geom)
    ops = st.operands(plan, qp, None, gp, None)
    shards.append((plan, lo, hi, gp, ops))

    def capture(t):
        blocks.append(t.clone())
        return torch.stack([t] * world).contiguous()
    R.collect_candidates(st, plan, ops, qp, gp, k, gather=capture, idx_offset=lo, rank=r, world=world)
gathered = torch.stack(blocks).contiguous()
tots = []
for r, (plan, lo, hi, gp, ops) in enumerate(shards):
    tots.append(R.collect_candidates(st, plan, ops, qp, gp, k, gather=lambda t: gathered, idx_offset=lo, rank=r, world=world)[3])
tot_all = torch.stack(tots).contiguous()
plan, lo, hi, _, _ = shards[me]
g_me = g[lo:hi].contiguous()
keys = torch.empty((Q, k), dtype=torch.int64, device=dev)


def step():
    qp = R.pack_codes(q)
    gp = R.pack_codes(g_me)
    ops = st.operands(plan, qp, None, gp, None)
    cap, cand, cnt, tot, meta = R.collect_candidates(st, plan, ops, qp, gp, k, gather=lambda t: gathered, idx_offset=lo, rank=me, world=world)
    need = torch.clamp(meta[:, plan.bins, 1].sum(), max=k)
    short = (tot_all[:, : plan.bins, :Q].sum(dim=(0, 1)) < need).any()
    bad = short | (tot_all[:, plan.bins, 0].max() != 0)
    keys.fill_(R.EMPTY_KEY)
    st.topk_place(plan, cap, cand, cnt, tot_all, world, me, k, lo, keys)
    return bad


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("placed keys of rank %d: %d of %d slots" % (me, int((keys != R.EMPTY_KEY).sum()), keys.numel()))
