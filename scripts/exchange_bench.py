"""Times the pieces of the fused top-k exchange on N GPUs (torchrun): symmetric-memory barriers, the place kernel with local /
per-peer / multicast stores, the NCCL all-reduce it replaces."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_based_cross_modal_hash_b200 import retrieval as R, synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ev = R.ShardedEvaluator()
st = ev.stages
Q, N, K, k = 10000, 1_000_000, 64, 1000
lo, hi = R.shard_bounds(N, world)[rank]
n_geom = max(b - a for a, b in R.shard_bounds(N, world))
qp = R.pack_codes(synth.random_codes(Q, K, 1).to(dev))
gp = R.pack_codes(synth.random_codes(N, K, 2)[lo:hi].contiguous().to(dev))
plan = st.make_plan(Q, hi - lo, K, 0, n_geom)
ops = st.operands(plan, qp, None, gp, None)
cap, cand, cnt, tot, _ = R.collect_candidates(st, plan, ops, qp, gp, k)
tot_all = ev._gather(tot)
buf, hdl = ev._symmetric_keys(Q, k, dev)
keys = torch.full((Q, k), -1, dtype=torch.int64, device=dev)


def timed(name, fn, n=20):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print("%-44s %8.3f ms" % (name, e0.elapsed_time(e1) / n), flush=True)


timed("2 x symmetric-memory barrier", lambda: (hdl.barrier(channel=0), hdl.barrier(channel=1)))
timed("place -> local buffer", lambda: st.topk_place(plan, cap, cand, cnt, tot_all, world, rank, k, lo, keys))
timed("place -> per-peer stores", lambda: st.topk_place(plan, cap, cand, cnt, tot_all, world, rank, k, lo, None,
                                                        peers_dev=int(hdl.buffer_ptrs_dev), npeers=world))
if getattr(hdl, "has_multicast_support", False):
    timed("place -> multimem.st", lambda: st.topk_place(plan, cap, cand, cnt, tot_all, world, rank, k, lo, None,
                                                        peers_dev=int(hdl.buffer_ptrs_dev), npeers=world, multicast=int(hdl.multicast_ptr)))
from clip_based_cross_modal_hash_b200 import _lib  # noqa: E402


def nvls():
    hdl.barrier(channel=1)
    _lib.check(_lib.lib().cmh_nvls_allreduce_max_s64(int(hdl.multicast_ptr), buf.numel(), rank, world, torch.cuda.current_stream().cuda_stream))
    hdl.barrier(channel=2)


def push():
    hdl.barrier(channel=0)
    _lib.check(_lib.lib().cmh_nvls_push_owned_s64(buf.data_ptr(), int(hdl.multicast_ptr), buf.numel(), torch.cuda.current_stream().cuda_stream))
    hdl.barrier(channel=1)


if getattr(hdl, "has_multicast_support", False):
    timed("barrier + NVLS all-reduce(max) kernel + barrier", nvls)
    timed("barrier + NVLS push-owned kernel + barrier", push)
    timed("NVLS gather of the totals (2 barriers + bcast)", lambda: ev._gather_small("totals", tot, True))
    timed("fill(-1) of the [Q, k] buffer", lambda: buf.fill_(-1))
timed("NCCL all-reduce(MAX) of the [Q, k] keys", lambda: dist.all_reduce(keys, op=dist.ReduceOp.MAX))
timed("NCCL all-gather of the totals", lambda: ev._gather(tot))
timed("clone of the [Q, k] buffer", lambda: buf.clone())
for m in ("nvls", "nvls_reduce", "rank_scatter"):
    for small in ("nvls", "nccl"):
        os.environ["CMH_SMALL_EXCHANGE"] = small
        timed("whole sharded top-k: keys %s, small exchanges %s" % (m, small),
              lambda: ev.topk(qp, gp, K, k, lo, n_geom, method=m, copy=False), n=10)
os.environ.pop("CMH_SMALL_EXCHANGE", None)
dist.destroy_process_group()
