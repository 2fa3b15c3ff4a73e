"""Time the 8-way partial top-k merge (C4 shapes) on one GPU and check it against a sort: python scripts/merge_bench.py"""
import sys, torch
sys.path.insert(0, '/root/repo')
from clip_based_cross_modal_hash_b200 import retrieval as R
st = R.CudaStages()
for world, Q, k in ((8, 10000, 1000), (2, 10000, 1000), (3, 777, 50), (8, 2100, 5000)):
    g = torch.Generator(device='cuda').manual_seed(1)
    dist = torch.randint(0, 65, (world, Q, k), device='cuda', generator=g, dtype=torch.int64)
    idx = torch.randperm(world * Q * k, device='cuda', generator=g).view(world, Q, k) % (1 << 31)
    keys = ((dist << 32) | idx).sort(dim=-1).values.contiguous()
    keys[0, :, k - 3:] = -1   # a few empty slots (all ones) at the end of rank 0's lists
    want = keys.permute(1, 0, 2).reshape(Q, world * k).view(torch.int64)
    want = torch.sort(want.to(torch.float64) if False else (want ^ (1 << 63)), dim=-1).values[:, :k] ^ (1 << 63)  # unsigned order
    for _ in range(2):
        got = st.topk_merge(keys)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        got = st.topk_merge(keys)
    e1.record(); torch.cuda.synchronize()
    print('world=%d Q=%d k=%d: %.3f ms  exact=%s' % (world, Q, k, e0.elapsed_time(e1) / 5, bool(torch.equal(got, want))))
