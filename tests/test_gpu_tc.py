"""GPU: the tensor-core ranking passes (csrc/cmh_tc.cu, tcgen05.mma.kind::i8) against the XOR+POPC kernels of
csrc/cmh_retrieval.cu and the numpy oracle — bit-exact integer stages for every operand width (32/64/128-byte swizzle rows),
with and without the label block, on ragged shapes."""
import numpy as np
import pytest
import torch

from clip_based_cross_modal_hash_b200 import retrieval as R
from clip_based_cross_modal_hash_b200 import synth
from oracle import hamming_oracle as ho

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _u32(t):
    return t.cpu().numpy().view(np.uint32)


def _inputs(Q, N, K, C, seed):
    qp = R.pack_codes(synth.random_codes(Q, K, seed).to(DEV))
    gp = R.pack_codes(synth.random_codes(N, K, seed + 1).to(DEV))
    if C == 0:
        return qp, gp, None, None
    qlp = R.pack_labels(synth.random_labels(Q, C, seed + 2, p=0.15).to(DEV))
    glp = R.pack_labels(synth.random_labels(N, C, seed + 3, p=0.15).to(DEV))
    return qp, gp, qlp, glp


@pytest.mark.parametrize("K", [8, 16, 32, 33, 64, 100, 128])
def test_expand_matches_numpy(K):
    n, rows = 37, 128
    codes = synth.random_codes(n, K, 3)
    packed = R.pack_codes(codes.to(DEV))
    out = R._expand(packed, rows, K, 0).cpu().numpy()
    KP = out.shape[1]
    assert KP == (32 if K <= 32 else 64 if K <= 64 else 128)
    want = np.zeros((rows, KP), dtype=np.int8)
    want[:n, :K] = codes.numpy().astype(np.int8)
    want[n:, :K] = -1
    assert np.array_equal(out, want)
    for kind, val in ((1, -128), (2, 8)):
        lab = synth.random_labels(n, K, 5)
        lp = R.pack_labels(lab.to(DEV))
        got = R._expand(lp, rows, K, kind).cpu().numpy()
        w = np.zeros((rows, KP), dtype=np.int8)
        w[:n, :K] = (lab.numpy() != 0) * val
        assert np.array_equal(got, w)


SHAPES = [(5, 70), (130, 1000), (128, 64), (257, 4097), (300, 70_001)]


@pytest.mark.parametrize("K", [16, 32, 48, 64, 96, 128])
@pytest.mark.parametrize("C", [0, 24, 40, 80])
@pytest.mark.parametrize("Q,N", SHAPES)
def test_tc_stages_equal_popc_stages(Q, N, K, C):
    qp, gp, qlp, glp = _inputs(Q, N, K, C, 11 * K + C + Q)
    tc, pc = R.CudaStages(True), R.CudaStages(False)
    plan = tc.make_plan(Q, N, K, C, None, 64 if N < 5000 else 0)
    ops = tc.operands(plan, qp, qlp, gp, glp)
    h_tc = tc.hist(plan, qp, qlp, gp, glp, ops=ops)
    h_pc = pc.hist(plan, qp, qlp, gp, glp)
    assert torch.equal(h_tc[:, :, :Q], h_pc[:, :, :Q])
    k = min(50, N)
    sc = pc.scan(plan, h_pc, 1, 0, k, with_rel=C > 0)
    keys_tc = tc.rank_topk(plan, qp, gp, sc, k, 7, ops=ops)
    keys_pc = pc.rank_topk(plan, qp, gp, sc, k, 7)
    assert torch.equal(keys_tc, keys_pc)
    if C > 0:
        sc = pc.scan(plan, h_pc, 1, 0, None)
        cap = 64
        t_tc = torch.zeros((Q, cap), dtype=torch.int32, device=DEV)
        t_pc = torch.zeros((Q, cap), dtype=torch.int32, device=DEV)
        ap_tc = tc.rank_map(plan, qp, qlp, gp, glp, sc, t_tc, ops=ops)
        ap_pc = pc.rank_map(plan, qp, qlp, gp, glp, sc, t_pc)
        assert torch.equal(t_tc, t_pc)
        assert torch.equal(ap_tc[:, :Q], ap_pc[:, :Q])          # same fp32 quotients, same fp64 summation order
        ap2 = tc.rank_map(plan, qp, qlp, gp, glp, sc, None, ops=ops)
        assert torch.equal(ap2[:, :Q], ap_pc[:, :Q])


@pytest.mark.parametrize("K,C", [(16, 24), (64, 80), (128, 21)])
def test_tc_histogram_matches_numpy_oracle(K, C):
    Q, N = 70, 3000
    qp, gp, qlp, glp = _inputs(Q, N, K, C, 5)
    st = R.CudaStages(True)
    plan = st.make_plan(Q, N, K, C, None, 8)
    ops = st.operands(plan, qp, qlp, gp, glp)
    h = _u32(st.hist(plan, qp, qlp, gp, glp, ops=ops))
    W = (K + 31) // 32
    d = ho.hamming_matrix(_u32(qp)[:, :W], _u32(gp)[:, :W]).astype(np.int64)
    rel = ((_u32(qlp)[:, None, :] & _u32(glp)[None, :, :]) != 0).any(axis=2)
    for c in range(plan.nchunks):
        lo, hi = c * plan.chunk_items, min((c + 1) * plan.chunk_items, N)
        for q in range(0, Q, 7):
            ha = np.bincount(d[q, lo:hi], minlength=plan.bins)
            hr = np.bincount(d[q, lo:hi][rel[q, lo:hi]], minlength=plan.bins)
            assert np.array_equal(h[c, :, q] & 0xFFFF, ha) and np.array_equal(h[c, :, q] >> 16, hr)


def test_tc_many_shared_classes_do_not_leak_into_the_distance():
    """every query shares ALL classes with every gallery item: the label term of the accumulator is at its maximum."""
    Q, N, K, C = 64, 2000, 128, 128
    qp = R.pack_codes(synth.random_codes(Q, K, 1).to(DEV))
    gp = R.pack_codes(synth.random_codes(N, K, 2).to(DEV))
    qlp = R.pack_labels(torch.ones(Q, C, dtype=torch.int64, device=DEV))
    glp = R.pack_labels(torch.ones(N, C, dtype=torch.int64, device=DEV))
    tc, pc = R.CudaStages(True), R.CudaStages(False)
    plan = tc.make_plan(Q, N, K, C, None, 4)
    ops = tc.operands(plan, qp, qlp, gp, glp)
    assert torch.equal(tc.hist(plan, qp, qlp, gp, glp, ops=ops)[:, :, :Q], pc.hist(plan, qp, qlp, gp, glp)[:, :, :Q])


def test_tc_is_the_default_path():
    assert R.CudaStages().tensor_cores is True


# ---- candidate path of the top-k ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K", [16, 32, 64, 128])
@pytest.mark.parametrize("k", [1, 100, 1000])
def test_candidate_topk_equals_exact_two_pass(K, k):
    Q, N = 700, 300_000
    qp = R.pack_codes(synth.random_codes(Q, K, 100 + K).to(DEV))
    gp = R.pack_codes(synth.random_codes(N, K, 101 + K).to(DEV))
    st = R.CudaStages(True)
    assert R.candidate_path_ok(st, st.make_plan(Q, N, K, 0), N, k)
    stages = []
    fast = R.topk(qp, gp, K, k, idx_offset=5, stages=stages)
    assert len(stages) == len(R.TOPK_FAST_STAGE_NAMES)        # i.i.d. codes: the candidate path verified and returned
    exact = R.topk(qp, gp, K, k, idx_offset=5, exact=True)
    assert torch.equal(fast, exact)


def test_candidate_topk_clustered_codes_and_fallback():
    """clustered codes: huge tie buckets at small distances (lists overflow -> verified fallback); result must not change."""
    Q, N, K, k = 256, 262_144, 64, 500
    qp = R.pack_codes(synth.clustered_codes(Q, K, 5).to(DEV))
    gp = R.pack_codes(synth.clustered_codes(N, K, 6).to(DEV))
    assert torch.equal(R.topk(qp, gp, K, k), R.topk(qp, gp, K, k, exact=True))
    # a gallery whose prefix looks nothing like the rest: the sampled cutoffs are far too tight -> detected, exact path
    far = synth.random_codes(N, K, 7)
    near = synth.random_codes(1, K, 8).expand(40_000, K)
    g2 = torch.cat([far[: N - 40_000], near]).contiguous()
    q2 = torch.cat([synth.random_codes(1, K, 8).expand(100, K), synth.random_codes(Q - 100, K, 9)]).contiguous()
    qp2, gp2 = R.pack_codes(q2.to(DEV)), R.pack_codes(g2.to(DEV))
    assert torch.equal(R.topk(qp2, gp2, K, k), R.topk(qp2, gp2, K, k, exact=True))


@pytest.mark.parametrize("world", [2, 3])
def test_candidate_stages_sharded_equal_single(world):
    """the sharded form of the candidate path, simulated on one GPU: per-shard collect + count, totals concatenated like the
    all-gather, per-shard place into one key buffer == single-GPU keys."""
    Q, N, K, k = 300, 400_000, 64, 1000
    qp = R.pack_codes(synth.random_codes(Q, K, 1).to(DEV))
    gp = R.pack_codes(synth.random_codes(N, K, 2).to(DEV))
    want = R.topk(qp, gp, K, k, exact=True)
    st = R.CudaStages(True)
    bounds = R.shard_bounds(N, world)
    n_geom = max(hi - lo for lo, hi in bounds)
    parts = []
    for lo, hi in bounds:
        plan = st.make_plan(Q, hi - lo, K, 0, n_geom)
        ops = st.operands(plan, qp, None, gp[lo:hi], None)
        parts.append((plan, lo) + R.collect_candidates(st, plan, ops, qp, gp[lo:hi], k)[:4])
    tot_all = torch.stack([p[5] for p in parts]).contiguous()
    assert int(tot_all[:, parts[0][0].bins, 0].max()) == 0
    keys = torch.full((Q, k), R.EMPTY_KEY, dtype=torch.int64, device=DEV)
    for r, (plan, lo, cap, cand, cnt, tot) in enumerate(parts):
        mine = torch.full((Q, k), R.EMPTY_KEY, dtype=torch.int64, device=DEV)
        st.topk_place(plan, cap, cand, cnt, tot_all, world, r, k, lo, mine)
        assert int(((mine != R.EMPTY_KEY) & (keys != R.EMPTY_KEY)).sum()) == 0   # every slot has exactly one owner
        keys = torch.maximum(keys, mine)
    assert torch.equal(keys, want)


@pytest.mark.parametrize("world,k", [(2, 1000), (4, 257), (3, 64)])
def test_candidate_stages_global_cutoff_equal_single(world, k):
    """the multi-GPU form with ONE global cutoff per query, simulated on one GPU: sample blocks of all shards gathered (stacked),
    every shard derives the same cutoff and keeps ~k/world candidates, totals stacked like the all-gather, per-shard place into one
    key buffer == single-GPU keys; the union of the candidates is enough for every query."""
    Q, N, K = 300, 400_000, 64
    qp = R.pack_codes(synth.random_codes(Q, K, 11).to(DEV))
    gp = R.pack_codes(synth.random_codes(N, K, 12).to(DEV))
    want = R.topk(qp, gp, K, k, exact=True)
    st = R.CudaStages(True)
    bounds = R.shard_bounds(N, world)
    n_geom = max(hi - lo for lo, hi in bounds)
    shards = []
    for lo, hi in bounds:
        plan = st.make_plan(Q, hi - lo, K, 0, n_geom)
        shards.append((plan, lo, hi, st.operands(plan, qp, None, gp[lo:hi], None)))
    blocks = []                                           # pass 1: every shard's own sample block (what it would send)
    for r, (plan, lo, hi, ops) in enumerate(shards):
        def capture(t):
            blocks.append(t.clone())
            return torch.stack([t] * world).contiguous()
        R.collect_candidates(st, plan, ops, qp, gp[lo:hi], k, gather=capture, idx_offset=lo, rank=r, world=world)
    gathered = torch.stack(blocks).contiguous()
    bins = shards[0][0].bins
    assert int(gathered[:, bins, 1].sum()) == N
    parts = []
    for r, (plan, lo, hi, ops) in enumerate(shards):       # pass 2: the real thing on the gathered blocks
        cap, cand, cnt, tot, meta = R.collect_candidates(st, plan, ops, qp, gp[lo:hi], k, gather=lambda t: gathered,
                                                         idx_offset=lo, rank=r, world=world)
        parts.append((plan, lo, cap, cand, cnt, tot))
    tot_all = torch.stack([p[5] for p in parts]).contiguous()
    assert int(tot_all[:, bins, 0].max()) == 0                                  # no list overflowed
    per_query = tot_all[:, :bins, :Q].sum(dim=(0, 1))
    assert int(per_query.min()) >= k                                            # together: enough for every query
    assert float(per_query.float().mean()) < 2.5 * k + 200                      # ... and about k of them, not world x k
    keys = torch.full((Q, k), R.EMPTY_KEY, dtype=torch.int64, device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    for r, (plan, lo, cap, cand, cnt, tot) in enumerate(parts):
        mine = torch.full((Q, k), R.EMPTY_KEY, dtype=torch.int64, device=DEV)
        st.topk_place(plan, cap, cand, cnt, tot_all, world, r, k, lo, mine, verify=(gathered, status))
        assert int(((mine != R.EMPTY_KEY) & (keys != R.EMPTY_KEY)).sum()) == 0
        keys = torch.maximum(keys, mine)
    assert torch.equal(keys, want)
    assert int(status.item()) == 0                                              # verified inside the place kernel
    plan, lo, cap, cand, cnt, tot = parts[0]
    mine = torch.empty((Q, k), dtype=torch.int64, device=DEV)
    short = tot_all.clone()
    short[:, :bins, 5] = 0                                                      # query 5: no candidates anywhere -> too few
    st.topk_place(plan, cap, cand, cnt, short, world, 0, k, lo, mine, verify=(gathered, status))
    assert int(status.item()) == 1
    status.zero_()
    over = tot_all.clone()
    over[world - 1, bins, 0] = 1                                                # the last rank flagged a list overflow
    st.topk_place(plan, cap, cand, cnt, over, world, 0, k, lo, mine, verify=(gathered, status))
    assert int(status.item()) == 1
    with pytest.raises(R.CmhError):                                             # one block instead of the gathered blocks
        st.topk_cutoff_sharded(shards[0][0], gathered[0], k, 0, world)


def test_topk_from_host_slab_pipeline_equals_resident_path():
    """host +-1 fp32 codes streamed in slabs (H2D overlapped with pack / expand / collect) == codes resident on the GPU."""
    from clip_based_cross_modal_hash_b200 import calc_utils as cu

    Q, N, K, k = 500, 330_001, 64, 300
    qB, rB = synth.random_codes(Q, K, 71), synth.random_codes(N, K, 72)
    want = R.topk(R.pack_codes(qB.to(DEV)), R.pack_codes(rB.to(DEV)), K, k, exact=True)
    got = R.topk_from_host(qB.pin_memory(), rB.pin_memory(), k, DEV)
    assert torch.equal(got, want)
    got2 = R.topk_from_host(qB, rB, k, DEV, slabs=3)          # pageable memory, other slab count
    assert torch.equal(got2, want)
    dist, idx = cu.hamming_topk(qB, rB, k)                     # the public call takes this path for host inputs
    wd, wi = R.split_keys(want)
    assert torch.equal(idx, wi.cpu()) and torch.equal(dist, wd.cpu().float())
    bad = rB.clone()
    bad[N - 5, 3] = 0.0                                        # a non +-1 element in the LAST slab is still reported
    with pytest.raises(ValueError):
        cu.hamming_topk(qB, bad, k)


def test_topk_graph_replays_equal_exact_path():
    """the whole step captured as one CUDA graph: replays on refreshed static inputs == the exact two-pass path; the device-side
    verification still routes a failed candidate pass to the exact path; codes that are not +-1 are reported."""
    Q, N, K, k = 400, 300_000, 64, 500
    q = synth.random_codes(Q, K, 81).to(DEV)
    g = synth.random_codes(N, K, 82).to(DEV)
    graph = R.TopkGraph(q, g, k)
    assert graph.kernels >= 8
    for seed in (82, 83, 84):
        g.copy_(synth.random_codes(N, K, seed).to(DEV))          # refreshed in place: same buffers, new gallery
        want = R.topk(R.pack_codes(q), R.pack_codes(g), K, k, exact=True)
        assert torch.equal(graph.run(), want)
    # a gallery whose sample prefix looks nothing like the rest: cutoffs too tight -> flagged on the device -> exact path
    near = synth.random_codes(1, K, 8).to(DEV)
    g[N - 40_000:] = near
    q[:100] = near
    want = R.topk(R.pack_codes(q), R.pack_codes(g), K, k, exact=True)
    assert torch.equal(graph.run(), want)
    g[17, 5] = 0.5
    with pytest.raises(ValueError):
        graph.run()
    with pytest.raises(R.CmhError):                                # too small for the candidate path
        R.TopkGraph(q[:10], g[:1000].contiguous(), 100)
