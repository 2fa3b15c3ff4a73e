"""TEST INFRASTRUCTURE ONLY — numpy stand-in for ``retrieval.CudaStages`` (same stage contracts, CPU tensors).

Lets the CPU suite exercise the host-side exchange logic of ``ShardedEvaluator`` (shard geometry, gather
layouts, merge) over gloo without a GPU.  Built on oracle/hamming_oracle.py; never used by the product.
"""
import numpy as np
import torch

from clip_based_cross_modal_hash_b200 import _lib
from oracle import hamming_oracle as ho

EMPTY = np.uint64(0xFFFFFFFFFFFFFFFF)


def _u32(t):
    return t.numpy().view(np.uint32)


class OracleStages:
    """``tensor_cores=True`` additionally offers the stage contracts of the single-pass candidate path (cmh_tc_topk_*), so that
    ``ShardedEvaluator._topk_candidates`` — gathers, global cutoff, verification, key exchange, fallback — runs over gloo."""

    def __init__(self, tensor_cores=False):
        self.tensor_cores = tensor_cores

    def make_plan(self, Q, N, nbits, ncls, N_geom=None, target_blocks=0):
        return _lib.make_plan(Q, N, nbits, ncls, N_geom, target_blocks or 64)

    # ---- candidate path (numpy restatement of the contracts in include/cmh.h, "Candidate path of the top-k") ----
    def operands(self, plan, qp, qlp, gp, glp):
        return (qp, gp) if self.tensor_cores else None

    def topk_sample_block(self, plan, plan_s, hist_s, idx_offset, rank, world, device):
        out = np.zeros((plan.bins + 1, plan.Qpad), dtype=np.uint32)
        n_s = 0
        if hist_s is not None:
            out[: plan.bins] = (_u32(hist_s).astype(np.int64) & 0xFFFF).sum(axis=0)
            n_s = plan_s.N
        out[plan.bins, 0], out[plan.bins, 1], out[plan.bins, 2 + rank] = n_s, plan.N, idx_offset
        return torch.from_numpy(out.view(np.int32))

    @staticmethod
    def _cutoff(tot, need, n_items):
        """T = first distance whose prefix count reaches ``need``; bound = ceil(fraction of bucket T still needed * n_items)."""
        cum, T, frac, found = 0, len(tot) - 1, 1.0, False
        for d, t in enumerate(tot):
            if not found and cum + t >= need:
                T, found, frac = d, True, (need - cum) / float(t)
            cum += t
        return T, (float(np.ceil(frac * n_items)) if found else float(n_items))

    def topk_cutoff(self, plan_s, hist_s, n_local, k):
        tot = (_u32(hist_s).astype(np.int64) & 0xFFFF).sum(axis=0)            # [bins, Qpad]
        ks = min(k, n_local) * plan_s.N / float(n_local)
        need = ks + 5.0 * np.sqrt(ks) + 2.0
        cut = np.full((2, plan_s.Qpad), -1, dtype=np.int32)
        for q in range(plan_s.Qpad):
            T, ib = self._cutoff(tot[:, q], need, n_local)
            cut[0, q] = T if q < plan_s.Q else -1
            cut[1, q] = int(min(ib, n_local))
        return torch.from_numpy(cut)

    def topk_cutoff_sharded(self, plan, sample_all, k, rank, world):
        s = _u32(sample_all).astype(np.int64)                                 # [world, bins + 1, Qpad]
        head = s[:, plan.bins]
        n_sample, n_total = head[:, 0].sum(), head[:, 1].sum()
        offs = np.array([head[r, 2 + r] for r in range(world)])
        shard_lo = offs[rank] - offs.min()
        ks = (min(k, n_total) * n_sample / float(n_total)) if n_total > 0 else 0.0
        need = ks + 5.0 * np.sqrt(ks) + 2.0
        tot = s[:, : plan.bins].sum(axis=0)
        cut = np.full((2, plan.Qpad), -1, dtype=np.int32)
        for q in range(plan.Qpad):
            T, ib = self._cutoff(tot[:, q], need, n_total)
            cut[0, q] = T if q < plan.Q else -1
            cut[1, q] = int(min(max(ib - shard_lo, -1.0), plan.N))
        return torch.from_numpy(cut)

    def topk_collect(self, plan, ops, cutoff, cap, out=None, chunks=None):
        qp, gp = ops
        cut = cutoff.numpy()
        cand = np.zeros((plan.nchunks, plan.Qpad, cap), dtype=np.uint32)
        cnt = np.zeros((plan.nchunks, plan.Qpad), dtype=np.uint32)
        for c, (lo, hi) in enumerate(self._chunks(plan)):
            if lo >= hi:
                continue
            d, _ = self._dist_rel(qp, None, gp[lo:hi], None)
            idx = np.arange(lo, hi)
            for q in range(plan.Q):
                keep = (d[q] < cut[0, q]) | ((d[q] == cut[0, q]) & (idx <= cut[1, q]))
                items = np.nonzero(keep)[0]
                if items.size > cap:
                    cnt[c, q] = 0xFFFFFFFF
                    continue
                cand[c, q, : items.size] = (d[q][items].astype(np.uint32) << 24) | items.astype(np.uint32)
                cnt[c, q] = items.size
        return torch.from_numpy(cand.view(np.int32)), torch.from_numpy(cnt.view(np.int32))

    def topk_count(self, plan, cap, cand, cnt, k):
        ca, cn = _u32(cand), _u32(cnt)
        tot = np.zeros((plan.bins + 1, plan.Qpad), dtype=np.uint32)
        need = min(k, plan.N)
        for q in range(plan.Q):
            for c in range(plan.nchunks):
                if cn[c, q] == 0xFFFFFFFF:
                    tot[plan.bins, 0] |= 1
                    continue
                tot[: plan.bins, q] += np.bincount(ca[c, q, : cn[c, q]] >> 24, minlength=plan.bins).astype(np.uint32)
            if k > 0 and tot[: plan.bins, q].sum() < need:
                tot[plan.bins, 0] |= 1
        return torch.from_numpy(tot.view(np.int32))

    def topk_place(self, plan, cap, cand, cnt, totals_all, world, rank, k, idx_offset, keys, peers_dev=None, npeers=0,
                   multicast=None, verify=None):
        ca, cn = _u32(cand), _u32(cnt)
        out = keys.numpy().view(np.uint64)
        if totals_all is None:
            local = _u32(self.topk_count(plan, cap, cand, cnt, 0))
            t = local[None].astype(np.int64)
        else:
            t = _u32(totals_all).astype(np.int64)                             # [world, bins + 1, Qpad]
        sample_all, status = verify if verify is not None else (None, None)
        if status is not None:
            n_total = plan.N if sample_all is None else int(_u32(sample_all).astype(np.int64)[:, plan.bins, 1].sum())
            if t[:, plan.bins, 0].max() != 0 or (t[:, : plan.bins, : plan.Q].sum(axis=(0, 1)) < min(k, n_total)).any():
                status |= 1
        alls = t[:, : plan.bins].sum(axis=0)
        below = np.cumsum(alls, axis=0) - alls + t[:rank, : plan.bins].sum(axis=0)
        for q in range(plan.Q):
            run = below[:, q].copy()
            for c in range(plan.nchunks):
                if cn[c, q] == 0xFFFFFFFF:
                    continue
                first = idx_offset + c * plan.chunk_items
                for e in ca[c, q, : cn[c, q]]:
                    d = int(e >> 24)
                    r = run[d]
                    run[d] += 1
                    if r < k:
                        out[q, r] = (np.uint64(d) << np.uint64(32)) | np.uint64(first + int(e & 0xFFFFFF))
        return keys

    @staticmethod
    def _chunks(plan):
        return [(c * plan.chunk_items, min((c + 1) * plan.chunk_items, plan.N)) for c in range(plan.nchunks)]

    @staticmethod
    def _dist_rel(qp, qlp, gp, glp):
        d = ho.hamming_matrix(_u32(qp), _u32(gp)).astype(np.int64)
        if qlp is None:
            return d, np.zeros_like(d, dtype=bool)
        rel = ((_u32(qlp)[:, None, :] & _u32(glp)[None, :, :]) != 0).any(axis=2)
        return d, rel

    def hist(self, plan, qp, qlp, gp, glp, ops=None):
        out = np.zeros((plan.nchunks, plan.bins, plan.Qpad), dtype=np.uint32)
        for c, (lo, hi) in enumerate(self._chunks(plan)):
            if lo >= hi:
                continue
            d, rel = self._dist_rel(qp, qlp, gp[lo:hi], None if glp is None else glp[lo:hi])
            for q in range(plan.Q):
                ha = np.bincount(d[q], minlength=plan.bins)
                hr = np.bincount(d[q][rel[q]], minlength=plan.bins)
                out[c, :, q] = (hr.astype(np.uint32) << 16) | ha.astype(np.uint32)
        return torch.from_numpy(out.view(np.int32))

    def scan(self, plan, hist_all, world, rank, k, with_rel=True):
        h = _u32(hist_all.reshape(-1, plan.bins, plan.Qpad)).astype(np.int64)
        ha, hr = h & 0xFFFF, h >> 16
        # exclusive prefix over chunks inside each bucket
        wa = np.cumsum(ha, axis=0) - ha
        wr = np.cumsum(hr, axis=0) - hr
        ta, tr = ha.sum(axis=0), hr.sum(axis=0)                 # [bins, Qpad]
        ba = np.cumsum(ta, axis=0) - ta
        br = np.cumsum(tr, axis=0) - tr
        lo, hi = rank * plan.nchunks, (rank + 1) * plan.nchunks
        tsum = tr.sum(axis=0)
        kk = int(k) if k else 0
        total = np.minimum(tsum, kk) if kk > 0 else tsum
        cum = np.cumsum(ta, axis=0)
        thresh = np.full(plan.Qpad, plan.bins - 1, dtype=np.int64)
        if kk > 0:
            for q in range(plan.Qpad):
                hit = np.nonzero(cum[:, q] >= kk)[0]
                if hit.size:
                    thresh[q] = hit[0]
        t32 = lambda a: torch.from_numpy(np.ascontiguousarray(a).astype(np.int32))
        return {"within_all": t32(wa[lo:hi]), "within_rel": t32(wr[lo:hi]) if with_rel else None,
                "below_all": t32(ba), "below_rel": t32(br) if with_rel else None,
                "tsum": t32(tsum), "total": t32(total), "thresh": t32(thresh)}

    def hist_totals(self, plan, hist):
        h = _u32(hist).astype(np.int64)
        return torch.from_numpy(np.stack([(h & 0xFFFF).sum(axis=0), (h >> 16).sum(axis=0)]).astype(np.int32))

    def scan_sharded(self, plan, hist_local, totals_all, world, rank, k):
        """Same contract as CudaStages.scan_sharded: rebuild the global view from the totals (every other rank's block is
        replaced by one pseudo-chunk holding its totals — only sums over lower ranks and over all ranks are used)."""
        h = _u32(hist_local).astype(np.int64)
        ha, hr = h & 0xFFFF, h >> 16
        t = totals_all.numpy().astype(np.int64)                 # [world, 2, bins, Qpad]
        base_a, base_r = t[:rank, 0].sum(axis=0), t[:rank, 1].sum(axis=0)
        wa = base_a[None] + np.cumsum(ha, axis=0) - ha
        wr = base_r[None] + np.cumsum(hr, axis=0) - hr
        ta, tr = t[:, 0].sum(axis=0), t[:, 1].sum(axis=0)
        ba = np.cumsum(ta, axis=0) - ta
        br = np.cumsum(tr, axis=0) - tr
        tsum = tr.sum(axis=0)
        kk = int(k) if k else 0
        total = np.minimum(tsum, kk) if kk > 0 else tsum
        cum = np.cumsum(ta, axis=0)
        thresh = np.full(plan.Qpad, plan.bins - 1, dtype=np.int64)
        if kk > 0:
            for q in range(plan.Qpad):
                hit = np.nonzero(cum[:, q] >= kk)[0]
                if hit.size:
                    thresh[q] = hit[0]
        t32 = lambda a: torch.from_numpy(np.ascontiguousarray(a).astype(np.int32))
        return {"within_all": t32(wa), "within_rel": t32(wr), "below_all": t32(ba), "below_rel": t32(br),
                "tsum": t32(tsum), "total": t32(total), "thresh": t32(thresh)}

    def _positions(self, drow, bins):
        """in-bucket position (index order) of every item of a chunk."""
        ranks = ho.stable_ranks(drow, bins)
        hist = np.bincount(drow, minlength=bins)
        below = np.concatenate(([0], np.cumsum(hist)[:-1]))
        return ranks - below[drow]

    def rank_map(self, plan, qp, qlp, gp, glp, sc, tindex=None, n_total=None, ops=None):
        ap = np.zeros((plan.nchunks, plan.Qpad), dtype=np.float64)
        wa, wr = sc["within_all"].numpy(), sc["within_rel"].numpy()
        ba, br = sc["below_all"].numpy(), sc["below_rel"].numpy()
        total = sc["total"].numpy()
        for c, (lo, hi) in enumerate(self._chunks(plan)):
            if lo >= hi:
                continue
            d, rel = self._dist_rel(qp, qlp, gp[lo:hi], glp[lo:hi])
            for q in range(plan.Q):
                pos_all = self._positions(d[q], plan.bins)
                a = ba[d[q], q] + wa[c, d[q], q] + pos_all                       # 0-based global stable rank
                dr = d[q][rel[q]]
                pos_rel = self._positions(dr, plan.bins)
                r = br[dr, q] + wr[c, dr, q] + pos_rel
                ar = a[rel[q]]
                keep = r < total[q]
                cnt = (r[keep] + 1).astype(np.float32)
                tix = ar[keep].astype(np.float32) + np.float32(1.0)
                ap[c, q] = (cnt / tix).astype(np.float64).sum()
                if tindex is not None:
                    sel = keep & (r < tindex.shape[1])
                    tindex[q, torch.from_numpy(r[sel])] = torch.from_numpy((ar[sel] + 1).astype(np.int32))
        return torch.from_numpy(ap)

    def ap_reduce(self, plan, ap_partial):
        parts = ap_partial.reshape(-1, plan.Qpad).numpy()
        out = np.zeros(plan.Qpad, dtype=np.float64)
        for p in parts:                                                           # chunk order, like ap_reduce_kernel
            out = out + p
        return torch.from_numpy(out[None, :].copy())

    def map_finish(self, plan, ap_partial_all, total):
        parts = ap_partial_all.reshape(-1, plan.Qpad).numpy()
        with np.errstate(invalid="ignore", divide="ignore"):
            ap = parts[:, : plan.Q].sum(axis=0) / total.numpy()[: plan.Q].astype(np.float64)
        return torch.from_numpy(ap), torch.tensor(ap.sum() / plan.Q, dtype=torch.float64)

    def rank_topk(self, plan, qp, gp, sc, k, idx_offset, keys=None, ops=None):
        out = np.full((plan.Q, k), EMPTY, dtype=np.uint64)
        wa, ba, th = sc["within_all"].numpy(), sc["below_all"].numpy(), sc["thresh"].numpy()
        for c, (lo, hi) in enumerate(self._chunks(plan)):
            if lo >= hi:
                continue
            d, _ = self._dist_rel(qp, None, gp[lo:hi], None)
            for q in range(plan.Q):
                a = ba[d[q], q] + wa[c, d[q], q] + self._positions(d[q], plan.bins)
                sel = (d[q] <= th[q]) & (a < k)
                idx = np.arange(lo, hi)[sel] + idx_offset
                out[q, a[sel]] = (d[q][sel].astype(np.uint64) << np.uint64(32)) | idx.astype(np.uint64)
        return torch.from_numpy(out.view(np.int64))

    def topk_merge(self, parts):
        world, Q, k = parts.shape
        p = parts.numpy().view(np.uint64)
        out = np.sort(np.concatenate([p[s] for s in range(world)], axis=1), axis=1)[:, :k]
        return torch.from_numpy(np.ascontiguousarray(out).view(np.int64))
