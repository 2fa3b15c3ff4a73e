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
    def make_plan(self, Q, N, nbits, ncls, N_geom=None, target_blocks=0):
        return _lib.make_plan(Q, N, nbits, ncls, N_geom, target_blocks or 64)

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

    def hist(self, plan, qp, qlp, gp, glp):
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

    def rank_map(self, plan, qp, qlp, gp, glp, sc, tindex=None, n_total=None):
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

    def rank_topk(self, plan, qp, gp, sc, k, idx_offset, keys=None):
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
