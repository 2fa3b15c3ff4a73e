/* TEST INFRASTRUCTURE ONLY — plain-C restatement of the retrieval evaluator's integer stages.
 *
 * Restates common/calc_utils.py:51-92 (reference: kalenforn/clip-based-cross-modal-hash) on
 * bit-packed codes so that parity tests can run at sizes where the torch/numpy oracles take too long
 * (e.g. 2k x 200k).  It is a checker: nothing in the product links or calls it.
 *
 *   calc_hammingDist (calc_utils.py:51-56)  0.5*(K - q.g) on +-1 codes  ==  popcount(q XOR g)
 *   gnds             (calc_utils.py:72)     (qL.rL > 0) on 0/1 labels    ==  (qmask AND gmask) != 0
 *   torch.sort       (calc_utils.py:77)     canonical stable order        ==  key (dist, index)
 *   tindex           (calc_utils.py:85-88)  1-based ranks of the first min(R,k) relevant items
 *
 * Formulation: two sequential sweeps per query over the gallery in index order with (K+1)-bin
 * counters — deliberately the dumbest correct thing, no tiling, no tricks.  Single-threaded; the Python wrapper (oracle/c_oracle.py) fans query
 * ranges out over host threads (ctypes drops the GIL).
 *
 * Layout: codes [n][W] uint32 (bit b of word w = column 32w+b, 1 iff value > 0); labels [n][4] uint32.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_LABEL_WORDS 4
#define ORACLE_MAX_BINS 1025

static inline int dist_of(const uint32_t* a, const uint32_t* b, int W) {
    int d = 0;
    for (int w = 0; w < W; ++w) d += __builtin_popcount(a[w] ^ b[w]);
    return d;
}

static inline int rel_of(const uint32_t* a, const uint32_t* b) {
    uint32_t r = 0;
    for (int w = 0; w < ORACLE_LABEL_WORDS; ++w) r |= a[w] & b[w];
    return r != 0;
}

/* +-1 floats -> packed words; returns the number of elements that are not exactly +1 or -1. */
int64_t oracle_pack_codes_f32(const float* codes, int64_t n, int K, uint32_t* out) {
    int W = (K + 31) / 32;
    int64_t bad = 0;
    memset(out, 0, (size_t)n * W * sizeof(uint32_t));
    for (int64_t i = 0; i < n; ++i)
        for (int c = 0; c < K; ++c) {
            float v = codes[i * K + c];
            if (v > 0.0f) out[i * W + c / 32] |= 1u << (c % 32);
            if (v != 1.0f && v != -1.0f) ++bad;
        }
    return bad;
}

/* int64 multi-hot labels -> 4-word masks; returns the number of entries outside {0,1}. */
int64_t oracle_pack_labels_i64(const int64_t* labels, int64_t n, int C, uint32_t* out) {
    int64_t bad = 0;
    memset(out, 0, (size_t)n * ORACLE_LABEL_WORDS * sizeof(uint32_t));
    for (int64_t i = 0; i < n; ++i)
        for (int c = 0; c < C; ++c) {
            int64_t v = labels[i * C + c];
            if (v != 0) out[i * ORACLE_LABEL_WORDS + c / 32] |= 1u << (c % 32);
            if (v != 0 && v != 1) ++bad;
        }
    return bad;
}

/* Full distance matrix, uint16 [Q][N]. */
void oracle_hamming_u16(const uint32_t* qp, int64_t Q, const uint32_t* gp, int64_t N, int W,
                        uint16_t* out) {
    for (int64_t q = 0; q < Q; ++q)
        for (int64_t j = 0; j < N; ++j)
            out[q * N + j] = (uint16_t)dist_of(qp + q * W, gp + j * W, W);
}

/* Integer stage of calc_map_k.
 *   tindex  [Q][cap] int32: ascending 1-based ranks of the first totals[q] relevant items
 *                           (entries >= totals[q] are left untouched)
 *   totals  [Q] = min(R, k), tsums [Q] = R
 *   hist_all/hist_rel [Q][K+1] may be NULL.
 * Returns 0, or -1 if cap < some totals[q] (nothing is written past cap).
 */
int oracle_map_tindex(const uint32_t* qp, const uint32_t* qlp, int64_t Q, const uint32_t* gp,
                      const uint32_t* glp, int64_t N, int K, int64_t k, int32_t* tindex,
                      int64_t cap, int32_t* totals, int32_t* tsums, int32_t* hist_all,
                      int32_t* hist_rel) {
    int W = (K + 31) / 32;
    int bins = K + 1;
    int overflow = 0;
    if (bins > ORACLE_MAX_BINS) return -2;
    for (int64_t q = 0; q < Q; ++q) {
        int64_t ha[ORACLE_MAX_BINS], hr[ORACLE_MAX_BINS], ba[ORACLE_MAX_BINS], br[ORACLE_MAX_BINS];
        memset(ha, 0, sizeof(ha));
        memset(hr, 0, sizeof(hr));
        const uint32_t* qc = qp + q * W;
        const uint32_t* ql = qlp + q * ORACLE_LABEL_WORDS;
        for (int64_t j = 0; j < N; ++j) { /* sweep 1: histograms */
            int d = dist_of(qc, gp + j * W, W);
            ha[d]++;
            hr[d] += rel_of(ql, glp + j * ORACLE_LABEL_WORDS);
        }
        int64_t ca = 0, cr = 0;
        for (int d = 0; d < bins; ++d) { /* exclusive prefix = rank base of each distance bucket */
            ba[d] = ca; br[d] = cr;
            ca += ha[d]; cr += hr[d];
            if (hist_all) hist_all[q * bins + d] = (int32_t)ha[d];
            if (hist_rel) hist_rel[q * bins + d] = (int32_t)hr[d];
        }
        int64_t total = cr < k ? cr : k;
        tsums[q] = (int32_t)cr;
        totals[q] = (int32_t)total;
        if (total > cap) { overflow |= 1; total = cap; }
        for (int64_t j = 0; j < N; ++j) { /* sweep 2: ranks, in gallery-index order */
            int d = dist_of(qc, gp + j * W, W);
            int64_t rank_all = ba[d]++; /* 0-based position in the stable (dist, index) order */
            if (rel_of(ql, glp + j * ORACLE_LABEL_WORDS)) {
                int64_t rank_rel = br[d]++;
                if (rank_rel < total) tindex[q * cap + rank_rel] = (int32_t)(rank_all + 1);
            }
        }
    }
    return overflow ? -1 : 0;
}

/* Per-query k smallest (dist, index): out_dist/out_idx [Q][k] int32, unused slots = -1. */
void oracle_topk(const uint32_t* qp, int64_t Q, const uint32_t* gp, int64_t N, int K, int64_t k,
                 int32_t* out_dist, int32_t* out_idx) {
    int W = (K + 31) / 32;
    int bins = K + 1;
    for (int64_t q = 0; q < Q; ++q) {
        int64_t base[ORACLE_MAX_BINS + 1];
        memset(base, 0, sizeof(base));
        const uint32_t* qc = qp + q * W;
        for (int64_t j = 0; j < N; ++j) base[dist_of(qc, gp + j * W, W) + 1]++;
        for (int d = 0; d < bins; ++d) base[d + 1] += base[d];
        for (int64_t s = 0; s < k; ++s) { out_dist[q * k + s] = -1; out_idx[q * k + s] = -1; }
        for (int64_t j = 0; j < N; ++j) {
            int d = dist_of(qc, gp + j * W, W);
            int64_t r = base[d]++;
            if (r < k) { out_dist[q * k + r] = d; out_idx[q * k + r] = (int32_t)j; }
        }
    }
}
