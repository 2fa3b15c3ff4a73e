/* cmh.h — C-ABI of the B200-native cross-modal-hash hot path (libcmh.so).
 *
 * The reference (kalenforn/clip-based-cross-modal-hash) is pure Python/PyTorch and has no FFI of its
 * own; its "operator API" for this path is a set of Python functions.  Each entry
 * point below states the reference function (file:line, relative to the reference root) whose
 * arithmetic it replaces.  The Python host shim (clip_based_cross_modal_hash_b200/calc_utils.py) binds
 * these with ctypes and re-exposes the reference signatures; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the library allocates nothing: callers pass output buffers and a workspace (sizes from cmh_plan);
 *   - every launch goes to the `stream` argument (a cudaStream_t passed as void*; NULL = default stream);
 *   - return value 0 = CMH_OK, negative = error; cmh_last_error() gives a per-thread message;
 *   - calls are asynchronous w.r.t. the host unless stated otherwise.
 *
 * Packed layouts ("data layout in HBM")
 *   codes   [n][W]  uint32, W = cmh_code_words(nbits) in {1,2,4}; bit b of word w = column 32w+b of the
 *                   reference's +-1 float code matrix, 1 iff value > 0 (runners/base.py:407-410 sign_()).
 *   labels  [n][LW] uint32, LW = cmh_label_words(ncls) in {1,2,4}; bit b of word w = class 32w+b, 1 iff != 0
 *                   (dataset/transformer_dataset.py:95-100 int64 multi-hot).
 *   keys    uint64 = (uint64(distance) << 32) | gallery_index   — total order == stable (dist, index) sort.
 */
#ifndef CMH_H_
#define CMH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMH_ABI_VERSION 2   /* 2: round 2 (tensor-core ranking, candidate top-k, multicast exchanges, training tail) */

#define CMH_OK 0
#define CMH_ERR_INVALID (-1)     /* bad argument (NULL pointer, negative size, misaligned buffer) */
#define CMH_ERR_CUDA (-2)        /* a CUDA runtime call or launch failed; see cmh_last_error()   */
#define CMH_ERR_WORKSPACE (-3)   /* workspace too small                                           */
#define CMH_ERR_UNSUPPORTED (-4) /* nbits > 128 or ncls > 128                                     */

#define CMH_QTILE 128            /* queries per thread block; Qpad = roundup(Q, CMH_QTILE)        */
#define CMH_MAX_BITS 128
#define CMH_MAX_CLASSES 128

/* label dtypes accepted by cmh_pack_labels */
#define CMH_DT_I64 0
#define CMH_DT_F32 1
#define CMH_DT_U8 2
#define CMH_DT_I32 3

int cmh_abi_version(void);
const char* cmh_last_error(void);
/* Kernels this library has launched in this process so far (every stream, every thread): what bench.py reports as
 * gpu_launches.  Host-side counter, no device synchronisation. */
unsigned long long cmh_launch_count(void);
/* SM count / compute capability of the current device (synchronous). */
int cmh_device_info(int* sm_count, int* cc_major, int* cc_minor);

int cmh_code_words(int nbits);  /* device words per code, <0 if unsupported  */
int cmh_label_words(int ncls);  /* device words per label mask, 0 if ncls==0 */

/* ---- R0: packing ------------------------------------------------------------------------------------
 * Replaces the +-1 fp32 code buffers of runners/base.py:245-257 (get_code) and the int64 label matrices
 * of dataset/transformer_dataset.py:95-100 as the evaluator's input format.
 * `ld` = row stride in elements.  `bad_count` (device u64, may be NULL) is incremented by the number of
 * elements that are not exactly +1/-1 (codes) or 0/1 (labels); it is NOT zeroed by the call. */
int cmh_pack_codes_f32(const float* codes, int64_t n, int nbits, int64_t ld, uint32_t* out,
                       unsigned long long* bad_count, void* stream);
int cmh_pack_labels(const void* labels, int dtype, int64_t n, int ncls, int64_t ld, uint32_t* out,
                    unsigned long long* bad_count, void* stream);
/* packed -> +-1 fp32 rows (the reference's code format, e.g. for save_mat runners/base.py:386-405). */
int cmh_unpack_codes_f32(const uint32_t* packed, int64_t n, int nbits, float* out, int64_t ld, void* stream);

/* ---- R1: calc_hammingDist (common/calc_utils.py:51-56) ----------------------------------------------
 * out[q][j] = popcount(qcodes[q] ^ gcodes[j]) as fp32 == 0.5*(K - B1.B2^T) for +-1 inputs.
 * Materialises Q x N floats (HBM-write bound); the evaluator below never does. */
int cmh_hamming_f32(const uint32_t* qcodes, int64_t Q, const uint32_t* gcodes, int64_t N, int nbits,
                    float* out, int64_t ld_out, void* stream);
/* General fp32 inputs (codes containing 0 after sign_(), or un-binarised activations):
 * out = 0.5*(K - B1.B2^T) accumulated in fp32 in ascending column order. */
int cmh_hamming_dense_f32(const float* B1, int64_t Q, const float* B2, int64_t N, int nbits, float* out,
                          void* stream);

/* ---- evaluator geometry -----------------------------------------------------------------------------
 * The gallery shard [0,N) of this rank is cut into `nchunks` contiguous chunks of `chunk_items` items
 * (<= 65024, multiple of 512); one thread block ranks CMH_QTILE queries against one chunk.
 * N_geom >= N is the largest shard of any rank (all ranks must use the same geometry so that the
 * all-gathered histograms line up; single GPU: N_geom = N). */
typedef struct cmh_plan {
    int64_t Q, N, N_geom, Qpad;
    int32_t nbits, ncls, W, LW, bins, nchunks;
    int64_t chunk_items;
    int64_t hist_elems;   /* uint32 elements of one rank's histogram block: nchunks*bins*Qpad         */
    int64_t within_elems; /* uint32 elements of within_all / within_rel (each): nchunks*bins*Qpad     */
    int64_t below_elems;  /* uint32 elements of below_all / below_rel (each): bins*Qpad               */
    int64_t ap_elems;     /* fp64 elements of one rank's AP partials: nchunks*Qpad                    */
    int64_t workspace_bytes; /* for the one-shot cmh_map_k / cmh_topk calls (world = 1)               */
} cmh_plan;

/* `target_blocks` <= 0 picks 16 blocks per SM of the current device (synchronous device query on the
 * first call only). */
int cmh_make_plan(int64_t Q, int64_t N, int64_t N_geom, int nbits, int ncls, int target_blocks, cmh_plan* plan);

/* ---- R1+R2 pass 1: per-(query, chunk) distance histograms ---------------------------------------------
 * Replaces hamms = calc_hammingDist(qB, rB) and gnds = (query_L.mm(retrieval_L.T) > 0)
 * (common/calc_utils.py:72,76) without materialising either.
 * hist[c][d][q] = (#relevant items of chunk c at distance d) << 16 | (#items at distance d).
 * qlabels/glabels may be NULL (top-k only: relevance counts are 0). */
int cmh_hist(const cmh_plan* plan, const uint32_t* qcodes, const uint32_t* qlabels, const uint32_t* gcodes,
             const uint32_t* glabels, uint32_t* hist, void* stream);

/* ---- R3 rank bases: exclusive prefix over (distance, rank, chunk) -------------------------------------
 * Replaces torch.sort(hamms) (common/calc_utils.py:77) by counting: stable rank of an item =
 * #items at smaller distance + #items at equal distance and lower gallery index.
 * hist_all = `world` rank blocks laid out [world][nchunks][bins][Qpad] (all-gathered; world=1: own block).
 * Outputs (this rank's chunks only): within_*[c][d][q], below_*[d][q], tsum[q] = R_q (calc_utils.py:75),
 * total[q] = min(R_q, k) (calc_utils.py:81), thresh[q] = distance bucket holding global rank k-1.
 * k <= 0 means "no limit" (k=None in the reference). */
int cmh_scan(const cmh_plan* plan, const uint32_t* hist_all, int world, int rank, int64_t k,
             uint32_t* within_all, uint32_t* within_rel, uint32_t* below_all, uint32_t* below_rel,
             int32_t* tsum, int32_t* total, int32_t* thresh, void* stream);

/* Sharded form of the same scan (what ShardedEvaluator.map_k uses): a rank's chunks follow every chunk of the lower ranks
 * in gallery order, so the other ranks only have to contribute their per-bucket totals.
 * cmh_hist_totals: totals[0][d][q] = #items of this rank at distance d, totals[1][d][q] = #relevant ones
 *                  (uint32 [2][bins][Qpad]; summed over this rank's chunks).
 * cmh_scan_sharded: hist_local = this rank's block [nchunks][bins][Qpad]; totals_all = the all-gathered totals
 *                  [world][2][bins][Qpad].  Outputs as cmh_scan.  Exchange volume per rank: 2*bins*Qpad*4 bytes
 *                  (C2: 2.6 MB) instead of nchunks*bins*Qpad*4 (C2: 78 MB). */
int cmh_hist_totals(const cmh_plan* plan, const uint32_t* hist, uint32_t* totals, void* stream);
int cmh_scan_sharded(const cmh_plan* plan, const uint32_t* hist_local, const uint32_t* totals_all, int world, int rank,
                     int64_t k, uint32_t* within_all, uint32_t* within_rel, uint32_t* below_all, uint32_t* below_rel,
                     int32_t* tsum, int32_t* total, int32_t* thresh, void* stream);

/* ---- R4 pass 2 (mAP): ranks of the relevant items and their AP terms ----------------------------------
 * Replaces the per-query loop common/calc_utils.py:84-89.
 * ap_partial[c][q] = fp64 sum over this chunk's relevant items j with relevant-rank r_j < total[q] of
 *                    fp32(r_j + 1) / fp32(rank_j + 1)      (the same fp32 quotients the reference forms)
 * tindex (may be NULL): tindex[q*cap + r_j] = rank_j + 1 for r_j < min(total[q], cap)  (int32; the
 * reference's `tindex` at calc_utils.py:88) — entries owned by other ranks are left untouched.
 * n_total = gallery items over ALL ranks (= plan->N on one GPU); below 2^24 the kernel keeps its running
 * ranks in fp32 (exact there), which removes the int->float conversions of the AP terms. */
int cmh_rank_map(const cmh_plan* plan, const uint32_t* qcodes, const uint32_t* qlabels,
                 const uint32_t* gcodes, const uint32_t* glabels, const uint32_t* within_all,
                 const uint32_t* within_rel, const uint32_t* below_all, const uint32_t* below_rel,
                 const int32_t* total, int64_t n_total, double* ap_partial, int32_t* tindex, int64_t cap,
                 void* stream);

/* Sharded runs: a rank folds its chunk partials into ONE fp64 per query before the exchange (41 KB per rank at C2 instead of
 * 2.4 MB): out[q] = sum_c ap_partial[c][q], chunk order, out = [Qpad]. */
int cmh_ap_reduce(const cmh_plan* plan, const double* ap_partial, int nparts, double* out, void* stream);

/* ap[q] = (sum over `nparts` chunk partials [nparts][Qpad], in index order) / total[q];
 * *map_out = (sum_q ap[q]) / Q in fp64, fixed order (calc_utils.py:89-90).  nan if any total is 0,
 * like the reference.  ap may be NULL. */
int cmh_map_finish(const cmh_plan* plan, const double* ap_partial, int nparts, const int32_t* total,
                   double* ap, double* map_out, void* stream);

/* ---- R3 pass 2 (top-k): the first k entries of the stable ranking --------------------------------------
 * keys[q*k + rank] = (dist << 32) | (idx_offset + local index) for every item of this shard with
 * rank < k; other slots untouched (pre-fill with cmh_fill_keys when the buffer is not owned by one
 * rank alone or when k > N). */
int cmh_rank_topk(const cmh_plan* plan, const uint32_t* qcodes, const uint32_t* gcodes,
                  const uint32_t* within_all, const uint32_t* below_all, const int32_t* thresh, int64_t k,
                  int64_t idx_offset, uint64_t* keys, void* stream);
int cmh_fill_keys(uint64_t* keys, int64_t count, uint64_t value, void* stream);

/* ---- T: the ranking passes on the tensor cores (tcgen05.mma.kind::i8) -------------------------------------------------
 * +-1 codes make calc_hammingDist a dense contraction (it IS one in the reference: common/calc_utils.py:55); so is the label
 * test query_L.mm(retrieval_L.T) > 0 (calc_utils.py:72).  The cmh_tc_* entry points compute both on the tensor pipe
 * (int8 operands, int32 accumulators in tensor memory) and keep the counting formulation of cmh_hist / cmh_rank_topk /
 * cmh_rank_map unchanged: same plan, same outputs bit for bit, the scan entry points above are shared.
 * Operands are int8 rows expanded from the bit-packed words (caller-owned buffers, 16-byte aligned):
 *   codes  [rows][code_bytes]   +1 / -1, zero beyond nbits; code_bytes  = cmh_tc_operand_bytes(nbits) in {32, 64, 128}
 *   labels [rows][label_bytes]  -128 (query side) / +8 (gallery side) per class; label_bytes = cmh_tc_operand_bytes(ncls)
 * Query operands have Qpad rows (rows >= Q are filled by cmh_tc_expand), gallery operands exactly N rows. */
typedef struct cmh_tc_operands {
    const int8_t* q_codes;
    const int8_t* q_labels; /* NULL for top-k */
    const int8_t* g_codes;
    const int8_t* g_labels; /* NULL for top-k */
    int32_t code_bytes, label_bytes;
} cmh_tc_operands;
int cmh_tc_operand_bytes(int n); /* 32, 64 or 128; <0 if n is outside 1..128 */
/* packed [n][nwords] -> out [rows][cmh_tc_operand_bytes(ncols)]; kind 0 = codes (either side), 1 = query labels, 2 = gallery labels;
 * rows >= n are padding (codes: the all -1 code, labels: zero). */
int cmh_tc_expand(const uint32_t* packed, int64_t n, int64_t rows, int nwords, int ncols, int kind, int8_t* out, void* stream);
/* == cmh_hist (with_labels = 0: relevance counts are 0) */
int cmh_tc_hist(const cmh_plan* plan, const cmh_tc_operands* ops, int with_labels, uint32_t* hist, void* stream);
/* == cmh_rank_topk */
int cmh_tc_rank_topk(const cmh_plan* plan, const cmh_tc_operands* ops, const uint32_t* within_all, const uint32_t* below_all,
                     const int32_t* thresh, int64_t k, int64_t idx_offset, uint64_t* keys, void* stream);
/* == cmh_rank_map for n_total < 2^24 (fp32 running ranks) */
int cmh_tc_rank_map(const cmh_plan* plan, const cmh_tc_operands* ops, const uint32_t* within_all, const uint32_t* within_rel,
                    const uint32_t* below_all, const uint32_t* below_rel, const int32_t* total, int64_t n_total,
                    double* ap_partial, int32_t* tindex, int64_t cap, void* stream);

/* Candidate path of the top-k (what retrieval.topk runs by default on large galleries; the two-pass path above is its exact
 * fallback).  Only ~k of N items can be in a query's top-k: with a per-query cutoff distance T >= (distance of the k-th
 * neighbour) ONE tensor-core pass appends the items with d <= T to per-(chunk, query) lists and skips everything else.
 *   cmh_tc_topk_cutoff   cutoff T[q] and index bound I[q] from the exact histogram (cmh_tc_hist on `sample_plan`) of a gallery
 *                        prefix, 5-sigma margin: candidates = { d < T } + { d == T, shard index <= I (rounded up to a 64-item tile) },
 *                        a prefix of the stable (distance, index) order
 *   cmh_tc_topk_collect  cand[c][q][cand_cap] = (distance << 24) | index inside chunk c, in gallery order; cand_count[c][q]
 *                        (0xFFFFFFFF = the list overflowed)
 *   cmh_tc_topk_count    totals[d][q] = #candidates at distance d ([bins][Qpad], what a rank exchanges); flags[0] |= 1 if a
 *                        list overflowed or a query has fewer than min(k, N) candidates -> caller must use the exact path
 *                        (k = 0: only overflow is flagged — sharded runs check the candidate count over all ranks instead)
 *   cmh_tc_topk_place    keys[q][rank] for this shard's candidates with global stable rank < k; totals_all = the all-gathered
 *                        totals, rank r's [bins][Qpad] block starting at r * rank_stride elements.
 *                        Fused exchange (replaces the all-reduce of the [Q][k] buffer): with multicast_keys != NULL every key is
 *                        stored once through that NVSwitch multicast address (multimem.st) and lands in the same slot of every
 *                        rank's buffer; else with npeers > 0, peer_keys = DEVICE array of npeers buffer addresses (one per rank,
 *                        peer-mapped) and the kernel stores the key into each of them over NVLink; else into `keys`.
 *                        The caller owns the synchronisation (a barrier over the ranks before and after).
 *                        status != NULL (sharded runs): the kernel also verifies the pass from the gathered data — status[0] |= 1
 *                        when a rank flagged a list overflow (row `bins` of its totals block) or the candidates of all ranks
 *                        together are fewer than min(k, gallery size) for a query; sample_all = the gathered sample blocks
 *                        (their header rows carry the shard sizes), same rank_stride = (bins + 1) * Qpad as the totals.
 *                        totals_all == NULL (one shard, world = 1; status required): the kernel sums its own per-chunk counts —
 *                        cmh_tc_topk_count is not needed — and status[0] |= 1 on a list overflow or fewer than min(k, N) candidates. */
int cmh_tc_topk_cutoff(const cmh_plan* sample_plan, const uint32_t* hist_sample, int64_t n_local, int64_t k, int32_t* cutoff,
                       int32_t* ibound, void* stream);
/* Sharded form of the cutoff: sample_sum = the all-gathered blocks of all ranks, uint32 [world][bins + 1][Qpad]; rank r's block =
 * its sample histogram totals [bins][Qpad] | row `bins`: [0] = its sample size, [1] = its shard size, [2 + r] = gallery index of its
 * first item (cmh_tc_topk_sample_block).  One GLOBAL cutoff and index bound per query, the bound translated into this rank's
 * shard: every rank keeps ~k/world candidates and their union is a prefix of the global (distance, index) order. */
/* This rank's contribution to that sum, built by ONE kernel from its sample histogram (hist_sample may be NULL for an empty
 * shard): out[d][q] = sum over the sample chunks, out[bins][0..] = the header described above, everything else 0. */
int cmh_tc_topk_sample_block(const cmh_plan* sample_plan, const uint32_t* hist_sample, int64_t Qpad, int bins, int64_t n_local,
                             int64_t idx_offset, int rank, int world, uint32_t* out, void* stream);
int cmh_tc_topk_cutoff_sharded(const cmh_plan* plan, const uint32_t* sample_sum, int64_t k, int rank, int world, int32_t* cutoff,
                               int32_t* ibound, void* stream);
/* chunk_begin / chunk_end: the launch covers chunks [chunk_begin, chunk_end) of the plan (chunk_end <= 0: all).  Lets a caller
 * whose gallery is still arriving from the host collect slab by slab while the next slab is in flight (calc_utils.hamming_topk). */
int cmh_tc_topk_collect(const cmh_plan* plan, const cmh_tc_operands* ops, const int32_t* cutoff, const int32_t* ibound,
                        int cand_cap, uint32_t* cand, uint32_t* cand_count, int chunk_begin, int chunk_end, void* stream);
int cmh_tc_topk_count(const cmh_plan* plan, int cand_cap, const uint32_t* cand, const uint32_t* cand_count, int64_t k,
                      uint32_t* totals, int32_t* flags, void* stream);
int cmh_tc_topk_place(const cmh_plan* plan, int cand_cap, const uint32_t* cand, const uint32_t* cand_count,
                      const uint32_t* totals_all, int64_t rank_stride, int world, int rank, int64_t k, int64_t idx_offset,
                      uint64_t* keys, const uint64_t* peer_keys, int npeers, uint64_t* multicast_keys, const uint32_t* sample_all,
                      int32_t* status, void* stream);

/* ---- X: the exchange step of the sharded top-k through NVSwitch multicast memory (NVLS) -------------------------------
 * `multicast_ptr` = the multicast address of a symmetric int64 buffer of `count` elements that every rank of the group has
 * mapped (torch.distributed._symmetric_memory: handle.multicast_ptr).  Rank `rank` reduces its 1/world slice with
 * multimem.ld_reduce(max.s64) — the switch reads the slot on every rank — and broadcasts the result with multimem.st, so after
 * a barrier every rank's buffer holds the element-wise maximum.  Replaces the NCCL all-reduce(MAX) of the [Q][k] key buffer
 * (unowned slots are -1).  The caller brackets the call with barriers over the ranks. */
int cmh_nvls_allreduce_max_s64(void* multicast_ptr, int64_t count, int rank, int world, void* stream);
/* The same exchange without a reduction: every slot has one owner, so a rank pushes the slots it owns (everything that is not -1 in
 * `local_keys`, its private placement buffer) into every rank's symmetric buffer with multicast stores.  The symmetric buffers
 * must hold -1 (EMPTY) or stale data that will be overwritten: every one of the `count` slots is written by its owner. */
int cmh_nvls_push_owned_s64(const void* local_keys, void* multicast_keys, int64_t count, void* stream);
/* Broadcast `bytes` (multiple of 16) from local `src` to the same offset of every rank's symmetric buffer (multicast_dst already
 * points at that offset): the all-gather of the small per-rank blocks (sample histograms, per-distance totals). */
int cmh_nvls_broadcast(const void* src, void* multicast_dst, int64_t bytes, void* stream);

/* ---- R5: merge of per-shard partial top-k after ONE all-gather -----------------------------------------
 * parts = [world][Q][k] sorted keys (0xFFFF...F = empty slot); out[q] = the k smallest keys. */
int cmh_topk_merge(const uint64_t* parts, int world, int64_t Q, int64_t k, uint64_t* out, void* stream);
/* keys -> (dist int32, index int64); either output may be NULL.  Empty slots give -1. */
int cmh_split_keys(const uint64_t* keys, int64_t count, int32_t* dist, int64_t* index, void* stream);

/* ---- one-shot single-GPU evaluator calls ---------------------------------------------------------------
 * cmh_map_k == calc_map_k(qB, rB, query_L, retrieval_L, k) (common/calc_utils.py:58-92) on packed inputs.
 * Outputs (each may be NULL except map_out): map_out fp64 scalar, ap[Q] fp64, tsum[Q], total[Q] int32,
 * tindex[Q*cap] int32.  workspace >= plan->workspace_bytes, 256-byte aligned. */
int cmh_map_k(const cmh_plan* plan, const uint32_t* qcodes, const uint32_t* qlabels, const uint32_t* gcodes,
              const uint32_t* glabels, int64_t k, void* workspace, size_t workspace_bytes, double* map_out,
              double* ap, int32_t* tsum, int32_t* total, int32_t* tindex, int64_t cap, void* stream);
/* cmh_topk == the first k columns of torch.sort(calc_hammingDist(qB, rB)) (calc_utils.py:76-77, stable). */
int cmh_topk(const cmh_plan* plan, const uint32_t* qcodes, const uint32_t* gcodes, int64_t k,
             int64_t idx_offset, void* workspace, size_t workspace_bytes, uint64_t* keys, void* stream);

/* ---- S: similarity helpers (fp32) -----------------------------------------------------------------------
 * cmh_label_sim   == calc_label_sim        (common/calc_utils.py:8-10)   out[i][j] = (a_i . b_j > 0)
 * cmh_cosine_sim  == cosine_similarity     (common/calc_utils.py:38-49)  no epsilon: zero row -> nan
 * cmh_euclid_sim  == euclidean_similarity  (common/calc_utils.py:28-36)  pairwise L2 distance        */
int cmh_label_sim_f32(const float* a, int64_t n, const float* b, int64_t m, int d, float* out, void* stream);
int cmh_cosine_sim_f32(const float* a, int64_t n, const float* b, int64_t m, int d, float* out, void* stream);
int cmh_euclid_sim_f32(const float* a, int64_t n, const float* b, int64_t m, int d, float* out, void* stream);

/* ---- E: encoder building blocks (models/CLIP/model.py) --------------------------------------------------------
 * cmh_gemm_bf16: out[M][N] = epilogue(A[M][K] . W[N][K]^T + bias[N]) on tcgen05 tensor cores (bf16 in, fp32
 * accumulate in TMEM).  A and W are bf16, K contiguous (W is exactly torch.nn.Linear.weight); lda/ldw/ldo/ldr in
 * elements.  Replaces the fp32 nn.Linear / in_proj / out_proj / c_fc / c_proj GEMMs of
 * ResidualAttentionBlock (models/CLIP/model.py:167-197) and conv1 (patch embed, :219,235). */
#define CMH_EPI_BF16 0       /* out bf16 = acc + bias                                   */
#define CMH_EPI_GELU_BF16 1  /* out bf16 = QuickGELU(acc + bias)   (model.py:162-164)   */
#define CMH_EPI_RESID_F32 2  /* out fp32 = resid + acc + bias      (residual stream)    */
#define CMH_EPI_F32 3        /* out fp32 = acc + bias                                   */
#define CMH_EPI_TANH_F32 4   /* out fp32 = tanh(acc + bias)        (DSPH head)          */
#define CMH_EPI_ERF_GELU_BF16 5 /* out bf16 = GELU_erf(acc + bias) (MITH residual MLPs, models/MITH/hash/hash.py:22) */
#define CMH_EPI_ERF_GELU_F32 6  /* out fp32 = GELU_erf(acc + bias) (split-precision token path of the MITH head)     */
int cmh_gemm_bf16(const void* A, int64_t M, int64_t K, int64_t lda, const void* W, int64_t N, int64_t ldw,
                  const float* bias, int epilogue, void* out, int64_t ldo, const float* resid, int64_t ldr,
                  void* stream);
/* (tuning / trace hooks of the GEMM are not part of the product ABI: include/cmh_debug.h) */


/* ---- E: CLIP encoders (models/CLIP/model.py) ---------------------------------------------------------------------
 * Weights are DEVICE pointers; matrices are bf16 in torch.nn.Linear layout ([out][in], `in` contiguous), vectors
 * (biases, LayerNorm gain/bias, embeddings) fp32.  The structs below are plain host structs filled by the caller
 * (the Python shim builds them from a reference `state_dict`; key names in the comments). */
typedef struct cmh_block_weights {      /* one ResidualAttentionBlock (model.py:167-197), prefix ...resblocks.<i>.  */
    const float* ln1_gain;  const float* ln1_bias;    /* ln_1.weight / ln_1.bias                                   */
    const void*  w_qkv;     const float* b_qkv;       /* attn.in_proj_weight [3D][D] / attn.in_proj_bias [3D]      */
    const void*  w_out;     const float* b_out;       /* attn.out_proj.weight [D][D] / .bias                       */
    const float* ln2_gain;  const float* ln2_bias;    /* ln_2.*                                                    */
    const void*  w_fc;      const float* b_fc;        /* mlp.c_fc.weight [4D][D] / .bias                           */
    const void*  w_proj;    const float* b_proj;      /* mlp.c_proj.weight [D][4D] / .bias                         */
} cmh_block_weights;

typedef struct cmh_tower {
    int32_t width, layers, heads, out_dim;            /* heads = width / 64 (model.py:300,465)                     */
    const cmh_block_weights* blocks;                  /* HOST array [layers]                                       */
    const float* ln_out_gain; const float* ln_out_bias; /* visual.ln_post.* | ln_final.*                           */
    const void*  w_out_proj;                          /* bf16 [out_dim][width] = visual.proj^T | text_projection^T */
    const float* pos_emb;                             /* visual.positional_embedding [grid^2+1][D] | positional_embedding [ctx][D] */
    /* image tower only (width of the text tower: leave zero/NULL) */
    int32_t patch, resolution;                        /* 32, 224                                                   */
    const void*  w_patch;                             /* bf16 [D][3*patch*patch] = visual.conv1.weight flattened   */
    const float* cls_emb;                             /* visual.class_embedding [D]                                */
    const float* ln_pre_gain; const float* ln_pre_bias; /* visual.ln_pre.*                                         */
    /* text tower only */
    int32_t vocab, context;                           /* 49408, 77                                                 */
    const float* tok_emb;                             /* token_embedding.weight [vocab][D] fp32                    */
    int64_t eot_id;                                   /* 49407 (model.py:384)                                      */
} cmh_tower;

/* Workspace bytes for `batch` samples of `seq_len` tokens (image tower: seq_len = grid^2 + 1). */
int64_t cmh_encoder_workspace_bytes(const cmh_tower* tower, int64_t batch, int32_t seq_len);

/* CLIP.encode_image (model.py:370 -> VisionTransformer.forward :232-268).
 * images [B][3][R][R] fp32 NCHW.  Outputs (fp32): cls_out [B][out_dim] (required);
 * tokens_out [B][L][out_dim] = ln_post + projection of ALL tokens (model.py:257-260; NULL = skip, the plain
 * encode_image only returns x[0]); attn_out [B][L-1] = last block's head-averaged CLS attention row without the CLS
 * column (model.py:265; NULL = skip).  The reference's seq_tokens [L-1][B][E] is tokens_out[:,1:].permute(1,0,2). */
int cmh_encode_image(const cmh_tower* tower, const float* images, int64_t batch, void* workspace, size_t workspace_bytes,
                     float* cls_out, float* tokens_out, float* attn_out, void* stream);
/* Same, from uint8 pixels [B][3][R][R] (what the dataloader holds after Resize / CenterCrop): ToTensor's /255 and
 * Normalize(mean, std) of dataset/transformer_dataset.py:41-45 run on the GPU, fused into the patch gather — a quarter of the
 * fp32 bytes cross PCIe and HBM.  mean_host / std_host = 3 HOST floats each (per channel). */
int cmh_encode_image_u8(const cmh_tower* tower, const uint8_t* images, const float* mean_host, const float* std_host, int64_t batch,
                        void* workspace, size_t workspace_bytes, float* cls_out, float* tokens_out, float* attn_out, void* stream);

/* CLIP.encode_text (model.py:373-396).  text [B][L] int64 token ids; key_padding_mask [B][L] uint8 (1 = pad) or NULL.
 * Outputs: eos_out [B][out_dim] fp32 (required); tokens_out [B][L][out_dim] (NULL = skip);
 * attn_out [B][L] = last block's head-averaged attention row of the EOS query with its own column zeroed
 * (model.py:381-382); new_mask_out [B][L] uint8 = key_padding_mask | (text == eot_id) (model.py:384).
 * L <= 128 and L <= tower->context. */
int cmh_encode_text(const cmh_tower* tower, const int64_t* text, const uint8_t* key_padding_mask, int64_t batch,
                    int32_t seq_len, void* workspace, size_t workspace_bytes, float* eos_out, float* tokens_out,
                    float* attn_out, uint8_t* new_mask_out, void* stream);

/* Building blocks, exported for the per-kernel parity tests:
 * cmh_layernorm: out[r] = LayerNorm(x[r]) (fp32 statistics, model.py:153-159); out bf16 or fp32.
 * cmh_attention_bf16: softmax(q.k^T/8 + masks).v per (sample, head); qkv bf16 [B*L][3*heads*64] (q|k|v), out bf16
 * [B*L][heads*64]; causal = build_attention_mask (model.py:358-364). */
int cmh_layernorm(const float* x, int64_t rows, int width, const float* gain, const float* bias, float eps, void* out,
                  int out_f32, void* stream);
int cmh_attention_bf16(const void* qkv, int64_t batch, int seq_len, int heads, const uint8_t* key_padding_mask, int causal,
                       void* out, void* stream);

/* ---- H: hash heads (fp32) -----------------------------------------------------------------------------------------
 * cmh_linear_f32: out[r][n] = act((x[r].W[n] + bias[n]) * scale[n] + shift[n]); W [out_dim][in_dim] fp32; bias, scale,
 * shift may be NULL (scale and shift together). */
#define CMH_ACT_NONE 0
#define CMH_ACT_TANH 1
#define CMH_ACT_RELU 2
int cmh_linear_f32(const float* x, int64_t rows, int in_dim, const float* weight, const float* bias, int out_dim,
                   const float* scale, const float* shift, int act, float* out, int64_t ldo, void* stream);

/* DSPH head, eval mode (models/DSPH/hash/hash.py:6-15): hash [rows][nbits] = tanh(feat.W^T + b);
 * packed (may be NULL) = bit-packed sign (runners/base.py:407-410: +1 iff hash > 0). */
int cmh_head_dsph(const float* feat, int64_t rows, int in_dim, const float* weight, const float* bias, int nbits, float* hash,
                  uint32_t* packed, void* stream);

/* DCMHT head, eval mode (models/DCMHT/hash/hash.py:35-46), one modality. */
typedef struct cmh_dcmht_head {
    const float* w_v;   const float* b_v;      /* atten.in_proj_weight[2D:3D] [D][D] / atten.in_proj_bias[2D:3D]       */
    const float* w_out; const float* b_out;    /* atten.out_proj.*                                                     */
    const float* bn_scale; const float* bn_shift; /* image: norm.weight/sqrt(running_var+eps), norm.bias - running_mean*scale; text: NULL */
    const float* norm_gain; const float* norm_bias; /* text: LayerNorm norm.weight / norm.bias; image: NULL            */
    float eps;                                  /* 1e-5                                                                 */
    const float* w_fc2; const float* b_fc2;    /* fc2.* [2*nbits][D]                                                   */
} cmh_dcmht_head;
/* scratch: fp32 [rows*(2*in_dim + 2*nbits)].  probs [rows][2*nbits] (may be NULL) = pairwise-softmax output of the
 * head; packed (may be NULL): bit j = 1 iff probs[2j+1] > probs[2j] (DCMHTTrainer.make_hash_code,
 * runners/DCMHT/runner.py:83-95). */
int cmh_head_dcmht(const float* feat, int64_t rows, int in_dim, const cmh_dcmht_head* head, int nbits, float* scratch,
                   float* probs, uint32_t* packed, void* stream);

/* ---- L: objective value ---------------------------------------------------------------------------------------------
 * cmh_hyp_loss_f32 == HyP.forward(x, y, label) (models/DSPH/loss/HyP.py:18-69), forward value only (evaluation /
 * monitoring; no gradient).  x, y [B][nbits] fp32 (image / text hash outputs), labels_packed [B][LW] from
 * cmh_pack_labels, proxies [ncls][nbits] fp32 (HyP.proxies), threshold / alpha as in the reference constructor.
 * workspace >= 256 + (2B + ncls) * nbits * 4 bytes, 256-byte aligned.  loss_out: one fp32 on the device. */
int cmh_hyp_loss_f32(const float* x, const float* y, const uint32_t* labels_packed, const float* proxies, int64_t batch,
                     int nbits, int ncls, float threshold, float alpha, void* workspace, size_t workspace_bytes,
                     float* loss_out, void* stream);

/* MITH head, eval mode (models/MITH/hash/hash.py:193-254), one modality (HashLayer.encode_img / encode_txt).
 * Matrices of the residual MLPs, of the concept transformer and the concept projection are bf16 ([out][in]); the concept
 * embedding, the per-bit hashing weights and every vector are fp32. */
#define CMH_MITH_MAX_MLP_LAYERS 4
typedef struct cmh_mith_head {
    int32_t dim, nbits, mlp_layers, top_k;           /* 512, k_bits, res_mlp_layers, top_k_label                           */
    const float* ln_gain[CMH_MITH_MAX_MLP_LAYERS];   /* gcl_*.mlp.lns.<i>.weight                                            */
    const float* ln_bias[CMH_MITH_MAX_MLP_LAYERS];
    const void*  w1[CMH_MITH_MAX_MLP_LAYERS];        /* gcl_*.mlp.mlps.<i>.0.weight [4D][D] bf16                            */
    const float* b1[CMH_MITH_MAX_MLP_LAYERS];
    const void*  w2[CMH_MITH_MAX_MLP_LAYERS];        /* gcl_*.mlp.mlps.<i>.3.weight [D][4D] bf16                            */
    const float* b2[CMH_MITH_MAX_MLP_LAYERS];
    /* Optional split-precision copies for the TOKEN path (the one that feeds the discontinuous top-k concept selection):
     * w = hi + lo with hi = bf16(w), lo = bf16(w - hi); w1_split = [hi | lo | hi] along K (bf16 [4D][3D]), w2_split likewise
     * (bf16 [D][12D]).  With A split the same way ([hi | hi | lo]) one K-concatenated GEMM accumulates hi.hi + hi.lo + lo.hi:
     * ~2^-16 relative instead of 2^-9.  NULL = plain bf16 (3x fewer FLOPs on the token MLPs). */
    const void*  w1_split[CMH_MITH_MAX_MLP_LAYERS];
    const void*  w2_split[CMH_MITH_MAX_MLP_LAYERS];
    const float* w_concept;                          /* gcl_*.common_concept_embedding.weight [K][D] fp32 (no bias)         */
    const float* pos;                                /* lct_*.position.pe[:, 0, :] [K][D] fp32                              */
    cmh_tower    transformer;                        /* lct_*.transformer: width, layers, heads, blocks (other fields unused) */
    const float* w_bits; const float* b_bits;        /* lct_*.hashing.fc_list.<k>.weight stacked [K][D] / bias [K]          */
    const void*  w_cproj; const float* b_cproj;      /* {img,txt}_concept_proj.weight [D][D] bf16 / bias                    */
} cmh_mith_head;

int64_t cmh_head_mith_workspace_bytes(const cmh_mith_head* head, int64_t batch, int32_t tokens);

/* cls [B][D] fp32; the L tokens of sample b are rows b*tokens_per_sample + first_token .. +L-1 of `tokens` ([.][D] fp32:
 * exactly the tokens_out buffer of cmh_encode_image (first_token = 1, L = 49) / cmh_encode_text (first_token = 0));
 * key_padding_mask [B][L] uint8 or NULL (text: the new_mask_out of cmh_encode_text).
 * Outputs fp32: res_cls [B][D] = normalize(mlp(cls)) (may be NULL), cls_hash [B][K], tokens_hash [B][K] (required),
 * trans_tokens [B][K][D] = normalize(concept_proj(transformer tokens)) (may be NULL; the reference returns it as
 * [K][B][D]); packed [B][W] (may be NULL) = bit-packed sign(cls_hash + tokens_hash)
 * (MITHTrainer.generate_hash, runners/MITH/runner.py:125-131). */
int cmh_head_mith(const cmh_mith_head* head, const float* cls, const float* tokens, int32_t tokens_per_sample,
                  int32_t first_token, int32_t L, const uint8_t* key_padding_mask, int64_t batch, void* workspace,
                  size_t workspace_bytes, float* res_cls, float* cls_hash, float* tokens_hash, float* trans_tokens,
                  uint32_t* packed, void* stream);

/* ---- TR: the tail of the DSPH training step (BASELINE config C5; runners/DSPH/runner.py:104-127) -------------------------
 * What a step needs besides the encoder's forward/backward: the objective's gradient, the hash head's backward and the
 * optimiser.  The backward pass through the CLIP towers is not part of this library.
 *
 * cmh_hyp_loss_grad_f32: HyP.forward + what loss.backward() leaves in x.grad, y.grad and hyp.proxies.grad
 *   (models/DSPH/loss/HyP.py:18-69).  Arguments as cmh_hyp_loss_f32 plus dx, dy [B][nbits] and dproxies [ncls][nbits] (fp32, written).
 * cmh_linear_tanh_backward_f32: backward of y = tanh(feat . W^T + b) (models/DSPH/hash/hash.py:6-15, dropout off):
 *   dW [nbits][in_dim], db [nbits] (may be NULL), dfeat [rows][in_dim] (may be NULL); dz_scratch = fp32 [rows][nbits]. */
int cmh_hyp_loss_grad_f32(const float* x, const float* y, const uint32_t* labels_packed, const float* proxies, int64_t batch, int nbits,
                          int ncls, float threshold, float alpha, void* workspace, size_t workspace_bytes, float* loss_out, float* dx,
                          float* dy, float* dproxies, void* stream);
int cmh_linear_tanh_backward_f32(const float* feat, const float* y, const float* dy, const float* W, int64_t rows, int in_dim, int nbits,
                                 float* dz_scratch, float* dW, float* db, float* dfeat, void* stream);

/* Fused multi-tensor optimiser steps.  One launch updates EVERY tensor (the reference loops over tensors in Python,
 * models/common/optimizer.py:118-163: ~300 tensors x ~10 small kernels per step).  The caller keeps a DEVICE array of
 * cmh_opt_tensor (lr = the scheduled learning rate of the tensor's group for this step) and a block table: block b works on
 * elements [block_chunk[b] * cmh_opt_chunk_elems(), ...) of tensor block_tensor[b].
 * cmh_bert_adam_step == BertAdam.step (:130-165): per-tensor clip_grad_norm_ (in place), next_m / next_v, update = m / (sqrt(v) + e)
 *   + weight_decay * p, p -= lr * update; no bias correction.  sumsq_dev = fp32 [ntensors] scratch.
 * cmh_sgd_momentum_step == torch.optim.SGD(momentum, weight_decay) as built for the HyP proxies (runners/DSPH/runner.py:86-89);
 *   `m` is the momentum buffer, first_step != 0 initialises it with the gradient like torch does. */
typedef struct cmh_opt_tensor {
    float* param;
    float* grad;
    float* m;
    float* v;
    int64_t n;
    float lr;
    float weight_decay;
} cmh_opt_tensor;
int cmh_opt_chunk_elems(void);
int cmh_bert_adam_step(const cmh_opt_tensor* tensors_dev, int ntensors, const int32_t* block_tensor_dev, const int32_t* block_chunk_dev,
                       int nblocks, float* sumsq_dev, float b1, float b2, float e, float max_grad_norm, void* stream);
int cmh_sgd_momentum_step(const cmh_opt_tensor* tensors_dev, int ntensors, const int32_t* block_tensor_dev, const int32_t* block_chunk_dev,
                          int nblocks, float momentum, int first_step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CMH_H_ */
