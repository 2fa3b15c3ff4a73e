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

#define CMH_ABI_VERSION 1

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
int cmh_gemm_bf16(const void* A, int64_t M, int64_t K, int64_t lda, const void* W, int64_t N, int64_t ldw,
                  const float* bias, int epilogue, void* out, int64_t ldo, const float* resid, int64_t ldr,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CMH_H_ */
