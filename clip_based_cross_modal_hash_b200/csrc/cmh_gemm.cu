// cmh_gemm.cu — bf16 x bf16 -> fp32 GEMM for the CLIP encoder blocks on 5th-generation tensor cores.
//
//   out[M][N] = epilogue( A[M][K] . W[N][K]^T + bias[N] )         (torch.nn.Linear layout: both operands K-major)
//
// replaces the fp32 cuBLAS SGEMMs behind nn.Linear / nn.MultiheadAttention / conv1 of models/CLIP/model.py:167-268.
//
// Persistent, warp-specialised, one CTA per SM.  CG = 2 pairs the two SMs of a TPC on one 256 x BN tile
// (tcgen05.mma.cta_group::2): each CTA stages its own 128 rows of A and HALF of the W tile (fewer bytes per MAC from L2
// and half the shared-memory operand reads per SM).
//   warps 0..7  epilogue       tcgen05.ld 32x32b.x32 -> registers -> bias (prefetched, one column per lane) / QuickGELU /
//                              erf-GELU / tanh -> swizzled shared staging -> coalesced 16-byte global stores, or
//                              red.global.add.v4.f32 into the fp32 residual stream
//   warp 8      TMA producer   cp.async.bulk.tensor.2d (128B swizzle) of a 128 x 64 A tile and a (BN/CG) x 64 W tile into a
//                              5..8-stage shared-memory ring, completion on mbarriers (expect_tx); L2 prefetch 10 k-blocks ahead
//   warp 9      MMA issuer     warp-converged, elect.sync-predicated tcgen05.mma.kind::f16 (M=128*CG, N=BN, K=16) x 4 per
//                              stage, at most 2 stages queued; tcgen05.commit releases the stage / publishes the accumulator
//   warp 10     TMEM owner     tcgen05.alloc of 2 x BN fp32 accumulator columns (double buffered) / dealloc
// The epilogue of tile i overlaps the MMAs of tile i+1 through the two TMEM accumulator stages; the tiles of the last,
// partial wave are cut into column slices; every launch is a programmatic dependent launch (prologue overlaps the
// previous kernel's tail).
#include <cuda.h>
#include <cuda_bf16.h>

#include "cmh_common.cuh"
#include "cmh_debug.h"
#include "cmh_encoder.h"
#include "cmh_tcgen05.cuh"

namespace cmh {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;        // two per TMEM lane quarter: they hide each other's TMEM-load latency
constexpr int L2_PREFETCH_KB = 10;  // k-blocks between the L2 prefetch of a tile and its TMA load (ring depth + 4)
constexpr int MMA_LOOKAHEAD = 2;  // k-blocks (of 4 MMAs) in flight in the tensor pipe
constexpr int PRODUCER_WARP = EPI_WARPS, MMA_WARP = EPI_WARPS + 1, TMEM_WARP = EPI_WARPS + 2;
// The latency-critical single-thread roles get the HIGHEST warp ids: the SM's warp arbiter prefers higher ids, and with
// the epilogue warps above them the MMA issue loop lost ~20 % once an epilogue ran concurrently (profiles/README.md).
constexpr int GEMM_THREADS = (EPI_WARPS + 3) * 32;
constexpr int STG_CHUNK_BYTES = 4096;  // one epilogue chunk: 32 rows x 128 B (64 bf16 or 32 fp32 columns), XOR-swizzled
// One staging buffer per warp.  The epilogue leaves through ordinary coalesced 16-byte stores / vector reductions, not TMA:
// a TMA store queues behind the operand loads in the SM's TMA unit, and with a deep operand ring its shared-memory read
// took ~2000 clk, which made the epilogue 3x slower than the main loop (profiles/README.md).
constexpr int STG_WARP_BYTES = STG_CHUNK_BYTES;
constexpr int SMEM_BUDGET = 227 * 1024 - EPI_WARPS * STG_WARP_BYTES - 1024 - 256;  // operand ring

template <int BN, int CG>
struct GemmCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = (BN / CG) * BK * 2;  // per CTA: with cta_group::2 each CTA holds half of the W tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES < 8 ? SMEM_BUDGET / STAGE_BYTES : 8;
    static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
    static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
    static constexpr int SMEM_BYTES = RING_BYTES + EPI_WARPS * STG_WARP_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

struct GemmParams {
    int64_t M, N, K;
    const float* bias;   // [N] or null
    void* out;           // bf16 or fp32 [M][ldo]
    int64_t ldo;
    const float* resid;  // fp32 [M][ldr] (EPI_RESID_F32) or null
    int64_t ldr;
    int epi;
    int tiles_m, tiles_n;
    // work items: [0, tail_start) are full (CG*128) x BN tiles; the tiles of the last, partial wave are cut into
    // 2^tail_shift column slices each (one item per slice) so that the wave ends after a slice, not after a tile
    int tail_start, tail_shift, num_items;
    int lookahead;  // k-blocks the MMA warp may have queued in the tensor pipe
    long long* trace;  // debug timeline (cmh_gemm_set_trace): [cta][64] SM clock stamps, or null
};

#define CMH_TRACE(slot)                                                        \
    do {                                                                       \
        if (p.trace && (slot) < 64) p.trace[size_t(blockIdx.x) * 64 + (slot)] = clock64(); \
    } while (0)

// x * sigmoid(1.702 x)   models/CLIP/model.py:162-164, as 0.5x + 0.5x * tanh(0.851 x): ONE MUFU op per element
// (tanh.approx.f32, relative error 2^-11 on the tanh => absolute error <= 2.5e-4 |x|, below the bf16 rounding of the
// result).  The epilogue owns 128 x 256 elements per tile and the SM has 16 MUFU lanes: an exp2 + reciprocal sigmoid
// (2 MUFU) alone costs 4096 clk of the 6144 clk a tile's MMAs take, and an IEEE division made the epilogue 3x slower
// than the main loop (profiles/README.md).
__device__ __forceinline__ float quick_gelu(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
    const float h = 0.5f * x;
    return fmaf(h, t, h);
}
template <int N>
__device__ __forceinline__ float pick_bias(const float (&b)[N], int i) {  // register array, runtime index
    float v = b[0];
#pragma unroll
    for (int k = 1; k < N; ++k) v = i == k ? b[k] : v;
    return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <int BN, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmB2, const GemmParams p) {
    using C = GemmCfg<BN, CG>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stg_all = smem + C::RING_BYTES;  // EPI_WARPS x 4 KB, 1024-aligned
    uint64_t* full = reinterpret_cast<uint64_t*>(stg_all + EPI_WARPS * STG_WARP_BYTES);
    uint64_t* empty = full + C::STAGES;
    uint64_t* tfull = empty + C::STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 1 ? 0u : cluster_ctarank();  // CTA inside the pair; rank 0 issues the MMAs
    if (threadIdx.x == MMA_WARP * 32) CMH_TRACE(0);
    const int unit = CG == 1 ? int(blockIdx.x) : int(blockIdx.x >> 1);
    const int units = CG == 1 ? int(gridDim.x) : int(gridDim.x >> 1);
    const int num_items = p.num_items;
    const int kblocks = int((p.K + BK - 1) / BK);
    // item -> (row block, first column, width)
    auto decode = [&](int j, int& m_idx, int& n_first, int& width) {
        int tile = j, sub = 0;
        width = BN;
        if (j >= p.tail_start) {
            const int r = j - p.tail_start;
            tile = p.tail_start + (r >> p.tail_shift);
            sub = r & ((1 << p.tail_shift) - 1);
            width = BN >> p.tail_shift;
        }
        // column tiles fastest: the CTAs that run together share a few row blocks of A (the big operand: activations)
        // and all of W, so A is read from HBM once — with row blocks fastest the K=3072 GEMM re-read the 79 MB
        // activation matrix once per column tile and ran at HBM speed (profiles/README.md)
        m_idx = tile / p.tiles_n;
        n_first = (tile % p.tiles_n) * BN + sub * width;
    };

    if (warp == PRODUCER_WARP && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        prefetch_tmap(&tmB2);
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full[s], 1);   // the leader's arrive.expect_tx; bytes come from both CTAs' TMA loads
            mbar_init(&empty[s], 1);  // tcgen05.commit (multicast to both CTAs when CG = 2)
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], EPI_WARPS * CG);  // one arrival per epilogue warp of every CTA of the pair
        }
        mbar_fence_init();
    }
    if (warp == TMEM_WARP) tmem_alloc<CG>(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncwarp();
    if (CG == 1) __syncthreads(); else cluster_sync_all();  // the peer's barriers must exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == MMA_WARP * 32) CMH_TRACE(1);
    pdl_wait();               // everything above overlapped the previous kernel's tail; its outputs are visible from here
    pdl_launch_dependents();  // the next kernel's blocks may take SMs as this grid's CTAs retire

    if (warp == PRODUCER_WARP) {
        int stage = 0;
        uint32_t phase = 0;
        for (int t = unit; t < num_items; t += units) {
            int m_idx, n_first, width;
            decode(t, m_idx, n_first, width);
            const int m0 = m_idx * (BM * CG) + int(rank) * BM;
            const int n0 = n_first + int(rank) * (width / CG);
            const CUtensorMap* mapB = width == BN ? &tmB : &tmB2;  // box = width/CG rows of W
            const uint32_t stage_tx = uint32_t(C::A_BYTES + (width / CG) * BK * 2);
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&empty[stage], phase ^ 1u);
                if (elect_one()) {
                    uint8_t* sa = smem + stage * C::STAGE_BYTES;
                    if (kb + L2_PREFETCH_KB < kblocks) {  // hide the HBM part of the load latency behind the ring
                        tma_prefetch_l2_2d(&tmA, (kb + L2_PREFETCH_KB) * BK, m0);
                        tma_prefetch_l2_2d(mapB, (kb + L2_PREFETCH_KB) * BK, n0);
                    }
                    if (CG == 1) {
                        mbar_arrive_expect_tx(&full[stage], stage_tx);
                        tma_load_2d(sa, &tmA, kb * BK, m0, &full[stage]);
                        tma_load_2d(sa + C::A_BYTES, mapB, kb * BK, n0, &full[stage]);
                    } else {
                        if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * stage_tx);
                        tma_load_2d_pair(sa, &tmA, kb * BK, m0, &full[stage]);
                        tma_load_2d_pair(sa + C::A_BYTES, mapB, kb * BK, n0, &full[stage]);
                    }
                }
                __syncwarp();
                if (++stage == C::STAGES) stage = 0, phase ^= 1u;
            }
        }
    } else if (warp == MMA_WARP) {
        if (rank == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0, issued = 0;
            for (int t = unit; t < num_items; t += units, ++it) {
                const uint32_t idesc = make_idesc_bf16(BM * CG, t >= p.tail_start ? (BN >> p.tail_shift) : BN);
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                if (lane == 0) CMH_TRACE(2 + it * 4);
                mbar_wait(&tempty[as], aphase ^ 1u);  // every epilogue warp (of both CTAs) has drained this stage
                tc_fence_after();
                if (lane == 0) CMH_TRACE(3 + it * 4);
                const uint32_t d_tmem = tmem_base + uint32_t(as * BN);
                for (int kb = 0; kb < kblocks; ++kb, ++issued) {
                    // Keep at most MMA_LOOKAHEAD k-blocks queued in the tensor pipe.  tcgen05.ld of the epilogue warps is
                    // served in order behind the queued MMAs: with the whole 6-stage ring issued ahead, every TMEM load
                    // waited ~3000 clk and the epilogue fell behind the main loop (profiles/README.md).
                    if (issued >= p.lookahead) {
                        const int g = issued - p.lookahead;
                        mbar_wait(&empty[g % C::STAGES], uint32_t(g / C::STAGES) & 1u);
                    }
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (kb == 0 && lane == 0) CMH_TRACE(4 + it * 4);
                    const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
                    const uint64_t adesc = make_desc_sw128(sa), bdesc = make_desc_sw128(sa + C::A_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            // advancing 16 bf16 = 32 B inside the swizzle row: +2 in 16-byte units
                            umma_f16<CG>(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (kb | k) != 0);
                        }
                        umma_commit<CG>(&empty[stage]);  // stage reusable once these MMAs have read it
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) stage = 0, phase ^= 1u;
                }
                if (elect_one()) umma_commit<CG>(&tfull[as]);  // accumulator complete
                __syncwarp();
                if (lane == 0) CMH_TRACE(5 + it * 4);
            }
        }
    } else if (warp < EPI_WARPS) {
        // Epilogue: warp w reads TMEM lanes 32*(w%4).. (its lane quarter = 32 accumulator rows); the two warps of a
        // quarter take alternate 128-byte column chunks (64 bf16 / 32 fp32 columns).  A chunk goes TMEM -> registers ->
        // (+bias, activation) -> swizzled shared staging -> one TMA tile store (or fp32 reduce-add into the residual
        // stream); staging is double buffered per warp, stores are asynchronous and fully coalesced.
        const int e = warp, q = warp & 3, half = e >> 2;
        const bool out_bf16 = p.epi == CMH_EPI_BF16 || p.epi == CMH_EPI_GELU_BF16 || p.epi == CMH_EPI_ERF_GELU_BF16;
        const int cw = out_bf16 ? 64 : 32;  // columns per chunk
        const uint32_t buf = smem_u32(stg_all + e * STG_WARP_BYTES);
        const uint32_t rowp = buf + lane * 128;
        const int sw = lane & 7;  // 16-byte chunk index ^= row & 7: conflict-free row writes and row reads
        const int esz = out_bf16 ? 2 : 4;
        int it = 0;
        for (int t = unit; t < num_items; t += units, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            int m_idx, n0, width;
            decode(t, m_idx, n0, width);
            const int m0 = m_idx * (BM * CG) + int(rank) * BM;
            // Bias of this warp's 32-column groups, one column per lane, fetched BEFORE the accumulator is ready: while the
            // operand ring is in flight a global load from this SM queues behind ~190 KB of TMA traffic (~3000 clk), and
            // loading the bias inside the chunk loop made the epilogue 4x slower than the main loop (profiles/README.md).
            constexpr int NB = 2 * ((BN + 127) / 128);  // 32-column groups one warp can own (alternate 128-byte chunks)
            float breg[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const int gcol = n0 + (out_bf16 ? (half + 2 * (i >> 1)) * 64 + (i & 1) * 32 : (half + 2 * i) * 32) + lane;
                breg[i] = (p.bias && gcol < p.N) ? __ldg(p.bias + gcol) : 0.f;
            }
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            if (e == 0 && lane == 0) CMH_TRACE(34 + it * 2);
            int grp = 0;  // index of the current 32-column group in breg
#pragma unroll 1
            for (int c0 = half * cw; c0 < width; c0 += 2 * cw) {
                if (n0 + c0 >= p.N) break;  // remaining chunks lie outside the matrix (uniform)
                const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN + c0);
                uint32_t r[32];
                tmem_ld32(taddr, r);
                if (out_bf16) {
                    uint32_t pk[32];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        if (hh == 1) tmem_ld32(taddr + 32, r);
                        float v[32];
                        const float bl = pick_bias<NB>(breg, grp++);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bl, j);
                        if (p.epi == CMH_EPI_GELU_BF16) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
                        } else if (p.epi == CMH_EPI_ERF_GELU_BF16) {  // nn.GELU() of the MITH residual MLPs (hash.py:22)
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = 0.5f * v[j] * (1.0f + erff(v[j] * 0.70710678118654752f));
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) pk[hh * 16 + j] = pack_bf16(v[2 * j], v[2 * j + 1]);
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c) sts128(rowp + ((c ^ sw) << 4), pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                } else {
                    float v[32];
                    const float bl = pick_bias<NB>(breg, grp++);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bl, j);
                    if (p.epi == CMH_EPI_TANH_F32) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
                    } else if (p.epi == CMH_EPI_ERF_GELU_F32) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.5f * v[j] * (1.0f + erff(v[j] * 0.70710678118654752f));
                    }
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        sts128(rowp + ((c ^ sw) << 4), __float_as_uint(v[4 * c]), __float_as_uint(v[4 * c + 1]),
                               __float_as_uint(v[4 * c + 2]), __float_as_uint(v[4 * c + 3]));
                }
                __syncwarp();
                // staging -> global: 8 lanes cover one 128-byte row, a warp instruction writes 4 full rows
                {
                    const int c16 = lane & 7;
                    const int col = n0 + c0 + c16 * (16 / esz);
                    const bool col_ok = col < p.N;  // N is a multiple of the 16-byte vector (checked on the host)
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = i * 4 + (lane >> 3);
                        const int64_t grow = int64_t(m0) + q * 32 + row;
                        const uint4 v = lds128(buf + row * 128 + ((c16 ^ (row & 7)) << 4));
                        if (col_ok && grow < p.M) {
                            uint8_t* dst = static_cast<uint8_t*>(p.out) + (grow * p.ldo + col) * esz;
                            if (p.epi == CMH_EPI_RESID_F32) red_add_f32x4(dst, v);
                            else stg128(dst, v);
                        }
                    }
                }
                __syncwarp();  // the buffer is free again
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 1) mbar_arrive(&tempty[as]);
                else mbar_arrive_remote(&tempty[as], 0);  // the leader's MMA thread waits for both CTAs
                if (e == 0) CMH_TRACE(35 + it * 2);
            }
        }
    }
    tc_fence_before();
    __syncwarp();
    if (CG == 1) __syncthreads(); else cluster_sync_all();  // the peer may still be reading this CTA's shared memory / TMEM
    if (warp == TMEM_WARP) {
        tc_fence_after();
        tmem_dealloc<CG>(tmem_base, C::TMEM_COLS);
    }
    if (threadIdx.x == MMA_WARP * 32) CMH_TRACE(63);
}

long long* g_trace = nullptr;
int g_force_bn = 0, g_force_cg = 0, g_force_units = 0, g_lookahead = MMA_LOOKAHEAD;
bool g_tail_slicing = true;  // test/bench hook (cmh_gemm_force_tile): 0 = automatic

// ---- host side ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

// 2-D bf16 tensor map: rows x cols (cols contiguous), box = box_rows x 64 columns, 128-byte swizzle
int make_tmap(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(CMH_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
    cuuint64_t strides[1] = {cuuint64_t(ld) * 2};
    cuuint32_t box[2] = {cuuint32_t(BK), cuuint32_t(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CMH_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
    return CMH_OK;
}

template <int BN, int CG>
int launch_gemm(const CUtensorMap& ta, const void* W, int64_t ldw, GemmParams p, cudaStream_t st) {
    using C = GemmCfg<BN, CG>;
    static PerDeviceOnce once;
    if (once.needs()) {
        CMH_CUDA_TRY(cudaFuncSetAttribute(gemm_bf16_kernel<BN, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    }
    p.tiles_m = int(ceil_div(p.M, BM * CG));
    p.tiles_n = int(ceil_div(p.N, BN));
    const int tiles = p.tiles_m * p.tiles_n;
    int units = sm_count_cached() / CG;
    if (g_force_units > 0 && g_force_units < units) units = g_force_units;
    // Tail slicing: the R tiles of the last partial wave become R * 2^shift column slices on as many units.  A slice is
    // at least one epilogue chunk wide (64 bf16 / 32 fp32 columns) and the slices must fit one wave.
    const bool out_bf16 = p.epi == CMH_EPI_BF16 || p.epi == CMH_EPI_GELU_BF16 || p.epi == CMH_EPI_ERF_GELU_BF16;
    const int rem = tiles % units;
    int shift = 0;
    if (g_tail_slicing && tiles > units && rem > 0) {
        const int min_w = out_bf16 ? 64 : 32;
        while ((BN >> (shift + 1)) >= min_w && (BN >> (shift + 1)) % min_w == 0 && (BN >> (shift + 1)) % (16 * CG) == 0 &&
               (rem << (shift + 1)) <= units)
            ++shift;
    }
    p.tail_shift = shift;
    p.tail_start = shift ? tiles - rem : tiles;
    p.num_items = p.tail_start + ((tiles - p.tail_start) << shift);
    CUtensorMap tb, tb2;
    if (int rc = make_tmap(&tb, W, p.N, p.K, ldw, BN / CG)) return rc;
    if (int rc = make_tmap(&tb2, W, p.N, p.K, ldw, (BN >> shift) / CG)) return rc;
    const int grid_units = p.num_items < units ? p.num_items : units;
    CMH_CUDA_TRY(launch_kernel(gemm_bf16_kernel<BN, CG>, dim3(unsigned(grid_units * CG)), dim3(GEMM_THREADS), C::SMEM_BYTES, st, CG,
                               ta, tb, tb2, p));
    return CMH_OK;
}

// Tile shape: fewest "waves x bytes one SM pulls from L2 per k-block" units.  The kernels are bound by L2->SM
// bandwidth before the tensor pipe (profiles/README.md), so a tile costs (128 + BN/CG) rows of 128 bytes per k-block.
void pick_tile(int64_t M, int64_t N, int* bn_out, int* cg_out) {
    const int sms = sm_count_cached();
    double best_cost = 1e30;
    *bn_out = 128, *cg_out = 1;
    for (int cg : {2, 1}) {
        if (cg == 2 && M <= BM) continue;  // a pair would leave the second CTA without rows
        for (int bn : {256, 192, 128}) {
            if (N < bn && bn != 128) continue;
            const int64_t tiles = ceil_div(M, BM * cg) * ceil_div(N, bn);
            const double cost = double(ceil_div(tiles, sms / cg)) * (128 + bn / cg);
            if (cost < best_cost - 1e-9) best_cost = cost, *bn_out = bn, *cg_out = cg;
        }
    }
}

}  // namespace

int gemm_bf16(const void* A, int64_t M, int64_t K, int64_t lda, const void* W, int64_t N, int64_t ldw,
              const float* bias, int epi, void* out, int64_t ldo, const float* resid, int64_t ldr, cudaStream_t st) {
    CMH_REQUIRE(A && W && out && M > 0 && N > 0 && K > 0, "gemm: bad arguments");
    CMH_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, "gemm: leading dimensions must be multiples of 8 elements");
    CMH_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "gemm: operands must be 16-byte aligned");
    CMH_REQUIRE(epi >= CMH_EPI_BF16 && epi <= CMH_EPI_ERF_GELU_F32, "gemm: unknown epilogue %d", epi);
    CMH_REQUIRE(N % ((epi == CMH_EPI_BF16 || epi == CMH_EPI_GELU_BF16 || epi == CMH_EPI_ERF_GELU_BF16) ? 8 : 4) == 0,
                "gemm: N = %lld must be a multiple of the 16-byte output vector", (long long)N);
    CMH_REQUIRE(epi != CMH_EPI_RESID_F32 || resid, "gemm: residual epilogue needs resid");
    const bool out_bf16 = epi == CMH_EPI_BF16 || epi == CMH_EPI_GELU_BF16 || epi == CMH_EPI_ERF_GELU_BF16;
    CMH_REQUIRE(ldo % (out_bf16 ? 8 : 4) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "gemm: output must be 16-byte aligned per row");
    CMH_REQUIRE(epi != CMH_EPI_RESID_F32 || (resid == out && ldr == ldo),
                "gemm: the residual epilogue works in place (resid == out): it is a vector red.add into the stream");
    int bn = 128, cg = 1;
    pick_tile(M, N, &bn, &cg);
    if (g_force_bn > 0) bn = g_force_bn;
    if (g_force_cg > 0) cg = g_force_cg;
    CUtensorMap ta;
    if (int rc = make_tmap(&ta, A, M, K, lda, BM)) return rc;
    GemmParams p{};
    p.M = M, p.N = N, p.K = K, p.bias = bias, p.out = out, p.ldo = ldo, p.resid = resid, p.ldr = ldr, p.epi = epi;
    p.trace = g_trace;
    p.lookahead = g_lookahead;
    if (cg == 2) {
        switch (bn) {
            case 256: return launch_gemm<256, 2>(ta, W, ldw, p, st);
            case 192: return launch_gemm<192, 2>(ta, W, ldw, p, st);
            default: return launch_gemm<128, 2>(ta, W, ldw, p, st);
        }
    }
    switch (bn) {
        case 256: return launch_gemm<256, 1>(ta, W, ldw, p, st);
        case 192: return launch_gemm<192, 1>(ta, W, ldw, p, st);
        default: return launch_gemm<128, 1>(ta, W, ldw, p, st);
    }
}

}  // namespace cmh

extern "C" int cmh_gemm_set_trace(long long* device_buffer) {  // [grid][64] clock stamps per CTA; NULL = off
    cmh::g_trace = device_buffer;
    return CMH_OK;
}

extern "C" int cmh_gemm_force_tile(int bn, int cta_group) {
    if (!(bn == 0 || bn == 128 || bn == 192 || bn == 256) || cta_group < 0 || cta_group > 2)
        return cmh::fail(CMH_ERR_INVALID, "gemm_force_tile: bn in {0,128,192,256}, cta_group in {0,1,2}");
    cmh::g_force_bn = bn, cmh::g_force_cg = cta_group;
    return CMH_OK;
}

extern "C" int cmh_gemm_mma_lookahead(int kblocks) {  // debug / tuning: 1..8, default MMA_LOOKAHEAD
    if (kblocks < 1 || kblocks > 8) return cmh::fail(CMH_ERR_INVALID, "gemm_mma_lookahead: 1..8");
    cmh::g_lookahead = kblocks;
    return CMH_OK;
}

extern "C" int cmh_gemm_tail_slicing(int on) {  // debug: 0 = every tile of the last wave stays whole
    cmh::g_tail_slicing = on != 0;
    return CMH_OK;
}

extern "C" int cmh_gemm_force_units(int units) {  // debug: cap the number of CTAs (CG=1) / CTA pairs (CG=2); 0 = all SMs
    cmh::g_force_units = units;
    return CMH_OK;
}

extern "C" int cmh_gemm_bf16(const void* A, int64_t M, int64_t K, int64_t lda, const void* W, int64_t N, int64_t ldw,
                             const float* bias, int epilogue, void* out, int64_t ldo, const float* resid, int64_t ldr,
                             void* stream) {
    return cmh::gemm_bf16(A, M, K, lda, W, N, ldw, bias, epilogue, out, ldo, resid, ldr, cmh::as_stream(stream));
}
