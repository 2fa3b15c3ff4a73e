// cmh_gemm.cu — bf16 x bf16 -> fp32 GEMM for the CLIP encoder blocks on 5th-generation tensor cores.
//
//   out[M][N] = epilogue( A[M][K] . W[N][K]^T + bias[N] )         (torch.nn.Linear layout: both operands K-major)
//
// replaces the fp32 cuBLAS SGEMMs behind nn.Linear / nn.MultiheadAttention / conv1 of models/CLIP/model.py:167-268.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer   cp.async.bulk.tensor.2d (128B swizzle) of a 128 x 64 A tile and a BN x 64 W tile
//                              into a 4..6-stage shared-memory ring, completion on mbarriers (expect_tx)
//   warp 1      MMA issuer     one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) x 4
//                              per stage; tcgen05.commit releases the stage / publishes the accumulator
//   warp 2      TMEM owner     tcgen05.alloc of 2 x BN fp32 accumulator columns (double buffered) / dealloc
//   warps 4..7  epilogue       tcgen05.ld 32x32b.x32 -> registers -> bias / QuickGELU / tanh -> swizzled smem staging
//                              -> TMA tile store (cp.async.bulk.tensor) or, for the residual stream, TMA fp32
//                              reduce-add (cp.reduce.async.bulk.tensor .add) straight into x
// The epilogue of tile i overlaps the MMAs of tile i+1 through the two TMEM accumulator stages.
#include <cuda.h>
#include <cuda_bf16.h>

#include "cmh_common.cuh"
#include "cmh_encoder.h"

namespace cmh {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 256;
constexpr int SMEM_BUDGET = 192 * 1024;   // operand ring; the epilogue staging takes another 16 KB + 4 KB
constexpr int STG_WARP_BYTES = 4096;      // per epilogue warp: 2 x (32 rows x 64 B bf16) or 1 x (32 rows x 128 B fp32)

template <int BN>
struct GemmCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES < 8 ? SMEM_BUDGET / STAGE_BYTES : 8;
    static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
    static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
    static constexpr int SMEM_BYTES = RING_BYTES + 4 * STG_WARP_BYTES + 4 * 256 * 4 /*bias*/ + 1024 /*align*/ + 256 /*barriers*/;
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
};

// ---- PTX wrappers ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// shared -> global tile store / fp32 reduce-add through the tensor map (clips rows/columns outside the tensor)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, int c0, int c1, const void* smem_src) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr >> 4) & 0x3FFF);
    d |= uint64_t(1) << 16;                  // leading byte offset (unused for swizzled K-major), 16 B units
    d |= uint64_t(1024 >> 4) << 32;          // stride byte offset between 8-row groups
    d |= uint64_t(1) << 46;                  // descriptor version
    d |= uint64_t(2) << 61;                  // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16(int n) {
    return (1u << 4)            // D = fp32
           | (1u << 7)          // A = bf16
           | (1u << 10)         // B = bf16
           | (uint32_t(n >> 3) << 17) | (uint32_t(BM >> 4) << 24);
}

__device__ __forceinline__ float quick_gelu(float x) {  // x * sigmoid(1.702 x)   models/CLIP/model.py:162-164
    return x / (1.0f + __expf(-1.702f * x));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const GemmParams p) {
    using C = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stg_all = smem + C::RING_BYTES;                                   // 4 warps x 4 KB, 1024-aligned
    float* bias_all = reinterpret_cast<float*>(stg_all + 4 * STG_WARP_BYTES);  // 4 warps x 256 floats
    uint64_t* full = reinterpret_cast<uint64_t*>(bias_all + 4 * 256);
    uint64_t* empty = full + C::STAGES;
    uint64_t* tfull = empty + C::STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = p.tiles_m * p.tiles_n;
    const int kblocks = int((p.K + BK - 1) / BK);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        prefetch_tmap(&tmO);
        for (int s = 0; s < C::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull[s], 1);
            mbar_init(&tempty[s], 4);  // one arrival per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m0 = (t % p.tiles_m) * BM, n0 = (t / p.tiles_m) * BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1u);
                    uint8_t* sa = smem + stage * C::STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[stage], C::STAGE_BYTES);
                    tma_load_2d(sa, &tmA, kb * BK, m0, &full[stage]);
                    tma_load_2d(sa + C::A_BYTES, &tmB, kb * BK, n0, &full[stage]);
                    if (++stage == C::STAGES) stage = 0, phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty[as], aphase ^ 1u);  // epilogue has drained this accumulator stage
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(as * BN);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
                    const uint64_t adesc = make_desc_sw128(sa), bdesc = make_desc_sw128(sa + C::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advancing 16 bf16 = 32 B inside the swizzle row: +2 in 16-byte units
                        umma_f16(d_tmem, adesc + uint64_t(2 * k), bdesc + uint64_t(2 * k), idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty[stage]);  // stage reusable once these MMAs have read it
                    if (++stage == C::STAGES) stage = 0, phase ^= 1u;
                }
                umma_commit(&tfull[as]);  // accumulator complete
            }
        }
    } else if (warp >= 4) {
        // Epilogue: each warp owns 32 accumulator rows (its TMEM lane quarter).  A 32-column chunk goes
        // TMEM -> registers -> (+bias, activation) -> swizzled shared staging -> TMA store (or fp32 reduce-add
        // into the residual stream), double buffered per warp; stores are asynchronous and fully coalesced.
        const int e = warp - 4;
        const bool out_bf16 = p.epi == CMH_EPI_BF16 || p.epi == CMH_EPI_GELU_BF16;
        uint8_t* stg = stg_all + e * STG_WARP_BYTES;
        float* bias_s = bias_all + e * 256;
        int it = 0;
        uint32_t chunk_no = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m0 = (t % p.tiles_m) * BM, n0 = (t / p.tiles_m) * BN;
            for (int j = lane; j < BN; j += 32) bias_s[j] = (p.bias && n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
            __syncwarp();
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                if (n0 + c0 >= p.N) break;  // remaining chunks lie outside the matrix (uniform)
                uint32_t r[32];
                tmem_ld32(tmem_base + (uint32_t(e * 32) << 16) + uint32_t(as * BN + c0), r);
                uint8_t* buf = out_bf16 ? stg + (chunk_no & 1u) * (STG_WARP_BYTES / 2) : stg;
                ++chunk_no;
                if (lane == 0) {  // the store that last used this buffer has finished reading it
                    if (out_bf16) bulk_wait_read<1>(); else bulk_wait_read<0>();
                }
                __syncwarp();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b = *reinterpret_cast<const float4*>(bias_s + c0 + j);
                    v[j] = __uint_as_float(r[j]) + b.x, v[j + 1] = __uint_as_float(r[j + 1]) + b.y;
                    v[j + 2] = __uint_as_float(r[j + 2]) + b.z, v[j + 3] = __uint_as_float(r[j + 3]) + b.w;
                }
                if (p.epi == CMH_EPI_GELU_BF16) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = quick_gelu(v[j]);
                } else if (p.epi == CMH_EPI_TANH_F32) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
                }
                if (out_bf16) {  // 32 rows x 64 B, SWIZZLE_64B: 16-byte chunk index ^= (row >> 1) & 3
                    uint8_t* rowp = buf + lane * 64;
                    const int sw = (lane >> 1) & 3;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 o;
                        o.x = pack_bf16(v[8 * c], v[8 * c + 1]), o.y = pack_bf16(v[8 * c + 2], v[8 * c + 3]);
                        o.z = pack_bf16(v[8 * c + 4], v[8 * c + 5]), o.w = pack_bf16(v[8 * c + 6], v[8 * c + 7]);
                        *reinterpret_cast<uint4*>(rowp + ((c ^ sw) << 4)) = o;
                    }
                } else {  // 32 rows x 128 B, SWIZZLE_128B: chunk index ^= row & 7
                    uint8_t* rowp = buf + lane * 128;
                    const int sw = lane & 7;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<float4*>(rowp + ((c ^ sw) << 4)) =
                            make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (p.epi == CMH_EPI_RESID_F32) tma_reduce_add_2d(&tmO, n0 + c0, m0 + e * 32, buf);
                    else tma_store_2d(&tmO, n0 + c0, m0 + e * 32, buf);
                    bulk_commit();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
        }
        if (lane == 0) bulk_wait_read<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
}

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

// output tile map: 32 rows x 32 columns per store, bf16 (64-byte rows, SWIZZLE_64B) or fp32 (128-byte rows, SWIZZLE_128B)
int make_out_tmap(CUtensorMap* map, void* base, int64_t rows, int64_t cols, int64_t ld, bool bf16) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(CMH_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {cuuint64_t(cols), cuuint64_t(rows)};
    cuuint64_t strides[1] = {cuuint64_t(ld) * (bf16 ? 2 : 4)};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, bf16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(CMH_ERR_CUDA, "cuTensorMapEncodeTiled (output) failed with CUresult %d", int(r));
    return CMH_OK;
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

template <int BN>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, GemmParams p, cudaStream_t st) {
    using C = GemmCfg<BN>;
    static bool configured = false;
    if (!configured) {
        CMH_CUDA_TRY(cudaFuncSetAttribute(gemm_bf16_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured = true;
    }
    p.tiles_m = int(ceil_div(p.M, BM));
    p.tiles_n = int(ceil_div(p.N, BN));
    const int tiles = p.tiles_m * p.tiles_n;
    const int grid = tiles < sm_count_cached() ? tiles : sm_count_cached();
    gemm_bf16_kernel<BN><<<grid, GEMM_THREADS, C::SMEM_BYTES, st>>>(ta, tb, to, p);
    CMH_LAUNCH_CHECK("gemm_bf16_kernel");
    return CMH_OK;
}

// tile width: fewest "wave x tile cost" units on this device
int pick_bn(int64_t M, int64_t N) {
    const int sms = sm_count_cached();
    const int64_t tm = ceil_div(M, BM);
    int best = 128;
    double best_cost = 1e30;
    for (int bn : {256, 192, 128}) {
        if (N < bn && bn != 128) continue;
        const int64_t tiles = tm * ceil_div(N, bn);
        const double cost = double(ceil_div(tiles, sms)) * bn;
        if (cost < best_cost - 1e-9) best_cost = cost, best = bn;
    }
    return best;
}

}  // namespace

int gemm_bf16(const void* A, int64_t M, int64_t K, int64_t lda, const void* W, int64_t N, int64_t ldw,
              const float* bias, int epi, void* out, int64_t ldo, const float* resid, int64_t ldr, cudaStream_t st) {
    CMH_REQUIRE(A && W && out && M > 0 && N > 0 && K > 0, "gemm: bad arguments");
    CMH_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, "gemm: leading dimensions must be multiples of 8 elements");
    CMH_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "gemm: operands must be 16-byte aligned");
    CMH_REQUIRE(epi >= CMH_EPI_BF16 && epi <= CMH_EPI_TANH_F32, "gemm: unknown epilogue %d", epi);
    CMH_REQUIRE(epi != CMH_EPI_RESID_F32 || resid, "gemm: residual epilogue needs resid");
    const bool out_bf16 = epi == CMH_EPI_BF16 || epi == CMH_EPI_GELU_BF16;
    CMH_REQUIRE(ldo % (out_bf16 ? 8 : 4) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "gemm: output must be 16-byte aligned per row");
    CMH_REQUIRE(epi != CMH_EPI_RESID_F32 || (resid == out && ldr == ldo),
                "gemm: the residual epilogue works in place (resid == out): it is a TMA fp32 reduce-add into the stream");
    const int bn = pick_bn(M, N);
    CUtensorMap ta, tb, to;
    if (int rc = make_tmap(&ta, A, M, K, lda, BM)) return rc;
    if (int rc = make_tmap(&tb, W, N, K, ldw, bn)) return rc;
    if (int rc = make_out_tmap(&to, out, M, N, ldo, out_bf16)) return rc;
    GemmParams p{};
    p.M = M, p.N = N, p.K = K, p.bias = bias, p.out = out, p.ldo = ldo, p.resid = resid, p.ldr = ldr, p.epi = epi;
    switch (bn) {
        case 256: return launch_gemm<256>(ta, tb, to, p, st);
        case 192: return launch_gemm<192>(ta, tb, to, p, st);
        default: return launch_gemm<128>(ta, tb, to, p, st);
    }
}

}  // namespace cmh

extern "C" int cmh_gemm_bf16(const void* A, int64_t M, int64_t K, int64_t lda, const void* W, int64_t N, int64_t ldw,
                             const float* bias, int epilogue, void* out, int64_t ldo, const float* resid, int64_t ldr,
                             void* stream) {
    return cmh::gemm_bf16(A, M, K, lda, W, N, ldw, bias, epilogue, out, ldo, resid, ldr, cmh::as_stream(stream));
}
