// cmh_attention.cu — multi-head self-attention for the short CLIP sequences (50 image tokens, 32 text tokens).
//
// replaces nn.MultiheadAttention(x, x, x, need_weights=True, attn_mask, key_padding_mask) as called from
// ResidualAttentionBlock.attention (models/CLIP/model.py:181-189) between the in_proj and out_proj GEMMs.
//
// One thread block per (sample, head): the whole head (L <= 128 rows of Q, K, V, head dim 64) lives in shared
// memory, one warp owns 16 query rows.  S = Q.K^T and O = P.V run on the warp-level tensor-core path
// (mma.sync m16n8k16, bf16 in / fp32 accumulate; tiles of 16 x L x 64 are far below a tcgen05 M=128 atom),
// softmax stays in fp32 registers, the probabilities never leave the SM — the reference materialises
// [B.H, L, L] probabilities plus their head average in HBM for every block (need_weights=True) although only the
// last block's CLS / EOS row is ever read (model.py:265, :381).  That one row is written on request.
// HBM traffic per layer = read qkv (M x 3D bf16) + write out (M x D bf16): the kernel is HBM-bound.
#include <cuda_bf16.h>

#include "cmh_common.cuh"
#include "cmh_encoder.h"

namespace cmh {
namespace {

constexpr int DH = 64;  // head dimension of every CLIP transformer (width / 64 heads, model.py:300,465)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
// byte offset of 16-byte chunk `c` of row `r` in a [rows][64] bf16 tile (128-byte rows, XOR swizzle => conflict-free ldmatrix)
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return uint32_t(r * 128 + ((c ^ (r & 7)) << 4)); }

struct AttnParams {
    const __nv_bfloat16* qkv;  // [B*L][3*D]: q | k | v, head h at columns h*64
    __nv_bfloat16* out;        // [B*L][D]
    const uint8_t* pad;        // [B][L] key_padding_mask (1 = ignore key) or null
    float* probs;              // [B][H][L] probabilities of one query row per sample, or null
    const int32_t* probs_row;  // [B] query row to report (text: EOS), null => row 0 (image: CLS)
    int L, H, D, causal;
};

template <int LP>
__global__ void __launch_bounds__(LP * 2, LP == 128 ? 2 : (LP == 64 ? 7 : 12)) attn_kernel(const AttnParams p) {
    constexpr int NT = LP / 8;  // key tiles of 8
    extern __shared__ __align__(128) uint8_t attn_smem[];
    uint8_t* sq = attn_smem;
    uint8_t* sk = sq + LP * 128;
    uint8_t* sv = sk + LP * 128;
    uint8_t* spad = sv + LP * 128;

    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int L = p.L, D = p.D;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const __nv_bfloat16* base = p.qkv + size_t(b) * L * 3 * D + h * DH;

    pdl_wait();
    // Q and K first (one cp.async group), V second: S = Q.K^T and the softmax run while V is still in flight.
    // A thread copies the same 16-byte column chunk of 4 rows (LP/4 apart) of each part: one swizzled shared offset and one
    // global pointer per thread, constants added per copy (the generic index arithmetic was half of the kernel's instructions).
    {
        constexpr int RPP = LP / 4;  // rows per pass = threads / 8
        const int r0 = tid >> 3, c = tid & 7;
        const uint32_t soff = uint32_t(r0 * 128 + ((c ^ (r0 & 7)) << 4));  // (r0 + j*RPP) & 7 == r0 & 7
        const __nv_bfloat16* g0 = base + size_t(r0) * 3 * D + c * 8;
        const size_t gstep = size_t(RPP) * 3 * D;
#pragma unroll
        for (int part = 0; part < 3; ++part) {
            uint8_t* dstp = (part == 0 ? sq : part == 1 ? sk : sv) + soff;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool valid = r0 + j * RPP < L;
                cp_async16(dstp + j * RPP * 128, valid ? g0 + part * D + j * gstep : base, valid);
            }
            if (part >= 1) cp_async_commit();
        }
    }
    for (int j = tid; j < LP; j += LP * 2) spad[j] = (j >= L) || (p.pad && p.pad[size_t(b) * L + j]);
    cp_async_wait_group<1>();
    __syncthreads();
    pdl_launch_dependents();
    const bool active = warp * 16 < L;  // warps without a valid query row only take part in the barriers

    // ---- S = Q K^T -------------------------------------------------------------------------------------
    float s[NT][4];
    float inv0 = 0.f, inv1 = 0.f;
    const int r0 = warp * 16 + (lane >> 2), r1 = r0 + 8;
    const uint32_t q_base = smem_u32(sq), k_base = smem_u32(sk), v_base = smem_u32(sv);
    if (active) {
#pragma unroll
    for (int n = 0; n < NT; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
    // ldmatrix lane addresses: row term once, the swizzled 16-byte chunk per k-step from (chunk ^ (row & 7)), row & 7 == lane & 7
    const uint32_t q_row = q_base + uint32_t((warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * 128);
    const uint32_t k_row = k_base + uint32_t(((lane >> 4) * 8 + (lane & 7)) * 128);
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) {
        uint32_t a0, a1, a2, a3;
        ldsm_x4(q_row + uint32_t(((ks * 2 + (lane >> 4)) ^ (lane & 7)) << 4), a0, a1, a2, a3);
        const uint32_t k_sw = uint32_t(((ks * 2 + ((lane >> 3) & 1)) ^ (lane & 7)) << 4);
#pragma unroll
        for (int n = 0; n < NT; n += 2) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(k_row + n * 1024 + k_sw, b0, b1, b2, b3);
            mma_bf16(s[n], a0, a1, a2, a3, b0, b1);
            mma_bf16(s[n + 1], a0, a1, a2, a3, b2, b3);
        }
    }

    // ---- masked softmax over keys (fp32; scale = 1/sqrt(64) folded into the exp2 argument) ---------------
    const float kscale = 0.125f * 1.4426950408889634f;
    // dead keys (padding, j >= L) as bit masks in registers: one shared-memory byte per lane instead of one per element
    uint32_t dead[LP / 32];
#pragma unroll
    for (int w = 0; w < LP / 32; ++w) dead[w] = __ballot_sync(0xffffffffu, spad[w * 32 + lane] != 0);
    float m0 = -INFINITY, m1 = -INFINITY;
    const bool any_mask = p.causal || p.pad != nullptr;  // image tower: only the tail tiles (columns >= L) need masking
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        if (any_mask || n * 8 + 8 > L) {  // warp-uniform
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = n * 8 + (lane & 3) * 2 + e;
                const bool gone = (dead[(n * 8) / 32] >> (j & 31)) & 1u;
                if (gone || (p.causal && j > r0)) s[n][e] = -INFINITY;
                if (gone || (p.causal && j > r1)) s[n][2 + e] = -INFINITY;
            }
        }
        m0 = fmaxf(m0, fmaxf(s[n][0], s[n][1]));
        m1 = fmaxf(m1, fmaxf(s[n][2], s[n][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    // exp((s - m) / 8) = 2^(s*kscale - m*kscale): one FFMA + one MUFU.EX2 per element (a fully masked row keeps
    // m = -inf and yields nan, like the reference's softmax)
    const float mk0 = m0 * kscale, mk1 = m1 * kscale;
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            s[n][e] = ex2_approx(fmaf(s[n][e], kscale, -mk0));
            s[n][2 + e] = ex2_approx(fmaf(s[n][2 + e], kscale, -mk1));
            sum0 += s[n][e];
            sum1 += s[n][2 + e];
        }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    inv0 = 1.f / sum0, inv1 = 1.f / sum1;  // a fully masked row gives nan, like the reference

    if (p.probs) {  // need_weights row (model.py:265 CLS row, :381 EOS row) for the head average
        const int want = p.probs_row ? p.probs_row[b] : 0;
        float* dst = p.probs + (size_t(b) * p.H + h) * L;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int j = n * 8 + (lane & 3) * 2 + e;
                if (j < L && r0 == want) dst[j] = s[n][e] * inv0;
                if (j < L && r1 == want) dst[j] = s[n][2 + e] * inv1;
            }
        }
    }

    }  // active
    cp_async_wait_group<0>();
    __syncthreads();  // V has landed for every thread's copies
    if (!active) return;

    // ---- O = P V ------------------------------------------------------------------------------------------
    float o[DH / 8][4];
#pragma unroll
    for (int n = 0; n < DH / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    const uint32_t v_row = v_base + uint32_t((((lane >> 3) & 1) * 8 + (lane & 7)) * 128);  // + kk*16 rows per step
#pragma unroll
    for (int kk = 0; kk < LP / 16; ++kk) {
        const uint32_t a0 = pack2(s[2 * kk][0], s[2 * kk][1]), a1 = pack2(s[2 * kk][2], s[2 * kk][3]);
        const uint32_t a2 = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]), a3 = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int dn = 0; dn < DH / 8; dn += 2) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(v_row + kk * 2048 + uint32_t(((dn + (lane >> 4)) ^ (lane & 7)) << 4), b0, b1, b2, b3);
            mma_bf16(o[dn], a0, a1, a2, a3, b0, b1);
            mma_bf16(o[dn + 1], a0, a1, a2, a3, b2, b3);
        }
    }

    // ---- normalise, stage through this warp's own (now dead) Q rows, store 16 bytes per lane --------------
    __syncwarp();
    {
        const uint32_t st0 = q_base + uint32_t(r0 * 128 + (lane & 3) * 4), st1 = st0 + 8 * 128;  // r1 = r0 + 8, same row & 7
        const int sw = r0 & 7;
#pragma unroll
        for (int n = 0; n < DH / 8; ++n) {
            const uint32_t off = uint32_t((n ^ sw) << 4);
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(st0 + off), "r"(pack2(o[n][0] * inv0, o[n][1] * inv0)) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(st1 + off), "r"(pack2(o[n][2] * inv1, o[n][3] * inv1)) : "memory");
        }
    }
    __syncwarp();
    {
        const int c = lane & 7, rr = lane >> 3;  // 8 lanes cover one 128-byte row; 4 rows per instruction
        __nv_bfloat16* orow = p.out + (size_t(b) * L + warp * 16 + rr) * D + h * DH + c * 8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = warp * 16 + rr + 4 * i;
            if (r < L) {
                uint4 v;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                             : "r"(q_base + uint32_t(r * 128 + ((c ^ (r & 7)) << 4))));
                *reinterpret_cast<uint4*>(orow + size_t(4 * i) * D) = v;
            }
        }
    }
}

// probabilities [B][H][L] -> head average [B][L] with the reported query's own column cleared when asked
// (model.py:382: attn_weight[arange, EOS] = 0), dropping the first `skip` columns (model.py:265: [:, 0, 1:]).
__global__ void attn_mean_kernel(const float* __restrict__ probs, const int32_t* __restrict__ zero_col, float* out, int B,
                                 int H, int L, int skip) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_wait();
    if (i >= B * (L - skip)) return;
    const int b = i / (L - skip), j = i % (L - skip) + skip;
    float acc = 0.f;
    for (int h = 0; h < H; ++h) acc += probs[(size_t(b) * H + h) * L + j];
    acc /= float(H);
    if (zero_col && zero_col[b] == j) acc = 0.f;
    out[i] = acc;
}

}  // namespace

int attention_bf16(const void* qkv, int64_t B, int L, int H, const uint8_t* pad, int causal, void* out, float* probs,
                   const int32_t* probs_row, cudaStream_t st) {
    CMH_REQUIRE(qkv && out && B > 0 && L > 0 && H > 0, "attention: bad arguments");
    CMH_REQUIRE(L <= 128, "attention: sequence length %d > 128 is not supported", L);
    CMH_REQUIRE(B * H < (int64_t(1) << 31), "attention: too many (sample, head) pairs");
    AttnParams p{};
    p.qkv = static_cast<const __nv_bfloat16*>(qkv), p.out = static_cast<__nv_bfloat16*>(out), p.pad = pad;
    p.probs = probs, p.probs_row = probs_row, p.L = L, p.H = H, p.D = H * DH, p.causal = causal;
    const unsigned grid = unsigned(B * H);
    if (L <= 32) {
        CMH_CUDA_TRY(launch_kernel(attn_kernel<32>, dim3(grid), dim3(64), 32 * 385, st, 1, p));
    } else if (L <= 64) {
        CMH_CUDA_TRY(launch_kernel(attn_kernel<64>, dim3(grid), dim3(128), 64 * 385, st, 1, p));
    } else {
        static PerDeviceOnce once;
        if (once.needs()) {
            CMH_CUDA_TRY(cudaFuncSetAttribute(attn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 385));
        }
        CMH_CUDA_TRY(launch_kernel(attn_kernel<128>, dim3(grid), dim3(256), 128 * 385, st, 1, p));
    }
    return CMH_OK;
}

int attention_mean(const float* probs, int64_t B, int H, int L, int skip, const int32_t* zero_col, float* out,
                   cudaStream_t st) {
    const int n = int(B) * (L - skip);
    CMH_CUDA_TRY(launch_kernel(attn_mean_kernel, dim3((n + 255) / 256), dim3(256), 0, st, 1, probs, zero_col, out, int(B), H, L, skip));
    return CMH_OK;
}

}  // namespace cmh

extern "C" int cmh_attention_bf16(const void* qkv, int64_t batch, int seq_len, int heads, const uint8_t* key_padding_mask,
                                  int causal, void* out, void* stream) {
    return cmh::attention_bf16(qkv, batch, seq_len, heads, key_padding_mask, causal, out, nullptr, nullptr,
                               cmh::as_stream(stream));
}
