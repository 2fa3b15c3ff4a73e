// cmh_mith.cu — kernels of the MITH hash head (models/MITH/hash/hash.py) that are not plain GEMM / LayerNorm / attention:
// localized token aggregation (top-k concepts per token, softmax over tokens, weighted token sum + positional encoding),
// the per-bit hashing Linear(dim, 1) layers, row normalisation and the final sign(cls_hash + tokens_hash) packing.
// The residual MLPs, the 2-block transformer over the K concept tokens and the concept projection reuse the encoder's
// tcgen05 GEMM, LayerNorm and attention kernels (cmh_encoder.cu: cmh_head_mith).
#include <cuda_bf16.h>

#include "cmh_common.cuh"
#include "cmh_encoder.h"

namespace cmh {
namespace {

// dst[b*L + l] = src[(b*tokens_per_sample + first + l)]   (fp32 rows of D floats; D % 4 == 0)
__global__ void gather_rows_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int64_t B, int L, int per_sample,
                                   int first, int d4) {
    const int64_t total = B * L * d4;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int c = int(i % d4);
        const int64_t row = i / d4, b = row / L;
        const int l = int(row % L);
        dst[i] = __ldg(src + ((b * per_sample + first + l) * d4 + c));
    }
}

// LocalizedTokenAggregation.forward (hash.py:109-170) + PositionalEncoding (hash.py:41-64), one block per sample.
//   concept [B*L][K] fp32 (tanh concept embedding of every token), tokens x [B][per_sample][D] fp32 (rows first..first+L),
//   pad [B][L] (1 = padded token) or null  ->  out [B*K][D] fp32 = sum_l p[l][k] x[l] + pos[k]
constexpr int LTA_THREADS = 256;
__global__ void __launch_bounds__(LTA_THREADS)
lta_kernel(const float* __restrict__ concept, const float* __restrict__ x, const uint8_t* __restrict__ pad,
           const float* __restrict__ pos, int L, int K, int D, int per_sample, int first, int top_k, float* __restrict__ out) {
    extern __shared__ float sim[];  // [L][K + 1]
    const int b = blockIdx.x, tid = threadIdx.x, ld = K + 1;
    const float NEG = -INFINITY;
    for (int i = tid; i < L * K; i += LTA_THREADS) {
        const int l = i / K, k = i % K;
        float v = concept[(size_t(b) * L + l) * K + k];
        if ((pad && pad[size_t(b) * L + l]) || !(v > 0.f)) v = NEG;  // hash.py:140-154
        sim[l * ld + k] = v;
    }
    __syncthreads();
    // each token keeps the concepts >= its top_k-th largest similarity (ties kept, hash.py:114-124)
    for (int l = tid; l < L; l += LTA_THREADS) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = NEG;
        for (int k = 0; k < K; ++k) {
            float v = sim[l * ld + k];
#pragma unroll
            for (int j = 0; j < 8; ++j) {  // sorted insert, t[0] >= ... >= t[7]
                if (j < top_k && v > t[j]) {
                    const float tmp = t[j];
                    t[j] = v;
                    v = tmp;
                }
            }
        }
        const float kth = t[top_k - 1];
        for (int k = 0; k < K; ++k)
            if (sim[l * ld + k] < kth) sim[l * ld + k] = NEG;
    }
    __syncthreads();
    // softmax over the tokens of every concept; a concept nobody selected gives 0 (hash.py:159-160)
    for (int k = tid; k < K; k += LTA_THREADS) {
        float m = NEG;
        for (int l = 0; l < L; ++l) m = fmaxf(m, sim[l * ld + k]);
        float s = 0.f;
        for (int l = 0; l < L; ++l) s += (m == NEG) ? 0.f : expf(sim[l * ld + k] - m);
        for (int l = 0; l < L; ++l) sim[l * ld + k] = (m == NEG) ? 0.f : expf(sim[l * ld + k] - m) / s;
    }
    __syncthreads();
    // merged[k] = sum_l p[l][k] * x[l]  (+ positional encoding); a thread owns one float4 column and 8 concepts at a time
    const int d4 = D / 4;
    const float4* xb = reinterpret_cast<const float4*>(x) + (size_t(b) * per_sample + first) * d4;
    for (int item = tid; item < d4 * ((K + 7) / 8); item += LTA_THREADS) {
        const int c = item % d4, k0 = (item / d4) * 8;
        float4 acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int l = 0; l < L; ++l) {
            const float4 xv = __ldg(xb + size_t(l) * d4 + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float pw = (k0 + j < K) ? sim[l * ld + k0 + j] : 0.f;
                acc[j].x = fmaf(pw, xv.x, acc[j].x), acc[j].y = fmaf(pw, xv.y, acc[j].y);
                acc[j].z = fmaf(pw, xv.z, acc[j].z), acc[j].w = fmaf(pw, xv.w, acc[j].w);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (k0 + j < K) {
                const float4 pe = __ldg(reinterpret_cast<const float4*>(pos) + size_t(k0 + j) * d4 + c);
                reinterpret_cast<float4*>(out)[(size_t(b) * K + k0 + j) * d4 + c] =
                    make_float4(acc[j].x + pe.x, acc[j].y + pe.y, acc[j].z + pe.z, acc[j].w + pe.w);
            }
        }
    }
}

// BitwiseHashing.forward (hash.py:67-83): out[b][k] = tanh(x[b*K + k] . w[k] + bias[k]); one warp per row
__global__ void __launch_bounds__(256)
bit_hash_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int64_t rows, int K, int D,
                float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int k = int(r % K);
    const float4* xr = reinterpret_cast<const float4*>(x + r * D);
    const float4* wr = reinterpret_cast<const float4*>(w + size_t(k) * D);
    float acc = 0.f;
    for (int i = lane; i < D / 4; i += 32) {
        const float4 a = xr[i], c = __ldg(wr + i);
        acc = fmaf(a.x, c.x, fmaf(a.y, c.y, fmaf(a.z, c.z, fmaf(a.w, c.w, acc))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[r] = tanhf(acc + bias[k]);
}

// F.normalize(x, dim=-1): x / max(||x||_2, 1e-12); one warp per row, in place or to `out`
__global__ void __launch_bounds__(256)
normalize_rows_kernel(const float* __restrict__ x, int64_t rows, int D, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + r * D);
    float q = 0.f;
    for (int i = lane; i < D / 4; i += 32) {
        const float4 a = xr[i];
        q += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float inv = 1.f / fmaxf(sqrtf(q), 1e-12f);
    float4* orow = reinterpret_cast<float4*>(out + r * D);
    for (int i = lane; i < D / 4; i += 32) {
        const float4 a = xr[i];
        orow[i] = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
    }
}

// fp32 rows -> bf16 rows (A operand of the concept projection GEMM)
__global__ void cast_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ out, int64_t n4) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
        const float4 v = x[i];
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        out[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
}

// fp32 rows [M][D] -> bf16 rows [M][3D] = [hi | hi | lo], hi = bf16(x), lo = bf16(x - hi): the A operand of a K-concatenated
// split-precision GEMM against W' = [W_hi | W_lo | W_hi]
__global__ void split3_kernel(const float4* __restrict__ x, uint2* __restrict__ out, int64_t M, int d4) {
    const int64_t total = M * d4;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t m = i / d4;
        const int c = int(i % d4);
        const float4 v = x[i];
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
        const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
        const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
        const uint2 hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
        const uint2 lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
        uint2* row = out + m * (3 * d4);
        row[c] = hi;
        row[d4 + c] = hi;
        row[2 * d4 + c] = lo;
    }
}

// MITHTrainer.generate_hash + make_hash_code (runners/MITH/runner.py:125-131): bit = (cls_hash + tokens_hash) > 0
__global__ void add_sign_pack_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t rows, int nbits, int W,
                                     uint32_t* __restrict__ packed) {
    const int64_t gw = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= rows * W) return;
    const int64_t r = gw / W;
    const int j = int(gw % W) * 32 + lane;
    const bool bit = j < nbits && (a[r * nbits + j] + b[r * nbits + j]) > 0.f;
    const uint32_t word = __ballot_sync(0xffffffffu, bit);
    if (lane == 0) packed[gw] = word;
}

unsigned grid_for(int64_t items, int per_block) {
    const int64_t blocks = ceil_div(items, per_block);
    const int64_t cap = int64_t(sm_count_cached()) * 32;
    return unsigned(blocks < cap ? blocks : cap);
}

}  // namespace

int gather_rows(const float* src, float* dst, int64_t B, int L, int per_sample, int first, int D, cudaStream_t st) {
    gather_rows_kernel<<<grid_for(B * L * (D / 4), 256), 256, 0, st>>>(reinterpret_cast<const float4*>(src),
                                                                     reinterpret_cast<float4*>(dst), B, L, per_sample, first, D / 4);
    CMH_LAUNCH_CHECK("gather_rows_kernel");
    return CMH_OK;
}

int token_aggregation(const float* concept, const float* x, const uint8_t* pad, const float* pos, int64_t B, int L, int K, int D,
                      int per_sample, int first, int top_k, float* out, cudaStream_t st) {
    CMH_REQUIRE(top_k >= 1 && top_k <= 8 && top_k <= K, "mith: top_k_label %d unsupported (1..8, <= k_bits)", top_k);
    const size_t smem = size_t(L) * (K + 1) * sizeof(float);
    CMH_REQUIRE(smem <= 48 * 1024, "mith: %d tokens x %d concepts exceed the aggregation kernel's shared memory", L, K);
    lta_kernel<<<unsigned(B), LTA_THREADS, smem, st>>>(concept, x, pad, pos, L, K, D, per_sample, first, top_k, out);
    CMH_LAUNCH_CHECK("lta_kernel");
    return CMH_OK;
}

int bit_hash(const float* x, const float* w, const float* bias, int64_t rows, int K, int D, float* out, cudaStream_t st) {
    bit_hash_kernel<<<unsigned(ceil_div(rows, 8)), 256, 0, st>>>(x, w, bias, rows, K, D, out);
    CMH_LAUNCH_CHECK("bit_hash_kernel");
    return CMH_OK;
}

int normalize_rows(const float* x, int64_t rows, int D, float* out, cudaStream_t st) {
    normalize_rows_kernel<<<unsigned(ceil_div(rows, 8)), 256, 0, st>>>(x, rows, D, out);
    CMH_LAUNCH_CHECK("normalize_rows_kernel");
    return CMH_OK;
}

int cast_bf16(const float* x, void* out, int64_t n, cudaStream_t st) {
    cast_bf16_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<uint2*>(out), n / 4);
    CMH_LAUNCH_CHECK("cast_bf16_kernel");
    return CMH_OK;
}

int split3_bf16(const float* x, void* out, int64_t M, int D, cudaStream_t st) {
    split3_kernel<<<grid_for(M * (D / 4), 256), 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<uint2*>(out), M, D / 4);
    CMH_LAUNCH_CHECK("split3_kernel");
    return CMH_OK;
}

int add_sign_pack(const float* a, const float* b, int64_t rows, int nbits, uint32_t* packed, cudaStream_t st) {
    const int W = cmh_code_words(nbits);
    CMH_REQUIRE(W > 0, "mith: %d bits unsupported", nbits);
    add_sign_pack_kernel<<<unsigned(ceil_div(rows * W * 32, 256)), 256, 0, st>>>(a, b, rows, nbits, W, packed);
    CMH_LAUNCH_CHECK("add_sign_pack_kernel");
    return CMH_OK;
}

}  // namespace cmh
