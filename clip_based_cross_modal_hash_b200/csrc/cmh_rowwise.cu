// cmh_rowwise.cu — the memory-bound row kernels of the CLIP encoders: LayerNorm, patch gathering, token/position
// embedding, CLS assembly.  One warp per row of the residual stream, 16-byte loads/stores, fp32 statistics
// (models/CLIP/model.py:153-159 evaluates LayerNorm in fp32 whatever the stream dtype is).
#include <cuda_bf16.h>

#include "cmh_common.cuh"
#include "cmh_encoder.h"

namespace cmh {
namespace {

constexpr int ROWS_PER_BLOCK = 8;  // 8 warps

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// LayerNorm of one row held as NV float4 per lane (D = NV*128): two-pass mean / biased variance in fp32,
// eps inside the square root (nn.LayerNorm), then gain and bias.
template <int NV>
__device__ __forceinline__ void warp_layernorm(float4 (&v)[NV], const float* __restrict__ g, const float* __restrict__ bta,
                                               int lane, float eps) {
    constexpr float invD = 1.0f / float(NV * 128);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mu = warp_sum(s) * invD;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i].x -= mu, v[i].y -= mu, v[i].z -= mu, v[i].w -= mu;
        q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float rstd = rsqrtf(warp_sum(q) * invD + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i * 32 + lane);
        const float4 bb = __ldg(reinterpret_cast<const float4*>(bta) + i * 32 + lane);
        v[i].x = v[i].x * rstd * gg.x + bb.x, v[i].y = v[i].y * rstd * gg.y + bb.y;
        v[i].z = v[i].z * rstd * gg.z + bb.z, v[i].w = v[i].w * rstd * gg.w + bb.w;
    }
}

template <int NV, bool OUT_F32>
__device__ __forceinline__ void store_row(const float4 (&v)[NV], void* out, int64_t row, int lane) {
    if (OUT_F32) {
        float4* o = reinterpret_cast<float4*>(out) + row * (NV * 32);
#pragma unroll
        for (int i = 0; i < NV; ++i) o[i * 32 + lane] = v[i];
    } else {
        uint2* o = reinterpret_cast<uint2*>(out) + row * (NV * 32);
#pragma unroll
        for (int i = 0; i < NV; ++i) o[i * 32 + lane] = make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
    }
}

// out[r] = LN(x[src(r)]);  src(r) = r*row_mul + (row_idx ? row_idx[r] : 0)   (CLS rows: mul = L; EOS rows: mul = L, idx = eos)
template <int NV, bool OUT_F32>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
layernorm_kernel(const float* __restrict__ x, int64_t rows, int64_t row_mul, const int32_t* __restrict__ row_idx,
                 const float* __restrict__ g, const float* __restrict__ b, float eps, void* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = int64_t(blockIdx.x) * ROWS_PER_BLOCK + (threadIdx.x >> 5);
    pdl_wait();
    if (r >= rows) return;
    const int64_t src = r * row_mul + (row_idx ? row_idx[r] : 0);
    const float4* xr = reinterpret_cast<const float4*>(x) + src * (NV * 32);
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = xr[i * 32 + lane];
    pdl_launch_dependents();  // the loads are in: the next kernel may start taking SMs as these blocks retire
    warp_layernorm<NV>(v, g, b, lane, eps);
    store_row<NV, OUT_F32>(v, out, r, lane);
}

// x[b*L + l] = ln_pre( (l == 0 ? class_embedding : patch_embed[b*(L-1) + l-1]) + positional_embedding[l] )
// models/CLIP/model.py:241-243
template <int NV>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
vit_assemble_kernel(const float* __restrict__ emb, const float* __restrict__ cls, const float* __restrict__ pos, int64_t rows,
                    int L, const float* __restrict__ g, const float* __restrict__ b, float eps, float* __restrict__ x) {
    const int lane = threadIdx.x & 31;
    const int64_t r = int64_t(blockIdx.x) * ROWS_PER_BLOCK + (threadIdx.x >> 5);
    pdl_wait();
    if (r >= rows) return;
    const int64_t bi = r / L;
    const int l = int(r % L);
    const float4* src = l == 0 ? reinterpret_cast<const float4*>(cls)
                               : reinterpret_cast<const float4*>(emb) + (bi * (L - 1) + (l - 1)) * (NV * 32);
    const float4* pr = reinterpret_cast<const float4*>(pos) + int64_t(l) * (NV * 32);
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 a = src[i * 32 + lane], p = __ldg(pr + i * 32 + lane);
        v[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    }
    warp_layernorm<NV>(v, g, b, lane, eps);
    store_row<NV, true>(v, x, r, lane);
}

// x[b*L + l] = token_embedding[text[b][l]] + positional_embedding[l]      models/CLIP/model.py:374-376
template <int NV>
__global__ void __launch_bounds__(ROWS_PER_BLOCK * 32)
text_embed_kernel(const int64_t* __restrict__ text, const float* __restrict__ tok, const float* __restrict__ pos, int64_t rows,
                  int L, int vocab, float* __restrict__ x) {
    const int lane = threadIdx.x & 31;
    const int64_t r = int64_t(blockIdx.x) * ROWS_PER_BLOCK + (threadIdx.x >> 5);
    pdl_wait();
    if (r >= rows) return;
    int64_t id = text[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);  // ids are validated on the host side of the ABI; never read outside the table
    const float4* tr = reinterpret_cast<const float4*>(tok) + id * (NV * 32);
    const float4* pr = reinterpret_cast<const float4*>(pos) + (r % L) * (NV * 32);
    float4* o = reinterpret_cast<float4*>(x) + r * (NV * 32);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const float4 a = __ldg(tr + i * 32 + lane), p = __ldg(pr + i * 32 + lane);
        o[i * 32 + lane] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    }
}

// eos[b] = argmax_l text[b][l] (first maximum, like torch.argmax; EOT 49407 is the largest id, model.py:379);
// new_mask[b][l] = pad[b][l] | (text[b][l] == eot_id)  (model.py:384)
__global__ void text_eos_kernel(const int64_t* __restrict__ text, const uint8_t* __restrict__ pad, int64_t B, int L,
                                int64_t eot_id, int32_t* __restrict__ eos, uint8_t* __restrict__ new_mask) {
    const int64_t b = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    pdl_wait();
    if (b >= B) return;
    const int64_t* t = text + b * L;
    int best = 0;
    int64_t bv = t[0];
    for (int l = 0; l < L; ++l) {
        const int64_t v = t[l];
        if (v > bv) bv = v, best = l;
        if (new_mask) new_mask[b * L + l] = uint8_t((pad && pad[b * L + l]) || v == eot_id);
    }
    eos[b] = best;
}

// patches[(b*g*g + py*g + px)][c*P*P + ky*P + kx] = bf16(image[b][c][py*P + ky][px*P + kx])
// Conv2d(3, width, kernel = stride = P, bias=False) (model.py:219,235) as a GEMM over non-overlapping patches.
// Thread = 8 consecutive pixels of one image row (32 B read, 16 B write); reads are fully coalesced along x.
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ img, int64_t B, int C, int R, int P, __nv_bfloat16* __restrict__ out) {
    const int g = R / P, xv = R / 8;
    const int64_t total = B * C * R * xv;
    pdl_wait();
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int x8 = int(i % xv);
        const int y = int((i / xv) % R);
        const int c = int((i / (int64_t(xv) * R)) % C);
        const int64_t b = i / (int64_t(xv) * R * C);
        const float4* src = reinterpret_cast<const float4*>(img + ((b * C + c) * R + y) * R + x8 * 8);
        const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
        const int x = x8 * 8, px = x / P, kx = x % P, py = y / P, ky = y % P;
        __nv_bfloat16* dst = out + ((b * g + py) * g + px) * int64_t(C * P * P) + (c * P + ky) * P + kx;
        *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w), pack_bf16x2(v1.x, v1.y),
                                                     pack_bf16x2(v1.z, v1.w));
    }
}

// uint8 pixels straight from the dataloader (dataset/transformer_dataset.py:41-45: ToTensor's /255 and Normalize(mean, std) are
// the only arithmetic left after Resize/CenterCrop): out = ((x / 255) - mean[c]) / std[c] as bf16 patches.  Thread = 16
// consecutive pixels of one image row (16 B read, 32 B write).  A quarter of the fp32 bytes cross PCIe and HBM.
__global__ void __launch_bounds__(256)
patchify_u8_kernel(const uint8_t* __restrict__ img, int64_t B, int C, int R, int P, float3 scale, float3 shift,
                   __nv_bfloat16* __restrict__ out) {
    const int g = R / P, xv = R / 16;
    const int64_t total = B * C * R * xv;
    pdl_wait();
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int x16 = int(i % xv);
        const int y = int((i / xv) % R);
        const int c = int((i / (int64_t(xv) * R)) % C);
        const int64_t b = i / (int64_t(xv) * R * C);
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(img + ((b * C + c) * R + y) * R + x16 * 16));
        const float sc = c == 0 ? scale.x : c == 1 ? scale.y : scale.z, sh = c == 0 ? shift.x : c == 1 ? shift.y : shift.z;
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float p0 = fmaf(float(w[j] & 0xFFu), sc, sh), p1 = fmaf(float((w[j] >> 8) & 0xFFu), sc, sh);
            const float p2 = fmaf(float((w[j] >> 16) & 0xFFu), sc, sh), p3 = fmaf(float(w[j] >> 24), sc, sh);
            o[2 * j] = pack_bf16x2(p0, p1), o[2 * j + 1] = pack_bf16x2(p2, p3);
        }
        const int x = x16 * 16, px = x / P, kx = x % P, py = y / P, ky = y % P;
        uint4* dst = reinterpret_cast<uint4*>(out + ((b * g + py) * g + px) * int64_t(C * P * P) + (c * P + ky) * P + kx);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

template <int NV, bool OUT_F32>
int launch_ln(const float* x, int64_t rows, int64_t row_mul, const int32_t* row_idx, const float* g, const float* b, float eps,
              void* out, cudaStream_t st) {
    CMH_CUDA_TRY(launch_kernel(layernorm_kernel<NV, OUT_F32>, dim3(unsigned(ceil_div(rows, ROWS_PER_BLOCK))),
                               dim3(ROWS_PER_BLOCK * 32), 0, st, 1, x, rows, row_mul, row_idx, g, b, eps, out));
    return CMH_OK;
}

#define CMH_DISPATCH_NV(D, CALL)                                                                    \
    switch ((D) / 128) {                                                                            \
        case 1: { constexpr int NV = 1; CALL; } break;                                              \
        case 2: { constexpr int NV = 2; CALL; } break;                                              \
        case 3: { constexpr int NV = 3; CALL; } break;                                              \
        case 4: { constexpr int NV = 4; CALL; } break;                                              \
        case 6: { constexpr int NV = 6; CALL; } break;                                              \
        case 8: { constexpr int NV = 8; CALL; } break;                                              \
        default: return fail(CMH_ERR_UNSUPPORTED, "width %d is not one of 128,256,384,512,768,1024", int(D)); \
    }

}  // namespace

int layernorm(const float* x, int64_t rows, int D, int64_t row_mul, const int32_t* row_idx, const float* g, const float* b,
              float eps, void* out, bool out_f32, cudaStream_t st) {
    CMH_REQUIRE(x && g && b && out && rows > 0, "layernorm: bad arguments");
    CMH_REQUIRE(D % 128 == 0, "layernorm: width %d is not a multiple of 128", D);
    if (out_f32) {
        CMH_DISPATCH_NV(D, return (launch_ln<NV, true>(x, rows, row_mul, row_idx, g, b, eps, out, st)));
    } else {
        CMH_DISPATCH_NV(D, return (launch_ln<NV, false>(x, rows, row_mul, row_idx, g, b, eps, out, st)));
    }
    return CMH_OK;
}

int vit_assemble(const float* emb, const float* cls, const float* pos, int64_t B, int L, int D, const float* g, const float* b,
                 float eps, float* x, cudaStream_t st) {
    const int64_t rows = B * L;
    const unsigned grid = unsigned(ceil_div(rows, ROWS_PER_BLOCK));
    CMH_DISPATCH_NV(D, CMH_CUDA_TRY(launch_kernel(vit_assemble_kernel<NV>, dim3(grid), dim3(ROWS_PER_BLOCK * 32), 0, st, 1, emb, cls,
                                                  pos, rows, L, g, b, eps, x)));
    return CMH_OK;
}

int text_embed(const int64_t* text, const float* tok, const float* pos, int64_t B, int L, int D, int vocab, float* x,
               cudaStream_t st) {
    const int64_t rows = B * L;
    const unsigned grid = unsigned(ceil_div(rows, ROWS_PER_BLOCK));
    CMH_DISPATCH_NV(D, CMH_CUDA_TRY(launch_kernel(text_embed_kernel<NV>, dim3(grid), dim3(ROWS_PER_BLOCK * 32), 0, st, 1, text, tok,
                                                  pos, rows, L, vocab, x)));
    return CMH_OK;
}

int text_eos(const int64_t* text, const uint8_t* pad, int64_t B, int L, int64_t eot_id, int32_t* eos, uint8_t* new_mask,
             cudaStream_t st) {
    CMH_CUDA_TRY(launch_kernel(text_eos_kernel, dim3(unsigned(ceil_div(B, 128))), dim3(128), 0, st, 1, text, pad, B, L, eot_id, eos,
                               new_mask));
    return CMH_OK;
}

int patchify(const float* img, int64_t B, int C, int R, int P, void* out, cudaStream_t st) {
    CMH_REQUIRE(R % P == 0 && P % 8 == 0, "patchify: resolution %d / patch %d unsupported", R, P);
    CMH_REQUIRE((reinterpret_cast<uintptr_t>(img) & 15) == 0, "patchify: images must be 16-byte aligned");
    const int64_t total = B * C * R * (R / 8);
    const int64_t blocks = ceil_div(total, 256);
    const int64_t cap = int64_t(sm_count_cached()) * 32;
    CMH_CUDA_TRY(launch_kernel(patchify_kernel, dim3(unsigned(blocks < cap ? blocks : cap)), dim3(256), 0, st, 1, img, B, C, R, P,
                               static_cast<__nv_bfloat16*>(out)));
    return CMH_OK;
}

int patchify_u8(const uint8_t* img, int64_t B, int C, int R, int P, const float* mean, const float* std, void* out, cudaStream_t st) {
    CMH_REQUIRE(C == 3 && R % P == 0 && P % 16 == 0 && R % 16 == 0, "patchify_u8: %d channels / resolution %d / patch %d unsupported", C, R, P);
    CMH_REQUIRE((reinterpret_cast<uintptr_t>(img) & 15) == 0, "patchify_u8: images must be 16-byte aligned");
    CMH_REQUIRE(mean && std && std[0] != 0.f && std[1] != 0.f && std[2] != 0.f, "patchify_u8: mean / std");
    // ((x / 255) - mean) / std  ==  x * scale + shift
    const float3 scale = make_float3(1.f / (255.f * std[0]), 1.f / (255.f * std[1]), 1.f / (255.f * std[2]));
    const float3 shift = make_float3(-mean[0] / std[0], -mean[1] / std[1], -mean[2] / std[2]);
    const int64_t total = B * C * R * (R / 16);
    const int64_t blocks = ceil_div(total, 256);
    const int64_t cap = int64_t(sm_count_cached()) * 32;
    CMH_CUDA_TRY(launch_kernel(patchify_u8_kernel, dim3(unsigned(blocks < cap ? blocks : cap)), dim3(256), 0, st, 1, img, B, C, R, P, scale,
                               shift, static_cast<__nv_bfloat16*>(out)));
    return CMH_OK;
}

}  // namespace cmh

extern "C" int cmh_layernorm(const float* x, int64_t rows, int width, const float* gain, const float* bias, float eps,
                             void* out, int out_f32, void* stream) {
    return cmh::layernorm(x, rows, width, 1, nullptr, gain, bias, eps, out, out_f32 != 0, cmh::as_stream(stream));
}
