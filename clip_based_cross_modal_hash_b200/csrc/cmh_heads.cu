// cmh_heads.cu — the per-method hash heads on top of the 512-d CLIP features, in fp32.
//
//   DSPH   models/DSPH/hash/hash.py:6-15     tanh(Linear(512, K)(x))           code = sign     (runners/base.py:407-410)
//   DCMHT  models/DCMHT/hash/hash.py:35-46   MHA over a length-1 sequence == out_proj(v_proj(x)) -> BatchNorm1d(eval) |
//                                            LayerNorm -> Linear(512, 2K) -> ReLU -> softmax over (2j, 2j+1)
//                                            code bit j = argmax of the pair (runners/DCMHT/runner.py:83-95)
//
// The heads are <0.1 % of the encoder FLOPs (B x 512 x K); they stay in fp32 on the SIMT pipes so that the sign /
// argmax decision is taken on the same arithmetic as the reference (no extra bf16 rounding right before the bit).
#include "cmh_common.cuh"
#include "cmh_encoder.h"

namespace cmh {
namespace {

constexpr int LIN_NPW = 4;  // outputs per warp

// out[r][n] = act( (x[r] . W[n] + bias[n]) * scale[n] + shift[n] )        x [rows][K], W [N][K] (nn.Linear.weight)
// one warp = one row x LIN_NPW consecutive outputs; lanes split K with float4 loads (coalesced over x and W rows).
template <int ACT>
__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ x, int64_t rows, int K, const float* __restrict__ W, const float* __restrict__ bias,
                  int N, const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ out, int64_t ldo) {
    const int lane = threadIdx.x & 31;
    const int ngroups = (N + LIN_NPW - 1) / LIN_NPW;
    const int64_t item = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (item >= rows * ngroups) return;
    const int64_t r = item / ngroups;
    const int n0 = int(item % ngroups) * LIN_NPW;
    const float4* xr = reinterpret_cast<const float4*>(x + r * K);
    float acc[LIN_NPW];
#pragma unroll
    for (int j = 0; j < LIN_NPW; ++j) acc[j] = 0.f;
    for (int k4 = lane; k4 < K / 4; k4 += 32) {
        const float4 a = __ldg(xr + k4);
#pragma unroll
        for (int j = 0; j < LIN_NPW; ++j) {
            if (n0 + j < N) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(W + int64_t(n0 + j) * K) + k4);
                acc[j] = fmaf(a.x, w.x, fmaf(a.y, w.y, fmaf(a.z, w.z, fmaf(a.w, w.w, acc[j]))));
            }
        }
    }
#pragma unroll
    for (int j = 0; j < LIN_NPW; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    }
    if (lane < LIN_NPW && n0 + lane < N) {
        const int n = n0 + lane;
        float v = acc[0];
#pragma unroll
        for (int j = 1; j < LIN_NPW; ++j) v = lane == j ? acc[j] : v;
        if (bias) v += bias[n];
        if (scale) v = v * scale[n] + shift[n];
        if (ACT == CMH_ACT_TANH) v = tanhf(v);
        if (ACT == CMH_ACT_RELU) v = fmaxf(v, 0.f);
        out[r * ldo + n] = v;
    }
}

// probs[r][2j], probs[r][2j+1] = softmax(logits[r][2j], logits[r][2j+1])   (models/common/hash.py:20-31)
// packed bit j = 1 iff probs[2j+1] > probs[2j]                             (argmax with ties -> index 0 -> -1)
__global__ void pair_softmax_pack_kernel(const float* __restrict__ logits, int64_t rows, int nbits, float* __restrict__ probs,
                                         uint32_t* __restrict__ packed, int W) {
    const int64_t gw = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;  // one warp = 32 bits of one row
    const int lane = threadIdx.x & 31;
    if (gw >= rows * W) return;
    const int64_t r = gw / W;
    const int w = int(gw % W), j = w * 32 + lane;
    bool bit = false;
    if (j < nbits) {
        const float2 e = *reinterpret_cast<const float2*>(logits + r * 2 * nbits + 2 * j);
        const float m = fmaxf(e.x, e.y);
        const float a = expf(e.x - m), b = expf(e.y - m);
        const float s = a + b;
        const float p0 = a / s, p1 = b / s;
        if (probs) *reinterpret_cast<float2*>(probs + r * 2 * nbits + 2 * j) = make_float2(p0, p1);
        bit = p1 > p0;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, bit);
    if (lane == 0 && packed) packed[r * W + w] = word;
}

}  // namespace

int linear_f32(const float* x, int64_t rows, int K, const float* W, const float* bias, int N, const float* scale,
               const float* shift, int act, float* out, int64_t ldo, cudaStream_t st) {
    CMH_REQUIRE(x && W && out && rows > 0 && N > 0 && K > 0 && K % 4 == 0, "linear: bad arguments (K must be a multiple of 4)");
    CMH_REQUIRE((scale == nullptr) == (shift == nullptr), "linear: scale and shift come together");
    const int64_t items = rows * ceil_div(N, LIN_NPW);
    const unsigned grid = unsigned(ceil_div(items, 8));
    switch (act) {
        case CMH_ACT_NONE: linear_f32_kernel<CMH_ACT_NONE><<<grid, 256, 0, st>>>(x, rows, K, W, bias, N, scale, shift, out, ldo); break;
        case CMH_ACT_TANH: linear_f32_kernel<CMH_ACT_TANH><<<grid, 256, 0, st>>>(x, rows, K, W, bias, N, scale, shift, out, ldo); break;
        case CMH_ACT_RELU: linear_f32_kernel<CMH_ACT_RELU><<<grid, 256, 0, st>>>(x, rows, K, W, bias, N, scale, shift, out, ldo); break;
        default: return fail(CMH_ERR_INVALID, "linear: unknown activation %d", act);
    }
    CMH_LAUNCH_CHECK("linear_f32_kernel");
    return CMH_OK;
}

int pair_softmax_pack(const float* logits, int64_t rows, int nbits, float* probs, uint32_t* packed, cudaStream_t st) {
    const int W = cmh_code_words(nbits);
    CMH_REQUIRE(W > 0, "pair_softmax: %d bits unsupported", nbits);
    const int64_t threads = rows * W * 32;
    pair_softmax_pack_kernel<<<unsigned(ceil_div(threads, 256)), 256, 0, st>>>(logits, rows, nbits, probs, packed, W);
    CMH_LAUNCH_CHECK("pair_softmax_pack_kernel");
    return CMH_OK;
}

}  // namespace cmh

extern "C" {

int cmh_linear_f32(const float* x, int64_t rows, int in_dim, const float* weight, const float* bias, int out_dim,
                   const float* scale, const float* shift, int act, float* out, int64_t ldo, void* stream) {
    return cmh::linear_f32(x, rows, in_dim, weight, bias, out_dim, scale, shift, act, out, ldo, cmh::as_stream(stream));
}

int cmh_head_dsph(const float* feat, int64_t rows, int in_dim, const float* weight, const float* bias, int nbits, float* hash,
                  uint32_t* packed, void* stream) {
    CMH_REQUIRE(hash, "head_dsph: the tanh output buffer is required");
    if (int rc = cmh::linear_f32(feat, rows, in_dim, weight, bias, nbits, nullptr, nullptr, CMH_ACT_TANH, hash, nbits,
                                 cmh::as_stream(stream)))
        return rc;
    if (packed) return cmh_pack_codes_f32(hash, rows, nbits, nbits, packed, nullptr, stream);
    return CMH_OK;
}

int cmh_head_dcmht(const float* feat, int64_t rows, int in_dim, const cmh_dcmht_head* head, int nbits, float* scratch,
                   float* probs, uint32_t* packed, void* stream) {
    CMH_REQUIRE(feat && head && scratch && rows > 0, "head_dcmht: bad arguments");
    CMH_REQUIRE(head->w_v && head->w_out && head->w_fc2, "head_dcmht: missing weights");
    CMH_REQUIRE((head->bn_scale && head->bn_shift) || (head->norm_gain && head->norm_bias),
                "head_dcmht: needs either the folded BatchNorm affine (image) or the LayerNorm gain/bias (text)");
    cudaStream_t st = cmh::as_stream(stream);
    float* v = scratch;                       // [rows][in_dim]
    float* e = v + rows * in_dim;             // [rows][in_dim]
    float* l = e + rows * in_dim;             // [rows][2*nbits]
    if (int rc = cmh::linear_f32(feat, rows, in_dim, head->w_v, head->b_v, in_dim, nullptr, nullptr, CMH_ACT_NONE, v, in_dim, st))
        return rc;
    if (head->bn_scale) {  // image branch: BatchNorm1d in eval mode is a per-feature affine map (folded by the caller)
        if (int rc = cmh::linear_f32(v, rows, in_dim, head->w_out, head->b_out, in_dim, head->bn_scale, head->bn_shift,
                                     CMH_ACT_NONE, e, in_dim, st))
            return rc;
    } else {               // text branch: LayerNorm
        if (int rc = cmh::linear_f32(v, rows, in_dim, head->w_out, head->b_out, in_dim, nullptr, nullptr, CMH_ACT_NONE, e,
                                     in_dim, st))
            return rc;
        if (int rc = cmh::layernorm(e, rows, in_dim, 1, nullptr, head->norm_gain, head->norm_bias, head->eps, v, true, st))
            return rc;
        e = v;
    }
    if (int rc = cmh::linear_f32(e, rows, in_dim, head->w_fc2, head->b_fc2, 2 * nbits, nullptr, nullptr, CMH_ACT_RELU, l,
                                 2 * nbits, st))
        return rc;
    return cmh::pair_softmax_pack(l, rows, nbits, probs, packed, st);
}
}
