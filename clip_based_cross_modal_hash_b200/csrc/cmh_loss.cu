// cmh_loss.cu — forward value of the DSPH objective, HyP.forward (models/DSPH/loss/HyP.py:18-69): the "pairwise cosine-sim
// loss" of BASELINE.json's C5 configuration.  Evaluation only (no gradient): the training step is outside this round's scope.
//
//   cos  = normalize(x) . normalize(proxies)^T                       [B][C]
//   loss = mean_{label=1}(1 - cos) + mean_{label=0} relu(cos - thr)   (image) + the same for y (text)
//        + alpha * mean over pairs (i, j) of multi-label rows with disjoint labels of relu(sim - thr),
//          for sim in { xn.xn^T, yn.yn^T, xn.yn^T }                   (0 when there is no such pair)
// Sums are accumulated in fp64 (block partials + atomics), so the result does not depend on the launch geometry beyond
// ~1e-15 relative; the reference sums in fp32.
#include "cmh_common.cuh"
#include "cmh_encoder.h"

namespace cmh {
namespace {

// acc: [0] pos_x [1] neg_x [2] pos_y [3] neg_y [4] P_num [5] N_num [6] reg_xx [7] reg_yy [8] reg_xy [9] zero_pairs
constexpr int NACC = 10;

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < int(blockDim.x >> 5); ++w) t += red[w];
    return t;  // valid in thread 0
}

// rows of `src` [n][K] -> F.normalize(src, dim=1) (x / max(||x||, 1e-12)); one warp per row
__global__ void __launch_bounds__(256) l2_normalize_kernel(const float* __restrict__ src, int n, int K, float* __restrict__ dst) {
    const int lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n) return;
    float q = 0.f;
    for (int k = lane; k < K; k += 32) q += src[size_t(r) * K + k] * src[size_t(r) * K + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float inv = 1.f / fmaxf(sqrtf(q), 1e-12f);
    for (int k = lane; k < K; k += 32) dst[size_t(r) * K + k] = src[size_t(r) * K + k] * inv;
}

// proxy terms: one thread per (sample, class)
__global__ void __launch_bounds__(256) hyp_proxy_kernel(const float* __restrict__ xn, const float* __restrict__ yn,
                                                        const float* __restrict__ pn, const uint32_t* __restrict__ lab, int LW,
                                                        int B, int K, int C, float thr, double* __restrict__ acc) {
    __shared__ double red[8];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double v[6] = {0, 0, 0, 0, 0, 0};
    if (i < B * C) {
        const int b = i / C, c = i % C;
        float cx = 0.f, cy = 0.f;
        for (int k = 0; k < K; ++k) {
            const float p = pn[size_t(c) * K + k];
            cx = fmaf(xn[size_t(b) * K + k], p, cx);
            cy = fmaf(yn[size_t(b) * K + k], p, cy);
        }
        const bool on = (lab[size_t(b) * LW + (c >> 5)] >> (c & 31)) & 1u;
        if (on) v[0] = 1.f - cx, v[2] = 1.f - cy, v[4] = 1.0;
        else v[1] = fmaxf(cx - thr, 0.f), v[3] = fmaxf(cy - thr, 0.f), v[5] = 1.0;
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const double t = block_sum(v[j], red);
        if (threadIdx.x == 0 && t != 0.0) atomicAdd(acc + j, t);
    }
}

// regulariser: one thread per ordered pair (i, j) of samples with more than one label each and no label in common
__global__ void __launch_bounds__(256) hyp_pair_kernel(const float* __restrict__ xn, const float* __restrict__ yn,
                                                       const uint32_t* __restrict__ lab, int LW, int B, int K, float thr,
                                                       float alpha, double* __restrict__ acc) {
    __shared__ double red[8];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    double v[4] = {0, 0, 0, 0};
    if (idx < B * B) {
        const int i = idx / B, j = idx % B;
        int ci = 0, cj = 0;
        uint32_t both = 0;
        for (int w = 0; w < LW; ++w) {
            const uint32_t a = lab[size_t(i) * LW + w], b = lab[size_t(j) * LW + w];
            ci += __popc(a), cj += __popc(b), both |= a & b;
        }
        if (ci > 1 && cj > 1 && both == 0) {
            float xx = 0.f, yy = 0.f, xy = 0.f;
            for (int k = 0; k < K; ++k) {
                const float xi = xn[size_t(i) * K + k], xj = xn[size_t(j) * K + k];
                const float yi = yn[size_t(i) * K + k], yj = yn[size_t(j) * K + k];
                xx = fmaf(xi, xj, xx), yy = fmaf(yi, yj, yy), xy = fmaf(xi, yj, xy);
            }
            v[0] = alpha * fmaxf(xx - thr, 0.f), v[1] = alpha * fmaxf(yy - thr, 0.f), v[2] = alpha * fmaxf(xy - thr, 0.f), v[3] = 1.0;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double t = block_sum(v[j], red);
        if (threadIdx.x == 0 && t != 0.0) atomicAdd(acc + 6 + j, t);
    }
}

__global__ void hyp_finish_kernel(const double* __restrict__ acc, float alpha, float* __restrict__ out) {
    // HyP.py:33-39: the four proxy terms (a missing class of terms divides by zero like the reference); :41-67 regulariser
    double loss = acc[0] / acc[4] + acc[1] / acc[5] + acc[2] / acc[4] + acc[3] / acc[5];
    if (alpha > 0.f && acc[9] > 0.0) loss += (acc[6] + acc[7] + acc[8]) / acc[9];
    out[0] = float(loss);
}

}  // namespace
}  // namespace cmh

extern "C" int cmh_hyp_loss_f32(const float* x, const float* y, const uint32_t* labels_packed, const float* proxies, int64_t B,
                                int nbits, int ncls, float threshold, float alpha, void* workspace, size_t workspace_bytes,
                                float* loss_out, void* stream) {
    using namespace cmh;
    CMH_REQUIRE(x && y && labels_packed && proxies && loss_out && B > 0 && nbits > 0 && ncls > 0 && B < 32768, "hyp_loss: bad arguments");
    const int LW = cmh_label_words(ncls);
    CMH_REQUIRE(LW > 0, "hyp_loss: %d classes unsupported", ncls);
    const size_t need = round_up(NACC * sizeof(double), 256) + (2 * size_t(B) + size_t(ncls)) * nbits * sizeof(float);
    if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255))
        return fail(CMH_ERR_WORKSPACE, "hyp_loss: workspace needs %zu bytes, 256-byte aligned", need);
    cudaStream_t st = as_stream(stream);
    double* acc = static_cast<double*>(workspace);
    float* xn = reinterpret_cast<float*>(static_cast<char*>(workspace) + round_up(NACC * sizeof(double), 256));
    float* yn = xn + size_t(B) * nbits;
    float* pn = yn + size_t(B) * nbits;
    CMH_CUDA_TRY(cudaMemsetAsync(acc, 0, NACC * sizeof(double), st));
    l2_normalize_kernel<<<unsigned(ceil_div(B, 8)), 256, 0, st>>>(x, int(B), nbits, xn);
    l2_normalize_kernel<<<unsigned(ceil_div(B, 8)), 256, 0, st>>>(y, int(B), nbits, yn);
    l2_normalize_kernel<<<unsigned(ceil_div(ncls, 8)), 256, 0, st>>>(proxies, ncls, nbits, pn);
    hyp_proxy_kernel<<<unsigned(ceil_div(B * ncls, 256)), 256, 0, st>>>(xn, yn, pn, labels_packed, LW, int(B), nbits, ncls, threshold, acc);
    if (alpha > 0.f)
        hyp_pair_kernel<<<unsigned(ceil_div(B * B, 256)), 256, 0, st>>>(xn, yn, labels_packed, LW, int(B), nbits, threshold, alpha, acc);
    hyp_finish_kernel<<<1, 1, 0, st>>>(acc, alpha, loss_out);
    CMH_LAUNCH_CHECK("hyp_loss kernels");
    return CMH_OK;
}
