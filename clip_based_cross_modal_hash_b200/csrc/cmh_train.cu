// cmh_train.cu — the tail of the DSPH training step (BASELINE.json config C5; runners/DSPH/runner.py:104-127):
//   HyP objective forward + backward                    models/DSPH/loss/HyP.py:18-69 (autograd in the reference)
//   tanh(Linear) hash head backward                     models/DSPH/hash/hash.py:6-15 (evaluation-mode head: dropout off)
//   fused multi-tensor BertAdam step                    models/common/optimizer.py:102-165 (a per-tensor Python loop there)
//   SGD with momentum for the HyP proxies               torch.optim.SGD at runners/DSPH/runner.py:86-89
// The backward pass through the CLIP towers is NOT here: this is what a step needs when the backbone is frozen,
// and the optimiser kernels work on any list of tensors (optim.FusedBertAdam is a drop-in for the reference's BertAdam).
// All arithmetic fp32, in the reference's order of operations where it matters (clip -> moments -> update -> decay -> lr).
#include "cmh_common.cuh"
#include "cmh_encoder.h"

namespace cmh {
namespace {

// ---- HyP backward --------------------------------------------------------------------------------------------------------
// Forward (cmh_hyp_loss_f32) leaves in the workspace: acc[4] = P_num, acc[5] = N_num, acc[9] = number of disjoint pairs, and the
// normalised rows xn, yn, pn.  With g = dL/dcos and c = cos(a, b) = an . bn:   d c / d a = (bn - c an) / ||a||.
__device__ __forceinline__ bool label_on(const uint32_t* lab, int LW, int row, int c) {
    return (lab[size_t(row) * LW + (c >> 5)] >> (c & 31)) & 1u;
}

// one block per sample i: gradient w.r.t. x_i and y_i
__global__ void __launch_bounds__(128) hyp_grad_feat_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            const float* __restrict__ xn, const float* __restrict__ yn,
                                                            const float* __restrict__ pn, const uint32_t* __restrict__ lab, int LW,
                                                            int B, int K, int C, float thr, float alpha,
                                                            const double* __restrict__ acc, float* __restrict__ dx,
                                                            float* __restrict__ dy) {
    extern __shared__ float sh[];
    float* gx = sh;            // [C] dL/dcos(x_i, p_c)
    float* gy = gx + C;        // [C]
    float* cx = gy + C;        // [C] cos values
    float* cy = cx + C;        // [C]
    float* wxx = cy + C;       // [B] coefficient of (xn_j - s_ij xn_i) in dL/dxn_i, from the x-x regulariser (both orders)
    float* sxx = wxx + B;      // [B] s_ij = xn_i . xn_j
    float* wyy = sxx + B;      // [B]
    float* syy = wyy + B;      // [B]
    float* wxy = syy + B;      // [B] x-t regulariser, pair (i, j): x side = i
    float* sxy = wxy + B;      // [B] xn_i . yn_j
    float* wyx = sxy + B;      // [B] x-t regulariser, pair (j, i): t side = i
    float* syx = wyx + B;      // [B] xn_j . yn_i
    __shared__ float red[8];
    const int i = blockIdx.x, tid = threadIdx.x;
    const float inv_p = 1.f / float(acc[4]), inv_n = 1.f / float(acc[5]);
    const float pairs = float(acc[9]);
    const float reg = (alpha > 0.f && pairs > 0.f) ? alpha / pairs : 0.f;
    int ci = 0;
    for (int w = 0; w < LW; ++w) ci += __popc(lab[size_t(i) * LW + w]);
    for (int c = tid; c < C; c += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int k = 0; k < K; ++k) {
            const float p = pn[size_t(c) * K + k];
            a = fmaf(xn[size_t(i) * K + k], p, a);
            b = fmaf(yn[size_t(i) * K + k], p, b);
        }
        const bool on = label_on(lab, LW, i, c);
        cx[c] = a, cy[c] = b;
        gx[c] = on ? -inv_p : (a > thr ? inv_n : 0.f);
        gy[c] = on ? -inv_p : (b > thr ? inv_n : 0.f);
    }
    for (int j = tid; j < B; j += blockDim.x) {
        int cj = 0;
        uint32_t both = 0;
        for (int w = 0; w < LW; ++w) {
            const uint32_t a = lab[size_t(i) * LW + w], b = lab[size_t(j) * LW + w];
            cj += __popc(b), both |= a & b;
        }
        float xx = 0.f, yy = 0.f, xy = 0.f, yx = 0.f;
        const bool pair = reg > 0.f && ci > 1 && cj > 1 && both == 0;
        if (pair) {
            for (int k = 0; k < K; ++k) {
                const float xi = xn[size_t(i) * K + k], xj = xn[size_t(j) * K + k];
                const float yi = yn[size_t(i) * K + k], yj = yn[size_t(j) * K + k];
                xx = fmaf(xi, xj, xx), yy = fmaf(yi, yj, yy), xy = fmaf(xi, yj, xy), yx = fmaf(xj, yi, yx);
            }
        }
        sxx[j] = xx, syy[j] = yy, sxy[j] = xy, syx[j] = yx;
        wxx[j] = (pair && xx > thr) ? 2.f * reg : 0.f;   // s_ij appears in the terms (i, j) and (j, i)
        wyy[j] = (pair && yy > thr) ? 2.f * reg : 0.f;
        wxy[j] = (pair && xy > thr) ? reg : 0.f;
        wyx[j] = (pair && yx > thr) ? reg : 0.f;
    }
    __syncthreads();
    // squared norms of x_i, y_i
    float qx = 0.f, qy = 0.f;
    for (int k = tid; k < K; k += blockDim.x) qx += x[size_t(i) * K + k] * x[size_t(i) * K + k], qy += y[size_t(i) * K + k] * y[size_t(i) * K + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qx += __shfl_xor_sync(0xffffffffu, qx, o), qy += __shfl_xor_sync(0xffffffffu, qy, o);
    if ((tid & 31) == 0) red[tid >> 5] = qx, red[4 + (tid >> 5)] = qy;
    __syncthreads();
    const int nw = blockDim.x >> 5;
    float nx = 0.f, ny = 0.f;
    for (int w = 0; w < nw; ++w) nx += red[w], ny += red[4 + w];
    const float inx = 1.f / fmaxf(sqrtf(nx), 1e-12f), iny = 1.f / fmaxf(sqrtf(ny), 1e-12f);
    for (int k = tid; k < K; k += blockDim.x) {
        const float xi = xn[size_t(i) * K + k], yi = yn[size_t(i) * K + k];
        float ax = 0.f, ay = 0.f;
        for (int c = 0; c < C; ++c) {
            const float p = pn[size_t(c) * K + k];
            ax = fmaf(gx[c], p - cx[c] * xi, ax);
            ay = fmaf(gy[c], p - cy[c] * yi, ay);
        }
        if (reg > 0.f) {
            for (int j = 0; j < B; ++j) {
                const float xj = xn[size_t(j) * K + k], yj = yn[size_t(j) * K + k];
                ax = fmaf(wxx[j], xj - sxx[j] * xi, ax);
                ax = fmaf(wxy[j], yj - sxy[j] * xi, ax);
                ay = fmaf(wyy[j], yj - syy[j] * yi, ay);
                ay = fmaf(wyx[j], xj - syx[j] * yi, ay);
            }
        }
        dx[size_t(i) * K + k] = ax * inx;
        dy[size_t(i) * K + k] = ay * iny;
    }
}

// one block per class c: gradient w.r.t. proxy p_c
__global__ void __launch_bounds__(128) hyp_grad_proxy_kernel(const float* __restrict__ proxies, const float* __restrict__ xn,
                                                             const float* __restrict__ yn, const float* __restrict__ pn,
                                                             const uint32_t* __restrict__ lab, int LW, int B, int K, int C, float thr,
                                                             const double* __restrict__ acc, float* __restrict__ dp) {
    extern __shared__ float sh[];
    float* gx = sh;       // [B]
    float* gy = gx + B;
    float* cx = gy + B;
    float* cy = cx + B;
    __shared__ float red[4];
    const int c = blockIdx.x, tid = threadIdx.x;
    const float inv_p = 1.f / float(acc[4]), inv_n = 1.f / float(acc[5]);
    for (int i = tid; i < B; i += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int k = 0; k < K; ++k) {
            const float p = pn[size_t(c) * K + k];
            a = fmaf(xn[size_t(i) * K + k], p, a);
            b = fmaf(yn[size_t(i) * K + k], p, b);
        }
        const bool on = label_on(lab, LW, i, c);
        cx[i] = a, cy[i] = b;
        gx[i] = on ? -inv_p : (a > thr ? inv_n : 0.f);
        gy[i] = on ? -inv_p : (b > thr ? inv_n : 0.f);
    }
    float q = 0.f;
    for (int k = tid; k < K; k += blockDim.x) q += proxies[size_t(c) * K + k] * proxies[size_t(c) * K + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if ((tid & 31) == 0) red[tid >> 5] = q;
    __syncthreads();
    float n2 = 0.f;
    for (int w = 0; w < int(blockDim.x >> 5); ++w) n2 += red[w];
    const float inp = 1.f / fmaxf(sqrtf(n2), 1e-12f);
    for (int k = tid; k < K; k += blockDim.x) {
        const float p = pn[size_t(c) * K + k];
        float a = 0.f;
        for (int i = 0; i < B; ++i) {
            a = fmaf(gx[i], xn[size_t(i) * K + k] - cx[i] * p, a);
            a = fmaf(gy[i], yn[size_t(i) * K + k] - cy[i] * p, a);
        }
        dp[size_t(c) * K + k] = a * inp;
    }
}

// ---- tanh(Linear) backward -------------------------------------------------------------------------------------------------
// y = tanh(feat . W^T + b);  dz = dy * (1 - y^2);  dW = dz^T . feat;  db = sum_b dz;  dfeat = dz . W
__global__ void __launch_bounds__(256) head_dz_kernel(const float* __restrict__ y, const float* __restrict__ dy, int64_t n, float* __restrict__ dz) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) dz[i] = dy[i] * (1.f - y[i] * y[i]);
}
__global__ void __launch_bounds__(256) head_dw_kernel(const float* __restrict__ dz, const float* __restrict__ feat, int B, int K, int D,
                                                      float* __restrict__ dW, float* __restrict__ db) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < K * D) {
        const int k = idx / D, d = idx % D;
        float a = 0.f;
        for (int b = 0; b < B; ++b) a = fmaf(dz[size_t(b) * K + k], feat[size_t(b) * D + d], a);
        dW[idx] = a;
    }
    if (idx < K && db) {
        float a = 0.f;
        for (int b = 0; b < B; ++b) a += dz[size_t(b) * K + idx];
        db[idx] = a;
    }
}
__global__ void __launch_bounds__(256) head_dfeat_kernel(const float* __restrict__ dz, const float* __restrict__ W, int B, int K, int D,
                                                         float* __restrict__ dfeat) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * D) return;
    const int b = idx / D, d = idx % D;
    float a = 0.f;
    for (int k = 0; k < K; ++k) a = fmaf(dz[size_t(b) * K + k], W[size_t(k) * D + d], a);
    dfeat[idx] = a;
}

// ---- fused multi-tensor optimiser steps --------------------------------------------------------------------------------------
constexpr int OPT_CHUNK = 16384;  // elements per block

// sum of squares of every gradient tensor (clip_grad_norm_ is applied per tensor, optimizer.py:138-139)
__global__ void __launch_bounds__(256) opt_norm_kernel(const cmh_opt_tensor* __restrict__ tensors, const int32_t* __restrict__ block_tensor,
                                                       const int32_t* __restrict__ block_chunk, float* __restrict__ sumsq) {
    __shared__ float red[8];
    const int t = block_tensor[blockIdx.x];
    const cmh_opt_tensor T = tensors[t];
    const int64_t lo = int64_t(block_chunk[blockIdx.x]) * OPT_CHUNK, hi = lo + OPT_CHUNK < T.n ? lo + OPT_CHUNK : T.n;
    float s = 0.f;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) s = fmaf(T.grad[i], T.grad[i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int w = 0; w < 8; ++w) tot += red[w];
        atomicAdd(sumsq + t, tot);
    }
}

// BertAdam.step (optimizer.py:130-165) for every tensor at once: clip (in place, like clip_grad_norm_) -> next_m, next_v ->
// update = m / (sqrt(v) + e) + weight_decay * p -> p -= lr_t * update.  No bias correction (BERT's Adam).
__global__ void __launch_bounds__(256) bert_adam_kernel(const cmh_opt_tensor* __restrict__ tensors, const int32_t* __restrict__ block_tensor,
                                                        const int32_t* __restrict__ block_chunk, const float* __restrict__ sumsq,
                                                        float b1, float b2, float e, float max_grad_norm) {
    const int t = block_tensor[blockIdx.x];
    const cmh_opt_tensor T = tensors[t];
    const int64_t lo = int64_t(block_chunk[blockIdx.x]) * OPT_CHUNK, hi = lo + OPT_CHUNK < T.n ? lo + OPT_CHUNK : T.n;
    float coef = 1.f;
    if (max_grad_norm > 0.f) {  // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
        coef = max_grad_norm / (sqrtf(sumsq[t]) + 1e-6f);
        coef = coef < 1.f ? coef : 1.f;
    }
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const float g = T.grad[i] * coef;
        const float m = T.m[i] * b1 + g * (1.f - b1);
        const float v = T.v[i] * b2 + g * g * (1.f - b2);
        float u = m / (sqrtf(v) + e);
        const float p = T.param[i];
        if (T.weight_decay > 0.f) u += T.weight_decay * p;
        T.grad[i] = g;
        T.m[i] = m, T.v[i] = v;
        T.param[i] = p - T.lr * u;
    }
}

// torch.optim.SGD with momentum (dampening 0, no nesterov): g += wd * p; buf = first ? g : mu * buf + g; p -= lr * buf
__global__ void __launch_bounds__(256) sgd_momentum_kernel(const cmh_opt_tensor* __restrict__ tensors, const int32_t* __restrict__ block_tensor,
                                                           const int32_t* __restrict__ block_chunk, float momentum, int first_step) {
    const int t = block_tensor[blockIdx.x];
    const cmh_opt_tensor T = tensors[t];
    const int64_t lo = int64_t(block_chunk[blockIdx.x]) * OPT_CHUNK, hi = lo + OPT_CHUNK < T.n ? lo + OPT_CHUNK : T.n;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const float p = T.param[i];
        float g = T.grad[i];
        if (T.weight_decay != 0.f) g = fmaf(T.weight_decay, p, g);
        const float buf = (first_step || momentum == 0.f) ? g : T.m[i] * momentum + g;
        if (momentum != 0.f) T.m[i] = buf;
        T.param[i] = p - T.lr * buf;
    }
}

}  // namespace
}  // namespace cmh

using namespace cmh;

extern "C" {

int cmh_hyp_loss_grad_f32(const float* x, const float* y, const uint32_t* labels_packed, const float* proxies, int64_t B, int nbits,
                          int ncls, float threshold, float alpha, void* workspace, size_t workspace_bytes, float* loss_out, float* dx,
                          float* dy, float* dproxies, void* stream) {
    CMH_REQUIRE(dx && dy && dproxies, "hyp_loss_grad: NULL gradient buffer");
    CMH_REQUIRE(nbits <= 1024 && ncls <= CMH_MAX_CLASSES && B <= 4096, "hyp_loss_grad: sizes outside the supported range");
    if (int rc = cmh_hyp_loss_f32(x, y, labels_packed, proxies, B, nbits, ncls, threshold, alpha, workspace, workspace_bytes, loss_out, stream))
        return rc;
    const int LW = cmh_label_words(ncls);
    cudaStream_t st = as_stream(stream);
    const double* acc = static_cast<const double*>(workspace);
    const float* xn = reinterpret_cast<const float*>(static_cast<const char*>(workspace) + round_up(10 * sizeof(double), 256));
    const float* yn = xn + size_t(B) * nbits;
    const float* pn = yn + size_t(B) * nbits;
    const size_t smem_f = (4 * size_t(ncls) + 8 * size_t(B)) * sizeof(float), smem_p = 4 * size_t(B) * sizeof(float);
    CMH_REQUIRE(smem_f <= 200 * 1024 && smem_p <= 200 * 1024, "hyp_loss_grad: batch too large for the shared-memory tables");
    if (smem_f > 48 * 1024) CMH_CUDA_TRY(cudaFuncSetAttribute(hyp_grad_feat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_f)));
    if (smem_p > 48 * 1024) CMH_CUDA_TRY(cudaFuncSetAttribute(hyp_grad_proxy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_p)));
    hyp_grad_feat_kernel<<<unsigned(B), 128, smem_f, st>>>(x, y, xn, yn, pn, labels_packed, LW, int(B), nbits, ncls, threshold, alpha, acc, dx, dy);
    CMH_LAUNCH_CHECK("hyp_grad_feat_kernel");
    hyp_grad_proxy_kernel<<<unsigned(ncls), 128, smem_p, st>>>(proxies, xn, yn, pn, labels_packed, LW, int(B), nbits, ncls, threshold, acc, dproxies);
    CMH_LAUNCH_CHECK("hyp_grad_proxy_kernel");
    return CMH_OK;
}

int cmh_linear_tanh_backward_f32(const float* feat, const float* y, const float* dy, const float* W, int64_t rows, int in_dim, int nbits,
                                 float* dz_scratch, float* dW, float* db, float* dfeat, void* stream) {
    CMH_REQUIRE(feat && y && dy && W && dz_scratch && dW && rows > 0 && in_dim > 0 && nbits > 0, "linear_tanh_backward: bad arguments");
    CMH_REQUIRE(rows * int64_t(in_dim) < (int64_t(1) << 31) && int64_t(nbits) * in_dim < (int64_t(1) << 31), "linear_tanh_backward: too large");
    cudaStream_t st = as_stream(stream);
    const int64_t n = rows * nbits;
    head_dz_kernel<<<unsigned(ceil_div(n, 256)), 256, 0, st>>>(y, dy, n, dz_scratch);
    CMH_LAUNCH_CHECK("head_dz_kernel");
    head_dw_kernel<<<unsigned(ceil_div(int64_t(nbits) * in_dim, 256)), 256, 0, st>>>(dz_scratch, feat, int(rows), nbits, in_dim, dW, db);
    CMH_LAUNCH_CHECK("head_dw_kernel");
    if (dfeat) {
        head_dfeat_kernel<<<unsigned(ceil_div(rows * in_dim, 256)), 256, 0, st>>>(dz_scratch, W, int(rows), nbits, in_dim, dfeat);
        CMH_LAUNCH_CHECK("head_dfeat_kernel");
    }
    return CMH_OK;
}

int cmh_opt_chunk_elems(void) { return OPT_CHUNK; }

int cmh_bert_adam_step(const cmh_opt_tensor* tensors_dev, int ntensors, const int32_t* block_tensor_dev, const int32_t* block_chunk_dev,
                       int nblocks, float* sumsq_dev, float b1, float b2, float e, float max_grad_norm, void* stream) {
    CMH_REQUIRE(tensors_dev && block_tensor_dev && block_chunk_dev && sumsq_dev && ntensors > 0 && nblocks > 0, "bert_adam_step: bad arguments");
    cudaStream_t st = as_stream(stream);
    if (max_grad_norm > 0.f) {
        CMH_CUDA_TRY(cudaMemsetAsync(sumsq_dev, 0, size_t(ntensors) * sizeof(float), st));
        opt_norm_kernel<<<unsigned(nblocks), 256, 0, st>>>(tensors_dev, block_tensor_dev, block_chunk_dev, sumsq_dev);
        CMH_LAUNCH_CHECK("opt_norm_kernel");
    }
    bert_adam_kernel<<<unsigned(nblocks), 256, 0, st>>>(tensors_dev, block_tensor_dev, block_chunk_dev, sumsq_dev, b1, b2, e, max_grad_norm);
    CMH_LAUNCH_CHECK("bert_adam_kernel");
    return CMH_OK;
}

int cmh_sgd_momentum_step(const cmh_opt_tensor* tensors_dev, int ntensors, const int32_t* block_tensor_dev, const int32_t* block_chunk_dev,
                          int nblocks, float momentum, int first_step, void* stream) {
    CMH_REQUIRE(tensors_dev && block_tensor_dev && block_chunk_dev && ntensors > 0 && nblocks > 0, "sgd_momentum_step: bad arguments");
    sgd_momentum_kernel<<<unsigned(nblocks), 256, 0, as_stream(stream)>>>(tensors_dev, block_tensor_dev, block_chunk_dev, momentum, first_step);
    CMH_LAUNCH_CHECK("sgd_momentum_kernel");
    return CMH_OK;
}

}  // extern "C"
