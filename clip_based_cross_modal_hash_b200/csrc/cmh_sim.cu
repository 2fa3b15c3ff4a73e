// cmh_sim.cu — S: the small fp32 similarity helpers of common/calc_utils.py:8-49 (label / cosine / euclid).
// Batch-sized operands (B = 128, d <= a few hundred): one 16x16 output tile per block, operands staged
// through padded shared memory, row norms accumulated in the same sweep.
#include "cmh_common.cuh"

namespace cmh {
namespace {

enum SimOp { SIM_LABEL = 0, SIM_COSINE = 1, SIM_EUCLID = 2 };

template <int OP>
__global__ void __launch_bounds__(256) sim_kernel(const float* __restrict__ a, int64_t n, const float* __restrict__ b,
                                                  int64_t m, int d, float* __restrict__ out) {
    __shared__ float As[16][33];
    __shared__ float Bs[16][33];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t i = int64_t(blockIdx.y) * 16 + ty, j = int64_t(blockIdx.x) * 16 + tx;
    float dot = 0.f, na = 0.f, nb = 0.f, dist = 0.f;
    for (int c0 = 0; c0 < d; c0 += 32) {
        for (int e = threadIdx.x; e < 16 * 32; e += 256) {
            const int r = e >> 5, c = e & 31;
            const int64_t ia = int64_t(blockIdx.y) * 16 + r, jb = int64_t(blockIdx.x) * 16 + r;
            As[r][c] = (ia < n && c0 + c < d) ? a[ia * d + c0 + c] : 0.f;
            Bs[r][c] = (jb < m && c0 + c < d) ? b[jb * d + c0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const float x = As[ty][c], y = Bs[tx][c];
            if (OP == SIM_EUCLID) {
                const float t = x - y;
                dist = __fmaf_rn(t, t, dist);
            } else {
                dot = __fmaf_rn(x, y, dot);
                if (OP == SIM_COSINE) {
                    na = __fmaf_rn(x, x, na);
                    nb = __fmaf_rn(y, y, nb);
                }
            }
        }
        __syncthreads();
    }
    if (i < n && j < m) {
        float r;
        if (OP == SIM_LABEL) r = dot > 0.f ? 1.f : 0.f;
        else if (OP == SIM_COSINE) r = dot / (sqrtf(na) * sqrtf(nb));  // zero row -> 0/0 = nan, like the reference
        else r = sqrtf(dist);
        out[i * m + j] = r;
    }
}

template <int OP>
int launch_sim(const float* a, int64_t n, const float* b, int64_t m, int d, float* out, void* stream) {
    CMH_REQUIRE(n >= 0 && m >= 0 && d > 0, "bad sizes");
    if (n == 0 || m == 0) return CMH_OK;
    CMH_REQUIRE(a && b && out, "NULL pointer");
    CMH_REQUIRE(ceil_div(n, 16) <= 65535, "n too large");
    dim3 grid(unsigned(ceil_div(m, 16)), unsigned(ceil_div(n, 16)));
    sim_kernel<OP><<<grid, 256, 0, as_stream(stream)>>>(a, n, b, m, d, out);
    CMH_LAUNCH_CHECK("sim_kernel");
    return CMH_OK;
}

}  // namespace
}  // namespace cmh

using namespace cmh;

extern "C" {
int cmh_label_sim_f32(const float* a, int64_t n, const float* b, int64_t m, int d, float* out, void* stream) {
    return launch_sim<SIM_LABEL>(a, n, b, m, d, out, stream);
}
int cmh_cosine_sim_f32(const float* a, int64_t n, const float* b, int64_t m, int d, float* out, void* stream) {
    return launch_sim<SIM_COSINE>(a, n, b, m, d, out, stream);
}
int cmh_euclid_sim_f32(const float* a, int64_t n, const float* b, int64_t m, int d, float* out, void* stream) {
    return launch_sim<SIM_EUCLID>(a, n, b, m, d, out, stream);
}
}
