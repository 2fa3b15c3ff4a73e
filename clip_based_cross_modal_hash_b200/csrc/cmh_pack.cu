// cmh_pack.cu — R0: +-1 fp32 codes / multi-hot labels -> bit-packed words (warp ballot), and back.
//
// A warp turns 32 consecutive output words (32 x 32 input elements) into one coalesced 128-byte store:
// lane l reads element l of word j (coalesced 128-byte load), __ballot_sync gathers the 32 predicate
// bits, lane j keeps the ballot of word j.
#include "cmh_common.cuh"

namespace cmh {
namespace {

template <class T>
struct Elem;
template <> struct Elem<float>   { static __device__ bool set(float v) { return v > 0.0f; } static __device__ bool bad_code(float v) { return v != 1.0f && v != -1.0f; } static __device__ bool nz(float v) { return v != 0.0f; } static __device__ bool bad_label(float v) { return v != 0.0f && v != 1.0f; } };
template <> struct Elem<int64_t> { static __device__ bool nz(int64_t v) { return v != 0; } static __device__ bool bad_label(int64_t v) { return v != 0 && v != 1; } };
template <> struct Elem<int32_t> { static __device__ bool nz(int32_t v) { return v != 0; } static __device__ bool bad_label(int32_t v) { return v != 0 && v != 1; } };
template <> struct Elem<uint8_t> { static __device__ bool nz(uint8_t v) { return v != 0; } static __device__ bool bad_label(uint8_t v) { return v > 1; } };

// IS_CODE: bit = v > 0, bad = not +-1.  else: bit = v != 0, bad = not 0/1.
template <class T, bool IS_CODE>
__global__ void __launch_bounds__(256) pack_kernel(const T* __restrict__ in, int64_t n, int ncols, int64_t ld,
                                                   int words_src, int words_dst, uint32_t* __restrict__ out,
                                                   unsigned long long* __restrict__ bad_count) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int64_t total_words = n * words_dst;
    unsigned bad = 0;
    for (int64_t g0 = warp * 32; g0 < total_words; g0 += nwarps * 32) {
        uint32_t mine = 0;
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
            const int64_t g = g0 + j;
            bool bit = false, isbad = false;
            if (g < total_words) {
                const int64_t row = g / words_dst;
                const int w = int(g - row * words_dst);
                const int col = w * 32 + lane;
                if (w < words_src && col < ncols) {
                    const T v = in[row * ld + col];
                    if (IS_CODE) {
                        bit = Elem<float>::set(float(v));
                        isbad = Elem<float>::bad_code(float(v));
                    } else {
                        bit = Elem<T>::nz(v);
                        isbad = Elem<T>::bad_label(v);
                    }
                }
            }
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, bit);
            bad += __popc(__ballot_sync(0xFFFFFFFFu, isbad));
            if (lane == j) mine = b;
        }
        if (g0 + lane < total_words) out[g0 + lane] = mine;
    }
    if (bad_count && lane == 0 && bad) atomicAdd(bad_count, (unsigned long long)bad);
}

__global__ void __launch_bounds__(256) unpack_kernel(const uint32_t* __restrict__ packed, int64_t n, int nbits, int W,
                                                     float* __restrict__ out, int64_t ld) {
    const int64_t total = n * nbits;
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += int64_t(gridDim.x) * blockDim.x) {
        const int64_t row = e / nbits;
        const int col = int(e - row * nbits);
        const uint32_t w = __ldg(packed + row * W + (col >> 5));
        out[row * ld + col] = ((w >> (col & 31)) & 1u) ? 1.0f : -1.0f;
    }
}

template <class T, bool IS_CODE>
int launch_pack(const T* in, int64_t n, int ncols, int64_t ld, int words_dst, uint32_t* out,
                unsigned long long* bad, cudaStream_t st) {
    const int words_src = (ncols + 31) / 32;
    const int64_t total_words = n * words_dst;
    int64_t blocks = ceil_div(ceil_div(total_words, 32), 8);  // 8 warps per block, one 32-word group per warp pass
    const int64_t cap = int64_t(sm_count_cached()) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    pack_kernel<T, IS_CODE><<<unsigned(blocks), 256, 0, st>>>(in, n, ncols, ld, words_src, words_dst, out, bad);
    CMH_LAUNCH_CHECK("pack_kernel");
    return CMH_OK;
}

}  // namespace
}  // namespace cmh

using namespace cmh;

extern "C" {

int cmh_pack_codes_f32(const float* codes, int64_t n, int nbits, int64_t ld, uint32_t* out,
                       unsigned long long* bad_count, void* stream) {
    const int W = cmh_code_words(nbits);
    if (W < 0) return fail(CMH_ERR_UNSUPPORTED, "nbits=%d outside 1..%d", nbits, CMH_MAX_BITS);
    CMH_REQUIRE(n >= 0 && ld >= nbits, "bad sizes n=%lld ld=%lld", (long long)n, (long long)ld);
    if (n == 0) return CMH_OK;
    CMH_REQUIRE(codes && out, "NULL pointer");
    return launch_pack<float, true>(codes, n, nbits, ld, W, out, bad_count, as_stream(stream));
}

int cmh_pack_labels(const void* labels, int dtype, int64_t n, int ncls, int64_t ld, uint32_t* out,
                    unsigned long long* bad_count, void* stream) {
    const int LW = cmh_label_words(ncls);
    if (LW <= 0) return fail(CMH_ERR_UNSUPPORTED, "ncls=%d outside 1..%d", ncls, CMH_MAX_CLASSES);
    CMH_REQUIRE(n >= 0 && ld >= ncls, "bad sizes n=%lld ld=%lld", (long long)n, (long long)ld);
    if (n == 0) return CMH_OK;
    CMH_REQUIRE(labels && out, "NULL pointer");
    cudaStream_t st = as_stream(stream);
    switch (dtype) {
        case CMH_DT_I64: return launch_pack<int64_t, false>(static_cast<const int64_t*>(labels), n, ncls, ld, LW, out, bad_count, st);
        case CMH_DT_F32: return launch_pack<float, false>(static_cast<const float*>(labels), n, ncls, ld, LW, out, bad_count, st);
        case CMH_DT_U8: return launch_pack<uint8_t, false>(static_cast<const uint8_t*>(labels), n, ncls, ld, LW, out, bad_count, st);
        case CMH_DT_I32: return launch_pack<int32_t, false>(static_cast<const int32_t*>(labels), n, ncls, ld, LW, out, bad_count, st);
    }
    return fail(CMH_ERR_INVALID, "unknown label dtype %d", dtype);
}

int cmh_unpack_codes_f32(const uint32_t* packed, int64_t n, int nbits, float* out, int64_t ld, void* stream) {
    const int W = cmh_code_words(nbits);
    if (W < 0) return fail(CMH_ERR_UNSUPPORTED, "nbits=%d", nbits);
    CMH_REQUIRE(n >= 0 && ld >= nbits, "bad sizes");
    if (n == 0) return CMH_OK;
    CMH_REQUIRE(packed && out, "NULL pointer");
    int64_t blocks = ceil_div(n * nbits, 256);
    const int64_t cap = int64_t(sm_count_cached()) * 16;
    if (blocks > cap) blocks = cap;
    unpack_kernel<<<unsigned(blocks), 256, 0, as_stream(stream)>>>(packed, n, nbits, W, out, ld);
    CMH_LAUNCH_CHECK("unpack_kernel");
    return CMH_OK;
}
}
