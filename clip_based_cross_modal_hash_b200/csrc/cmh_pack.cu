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

// Vector path (16-byte aligned rows, ncols a multiple of the 16-byte vector): a block stages whole rows.
//   load   every thread issues 8 independent, fully coalesced 16-byte loads (a "slot" = one vector of V elements) before it
//          touches any of them: 32 KB in flight per block, enough memory-level parallelism to stream at HBM speed;
//   stage  the V sign / non-zero bits of a slot go to shared memory (one uint16 per slot);
//   build  one thread per output word ORs the 32/V slots of its word together -> coalesced word stores.
// Replaces the one-element-per-lane ballot loop for the bulk inputs (C4: 256 MB of +-1 fp32 gallery codes, C2: 75 MB of
// int64 labels).
template <class T>
struct VecBits;  // bits of the V = 16/sizeof(T) elements of one 16-byte vector + number of "bad" elements
template <>
struct VecBits<float> {
    static constexpr int V = 4;
    template <bool IS_CODE>
    static __device__ __forceinline__ void eval(const uint4& v, uint32_t& bits, unsigned& bad) {
        const float f[4] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (IS_CODE) {
                bits |= uint32_t(f[j] > 0.0f) << j;
                bad += (f[j] != 1.0f && f[j] != -1.0f);
            } else {
                bits |= uint32_t(f[j] != 0.0f) << j;
                bad += (f[j] != 0.0f && f[j] != 1.0f);
            }
        }
    }
};
template <>
struct VecBits<int32_t> {
    static constexpr int V = 4;
    template <bool IS_CODE>
    static __device__ __forceinline__ void eval(const uint4& v, uint32_t& bits, unsigned& bad) {
        const uint32_t e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) bits |= uint32_t(e[j] != 0u) << j, bad += e[j] > 1u;
    }
};
template <>
struct VecBits<int64_t> {
    static constexpr int V = 2;
    template <bool IS_CODE>
    static __device__ __forceinline__ void eval(const uint4& v, uint32_t& bits, unsigned& bad) {
        bits |= uint32_t((v.x | v.y) != 0u) | (uint32_t((v.z | v.w) != 0u) << 1);
        bad += (v.y != 0u || v.x > 1u) + (v.w != 0u || v.z > 1u);
    }
};
template <>
struct VecBits<uint8_t> {
    static constexpr int V = 16;
    template <bool IS_CODE>
    static __device__ __forceinline__ void eval(const uint4& v, uint32_t& bits, unsigned& bad) {
        const uint32_t e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const uint32_t x = (e[j] >> (8 * b)) & 0xFFu;
                bits |= uint32_t(x != 0u) << (4 * j + b);
                bad += x > 1u;
            }
    }
};

constexpr int PACK_THREADS = 256;
constexpr int PACK_LOADS = 8;                          // 16-byte loads in flight per thread
constexpr int PACK_SLOTS = PACK_THREADS * PACK_LOADS;  // slots staged per block iteration

template <class T, bool IS_CODE>
__global__ void __launch_bounds__(PACK_THREADS) pack_vec_kernel(const T* __restrict__ in, int64_t n, int ncols, int64_t ld,
                                                                int words_dst, uint32_t* __restrict__ out,
                                                                unsigned long long* __restrict__ bad_count) {
    constexpr int V = VecBits<T>::V;
    constexpr int SPW = 32 / V;  // slots per output word
    __shared__ uint16_t nib[PACK_SLOTS];
    const int tid = threadIdx.x;
    const int spr = ncols / V;  // slots per row (ncols % V == 0, checked on the host)
    const int rows_per_tile = PACK_SLOTS / spr;
    const int words_src = (ncols + 31) / 32;
    const int wshift = words_dst == 1 ? 0 : words_dst == 2 ? 1 : 2;
    unsigned bad = 0;
    for (int64_t row0 = int64_t(blockIdx.x) * rows_per_tile; row0 < n; row0 += int64_t(gridDim.x) * rows_per_tile) {
        const int rows = n - row0 < rows_per_tile ? int(n - row0) : rows_per_tile;
        const int nslots = rows * spr;
        uint4 v[PACK_LOADS];
#pragma unroll
        for (int i = 0; i < PACK_LOADS; ++i) {
            const int li = i * PACK_THREADS + tid;
            v[i] = make_uint4(0u, 0u, 0u, 0u);
            if (li < nslots) {
                const int r = li / spr, s = li - r * spr;
                v[i] = __ldcs(reinterpret_cast<const uint4*>(in + (row0 + r) * ld + int64_t(s) * V));  // streamed once
            }
        }
#pragma unroll
        for (int i = 0; i < PACK_LOADS; ++i) {
            const int li = i * PACK_THREADS + tid;
            uint32_t bits = 0;
            unsigned b = 0;
            VecBits<T>::template eval<IS_CODE>(v[i], bits, b);
            if (li < nslots) nib[li] = uint16_t(bits), bad += b;
        }
        __syncthreads();
        const int nwords = rows << wshift;
        for (int wi = tid; wi < nwords; wi += PACK_THREADS) {
            const int r = wi >> wshift, w = wi & (words_dst - 1);
            uint32_t word = 0;
            if (w < words_src) {
                const int first = w * SPW;
                const int cnt = spr - first < SPW ? spr - first : SPW;
                const uint16_t* src = nib + r * spr + first;
#pragma unroll
                for (int j = 0; j < SPW; ++j)
                    if (j < cnt) word |= uint32_t(src[j]) << (j * V);
            }
            out[((row0 + r) << wshift) + w] = word;
        }
        __syncthreads();
    }
    for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xFFFFFFFFu, bad, o);
    if (bad_count && (tid & 31) == 0 && bad) atomicAdd(bad_count, (unsigned long long)bad);
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
    constexpr int V = VecBits<T>::V;
    if (ncols % V == 0 && ld % V == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && ncols / V <= PACK_SLOTS) {
        const int rows_per_tile = PACK_SLOTS / (ncols / V);
        int64_t blocks = ceil_div(n, rows_per_tile);
        const int64_t cap = int64_t(sm_count_cached()) * 8;
        if (blocks > cap) blocks = cap;
        pack_vec_kernel<T, IS_CODE><<<unsigned(blocks), PACK_THREADS, 0, st>>>(in, n, ncols, ld, words_dst, out, bad);
        CMH_LAUNCH_CHECK("pack_vec_kernel");
        return CMH_OK;
    }
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
