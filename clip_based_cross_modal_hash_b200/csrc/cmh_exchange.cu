// cmh_exchange.cu — the one exchange step of the sharded top-k through NVSwitch multicast memory.
//
// After cand_place_kernel / rank_topk every rank holds a [Q][k] key buffer in which exactly the slots it owns are filled and all
// others are EMPTY (-1 as int64, below every real key).  The buffers live in symmetric memory that is also mapped at ONE multicast
// address: a multimem.ld_reduce on that address makes the switch fetch the slot from every rank and return the maximum, a
// multimem.st broadcasts a value into the slot of every rank.  Each rank reduces 1/world of the buffer and broadcasts it: 2 x 1/world
// of the buffer crosses its NVLink ports instead of the 2 x (world-1)/world of a ring all-reduce, in one kernel, with coalesced
// 8-byte accesses (scattered per-key remote stores from the place kernel were measured 4-8x slower, profiles/README.md).
#include "cmh_common.cuh"

namespace cmh {
namespace {

// Four independent switch reductions in flight per thread, then the four broadcasts (a reduction is a round trip through the
// NVSwitch: issuing them one at a time left the kernel latency-bound at a quarter of the link rate).
constexpr int NVLS_ILP = 4;
__global__ void __launch_bounds__(256) nvls_allreduce_max_kernel(long long* mc, int64_t begin, int64_t end) {
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i0 = begin + int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i0 < end; i0 += stride * NVLS_ILP) {
        long long v[NVLS_ILP];
#pragma unroll
        for (int u = 0; u < NVLS_ILP; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < end) asm volatile("multimem.ld_reduce.relaxed.sys.global.max.s64 %0, [%1];" : "=l"(v[u]) : "l"(mc + i));
        }
#pragma unroll
        for (int u = 0; u < NVLS_ILP; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < end) asm volatile("multimem.st.relaxed.sys.global.s64 [%0], %1;" ::"l"(mc + i), "l"(v[u]) : "memory");
        }
    }
}

// Broadcast of a rank's own block into the same place of EVERY rank's symmetric buffer: plain 16-byte loads of the local source,
// 16-byte multicast stores (fire and forget — no round trip through the switch).  Used for the small exchanges of the sharded
// top-k (sample histograms, per-distance totals: 2.7 MB per rank), where an NCCL collective is launch- and latency-bound.
__global__ void __launch_bounds__(256) nvls_broadcast_kernel(const float4* __restrict__ src, float4* mc_dst, int64_t n16) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += int64_t(gridDim.x) * blockDim.x) {
        const float4 v = __ldg(src + i);
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
    }
}

// Result exchange without a reduction: every slot of the [Q][k] key buffer has exactly one owner, so each rank pushes the slots
// it owns (everything that is not EMPTY in its local placement buffer) to all ranks with multicast stores; coalesced over slots.
__global__ void __launch_bounds__(256) nvls_push_owned_kernel(const long long* __restrict__ local, long long* mc, int64_t n) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        const long long v = __ldg(local + i);
        if (v != -1) asm volatile("multimem.st.relaxed.sys.global.s64 [%0], %1;" ::"l"(mc + i), "l"(v) : "memory");
    }
}

}  // namespace
}  // namespace cmh

using namespace cmh;

extern "C" int cmh_nvls_broadcast(const void* src, void* multicast_dst, int64_t bytes, void* stream) {
    CMH_REQUIRE(src && multicast_dst && bytes >= 0 && bytes % 16 == 0, "nvls_broadcast: bytes must be a multiple of 16");
    CMH_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(multicast_dst)) & 15) == 0, "nvls_broadcast: 16-byte alignment");
    if (bytes == 0) return CMH_OK;
    const int64_t n16 = bytes / 16;
    int64_t blocks = ceil_div(n16, 256);
    const int64_t cap = int64_t(sm_count_cached()) * 4;
    if (blocks > cap) blocks = cap;
    nvls_broadcast_kernel<<<unsigned(blocks), 256, 0, as_stream(stream)>>>(static_cast<const float4*>(src), static_cast<float4*>(multicast_dst), n16);
    CMH_LAUNCH_CHECK("nvls_broadcast_kernel");
    return CMH_OK;
}

extern "C" int cmh_nvls_push_owned_s64(const void* local_keys, void* multicast_keys, int64_t count, void* stream) {
    CMH_REQUIRE(local_keys && multicast_keys && count >= 0, "nvls_push_owned: bad arguments");
    if (count == 0) return CMH_OK;
    int64_t blocks = ceil_div(count, 256 * 4);
    const int64_t cap = int64_t(sm_count_cached()) * 8;
    if (blocks > cap) blocks = cap;
    nvls_push_owned_kernel<<<unsigned(blocks), 256, 0, as_stream(stream)>>>(static_cast<const long long*>(local_keys),
                                                                            static_cast<long long*>(multicast_keys), count);
    CMH_LAUNCH_CHECK("nvls_push_owned_kernel");
    return CMH_OK;
}

extern "C" int cmh_nvls_allreduce_max_s64(void* multicast_ptr, int64_t count, int rank, int world, void* stream) {
    CMH_REQUIRE(multicast_ptr && count >= 0 && world >= 1 && rank >= 0 && rank < world, "nvls_allreduce_max: bad arguments");
    const int64_t per = ceil_div(count, world);
    const int64_t begin = per * rank < count ? per * rank : count;
    const int64_t end = begin + per < count ? begin + per : count;
    if (end <= begin) return CMH_OK;
    int64_t blocks = ceil_div(end - begin, 256 * NVLS_ILP);
    const int64_t cap = int64_t(sm_count_cached()) * 8;
    if (blocks > cap) blocks = cap;
    nvls_allreduce_max_kernel<<<unsigned(blocks), 256, 0, as_stream(stream)>>>(static_cast<long long*>(multicast_ptr), begin, end);
    CMH_LAUNCH_CHECK("nvls_allreduce_max_kernel");
    return CMH_OK;
}
