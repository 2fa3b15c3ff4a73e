// cmh_tcgen05.cuh — PTX wrappers for the 5th-generation tensor-core path (tcgen05.mma / TMEM / 2-D TMA), shared by the
// encoder GEMM (cmh_gemm.cu, kind::f16) and the retrieval evaluator (cmh_tc.cu, kind::i8).
#pragma once

#include <cuda.h>

#include "cmh_common.cuh"

namespace cmh {

// ---- PTX wrappers ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// cta_group::2 flavour: executed by both CTAs of the pair; the transaction bytes land on the LEADER's barrier
// (bit 24 of a shared::cluster address is the CTA rank inside the pair)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
// bring a tile into L2 ahead of its TMA load (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster.  Default (.release.cta) semantics:
// the accumulator hand-over is ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync, not by generic memory;
// a .release.cluster arrive compiled to MEMBAR.ALL.CTA + ERRBAR and cost ~1/3 of the epilogue (profiles/README.md)
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
        "r"(rank)
        : "memory");
}
// One lane of a converged warp.  The single-thread roles run with the WHOLE warp converged and predicate only the
// tcgen05 / TMA instruction on this: inside an `if (lane == 0)` region the compiler cannot prove the operands uniform
// and wraps every UTCHMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~145 clk per MMA, issue-bound; profiles/README.md).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// shared -> global tile store / fp32 reduce-add through the tensor map (clips rows/columns outside the tensor)
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void stg128(void* gptr, uint4 v) {
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(gptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// x[0..3] += v : one 16-byte fp32 vector reduction at L2 (the residual stream add of model.py:195-196)
__device__ __forceinline__ void red_add_f32x4(void* gptr, uint4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gptr), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)),
                 "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t cols) {
    if (CG == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {  // issued by the same warp of both CTAs of the pair
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if (CG == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
    else  // M = 256 over the CTA pair: descriptors address the same offsets in both CTAs' shared memory
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
            : "memory");
}
// arrives on the mbarrier once every previously issued tcgen05.mma of this thread has completed
// (CG = 2: on the barrier at this offset in BOTH CTAs of the pair)
template <int CG>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                     : "memory");
    else
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                smem_u32(bar)),
            "h"(uint16_t(3))
            : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= uint64_t((smem_addr >> 4) & 0x3FFF);
    d |= uint64_t(1) << 16;                  // leading byte offset (unused for swizzled K-major), 16 B units
    d |= uint64_t(1024 >> 4) << 32;          // stride byte offset between 8-row groups
    d |= uint64_t(1) << 46;                  // descriptor version
    d |= uint64_t(2) << 61;                  // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
    return (1u << 4)            // D = fp32
           | (1u << 7)          // A = bf16
           | (1u << 10)         // B = bf16
           | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}


// ---- kind::i8 (int8 x int8 -> int32 in TMEM): the +-1 Hamming contraction of the retrieval evaluator -----------------
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_s8(int m, int n) {
    return (2u << 4)            // D = s32
           | (1u << 7)          // A = s8
           | (1u << 10)         // B = s8
           | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}
// K-major operand tile whose rows are SWZ bytes long (SWZ in {32, 64, 128} = the TMA/UMMA swizzle span): 8-row groups are
// 8*SWZ bytes apart (SBO); layout type 6 / 4 / 2 = SWIZZLE_32B / 64B / 128B; descriptor version 1 (sm_100)
template <int SWZ>
__device__ __forceinline__ uint64_t make_desc_kmajor(uint32_t smem_addr) {
    static_assert(SWZ == 32 || SWZ == 64 || SWZ == 128, "swizzle span");
    uint64_t d = 0;
    d |= uint64_t((smem_addr >> 4) & 0x3FFF);
    d |= uint64_t(1) << 16;
    d |= uint64_t((8 * SWZ) >> 4) << 32;
    d |= uint64_t(1) << 46;
    d |= uint64_t(SWZ == 128 ? 2 : SWZ == 64 ? 4 : 6) << 61;
    return d;
}
// TMEM -> registers without the trailing wait (the caller overlaps several loads, then tcgen05.wait::ld once)
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 64 accumulator columns into 32 registers: register j = (low 16 bits of column 2j) | (low 16 bits of column 2j+1) << 16.
// For int32 accumulators of small magnitude the halves are the values themselves as int16.
__device__ __forceinline__ void tmem_ld32_pack16_async(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace cmh
