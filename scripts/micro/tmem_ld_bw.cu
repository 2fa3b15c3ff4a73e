// Micro-benchmark: tcgen05.ld throughput per SM (B200), for the shapes the retrieval epilogue could use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu && ./tmem_ld_bw
// Each CTA allocates 128 TMEM columns; `warps` warps (warp w -> lane quarter w % 4) loop over tcgen05.ld + wait::ld.
// Reported: bytes of accumulator data (32-bit columns x lanes) fetched per SM clock, for 1..4 resident CTAs per SM.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int MODE>
__global__ void __launch_bounds__(512) kernel(int iters, long long* out, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t r[32];
        const uint32_t a = base + uint32_t((i & 1) * 64);
        if (MODE == 0) {  // 32x32b.x32: 32 columns
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(a) : "memory");
        } else {  // 32x32b.x32 with pack::16b: 64 columns, two 16-bit halves per register
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(base) : "memory");
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r[j];
    }
    const long long t1 = clock64();
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(128) : "memory");
}

template <int MODE>
void run(const char* name, int warps, int ctas_per_sm) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 20000, grid = sms * ctas_per_sm;
    long long* out;
    uint32_t* sink;
    cudaMalloc(&out, grid * sizeof(long long));
    cudaMalloc(&sink, 4);
    kernel<MODE><<<grid, warps * 32, 0>>>(iters, out, sink);
    kernel<MODE><<<grid, warps * 32, 0>>>(iters, out, sink);
    cudaDeviceSynchronize();
    long long h[2048];
    cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; ++i) avg += double(h[i]);
    avg /= grid;
    const double cols = MODE == 0 ? 32 : 64;
    const double bytes = double(iters) * warps * ctas_per_sm * 32 * cols * 4;  // 32-bit accumulator columns fetched per SM
    printf("%-28s warps/CTA %2d CTAs/SM %d : %7.1f clk/iter  %6.1f accumulator-bytes/clk/SM  (%5.1f accumulators/clk/SM)  err=%s\n", name, warps,
           ctas_per_sm, avg / iters, bytes / avg, bytes / avg / 4, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(sink);
}

int main() {
    for (int c = 1; c <= 4; c *= 2)
        for (int w = 4; w <= 16; w *= 2) {
            run<0>("32x32b.x32", w, c);
            run<1>("32x32b.x32.pack::16b", w, c);
        }
    return 0;
}
