// POPC pipe throughput on this GPU: nvcc -O3 -gencode arch=compute_100a,code=sm_100a popc_peak.cu -o popc_peak && ./popc_peak
#include <cstdio>
#include <cuda_runtime.h>
__global__ void popc_kernel(unsigned* out, unsigned seed, int iters) {
    unsigned a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed * (threadIdx.x + 1) + j * 0x9E3779B9u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = __popc(a[j] ^ seed) + (a[j] << 3);  // popc + one cheap ALU op keeps the value alive
    }
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int blocks = sms * 8, threads = 256, iters = 1 << 14;
    unsigned* out;
    cudaMalloc(&out, size_t(blocks) * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    popc_kernel<<<blocks, threads>>>(out, 12345u, iters);
    cudaEventRecord(e0);
    popc_kernel<<<blocks, threads>>>(out, 12345u, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = double(blocks) * threads * iters * 8;
    printf("SMs %d, max clock %.0f MHz: %.3e popc/s = %.1f popc/clk/SM at the max clock (%.3f ms)\n", sms, khz / 1e3, ops / (ms * 1e-3),
           ops / (ms * 1e-3) / sms / (khz * 1e3), ms);
    return 0;
}
