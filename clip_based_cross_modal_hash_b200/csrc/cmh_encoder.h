// cmh_encoder.h — internal C++ declarations shared by the encoder translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cmh.h"

namespace cmh {
// out[M][N] = epilogue(A[M][K] . W[N][K]^T + bias); see cmh_gemm.cu
int gemm_bf16(const void* A, int64_t M, int64_t K, int64_t lda, const void* W, int64_t N, int64_t ldw,
              const float* bias, int epi, void* out, int64_t ldo, const float* resid, int64_t ldr, cudaStream_t st);
}  // namespace cmh
