// cmh_encoder.h — internal C++ declarations shared by the encoder translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cmh.h"

namespace cmh {
// out[M][N] = epilogue(A[M][K] . W[N][K]^T + bias); see cmh_gemm.cu
int gemm_bf16(const void* A, int64_t M, int64_t K, int64_t lda, const void* W, int64_t N, int64_t ldw,
              const float* bias, int epi, void* out, int64_t ldo, const float* resid, int64_t ldr, cudaStream_t st);

// cmh_rowwise.cu
int layernorm(const float* x, int64_t rows, int D, int64_t row_mul, const int32_t* row_idx, const float* g, const float* b,
              float eps, void* out, bool out_f32, cudaStream_t st);
int vit_assemble(const float* emb, const float* cls, const float* pos, int64_t B, int L, int D, const float* g, const float* b,
                 float eps, float* x, cudaStream_t st);
int text_embed(const int64_t* text, const float* tok, const float* pos, int64_t B, int L, int D, int vocab, float* x,
               cudaStream_t st);
int text_eos(const int64_t* text, const uint8_t* pad, int64_t B, int L, int64_t eot_id, int32_t* eos, uint8_t* new_mask,
             cudaStream_t st);
int patchify(const float* img, int64_t B, int C, int R, int P, void* out, cudaStream_t st);
int patchify_u8(const uint8_t* img, int64_t B, int C, int R, int P, const float* mean_host, const float* std_host, void* out,
                cudaStream_t st);

// cmh_attention.cu
int attention_bf16(const void* qkv, int64_t B, int L, int H, const uint8_t* pad, int causal, void* out, float* probs,
                   const int32_t* probs_row, cudaStream_t st);
int attention_mean(const float* probs, int64_t B, int H, int L, int skip, const int32_t* zero_col, float* out, cudaStream_t st);

// cmh_heads.cu
int linear_f32(const float* x, int64_t rows, int K, const float* W, const float* bias, int N, const float* scale,
               const float* shift, int act, float* out, int64_t ldo, cudaStream_t st);
int pair_softmax_pack(const float* logits, int64_t rows, int nbits, float* probs, uint32_t* packed, cudaStream_t st);

// cmh_mith.cu
int gather_rows(const float* src, float* dst, int64_t B, int L, int per_sample, int first, int D, cudaStream_t st);
int token_aggregation(const float* concept, const float* x, const uint8_t* pad, const float* pos, int64_t B, int L, int K, int D,
                      int per_sample, int first, int top_k, float* out, cudaStream_t st);
int bit_hash(const float* x, const float* w, const float* bias, int64_t rows, int K, int D, float* out, cudaStream_t st);
int normalize_rows(const float* x, int64_t rows, int D, float* out, cudaStream_t st);
int cast_bf16(const float* x, void* out, int64_t n, cudaStream_t st);
int split3_bf16(const float* x, void* out, int64_t M, int D, cudaStream_t st);
int add_sign_pack(const float* a, const float* b, int64_t rows, int nbits, uint32_t* packed, cudaStream_t st);
}  // namespace cmh
