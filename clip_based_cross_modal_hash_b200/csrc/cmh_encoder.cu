// cmh_encoder.cu — host-side schedule of the CLIP ViT-B/32 image and text towers on one stream.
//
// replaces CLIP.encode_image / VisionTransformer.forward (models/CLIP/model.py:232-268,370) and CLIP.encode_text
// (:373-396).  Activations are token-major [B*L][D] (sample-major, so one sample's tokens are contiguous — the
// reference permutes to [L][B][D] for nn.MultiheadAttention, :246,377):
//
//   x    fp32 [M][D]    residual stream; out_proj / c_proj add into it through the GEMM's TMA reduce-add epilogue
//   h    bf16 [M][D]    LayerNorm output = A operand of the next GEMM
//   qkv  bf16 [M][3D]   in_proj output;  att bf16 [M][D] attention output;  fc bf16 [M][4D] QuickGELU(c_fc)
//
// Per block: LN -> GEMM(qkv) -> attention -> GEMM(+x) -> LN -> GEMM(GELU) -> GEMM(+x): 7 launches, no host sync.
#include <cuda_bf16.h>

#include "cmh_common.cuh"
#include "cmh_encoder.h"

namespace cmh {
namespace {

constexpr float LN_EPS = 1e-5f;  // nn.LayerNorm default, used by every LayerNorm of model.py

struct Workspace {
    float* x;
    void *h, *qkv, *att, *fc;
    float* probs;      // [B][H][L]
    int32_t* eos;      // [B]
    void* patches;     // aliases fc (dead before the first block)
    float* emb;        // aliases qkv
    int64_t bytes;
};

int64_t align256(int64_t v) { return round_up(v, 256); }

Workspace carve(const cmh_tower* t, int64_t B, int L, void* base) {
    const int64_t M = B * L, D = t->width;
    Workspace w{};
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        void* p = base ? static_cast<char*>(base) + off : nullptr;
        off += align256(bytes);
        return p;
    };
    w.x = static_cast<float*>(take(M * D * 4));
    w.h = take(M * D * 2);
    w.qkv = take(M * 3 * D * 2);
    w.att = take(M * D * 2);
    int64_t fc_bytes = M * 4 * D * 2;
    if (t->patch > 0) {
        const int64_t pk = int64_t(3) * t->patch * t->patch;
        if (M * pk * 2 > fc_bytes) fc_bytes = M * pk * 2;
    }
    w.fc = take(fc_bytes);
    w.probs = static_cast<float*>(take(B * t->heads * L * 4));
    w.eos = static_cast<int32_t*>(take(B * 4));
    w.patches = w.fc;
    w.emb = static_cast<float*>(w.qkv);  // (L-1)*D fp32 per sample <= 3*L*D bf16
    w.bytes = off;
    return w;
}

int check_tower(const cmh_tower* t) {
    CMH_REQUIRE(t && t->blocks && t->layers > 0, "encoder: tower without blocks");
    CMH_REQUIRE(t->width % 128 == 0 && t->width <= 1024 && t->heads * 64 == t->width,
                "encoder: width %d / heads %d unsupported (width = 64*heads, multiple of 128, <= 1024)", t->width, t->heads);
    CMH_REQUIRE(t->out_dim % 8 == 0 && t->out_dim > 0, "encoder: out_dim %d must be a multiple of 8", t->out_dim);
    CMH_REQUIRE(t->ln_out_gain && t->ln_out_bias && t->w_out_proj && t->pos_emb, "encoder: missing output weights");
    return CMH_OK;
}

// the 12 (or fewer) residual attention blocks, model.py:191-211
int run_blocks(const cmh_tower* t, const Workspace& w, int64_t B, int L, const uint8_t* pad, int causal, bool want_probs,
               const int32_t* probs_row, cudaStream_t st) {
    const int64_t M = B * L, D = t->width;
    for (int i = 0; i < t->layers; ++i) {
        const cmh_block_weights& k = t->blocks[i];
        CMH_REQUIRE(k.w_qkv && k.w_out && k.w_fc && k.w_proj && k.ln1_gain && k.ln2_gain, "encoder: block %d has null weights", i);
        const bool last = i == t->layers - 1;
        if (int rc = layernorm(w.x, M, int(D), 1, nullptr, k.ln1_gain, k.ln1_bias, LN_EPS, w.h, false, st)) return rc;
        if (int rc = gemm_bf16(w.h, M, D, D, k.w_qkv, 3 * D, D, k.b_qkv, CMH_EPI_BF16, w.qkv, 3 * D, nullptr, 0, st)) return rc;
        if (int rc = attention_bf16(w.qkv, B, L, t->heads, pad, causal, w.att, (last && want_probs) ? w.probs : nullptr,
                                    probs_row, st))
            return rc;
        if (int rc = gemm_bf16(w.att, M, D, D, k.w_out, D, D, k.b_out, CMH_EPI_RESID_F32, w.x, D, w.x, D, st)) return rc;
        if (int rc = layernorm(w.x, M, int(D), 1, nullptr, k.ln2_gain, k.ln2_bias, LN_EPS, w.h, false, st)) return rc;
        if (int rc = gemm_bf16(w.h, M, D, D, k.w_fc, 4 * D, D, k.b_fc, CMH_EPI_GELU_BF16, w.fc, 4 * D, nullptr, 0, st)) return rc;
        if (int rc = gemm_bf16(w.fc, M, 4 * D, 4 * D, k.w_proj, D, 4 * D, k.b_proj, CMH_EPI_RESID_F32, w.x, D, w.x, D, st)) return rc;
    }
    return CMH_OK;
}

// final LayerNorm + projection of one row per sample (and, on request, of every token)
int project(const cmh_tower* t, const Workspace& w, int64_t B, int L, const int32_t* row_idx, float* one_out, float* tokens_out,
            cudaStream_t st) {
    const int64_t D = t->width, E = t->out_dim;
    if (tokens_out) {
        if (int rc = layernorm(w.x, B * L, int(D), 1, nullptr, t->ln_out_gain, t->ln_out_bias, LN_EPS, w.h, false, st)) return rc;
        if (int rc = gemm_bf16(w.h, B * L, D, D, t->w_out_proj, E, D, nullptr, CMH_EPI_F32, tokens_out, E, nullptr, 0, st)) return rc;
    }
    // rows picked straight from the residual stream: CLS (row 0 of each sample) or EOS (row_idx[b])
    if (int rc = layernorm(w.x, B, int(D), L, row_idx, t->ln_out_gain, t->ln_out_bias, LN_EPS, w.att, false, st)) return rc;
    return gemm_bf16(w.att, B, D, D, t->w_out_proj, E, D, nullptr, CMH_EPI_F32, one_out, E, nullptr, 0, st);
}

}  // namespace
}  // namespace cmh

extern "C" {

int64_t cmh_encoder_workspace_bytes(const cmh_tower* tower, int64_t batch, int32_t seq_len) {
    if (!tower || batch <= 0 || seq_len <= 0) return 0;
    return cmh::carve(tower, batch, seq_len, nullptr).bytes;
}

int cmh_encode_image(const cmh_tower* t, const float* images, int64_t B, void* workspace, size_t workspace_bytes,
                     float* cls_out, float* tokens_out, float* attn_out, void* stream) {
    using namespace cmh;
    if (int rc = check_tower(t)) return rc;
    CMH_REQUIRE(images && cls_out && B > 0, "encode_image: bad arguments");
    CMH_REQUIRE(t->patch > 0 && t->resolution % t->patch == 0 && t->w_patch && t->cls_emb && t->ln_pre_gain && t->ln_pre_bias,
                "encode_image: not an image tower");
    const int g = t->resolution / t->patch, L = g * g + 1;
    CMH_REQUIRE(L <= 128, "encode_image: %d tokens > 128 unsupported", L);
    const Workspace w = carve(t, B, L, workspace);
    if (!workspace || int64_t(workspace_bytes) < w.bytes || (reinterpret_cast<uintptr_t>(workspace) & 255))
        return fail(CMH_ERR_WORKSPACE, "encode_image: workspace needs %lld bytes, 256-byte aligned", (long long)w.bytes);
    cudaStream_t st = as_stream(stream);
    const int64_t D = t->width, PK = int64_t(3) * t->patch * t->patch, MP = B * (L - 1);
    // conv1 as a GEMM over non-overlapping patches (model.py:235-238)
    if (int rc = patchify(images, B, 3, t->resolution, t->patch, w.patches, st)) return rc;
    if (int rc = gemm_bf16(w.patches, MP, PK, PK, t->w_patch, D, PK, nullptr, CMH_EPI_F32, w.emb, D, nullptr, 0, st)) return rc;
    // [CLS; patches] + positional embedding -> ln_pre (model.py:241-243)
    if (int rc = vit_assemble(w.emb, t->cls_emb, t->pos_emb, B, L, int(D), t->ln_pre_gain, t->ln_pre_bias, LN_EPS, w.x, st)) return rc;
    if (int rc = run_blocks(t, w, B, L, nullptr, 0, attn_out != nullptr, nullptr, st)) return rc;
    if (attn_out) {
        if (int rc = attention_mean(w.probs, B, t->heads, L, 1, nullptr, attn_out, st)) return rc;
    }
    return project(t, w, B, L, nullptr, cls_out, tokens_out, st);
}

int cmh_encode_text(const cmh_tower* t, const int64_t* text, const uint8_t* key_padding_mask, int64_t B, int32_t L,
                    void* workspace, size_t workspace_bytes, float* eos_out, float* tokens_out, float* attn_out,
                    uint8_t* new_mask_out, void* stream) {
    using namespace cmh;
    if (int rc = check_tower(t)) return rc;
    CMH_REQUIRE(text && eos_out && B > 0 && L > 0, "encode_text: bad arguments");
    CMH_REQUIRE(t->tok_emb && t->vocab > 0, "encode_text: not a text tower");
    CMH_REQUIRE(L <= 128 && L <= t->context, "encode_text: %d tokens exceed the context (%d) or 128", L, t->context);
    const Workspace w = carve(t, B, L, workspace);
    if (!workspace || int64_t(workspace_bytes) < w.bytes || (reinterpret_cast<uintptr_t>(workspace) & 255))
        return fail(CMH_ERR_WORKSPACE, "encode_text: workspace needs %lld bytes, 256-byte aligned", (long long)w.bytes);
    cudaStream_t st = as_stream(stream);
    if (int rc = text_embed(text, t->tok_emb, t->pos_emb, B, L, t->width, t->vocab, w.x, st)) return rc;
    if (int rc = text_eos(text, key_padding_mask, B, L, t->eot_id, w.eos, new_mask_out, st)) return rc;
    if (int rc = run_blocks(t, w, B, L, key_padding_mask, 1, attn_out != nullptr, w.eos, st)) return rc;
    if (attn_out) {
        if (int rc = attention_mean(w.probs, B, t->heads, L, 0, w.eos, attn_out, st)) return rc;
    }
    return project(t, w, B, L, w.eos, eos_out, tokens_out, st);
}
}
