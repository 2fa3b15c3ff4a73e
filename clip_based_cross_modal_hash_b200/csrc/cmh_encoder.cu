// cmh_encoder.cu — host-side schedule of the CLIP ViT-B/32 image and text towers on one stream.
//
// replaces CLIP.encode_image / VisionTransformer.forward (models/CLIP/model.py:232-268,370) and CLIP.encode_text
// (:373-396).  Activations are token-major [B*L][D] (sample-major, so one sample's tokens are contiguous — the
// reference permutes to [L][B][D] for nn.MultiheadAttention, :246,377):
//
//   x    fp32 [M][D]    residual stream; out_proj / c_proj add into it through the GEMM epilogue (vector red.add)
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

// ---- MITH head -------------------------------------------------------------------------------------------------
struct MithWs {
    float *xc, *xt, *concept, *projout, *h32, *g32;
    void *hc, *gc, *ht, *gt, *x2b;
    void* tw;  // transformer workspace (carve)
    int64_t tw_bytes, bytes;
};

MithWs carve_mith(const cmh_mith_head* h, int64_t B, int L, void* base) {
    const int64_t D = h->dim, K = h->nbits, M = B * L, M2 = B * K;
    MithWs w{};
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        void* p = base ? static_cast<char*>(base) + off : nullptr;
        off += align256(bytes);
        return p;
    };
    w.xc = static_cast<float*>(take(B * D * 4));
    w.hc = take(B * D * 2);
    w.gc = take(B * 4 * D * 2);
    w.xt = static_cast<float*>(take(M * D * 4));
    const bool split = h->mlp_layers > 0 && h->w1_split[0] != nullptr;
    w.ht = take(M * D * 2 * (split ? 3 : 1));          // split precision: [hi | hi | lo]
    w.gt = take(M * 4 * D * 2 * (split ? 3 : 1));
    w.h32 = static_cast<float*>(take(split ? M * D * 4 : 0));
    w.g32 = static_cast<float*>(take(split ? M * 4 * D * 4 : 0));
    w.concept = static_cast<float*>(take(M * K * 4));
    w.x2b = take(M2 * D * 2);
    w.projout = static_cast<float*>(take(M2 * D * 4));
    w.tw_bytes = carve(&h->transformer, B, int(K), nullptr).bytes;
    w.tw = take(w.tw_bytes);
    w.bytes = off;
    return w;
}

// ResidualMLPs.forward (hash.py:34-37) on `rows` rows of the fp32 stream x, in place.  h32/g32 non-null: split precision
// (fp32 LayerNorm and GELU outputs are split into bf16 hi/lo parts and multiplied with the split weights, K concatenated).
int res_mlps(const cmh_mith_head* h, float* x, void* hb, void* gb, float* h32, float* g32, int64_t rows, cudaStream_t st) {
    const int64_t D = h->dim;
    for (int i = 0; i < h->mlp_layers; ++i) {
        CMH_REQUIRE(h->w1[i] && h->w2[i] && h->ln_gain[i] && h->ln_bias[i], "mith: residual MLP %d has null weights", i);
        if (h32) {
            CMH_REQUIRE(h->w1_split[i] && h->w2_split[i], "mith: residual MLP %d lacks the split weights", i);
            if (int rc = layernorm(x, rows, int(D), 1, nullptr, h->ln_gain[i], h->ln_bias[i], LN_EPS, h32, true, st)) return rc;
            if (int rc = split3_bf16(h32, hb, rows, int(D), st)) return rc;
            if (int rc = gemm_bf16(hb, rows, 3 * D, 3 * D, h->w1_split[i], 4 * D, 3 * D, h->b1[i], CMH_EPI_ERF_GELU_F32, g32, 4 * D,
                                   nullptr, 0, st))
                return rc;
            if (int rc = split3_bf16(g32, gb, rows, int(4 * D), st)) return rc;
            if (int rc = gemm_bf16(gb, rows, 12 * D, 12 * D, h->w2_split[i], D, 12 * D, h->b2[i], CMH_EPI_RESID_F32, x, D, x, D, st))
                return rc;
            continue;
        }
        if (int rc = layernorm(x, rows, int(D), 1, nullptr, h->ln_gain[i], h->ln_bias[i], LN_EPS, hb, false, st)) return rc;
        if (int rc = gemm_bf16(hb, rows, D, D, h->w1[i], 4 * D, D, h->b1[i], CMH_EPI_ERF_GELU_BF16, gb, 4 * D, nullptr, 0, st)) return rc;
        if (int rc = gemm_bf16(gb, rows, 4 * D, 4 * D, h->w2[i], D, 4 * D, h->b2[i], CMH_EPI_RESID_F32, x, D, x, D, st)) return rc;
    }
    return CMH_OK;
}

}  // namespace
}  // namespace cmh

extern "C" {

int64_t cmh_head_mith_workspace_bytes(const cmh_mith_head* head, int64_t batch, int32_t tokens) {
    if (!head || batch <= 0 || tokens <= 0 || head->nbits <= 0 || head->dim <= 0 || !head->transformer.blocks) return 0;
    return cmh::carve_mith(head, batch, tokens, nullptr).bytes;
}

int cmh_head_mith(const cmh_mith_head* h, const float* cls, const float* tokens, int32_t per_sample, int32_t first, int32_t L,
                  const uint8_t* pad, int64_t B, void* workspace, size_t workspace_bytes, float* res_cls, float* cls_hash,
                  float* tokens_hash, float* trans_tokens, uint32_t* packed, void* stream) {
    using namespace cmh;
    CMH_REQUIRE(h && cls && tokens && cls_hash && tokens_hash && B > 0 && L > 0 && first >= 0 && first + L <= per_sample,
                "head_mith: bad arguments");
    CMH_REQUIRE(h->dim % 128 == 0 && h->nbits >= 1 && h->nbits <= 128 && h->mlp_layers >= 0 && h->mlp_layers <= CMH_MITH_MAX_MLP_LAYERS,
                "head_mith: dim %d / k_bits %d / res_mlp_layers %d unsupported", h->dim, h->nbits, h->mlp_layers);
    CMH_REQUIRE(h->w_concept && h->pos && h->w_bits && h->b_bits, "head_mith: missing weights");
    CMH_REQUIRE(h->transformer.width == h->dim && h->transformer.heads * 64 == h->dim && h->transformer.blocks && h->transformer.layers > 0,
                "head_mith: the concept transformer must have the head's width");
    const int64_t D = h->dim, K = h->nbits;
    const MithWs w = carve_mith(h, B, L, workspace);
    if (!workspace || int64_t(workspace_bytes) < w.bytes || (reinterpret_cast<uintptr_t>(workspace) & 255))
        return fail(CMH_ERR_WORKSPACE, "head_mith: workspace needs %lld bytes, 256-byte aligned", (long long)w.bytes);
    cudaStream_t st = as_stream(stream);
    // global branch: res_cls, cls_hash = gcl(cls)   (hash.py:233-234, 98-106)
    if (int rc = gather_rows(cls, w.xc, B, 1, 1, 0, int(D), st)) return rc;
    if (int rc = res_mlps(h, w.xc, w.hc, w.gc, nullptr, nullptr, B, st)) return rc;
    if (int rc = linear_f32(w.xc, B, int(D), h->w_concept, nullptr, int(K), nullptr, nullptr, CMH_ACT_TANH, cls_hash, K, st)) return rc;
    if (res_cls) {
        if (int rc = normalize_rows(w.xc, B, int(D), res_cls, st)) return rc;
    }
    // local branch: concept embedding of every token (hash.py:235), aggregation to K concept tokens (+ position)
    if (int rc = gather_rows(tokens, w.xt, B, L, per_sample, first, int(D), st)) return rc;
    const bool split = h->mlp_layers > 0 && h->w1_split[0] != nullptr;
    if (int rc = res_mlps(h, w.xt, w.ht, w.gt, split ? w.h32 : nullptr, split ? w.g32 : nullptr, B * L, st)) return rc;
    if (int rc = linear_f32(w.xt, B * L, int(D), h->w_concept, nullptr, int(K), nullptr, nullptr, CMH_ACT_TANH, w.concept, K, st)) return rc;
    const Workspace tw = carve(&h->transformer, B, int(K), w.tw);
    if (int rc = token_aggregation(w.concept, tokens, pad, h->pos, B, L, int(K), int(D), per_sample, first, h->top_k, tw.x, st)) return rc;
    // Transformer over the K concept tokens (hash.py:184-190), no mask
    if (int rc = run_blocks(&h->transformer, tw, B, int(K), nullptr, 0, false, nullptr, st)) return rc;
    // BitwiseHashing (hash.py:67-83)
    if (int rc = bit_hash(tw.x, h->w_bits, h->b_bits, B * K, int(K), int(D), tokens_hash, st)) return rc;
    if (trans_tokens) {  // normalize(concept_proj(x)) (hash.py:236, 244)
        CMH_REQUIRE(h->w_cproj, "head_mith: trans_tokens requested without the concept projection weights");
        if (int rc = cast_bf16(tw.x, w.x2b, B * K * D, st)) return rc;
        if (int rc = gemm_bf16(w.x2b, B * K, D, D, h->w_cproj, D, D, h->b_cproj, CMH_EPI_F32, w.projout, D, nullptr, 0, st)) return rc;
        if (int rc = normalize_rows(w.projout, B * K, int(D), trans_tokens, st)) return rc;
    }
    if (packed) return add_sign_pack(cls_hash, tokens_hash, B, int(K), packed, st);
    return CMH_OK;
}

int64_t cmh_encoder_workspace_bytes(const cmh_tower* tower, int64_t batch, int32_t seq_len) {
    if (!tower || batch <= 0 || seq_len <= 0) return 0;
    return cmh::carve(tower, batch, seq_len, nullptr).bytes;
}

static int encode_image_any(const cmh_tower* t, const float* images_f32, const uint8_t* images_u8, const float* mean, const float* stdv,
                            int64_t B, void* workspace, size_t workspace_bytes, float* cls_out, float* tokens_out, float* attn_out,
                            void* stream) {
    using namespace cmh;
    if (int rc = check_tower(t)) return rc;
    CMH_REQUIRE((images_f32 || images_u8) && cls_out && B > 0, "encode_image: bad arguments");
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
    if (images_u8) {
        if (int rc = patchify_u8(images_u8, B, 3, t->resolution, t->patch, mean, stdv, w.patches, st)) return rc;
    } else {
        if (int rc = patchify(images_f32, B, 3, t->resolution, t->patch, w.patches, st)) return rc;
    }
    if (int rc = gemm_bf16(w.patches, MP, PK, PK, t->w_patch, D, PK, nullptr, CMH_EPI_F32, w.emb, D, nullptr, 0, st)) return rc;
    // [CLS; patches] + positional embedding -> ln_pre (model.py:241-243)
    if (int rc = vit_assemble(w.emb, t->cls_emb, t->pos_emb, B, L, int(D), t->ln_pre_gain, t->ln_pre_bias, LN_EPS, w.x, st)) return rc;
    if (int rc = run_blocks(t, w, B, L, nullptr, 0, attn_out != nullptr, nullptr, st)) return rc;
    if (attn_out) {
        if (int rc = attention_mean(w.probs, B, t->heads, L, 1, nullptr, attn_out, st)) return rc;
    }
    return project(t, w, B, L, nullptr, cls_out, tokens_out, st);
}

int cmh_encode_image(const cmh_tower* t, const float* images, int64_t B, void* workspace, size_t workspace_bytes,
                     float* cls_out, float* tokens_out, float* attn_out, void* stream) {
    return encode_image_any(t, images, nullptr, nullptr, nullptr, B, workspace, workspace_bytes, cls_out, tokens_out, attn_out, stream);
}

int cmh_encode_image_u8(const cmh_tower* t, const uint8_t* images, const float* mean_host, const float* std_host, int64_t B,
                        void* workspace, size_t workspace_bytes, float* cls_out, float* tokens_out, float* attn_out, void* stream) {
    return encode_image_any(t, nullptr, images, mean_host, std_host, B, workspace, workspace_bytes, cls_out, tokens_out, attn_out, stream);
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
