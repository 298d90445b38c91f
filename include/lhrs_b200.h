/*
 * lhrs_b200.h — C ABI of the B200-native LHRS-Bot hot path (ViT-L/14 -> AttnPooler bridge -> LLaMA-2-7B).
 *
 * The reference (NJU-LHRS/LHRS-Bot) has no FFI of its own: its "operator API" for this path is the
 * Python nn.Module surface of lhrs/models.  Each entry point below names the reference
 * function whose arithmetic it replaces (paths relative to the reference checkout).  All pointers are
 * raw DEVICE pointers borrowed from the caller (torch storage); nothing here allocates parameter memory.
 * Every function returns 0 on success or a non-zero code; lhrs_last_error() returns the message of the
 * last failure on the calling thread.  `stream` is a cudaStream_t passed as void*.
 *
 * bf16 = IEEE bfloat16 stored as uint16_t.  Unless said otherwise matrices are row-major.
 */
#ifndef LHRS_B200_H
#define LHRS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LHRS_OK 0
#define LHRS_ERR_INVALID 1
#define LHRS_ERR_CUDA 2
#define LHRS_ERR_UNSUPPORTED 3

/* library/version + error plumbing */
const char* lhrs_last_error(void);
int lhrs_version(void);
/* Number of kernels this library has launched in this process (bench.py's `gpu_launches`). */
uint64_t lhrs_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Dense contraction (tcgen05 + TMEM accumulators + TMA operand staging).
 *   D[M,N] = epilogue( alpha * A[M,K] · B[N,K]^T )
 * Replaces every nn.Linear / F.linear on the hot path:
 *   - CLIP ViT q/k/v/out/fc1/fc2 + patch-embed conv-as-GEMM (lhrs/models/rgb_vision_modal.py:168-172 -> HF CLIPVisionModel)
 *   - AttnPooler in_proj/out_proj/c_fc/c_proj/out_proj      (lhrs/models/common_arch.py:134-173, 302-333)
 *   - LLaMA q/k/v/o/gate/up/down/lm_head                      (lhrs/models/text_modal.py:281-290 -> HF LlamaForCausalLM)
 *   - backward dX (b_mn_major=1) and dW (a_mn_major=1,b_mn_major=1) forms for SURVEY §8a row a11.
 * ---------------------------------------------------------------------------------------------- */
enum {
    LHRS_EPI_LINEAR = 0, /* optional bias, activation, residual                                  */
    LHRS_EPI_SWIGLU = 1, /* B = {gate, up}; D[M,N/2] = silu(A·gate^T) * (A·up^T)  (HF LlamaMLP) */
    LHRS_EPI_ROPE = 2    /* B = {q,k,v}; rotate-half RoPE applied to the q and k column blocks   */
};
enum { LHRS_ACT_NONE = 0, LHRS_ACT_GELU_ERF = 1, LHRS_ACT_QUICK_GELU = 2 };

typedef struct LhrsGemm {
    int32_t M, N, K;      /* N = total accumulator columns (sum over B segments)                       */
    const void* A;        /* bf16. K-major: [M, lda];  MN-major (a_mn_major): [K, lda] with lda >= M   */
    int64_t lda;          /* leading dimension in elements                                              */
    int32_t a_mn_major;
    const void* B[3];     /* bf16 segments. K-major: each [seg_rows, ldb]; MN-major: one [K, ldb>=N]   */
    int32_t num_b;        /* 1..3.  LINEAR: concatenated along N; SWIGLU: exactly 2; ROPE: exactly 3   */
    int32_t seg_rows;     /* rows (output columns) per segment when num_b > 1                          */
    int64_t ldb;
    int32_t b_mn_major;
    int32_t epilogue;     /* LHRS_EPI_*                                                                 */
    int32_t act;          /* LHRS_ACT_* (LINEAR only)                                                   */
    float alpha;          /* accumulator scale, applied first                                           */
    const void* bias;     /* bf16 [N] or NULL                                                           */
    const void* residual; /* bf16 [*, ldr] or NULL; added after activation; indexed by the SOURCE row   */
    int64_t ldr;
    void* D;              /* bf16 (or fp32 if d_f32) [*, ldd]                                           */
    int64_t ldd;
    int32_t d_f32;
    const int32_t* row_map; /* NULL or [M]: destination row of source row m (scatter epilogue; <0 skips) */
    /* RoPE epilogue (LHRS_EPI_ROPE): head_dim must be 128; tables are [max_pos, 64] fp32 */
    const float* rope_cos;
    const float* rope_sin;
    const int32_t* positions; /* [M] or NULL -> position = m % rope_seq_len                            */
    int32_t rope_seq_len;
    /* SwiGLU epilogue: optionally keep the raw gate / up projections for backward (bf16 [M, N/2]) */
    void* pre_gate;
    void* pre_up;
} LhrsGemm;

int lhrs_gemm_bf16(const LhrsGemm* g, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused attention forward: O = softmax(scale * Q K^T + causal/key-padding mask) V, online softmax.
 * Replaces HF LlamaAttention / CLIPAttention core and nn.MultiheadAttention's SDPA
 * (lhrs/models/common_arch.py:302-313; HF modules reached from rgb_vision_modal.py:168 and text_modal.py:281).
 * Element (b, s, h, d) of X lives at X + b*x_bs + s*x_rs + h*x_hs + d  (strides in elements, multiples of 8).
 * ---------------------------------------------------------------------------------------------- */
typedef struct LhrsAttention {
    const void* q;
    const void* k;
    const void* v;
    void* o;                 /* bf16 */
    float* lse;              /* [B, H, Sq] fp32 log-sum-exp (kept for backward) or NULL */
    const uint8_t* key_mask; /* [B, Skv], 1 = attend (HF attention_mask), or NULL */
    int64_t q_bs, q_rs, q_hs;
    int64_t k_bs, k_rs, k_hs;
    int64_t v_bs, v_rs, v_hs;
    int64_t o_bs, o_rs, o_hs;
    int32_t B, H, Sq, Skv, head_dim; /* head_dim 64 or 128 */
    int32_t causal;                  /* query i attends keys <= i + (Skv - Sq) */
    float scale;
} LhrsAttention;

int lhrs_attention_fwd(const LhrsAttention* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row-wise normalisations and small fused elementwise steps (HBM-bound, one pass each).
 * ---------------------------------------------------------------------------------------------- */
/* HF LlamaRMSNorm: y = w * bf16(x * rsqrt(mean(x^2) + eps)), statistics in fp32.  rstd_out optional [rows]. */
int lhrs_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd_out, int64_t rows, int32_t dim, float eps,
                     void* stream);
/* nn.LayerNorm (lhrs/models/common_arch.py:253-259, CLIP layer norms): fp32 statistics, bf16 in/out.
 * x rows are `ldx` elements apart (lets the ViT tap skip the CLS row without a copy). */
int lhrs_layernorm_fwd(const void* x, int64_t ldx, const void* w, const void* b, void* y, int64_t ldy,
                       float* mean_out, float* rstd_out, int64_t rows, int32_t dim, float eps, void* stream);
/* CLIP patch extraction for the conv-as-GEMM: pixels (B,3,H,W) bf16 -> patches [B*(H/P)*(W/P), kpad] bf16,
 * column order (c, kh, kw) to match Conv2d weight.flatten(1); columns >= 3*P*P are zero. */
int lhrs_vit_im2col(const void* pixels, void* patches, int32_t B, int32_t H, int32_t W, int32_t P, int32_t kpad,
                    void* stream);
/* CLIPVisionEmbeddings + pre_layrnorm: tokens[b,0] = cls + pos[0]; tokens[b,1+p] = patch[b,p] + pos[1+p]; then LN. */
int lhrs_vit_embed_ln(const void* patch_emb, const void* cls, const void* pos, const void* ln_w, const void* ln_b,
                      void* tokens, int32_t B, int32_t num_patches, int32_t dim, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Embed lookup + image splice: TextModal.prepare_inputs_for_multimodal (lhrs/models/text_modal.py:296-526).
 * Integer outputs (labels, mask, lengths) are bit-exact with the reference; embeds are pure gathers.
 *   scan: per sample, count IMAGE_TOKEN_INDEX (-200) occurrences, new length, first image slot.
 *         info[b] = {n_img, new_len, slot_base}; info[B] = {total_slots, max_len, 0}
 *   fill: write embeds (B, S_out, dim), labels (B, S_out) int64, mask (B, S_out) uint8,
 *         row_of_slot [n_slots*num_query] int32 (destination row of each image-feature row, -1 = unused slot).
 * ---------------------------------------------------------------------------------------------- */
int lhrs_splice_scan(const int64_t* input_ids, int32_t B, int32_t T, int32_t num_query, int32_t* info, void* stream);
int lhrs_splice_fill(const int64_t* input_ids, const int64_t* labels /*nullable*/, const uint8_t* attn_mask /*nullable*/,
                     const int32_t* info, const void* embed_table, const void* image_feats /*[slots,num_query,dim] or NULL*/,
                     int32_t B, int32_t T, int32_t S_out, int32_t num_query, int32_t dim, int32_t n_slots,
                     void* embeds_out, int64_t* labels_out, uint8_t* mask_out, int32_t* row_of_slot /*nullable*/,
                     void* stream);
/* backward of the splice w.r.t. the image features: d_image[slot,q,:] = d_embeds[row_of_slot[slot*nq+q], :] (0 if unused) */
int lhrs_splice_bwd(const void* d_embeds, const int32_t* row_of_slot, void* d_image, int64_t n_rows, int32_t dim,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Shifted cross-entropy over bf16 logits (HF LlamaForCausalLM loss, text_modal.py:281-294):
 * row r = (b, s) predicts labels[b, s+1]; ignore_index = -100; mean over counted rows.
 *   fwd: loss_sum[0] += sum of row losses, count[0] += rows counted, lse_out[r] (fp32) kept for backward.
 *   bwd: d_logits[r, :] = (softmax - onehot) * grad_scale / count   (bf16, zero rows where ignored)
 * ---------------------------------------------------------------------------------------------- */
int lhrs_ce_fwd(const void* logits, int64_t ld, const int64_t* labels, int32_t B, int32_t S, int32_t V,
                float* row_lse, float* loss_sum, int32_t* count, void* stream);
int lhrs_ce_bwd(const void* logits, int64_t ld, const int64_t* labels, int32_t B, int32_t S, int32_t V,
                const float* row_lse, const int32_t* count, float grad_scale, void* d_logits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LHRS_B200_H */
