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
/* Optional per-kernel timing for the roofline report: when enabled, CUDA events bracket every launch of the
 * GEMM (kind 0: the 2-CTA 256x256 instantiations gemm_bf16_kernel<256,*,*,*,2> — the dominant kernel; kind 3: the single-CTA
 * instantiations used for small / skinny problems), attention (kind 1), decode GEMV (kind 2) and LoRA streaming (kind 4) kernels.  lhrs_prof_summary synchronises the device and
 * returns the summed durations and the summed ALGORITHMIC flops / bytes of the launches recorded since enable. */
int lhrs_prof_enable(int on);
int lhrs_prof_summary(int kind, double* ms, double* flops, double* bytes, int64_t* launches);

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
    const void* bias[3];  /* bf16, one per B segment ([seg_rows] each; bias[0] is [N] when num_b == 1) or NULL  */
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
    void* pre_gate;       /* LINEAR: optional copy of the pre-activation (alpha*acc + bias), bf16 [M, N] */
    void* pre_up;         /* LINEAR with BOTH set (N % 32 == 0): fused SwiGLU backward — the accumulator is d_act, pre_gate /
                             pre_up are the stashed forward pre-activations (inputs), D [M, 2N] receives [d_gate | d_up] */
    /* K-extension (LoRA, common_arch-independent; text_modal.py:133-151): after the K loop over A/B the kernel keeps
     * accumulating A2[:, s*ext_k:(s+1)*ext_k] · B2[s]^T into the same TMEM tile, s = B segment of the tile.
     * A2 is bf16 [M, lda2] holding num_b blocks of ext_k columns (T = scale * x·lora_A^T); B2[s] is lora_B [seg_rows, ldb2]. */
    const void* A2;
    int64_t lda2;
    const void* B2[3];
    int64_t ldb2;
    int32_t ext_k;
    /* MN-major B with num_b > 1 = segments stacked along K.  b_seg_nshift > 0 makes them block-diagonal: segment s is
     * [K/num_b, b_seg_nshift] and contributes only to output columns [s*b_seg_nshift, (s+1)*b_seg_nshift). */
    int32_t b_seg_nshift;
    /* split_k > 1: split the K loop over that many CTAs per output tile (skinny problems that would otherwise occupy few
     * SMs).  Partial sums are added with fp32 atomics: D must be fp32 (d_f32) and zero-initialised; plain LINEAR epilogue. */
    int32_t split_k;
    /* LINEAR epilogue, drop_t > 0: zero the accumulator elements the LoRA dropout mask of drop_key drops (row = output row,
     * column = output column, ld = N; see lhrs_lora_dropout_mask) before the residual is added.  N % 32 == 0. */
    uint32_t drop_key;
    int32_t drop_t;
} LhrsGemm;

int lhrs_gemm_bf16(const LhrsGemm* g, void* stream);
/* fp32 -> bf16 (round to nearest even), n elements, n % 4 == 0 (finishes a split-K accumulation) */
int lhrs_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);

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
    /* Ragged ("padding-free") batch, optional: device int32 [B+1] row offsets.  Sequence b then occupies rows
     * [seq_off[b], seq_off[b+1]) of ONE [total_rows, *] buffer per operand (x_bs is ignored, row s of sequence b is row
     * seq_off[b] + s), its length replaces Sq = Skv for that batch entry, and Sq / Skv become the LONGEST length: they size the
     * grid and the [B, H, Sq] lse / delta arrays.  Needs head_dim 128, causal, Sq == Skv >= 128, key_mask == NULL (a right-padded
     * HF attention_mask is exactly what the lengths say).  NULL = dense [B, S] layout. */
    const int32_t* seq_off;
    int64_t total_rows;              /* seq_off[B], on the host (tensor maps are built from it) */
} LhrsAttention;

int lhrs_attention_fwd(const LhrsAttention* a, void* stream);
/* n (1..3) problems that share B, H, head_dim 64 and the causal flag as ONE launch: the three query groups of the AttnPooler's
 * cross-attention (common_arch.py:159-166).  Falls back to n single launches when the problems cannot be grouped. */
int lhrs_attention_fwd_grouped(const LhrsAttention* a, int32_t n, void* stream);

/* Backward of lhrs_attention_fwd (recompute-based; needs the forward's O and lse).  dQ/dK/dV are bf16 with their own
 * (batch,row,head) strides so they can land in a packed [rows, 3*H*hd] buffer.  delta: fp32 scratch of
 * lhrs_attention_bwd_scratch_floats(B, H, Sq, Skv) elements ([B,H,Sq] row sums of dO*O, then the key mask packed to bits). */
int64_t lhrs_attention_bwd_scratch_floats(int32_t B, int32_t H, int32_t Sq, int32_t Skv);
typedef struct LhrsAttentionBwd {
    LhrsAttention fwd;   /* the forward problem: q,k,v,o,lse,key_mask,strides,sizes */
    const void* d_o;     /* same layout as fwd.o */
    void* dq; void* dk; void* dv;
    float* delta;
    int64_t dq_bs, dq_rs, dq_hs, dk_bs, dk_rs, dk_hs, dv_bs, dv_rs, dv_hs;
    /* optional (head_dim 128): undo the RoPE rotation on dQ and dK before they are stored (position = row index), so the
     * gradients come out w.r.t. the un-rotated projections; tables [max_pos, 64] fp32 as in LhrsGemm */
    const float* rope_cos;
    const float* rope_sin;
} LhrsAttentionBwd;
int lhrs_attention_bwd(const LhrsAttentionBwd* a, void* stream);
/* grouped form (see lhrs_attention_fwd_grouped): delta, dQ and dK/dV of up to three problems as one launch each */
int lhrs_attention_bwd_grouped(const LhrsAttentionBwd* a, int32_t n, void* stream);

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
                const float* row_lse, const int32_t* count, float grad_scale, const float* grad_scale_dev /*nullable, x[0]*/,
                void* d_logits, void* stream);


/* ================================================================================================
 * Model-level entry points: the layer loops run inside the library (one call per module forward), so
 * Python adds no per-layer overhead.  Weight tables hold borrowed device pointers (bf16 unless noted),
 * one entry per layer, in the reference's parameter layout (HF names in comments).
 * ============================================================================================== */

/* CLIP ViT-L/14 encoder up to the last tap.  VisionModal.encode, lhrs/models/rgb_vision_modal.py:166-184. */
typedef struct LhrsVitWeights {
    int32_t num_layers;  /* layers to evaluate = last tap index (22 of 24: layers after the last tap are never read) */
    int32_t dim, ffn, heads, patch, image, kpad; /* 1024, 4096, 16, 14, 224, 640 (3*14*14=588 zero-padded) */
    float eps;
    const void* patch_w;  /* [dim, kpad]  embeddings.patch_embedding.weight.flatten(1), zero-padded columns */
    const void* cls;      /* [dim]        embeddings.class_embedding */
    const void* pos;      /* [1+patches, dim] embeddings.position_embedding.weight */
    const void* pre_ln_w; const void* pre_ln_b;               /* pre_layrnorm */
    const void* const* ln1_w; const void* const* ln1_b;       /* encoder.layers.N.layer_norm1 */
    const void* const* q_w; const void* const* q_b;           /* self_attn.q_proj [dim,dim] / [dim] */
    const void* const* k_w; const void* const* k_b;
    const void* const* v_w; const void* const* v_b;
    const void* const* o_w; const void* const* o_b;           /* self_attn.out_proj */
    const void* const* ln2_w; const void* const* ln2_b;
    const void* const* fc1_w; const void* const* fc1_b;       /* mlp.fc1 [ffn,dim] */
    const void* const* fc2_w; const void* const* fc2_b;       /* mlp.fc2 [dim,ffn] */
} LhrsVitWeights;

size_t lhrs_vit_workspace_bytes(const LhrsVitWeights* w, int32_t B);
/* pixels (B,3,image,image) bf16 -> out (B, n_taps*patches, dim) bf16 = cat_i hidden_states[taps[i]][:,1:,:].
 * `taps` is a HOST array, ascending, taps[n_taps-1] == num_layers. */
int lhrs_vit_fwd(const LhrsVitWeights* w, const void* pixels, int32_t B, const int32_t* taps, int32_t n_taps,
                 void* out, void* workspace, size_t workspace_bytes, void* stream);

/* AttnPooler (multi-level query perceiver).  lhrs/models/common_arch.py:134-173, 315-333. */
typedef struct LhrsPoolerWeights {
    int32_t num_layers, dim, ffn, heads, out_dim; /* 6, 1024, 4096, 16, 4096 */
    int32_t num_groups;                           /* 3 */
    int32_t stage_num[4];                         /* queries per group   {64,48,32} */
    int32_t split_part[4];                        /* image tokens per group {256,256,256} */
    float eps;
    const void* query;                                         /* [sum(stage_num), dim] */
    const void* const* ln1_w; const void* const* ln1_b;        /* layers.N.ln_1 */
    const void* const* lnkv_w; const void* const* lnkv_b;      /* layers.N.ln_1_kv */
    const void* const* in_w; const void* const* in_b;          /* layers.N.attn.in_proj_{weight,bias} [3*dim,dim] */
    const void* const* ao_w; const void* const* ao_b;          /* layers.N.attn.out_proj */
    const void* const* ln2_w; const void* const* ln2_b;
    const void* const* fc_w; const void* const* fc_b;          /* layers.N.mlp.c_fc   [ffn,dim] */
    const void* const* pj_w; const void* const* pj_b;          /* layers.N.mlp.c_proj [dim,ffn] */
    const void* out_w; const void* out_b;                      /* out_proj [out_dim, dim] */
} LhrsPoolerWeights;

size_t lhrs_pooler_workspace_bytes(const LhrsPoolerWeights* w, int32_t B);
/* bytes of the activation stash a later lhrs_pooler_bwd needs (0 is never returned; pass stash=NULL for inference) */
size_t lhrs_pooler_stash_bytes(const LhrsPoolerWeights* w, int32_t B);
/* image_embs (B, sum(split_part), dim) bf16 -> out.  Rows are written at out + row_map[b*nq + i]*ldo when row_map is
 * given (scatter straight into the LLaMA inputs_embeds buffer, fusing the splice), else at (b*nq + i)*ldo. */
int lhrs_pooler_fwd(const LhrsPoolerWeights* w, const void* image_embs, int32_t B, void* out, int64_t ldo,
                    const int32_t* row_map, void* stash, void* workspace, size_t workspace_bytes, void* stream);

/* Paged KV cache (HBM): pool is bf16 [layers][2 (k,v)][num_pages][heads][page_size][head_dim]; sequence b owns pages
 * block_table[b*max_pages + i] (device int32).  K is stored post-RoPE.  Memory is caller-owned (torch). */
typedef struct LhrsKvCache {
    void* pool;
    const int32_t* block_table;
    int32_t layers, heads, head_dim, page_size, num_pages, max_pages;
} LhrsKvCache;

/* LLaMA-2 decoder stack.  HF LlamaForCausalLM as called from lhrs/models/text_modal.py:281-290 (train) and
 * :600-612 (generate).  Optional LoRA on the seven projections (text_modal.py:133-151): A [r,in], B [out,r]. */
typedef struct LhrsLlamaWeights {
    int32_t num_layers, dim, ffn, heads, vocab, max_pos; /* 32, 4096, 11008, 32, 32000, 2048 */
    float eps;
    const void* const* ln1_w;   /* model.layers.N.input_layernorm.weight [dim] */
    const void* const* q_w; const void* const* k_w; const void* const* v_w; const void* const* o_w; /* [dim,dim] */
    const void* const* ln2_w;   /* post_attention_layernorm */
    const void* const* gate_w; const void* const* up_w;  /* [ffn,dim] */
    const void* const* down_w;                            /* [dim,ffn] */
    const void* norm_w;         /* model.norm.weight */
    const void* lm_head;        /* lm_head.weight [vocab, dim] */
    const void* embed;          /* model.embed_tokens.weight [vocab, dim] */
    const float* rope_cos; const float* rope_sin; /* [max_pos, 64] fp32 tables (bf16-rounded values, as HF casts them) */
    /* LoRA (all NULL when disabled).  Index [layer*7 + p], p in {q,k,v,o,gate,up,down}. */
    int32_t lora_r; float lora_scale;
    const void* const* lora_a; const void* const* lora_b;
    /* peft's input dropout of the LoRA branch for THIS call (training forward and its backward: same seed); 0 = off (eval).
     * The mask is a pure function of (lora_seed, layer * 7 + projection, row, column): lhrs_lora_dropout_mask. */
    float lora_dropout;
    uint64_t lora_seed;
    /* Optional TRANSPOSED copies of the frozen projection weights for the dX GEMMs of lhrs_llama_bwd (all NULL = read the
     * weights MN-major in place).  180 GB of HBM affords a second, K-major copy (+13 GB for LLaMA-2-7B) and the K-major form of
     * the tcgen05 GEMM runs the dX contractions 10-20 % faster than the MN-major one:
     *   qkv_wt[l] = [Wq;Wk;Wv]^T [dim, 3*dim]   o_wt[l] = Wo^T [dim, dim]   gu_wt[l] = [Wgate;Wup]^T [dim, 2*ffn]
     *   down_wt[l] = Wdown^T [ffn, dim]          lm_head_wt = lm_head^T [dim, vocab] */
    const void* const* qkv_wt; const void* const* o_wt; const void* const* gu_wt; const void* const* down_wt;
    const void* lm_head_wt;
} LhrsLlamaWeights;

size_t lhrs_llama_workspace_bytes(const LhrsLlamaWeights* w, int32_t B, int32_t S);
size_t lhrs_llama_stash_bytes(const LhrsLlamaWeights* w, int32_t B, int32_t S);
/* inputs_embeds (B,S,dim) bf16 -> hidden_out (B,S,dim) bf16 = model.norm(last layer).  key_mask (B,S) uint8 or NULL.
 * stash: NULL (inference) or a buffer of lhrs_llama_stash_bytes kept for lhrs_llama_bwd_dx.
 * kv_cache: NULL, or the paged cache that receives K/V of positions 0..S-1 (prefill before lhrs_llama_decode_step). */
int lhrs_llama_fwd(const LhrsLlamaWeights* w, const void* inputs_embeds, int32_t B, int32_t S, const uint8_t* key_mask,
                   void* hidden_out, void* stash, const LhrsKvCache* kv_cache, void* workspace, size_t workspace_bytes,
                   void* stream);
/* The same stack over a RAGGED batch ("padding-free"): the B sequences sit back to back in `rows` = seq_off[B] rows, without
 * their right padding — HF's LlamaModel under a right-padded attention_mask computes the padded positions and nobody reads them
 * (text_modal.py:398-412 passes the mask; the shifted CE ignores label -100, modeling_llama's causal mask keeps padded keys out
 * of every real query row), so every result on a real position is the same.  seq_off: device int32 [B+1]; positions: device
 * int32 [rows], each row's index inside its sequence (RoPE); S_max: the longest length (>= 128: tcgen05 attention only).
 * inputs_embeds / hidden_out are [rows, dim].  Workspace and stash are sized by the dense functions with (B, S_max). */
int lhrs_llama_fwd_ragged(const LhrsLlamaWeights* w, const void* inputs_embeds, int32_t B, int32_t S_max, int64_t rows,
                          const int32_t* seq_off, const int32_t* positions, void* hidden_out, void* stash, void* workspace,
                          size_t workspace_bytes, void* stream);
/* logits = hidden · lm_head^T (bf16 [rows, vocab]) */
int lhrs_lm_head(const LhrsLlamaWeights* w, const void* hidden, int64_t rows, void* logits, void* stream);

/* ================================================================================================
 * Backward and the flat-buffer optimizer step (§8f-2).
 * ============================================================================================== */
int lhrs_rmsnorm_bwd(const void* x, const void* w, const float* rstd, const void* dy, const void* dres /*nullable: added*/,
                     void* dx, int64_t rows, int32_t dim, void* stream);
size_t lhrs_layernorm_bwd_scratch_bytes(int32_t dim);
/* dx = dres + LN'(dy); dw/db (bf16 [dim], nullable) = column reductions; scratch: lhrs_layernorm_bwd_scratch_bytes */
int lhrs_layernorm_bwd(const void* x, int64_t ldx, const void* w, const float* mean, const float* rstd, const void* dy,
                       const void* dres, void* dx, void* dw, void* db, int32_t accumulate, float* scratch, int64_t rows,
                       int32_t dim, void* stream);
size_t lhrs_colsum_scratch_bytes(int32_t n);
int lhrs_colsum(const void* a, int64_t ld, int64_t rows, int32_t n, void* out, int32_t accumulate, float* scratch, void* stream);
/* LoRA input dropout (peft lora.Linear, text_modal.py:136-143): out = x with the dropped elements zeroed (survivors NOT scaled;
 * callers fold 256 / (256 - T) into the product that consumes it).  keep(row, col) is the counter-based hash documented in
 * csrc/dropout.cuh: drop probability T / 256 with T = round(p * 256); module = layer * 7 + {q,k,v,o,gate,up,down}. */
int lhrs_lora_dropout_mask(const void* x, int64_t ldx, int64_t rows, int32_t cols, uint64_t seed, int32_t module, float p,
                           void* out, int64_t ldo, void* stream);
/* Fused-mask forms of the rank-16 streaming side products (no masked copy of the activation is written):
 *   out[M, n] = alpha * 256/(256-T) * [ (mask_p o x) · w_p^T ]_p      w = [n, ldw] = lora_A of n/16 projections stacked
 *   dst[n, C] = 256/(256-T) * [ q_p^T · (mask_p o x) ]_p               q = dT [M, n]  (dA of the n/16 projections, stacked) */
int lhrs_lora_panel_dropout(const void* x, int64_t ldx, int64_t M, int32_t K, const void* w, int64_t ldw, int32_t n, float alpha,
                            uint64_t seed, int32_t module0, float p, void* out, int64_t ldo, void* stream);
int lhrs_lora_rowreduce_dropout(const void* x, int64_t ldx, int64_t M, int32_t C, const void* q, int64_t ldq, int32_t n, void* dst,
                                int64_t ldd, uint64_t seed, int32_t module0, float p, float* scratch, size_t scratch_bytes,
                                void* stream);
/* dx[M, C] += 256/(256-T) * sum_p mask_p o (dT[:, 16p:16p+16] · lora_a[p][16, C]): the LoRA branch's input gradient under
 * dropout (rank 16; module of projection p = module0 + p).  One streaming read-modify-write pass over dx. */
int lhrs_lora_dx_dropout(void* dx, int64_t ldx, int64_t M, int32_t C, const void* dT, int64_t ldt, const void* const* lora_a,
                         int32_t nproj, uint64_t seed, int32_t module0, float p, void* stream);
/* d_gu[rows, 2f] = [d_gate | d_up] from d_act and the stashed pre-activations (HF LlamaMLP backward) */
int lhrs_swiglu_bwd(const void* d_act, const void* pre_gate, const void* pre_up, void* d_gu, int64_t rows, int32_t f, void* stream);
int lhrs_gelu_bwd(void* d /*in place*/, const void* pre, int64_t n, void* stream);
/* inverse rotation, in place on the q and k blocks of a packed [rows, 3*dim] gradient (head_dim 128) */
int lhrs_rope_bwd(void* dqkv, int64_t ld, int64_t rows, int32_t dim, const float* cos, const float* sin,
                  const int32_t* positions, int32_t seq_len, void* stream);

/* Rank-r side products of LoRA as streaming kernels (HBM-bound: one pass over the large activation; csrc/skinny.cu).  They carry the
 * arithmetic peft's lora.Linear adds around every wrapped projection (lhrs/models/text_modal.py:133-151) and its autograd:
 *   lhrs_lora_panel      out[M, n] = alpha * X[M, K] · W,  n in {16, 32, 48}, K % 64 == 0
 *        w_kn = 0: w[0] = [n, ldw] K-major                     forward   T  = s · x · [A_0;A_1;..]^T
 *        w_kn = 1: w[s] = [K/nseg, 16], block diagonal          backward  dT_s = s · dy_s · B_s   (B_s = lora_B_s [out, r = 16])
 *   lhrs_lora_rowreduce  G = P[M, C]^T · Q[M, n_q]  reduced over the M rows, fp32 row-split partials in `scratch`
 *        transpose = 1, seg_c = 0:  dst[0][j, c] (ld = ldd) = G[c, j], n = n_q            dA = dT^T · x   as [n, in]
 *        transpose = 0, seg_c > 0:  column block c uses Q columns [16*(c / seg_c), +16), n = 16;
 *                                   dst[s][c - s*seg_c, j] (ld = ldd)                   dB_s = dy_s^T · T_s as [out, 16]
 * `w` and `dst` are HOST arrays of device pointers. */
int lhrs_lora_panel(const void* x, int64_t ldx, int64_t M, int32_t K, const void* const* w, int32_t nseg, int32_t w_kn, int64_t ldw,
                    int32_t n, float alpha, void* out, int64_t ldo, void* stream);
size_t lhrs_lora_rowreduce_scratch_bytes(int64_t M, int32_t C, int32_t n);
int lhrs_lora_rowreduce(const void* p, int64_t ldp, int64_t M, int32_t C, const void* q, int64_t ldq, int32_t n, int32_t seg_c,
                        int32_t transpose, void* const* dst, int64_t ldd, float alpha, float* scratch, size_t scratch_bytes, void* stream);

/* dX-only backward of the LLaMA stack (weights frozen) + LoRA factor gradients.
 * lora_a_grads / lora_b_grads: arrays [layers*7] of bf16 destinations (entries or the arrays themselves may be NULL).
 * d_hidden: grad w.r.t. the output of lhrs_llama_fwd (post final norm).  Writes d_inputs_embeds (B,S,dim) bf16. */
size_t lhrs_llama_bwd_workspace_bytes(const LhrsLlamaWeights* w, int32_t B, int32_t S);
int lhrs_llama_bwd(const LhrsLlamaWeights* w, void* const* lora_a_grads, void* const* lora_b_grads, const void* d_hidden,
                   int32_t B, int32_t S, const uint8_t* key_mask, const void* stash, void* d_inputs_embeds, void* workspace,
                   size_t workspace_bytes, void* stream);
/* backward of lhrs_llama_fwd_ragged: d_hidden / d_inputs_embeds are [rows, dim]; workspace of lhrs_llama_bwd_workspace_bytes(B, S_max) */
int lhrs_llama_bwd_ragged(const LhrsLlamaWeights* w, void* const* lora_a_grads, void* const* lora_b_grads, const void* d_hidden,
                          int32_t B, int32_t S_max, int64_t rows, const int32_t* seq_off, const void* stash, void* d_inputs_embeds,
                          void* workspace, size_t workspace_bytes, void* stream);
int lhrs_lm_head_bwd(const LhrsLlamaWeights* w, const void* d_logits, int64_t rows, void* d_hidden, void* stream);

/* Full backward of the AttnPooler.  `grads` has the layout of the weight table; each pointer is the bf16 DESTINATION of
 * that parameter's gradient (overwritten), NULL to skip.  d_out: (B, nq, out_dim) rows ldo apart.  d_image: nullable. */
size_t lhrs_pooler_bwd_workspace_bytes(const LhrsPoolerWeights* w, int32_t B);
int lhrs_pooler_bwd(const LhrsPoolerWeights* w, const LhrsPoolerWeights* grads, const void* d_out, int64_t ldo, int32_t B,
                    const void* stash, void* d_image, void* workspace, size_t workspace_bytes, void* stream);

/* sum of squares of a bf16 flat gradient buffer -> out[0] (fp32); scratch >= 1024 floats */
int lhrs_grad_sumsq(const void* g, int64_t n, float* out, float* scratch, void* stream);
/* AdamW (decoupled decay, bias-corrected) on flat buffers: fp32 master/m/v, bf16 grad in, bf16 params out.
 * grad_scale multiplies the gradient first (1/world for a summed allreduce); if max_norm > 0 the gradient is clipped to
 * max_norm using gnorm_sq[0] (sum of squares of the UNSCALED gradient).  decay_mask: per-element 0/1 or NULL (= all 1). */
int lhrs_adamw_step(float* master, float* m, float* v, const void* grad, void* param_bf16, const float* decay_mask, int64_t n,
                    float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step, const float* gnorm_sq,
                    float max_norm, float grad_scale, void* stream);

/* Adan on flat buffers: the reference's stage-1 optimizer (`optimizer: adanp`, Config/multi_modal_stage1.yaml:89, built by timm
 * create_optimizer_v2 in lhrs/optimizer/build_optimizer.py:76-86; `adanp` = Adan(no_prox=0), `adanw` = Adan(no_prox=1)).  Betas in
 * timm's convention (defaults 0.98, 0.92, 0.99).  exp_avg / exp_avg_diff / exp_avg_sq / pre_grad are fp32 state ([n], zero-initialised;
 * pre_grad is taken from the first gradient when step == 1).  Clipping / grad_scale / decay_mask as in lhrs_adamw_step. */
int lhrs_adan_step(float* master, float* exp_avg, float* exp_avg_diff, float* exp_avg_sq, float* pre_grad, const void* grad,
                   void* param_bf16, const float* decay_mask, int64_t n, float lr, float beta1, float beta2, float beta3, float eps,
                   float weight_decay, int32_t step, int32_t no_prox, const float* gnorm_sq, float max_norm, float grad_scale,
                   void* stream);

/* Gradient exchange + optimizer step over NVLink peer memory — what the reference obtains from DeepSpeed ZeRO-2
 * (main_pretrain_stage1.py:28-85, 215-220; hook/deepspeed_hook.py:5-9): reduce-scatter of the gradients, optimizer state sharded
 * over the data-parallel ranks, all-gather of the updated parameters.  grads / params / norm_slots hold, for every rank of the
 * node, the PEER-MAPPED device pointer of that rank's flat bf16 gradient buffer, flat bf16 parameter buffer and float[world] norm
 * table (symmetric allocations; entry [rank] is the local buffer).  Rank r owns elements [slice_offset, slice_offset + slice_n).
 *   lhrs_p2p_reduce_slice  grad_sum[slice_n] (fp32) = sum over ranks of the slice (fixed rank order), and the slice's sum of
 *                          squares stored into norm_slots[q][rank] of every rank q.          scratch: >= 1024 floats
 *   lhrs_p2p_adamw_slice   clip by the global norm (sum of this rank's norm table) and grad_scale, AdamW on the slice's fp32
 *                          master / m / v (local, [slice_n]), bf16 result stored into params[q] of every rank q.
 * The caller separates the two calls — and brackets the pair — with a cross-rank barrier on the stream. */
typedef struct LhrsPeerExchange {
    int32_t world, rank;
    void* grads[16];
    void* params[16];
    float* norm_slots[16];
    int64_t slice_offset, slice_n;
    void* mc_grads;   /* optional NVLS multicast mapping of the gradient buffers: the slice is reduced in the NVSwitch (multimem.ld_reduce) */
    void* mc_params;  /* optional NVLS multicast mapping of the parameter buffers: the updated slice is broadcast by one multimem.st     */
} LhrsPeerExchange;
int lhrs_p2p_reduce_slice(const LhrsPeerExchange* x, float* grad_sum, float* scratch, void* stream);
int lhrs_p2p_adamw_slice(const LhrsPeerExchange* x, float* master, float* m, float* v, const float* grad_sum, const float* decay_mask,
                         float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step, float max_norm,
                         float grad_scale, void* stream);
/* same with the stage-1 optimizer (lhrs_adan_step's update) on the slice */
int lhrs_p2p_adan_slice(const LhrsPeerExchange* x, float* master, float* exp_avg, float* exp_avg_diff, float* exp_avg_sq, float* pre_grad,
                        const float* grad_sum, const float* decay_mask, float lr, float beta1, float beta2, float beta3, float eps,
                        float weight_decay, int32_t step, int32_t no_prox, float max_norm, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Single-sequence decode (HF generate loop reached from TextModal.generate, lhrs/models/text_modal.py:600-612, with the
 * generation-input rule of :36-60).  All buffers are caller-owned device memory (bf16 unless noted):
 *   xbuf [dim] residual stream of the token being fed, qkv [3*dim], obuf [dim], act [ffn], logits fp32 [vocab],
 *   attn_part / attn_count: scratch of the split ("flash decoding") attention, 4 context slices per head,
 *   part_val fp32 / part_idx int32 [>= 8*SMs] argmax partials, state int32[4] = {next token, ctx_len, tokens emitted, finished flag},
 *   tokens_out int32 [max_tokens].
 * first_token: logits/argmax of the prefill's last (post-norm) hidden row; sets ctx_len and feeds the token's embedding.
 * decode_step: one token through every layer (5 launches per layer), K/V appended to the paged cache at ctx_len.
 * greedy=1 commits argmax on the device (no host sync per token); greedy=0 leaves the fp32 logits for a host-side sampler,
 * which then calls lhrs_decode_commit_token(token, set_ctx): set_ctx = prompt length after first_token, -2 after a step.
 * ---------------------------------------------------------------------------------------------- */
typedef struct LhrsDecodeBuffers {
    void* xbuf; void* qkv; void* obuf; void* act;
    float* logits; float* part_val; int32_t* part_idx;
    int32_t* state; int32_t* tokens_out; int32_t max_tokens;
    float* attn_part;    /* fp32 [heads * 4 * 132]: per-slice attention partials (max, sum, un-normalised output) */
    int32_t* attn_count; /* int32 [heads], ZERO-initialised by the caller; the kernels return it to zero after every use */
} LhrsDecodeBuffers;
int lhrs_llama_first_token(const LhrsLlamaWeights* w, const void* hidden_last, int32_t ctx_len, const LhrsDecodeBuffers* b,
                           int32_t greedy, void* stream);
int lhrs_llama_decode_step(const LhrsLlamaWeights* w, const LhrsKvCache* kv, const LhrsDecodeBuffers* b, int32_t greedy,
                           int32_t max_ctx, void* stream);
int lhrs_decode_commit_token(const LhrsLlamaWeights* w, const LhrsDecodeBuffers* b, int32_t token, int32_t set_ctx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * CLIP image preprocessing on the device.  Replaces the per-image host work of the reference's input pipeline:
 * lhrs/Dataset/build_transform.py:43-45 (CLIPImageProcessor for every ViT config), called per sample from
 * lhrs/Dataset/cap_dataset.py:167-175 and cli_qa.py:119-126 — resize so the shortest edge is out_size (PIL BICUBIC on uint8:
 * Pillow ImagingResample, antialiased, 22-bit fixed point, horizontal then vertical pass), center crop out_size x out_size,
 * x/255, (x - mean) / std, channels first.  The uint8 stage is bit-exact with Pillow, the float stage with numpy's float32
 * arithmetic.  images: device uint8 [B, H, W, 3] (RGB, one geometry per call); out: [B, 3, out_size, out_size] bf16 (or fp32);
 * mean3 / std3: HOST float[3]; resized_u8: nullable device uint8 [B, out_size, out_size, 3] receiving the cropped resize result.
 * ---------------------------------------------------------------------------------------------- */
size_t lhrs_clip_preprocess_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t out_size);
int lhrs_clip_preprocess(const uint8_t* images, int32_t B, int32_t H, int32_t W, int32_t out_size, const float* mean3,
                         const float* std3, void* out, int32_t out_f32, uint8_t* resized_u8, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * Device-side token selection: the HF logits processors the reference's callers enable — repetition penalty,
 * temperature, top-k, top-p (cli_qa.py:176-186, lhrs_webui.py:206-218, main_vqa.py:205-214) — the multinomial draw, and the
 * EOS / token-suffix stop test of KeywordsStoppingCriteria (lhrs/utils/eval_utils.py:24-56), all without a host round trip.
 * Probabilities are handled as 2^-40 fixed-point integers and the draw is Philox4x32-10(seed, counter = draw index), so the
 * selection is order-independent and restated bit for bit in oracle/sampling.py (csrc/sampling.cuh states the rule).
 * With state[3] (raised on EOS / stop) the decode kernels of later, already enqueued steps return immediately.
 * ---------------------------------------------------------------------------------------------- */
typedef struct LhrsSampling {
    int32_t do_sample;         /* 0: argmax of the repetition-penalised logits (HF greedy); 1: sample                     */
    float temperature;         /* TemperatureLogitsWarper; <= 0 or 1 disables                                            */
    int32_t top_k;             /* TopKLogitsWarper; <= 0 disables                                                        */
    float top_p;               /* TopPLogitsWarper; <= 0 or >= 1 disables                                                */
    float repetition_penalty;  /* RepetitionPenaltyLogitsProcessor over the tokens generated so far; <= 0 or 1 disables  */
    int32_t eos_token;         /* -1: none                                                                               */
    uint64_t seed;             /* Philox key; draw t of a sequence uses counter t                                        */
    const uint64_t* seed_dev;  /* nullable: device uint64[1] read instead of `seed` (lets a captured CUDA graph be re-seeded) */
    const int32_t* stop_seqs;  /* device int32 [n_stop, stop_len], right-aligned, left-padded with -1; NULL: none        */
    int32_t n_stop, stop_len;
    float* work;               /* device scratch, >= vocab floats                                                        */
} LhrsSampling;
/* One selection from fp32 logits [vocab].  history: device int32 [n_history] (penalty set).  token_out: device int32[1].
 * debug4 (nullable, device uint64[4]): {sum of masses, kept mass, selection key, target} for parity tests. */
int lhrs_sample_logits(const float* logits, int32_t vocab, const int32_t* history, int32_t n_history, const LhrsSampling* s,
                       uint64_t draw, int32_t* token_out, uint64_t* debug4, void* stream);
/* lhrs_llama_first_token / lhrs_llama_decode_step with the selection done on the device by `s` (history = tokens emitted so
 * far, draw index = their count).  The host polls tokens_out / state[2] / state[3] whenever it likes. */
int lhrs_llama_first_token_sampled(const LhrsLlamaWeights* w, const void* hidden_last, int32_t ctx_len, const LhrsDecodeBuffers* b,
                                   const LhrsSampling* s, void* stream);
int lhrs_llama_decode_step_sampled(const LhrsLlamaWeights* w, const LhrsKvCache* kv, const LhrsDecodeBuffers* b,
                                   const LhrsSampling* s, int32_t max_ctx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LHRS_B200_H */
