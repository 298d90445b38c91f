// Layer loops of the three modules of the hot path, run inside the library so one C call = one module forward:
//   lhrs_vit_fwd     CLIP ViT-L/14 up to the last tap            (VisionModal.encode, rgb_vision_modal.py:166-184)
//   lhrs_pooler_fwd  AttnPooler, three query groups batched       (common_arch.py:134-173, 315-333)
//   lhrs_llama_fwd   LLaMA-2 decoder stack + final norm           (HF LlamaModel via text_modal.py:281-290)
// Every dense contraction is the tcgen05 GEMM of gemm_tcgen05.cu with its fused epilogues; the rest are the
// row kernels of elementwise.cu and the flash attention of attention.cu.
#include <stdlib.h>
#include "host_common.h"
#include "ptx.cuh"
#include "dropout.cuh"
#include "models_common.h"

namespace lhrs {

// ------------------------------------------------------------------ small kernels private to the orchestrators
// AttnPooler input staging: group-major query state and the constant kv = cat([initial queries, image tokens]).
__global__ void __launch_bounds__(128)
pooler_gather_kernel(const __nv_bfloat16* __restrict__ query, const __nv_bfloat16* __restrict__ img, PoolerGeom geo,
                     int B, int dim, __nv_bfloat16* __restrict__ xq, __nv_bfloat16* __restrict__ kv) {
    // grid: one block per kv row (group-major); blocks for j < stage[g] also write the query-state row
    const long long r = blockIdx.x;
    int g = 0;
    while (g + 1 < geo.G && r >= static_cast<long long>(B) * geo.kv_off[g + 1]) ++g;
    const long long rl = r - static_cast<long long>(B) * geo.kv_off[g];
    const int lkv = geo.stage[g] + geo.split[g];
    const int b = rl / lkv, j = rl % lkv;
    const int nchunks = dim / 8;
    const uint4* src;
    if (j < geo.stage[g]) src = reinterpret_cast<const uint4*>(query + static_cast<long long>(geo.q_off[g] + j) * dim);
    else src = reinterpret_cast<const uint4*>(img + (static_cast<long long>(b) * geo.img_total + geo.img_off[g] + (j - geo.stage[g])) * dim);
    uint4* dkv = reinterpret_cast<uint4*>(kv + r * dim);
    uint4* dq = nullptr;
    if (j < geo.stage[g])
        dq = reinterpret_cast<uint4*>(xq + (static_cast<long long>(B) * geo.q_off[g] + static_cast<long long>(b) * geo.stage[g] + j) * dim);
    for (int c = threadIdx.x; c < nchunks; c += blockDim.x) {
        const uint4 v = __ldg(src + c);
        dkv[c] = v;
        if (dq != nullptr) dq[c] = v;
    }
}

// group-major row (g, b, i) -> destination row: batch-major b*nq + q_off[g] + i, optionally through the caller's map
__global__ void pooler_rowmap_kernel(PoolerGeom geo, int B, const int* __restrict__ user_map, int* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B * geo.nq) return;
    int g = 0;
    while (g + 1 < geo.G && r >= B * geo.q_off[g + 1]) ++g;
    const int rl = r - B * geo.q_off[g];
    const int b = rl / geo.stage[g], i = rl % geo.stage[g];
    const int dst = b * geo.nq + geo.q_off[g] + i;
    out[r] = (user_map != nullptr) ? user_map[dst] : dst;
}

// prefill: copy this layer's K and V (post-RoPE) rows of the packed qkv buffer into the paged cache
__global__ void __launch_bounds__(128)
kv_write_kernel(const __nv_bfloat16* __restrict__ qkv, long long ldq, int dim, int S, int pos0, KvGeom kv, int layer) {
    // grid (S, B); one block copies K and V of one token (all heads)
    const int s = blockIdx.x, b = blockIdx.y;
    const int pos = pos0 + s;
    const int page = kv.block_table[b * kv.max_pages + pos / kv.page_size];
    const int slot = pos % kv.page_size;
    const __nv_bfloat16* src = qkv + (static_cast<long long>(b) * S + s) * ldq;
    const int hd = kv.head_dim, H = kv.heads;
    const int chunks_per_head = hd / 8;
    for (int which = 0; which < 2; ++which) {
        const uint4* sp = reinterpret_cast<const uint4*>(src + (which + 1) * dim);
        for (int c = threadIdx.x; c < H * chunks_per_head; c += blockDim.x) {
            const int h = c / chunks_per_head, cc = c % chunks_per_head;
            __nv_bfloat16* dst = kv_ptr(kv, layer, which, page, h, slot);
            reinterpret_cast<uint4*>(dst)[cc] = sp[c];
        }
    }
}

// true when the nproj LoRA-A factors starting at idx are laid out back to back ([nproj*r, in] contiguous) — the flat
// parameter buffer of training.SftStepper orders them that way so that the side GEMMs batch over projections
bool lora_a_adjacent(const void* const* arr, int idx, int nproj, long long elems_each) {
    if (arr == nullptr) return false;
    for (int p = 0; p < nproj; ++p) {
        if (arr[idx + p] == nullptr) return false;
        if (reinterpret_cast<const __nv_bfloat16*>(arr[idx + p]) != reinterpret_cast<const __nv_bfloat16*>(arr[idx]) + p * elems_each) return false;
    }
    return true;
}

int skinny_gemm(LhrsGemm& g, float* scratch, void* stream) {
    const long long tiles = ((g.M + 127) / 128) * (long long)((g.N + 127) / 128);
    const int nkb = (g.K + 63) / 64;
    // measured on B200: the split pays only when a launch would otherwise occupy < 1/3 of the SMs (its fixed cost is a
    // memset + a cast); 4 splits of >= 8 k-blocks each is the sweet spot for the token-reduction (TN) forms
    int sk = (tiles * 3 <= num_sms()) ? 4 : 1;
    if (sk > nkb / 8) sk = nkb / 8;
    if (scratch == nullptr || sk < 2 || g.ldd != g.N || g.d_f32 || g.bias[0] || g.residual || g.act != LHRS_ACT_NONE || g.A2 ||
        (static_cast<long long>(g.M) * g.N) % 4 != 0)
        return lhrs_gemm_bf16(&g, stream);
    void* dst = g.D;
    const long long n = static_cast<long long>(g.M) * g.N;
    LHRS_CUDA(cudaMemsetAsync(scratch, 0, n * sizeof(float), (cudaStream_t)stream));
    g.D = scratch; g.d_f32 = 1; g.split_k = sk;
    int rc = lhrs_gemm_bf16(&g, stream);
    g.D = dst; g.d_f32 = 0; g.split_k = 0;
    if (rc) return rc;
    return lhrs_cast_f32_bf16(scratch, dst, n, stream);
}

bool lora_stream_ok(const LhrsLlamaWeights* w, int in_dim, int out_dim, int nproj) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("LHRS_LORA_STREAM"); on = e ? atoi(e) : 1; }
    return on && w->lora_r == 16 && nproj >= 1 && nproj <= 3 && in_dim % 128 == 0 && out_dim % 128 == 0;
}

int lora_attach(LhrsGemm& g, const LhrsLlamaWeights* w, int layer, int first_proj, int nproj, const void* x, long long ldx,
                long long M, __nv_bfloat16* t_buf, float* scratch, __nv_bfloat16* drop_x, void* stream) {
    if (w->lora_r <= 0 || w->lora_a == nullptr || w->lora_b == nullptr) return LHRS_OK;
    const int r = w->lora_r;
    const int idx = layer * 7 + first_proj;
    int rc;
    const int drop_t = drop_threshold(w->lora_dropout);
    if (drop_t > 0) {
        // training forward with peft's input dropout: every LoRA module draws its own mask of the shared input, so the
        // projections of a group no longer share one pass: T_p = (s / keep) * (mask_p o x) · A_p^T from a masked copy of x
        LHRS_CHECK_ARG(drop_x != nullptr && ldx % 8 == 0, "lora_attach: dropout needs the masked-input scratch");
        const float alpha = w->lora_scale * drop_inv_keep(drop_t);
        if (lora_stream_ok(w, g.K, 128, nproj) && lora_a_adjacent(w->lora_a, idx, nproj, (long long)r * g.K)) {
            // rank 16, flat parameter layout: ONE pass over x for the whole group, masks applied to the mma fragments
            if ((rc = lhrs_lora_panel_dropout(x, ldx, M, g.K, w->lora_a[idx], g.K, nproj * r, w->lora_scale, w->lora_seed, idx,
                                              w->lora_dropout, t_buf, (long long)nproj * r, stream))) return rc;
        } else
        for (int p = 0; p < nproj; ++p) {
            if ((rc = lhrs_lora_dropout_mask(x, ldx, M, g.K, w->lora_seed, idx + p, w->lora_dropout, drop_x, g.K, stream))) return rc;
            if (lora_stream_ok(w, g.K, 128, 1)) {
                const void* wa[1] = {w->lora_a[idx + p]};
                if ((rc = lhrs_lora_panel(drop_x, g.K, M, g.K, wa, 1, 0, g.K, r, alpha, t_buf + p * r, (long long)nproj * r, stream))) return rc;
            } else {
                LhrsGemm t = gemm_desc(M, r, g.K, drop_x, g.K, w->lora_a[idx + p], g.K, t_buf + p * r, (long long)nproj * r);
                t.alpha = alpha;
                if ((rc = lhrs_gemm_bf16(&t, stream))) return rc;
            }
        }
    } else
    if (lora_a_adjacent(w->lora_a, idx, nproj, (long long)r * g.K)) {
        // T = (alpha/r) * x · [A_0; A_1; ..]^T in ONE skinny GEMM (one pass over x)
        if (lora_stream_ok(w, g.K, 128, nproj) && ldx % 8 == 0) {
            const void* wa[1] = {w->lora_a[idx]};
            if ((rc = lhrs_lora_panel(x, ldx, M, g.K, wa, 1, 0, g.K, nproj * r, w->lora_scale, t_buf, (long long)nproj * r, stream))) return rc;
        } else {
            LhrsGemm t = gemm_desc(M, nproj * r, g.K, x, ldx, w->lora_a[idx], g.K, t_buf, (long long)nproj * r);
            t.alpha = w->lora_scale;
            if ((rc = skinny_gemm(t, scratch, stream))) return rc;
        }
    } else {
        for (int p = 0; p < nproj; ++p) {
            LhrsGemm t = gemm_desc(M, r, g.K, x, ldx, w->lora_a[idx + p], g.K, t_buf + p * r, (long long)nproj * r);
            t.alpha = w->lora_scale;
            if ((rc = lhrs_gemm_bf16(&t, stream))) return rc;
        }
    }
    for (int p = 0; p < nproj; ++p) g.B2[p] = w->lora_b[idx + p];
    g.A2 = t_buf; g.lda2 = (long long)nproj * r; g.ldb2 = r; g.ext_k = r;
    return LHRS_OK;
}

}  // namespace lhrs

using namespace lhrs;
typedef __nv_bfloat16 bf16;

// ================================================================================================ ViT
namespace {
struct VitBufs { bf16 *patches, *pe, *x, *h, *qkv, *o, *f; };
VitBufs vit_plan(Arena& a, const LhrsVitWeights* w, int B) {
    const long long P = (long long)(w->image / w->patch) * (w->image / w->patch);
    const long long M = B * (P + 1);
    VitBufs v;
    v.patches = a.take<bf16>(B * P * w->kpad);
    v.pe = a.take<bf16>(B * P * w->dim);
    v.x = a.take<bf16>(M * w->dim);
    v.h = a.take<bf16>(M * w->dim);
    v.qkv = a.take<bf16>(M * 3 * w->dim);
    v.o = a.take<bf16>(M * w->dim);
    v.f = a.take<bf16>(M * w->ffn);
    return v;
}
}  // namespace

extern "C" size_t lhrs_vit_workspace_bytes(const LhrsVitWeights* w, int32_t B) {
    Arena a(nullptr, 0);
    vit_plan(a, w, B);
    return a.used();
}

extern "C" int lhrs_vit_fwd(const LhrsVitWeights* w, const void* pixels, int32_t B, const int32_t* taps, int32_t n_taps,
                            void* out, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LHRS_CHECK_ARG(w && pixels && out && taps && n_taps > 0 && B > 0, "lhrs_vit_fwd: null/empty");
    LHRS_CHECK_ARG(w->dim % (w->heads * 8) == 0 && w->dim / w->heads == 64, "lhrs_vit_fwd: head_dim must be 64");
    LHRS_CHECK_ARG(taps[n_taps - 1] == w->num_layers, "lhrs_vit_fwd: last tap %d must equal num_layers %d", taps[n_taps - 1], w->num_layers);
    Arena a(workspace, workspace_bytes);
    VitBufs v = vit_plan(a, w, B);
    LHRS_CHECK_ARG(a.fits(), "lhrs_vit_fwd: workspace too small (%zu < %zu)", workspace_bytes, a.used());
    const int P = (w->image / w->patch) * (w->image / w->patch);
    const int T = P + 1, D = w->dim;
    const long long M = (long long)B * T;
    int rc;
    if ((rc = lhrs_vit_im2col(pixels, v.patches, B, w->image, w->image, w->patch, w->kpad, stream))) return rc;
    {
        LhrsGemm g = gemm_desc(B * P, D, w->kpad, v.patches, w->kpad, w->patch_w, w->kpad, v.pe, D);
        if ((rc = lhrs_gemm_bf16(&g, stream))) return rc;
    }
    if ((rc = lhrs_vit_embed_ln(v.pe, w->cls, w->pos, w->pre_ln_w, w->pre_ln_b, v.x, B, P, D, w->eps, stream))) return rc;

    int next_tap = 0;
    auto emit_tap = [&](int hidden_index) -> int {
        while (next_tap < n_taps && taps[next_tap] == hidden_index) {
            // hidden_states[s][:, 1:, :] -> out[:, tap*P:(tap+1)*P, :]; one 2-D copy, a "row" = one image's P tokens
            cudaError_t e = cudaMemcpy2DAsync(reinterpret_cast<bf16*>(out) + (long long)next_tap * P * D, (size_t)n_taps * P * D * 2,
                                              v.x + D, (size_t)T * D * 2, (size_t)P * D * 2, B, cudaMemcpyDeviceToDevice, stream);
            if (e != cudaSuccess) { set_error("vit tap copy failed: %s", cudaGetErrorString(e)); return LHRS_ERR_CUDA; }
            ++next_tap;
        }
        return LHRS_OK;
    };
    if ((rc = emit_tap(0))) return rc;
    for (int l = 0; l < w->num_layers; ++l) {
        if ((rc = lhrs_layernorm_fwd(v.x, D, w->ln1_w[l], w->ln1_b[l], v.h, D, nullptr, nullptr, M, D, w->eps, stream))) return rc;
        {
            LhrsGemm g = gemm_desc(M, 3 * D, D, v.h, D, w->q_w[l], D, v.qkv, 3 * D);
            g.B[1] = w->k_w[l]; g.B[2] = w->v_w[l]; g.num_b = 3; g.seg_rows = D;
            g.bias[0] = w->q_b[l]; g.bias[1] = w->k_b[l]; g.bias[2] = w->v_b[l];
            if ((rc = lhrs_gemm_bf16(&g, stream))) return rc;
        }
        {
            LhrsAttention at = attn_desc(v.qkv, v.qkv + D, v.qkv + 2 * D, 3 * D, (long long)T * 3 * D, v.o, D, (long long)T * D,
                                         B, w->heads, T, T, 64, 0);
            if ((rc = lhrs_attention_fwd(&at, stream))) return rc;
        }
        {
            LhrsGemm g = gemm_desc(M, D, D, v.o, D, w->o_w[l], D, v.x, D);
            g.bias[0] = w->o_b[l]; g.residual = v.x; g.ldr = D;
            if ((rc = lhrs_gemm_bf16(&g, stream))) return rc;
        }
        if ((rc = lhrs_layernorm_fwd(v.x, D, w->ln2_w[l], w->ln2_b[l], v.h, D, nullptr, nullptr, M, D, w->eps, stream))) return rc;
        {
            LhrsGemm g = gemm_desc(M, w->ffn, D, v.h, D, w->fc1_w[l], D, v.f, w->ffn);
            g.bias[0] = w->fc1_b[l]; g.act = LHRS_ACT_QUICK_GELU;
            if ((rc = lhrs_gemm_bf16(&g, stream))) return rc;
        }
        {
            LhrsGemm g = gemm_desc(M, D, w->ffn, v.f, w->ffn, w->fc2_w[l], w->ffn, v.x, D);
            g.bias[0] = w->fc2_b[l]; g.residual = v.x; g.ldr = D;
            if ((rc = lhrs_gemm_bf16(&g, stream))) return rc;
        }
        if ((rc = emit_tap(l + 1))) return rc;
    }
    return LHRS_OK;
}

// ================================================================================================ AttnPooler
namespace {
struct PoolBufs { bf16 *xq, *kv_raw, *kvn, *kvp, *hq, *qp, *ao, *f; int* rowmap; };
PoolBufs pool_plan(Arena& a, const LhrsPoolerWeights* w, int B, const PoolerGeom& geo) {
    PoolBufs p;
    const long long RQ = (long long)B * geo.nq, RKV = (long long)B * geo.kv_total;
    p.xq = a.take<bf16>(RQ * w->dim);
    p.kv_raw = a.take<bf16>(RKV * w->dim);
    p.kvn = a.take<bf16>(RKV * w->dim);
    p.kvp = a.take<bf16>(RKV * 2 * w->dim);
    p.hq = a.take<bf16>(RQ * w->dim);
    p.qp = a.take<bf16>(RQ * w->dim);
    p.ao = a.take<bf16>(RQ * w->dim);
    p.f = a.take<bf16>(RQ * w->ffn);
    p.rowmap = a.take<int>(RQ);
    return p;
}
}  // namespace

extern "C" size_t lhrs_pooler_workspace_bytes(const LhrsPoolerWeights* w, int32_t B) {
    Arena a(nullptr, 0);
    pool_plan(a, w, B, pooler_geom(w));
    return a.used();
}

extern "C" size_t lhrs_pooler_stash_bytes(const LhrsPoolerWeights* w, int32_t B) {
    Arena a(nullptr, 0);
    pooler_stash_plan(a, w, B, pooler_geom(w));
    return a.used();
}

extern "C" int lhrs_pooler_fwd(const LhrsPoolerWeights* w, const void* image_embs, int32_t B, void* out, int64_t ldo,
                               const int32_t* row_map, void* stash, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LHRS_CHECK_ARG(w && image_embs && out && B > 0, "lhrs_pooler_fwd: null/empty");
    LHRS_CHECK_ARG(w->num_groups >= 1 && w->num_groups <= 4 && w->dim / w->heads == 64, "lhrs_pooler_fwd: groups %d / head_dim", w->num_groups);
    const PoolerGeom geo = pooler_geom(w);
    Arena a(workspace, workspace_bytes);
    PoolBufs p = pool_plan(a, w, B, geo);
    LHRS_CHECK_ARG(a.fits(), "lhrs_pooler_fwd: workspace too small (%zu < %zu)", workspace_bytes, a.used());
    PoolerStash st;
    if (stash != nullptr) {
        Arena sa(stash, (size_t)-1);
        st = pooler_stash_plan(sa, w, B, geo);
    }
    const int D = w->dim, F = w->ffn;
    const long long RQ = (long long)B * geo.nq, RKV = (long long)B * geo.kv_total;
    int rc;
    bf16* kv_raw = stash ? st.kv_raw : p.kv_raw;
    pooler_gather_kernel<<<(unsigned)RKV, 128, 0, stream>>>((const bf16*)w->query, (const bf16*)image_embs, geo, B, D, p.xq, kv_raw);
    LHRS_LAUNCH_CHECK("pooler_gather_kernel");
    pooler_rowmap_kernel<<<(unsigned)((RQ + 255) / 256), 256, 0, stream>>>(geo, B, row_map, p.rowmap);
    LHRS_LAUNCH_CHECK("pooler_rowmap_kernel");

    for (int l = 0; l < w->num_layers; ++l) {
        PoolerLayerStash* ls = stash ? &st.layer[l] : nullptr;
        bf16* kvn = ls ? ls->kvn : p.kvn;
        bf16* kvp = ls ? ls->kvp : p.kvp;
        bf16* hq = ls ? ls->hq : p.hq;
        bf16* qp = ls ? ls->qp : p.qp;
        bf16* ao = ls ? ls->ao : p.ao;
        if (ls) LHRS_CUDA(cudaMemcpyAsync(ls->x_in, p.xq, RQ * D * 2, cudaMemcpyDeviceToDevice, stream));
        if ((rc = lhrs_layernorm_fwd(kv_raw, D, w->lnkv_w[l], w->lnkv_b[l], kvn, D, ls ? ls->kv_mean : nullptr, ls ? ls->kv_rstd : nullptr, RKV, D, w->eps, stream))) return rc;
        {
            const bf16* in_w = (const bf16*)w->in_w[l];
            const bf16* in_b = (const bf16*)w->in_b[l];
            LhrsGemm g = gemm_desc(RKV, 2 * D, D, kvn, D, in_w + (long long)D * D, D, kvp, 2 * D);
            g.bias[0] = in_b + D;
            if ((rc = lhrs_gemm_bf16(&g, stream))) return rc;
            if ((rc = lhrs_layernorm_fwd(p.xq, D, w->ln1_w[l], w->ln1_b[l], hq, D, ls ? ls->q_mean : nullptr, ls ? ls->q_rstd : nullptr, RQ, D, w->eps, stream))) return rc;
            LhrsGemm gq = gemm_desc(RQ, D, D, hq, D, in_w, D, qp, D);
            gq.bias[0] = in_b;
            if ((rc = lhrs_gemm_bf16(&gq, stream))) return rc;
        }
        {   // the cross-attention of all query groups as ONE launch (group-major rows: each group is its own (B, Lq) x (B, Lkv) problem)
            LhrsAttention at[4];
            for (int g = 0; g < geo.G; ++g) {
                const int Lq = geo.stage[g], Lkv = geo.stage[g] + geo.split[g];
                const bf16* q = qp + (long long)B * geo.q_off[g] * D;
                const bf16* k = kvp + (long long)B * geo.kv_off[g] * 2 * D;
                bf16* o = ao + (long long)B * geo.q_off[g] * D;
                at[g] = attn_desc(q, k, k + D, D, (long long)Lq * D, o, D, (long long)Lq * D, B, w->heads, Lq, Lkv, 64, 0);
                at[g].k_rs = at[g].v_rs = 2 * D; at[g].k_bs = at[g].v_bs = (long long)Lkv * 2 * D;
                if (ls) at[g].lse = ls->lse + (long long)B * geo.q_off[g] * w->heads;
            }
            if (geo.G <= 3) {
                if ((rc = lhrs_attention_fwd_grouped(at, geo.G, stream))) return rc;
            } else {
                for (int g = 0; g < geo.G; ++g) if ((rc = lhrs_attention_fwd(&at[g], stream))) return rc;
            }
        }
        {
            LhrsGemm g = gemm_desc(RQ, D, D, ao, D, w->ao_w[l], D, p.xq, D);
            g.bias[0] = w->ao_b[l]; g.residual = p.xq; g.ldr = D;
            if ((rc = lhrs_gemm_bf16(&g, stream))) return rc;
        }
        bf16* h2 = ls ? ls->h2 : p.hq;
        if (ls) LHRS_CUDA(cudaMemcpyAsync(ls->x_mid, p.xq, RQ * D * 2, cudaMemcpyDeviceToDevice, stream));
        if ((rc = lhrs_layernorm_fwd(p.xq, D, w->ln2_w[l], w->ln2_b[l], h2, D, ls ? ls->m_mean : nullptr, ls ? ls->m_rstd : nullptr, RQ, D, w->eps, stream))) return rc;
        bf16* f = ls ? ls->f_act : p.f;
        {
            LhrsGemm g = gemm_desc(RQ, F, D, h2, D, w->fc_w[l], D, f, F);
            g.bias[0] = w->fc_b[l]; g.act = LHRS_ACT_GELU_ERF;
            if (ls) g.pre_gate = ls->f_pre;  // LINEAR epilogue: pre-activation copy for the GELU backward
            if ((rc = lhrs_gemm_bf16(&g, stream))) return rc;
        }
        {
            LhrsGemm g = gemm_desc(RQ, D, F, f, F, w->pj_w[l], F, p.xq, D);
            g.bias[0] = w->pj_b[l]; g.residual = p.xq; g.ldr = D;
            if ((rc = lhrs_gemm_bf16(&g, stream))) return rc;
        }
    }
    if (stash) LHRS_CUDA(cudaMemcpyAsync(st.x_final, p.xq, RQ * D * 2, cudaMemcpyDeviceToDevice, stream));
    {
        LhrsGemm g = gemm_desc(RQ, w->out_dim, D, p.xq, D, w->out_w, D, out, ldo);
        g.bias[0] = w->out_b; g.row_map = p.rowmap;
        if ((rc = lhrs_gemm_bf16(&g, stream))) return rc;
    }
    return LHRS_OK;
}

// ================================================================================================ LLaMA
namespace {
struct LlamaBufs { bf16 *x, *h, *qkv, *o, *act, *lora_t, *drop_x; float* skinny; };
LlamaBufs llama_plan(Arena& a, const LhrsLlamaWeights* w, long long M, bool have_stash) {
    LlamaBufs b;
    b.x = a.take<bf16>(M * w->dim);
    b.h = a.take<bf16>(M * w->dim);
    b.qkv = have_stash ? nullptr : a.take<bf16>(M * 3 * w->dim);
    b.o = have_stash ? nullptr : a.take<bf16>(M * w->dim);
    b.act = have_stash ? nullptr : a.take<bf16>(M * w->ffn);
    b.lora_t = (w->lora_r > 0) ? a.take<bf16>(M * 3 * w->lora_r) : nullptr;
    b.skinny = (w->lora_r > 0) ? a.take<float>(skinny_scratch_elems(w, M)) : nullptr;
    b.drop_x = (w->lora_r > 0 && w->lora_dropout > 0.f) ? a.take<bf16>(M * (w->ffn > w->dim ? w->ffn : w->dim)) : nullptr;
    return b;
}
}  // namespace

extern "C" size_t lhrs_llama_workspace_bytes(const LhrsLlamaWeights* w, int32_t B, int32_t S) {
    Arena a(nullptr, 0);
    llama_plan(a, w, (long long)B * S, false);
    return a.used();
}

extern "C" size_t lhrs_llama_stash_bytes(const LhrsLlamaWeights* w, int32_t B, int32_t S) {
    Arena a(nullptr, 0);
    llama_stash_plan(a, w, B, S);
    return a.used();
}

// Dense batch: M = B * S rows, row b * S + s.  Ragged batch (seq_off != nullptr): M = seq_off[B] rows, the sequences back to back
// without their right padding (S is then the longest length); `positions` [M] gives each row's index inside its sequence.
static int llama_fwd_impl(const LhrsLlamaWeights* w, const void* inputs_embeds, int32_t B, int32_t S, long long M,
                          const uint8_t* key_mask, const int32_t* seq_off, const int32_t* positions, void* hidden_out, void* stash,
                          const LhrsKvCache* kv, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    LHRS_CHECK_ARG(w && inputs_embeds && hidden_out && B > 0 && S > 0, "lhrs_llama_fwd: null/empty");
    LHRS_CHECK_ARG(w->dim / w->heads == 128 && w->dim % 256 == 0 && w->ffn % 128 == 0, "lhrs_llama_fwd: needs head_dim 128, dim %% 256 == 0, ffn %% 128 == 0");
    LHRS_CHECK_ARG(S <= w->max_pos, "lhrs_llama_fwd: S=%d exceeds max_position_embeddings=%d", S, w->max_pos);
    const int D = w->dim, F = w->ffn;
    Arena a(workspace, workspace_bytes);
    LlamaBufs b = llama_plan(a, w, (long long)B * S, stash != nullptr);   // (a ragged batch has fewer rows)
    LHRS_CHECK_ARG(a.fits(), "lhrs_llama_fwd: workspace too small (%zu < %zu)", workspace_bytes, a.used());
    LlamaStash st;
    if (stash) {
        Arena sa(stash, (size_t)-1);
        st = llama_stash_plan(sa, w, B, S);
    }
    int rc;
    // Residual stream: in inference it is updated in place in b.x; with a stash every GEMM that produces the next state of
    // the stream writes it straight into its stash slot (x_in[l] -> x_mid[l] -> x_in[l+1] ... -> x_final), no copies.
    const bf16* x_cur = stash ? st.layer[0].x_in : b.x;
    LHRS_CUDA(cudaMemcpyAsync(const_cast<bf16*>(x_cur), inputs_embeds, M * D * 2, cudaMemcpyDeviceToDevice, stream));
    for (int l = 0; l < w->num_layers; ++l) {
        LlamaLayerStash* ls = stash ? &st.layer[l] : nullptr;
        bf16* qkv = ls ? ls->qkv : b.qkv;
        bf16* o = ls ? ls->o : b.o;
        bf16* x_mid = ls ? ls->x_mid : b.x;
        bf16* x_next = !stash ? b.x : (l + 1 < w->num_layers ? st.layer[l + 1].x_in : st.x_final);
        bf16* h1 = (ls && ls->h1) ? ls->h1 : b.h;   // normalised inputs go straight into the stash when backward will need them
        bf16* h2 = (ls && ls->h2) ? ls->h2 : b.h;
        if ((rc = lhrs_rmsnorm_fwd(x_cur, w->ln1_w[l], h1, ls ? ls->rstd1 : nullptr, M, D, w->eps, stream))) return rc;
        {
            LhrsGemm g = gemm_desc(M, 3 * D, D, h1, D, w->q_w[l], D, qkv, 3 * D);
            g.B[1] = w->k_w[l]; g.B[2] = w->v_w[l]; g.num_b = 3; g.seg_rows = D;
            g.epilogue = LHRS_EPI_ROPE; g.rope_cos = w->rope_cos; g.rope_sin = w->rope_sin; g.rope_seq_len = S;
            g.positions = positions;       // ragged batch: the row index no longer gives the position
            if ((rc = lora_attach(g, w, l, 0, 3, h1, D, M, ls ? ls->lora_t[0] : b.lora_t, b.skinny, b.drop_x, stream))) return rc;
            if ((rc = lhrs_gemm_bf16(&g, stream))) return rc;
        }
        if (kv != nullptr) {
            kv_write_kernel<<<dim3(S, B), 128, 0, stream>>>(qkv, 3 * D, D, S, 0, *kv, l);
            LHRS_LAUNCH_CHECK("kv_write_kernel");
        }
        {
            LhrsAttention at = attn_desc(qkv, qkv + D, qkv + 2 * D, 3 * D, (long long)S * 3 * D, o, D, (long long)S * D, B,
                                         w->heads, S, S, 128, 1);
            at.key_mask = key_mask;
            at.seq_off = seq_off; at.total_rows = seq_off ? M : 0;
            if (ls) at.lse = ls->lse;
            if ((rc = lhrs_attention_fwd(&at, stream))) return rc;
        }
        {
            LhrsGemm g = gemm_desc(M, D, D, o, D, w->o_w[l], D, x_mid, D);
            g.residual = x_cur; g.ldr = D;
            if ((rc = lora_attach(g, w, l, 3, 1, o, D, M, ls ? ls->lora_t[1] : b.lora_t, b.skinny, b.drop_x, stream))) return rc;
            if ((rc = lhrs_gemm_bf16(&g, stream))) return rc;
        }
        if ((rc = lhrs_rmsnorm_fwd(x_mid, w->ln2_w[l], h2, ls ? ls->rstd2 : nullptr, M, D, w->eps, stream))) return rc;
        bf16* act = ls ? ls->act : b.act;
        {
            LhrsGemm g = gemm_desc(M, 2 * F, D, h2, D, w->gate_w[l], D, act, F);
            g.B[1] = w->up_w[l]; g.num_b = 2; g.seg_rows = F; g.epilogue = LHRS_EPI_SWIGLU;
            if (ls) { g.pre_gate = ls->pre_gate; g.pre_up = ls->pre_up; }
            if ((rc = lora_attach(g, w, l, 4, 2, h2, D, M, ls ? ls->lora_t[2] : b.lora_t, b.skinny, b.drop_x, stream))) return rc;
            if ((rc = lhrs_gemm_bf16(&g, stream))) return rc;
        }
        {
            LhrsGemm g = gemm_desc(M, D, F, act, F, w->down_w[l], F, x_next, D);
            g.residual = x_mid; g.ldr = D;
            if ((rc = lora_attach(g, w, l, 6, 1, act, F, M, ls ? ls->lora_t[3] : b.lora_t, b.skinny, b.drop_x, stream))) return rc;
            if ((rc = lhrs_gemm_bf16(&g, stream))) return rc;
        }
        x_cur = x_next;
    }
    if ((rc = lhrs_rmsnorm_fwd(x_cur, w->norm_w, hidden_out, stash ? st.rstd_final : nullptr, M, D, w->eps, stream))) return rc;
    return LHRS_OK;
}

extern "C" int lhrs_llama_fwd(const LhrsLlamaWeights* w, const void* inputs_embeds, int32_t B, int32_t S,
                              const uint8_t* key_mask, void* hidden_out, void* stash, const LhrsKvCache* kv, void* workspace,
                              size_t workspace_bytes, void* stream_) {
    return llama_fwd_impl(w, inputs_embeds, B, S, (long long)B * S, key_mask, nullptr, nullptr, hidden_out, stash, kv, workspace,
                          workspace_bytes, stream_);
}

extern "C" int lhrs_llama_fwd_ragged(const LhrsLlamaWeights* w, const void* inputs_embeds, int32_t B, int32_t S_max, int64_t rows,
                                     const int32_t* seq_off, const int32_t* positions, void* hidden_out, void* stash,
                                     void* workspace, size_t workspace_bytes, void* stream_) {
    LHRS_CHECK_ARG(seq_off && positions && rows > 0 && rows <= (long long)B * S_max, "lhrs_llama_fwd_ragged: seq_off / positions / rows (<= B * S_max)");
    LHRS_CHECK_ARG(S_max >= 128, "lhrs_llama_fwd_ragged: the ragged attention kernels need a longest length >= 128 (got %d)", S_max);
    return llama_fwd_impl(w, inputs_embeds, B, S_max, rows, nullptr, seq_off, positions, hidden_out, stash, nullptr, workspace,
                          workspace_bytes, stream_);
}

extern "C" int lhrs_lm_head(const LhrsLlamaWeights* w, const void* hidden, int64_t rows, void* logits, void* stream_) {
    LHRS_CHECK_ARG(w && hidden && logits && rows > 0, "lhrs_lm_head: null/empty");
    LhrsGemm g = gemm_desc(rows, w->vocab, w->dim, hidden, w->dim, w->lm_head, w->dim, logits, w->vocab);
    return lhrs_gemm_bf16(&g, stream_);
}
