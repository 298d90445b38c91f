// Shared between the forward and backward orchestrators: workspace arena, geometry helpers, activation-stash layouts.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include <string.h>

#include "host_common.h"

namespace lhrs {

// Bump allocator over a caller-provided buffer.  With base == nullptr it only measures (used()).
struct Arena {
    uint8_t* base;
    size_t cap, off;
    Arena(void* b, size_t c) : base(reinterpret_cast<uint8_t*>(b)), cap(c), off(0) {}
    template <typename T>
    T* take(long long n) {
        off = (off + 255) & ~size_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += static_cast<size_t>(n) * sizeof(T);
        return p;
    }
    size_t used() const { return (off + 255) & ~size_t(255); }
    bool fits() const { return base != nullptr && off <= cap; }
};

inline LhrsGemm gemm_desc(long long M, int N, int K, const void* A, long long lda, const void* B0, long long ldb, void* D,
                          long long ldd) {
    LhrsGemm g;
    memset(&g, 0, sizeof(g));
    g.M = static_cast<int32_t>(M); g.N = N; g.K = K;
    g.A = A; g.lda = lda;
    g.B[0] = B0; g.num_b = 1; g.seg_rows = N; g.ldb = ldb;
    g.epilogue = LHRS_EPI_LINEAR; g.act = LHRS_ACT_NONE; g.alpha = 1.0f;
    g.D = D; g.ldd = ldd;
    return g;
}

// q/k/v/o given as (B, S, H, hd) views with a common row stride for q,k,v (packed buffers) and head stride hd
inline LhrsAttention attn_desc(const void* q, const void* k, const void* v, long long qkv_rs, long long qkv_bs, void* o,
                               long long o_rs, long long o_bs, int B, int H, int Sq, int Skv, int hd, int causal) {
    LhrsAttention a;
    memset(&a, 0, sizeof(a));
    a.q = q; a.k = k; a.v = v; a.o = o;
    a.q_rs = a.k_rs = a.v_rs = qkv_rs; a.q_bs = a.k_bs = a.v_bs = qkv_bs;
    a.q_hs = a.k_hs = a.v_hs = a.o_hs = hd;
    a.o_rs = o_rs; a.o_bs = o_bs;
    a.B = B; a.H = H; a.Sq = Sq; a.Skv = Skv; a.head_dim = hd; a.causal = causal;
    a.scale = 1.0f / sqrtf(static_cast<float>(hd));
    return a;
}

// ------------------------------------------------------------------ AttnPooler geometry (group-major row layout)
struct PoolerGeom {
    int G;
    int stage[4], split[4];
    int q_off[5];    // prefix sums of stage (per image)
    int kv_off[5];   // prefix sums of stage+split (per image)
    int img_off[5];  // prefix sums of split (per image)
    int nq, kv_total, img_total;
};
inline PoolerGeom pooler_geom(const LhrsPoolerWeights* w) {
    PoolerGeom g;
    memset(&g, 0, sizeof(g));
    g.G = w->num_groups;
    for (int i = 0; i < g.G; ++i) {
        g.stage[i] = w->stage_num[i];
        g.split[i] = w->split_part[i];
        g.q_off[i + 1] = g.q_off[i] + g.stage[i];
        g.kv_off[i + 1] = g.kv_off[i] + g.stage[i] + g.split[i];
        g.img_off[i + 1] = g.img_off[i] + g.split[i];
    }
    g.nq = g.q_off[g.G]; g.kv_total = g.kv_off[g.G]; g.img_total = g.img_off[g.G];
    return g;
}

struct PoolerLayerStash {
    __nv_bfloat16 *x_in, *kvn, *kvp, *hq, *qp, *ao, *x_mid, *h2, *f_pre, *f_act;
    float *kv_mean, *kv_rstd, *q_mean, *q_rstd, *m_mean, *m_rstd, *lse;
};
struct PoolerStash {
    __nv_bfloat16* kv_raw;
    PoolerLayerStash layer[16];
    __nv_bfloat16* x_final;
};
inline PoolerStash pooler_stash_plan(Arena& a, const LhrsPoolerWeights* w, int B, const PoolerGeom& geo) {
    PoolerStash s;
    const long long RQ = (long long)B * geo.nq, RKV = (long long)B * geo.kv_total;
    const int D = w->dim, F = w->ffn;
    s.kv_raw = a.take<__nv_bfloat16>(RKV * D);
    for (int l = 0; l < w->num_layers && l < 16; ++l) {
        PoolerLayerStash& t = s.layer[l];
        t.x_in = a.take<__nv_bfloat16>(RQ * D);
        t.kvn = a.take<__nv_bfloat16>(RKV * D);
        t.kvp = a.take<__nv_bfloat16>(RKV * 2 * D);
        t.hq = a.take<__nv_bfloat16>(RQ * D);
        t.qp = a.take<__nv_bfloat16>(RQ * D);
        t.ao = a.take<__nv_bfloat16>(RQ * D);
        t.x_mid = a.take<__nv_bfloat16>(RQ * D);
        t.h2 = a.take<__nv_bfloat16>(RQ * D);
        t.f_pre = a.take<__nv_bfloat16>(RQ * F);
        t.f_act = a.take<__nv_bfloat16>(RQ * F);
        t.kv_mean = a.take<float>(RKV); t.kv_rstd = a.take<float>(RKV);
        t.q_mean = a.take<float>(RQ); t.q_rstd = a.take<float>(RQ);
        t.m_mean = a.take<float>(RQ); t.m_rstd = a.take<float>(RQ);
        t.lse = a.take<float>(RQ * w->heads);
    }
    s.x_final = a.take<__nv_bfloat16>(RQ * D);
    return s;
}

// ------------------------------------------------------------------ LLaMA activation stash (kept for dX / LoRA backward)
struct LlamaLayerStash {
    __nv_bfloat16 *x_in, *x_mid, *qkv, *o, *pre_gate, *pre_up, *act;
    float *rstd1, *rstd2, *lse;
    __nv_bfloat16* lora_t[4];   // T = s*x*A^T of the four projection groups {qkv, o, gate/up, down} (LoRA only)
    __nv_bfloat16 *h1, *h2;     // RMSNorm outputs feeding qkv / gate-up (LoRA only: dA = dT^T h needs them; saves a recompute)
};
struct LlamaStash {
    LlamaLayerStash layer[80];
    __nv_bfloat16* x_final;
    float* rstd_final;
};
inline LlamaStash llama_stash_plan(Arena& a, const LhrsLlamaWeights* w, int B, int S) {
    LlamaStash s;
    const long long M = (long long)B * S;
    const int D = w->dim, F = w->ffn;
    for (int l = 0; l < w->num_layers && l < 80; ++l) {
        LlamaLayerStash& t = s.layer[l];
        t.x_in = a.take<__nv_bfloat16>(M * D);
        t.x_mid = a.take<__nv_bfloat16>(M * D);
        t.qkv = a.take<__nv_bfloat16>(M * 3 * D);
        t.o = a.take<__nv_bfloat16>(M * D);
        t.pre_gate = a.take<__nv_bfloat16>(M * F);
        t.pre_up = a.take<__nv_bfloat16>(M * F);
        t.act = a.take<__nv_bfloat16>(M * F);
        t.rstd1 = a.take<float>(M);
        t.rstd2 = a.take<float>(M);
        t.lse = a.take<float>(M * w->heads);
        const int np[4] = {3, 1, 2, 1};
        for (int g = 0; g < 4; ++g) t.lora_t[g] = (w->lora_r > 0) ? a.take<__nv_bfloat16>(M * np[g] * w->lora_r) : nullptr;
        t.h1 = (w->lora_r > 0) ? a.take<__nv_bfloat16>(M * D) : nullptr;
        t.h2 = (w->lora_r > 0) ? a.take<__nv_bfloat16>(M * D) : nullptr;
    }
    s.x_final = a.take<__nv_bfloat16>(M * D);
    s.rstd_final = a.take<float>(M);
    return s;
}

// ------------------------------------------------------------------ paged KV cache (see LhrsKvCache in lhrs_b200.h)
typedef LhrsKvCache KvGeom;
// element (layer, k|v, page, head, slot, 0)
__host__ __device__ inline __nv_bfloat16* kv_ptr(const KvGeom& kv, int layer, int which, int page, int head, int slot) {
    const long long per_page = (long long)kv.heads * kv.page_size * kv.head_dim;
    const long long off = (((long long)layer * 2 + which) * kv.num_pages + page) * per_page +
                          ((long long)head * kv.page_size + slot) * kv.head_dim;
    return reinterpret_cast<__nv_bfloat16*>(kv.pool) + off;
}

// LoRA: compute T = scale * x · A^T for `nproj` projections sharing the input x and attach the K-extension
// (A2 = T, B2 = lora_B) to the parent GEMM so that B·T accumulates into the same TMEM tile.  No-op when LoRA is off.
int lora_attach(LhrsGemm& g, const LhrsLlamaWeights* w, int layer, int first_proj, int nproj, const void* x, long long ldx,
                long long M, __nv_bfloat16* t_buf, float* scratch, __nv_bfloat16* drop_x, void* stream);

bool lora_a_adjacent(const void* const* arr, int idx, int nproj, long long elems_each);

// Run a skinny GEMM (few output tiles, long K) with a K split sized to fill the SMs: fp32 atomics into `scratch`
// (>= M*N floats), then one cast into the bf16 destination (which must be contiguous, ldd == N).  Falls back to the
// plain launch when the problem already has enough tiles or no scratch is given.
int skinny_gemm(LhrsGemm& g, float* scratch, void* stream);
inline long long skinny_scratch_elems(const LhrsLlamaWeights* w, long long M) {
    long long rows = M;
    if (2LL * w->ffn > rows) rows = 2LL * w->ffn;
    if (3LL * w->dim > rows) rows = 3LL * w->dim;
    long long a = rows * 3 * w->lora_r, b = 3LL * w->lora_r * (w->ffn > w->dim ? w->ffn : w->dim);
    // the streaming row-reduce kernels (skinny.cu) keep up to 8 row-split partials of [C, n] fp32
    long long widest = (3LL * w->dim > 2LL * w->ffn) ? 3LL * w->dim : 2LL * w->ffn;
    long long c = 8 * widest * 3 * w->lora_r;
    if (b > a) a = b;
    return a > c ? a : c;
}
// LoRA side products on the streaming kernels of skinny.cu (r == 16, dims multiples of 128); LHRS_LORA_STREAM=0 turns them off
bool lora_stream_ok(const LhrsLlamaWeights* w, int in_dim, int out_dim, int nproj);

}  // namespace lhrs
