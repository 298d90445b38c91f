// Single-sequence decode step of LLaMA-2 (HF generate loop reached from text_modal.py:600-612) — HBM-bound.
// One token streams every weight once (13.5 GB at 7B): the kernels are warp-per-row GEMVs with 16-byte loads and the
// small ops fused around them, 5 launches per layer:
//   gemv<QKV>     rmsnorm(x) -> q|k|v                         (reads Wq, Wk, Wv in place, no packed copy)
//   attn_decode   RoPE(q,k) at the device-side position, append K/V to the PAGED cache, softmax(qK^T)V over the pages
//   gemv<O>       x += Wo · o
//   gemv<GATEUP>  act = silu(Wg·rmsnorm(x)) * (Wu·rmsnorm(x))
//   gemv<DOWN>    x += Wd · act
// then gemv<LMHEAD> (final norm fused) + argmax.  Position, context length and the sampled token live in a device-side
// state block, so the host enqueues steps back to back without a per-token synchronisation.
#include "host_common.h"
#include "ptx.cuh"
#include "models_common.h"

namespace lhrs {

enum { GV_QKV = 0, GV_O = 1, GV_GATEUP = 2, GV_DOWN = 3, GV_LMHEAD = 4 };

struct GemvArgs {
    const __nv_bfloat16* w0; const __nv_bfloat16* w1; const __nv_bfloat16* w2;  // weight matrices [rows, K]
    int rows;          // output rows handled per matrix (QKV: dim; GATEUP: ffn; others: all)
    int K;
    const __nv_bfloat16* x;       // input vector [K] (bf16)
    const __nv_bfloat16* norm_w;  // rmsnorm weight (or null: x is used as is)
    float eps;
    __nv_bfloat16* out;           // QKV: [3*rows]; GATEUP: [rows]; O/DOWN: residual stream [rows] updated in place
    float* logits;                // LMHEAD
    float* part_val; int* part_idx;
};

__device__ __forceinline__ float dot8(const uint4& w, const uint4& x) {
    return bf16_lo(w.x) * bf16_lo(x.x) + bf16_hi(w.x) * bf16_hi(x.x) + bf16_lo(w.y) * bf16_lo(x.y) + bf16_hi(w.y) * bf16_hi(x.y) +
           bf16_lo(w.z) * bf16_lo(x.z) + bf16_hi(w.z) * bf16_hi(x.z) + bf16_lo(w.w) * bf16_lo(x.w) + bf16_hi(w.w) * bf16_hi(x.w);
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float warp_red(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// two weight rows against the smem-resident input, 4 x 16-byte loads in flight per row
__device__ __forceinline__ void dot_rows2(const __nv_bfloat16* r0, const __nv_bfloat16* r1, const uint4* xs, int nchunks, int lane,
                                          float& a0, float& a1) {
    const uint4* p0 = reinterpret_cast<const uint4*>(r0);
    const uint4* p1 = reinterpret_cast<const uint4*>(r1);
    a0 = 0.f; a1 = 0.f;
    int c = lane;
    for (; c + 96 < nchunks; c += 128) {
        uint4 w00 = ldg_stream(p0 + c), w01 = ldg_stream(p0 + c + 32), w02 = ldg_stream(p0 + c + 64), w03 = ldg_stream(p0 + c + 96);
        uint4 w10 = ldg_stream(p1 + c), w11 = ldg_stream(p1 + c + 32), w12 = ldg_stream(p1 + c + 64), w13 = ldg_stream(p1 + c + 96);
        a0 += dot8(w00, xs[c]) + dot8(w01, xs[c + 32]) + dot8(w02, xs[c + 64]) + dot8(w03, xs[c + 96]);
        a1 += dot8(w10, xs[c]) + dot8(w11, xs[c + 32]) + dot8(w12, xs[c + 64]) + dot8(w13, xs[c + 96]);
    }
    for (; c < nchunks; c += 32) {
        a0 += dot8(ldg_stream(p0 + c), xs[c]);
        a1 += dot8(ldg_stream(p1 + c), xs[c]);
    }
    a0 = warp_red(a0);
    a1 = warp_red(a1);
}

template <int MODE>
__global__ void __launch_bounds__(256)
gemv_kernel(const GemvArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem);
    __shared__ float red[8];
    __shared__ float bval[8];
    __shared__ int bidx[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunks = a.K / 8;
    // ---- stage the input vector (with the HF RMSNorm fused: w * bf16(x * rstd))
    if (a.norm_w != nullptr) {
        float ss = 0.f;
        for (int i = tid; i < a.K; i += 256) { const float v = __bfloat162float(a.x[i]); ss += v * v; }
        ss = warp_red(ss);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += red[i];
        const float rstd = rsqrtf(tot / a.K + a.eps);
        for (int i = tid; i < a.K; i += 256)
            xs[i] = __float2bfloat16_rn(__bfloat162float(a.norm_w[i]) * bf16_round(__bfloat162float(a.x[i]) * rstd));
    } else {
        for (int i = tid; i < nchunks; i += 256) reinterpret_cast<uint4*>(xs)[i] = reinterpret_cast<const uint4*>(a.x)[i];
    }
    __syncthreads();
    const uint4* xv = reinterpret_cast<const uint4*>(xs);

    float best = -INFINITY;
    int best_i = 0x7fffffff;
    if constexpr (MODE == GV_GATEUP) {
        for (int n = blockIdx.x * 8 + warp; n < a.rows; n += gridDim.x * 8) {
            float g, u;
            dot_rows2(a.w0 + static_cast<long long>(n) * a.K, a.w1 + static_cast<long long>(n) * a.K, xv, nchunks, lane, g, u);
            if (lane == 0) {
                g = bf16_round(g); u = bf16_round(u);
                a.out[n] = __float2bfloat16_rn(bf16_round(g / (1.f + __expf(-g))) * u);
            }
        }
    } else {
        const int total = (MODE == GV_QKV) ? 3 * a.rows : a.rows;
        // each warp takes two rows per trip (independent load streams)
        for (int n = (blockIdx.x * 8 + warp) * 2; n < total; n += gridDim.x * 16) {
            const int n1 = min(n + 1, total - 1);
            const __nv_bfloat16 *r0, *r1;
            if constexpr (MODE == GV_QKV) {
                const int s0 = n / a.rows, s1 = n1 / a.rows;
                const __nv_bfloat16* m0 = s0 == 0 ? a.w0 : (s0 == 1 ? a.w1 : a.w2);
                const __nv_bfloat16* m1 = s1 == 0 ? a.w0 : (s1 == 1 ? a.w1 : a.w2);
                r0 = m0 + static_cast<long long>(n - s0 * a.rows) * a.K;
                r1 = m1 + static_cast<long long>(n1 - s1 * a.rows) * a.K;
            } else {
                r0 = a.w0 + static_cast<long long>(n) * a.K;
                r1 = a.w0 + static_cast<long long>(n1) * a.K;
            }
            float v0, v1;
            dot_rows2(r0, r1, xv, nchunks, lane, v0, v1);
            if (lane == 0) {
                if constexpr (MODE == GV_QKV) {
                    a.out[n] = __float2bfloat16_rn(v0);
                    if (n1 != n) a.out[n1] = __float2bfloat16_rn(v1);
                } else if constexpr (MODE == GV_O || MODE == GV_DOWN) {
                    a.out[n] = __float2bfloat16_rn(bf16_round(v0) + __bfloat162float(a.out[n]));
                    if (n1 != n) a.out[n1] = __float2bfloat16_rn(bf16_round(v1) + __bfloat162float(a.out[n1]));
                } else {  // LMHEAD: HF casts the bf16 logits to fp32
                    v0 = bf16_round(v0); v1 = bf16_round(v1);
                    a.logits[n] = v0;
                    if (v0 > best || (v0 == best && n < best_i)) { best = v0; best_i = n; }
                    if (n1 != n) {
                        a.logits[n1] = v1;
                        if (v1 > best || (v1 == best && n1 < best_i)) { best = v1; best_i = n1; }
                    }
                }
            }
        }
        if constexpr (MODE == GV_LMHEAD) {
            if (lane == 0) { bval[warp] = best; bidx[warp] = best_i; }
            __syncthreads();
            if (tid == 0) {
                for (int i = 1; i < 8; ++i)
                    if (bval[i] > bval[0] || (bval[i] == bval[0] && bidx[i] < bidx[0])) { bval[0] = bval[i]; bidx[0] = bidx[i]; }
                a.part_val[blockIdx.x] = bval[0];
                a.part_idx[blockIdx.x] = bidx[0];
            }
        }
    }
}

// state: [0] token to feed next, [1] ctx_len (positions already in the cache), [2] number of tokens emitted, [3] unused
__global__ void __launch_bounds__(256)
commit_token_kernel(const float* __restrict__ part_val, const int* __restrict__ part_idx, int nparts, int forced_token,
                    int* __restrict__ state, int* __restrict__ tokens_out, int max_tokens, int set_ctx,
                    const __nv_bfloat16* __restrict__ embed, int dim, __nv_bfloat16* __restrict__ xbuf) {
    __shared__ int s_tok;
    if (threadIdx.x == 0) {
        int tok = forced_token;
        if (tok < 0) {
            float bv = -INFINITY;
            int bi = 0x7fffffff;
            for (int i = 0; i < nparts; ++i) {
                const float v = part_val[i];
                const int ix = part_idx[i];
                if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
            }
            tok = bi;
        }
        s_tok = tok;
        const int k = state[2];
        if (k < max_tokens) tokens_out[k] = tok;
        state[0] = tok;
        state[2] = k + 1;
        if (set_ctx >= 0) state[1] = set_ctx;        // after prefill: the prompt's length
        else if (set_ctx == -2) state[1] += 1;      // after a decode step: the fed position is now in the cache
    }
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(embed + static_cast<long long>(s_tok) * dim);
    for (int c = threadIdx.x; c < dim / 8; c += blockDim.x) reinterpret_cast<uint4*>(xbuf)[c] = src[c];
}

// one block per head: RoPE + KV append + attention over the paged cache for the single new query
__global__ void __launch_bounds__(128)
attn_decode_kernel(const __nv_bfloat16* __restrict__ qkv, int dim, KvGeom kv, int layer, int* __restrict__ state,
                   const float* __restrict__ cosT, const float* __restrict__ sinT, __nv_bfloat16* __restrict__ obuf) {
    extern __shared__ float sc[];            // scores [ctx+1]
    __shared__ float qs[128];
    __shared__ float red[4];
    const int h = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pos = state[1];
    const int n = pos + 1;
    const float scale = rsqrtf(128.f);
    const __nv_bfloat16* q = qkv + h * 128;
    const __nv_bfloat16* k = qkv + dim + h * 128;
    const __nv_bfloat16* v = qkv + 2 * dim + h * 128;
    const int page_new = kv.block_table[pos / kv.page_size];
    const int slot_new = pos % kv.page_size;
    if (tid < 64) {
        const float c = cosT[static_cast<long long>(pos) * 64 + tid], s = sinT[static_cast<long long>(pos) * 64 + tid];
        const float q1 = __bfloat162float(q[tid]), q2 = __bfloat162float(q[tid + 64]);
        const float k1 = __bfloat162float(k[tid]), k2 = __bfloat162float(k[tid + 64]);
        // HF apply_rotary_pos_emb in bf16: each product rounded, then the sum rounded
        qs[tid] = bf16_round(bf16_round(q1 * c) - bf16_round(q2 * s));
        qs[tid + 64] = bf16_round(bf16_round(q2 * c) + bf16_round(q1 * s));
        __nv_bfloat16* kd = kv_ptr(kv, layer, 0, page_new, h, slot_new);
        kd[tid] = __float2bfloat16_rn(bf16_round(k1 * c) - bf16_round(k2 * s));
        kd[tid + 64] = __float2bfloat16_rn(bf16_round(k2 * c) + bf16_round(k1 * s));
    } else {
        __nv_bfloat16* vd = kv_ptr(kv, layer, 1, page_new, h, slot_new);
        const int d = tid - 64;
        vd[d] = v[d];
        vd[d + 64] = v[d + 64];
    }
    __syncthreads();
    // scores
    float mx = -INFINITY;
    for (int p = tid; p < n; p += 128) {
        const __nv_bfloat16* kp = kv_ptr(kv, layer, 0, kv.block_table[p / kv.page_size], h, p % kv.page_size);
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const uint4 w = reinterpret_cast<const uint4*>(kp)[c];
            const float* qq = qs + c * 8;
            acc += bf16_lo(w.x) * qq[0] + bf16_hi(w.x) * qq[1] + bf16_lo(w.y) * qq[2] + bf16_hi(w.y) * qq[3] +
                   bf16_lo(w.z) * qq[4] + bf16_hi(w.z) * qq[5] + bf16_lo(w.w) * qq[6] + bf16_hi(w.w) * qq[7];
        }
        acc *= scale;                    // scores stay fp32 (as in the prefill flash kernel)
        sc[p] = acc;
        mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float sum = 0.f;
    for (int p = tid; p < n; p += 128) {
        const float e = __expf(sc[p] - mx);
        sc[p] = e;
        sum += e;
    }
    sum = warp_red(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    const float inv = 1.f / (red[0] + red[1] + red[2] + red[3]);
    // out[d] = sum_p P[p] V[p][d]; thread d owns one output dim (coalesced 256-byte rows)
    float acc0 = 0.f, acc1 = 0.f;
    const int npages = (n + kv.page_size - 1) / kv.page_size;
    for (int pg = 0; pg < npages; ++pg) {
        const __nv_bfloat16* vbase = kv_ptr(kv, layer, 1, kv.block_table[pg], h, 0) + tid;
        const int p0 = pg * kv.page_size;
        const int cnt = min(kv.page_size, n - p0);
        int s = 0;
        for (; s + 1 < cnt; s += 2) {   // independent loads / accumulators hide the L2 latency
            acc0 += sc[p0 + s] * __bfloat162float(vbase[s * 128]);
            acc1 += sc[p0 + s + 1] * __bfloat162float(vbase[(s + 1) * 128]);
        }
        if (s < cnt) acc0 += sc[p0 + s] * __bfloat162float(vbase[s * 128]);
    }
    obuf[h * 128 + tid] = __float2bfloat16_rn((acc0 + acc1) * inv);
}

}  // namespace lhrs

using namespace lhrs;
typedef __nv_bfloat16 bf16;

template <int MODE>
static int launch_gemv(const GemvArgs& a, cudaStream_t st) {
    const int units = (MODE == GV_GATEUP) ? a.rows : ((MODE == GV_QKV ? 3 * a.rows : a.rows) + 1) / 2;
    int grid = (units + 7) / 8;
    const int cap = num_sms() * 8;
    if (grid > cap) grid = cap;
    const size_t smem = (size_t)a.K * 2;
    auto kern = gemv_kernel<MODE>;
    static bool attr = false;
    if (!attr) {
        LHRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr = true;
    }
    LHRS_CHECK_ARG(smem <= 64 * 1024 && a.K % 8 == 0, "gemv: K=%d unsupported", a.K);
    const bool prof = prof_on();
    if (prof) prof_begin(PROF_OTHER, 0.0, 2.0 * (double)a.K * (MODE == GV_QKV ? 3.0 * a.rows : (MODE == GV_GATEUP ? 2.0 * a.rows : (double)a.rows)), st);
    kern<<<grid, 256, smem, st>>>(a);
    if (prof) prof_end(st);
    LHRS_LAUNCH_CHECK("gemv_kernel");
    return grid;
}

static int lm_head_and_commit(const LhrsLlamaWeights* w, const LhrsDecodeBuffers* b, const bf16* x, const bf16* norm_w, int greedy,
                              int set_ctx, cudaStream_t st) {
    GemvArgs a;
    memset(&a, 0, sizeof(a));
    a.w0 = (const bf16*)w->lm_head; a.rows = w->vocab; a.K = w->dim; a.x = x; a.norm_w = norm_w; a.eps = w->eps;
    a.logits = b->logits; a.part_val = b->part_val; a.part_idx = b->part_idx;
    const int grid = launch_gemv<GV_LMHEAD>(a, st);
    if (grid <= 0) return LHRS_ERR_CUDA;
    if (greedy) {
        commit_token_kernel<<<1, 256, 0, st>>>(b->part_val, b->part_idx, grid, -1, b->state, b->tokens_out, b->max_tokens, set_ctx,
                                               (const bf16*)w->embed, w->dim, (bf16*)b->xbuf);
        LHRS_LAUNCH_CHECK("commit_token_kernel");
    }
    return LHRS_OK;
}

extern "C" int lhrs_decode_commit_token(const LhrsLlamaWeights* w, const LhrsDecodeBuffers* b, int32_t token, int32_t set_ctx, void* stream) {
    LHRS_CHECK_ARG(w && b && token >= 0 && token < w->vocab, "lhrs_decode_commit_token: bad token %d", token);
    commit_token_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(b->part_val, b->part_idx, 0, token, b->state, b->tokens_out, b->max_tokens,
                                                            set_ctx, (const bf16*)w->embed, w->dim, (bf16*)b->xbuf);
    LHRS_LAUNCH_CHECK("commit_token_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_llama_first_token(const LhrsLlamaWeights* w, const void* hidden_last, int32_t ctx_len, const LhrsDecodeBuffers* b,
                                      int32_t greedy, void* stream) {
    LHRS_CHECK_ARG(w && hidden_last && b && ctx_len > 0, "lhrs_llama_first_token: bad args");
    LHRS_CUDA(cudaMemsetAsync(b->state, 0, 4 * sizeof(int), (cudaStream_t)stream));
    return lm_head_and_commit(w, b, (const bf16*)hidden_last, nullptr, greedy, ctx_len, (cudaStream_t)stream);
}

extern "C" int lhrs_llama_decode_step(const LhrsLlamaWeights* w, const LhrsKvCache* kv, const LhrsDecodeBuffers* b, int32_t greedy,
                                      int32_t max_ctx, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LHRS_CHECK_ARG(w && kv && b, "lhrs_llama_decode_step: null");
    LHRS_CHECK_ARG(w->dim / w->heads == 128 && kv->head_dim == 128 && kv->heads == w->heads, "lhrs_llama_decode_step: head_dim must be 128");
    LHRS_CHECK_ARG(w->lora_r == 0, "lhrs_llama_decode_step: merge the LoRA adapters first (merge_and_unload; the reference does so for eval, UniBind.py:114)");
    LHRS_CHECK_ARG(max_ctx > 0 && max_ctx <= kv->max_pages * kv->page_size && max_ctx <= w->max_pos, "lhrs_llama_decode_step: max_ctx %d", max_ctx);
    const int D = w->dim, F = w->ffn;
    bf16* x = (bf16*)b->xbuf;
    for (int l = 0; l < w->num_layers; ++l) {
        GemvArgs a;
        memset(&a, 0, sizeof(a));
        a.w0 = (const bf16*)w->q_w[l]; a.w1 = (const bf16*)w->k_w[l]; a.w2 = (const bf16*)w->v_w[l];
        a.rows = D; a.K = D; a.x = x; a.norm_w = (const bf16*)w->ln1_w[l]; a.eps = w->eps; a.out = (bf16*)b->qkv;
        if (launch_gemv<GV_QKV>(a, st) <= 0) return LHRS_ERR_CUDA;
        attn_decode_kernel<<<w->heads, 128, (size_t)max_ctx * sizeof(float), st>>>((const bf16*)b->qkv, D, *kv, l, b->state, w->rope_cos,
                                                                                 w->rope_sin, (bf16*)b->obuf);
        LHRS_LAUNCH_CHECK("attn_decode_kernel");
        memset(&a, 0, sizeof(a));
        a.w0 = (const bf16*)w->o_w[l]; a.rows = D; a.K = D; a.x = (const bf16*)b->obuf; a.out = x;
        if (launch_gemv<GV_O>(a, st) <= 0) return LHRS_ERR_CUDA;
        memset(&a, 0, sizeof(a));
        a.w0 = (const bf16*)w->gate_w[l]; a.w1 = (const bf16*)w->up_w[l]; a.rows = F; a.K = D; a.x = x;
        a.norm_w = (const bf16*)w->ln2_w[l]; a.eps = w->eps; a.out = (bf16*)b->act;
        if (launch_gemv<GV_GATEUP>(a, st) <= 0) return LHRS_ERR_CUDA;
        memset(&a, 0, sizeof(a));
        a.w0 = (const bf16*)w->down_w[l]; a.rows = D; a.K = F; a.x = (const bf16*)b->act; a.out = x;
        if (launch_gemv<GV_DOWN>(a, st) <= 0) return LHRS_ERR_CUDA;
    }
    // the new position is now in the cache for every layer
    {
        int rc = lm_head_and_commit(w, b, x, (const bf16*)w->norm_w, greedy, -2, st);
        if (rc) return rc;
    }
    return LHRS_OK;
}
