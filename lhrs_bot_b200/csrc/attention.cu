// Fused softmax(Q K^T * scale + mask) V ("flash" style, online softmax, nothing S x S ever reaches HBM).
//   - causal + key-padding mask, head_dim 128 : LLaMA self-attention      (HF LlamaAttention, called from lhrs/models/text_modal.py:281-290)
//   - non-causal, head_dim 64                  : CLIP ViT self-attention   (HF CLIPAttention, lhrs/models/rgb_vision_modal.py:168-172)
//                                                and the AttnPooler cross-attention (nn.MultiheadAttention, lhrs/models/common_arch.py:302-313)
// Attention is ~1% of the path's FLOPs at S=512, so this first version uses warp-level mma.sync tiles
// (cp.async double-buffered K/V, ldmatrix fragments); the dense projections around it are tcgen05.
#include "attention_common.h"
#include "host_common.h"
#include "ptx.cuh"

namespace lhrs {

struct AttnArgs {
    const __nv_bfloat16* q;
    const __nv_bfloat16* k;
    const __nv_bfloat16* v;
    __nv_bfloat16* o;
    float* lse;              // [B, H, Sq] natural-log-sum-exp of the scaled scores, or null
    const uint8_t* kmask;    // [B, Skv] 1 = attend, 0 = masked; or null
    long long q_bs, q_rs, q_hs;  // batch / row / head strides in elements
    long long k_bs, k_rs, k_hs;
    long long v_bs, v_rs, v_hs;
    long long o_bs, o_rs, o_hs;
    int B, H, Sq, Skv;
    float scale_log2;  // softmax scale * log2(e)
    float scale;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// smem tile [rows][HD] bf16, 16-byte chunks XOR-swizzled within each 128-byte group
template <int HD>
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
    return static_cast<uint32_t>(row * HD * 2 + (((chunk & ~7) | ((chunk & 7) ^ (row & 7))) << 4));
}

template <int HD, int ROWS, int THREADS>
__device__ __forceinline__ void load_tile(uint32_t smem_base, const __nv_bfloat16* gptr, long long row_stride,
                                          int row0, int nrows_total, int tid) {
    constexpr int CHUNKS = HD / 8;
    constexpr int TOTAL = ROWS * CHUNKS;
#pragma unroll
    for (int i = 0; i < TOTAL / THREADS; ++i) {
        const int idx = tid + i * THREADS;
        const int r = idx / CHUNKS;
        const int c = idx % CHUNKS;
        const int gr = row0 + r;
        const bool ok = gr < nrows_total;
        const __nv_bfloat16* src = gptr + static_cast<long long>(ok ? gr : 0) * row_stride + c * 8;
        cp_async16(smem_base + tile_off<HD>(r, c), src, ok);
    }
}

// Up to three problems that share (B, H) in ONE launch: blockIdx.z = group * B + batch.  The AttnPooler's three query groups
// (64/48/32 queries against 320/304/288 keys, common_arch.py:159-166) run as one grid instead of three.
struct AttnArgsG {
    AttnArgs a[3];
    int n;
};

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const AttnArgsG G) {
    constexpr int BQ = 64, BKV = 64, THREADS = 128;
    const int gi = static_cast<int>(blockIdx.z) / G.a[0].B;
    const AttnArgs p = (gi == 0) ? G.a[0] : (gi == 1 ? G.a[1] : G.a[2]);
    if (static_cast<int>(blockIdx.x) * BQ >= p.Sq) return;     // shorter groups own fewer query blocks
    constexpr int KSTEPS = HD / 16;   // k-steps of QK^T
    constexpr int DTILES = HD / 8;    // n-tiles of the output
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sK = sQ + BQ * HD * 2;
    const uint32_t sV = sK + 2 * BKV * HD * 2;
    __shared__ uint8_t sMask[2][BKV];   // key-padding mask of the current / next key tile

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int qb = blockIdx.x, h = blockIdx.y, b = static_cast<int>(blockIdx.z) - gi * p.B;
    const int q0 = qb * BQ;

    const __nv_bfloat16* qg = p.q + b * p.q_bs + h * p.q_hs;
    const __nv_bfloat16* kg = p.k + b * p.k_bs + h * p.k_hs;
    const __nv_bfloat16* vg = p.v + b * p.v_bs + h * p.v_hs;
    const uint8_t* km = p.kmask ? p.kmask + static_cast<long long>(b) * p.Skv : nullptr;

    const int causal_shift = p.Skv - p.Sq;  // query i sees keys <= i + shift
    int kv_end = p.Skv;
    if (CAUSAL) kv_end = min(p.Skv, q0 + BQ + causal_shift);
    const int nblk = (kv_end + BKV - 1) / BKV;

    load_tile<HD, BQ, THREADS>(sQ, qg, p.q_rs, q0, p.Sq, tid);
    load_tile<HD, BKV, THREADS>(sK, kg, p.k_rs, 0, p.Skv, tid);
    load_tile<HD, BKV, THREADS>(sV, vg, p.v_rs, 0, p.Skv, tid);
    cp_async_commit();
    if (km != nullptr && tid < BKV) sMask[0][tid] = (tid < p.Skv) ? km[tid] : 0;

    uint32_t qf[KSTEPS][4];
    float o[DTILES][4];
#pragma unroll
    for (int i = 0; i < DTILES; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};

    const int g = lane >> 2, t4 = lane & 3;
    const int qrow0 = q0 + warp * 16 + g;  // rows qrow0 and qrow0 + 8

    for (int j = 0; j < nblk; ++j) {
        const int buf = j & 1;
        if (j + 1 < nblk) {
            load_tile<HD, BKV, THREADS>(sK + (buf ^ 1) * BKV * HD * 2, kg, p.k_rs, (j + 1) * BKV, p.Skv, tid);
            load_tile<HD, BKV, THREADS>(sV + (buf ^ 1) * BKV * HD * 2, vg, p.v_rs, (j + 1) * BKV, p.Skv, tid);
            cp_async_commit();
            if (km != nullptr && tid < BKV) {
                const int key = (j + 1) * BKV + tid;
                sMask[buf ^ 1][tid] = (key < p.Skv) ? km[key] : 0;
            }
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        if (j == 0) {
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const int r = warp * 16 + (lane & 15);
                const int c = ks * 2 + (lane >> 4);
                ldsm_x4(sQ + tile_off<HD>(r, c), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
            }
        }

        const uint32_t kb = sK + buf * BKV * HD * 2;
        const uint32_t vb = sV + buf * BKV * HD * 2;

        // ---- S = Q K^T  (16 x 64 per warp)
        float s[BKV / 8][4];
#pragma unroll
        for (int i = 0; i < BKV / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
            for (int np = 0; np < BKV / 16; ++np) {
                const int mat = lane >> 3;
                const int key = np * 16 + (mat >> 1) * 8 + (lane & 7);
                const int c = ks * 2 + (mat & 1);
                uint32_t b0, b1, b2, b3;
                ldsm_x4(kb + tile_off<HD>(key, c), b0, b1, b2, b3);
                mma_bf16_16816(s[np * 2], qf[ks], b0, b1);
                mma_bf16_16816(s[np * 2 + 1], qf[ks], b2, b3);
            }
        }

        // ---- mask + online softmax (rows g and g+8 of this warp's 16)
        const int kbase = j * BKV;
        float mx[2] = {-INFINITY, -INFINITY};
        // masking is only evaluated on tiles that need it: the diagonal tile (causal), the ragged last tile, a padding mask
        const bool tile_masked = (km != nullptr) || (kbase + BKV > p.Skv) || (CAUSAL && kbase + BKV - 1 > q0 + warp * 16 + causal_shift);
        if (tile_masked) {
#pragma unroll
            for (int i = 0; i < BKV / 8; ++i) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int kl = i * 8 + t4 * 2 + (e & 1);
                    const int key = kbase + kl;
                    const int qr = qrow0 + (e >> 1) * 8;
                    bool ok = key < p.Skv;
                    if (CAUSAL) ok = ok && (key <= qr + causal_shift);
                    if (km != nullptr) ok = ok && (sMask[buf][kl] != 0);
                    if (!ok) s[i][e] = -INFINITY;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < BKV / 8; ++i) {
            mx[0] = fmaxf(mx[0], fmaxf(s[i][0], s[i][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[i][2], s[i][3]));
        }
        float scale_o[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_run[r], mx[r]);
            // fully-masked-so-far rows keep m = -inf; use 0 as the subtraction point to avoid inf - inf
            const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
            scale_o[r] = exp2f((m_run[r] - m_use) * p.scale_log2);
            if (m_run[r] == -INFINITY) scale_o[r] = 0.f;
            m_run[r] = m_new;
            mx[r] = m_use;
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pf[BKV / 16][4];
#pragma unroll
        for (int i = 0; i < BKV / 8; ++i) {
            float e0 = exp2f((s[i][0] - mx[0]) * p.scale_log2);
            float e1 = exp2f((s[i][1] - mx[0]) * p.scale_log2);
            float e2 = exp2f((s[i][2] - mx[1]) * p.scale_log2);
            float e3 = exp2f((s[i][3] - mx[1]) * p.scale_log2);
            rs[0] += e0 + e1;
            rs[1] += e2 + e3;
            // C-fragment of S -> A-fragment of P for the PV product
            pf[i >> 1][(i & 1) * 2 + 0] = pack_bf16(e0, e1);
            pf[i >> 1][(i & 1) * 2 + 1] = pack_bf16(e2, e3);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * scale_o[r] + rs[r];
#pragma unroll
        for (int i = 0; i < DTILES; ++i) {
            o[i][0] *= scale_o[0]; o[i][1] *= scale_o[0];
            o[i][2] *= scale_o[1]; o[i][3] *= scale_o[1];
        }

        // ---- O += P V
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
#pragma unroll
            for (int dp = 0; dp < DTILES / 2; ++dp) {
                const int mat = lane >> 3;
                const int key = kk * 16 + (mat & 1) * 8 + (lane & 7);
                const int c = dp * 2 + (mat >> 1);
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(vb + tile_off<HD>(key, c), b0, b1, b2, b3);
                mma_bf16_16816(o[dp * 2], pf[kk], b0, b1);
                mma_bf16_16816(o[dp * 2 + 1], pf[kk], b2, b3);
            }
        }
        __syncthreads();
    }

    // ---- finalize: O /= l, write bf16; lse = m*scale + ln(l)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
    const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
    __nv_bfloat16* og = p.o + b * p.o_bs + h * p.o_hs;
#pragma unroll
    for (int i = 0; i < DTILES; ++i) {
        const int d = i * 8 + t4 * 2;
        if (qrow0 < p.Sq)
            *reinterpret_cast<uint32_t*>(og + static_cast<long long>(qrow0) * p.o_rs + d) = pack_bf16(o[i][0] * inv0, o[i][1] * inv0);
        if (qrow0 + 8 < p.Sq)
            *reinterpret_cast<uint32_t*>(og + static_cast<long long>(qrow0 + 8) * p.o_rs + d) = pack_bf16(o[i][2] * inv1, o[i][3] * inv1);
    }
    if (p.lse != nullptr && t4 == 0) {
        float* lg = p.lse + (static_cast<long long>(b) * p.H + h) * p.Sq;
        if (qrow0 < p.Sq) lg[qrow0] = (l_run[0] > 0.f) ? m_run[0] * p.scale + logf(l_run[0]) : -INFINITY;
        if (qrow0 + 8 < p.Sq) lg[qrow0 + 8] = (l_run[1] > 0.f) ? m_run[1] * p.scale + logf(l_run[1]) : -INFINITY;
    }
}

template <int HD, bool CAUSAL>
static int launch_attn(const AttnArgsG& G, cudaStream_t stream) {
    const AttnArgs& a = G.a[0];
    constexpr int SMEM = (64 + 4 * 64) * HD * 2;
    auto kern = attn_fwd_kernel<HD, CAUSAL>;
    static bool attr_set = false;
    if (!attr_set) {
        LHRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set = true;
    }
    int max_sq = 0;
    double pairs = 0, rows = 0;
    for (int i = 0; i < G.n; ++i) {
        max_sq = G.a[i].Sq > max_sq ? G.a[i].Sq : max_sq;
        pairs += CAUSAL ? 0.5 * G.a[i].Sq * (double)G.a[i].Skv : (double)G.a[i].Sq * G.a[i].Skv;  // causal counted at half
        rows += 2.0 * G.a[i].Sq + 2.0 * G.a[i].Skv;
    }
    dim3 grid((max_sq + 63) / 64, a.H, a.B * G.n);
    const bool prof = prof_on();
    if (prof) prof_begin(PROF_ATTN, 4.0 * a.B * a.H * pairs * HD, 2.0 * a.B * a.H * HD * rows, stream);
    kern<<<grid, 128, SMEM, stream>>>(G);
    if (prof) prof_end(stream);
    LHRS_LAUNCH_CHECK("attn_fwd_kernel");
    return LHRS_OK;
}

bool use_tc_attention() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("LHRS_ATTN_TC"); v = e ? atoi(e) : 1; }
    return v != 0;
}

}  // namespace lhrs

using namespace lhrs;

extern "C" int lhrs_attention_fwd(const LhrsAttention* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    LHRS_CHECK_ARG(d != nullptr && d->q && d->k && d->v && d->o, "lhrs_attention_fwd: null operand");
    LHRS_CHECK_ARG(d->head_dim == 64 || d->head_dim == 128, "lhrs_attention_fwd: head_dim %d (only 64, 128)", d->head_dim);
    LHRS_CHECK_ARG(d->B > 0 && d->H > 0 && d->Sq > 0 && d->Skv > 0, "lhrs_attention_fwd: empty problem");
    const long long strides[] = {d->q_bs, d->q_rs, d->q_hs, d->k_bs, d->k_rs, d->k_hs, d->v_bs, d->v_rs, d->v_hs, d->o_rs, d->o_hs, d->o_bs};
    for (long long s : strides) LHRS_CHECK_ARG((s % 8) == 0, "lhrs_attention_fwd: strides must be multiples of 8 elements");
    if (d->seq_off != nullptr)   // ragged batch: tcgen05 kernel only
        LHRS_CHECK_ARG(d->head_dim == 128 && d->causal && d->Sq == d->Skv && d->Sq >= 128 && d->key_mask == nullptr &&
                           d->total_rows > 0 && d->total_rows < (1ll << 31) && use_tc_attention(),
                       "lhrs_attention_fwd: seq_off needs head_dim 128, causal, Sq == Skv >= 128, no key_mask, total_rows (and LHRS_ATTN_TC on)");
    // head_dim 128 with at least one full 128-row query tile runs on the tcgen05 kernel (attention_tc.cu)
    if (d->head_dim == 128 && d->Sq >= 128 && use_tc_attention()) return attention_fwd_tc(d, stream);
    AttnArgsG G;
    G.n = 1;
    AttnArgs& a = G.a[0];
    a.q = reinterpret_cast<const __nv_bfloat16*>(d->q);
    a.k = reinterpret_cast<const __nv_bfloat16*>(d->k);
    a.v = reinterpret_cast<const __nv_bfloat16*>(d->v);
    a.o = reinterpret_cast<__nv_bfloat16*>(d->o);
    a.lse = d->lse;
    a.kmask = d->key_mask;
    a.q_bs = d->q_bs; a.q_rs = d->q_rs; a.q_hs = d->q_hs;
    a.k_bs = d->k_bs; a.k_rs = d->k_rs; a.k_hs = d->k_hs;
    a.v_bs = d->v_bs; a.v_rs = d->v_rs; a.v_hs = d->v_hs;
    a.o_bs = d->o_bs; a.o_rs = d->o_rs; a.o_hs = d->o_hs;
    a.B = d->B; a.H = d->H; a.Sq = d->Sq; a.Skv = d->Skv;
    a.scale = d->scale;
    a.scale_log2 = d->scale * 1.4426950408889634f;
    G.a[1] = a; G.a[2] = a;
    if (d->head_dim == 128) return d->causal ? launch_attn<128, true>(G, stream) : launch_attn<128, false>(G, stream);
    return d->causal ? launch_attn<64, true>(G, stream) : launch_attn<64, false>(G, stream);
}

// n (<= 3) forward problems with the same B, H, head_dim 64 and causal flag in one launch (the AttnPooler's query groups)
extern "C" int lhrs_attention_fwd_grouped(const LhrsAttention* d, int32_t n, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    LHRS_CHECK_ARG(d != nullptr && n >= 1 && n <= 3, "lhrs_attention_fwd_grouped: 1..3 problems");
    bool same = true;
    for (int i = 0; i < n; ++i) {
        LHRS_CHECK_ARG(d[i].q && d[i].k && d[i].v && d[i].o && d[i].B > 0 && d[i].H > 0 && d[i].Sq > 0 && d[i].Skv > 0, "lhrs_attention_fwd_grouped: null/empty problem %d", i);
        same = same && d[i].B == d[0].B && d[i].H == d[0].H && d[i].head_dim == 64 && d[i].causal == d[0].causal && d[i].seq_off == nullptr;
        const long long strides[] = {d[i].q_bs, d[i].q_rs, d[i].q_hs, d[i].k_bs, d[i].k_rs, d[i].k_hs, d[i].v_bs, d[i].v_rs, d[i].v_hs, d[i].o_rs, d[i].o_hs, d[i].o_bs};
        for (long long s : strides) LHRS_CHECK_ARG((s % 8) == 0, "lhrs_attention_fwd_grouped: strides must be multiples of 8 elements");
    }
    if (!same || n == 1) {       // not groupable: one launch each
        for (int i = 0; i < n; ++i) { const int rc = lhrs_attention_fwd(&d[i], stream_); if (rc) return rc; }
        return LHRS_OK;
    }
    AttnArgsG G;
    G.n = n;
    for (int i = 0; i < 3; ++i) {
        const LhrsAttention& s = d[i < n ? i : 0];
        AttnArgs& a = G.a[i];
        a.q = reinterpret_cast<const __nv_bfloat16*>(s.q); a.k = reinterpret_cast<const __nv_bfloat16*>(s.k);
        a.v = reinterpret_cast<const __nv_bfloat16*>(s.v); a.o = reinterpret_cast<__nv_bfloat16*>(s.o);
        a.lse = s.lse; a.kmask = s.key_mask;
        a.q_bs = s.q_bs; a.q_rs = s.q_rs; a.q_hs = s.q_hs; a.k_bs = s.k_bs; a.k_rs = s.k_rs; a.k_hs = s.k_hs;
        a.v_bs = s.v_bs; a.v_rs = s.v_rs; a.v_hs = s.v_hs; a.o_bs = s.o_bs; a.o_rs = s.o_rs; a.o_hs = s.o_hs;
        a.B = s.B; a.H = s.H; a.Sq = s.Sq; a.Skv = s.Skv; a.scale = s.scale; a.scale_log2 = s.scale * 1.4426950408889634f;
    }
    return d[0].causal ? launch_attn<64, true>(G, stream) : launch_attn<64, false>(G, stream);
}
