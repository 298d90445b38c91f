// Flash attention forward on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), head_dim 128: the LLaMA self-attention of the
// hot path (HF LlamaAttention as driven by lhrs/models/text_modal.py:398-412 and, through it, modeling_llama's eager
// softmax(QK^T/sqrt(d) + mask)V).  Same contract as attn_fwd_kernel in attention.cu (which keeps head_dim 64 and short
// query blocks): strided Q/K/V views, causal with the (Skv - Sq) offset, optional [B, Skv] key mask, fp32 lse.
// Ragged batches (LhrsAttention::seq_off): the tensor maps span the whole batch's row space as one entry, a tile of sequence b
// starts at row seq_off[b] + q0, and the sequence length takes the place of Sq = Skv in every mask and store predicate.
//
// One CTA = one 128-row query tile of one (batch, head); two CTAs are resident per SM so one tile's softmax overlaps
// the other's MMAs.  Per 64-key step:   S = Q K^T   (128x64x128, accumulator in TMEM, double buffered)
//                                        P = exp2(S*c - m*c)  by 128 threads, one query row each, written to smem (SW128)
//                                        O += P V    (128x128x64, accumulator stays in TMEM across the whole row of steps)
// The running max is only raised when it grows by more than 2^8 (then O is rescaled in TMEM), so the O rescale is rare.
//   warp 0: TMA producer (Q once, K/V 2-stage ring)   warp 1: TMEM alloc + MMA issue   warps 2-5: softmax / epilogue
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/lhrs_b200.h"
#include "attention_common.h"
#include "host_common.h"
#include "ptx.cuh"

namespace lhrs {

struct AttnTcArgs {
    __nv_bfloat16* o;
    float* lse;
    const uint8_t* kmask;
    const int* seq_off;                 // ragged batch (LhrsAttention::seq_off): per-sequence row offsets, tensor maps over all rows
    long long o_bs, o_rs, o_hs;
    int B, H, Sq, Skv;
    int q_hfirst, k_hfirst, v_hfirst;   // tensor-map dimension order: {d, head, row, batch} instead of {d, row, head, batch}
    float scale, scale_log2;
};

namespace atc {
constexpr int BQ = 128, BKV = 64, HD = 128;
constexpr int Q_BYTES = BQ * HD * 2;          // 2 k-blocks (64 dims each) of 128 rows x 128 B
constexpr int K_BYTES = BKV * HD * 2;         // 2 k-blocks of 64 rows x 128 B
constexpr int V_BYTES = BKV * HD * 2;         // 2 atom columns (64 dims each) of 64 keys x 128 B
constexpr int P_BYTES = BQ * BKV * 2;         // 128 rows x 128 B
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + Q_BYTES;
constexpr int OFF_V = OFF_K + 2 * K_BYTES;
constexpr int OFF_P = OFF_V + 2 * V_BYTES;
constexpr int OFF_BAR = OFF_P + P_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 128;
constexpr int TMEM_COLS = 256;                // S0 [0,64) | S1 [64,128) | O [128,256)
constexpr float RESCALE_LOG2 = 8.f;
}  // namespace atc

template <bool CAUSAL>
__global__ void __launch_bounds__(192, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnTcArgs p) {
    using namespace atc;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();   // SW128 atoms need 1 KB alignment; no slack is budgeted (2 CTAs/SM)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* q_full = bars;           // 1
    uint64_t* k_full = bars + 1;       // 2   K and V rings are released separately: K_j is dead as soon as S_j has been
    uint64_t* k_empty = bars + 3;      // 2   computed, a whole softmax + PV step before V_j, so its refill has time to land
    uint64_t* s_full = bars + 5;       // 2
    uint64_t* s_empty = bars + 7;      // 2
    uint64_t* p_full = bars + 9;       // 1
    uint64_t* pv_full = bars + 10;     // 1
    uint64_t* v_full = bars + 11;      // 2
    uint64_t* v_empty = bars + 13;     // 2
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 15);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int qt = static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x);   // longest (latest) query tiles first
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = qt * BQ;
    // ragged batch: this sequence's rows start at rb in the shared row space and there are Sq = Skv of them (rows past the end
    // belong to the next sequence: finite values that the length masks below keep out of every sum, and that are never stored)
    int Sq = p.Sq, Skv = p.Skv, rb = 0, bc = b;
    if (p.seq_off != nullptr) { rb = p.seq_off[b]; Sq = Skv = p.seq_off[b + 1] - rb; bc = 0; }
    const int off = Skv - Sq;
    const int kv_end = CAUSAL ? min(Skv, q0 + BQ + off) : Skv;   // keys some row of this tile may attend
    const int n = (kv_end > 0 && q0 < Sq) ? (kv_end + BKV - 1) / BKV : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 1);
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], 128);
        }
        mbar_init(p_full, 128);
        mbar_init(pv_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (n > 0 && elect_one()) {
            mbar_arrive_expect_tx(q_full, Q_BYTES);
            // the tensor maps list (row, head) in increasing-stride order; pick the coordinate order to match
            auto load = [&](void* dst, const CUtensorMap* tm, uint64_t* bar, int d0, int r0, int hfirst) {
                tma_load_4d(dst, tm, bar, d0, hfirst ? h : r0 + rb, hfirst ? r0 + rb : h, bc);
            };
            load(smem + OFF_Q, &tmQ, q_full, 0, q0, p.q_hfirst);
            load(smem + OFF_Q + Q_BYTES / 2, &tmQ, q_full, 64, q0, p.q_hfirst);
            for (int j = 0; j < n; ++j) {
                const int st = j & 1;
                uint8_t* sk = smem + OFF_K + st * K_BYTES;
                uint8_t* sv = smem + OFF_V + st * V_BYTES;
                mbar_wait(&k_empty[st], ((j >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&k_full[st], K_BYTES);
                load(sk, &tmK, &k_full[st], 0, j * BKV, p.k_hfirst);
                load(sk + K_BYTES / 2, &tmK, &k_full[st], 64, j * BKV, p.k_hfirst);
                mbar_wait(&v_empty[st], ((j >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&v_full[st], V_BYTES);
                load(sv, &tmV, &v_full[st], 0, j * BKV, p.v_hfirst);
                load(sv + V_BYTES / 2, &tmV, &v_full[st], 64, j * BKV, p.v_hfirst);
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (n > 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0u, 0u);    // S = Q K^T : both operands K-major
            constexpr uint32_t idesc_o = make_idesc_bf16(BQ, HD, 0u, 1u);     // O = P V   : V is MN-major (dims contiguous)
            constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (UMMA_LAYOUT_SW128 << 29);
            const uint32_t q_lo = (smem_u32(smem + OFF_Q) >> 4) & 0x3FFFu;
            const uint32_t k_lo = (smem_u32(smem + OFF_K) >> 4) & 0x3FFFu;
            const uint32_t v_lo = ((smem_u32(smem + OFF_V) >> 4) & 0x3FFFu) | ((8192u >> 4) << 16);   // LBO: 64-dim atoms 8 KB apart
            const uint32_t p_lo = (smem_u32(smem + OFF_P) >> 4) & 0x3FFFu;
            const bool issuer = elect_one();
            mbar_wait(q_full, 0);
            auto issue_s = [&](int j) {
                const int st = j & 1;
                mbar_wait(&k_full[st], (j >> 1) & 1);
                mbar_wait(&s_empty[st], ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                if (issuer) {
#pragma unroll
                    for (int kk = 0; kk < HD / 16; ++kk) {
                        const uint32_t a = q_lo + (kk >> 2) * (Q_BYTES / 2 >> 4) + (kk & 3) * 2;
                        const uint32_t bb = k_lo + st * (K_BYTES >> 4) + (kk >> 2) * (K_BYTES / 2 >> 4) + (kk & 3) * 2;
                        umma_bf16_w(tmem_base + st * BKV, a, bb, desc_hi, idesc_s, kk ? 1u : 0u);
                    }
                    umma_commit(&k_empty[st]);
                    umma_commit(&s_full[st]);
                }
                __syncwarp();
            };
            issue_s(0);
            for (int j = 0; j < n; ++j) {
                if (j + 1 < n) issue_s(j + 1);
                const int st = j & 1;
                mbar_wait(&v_full[st], (j >> 1) & 1);
                mbar_wait(p_full, j & 1);
                tc_fence_after();
                if (issuer) {
#pragma unroll
                    for (int kk = 0; kk < BKV / 16; ++kk)
                        umma_bf16_w(tmem_base + 2 * BKV, p_lo + kk * 2, v_lo + st * (V_BYTES >> 4) + kk * (2048 >> 4), desc_hi,
                                    idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(&v_empty[st]);
                    umma_commit(pv_full);
                }
                __syncwarp();
            }
        }
    } else {
        // ===================================================================== softmax + epilogue (one query row per thread)
        const int quarter = warp & 3;                     // TMEM lane quarter this warp may touch
        const int row = quarter * 32 + lane;
        const int qi = q0 + row;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const int row_lim = CAUSAL ? qi + off : 0x7fffffff;   // last key this row may attend
        const float c = p.scale_log2;
        float m = -INFINITY, l = 0.f;
        uint8_t* prow = smem + OFF_P + row * 128;
        for (int j = 0; j < n; ++j) {
            const int k0 = j * BKV;
            const int sb = j & 1;
            mbar_wait(&s_full[sb], (j >> 1) & 1);
            tc_fence_after();
            uint32_t r[64];
            tmem_ld_32x32(t_lane + sb * BKV, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
            tmem_ld_32x32(t_lane + sb * BKV + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&s_empty[sb]);

            // ---- allowed-key bitmask of this step (key mask and sequence end are row-independent; causal is per row)
            uint32_t w0 = 0xffffffffu, w1 = 0xffffffffu;
            if (p.kmask != nullptr) {
                const uint8_t* km = p.kmask + static_cast<long long>(b) * Skv + k0;
                const uint32_t b0 = (k0 + lane < Skv) ? km[lane] : 0u;
                const uint32_t b1 = (k0 + 32 + lane < Skv) ? km[32 + lane] : 0u;
                w0 = __ballot_sync(0xffffffffu, b0 != 0u);
                w1 = __ballot_sync(0xffffffffu, b1 != 0u);
            } else if (k0 + BKV > Skv) {
                const int v = Skv - k0;   // 1..63 valid keys
                w0 = v >= 32 ? 0xffffffffu : ((1u << v) - 1u);
                w1 = v > 32 ? ((1u << (v - 32)) - 1u) : 0u;
            }
            if (CAUSAL) {
                const int nb = row_lim - k0 + 1;   // keys [k0, k0 + nb) allowed
                if (nb < 64) {
                    w0 &= nb >= 32 ? 0xffffffffu : (nb <= 0 ? 0u : ((1u << nb) - 1u));
                    w1 &= nb <= 32 ? 0u : ((1u << (nb - 32)) - 1u);
                }
            }
            float s[64];
#pragma unroll
            for (int i = 0; i < 64; ++i) s[i] = __uint_as_float(r[i]);
            if ((w0 & w1) != 0xffffffffu) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (!((w0 >> i) & 1u)) s[i] = -INFINITY;
                    if (!((w1 >> i) & 1u)) s[32 + i] = -INFINITY;
                }
            }
            float mx4[4] = {s[0], s[1], s[2], s[3]};   // four independent chains: a single 63-deep fmax chain is pure latency
#pragma unroll
            for (int i = 4; i < 64; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], s[i]);
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
            // ---- lazy running max: raise it only when it would grow by more than 2^RESCALE_LOG2
            const float m_new = fmaxf(m, mx);
            float alpha = 1.f;
            const bool grow = (m_new - m) * c > RESCALE_LOG2;   // (-inf) - (-inf) = NaN -> false
            if (grow) {
                alpha = (m == -INFINITY) ? 0.f : exp2f((m - m_new) * c);
                m = m_new;
                l *= alpha;
            }
            const float neg = (m == -INFINITY) ? 0.f : -m * c;
            float sum4[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t pk[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float e0 = ex2_approx(fmaf(s[2 * i], c, neg));
                const float e1 = ex2_approx(fmaf(s[2 * i + 1], c, neg));
                sum4[i & 3] += e0 + e1;
                pk[i] = pack_bf16(e0, e1);
            }
            l += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);

            if (j > 0) {
                // P smem and the O accumulator are free once the previous step's PV MMAs have completed
                mbar_wait(pv_full, (j - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
                    for (int ch = 0; ch < HD / 32; ++ch) {
                        uint32_t o[32];
                        tmem_ld_32x32(t_lane + 2 * BKV + ch * 32, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st_32x32(t_lane + 2 * BKV + ch * 32, o);
                    }
                    tmem_st_wait();
                }
            }
            // ---- P row -> smem, 128B-swizzled K-major (16-byte chunk ch of row r lives at chunk ch ^ (r & 7))
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                uint4 v = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
                *reinterpret_cast<uint4*>(prow + ((ch ^ (row & 7)) << 4)) = v;
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_full);
        }

        // ---- finalize: O /= l, write bf16; lse = m*scale + ln(l)
        const float inv = l > 0.f ? 1.f / l : 0.f;
        if (n > 0) {
            mbar_wait(pv_full, (n - 1) & 1);
            tc_fence_after();
        }
        __nv_bfloat16* og = p.o + bc * p.o_bs + h * p.o_hs + static_cast<long long>(rb + qi) * p.o_rs;
#pragma unroll 1
        for (int ch = 0; ch < HD / 32; ++ch) {
            uint32_t o[32];
            if (n > 0) {
                tmem_ld_32x32(t_lane + 2 * BKV + ch * 32, o);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = 0u;
            }
            if (qi < Sq) {
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4) {
                    uint4 v;
                    v.x = pack_bf16(__uint_as_float(o[v4 * 8 + 0]) * inv, __uint_as_float(o[v4 * 8 + 1]) * inv);
                    v.y = pack_bf16(__uint_as_float(o[v4 * 8 + 2]) * inv, __uint_as_float(o[v4 * 8 + 3]) * inv);
                    v.z = pack_bf16(__uint_as_float(o[v4 * 8 + 4]) * inv, __uint_as_float(o[v4 * 8 + 5]) * inv);
                    v.w = pack_bf16(__uint_as_float(o[v4 * 8 + 6]) * inv, __uint_as_float(o[v4 * 8 + 7]) * inv);
                    *reinterpret_cast<uint4*>(og + ch * 32 + v4 * 8) = v;
                }
            }
        }
        if (p.lse != nullptr && qi < Sq)
            p.lse[(static_cast<long long>(b) * p.H + h) * p.Sq + qi] = (l > 0.f) ? m * p.scale + logf(l) : -INFINITY;
        tc_fence_before();
    }

    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, atc::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// 4-D bf16 view {head_dim, rows | heads (smaller stride first), batch} with element strides (1, rs, hs, bs); the box is 64 dims x
// box_rows rows of one head, 128B swizzle.  Rows past S are zero-filled by TMA and a box never crosses into another head or
// batch entry.  *hfirst tells the kernel which coordinate order the map expects.
int make_tmap_bshd(CUtensorMap* out, int* hfirst, const void* ptr, int hd, int S, int H, int B, long long rs, long long hs,
                          long long bs, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return LHRS_ERR_CUDA;
    }
    if (H == 1) hs = rs * S;            // a dimension of extent 1 never uses its stride; keep the list increasing
    if (B == 1) bs = (rs > hs ? rs * S : hs * H);
    *hfirst = hs < rs ? 1 : 0;
    cuuint64_t dims[4], strides[3];
    cuuint32_t box[4];
    dims[0] = (cuuint64_t)hd; box[0] = 64;
    const int ri = *hfirst ? 2 : 1, hi = *hfirst ? 1 : 2;
    dims[ri] = (cuuint64_t)S; strides[ri - 1] = (cuuint64_t)rs * 2; box[ri] = (cuuint32_t)box_rows;
    dims[hi] = (cuuint64_t)H; strides[hi - 1] = (cuuint64_t)hs * 2; box[hi] = 1;
    dims[3] = (cuuint64_t)B; strides[2] = (cuuint64_t)bs * 2; box[3] = 1;
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (attention) failed (%d): ptr=%p S=%d H=%d B=%d rs=%lld hs=%lld bs=%lld", (int)r, ptr, S, H,
                  B, rs, hs, bs);
        return LHRS_ERR_CUDA;
    }
    return LHRS_OK;
}

template <bool CAUSAL>
static int launch_tc(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnTcArgs& a, cudaStream_t stream) {
    auto kern = attn_fwd_tc_kernel<CAUSAL>;
    static bool attr_set = false;
    if (!attr_set) {
        LHRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid((a.Sq + atc::BQ - 1) / atc::BQ, a.H, a.B);
    const bool prof = prof_on();
    if (prof) {
        const double pairs = CAUSAL ? 0.5 * a.Sq * (double)a.Skv : (double)a.Sq * a.Skv;   // causal counted at half
        prof_begin(PROF_ATTN, 4.0 * a.B * a.H * pairs * atc::HD, 2.0 * a.B * a.H * atc::HD * (2.0 * a.Sq + 2.0 * a.Skv), stream);
    }
    kern<<<grid, 192, atc::SMEM_BYTES, stream>>>(tq, tk, tv, a);
    if (prof) prof_end(stream);
    LHRS_LAUNCH_CHECK("attn_fwd_tc_kernel");
    return LHRS_OK;
}

// Called by lhrs_attention_fwd for head_dim 128 problems with at least one full query tile.  Arguments are already validated.
int attention_fwd_tc(const LhrsAttention* d, cudaStream_t stream) {
    CUtensorMap tq, tk, tv;
    int rc;
    AttnTcArgs a;
    const bool ragged = d->seq_off != nullptr;      // one row space for the whole batch: maps with a single batch entry of total_rows rows
    const int mB = ragged ? 1 : d->B, mSq = ragged ? (int)d->total_rows : d->Sq, mSkv = ragged ? (int)d->total_rows : d->Skv;
    if ((rc = make_tmap_bshd(&tq, &a.q_hfirst, d->q, 128, mSq, d->H, mB, d->q_rs, d->q_hs, d->q_bs, atc::BQ))) return rc;
    if ((rc = make_tmap_bshd(&tk, &a.k_hfirst, d->k, 128, mSkv, d->H, mB, d->k_rs, d->k_hs, d->k_bs, atc::BKV))) return rc;
    if ((rc = make_tmap_bshd(&tv, &a.v_hfirst, d->v, 128, mSkv, d->H, mB, d->v_rs, d->v_hs, d->v_bs, atc::BKV))) return rc;
    a.seq_off = d->seq_off;
    a.o = reinterpret_cast<__nv_bfloat16*>(d->o);
    a.lse = d->lse;
    a.kmask = d->key_mask;
    a.o_bs = d->o_bs; a.o_rs = d->o_rs; a.o_hs = d->o_hs;
    a.B = d->B; a.H = d->H; a.Sq = d->Sq; a.Skv = d->Skv;
    a.scale = d->scale;
    a.scale_log2 = d->scale * 1.4426950408889634f;
    return d->causal ? launch_tc<true>(tq, tk, tv, a, stream) : launch_tc<false>(tq, tk, tv, a, stream);
}

}  // namespace lhrs
