// Flash attention backward on tcgen05 (head_dim 128), the recompute-based gradient of attention_tc.cu's forward
// (what autograd derives for HF LlamaAttention's softmax(QK^T/sqrt(d) + mask)V; lhrs/models/text_modal.py:398-412 drives it).
// Two kernels, no atomics, deterministic:
//
//   dQ kernel     CTA = 128 query rows (TMEM lanes = queries), loops over 64-key steps:
//                   S = Q K^T, dP = dO V^T  (TMEM)  ->  dS = P o (dP - delta)  (bf16, smem SW128)  ->  dQ += dS K   (TMEM)
//   dK/dV kernel  CTA = 128 keys (TMEM lanes = keys), loops over 64-query steps:
//                   S^T = K Q^T, dP^T = V dO^T  ->  P^T, dS^T (smem)  ->  dV += P^T dO,  dK += dS^T Q   (TMEM)
//
// The backward needs no row reductions, so each 64-column step is split between two warps per TMEM lane quarter (8 compute
// warps).  The S/dP accumulators are double buffered so the next step's MMAs run under this step's exponentials.  K and Q tiles
// are read twice from the same SW128 smem bytes: K-major for the score products, MN-major for the gradient products.
// scale and the optional inverse RoPE (dQ, dK) are applied once, in the epilogue.
//   warp 0: TMA producer   warp 1: TMEM alloc + MMA issue   warps 2-9: compute / epilogue
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/lhrs_b200.h"
#include "attention_common.h"
#include "host_common.h"
#include "ptx.cuh"
#include "attention_bwd_tc.cuh"

namespace lhrs {

// ======================================================================================================== dQ
namespace dqk {
using namespace abt;
constexpr int OFF_Q = 0;
constexpr int OFF_DO = OFF_Q + T128;
constexpr int NST = 4;                       // K/V ring depth: a stage is only free once dQ(j) has read K_j, and the refill
                                             // takes a TMA round trip, so two stages would expose that latency every step
constexpr int OFF_K = OFF_DO + T128;         // NST stages
constexpr int OFF_V = OFF_K + NST * T64;     // NST stages
constexpr int OFF_DS = OFF_V + NST * T64;
constexpr int OFF_BAR = OFF_DS + PS;
constexpr int SMEM_BYTES = OFF_BAR + 256;
constexpr int TM_S = 0, TM_DP = 128, TM_DQ = 256;
}  // namespace dqk

template <bool CAUSAL>
__global__ void __launch_bounds__(320, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                      const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const AttnBwdTcArgs pa) {
    using namespace dqk;
    const AttnBwdArgs& p = pa.a;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* qdo_full = bars;        // 1
    uint64_t* sdp_full = bars + 1;    // 2
    uint64_t* sdp_empty = bars + 3;   // 2
    uint64_t* ds_full = bars + 5;     // 1
    uint64_t* dq_done = bars + 6;     // 1
    uint64_t* kv_full = bars + 8;     // NST
    uint64_t* kv_empty = bars + 8 + NST;   // NST
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8 + 2 * NST);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int qt = static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x);
    const int h = blockIdx.y, b = blockIdx.z;
    const int q0 = qt * 128;
    const int off = p.Skv - p.Sq;
    const int kv_end = CAUSAL ? min(p.Skv, q0 + 128 + off) : p.Skv;
    const int n = kv_end > 0 ? (kv_end + 63) / 64 : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        mbar_init(qdo_full, 1);
        for (int s = 0; s < NST; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&sdp_full[s], 1);
            mbar_init(&sdp_empty[s], NCOMPUTE);
        }
        mbar_init(ds_full, NCOMPUTE);
        mbar_init(dq_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (n > 0 && elect_one()) {
            auto load = [&](void* dst, const CUtensorMap* tm, uint64_t* bar, int d0, int r0, int hfirst) {
                tma_load_4d(dst, tm, bar, d0, hfirst ? h : r0, hfirst ? r0 : h, b);
            };
            mbar_arrive_expect_tx(qdo_full, 2 * T128);
            load(smem + OFF_Q, &tmQ, qdo_full, 0, q0, pa.q_hfirst);
            load(smem + OFF_Q + T128 / 2, &tmQ, qdo_full, 64, q0, pa.q_hfirst);
            load(smem + OFF_DO, &tmDO, qdo_full, 0, q0, pa.o_hfirst);
            load(smem + OFF_DO + T128 / 2, &tmDO, qdo_full, 64, q0, pa.o_hfirst);
            for (int j = 0; j < n; ++j) {
                const int st = j % NST;
                mbar_wait(&kv_empty[st], ((j / NST) & 1) ^ 1);
                mbar_arrive_expect_tx(&kv_full[st], 2 * T64);
                uint8_t* sk = smem + OFF_K + st * T64;
                uint8_t* sv = smem + OFF_V + st * T64;
                load(sk, &tmK, &kv_full[st], 0, j * 64, pa.k_hfirst);
                load(sk + T64 / 2, &tmK, &kv_full[st], 64, j * 64, pa.k_hfirst);
                load(sv, &tmV, &kv_full[st], 0, j * 64, pa.v_hfirst);
                load(sv + T64 / 2, &tmV, &kv_full[st], 64, j * 64, pa.v_hfirst);
            }
        }
    } else if (warp == 1) {
        if (n > 0) {
            constexpr uint32_t idesc_sdp = make_idesc_bf16(128, 64, 0u, 0u);
            constexpr uint32_t idesc_dq = make_idesc_bf16(128, 128, 0u, 1u);   // B = K_j read MN-major (dims contiguous)
            const uint32_t q_lo = (smem_u32(smem + OFF_Q) >> 4) & 0x3FFFu;
            const uint32_t do_lo = (smem_u32(smem + OFF_DO) >> 4) & 0x3FFFu;
            const uint32_t k_lo = (smem_u32(smem + OFF_K) >> 4) & 0x3FFFu;
            const uint32_t v_lo = (smem_u32(smem + OFF_V) >> 4) & 0x3FFFu;
            const uint32_t ds_lo = (smem_u32(smem + OFF_DS) >> 4) & 0x3FFFu;
            const bool issuer = elect_one();
            mbar_wait(qdo_full, 0);
            auto issue_sdp = [&](int j) {
                const int st = j % NST, ab = j & 1;      // K/V ring stage, S/dP accumulator buffer
                mbar_wait(&kv_full[st], (j / NST) & 1);
                mbar_wait(&sdp_empty[ab], ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                if (issuer) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint32_t ao = (kk >> 2) * (T128 / 2 >> 4) + (kk & 3) * 2;
                        const uint32_t bo = st * (T64 >> 4) + (kk >> 2) * (T64 / 2 >> 4) + (kk & 3) * 2;
                        umma_bf16_w(tmem_base + TM_S + ab * 64, q_lo + ao, k_lo + bo, DESC_HI, idesc_sdp, kk ? 1u : 0u);
                    }
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint32_t ao = (kk >> 2) * (T128 / 2 >> 4) + (kk & 3) * 2;
                        const uint32_t bo = st * (T64 >> 4) + (kk >> 2) * (T64 / 2 >> 4) + (kk & 3) * 2;
                        umma_bf16_w(tmem_base + TM_DP + ab * 64, do_lo + ao, v_lo + bo, DESC_HI, idesc_sdp, kk ? 1u : 0u);
                    }
                    umma_commit(&sdp_full[ab]);
                }
                __syncwarp();
            };
            issue_sdp(0);
            for (int j = 0; j < n; ++j) {
                if (j + 1 < n) issue_sdp(j + 1);
                const int st = j % NST;
                mbar_wait(ds_full, j & 1);
                tc_fence_after();
                if (issuer) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16_w(tmem_base + TM_DQ, ds_lo + kk * 2, (k_lo + st * (T64 >> 4) + kk * (2048 >> 4)) | LBO_8K, DESC_HI,
                                    idesc_dq, (j > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(&kv_empty[st]);
                    umma_commit(dq_done);
                }
                __syncwarp();
            }
        }
    } else {
        const int cw = warp - 2;
        const int quarter = warp & 3;
        const int half = cw >> 2;                         // which 32 of the step's 64 key columns
        const int row = quarter * 32 + lane;
        const int qi = q0 + row;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const int row_lim = CAUSAL ? qi + off : 0x7fffffff;
        const float c = p.scale_log2;
        float lse2 = INFINITY, dl = 0.f;                  // +inf: exp2(s*c - inf) = 0 for padded / fully-masked query rows
        if (qi < p.Sq) {
            const long long si = (static_cast<long long>(b) * p.H + h) * p.Sq + qi;
            const float l = p.lse[si];
            if (l != -INFINITY) lse2 = l * 1.4426950408889634f;
            dl = p.delta[si];
        }
        for (int j = 0; j < n; ++j) {
            const int kc = j * 64 + half * 32;            // first key of this warp's columns
            const int st = j & 1;
            mbar_wait(&sdp_full[st], (j >> 1) & 1);
            tc_fence_after();
            uint32_t rs[32], rd[32];
            tmem_ld_32x32(t_lane + TM_S + st * 64 + half * 32, rs);
            tmem_ld_32x32(t_lane + TM_DP + st * 64 + half * 32, rd);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&sdp_empty[st]);

            uint32_t w = 0xffffffffu;
            if (p.kmask != nullptr) {
                const uint32_t bit = (kc + lane < p.Skv) ? p.kmask[static_cast<long long>(b) * p.Skv + kc + lane] : 0u;
                w = __ballot_sync(0xffffffffu, bit != 0u);
            } else if (kc + 32 > p.Skv) {
                const int v = p.Skv - kc;
                w = v <= 0 ? 0u : ((1u << v) - 1u);
            }
            if (CAUSAL) {
                const int nb = row_lim - kc + 1;
                if (nb < 32) w &= nb <= 0 ? 0u : ((1u << nb) - 1u);
            }
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float p0 = ex2_approx(fmaf(__uint_as_float(rs[2 * i]), c, -lse2));
                float p1 = ex2_approx(fmaf(__uint_as_float(rs[2 * i + 1]), c, -lse2));
                if (!((w >> (2 * i)) & 1u)) p0 = 0.f;
                if (!((w >> (2 * i + 1)) & 1u)) p1 = 0.f;
                pk[i] = pack_bf16(p0 * (__uint_as_float(rd[2 * i]) - dl), p1 * (__uint_as_float(rd[2 * i + 1]) - dl));
            }
            if (j > 0) mbar_wait(dq_done, (j - 1) & 1);   // previous step's dQ MMAs have consumed the dS tile
            store_half_row(smem + OFF_DS, row, half, pk);
            fence_proxy_async_smem();
            mbar_arrive(ds_full);
        }
        if (n > 0) {
            mbar_wait(dq_done, (n - 1) & 1);
            tc_fence_after();
        }
        __nv_bfloat16* dst = p.dq + b * p.dq_bs + h * p.dq_hs + static_cast<long long>(qi) * p.dq_rs;
        store_grad_row(t_lane + TM_DQ, half, p.scale, p.rope_cos, p.rope_sin, qi, dst, qi < p.Sq, n > 0);
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ======================================================================================================== dK / dV
namespace dkv {
using namespace abt;
constexpr int OFF_K = 0;
constexpr int OFF_V = OFF_K + T128;
constexpr int NST = 3;                       // Q/dO ring depth (see dqk::NST; three stages is what fits beside K, V, P^T, dS^T)
constexpr int OFF_Q = OFF_V + T128;          // NST stages
constexpr int OFF_DO = OFF_Q + NST * T64;    // NST stages
constexpr int OFF_PT = OFF_DO + NST * T64;
constexpr int OFF_DST = OFF_PT + PS;
constexpr int OFF_STAT = OFF_DST + PS;       // [2 stages][lse2 64 | delta 64] fp32
constexpr int OFF_BAR = OFF_STAT + 2 * 128 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256;
constexpr int TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 384;
}  // namespace dkv

template <bool CAUSAL>
__global__ void __launch_bounds__(320, 1)
attn_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                       const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const AttnBwdTcArgs pa) {
    using namespace dkv;
    const AttnBwdArgs& p = pa.a;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* kv_full = bars;          // 1
    uint64_t* sdp_full = bars + 1;     // 2
    uint64_t* sdp_empty = bars + 3;    // 2
    uint64_t* ps_full = bars + 5;      // 1
    uint64_t* dvdk_done = bars + 6;    // 1
    uint64_t* qdo_full = bars + 8;     // NST
    uint64_t* qdo_empty = bars + 8 + NST;   // NST
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8 + 2 * NST);
    float* stat = reinterpret_cast<float*>(smem + OFF_STAT);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int kt = blockIdx.x;                 // early key tiles see the most query steps under the causal mask: first
    const int h = blockIdx.y, b = blockIdx.z;
    const int k0 = kt * 128;
    const int off = p.Skv - p.Sq;
    const int nq = (p.Sq + 63) / 64;
    int i_begin = 0;
    if (CAUSAL) i_begin = max(0, k0 - off) / 64;   // first query step with a row that can see key k0
    const int n = max(0, nq - i_begin);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        mbar_init(kv_full, 1);
        for (int s = 0; s < NST; ++s) {
            mbar_init(&qdo_full[s], 1);
            mbar_init(&qdo_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&sdp_full[s], 1);
            mbar_init(&sdp_empty[s], NCOMPUTE);
        }
        mbar_init(ps_full, NCOMPUTE);
        mbar_init(dvdk_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        if (n > 0 && elect_one()) {
            auto load = [&](void* dst, const CUtensorMap* tm, uint64_t* bar, int d0, int r0, int hfirst) {
                tma_load_4d(dst, tm, bar, d0, hfirst ? h : r0, hfirst ? r0 : h, b);
            };
            mbar_arrive_expect_tx(kv_full, 2 * T128);
            load(smem + OFF_K, &tmK, kv_full, 0, k0, pa.k_hfirst);
            load(smem + OFF_K + T128 / 2, &tmK, kv_full, 64, k0, pa.k_hfirst);
            load(smem + OFF_V, &tmV, kv_full, 0, k0, pa.v_hfirst);
            load(smem + OFF_V + T128 / 2, &tmV, kv_full, 64, k0, pa.v_hfirst);
            for (int t = 0; t < n; ++t) {
                const int st = t % NST;
                const int i0 = (i_begin + t) * 64;
                mbar_wait(&qdo_empty[st], ((t / NST) & 1) ^ 1);
                mbar_arrive_expect_tx(&qdo_full[st], 2 * T64);
                uint8_t* sq = smem + OFF_Q + st * T64;
                uint8_t* sd = smem + OFF_DO + st * T64;
                load(sq, &tmQ, &qdo_full[st], 0, i0, pa.q_hfirst);
                load(sq + T64 / 2, &tmQ, &qdo_full[st], 64, i0, pa.q_hfirst);
                load(sd, &tmDO, &qdo_full[st], 0, i0, pa.o_hfirst);
                load(sd + T64 / 2, &tmDO, &qdo_full[st], 64, i0, pa.o_hfirst);
            }
        }
    } else if (warp == 1) {
        if (n > 0) {
            constexpr uint32_t idesc_sdp = make_idesc_bf16(128, 64, 0u, 0u);
            constexpr uint32_t idesc_g = make_idesc_bf16(128, 128, 0u, 1u);   // B = dO_i / Q_i read MN-major (dims contiguous)
            const uint32_t k_lo = (smem_u32(smem + OFF_K) >> 4) & 0x3FFFu;
            const uint32_t v_lo = (smem_u32(smem + OFF_V) >> 4) & 0x3FFFu;
            const uint32_t q_lo = (smem_u32(smem + OFF_Q) >> 4) & 0x3FFFu;
            const uint32_t do_lo = (smem_u32(smem + OFF_DO) >> 4) & 0x3FFFu;
            const uint32_t pt_lo = (smem_u32(smem + OFF_PT) >> 4) & 0x3FFFu;
            const uint32_t dst_lo = (smem_u32(smem + OFF_DST) >> 4) & 0x3FFFu;
            const bool issuer = elect_one();
            mbar_wait(kv_full, 0);
            auto issue_sdp = [&](int t) {
                const int st = t % NST, ab = t & 1;      // Q/dO ring stage, S^T/dP^T accumulator buffer
                mbar_wait(&qdo_full[st], (t / NST) & 1);
                mbar_wait(&sdp_empty[ab], ((t >> 1) & 1) ^ 1);
                tc_fence_after();
                if (issuer) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {   // S^T = K Q^T
                        const uint32_t ao = (kk >> 2) * (T128 / 2 >> 4) + (kk & 3) * 2;
                        const uint32_t bo = st * (T64 >> 4) + (kk >> 2) * (T64 / 2 >> 4) + (kk & 3) * 2;
                        umma_bf16_w(tmem_base + TM_S + ab * 64, k_lo + ao, q_lo + bo, DESC_HI, idesc_sdp, kk ? 1u : 0u);
                    }
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {   // dP^T = V dO^T
                        const uint32_t ao = (kk >> 2) * (T128 / 2 >> 4) + (kk & 3) * 2;
                        const uint32_t bo = st * (T64 >> 4) + (kk >> 2) * (T64 / 2 >> 4) + (kk & 3) * 2;
                        umma_bf16_w(tmem_base + TM_DP + ab * 64, v_lo + ao, do_lo + bo, DESC_HI, idesc_sdp, kk ? 1u : 0u);
                    }
                    umma_commit(&sdp_full[ab]);
                }
                __syncwarp();
            };
            issue_sdp(0);
            for (int t = 0; t < n; ++t) {
                if (t + 1 < n) issue_sdp(t + 1);
                const int st = t % NST;
                mbar_wait(ps_full, t & 1);
                tc_fence_after();
                if (issuer) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)     // dV += P^T dO
                        umma_bf16_w(tmem_base + TM_DV, pt_lo + kk * 2, (do_lo + st * (T64 >> 4) + kk * (2048 >> 4)) | LBO_8K, DESC_HI,
                                    idesc_g, (t > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)     // dK += dS^T Q
                        umma_bf16_w(tmem_base + TM_DK, dst_lo + kk * 2, (q_lo + st * (T64 >> 4) + kk * (2048 >> 4)) | LBO_8K, DESC_HI,
                                    idesc_g, (t > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(&qdo_empty[st]);
                    umma_commit(dvdk_done);
                }
                __syncwarp();
            }
        }
    } else {
        const int cw = warp - 2;
        const int ct = cw * 32 + lane;                    // 0..255 among the compute threads
        const int quarter = warp & 3;
        const int half = cw >> 2;                         // which 32 of the step's 64 query columns
        const int row = quarter * 32 + lane;
        const int kj = k0 + row;                          // this thread's key
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const float c = p.scale_log2;
        bool key_ok = kj < p.Skv;
        if (key_ok && p.kmask != nullptr) key_ok = p.kmask[static_cast<long long>(b) * p.Skv + kj] != 0;
        const long long stat_base = (static_cast<long long>(b) * p.H + h) * p.Sq;
        // statistics of a 64-query step: threads 0..63 fetch lse (as log2, +inf = "no contribution"), 64..127 fetch delta
        auto fetch_stat = [&](int t) -> float {
            const int qr = (i_begin + t) * 64 + (ct & 63);
            if (ct >= 128 || t >= n) return 0.f;
            if (ct < 64) {
                if (qr >= p.Sq) return INFINITY;
                const float l = p.lse[stat_base + qr];
                return l == -INFINITY ? INFINITY : l * 1.4426950408889634f;
            }
            return qr < p.Sq ? p.delta[stat_base + qr] : 0.f;
        };
        float nxt = fetch_stat(0);
        for (int t = 0; t < n; ++t) {
            const int st = t & 1;
            const int i0 = (i_begin + t) * 64;
            if (ct < 128) stat[st * 128 + ct] = nxt;
            nxt = fetch_stat(t + 1);                      // in flight under this step's work
            bar_sync_compute();
            mbar_wait(&sdp_full[st], (t >> 1) & 1);
            tc_fence_after();
            uint32_t rs[32], rd[32];
            tmem_ld_32x32(t_lane + TM_S + st * 64 + half * 32, rs);
            tmem_ld_32x32(t_lane + TM_DP + st * 64 + half * 32, rd);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&sdp_empty[st]);

            // causal: query column cc (row i0 + cc) sees this key iff kj <= i0 + cc + off
            int cmin = 0;
            if (CAUSAL) cmin = kj - off - i0 - half * 32;
            if (!key_ok) cmin = 32;
            const float4* l4 = reinterpret_cast<const float4*>(stat + st * 128 + half * 32);
            const float4* d4 = reinterpret_cast<const float4*>(stat + st * 128 + 64 + half * 32);
            uint32_t pp[16], pd[16];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const float4 lv = l4[v], dv = d4[v];
                const float ll[4] = {lv.x, lv.y, lv.z, lv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
                float pr[4], ds[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int cc = v * 4 + e;
                    float x = ex2_approx(fmaf(__uint_as_float(rs[cc]), c, -ll[e]));
                    if (cc < cmin) x = 0.f;
                    pr[e] = x;
                    ds[e] = x * (__uint_as_float(rd[cc]) - dd[e]);
                }
                pp[v * 2] = pack_bf16(pr[0], pr[1]); pp[v * 2 + 1] = pack_bf16(pr[2], pr[3]);
                pd[v * 2] = pack_bf16(ds[0], ds[1]); pd[v * 2 + 1] = pack_bf16(ds[2], ds[3]);
            }
            if (t > 0) mbar_wait(dvdk_done, (t - 1) & 1);   // previous step's MMAs have consumed P^T / dS^T
            store_half_row(smem + OFF_PT, row, half, pp);
            store_half_row(smem + OFF_DST, row, half, pd);
            fence_proxy_async_smem();
            mbar_arrive(ps_full);
        }
        if (n > 0) {
            mbar_wait(dvdk_done, (n - 1) & 1);
            tc_fence_after();
        }
        const bool ok = kj < p.Skv;
        __nv_bfloat16* dvp = p.dv + b * p.dv_bs + h * p.dv_hs + static_cast<long long>(kj) * p.dv_rs;
        __nv_bfloat16* dkp = p.dk + b * p.dk_bs + h * p.dk_hs + static_cast<long long>(kj) * p.dk_rs;
        store_grad_row(t_lane + TM_DV, half, 1.f, nullptr, nullptr, kj, dvp, ok, n > 0);
        store_grad_row(t_lane + TM_DK, half, p.scale, p.rope_cos, p.rope_sin, kj, dkp, ok, n > 0);
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ host
template <bool CAUSAL>
static int launch_bwd_tc(const AttnBwdArgs& a, cudaStream_t stream) {
    auto kq = attn_bwd_dq_tc_kernel<CAUSAL>;
    auto kkv = attn_bwd_dkv_tc_kernel<CAUSAL>;
    static bool attr_set = false;
    if (!attr_set) {
        LHRS_CUDA(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, dqk::SMEM_BYTES));
        LHRS_CUDA(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, dkv::SMEM_BYTES));
        attr_set = true;
    }
    AttnBwdTcArgs pa;
    pa.a = a;
    CUtensorMap q128, do128, k64, v64, q64, do64, k128, v128;
    int rc, f;
    if ((rc = make_tmap_bshd(&q128, &pa.q_hfirst, a.q, 128, a.Sq, a.H, a.B, a.q_rs, a.q_hs, a.q_bs, 128))) return rc;
    if ((rc = make_tmap_bshd(&do128, &pa.o_hfirst, a.d_o, 128, a.Sq, a.H, a.B, a.o_rs, a.o_hs, a.o_bs, 128))) return rc;
    if ((rc = make_tmap_bshd(&k64, &pa.k_hfirst, a.k, 128, a.Skv, a.H, a.B, a.k_rs, a.k_hs, a.k_bs, 64))) return rc;
    if ((rc = make_tmap_bshd(&v64, &pa.v_hfirst, a.v, 128, a.Skv, a.H, a.B, a.v_rs, a.v_hs, a.v_bs, 64))) return rc;
    if ((rc = make_tmap_bshd(&q64, &f, a.q, 128, a.Sq, a.H, a.B, a.q_rs, a.q_hs, a.q_bs, 64))) return rc;
    if ((rc = make_tmap_bshd(&do64, &f, a.d_o, 128, a.Sq, a.H, a.B, a.o_rs, a.o_hs, a.o_bs, 64))) return rc;
    if ((rc = make_tmap_bshd(&k128, &f, a.k, 128, a.Skv, a.H, a.B, a.k_rs, a.k_hs, a.k_bs, 128))) return rc;
    if ((rc = make_tmap_bshd(&v128, &f, a.v, 128, a.Skv, a.H, a.B, a.v_rs, a.v_hs, a.v_bs, 128))) return rc;
    kq<<<dim3((a.Sq + 127) / 128, a.H, a.B), 320, dqk::SMEM_BYTES, stream>>>(q128, do128, k64, v64, pa);
    LHRS_LAUNCH_CHECK("attn_bwd_dq_tc_kernel");
    kkv<<<dim3((a.Skv + 127) / 128, a.H, a.B), 320, dkv::SMEM_BYTES, stream>>>(q64, do64, k128, v128, pa);
    LHRS_LAUNCH_CHECK("attn_bwd_dkv_tc_kernel");
    return LHRS_OK;
}

int attention_bwd_tc(const AttnBwdArgs& a, bool causal, cudaStream_t stream) {
    return causal ? launch_bwd_tc<true>(a, stream) : launch_bwd_tc<false>(a, stream);
}

}  // namespace lhrs
