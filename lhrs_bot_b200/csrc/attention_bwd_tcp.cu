// Flash attention backward on tcgen05 (head_dim 128), PERSISTENT and tile-pipelined form of attention_bwd_tc.cu — same
// arithmetic, same two deterministic passes (no atomics), same epilogues (scale, inverse RoPE, packed [rows, 3*dim] stores):
//
//   dQ kernel     work item = 128 query rows of one (batch, head); per 64-key step
//                   S = Q K^T, dP = dO V^T  (TMEM)  ->  dS = P o (dP - delta)  (bf16, smem SW128)  ->  dQ += dS K   (TMEM)
//   dK/dV kernel  work item = 128 keys of one (batch, head); per 64-query step
//                   S^T = K Q^T, dP^T = V dO^T  ->  P^T, dS^T (smem)  ->  dV += P^T dO,  dK += dS^T Q   (TMEM)
//
// Why persistent: with one CTA per tile (attention_bwd_tc.cu) a CTA lives for only ~5 steps at S = 512, and ncu showed the tensor
// pipe 20 % busy: 42 % of all warp samples sat on the first mbarrier wait — launch, tensor-map fetch, TMEM allocation, the first
// TMA round trip and the epilogue were paid per tile and hidden by nothing.  Here one CTA per SM walks a static list of items:
//   * every ring (K / V / Q / dO stages, S/dP accumulators, dS tiles) is indexed by a GLOBAL step counter that runs across
//     items, so the TMA producer and the MMA issuer flow from the last step of one item straight into the first of the next;
//   * dQ kernel: the dQ accumulator is double buffered in TMEM and drained by four dedicated epilogue warps while the next
//     item's steps run; dK/dV kernel: TMEM is full (S, dP double buffered + dV + dK), its compute warps drain between items;
//   * items are ordered so that the tiles running at the same time belong to the same few (batch, head) pairs — their K/V
//     (resp. Q/dO) tiles are then read from HBM once and hit in L2 for the other tiles — while every CTA still sees an even mix
//     of long and short causal tiles (`sched_item`).
//   * TWO MMA-issuing warps.  Measured on B200 (tools/ubench_tc.cu, profiles/r2_ubench_tc.txt): tcgen05.mma issue BLOCKS at the
//     execution rate (an N=64 MMA takes 48 cycles — smem-read bound — and the issuing thread gets the next one in ~48 cycles
//     later), a satisfied mbarrier wait costs ~82 cycles, and a batch's completion is seen ~250 cycles after its last issue.  A
//     single issuer that also waits on 4-5 barriers per step therefore left the tensor pipe idle for half of every step (the
//     device-side timeline, tools/attn_trace.py, showed 2600 cycles per step against 1024 cycles of MMAs).  Warp A issues the
//     score products (S, dP), warp B the gradient products (dQ, or dV and dK): each one's waits hide under the other's MMAs.
//   warp 0: TMA producer   warps 1-2: S and dP MMAs (warp 1 also owns the TMEM allocation)   warp 3: gradient MMAs
//   warps 4-11: compute   warps 12-15 (dQ kernel): epilogue
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <algorithm>

#include "../../include/lhrs_b200.h"
#include "attention_common.h"
#include "host_common.h"
#include "ptx.cuh"
#include "attention_bwd_tc.cuh"

namespace lhrs {

// Optional device-side timeline (debug builds only, -DLHRS_ATTN_TRACE; see tools/attn_trace.py): CTA 0 records
// (event, global step, clock64) per role into a global buffer.  Compiled out of the product library.
#ifdef LHRS_ATTN_TRACE
__device__ long long g_attn_trace[4 * 8192];
#define TRACE_DECL(role) long long* tr_ = g_attn_trace + (role) * 8192; int tr_n_ = 0; bool tr_on_ = (blockIdx.x == 0 && lane == 0);
#define TRACE(ev, g)                                                                                       \
    do {                                                                                                   \
        if (tr_on_ && tr_n_ < 4095) {                                                                      \
            tr_[2 * tr_n_] = (static_cast<long long>(ev) << 32) | static_cast<long long>(static_cast<uint32_t>(g)); \
            tr_[2 * tr_n_ + 1] = clock64();                                                                \
            ++tr_n_;                                                                                       \
        }                                                                                                  \
    } while (0)
#define TRACE_END() do { if (tr_on_) { tr_[2 * tr_n_] = -1; } } while (0)
#else
#define TRACE_DECL(role)
#define TRACE(ev, g)
#define TRACE_END()
#endif

// Static item order.  Items are (head pair bh, tile t).  A chunk = C heads x nt tiles; inside a chunk the tile index of
// slot s is rotated by the chunk number, so a CTA that always lands on the same slot still cycles through every tile index
// (short and long causal rows) as it moves from chunk to chunk.
struct BwdSched {
    int nt;      // tiles per (batch, head)
    int C;       // heads per chunk
    int total;   // item slots (some past the last head: skipped); < 2^20
    int BH, H;
    // every role decodes every item: the four divisions are multiplications by ceil(2^40 / d) (exact for n, d < 2^20) —
    // as hardware divisions they cost ~600 cycles per decode on the device-side timeline
    unsigned long long m_per, m_C, m_nt, m_H;
};
__device__ __forceinline__ int fast_div(int n, unsigned long long m) {
    return static_cast<int>((static_cast<unsigned long long>(n) * m) >> 40);
}
__device__ __forceinline__ bool sched_item(const BwdSched& s, int w, int& b, int& h, int& t) {
    const int per = s.C * s.nt;
    const int chunk = fast_div(w, s.m_per), r = w - chunk * per;
    const int slot = fast_div(r, s.m_C), hc = r - slot * s.C;
    const int bh = chunk * s.C + hc;
    if (bh >= s.BH) return false;
    const int sc = slot + chunk;
    t = sc - fast_div(sc, s.m_nt) * s.nt;
    b = fast_div(bh, s.m_H);
    h = bh - b * s.H;
    return true;
}

// ======================================================================================================== dQ
namespace dqp {
using namespace abt;
constexpr int NKV = 4;                       // K/V ring (one barrier per stage): a stage stays until dQ(j) has read K_j
constexpr int OFF_Q = 0;
constexpr int OFF_DO = OFF_Q + T128;
constexpr int OFF_K = OFF_DO + T128;         // NKV stages
constexpr int OFF_V = OFF_K + NKV * T64;     // NKV stages
constexpr int OFF_DS = OFF_V + NKV * T64;    // 2 buffers
constexpr int OFF_KBITS = OFF_DS + 2 * PS;   // key-padding mask of the item's batch row as bits, one copy per compute group (Skv <= 4096)
constexpr int OFF_BAR = OFF_KBITS + 1024;
constexpr int SMEM_BYTES = OFF_BAR + 512;
constexpr int TM_S = 0, TM_DP = 128, TM_DQ = 256;   // S 2x64 | dP 2x64 | dQ 2x128
constexpr int THREADS = 512;
constexpr int MAX_SKV = 4096;
}  // namespace dqp

template <bool CAUSAL>
__global__ void __launch_bounds__(dqp::THREADS, 1)
attn_bwd_dq_tcp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                       const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const AttnBwdTcArgs pa,
                       const BwdSched sc) {
    using namespace dqp;
    const AttnBwdArgs& p = pa.a;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* qdo_full = bars;            // 1
    uint64_t* qdo_empty = bars + 1;       // 1
    uint64_t* sdp_full = bars + 2;        // 2
    uint64_t* sdp_empty = bars + 4;       // 2
    uint64_t* ds_full = bars + 6;         // 2
    uint64_t* ds_empty = bars + 8;        // 2
    uint64_t* dq_full = bars + 10;        // 2
    uint64_t* dq_empty = bars + 12;       // 2
    uint64_t* kv_full = bars + 14;          // NKV
    uint64_t* kv_empty = bars + 14 + NKV;   // NKV
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14 + 2 * NKV);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int off = p.Skv - p.Sq;
    const int G = static_cast<int>(gridDim.x);

    // item -> geometry; n = number of 64-key steps (0: the tile sees no key)
    // ragged batch (p.seq_off): the sequence's rows start at row rb of the one row space every operand lives in, and there are
    // lq = lk of them; rows past the end are the next sequence's (finite) values — the length masks keep them out of every sum
    auto item = [&](int w, int& b, int& h, int& q0, int& n, int& rb, int& lq, int& lk) -> bool {
        int t;
        if (!sched_item(sc, w, b, h, t)) return false;
        q0 = t * 128;
        rb = 0; lq = p.Sq; lk = p.Skv;
        if (p.seq_off != nullptr) { rb = p.seq_off[b]; lq = lk = p.seq_off[b + 1] - rb; }
        const int kv_end = CAUSAL ? min(lk, q0 + 128 + off) : lk;
        n = (kv_end > 0 && q0 < lq) ? (kv_end + 63) / 64 : 0;
        return true;
    };
    const bool ragged = p.seq_off != nullptr;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        mbar_init(qdo_full, 1);
        mbar_init(qdo_empty, 2);               // issuers A1 (Q) and A2 (dO)
        for (int s = 0; s < 2; ++s) {
            mbar_init(&sdp_full[s], 2);        // S from issuer A1, dP from issuer A2
            mbar_init(&sdp_empty[s], 128);     // S/dP accumulator s, dS buffer s belong to compute group s (128 threads)
            mbar_init(&ds_full[s], 128);
            mbar_init(&ds_empty[s], 1);
            mbar_init(&dq_full[s], 1);
            mbar_init(&dq_empty[s], 128);
        }
        for (int s = 0; s < NKV; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
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
        // ===================================================================== TMA producer (one thread)
        if (elect_one()) {
            TRACE_DECL(0)
            int it = 0;
            uint32_t g = 0;
            for (int w = blockIdx.x; w < sc.total; w += G) {
                int b, h, q0, n, rb, lq, lk;
                if (!item(w, b, h, q0, n, rb, lq, lk) || n == 0) continue;
                auto load = [&](void* dst, const CUtensorMap* tm, uint64_t* bar, int d0, int r0, int hfirst) {
                    tma_load_4d(dst, tm, bar, d0, hfirst ? h : r0 + rb, hfirst ? r0 + rb : h, ragged ? 0 : b);
                };
                auto load_kv = [&](int j) {
                    const uint32_t st = g % NKV;
                    uint8_t* kd = smem + OFF_K + st * T64;
                    uint8_t* vd = smem + OFF_V + st * T64;
                    mbar_wait(&kv_empty[st], ((g / NKV) & 1) ^ 1);
                    TRACE(1, g);
                    mbar_arrive_expect_tx(&kv_full[st], 2 * T64);
                    load(kd, &tmK, &kv_full[st], 0, j * 64, pa.k_hfirst);
                    load(kd + T64 / 2, &tmK, &kv_full[st], 64, j * 64, pa.k_hfirst);
                    load(vd, &tmV, &kv_full[st], 0, j * 64, pa.v_hfirst);
                    load(vd + T64 / 2, &tmV, &kv_full[st], 64, j * 64, pa.v_hfirst);
                    ++g;
                };
                // the first K / V steps only need free ring slots: get them in flight before waiting for the Q / dO buffer
                const int pre = n < 2 ? n : 2;
                for (int j = 0; j < pre; ++j) load_kv(j);
                mbar_wait(qdo_empty, (it & 1) ^ 1);          // the previous item's last S/dP MMAs have read Q / dO
                TRACE(2, g);
                mbar_arrive_expect_tx(qdo_full, 2 * T128);
                load(smem + OFF_Q, &tmQ, qdo_full, 0, q0, pa.q_hfirst);
                load(smem + OFF_Q + T128 / 2, &tmQ, qdo_full, 64, q0, pa.q_hfirst);
                load(smem + OFF_DO, &tmDO, qdo_full, 0, q0, pa.o_hfirst);
                load(smem + OFF_DO + T128 / 2, &tmDO, qdo_full, 64, q0, pa.o_hfirst);
                for (int j = pre; j < n; ++j) load_kv(j);
                ++it;
            }
            TRACE_END();
        }
    } else if (warp == 1 || warp == 2) {
        // ===================================================================== MMA issuers A1 (S = Q K^T) and A2 (dP = dO V^T)
        // Two warps, not one: a tcgen05.mma issue blocks for about its execution time, so a lone issuer's barrier waits
        // (~80 cycles each even when satisfied) are dead tensor-pipe time.  Both arrive on sdp_full / qdo_empty (count 2).
        // (Measured alternatives, tools/attn_bench.py at b16 s512 / b4 s2048: one issuer for S, dP and dQ 304 / 706 us; S+dP and
        //  dQ issuers 294 / 665 us; this split 284 / 646 us; issuers split by step parity instead 292 / 677 us.)
        TRACE_DECL(1)
        const bool is_dp = (warp == 2);
        constexpr uint32_t idesc_sdp = make_idesc_bf16(128, 64, 0u, 0u);
        const uint32_t a_lo = (smem_u32(smem + (is_dp ? OFF_DO : OFF_Q)) >> 4) & 0x3FFFu;
        const uint32_t b_lo = (smem_u32(smem + (is_dp ? OFF_V : OFF_K)) >> 4) & 0x3FFFu;
        const uint32_t t_acc = tmem_base + (is_dp ? TM_DP : TM_S);
        const bool issuer = elect_one();
#ifdef LHRS_ATTN_TRACE
        if (is_dp) tr_on_ = false;
#endif
        int it = 0;
        uint32_t g = 0;
        for (int w = blockIdx.x; w < sc.total; w += G) {
            int b, h, q0, n, rb, lq, lk;
            if (!item(w, b, h, q0, n, rb, lq, lk) || n == 0) continue;
            mbar_wait(qdo_full, it & 1);
            for (int j = 0; j < n; ++j, ++g) {
                const uint32_t st = g % NKV, ab = g & 1;
                TRACE(11, g);
                mbar_wait(&kv_full[st], (g / NKV) & 1);
                mbar_wait(&sdp_empty[ab], ((g >> 1) & 1) ^ 1);
                TRACE(13, g);
                tc_fence_after();
                if (issuer) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint32_t ao = (kk >> 2) * (T128 / 2 >> 4) + (kk & 3) * 2;
                        const uint32_t bo = st * (T64 >> 4) + (kk >> 2) * (T64 / 2 >> 4) + (kk & 3) * 2;
                        umma_bf16_w(t_acc + ab * 64, a_lo + ao, b_lo + bo, DESC_HI, idesc_sdp, kk ? 1u : 0u);
                    }
                    umma_commit(&sdp_full[ab]);
                    if (j == n - 1) umma_commit(qdo_empty);   // Q (resp. dO) of this item is dead: the producer may refill it
                }
                __syncwarp();
                TRACE(14, g);
            }
            ++it;
        }
        TRACE_END();
    } else if (warp == 3) {
        // ===================================================================== MMA issuer B: dQ += dS K
        // (every operand of dQ(g) is ready when ds_full(g) completes: dS(g) was computed from S/dP(g), whose MMAs read the
        //  same K/V stage — so releasing the stage after dQ(g) also covers warp A's reads of it)
        TRACE_DECL(3)
        constexpr uint32_t idesc_dq = make_idesc_bf16(128, 128, 0u, 1u);   // B = K_j read MN-major (dims contiguous)
        const uint32_t k_lo = (smem_u32(smem + OFF_K) >> 4) & 0x3FFFu;
        const uint32_t ds_lo = (smem_u32(smem + OFF_DS) >> 4) & 0x3FFFu;
        const bool issuer = elect_one();
        int it = 0;
        uint32_t g = 0;
        for (int w = blockIdx.x; w < sc.total; w += G) {
            int b, h, q0, n, rb, lq, lk;
            if (!item(w, b, h, q0, n, rb, lq, lk) || n == 0) continue;
            const uint32_t buf = it & 1;
            mbar_wait(&dq_empty[buf], ((it >> 1) & 1) ^ 1);       // the epilogue warps have drained this accumulator
            for (int j = 0; j < n; ++j, ++g) {
                const uint32_t db = g & 1, st = g % NKV;
                TRACE(15, g);
                mbar_wait(&ds_full[db], (g >> 1) & 1);
                TRACE(16, g);
                tc_fence_after();
                if (issuer) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16_w(tmem_base + TM_DQ + buf * 128, ds_lo + db * (PS >> 4) + kk * 2,
                                    (k_lo + st * (T64 >> 4) + kk * (2048 >> 4)) | LBO_8K, DESC_HI, idesc_dq, (j > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(&kv_empty[st]);
                    umma_commit(&ds_empty[db]);
                    if (j == n - 1) umma_commit(&dq_full[buf]);
                }
                __syncwarp();
                TRACE(17, g);
            }
            ++it;
        }
        TRACE_END();
    } else if (warp >= 4 && warp < 12) {
        // ===================================================================== compute warps: two groups of four, ping-pong
        // Group k owns the global steps g with (g & 1) == k, and with them S/dP accumulator k and dS buffer k: while one
        // group runs its exponentials the other one is in its waits / TMEM loads / smem publish, so the per-step fixed costs
        // (two barrier waits, fence.proxy.async, arrive: ~500 cycles of a ~1400-cycle step when all eight warps worked on
        // the same step) are paid once per 64 columns instead of once per 32 and overlap the other group's MUFU work.
        const int cw = warp - 4;
        const int grp = cw >> 2;
        const int gt = (cw & 3) * 32 + lane;              // 0..127 inside the group
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const float c = p.scale_log2;
        uint32_t g = 0;
#ifdef LHRS_ATTN_TRACE
        long long* tr_ = g_attn_trace + 2 * 8192; int tr_n_ = 0; const bool tr_on_ = (blockIdx.x == 0 && threadIdx.x == 128);
#endif
        // per-item row data (lse, delta, mask bytes) is pulled towards L1 one item ahead: its ~1500-cycle miss latency would
        // otherwise sit at the head of every item, three times in a row (measured: 5500 cycles per item)
        auto prefetch_item = [&](int w0) {
            for (int w = w0; w < sc.total; w += G) {
                int b, h, q0, n, rb, lq, lk;
                if (!item(w, b, h, q0, n, rb, lq, lk) || n == 0) continue;
                const int qi = q0 + row;
                if (qi < lq) {
                    const long long si = (static_cast<long long>(b) * p.H + h) * p.Sq + qi;
                    prefetch_l1(p.lse + si);
                    prefetch_l1(p.delta + si);
                }
                if (p.kmask != nullptr && gt * 32 < 2 * n) prefetch_l1(p.kbits + static_cast<long long>(b) * p.kbits_w + gt * 32);
                return;
            }
        };
        prefetch_item(blockIdx.x);
        for (int w = blockIdx.x; w < sc.total; w += G) {
            int b, h, q0, n, rb, lq, lk;
            if (!item(w, b, h, q0, n, rb, lq, lk) || n == 0) continue;
            TRACE(20, g);
            const int qi = q0 + row;
            const int row_lim = CAUSAL ? qi + off : 0x7fffffff;
            float lse2 = INFINITY, dl = 0.f;              // +inf: exp2(s*c - inf) = 0 for padded / fully-masked query rows
            if (qi < lq) {
                const long long si = (static_cast<long long>(b) * p.H + h) * p.Sq + qi;
                const float l = p.lse[si];
                if (l != -INFINITY) lse2 = l * 1.4426950408889634f;
                dl = p.delta[si];
            }
            // key-padding mask of batch row b as bits (packed by the delta pre-pass): word t covers keys [32t, 32t+32)
            const uint32_t* kb = (p.kmask != nullptr) ? p.kbits + static_cast<long long>(b) * p.kbits_w : nullptr;
#ifdef LHRS_ATTN_TRACE
            if (lse2 == 12345.f || dl == 12345.f) printf("x");   // force the lse / delta loads to have landed before the next event
            TRACE(33, g);
#endif
            prefetch_item(w + G);
            TRACE(34, g);
            for (int j = 0; j < n; ++j, ++g) {
                if ((g & 1u) != static_cast<uint32_t>(grp)) continue;
                const uint32_t st = grp;
                uint2 kw = make_uint2(0xffffffffu, 0xffffffffu);       // mask words of this step's two halves (in flight under the wait)
                if (kb != nullptr) kw = __ldg(reinterpret_cast<const uint2*>(kb + 2 * j));
                TRACE(21, g);
                mbar_wait(&sdp_full[st], (g >> 1) & 1);
                TRACE(22, g);
                tc_fence_after();
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    const int kc = j * 64 + half * 32;        // first key of these 32 columns
                    uint32_t rs[32], rd[32];
                    tmem_ld_32x32(t_lane + TM_S + st * 64 + half * 32, rs);
                    tmem_ld_32x32(t_lane + TM_DP + st * 64 + half * 32, rd);
                    tmem_ld_wait();
                    if (half == 1) {                          // both halves are in registers: warp A may overwrite S/dP[st]
                        tc_fence_before();
                        mbar_arrive(&sdp_empty[st]);
                    }
                    uint32_t wm = half ? kw.y : kw.x;
                    if (kb == nullptr && kc + 32 > lk) {
                        const int v = lk - kc;
                        wm = v <= 0 ? 0u : ((1u << v) - 1u);
                    }
                    if (CAUSAL) {
                        const int nb = row_lim - kc + 1;
                        if (nb < 32) wm &= nb <= 0 ? 0u : ((1u << nb) - 1u);
                    }
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        float p0 = ex2_approx(fmaf(__uint_as_float(rs[2 * i]), c, -lse2));
                        float p1 = ex2_approx(fmaf(__uint_as_float(rs[2 * i + 1]), c, -lse2));
                        if (!((wm >> (2 * i)) & 1u)) p0 = 0.f;
                        if (!((wm >> (2 * i + 1)) & 1u)) p1 = 0.f;
                        pk[i] = pack_bf16(p0 * (__uint_as_float(rd[2 * i]) - dl), p1 * (__uint_as_float(rd[2 * i + 1]) - dl));
                    }
                    if (half == 0) {
                        TRACE(24, g);
                        mbar_wait(&ds_empty[st], ((g >> 1) & 1) ^ 1);   // the dQ MMAs of global step g-2 have consumed this dS buffer
                        TRACE(25, g);
                    }
                    store_half_row(smem + OFF_DS + st * PS, row, half, pk);
                }
                fence_proxy_async_smem();
                mbar_arrive(&ds_full[st]);
                TRACE(26, g);
            }
        }
        TRACE_END();
    } else if (warp >= 12) {
        // ===================================================================== epilogue warps (4): drain dQ while the next item runs
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        int it = 0;
#ifdef LHRS_ATTN_TRACE
        long long* tr_ = g_attn_trace + 3 * 8192; int tr_n_ = 0; const bool tr_on_ = false;   // (role 3 is MMA issuer B)
#endif
        for (int w = blockIdx.x; w < sc.total; w += G) {
            int b, h, q0, n, rb, lq, lk;
            if (!item(w, b, h, q0, n, rb, lq, lk)) continue;
            const int qi = q0 + row;
            const bool ok = qi < lq;
            __nv_bfloat16* dst = p.dq + b * p.dq_bs + h * p.dq_hs + static_cast<long long>(rb + qi) * p.dq_rs;
            if (n == 0) {                                  // no key visible from this tile: the gradient is zero
                store_grad_row(0u, 0, 0.f, nullptr, nullptr, qi, dst, ok, false);
                store_grad_row(0u, 1, 0.f, nullptr, nullptr, qi, dst, ok, false);
                continue;
            }
            const uint32_t buf = it & 1;
            TRACE(30, it);
            mbar_wait(&dq_full[buf], (it >> 1) & 1);
            TRACE(31, it);
            tc_fence_after();
            const uint32_t taddr = t_lane + TM_DQ + buf * 128;
            store_grad_row(taddr, 0, p.scale, p.rope_cos, p.rope_sin, qi, dst, ok, true);
            store_grad_row(taddr, 1, p.scale, p.rope_cos, p.rope_sin, qi, dst, ok, true);
            tc_fence_before();
            mbar_arrive(&dq_empty[buf]);
            TRACE(32, it);
            ++it;
        }
        TRACE_END();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ======================================================================================================== dK / dV
namespace dkp {
using namespace abt;
constexpr int NST = 3;                       // Q/dO ring: a stage stays until dV / dK of its step have read it
constexpr int OFF_K = 0;
constexpr int OFF_V = OFF_K + T128;
constexpr int OFF_Q = OFF_V + T128;          // NST stages
constexpr int OFF_DO = OFF_Q + NST * T64;    // NST stages
constexpr int OFF_PT = OFF_DO + NST * T64;   // 2 buffers (one per compute group)
constexpr int OFF_DST = OFF_PT + 2 * PS;     // 2 buffers
constexpr int OFF_STAT = OFF_DST + 2 * PS;   // [2 groups][2 buffers][lse2 64 | delta 64] fp32
constexpr int OFF_BAR = OFF_STAT + 2 * 2 * 128 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 512;
constexpr int TM_S = 0, TM_DP = 128, TM_DV = 256, TM_DK = 384;
constexpr int THREADS = 384;
}  // namespace dkp

template <bool CAUSAL>
__global__ void __launch_bounds__(dkp::THREADS, 1)
attn_bwd_dkv_tcp_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                        const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const AttnBwdTcArgs pa,
                        const BwdSched sc) {
    using namespace dkp;
    const AttnBwdArgs& p = pa.a;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* kv_full = bars;             // 1
    uint64_t* kv_empty = bars + 1;        // 1
    uint64_t* sdp_full = bars + 2;        // 2
    uint64_t* sdp_empty = bars + 4;       // 2
    uint64_t* ps_full = bars + 6;         // 2: P^T / dS^T buffer k published by compute group k
    uint64_t* ps_empty = bars + 8;        // 2: ... consumed by the dV / dK MMAs
    uint64_t* acc_full = bars + 10;       // 1: the item's last dV / dK MMAs have landed
    uint64_t* acc_empty = bars + 11;      // 1: the compute warps have drained dV / dK of the item
    uint64_t* qdo_full = bars + 12;       // NST
    uint64_t* qdo_empty = bars + 12 + NST;   // NST
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12 + 2 * NST);
    float* stat = reinterpret_cast<float*>(smem + OFF_STAT);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int off = p.Skv - p.Sq;
    const int nq = (p.Sq + 63) / 64;
    const int G = static_cast<int>(gridDim.x);

    // item -> geometry; the item's query steps are i_begin .. i_begin + n - 1 (64 rows each)
    auto item = [&](int w, int& b, int& h, int& k0, int& i_begin, int& n, int& rb, int& lq, int& lk) -> bool {
        int t;
        if (!sched_item(sc, w, b, h, t)) return false;
        k0 = t * 128;
        rb = 0; lq = p.Sq; lk = p.Skv;
        int nq_b = nq;
        if (p.seq_off != nullptr) { rb = p.seq_off[b]; lq = lk = p.seq_off[b + 1] - rb; nq_b = (lq + 63) / 64; }   // (see the dQ kernel)
        i_begin = 0;
        if (CAUSAL) i_begin = max(0, k0 - off) / 64;      // first query step with a row that can see key k0
        n = k0 < lk ? max(0, nq_b - i_begin) : 0;
        return true;
    };
    const bool ragged = p.seq_off != nullptr;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        mbar_init(kv_full, 1);
        mbar_init(kv_empty, 2);                // issuers A1 (K) and A2 (V)
        for (int s = 0; s < NST; ++s) {
            mbar_init(&qdo_full[s], 1);
            mbar_init(&qdo_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&sdp_full[s], 2);        // S^T from issuer A1, dP^T from issuer A2
            mbar_init(&sdp_empty[s], 128);
            mbar_init(&ps_full[s], 128);
            mbar_init(&ps_empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, NCOMPUTE);
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
        // ===================================================================== TMA producer (one thread)
        if (elect_one()) {
            int it = 0;
            uint32_t g = 0;
            for (int w = blockIdx.x; w < sc.total; w += G) {
                int b, h, k0, i_begin, n, rb, lq, lk;
                if (!item(w, b, h, k0, i_begin, n, rb, lq, lk) || n == 0) continue;
                auto load = [&](void* dst, const CUtensorMap* tm, uint64_t* bar, int d0, int r0, int hfirst) {
                    tma_load_4d(dst, tm, bar, d0, hfirst ? h : r0 + rb, hfirst ? r0 + rb : h, ragged ? 0 : b);
                };
                auto load_qdo = [&](int t) {
                    const uint32_t st = g % NST;
                    const int i0 = (i_begin + t) * 64;
                    mbar_wait(&qdo_empty[st], ((g / NST) & 1) ^ 1);
                    mbar_arrive_expect_tx(&qdo_full[st], 2 * T64);
                    uint8_t* sq = smem + OFF_Q + st * T64;
                    uint8_t* sd = smem + OFF_DO + st * T64;
                    load(sq, &tmQ, &qdo_full[st], 0, i0, pa.q_hfirst);
                    load(sq + T64 / 2, &tmQ, &qdo_full[st], 64, i0, pa.q_hfirst);
                    load(sd, &tmDO, &qdo_full[st], 0, i0, pa.o_hfirst);
                    load(sd + T64 / 2, &tmDO, &qdo_full[st], 64, i0, pa.o_hfirst);
                    ++g;
                };
                // the first Q / dO steps only need free ring slots: get them in flight before waiting for the K / V buffer
                const int pre = n < 2 ? n : 2;
                for (int t = 0; t < pre; ++t) load_qdo(t);
                mbar_wait(kv_empty, (it & 1) ^ 1);           // the previous item's last S^T/dP^T MMAs have read K / V
                mbar_arrive_expect_tx(kv_full, 2 * T128);
                load(smem + OFF_K, &tmK, kv_full, 0, k0, pa.k_hfirst);
                load(smem + OFF_K + T128 / 2, &tmK, kv_full, 64, k0, pa.k_hfirst);
                load(smem + OFF_V, &tmV, kv_full, 0, k0, pa.v_hfirst);
                load(smem + OFF_V + T128 / 2, &tmV, kv_full, 64, k0, pa.v_hfirst);
                for (int t = pre; t < n; ++t) load_qdo(t);
                ++it;
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ===================================================================== MMA issuers A1 (S^T = K Q^T) and A2 (dP^T = V dO^T)
        // (two warps so that one's barrier waits fall under the other's MMAs: see the dQ kernel)
        const bool is_dp = (warp == 2);
        constexpr uint32_t idesc_sdp = make_idesc_bf16(128, 64, 0u, 0u);
        const uint32_t a_lo = (smem_u32(smem + (is_dp ? OFF_V : OFF_K)) >> 4) & 0x3FFFu;
        const uint32_t b_lo = (smem_u32(smem + (is_dp ? OFF_DO : OFF_Q)) >> 4) & 0x3FFFu;
        const uint32_t t_acc = tmem_base + (is_dp ? TM_DP : TM_S);
        const bool issuer = elect_one();
        int it = 0;
        uint32_t g = 0;
        for (int w = blockIdx.x; w < sc.total; w += G) {
            int b, h, k0, i_begin, n, rb, lq, lk;
            if (!item(w, b, h, k0, i_begin, n, rb, lq, lk) || n == 0) continue;
            mbar_wait(kv_full, it & 1);
            for (int t = 0; t < n; ++t, ++g) {
                const uint32_t st = g % NST, ab = g & 1;
                mbar_wait(&qdo_full[st], (g / NST) & 1);
                mbar_wait(&sdp_empty[ab], ((g >> 1) & 1) ^ 1);
                tc_fence_after();
                if (issuer) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint32_t ao = (kk >> 2) * (T128 / 2 >> 4) + (kk & 3) * 2;
                        const uint32_t bo = st * (T64 >> 4) + (kk >> 2) * (T64 / 2 >> 4) + (kk & 3) * 2;
                        umma_bf16_w(t_acc + ab * 64, a_lo + ao, b_lo + bo, DESC_HI, idesc_sdp, kk ? 1u : 0u);
                    }
                    umma_commit(&sdp_full[ab]);
                    if (t == n - 1) umma_commit(kv_empty);        // K (resp. V) of this item is dead: the producer may refill it
                }
                __syncwarp();
            }
            ++it;
        }
    } else if (warp == 3) {
        // ===================================================================== MMA issuer B: dV += P^T dO, dK += dS^T Q
        // (ps_full(g) implies the S^T/dP^T MMAs of step g — which read the same Q/dO stage — have completed)
        constexpr uint32_t idesc_g = make_idesc_bf16(128, 128, 0u, 1u);   // B = dO_i / Q_i read MN-major (dims contiguous)
        const uint32_t q_lo = (smem_u32(smem + OFF_Q) >> 4) & 0x3FFFu;
        const uint32_t do_lo = (smem_u32(smem + OFF_DO) >> 4) & 0x3FFFu;
        const uint32_t pt_lo = (smem_u32(smem + OFF_PT) >> 4) & 0x3FFFu;
        const uint32_t dst_lo = (smem_u32(smem + OFF_DST) >> 4) & 0x3FFFu;
        const bool issuer = elect_one();
        int it = 0;
        uint32_t g = 0;
        for (int w = blockIdx.x; w < sc.total; w += G) {
            int b, h, k0, i_begin, n, rb, lq, lk;
            if (!item(w, b, h, k0, i_begin, n, rb, lq, lk) || n == 0) continue;
            mbar_wait(acc_empty, (it & 1) ^ 1);                   // dV / dK of the previous item have been drained
            for (int t = 0; t < n; ++t, ++g) {
                const uint32_t st = g % NST, pb = g & 1;
                mbar_wait(&ps_full[pb], (g >> 1) & 1);
                tc_fence_after();
                if (issuer) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)     // dV += P^T dO
                        umma_bf16_w(tmem_base + TM_DV, pt_lo + pb * (PS >> 4) + kk * 2,
                                    (do_lo + st * (T64 >> 4) + kk * (2048 >> 4)) | LBO_8K, DESC_HI, idesc_g, (t > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)     // dK += dS^T Q
                        umma_bf16_w(tmem_base + TM_DK, dst_lo + pb * (PS >> 4) + kk * 2,
                                    (q_lo + st * (T64 >> 4) + kk * (2048 >> 4)) | LBO_8K, DESC_HI, idesc_g, (t > 0 || kk > 0) ? 1u : 0u);
                    umma_commit(&qdo_empty[st]);
                    umma_commit(&ps_empty[pb]);
                    if (t == n - 1) umma_commit(acc_full);
                }
                __syncwarp();
            }
            ++it;
        }
    } else if (warp >= 4) {
        // ===================================================================== compute warps: two groups of four, ping-pong
        // (group k owns the global steps with (g & 1) == k, S^T/dP^T accumulator k and P^T/dS^T buffer k; see the dQ kernel).
        // After an item's last step all eight warps drain dV / dK (group k stores rotation pairs [32k, 32k+32)).
        const int cw = warp - 4;
        const int grp = cw >> 2;
        const int gt = (cw & 3) * 32 + lane;              // 0..127 inside the group
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const float c = p.scale_log2;
        float* gstat2 = stat + grp * 256;                 // two buffers per group: one group barrier per step instead of two
        uint32_t g = 0, ks = 0;                           // ks: this group's step counter (selects the statistics buffer)
        int it = 0;
        auto prefetch_item = [&](int w0) {
            for (int w = w0; w < sc.total; w += G) {
                int b, h, k0, i_begin, n, rb, lq, lk;
                if (!item(w, b, h, k0, i_begin, n, rb, lq, lk) || n == 0) continue;
                const int kj = k0 + row;
                if (p.kmask != nullptr && kj < p.Skv) prefetch_l1(p.kmask + static_cast<long long>(b) * p.Skv + kj);
                const long long stat_base = (static_cast<long long>(b) * p.H + h) * p.Sq;
                const int qr = i_begin * 64 + gt;          // the first two steps' statistics (128 consecutive rows)
                if (qr < lq) { prefetch_l1(p.lse + stat_base + qr); prefetch_l1(p.delta + stat_base + qr); }
                return;
            }
        };
        prefetch_item(blockIdx.x);
        for (int w = blockIdx.x; w < sc.total; w += G) {
            int b, h, k0, i_begin, n, rb, lq, lk;
            if (!item(w, b, h, k0, i_begin, n, rb, lq, lk)) continue;
            const int kj = k0 + row;                          // this thread's key
            const bool ok = kj < lk;
            __nv_bfloat16* dvp = p.dv + b * p.dv_bs + h * p.dv_hs + static_cast<long long>(rb + kj) * p.dv_rs;
            __nv_bfloat16* dkp = p.dk + b * p.dk_bs + h * p.dk_hs + static_cast<long long>(rb + kj) * p.dk_rs;
            if (n == 0) {                                     // no query sees this key tile: zero gradients
                store_grad_row(0u, grp, 0.f, nullptr, nullptr, kj, dvp, ok, false);
                store_grad_row(0u, grp, 0.f, nullptr, nullptr, kj, dkp, ok, false);
                continue;
            }
            bool key_ok = ok;
            if (key_ok && p.kmask != nullptr) key_ok = p.kmask[static_cast<long long>(b) * p.Skv + kj] != 0;
            const long long stat_base = (static_cast<long long>(b) * p.H + h) * p.Sq;
            // statistics of a 64-query step: the group's threads 0..63 fetch lse (as log2, +inf = "no contribution"), 64..127 delta
            auto fetch_stat = [&](int t) -> float {
                const int qr = (i_begin + t) * 64 + (gt & 63);
                if (t >= n) return 0.f;
                if (gt < 64) {
                    if (qr >= lq) return INFINITY;
                    const float l = p.lse[stat_base + qr];
                    return l == -INFINITY ? INFINITY : l * 1.4426950408889634f;
                }
                return qr < lq ? p.delta[stat_base + qr] : 0.f;
            };
            // this group's first step of the item: local index t0 with (g + t0) & 1 == grp
            const int t0 = ((g & 1u) == static_cast<uint32_t>(grp)) ? 0 : 1;
            float nxt = fetch_stat(t0);
            prefetch_item(w + G);
            for (int t = 0; t < n; ++t, ++g) {
                if ((g & 1u) != static_cast<uint32_t>(grp)) continue;
                const uint32_t st = grp;
                const int i0 = (i_begin + t) * 64;
                // the buffer written here was last read two of this group's steps ago, and every thread has passed the previous
                // step's barrier since: no barrier is needed in front of the write
                float* gstat = gstat2 + (ks & 1u) * 128;
                ++ks;
                gstat[gt] = nxt;
                nxt = fetch_stat(t + 2);                      // in flight under this step's work
                bar_sync_group(grp);
                mbar_wait(&sdp_full[st], (g >> 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    uint32_t rs[32], rd[32];
                    tmem_ld_32x32(t_lane + TM_S + st * 64 + half * 32, rs);
                    tmem_ld_32x32(t_lane + TM_DP + st * 64 + half * 32, rd);
                    tmem_ld_wait();
                    if (half == 1) {                          // both halves are in registers: warp A may overwrite S^T/dP^T[st]
                        tc_fence_before();
                        mbar_arrive(&sdp_empty[st]);
                    }
                    // causal: query column cc (row i0 + cc) sees this key iff kj <= i0 + cc + off
                    int cmin = 0;
                    if (CAUSAL) cmin = kj - off - i0 - half * 32;
                    if (!key_ok) cmin = 32;
                    const float4* l4 = reinterpret_cast<const float4*>(gstat + half * 32);
                    const float4* d4 = reinterpret_cast<const float4*>(gstat + 64 + half * 32);
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
                    if (half == 0) mbar_wait(&ps_empty[st], ((g >> 1) & 1) ^ 1);   // dV / dK of global step g-2 have consumed this buffer
                    store_half_row(smem + OFF_PT + st * PS, row, half, pp);
                    store_half_row(smem + OFF_DST + st * PS, row, half, pd);
                }
                fence_proxy_async_smem();
                mbar_arrive(&ps_full[st]);
            }
            mbar_wait(acc_full, it & 1);                       // the item's last dV / dK MMAs have landed
            tc_fence_after();
            store_grad_row(t_lane + TM_DV, grp, 1.f, nullptr, nullptr, kj, dvp, ok, true);
            store_grad_row(t_lane + TM_DK, grp, p.scale, p.rope_cos, p.rope_sin, kj, dkp, ok, true);
            tc_fence_before();
            mbar_arrive(acc_empty);
            ++it;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ host
static BwdSched make_sched(int nt, int BH, int H, int grid) {
    BwdSched s;
    s.nt = nt; s.BH = BH; s.H = H;
    s.C = grid / nt > 0 ? grid / nt : 1;
    const int nchunks = (BH + s.C - 1) / s.C;
    s.total = nchunks * s.C * nt;
    auto magic = [](int d) { return ((1ull << 40) + static_cast<unsigned long long>(d) - 1) / static_cast<unsigned long long>(d); };
    s.m_per = magic(s.C * nt); s.m_C = magic(s.C); s.m_nt = magic(nt); s.m_H = magic(H);
    return s;
}

template <bool CAUSAL>
static int launch_bwd_tcp(const AttnBwdArgs& a, cudaStream_t stream) {
    auto kq = attn_bwd_dq_tcp_kernel<CAUSAL>;
    auto kkv = attn_bwd_dkv_tcp_kernel<CAUSAL>;
    static bool attr_set = false;
    if (!attr_set) {
        LHRS_CUDA(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, dqp::SMEM_BYTES));
        LHRS_CUDA(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, dkp::SMEM_BYTES));
        attr_set = true;
    }
    AttnBwdTcArgs pa;
    pa.a = a;
    CUtensorMap q128, do128, k64, v64, q64, do64, k128, v128;
    int rc, f;
    const bool ragged = a.seq_off != nullptr;       // one row space for the whole batch: a single batch entry of total_rows rows
    const int mB = ragged ? 1 : a.B, mSq = ragged ? (int)a.total_rows : a.Sq, mSkv = ragged ? (int)a.total_rows : a.Skv;
    if ((rc = make_tmap_bshd(&q128, &pa.q_hfirst, a.q, 128, mSq, a.H, mB, a.q_rs, a.q_hs, a.q_bs, 128))) return rc;
    if ((rc = make_tmap_bshd(&do128, &pa.o_hfirst, a.d_o, 128, mSq, a.H, mB, a.o_rs, a.o_hs, a.o_bs, 128))) return rc;
    if ((rc = make_tmap_bshd(&k64, &pa.k_hfirst, a.k, 128, mSkv, a.H, mB, a.k_rs, a.k_hs, a.k_bs, 64))) return rc;
    if ((rc = make_tmap_bshd(&v64, &pa.v_hfirst, a.v, 128, mSkv, a.H, mB, a.v_rs, a.v_hs, a.v_bs, 64))) return rc;
    if ((rc = make_tmap_bshd(&q64, &f, a.q, 128, mSq, a.H, mB, a.q_rs, a.q_hs, a.q_bs, 64))) return rc;
    if ((rc = make_tmap_bshd(&do64, &f, a.d_o, 128, mSq, a.H, mB, a.o_rs, a.o_hs, a.o_bs, 64))) return rc;
    if ((rc = make_tmap_bshd(&k128, &f, a.k, 128, mSkv, a.H, mB, a.k_rs, a.k_hs, a.k_bs, 128))) return rc;
    if ((rc = make_tmap_bshd(&v128, &f, a.v, 128, mSkv, a.H, mB, a.v_rs, a.v_hs, a.v_bs, 128))) return rc;
    const int BH = a.B * a.H;
    const int nqt = (a.Sq + 127) / 128, nkt = (a.Skv + 127) / 128;
    {
        const long long items = static_cast<long long>(BH) * nqt;
        const int grid = static_cast<int>(items < num_sms() ? items : num_sms());
        const BwdSched sc = make_sched(nqt, BH, a.H, grid);
        kq<<<grid, dqp::THREADS, dqp::SMEM_BYTES, stream>>>(q128, do128, k64, v64, pa, sc);
        LHRS_LAUNCH_CHECK("attn_bwd_dq_tcp_kernel");
    }
    {
        const long long items = static_cast<long long>(BH) * nkt;
        const int grid = static_cast<int>(items < num_sms() ? items : num_sms());
        const BwdSched sc = make_sched(nkt, BH, a.H, grid);
        kkv<<<grid, dkp::THREADS, dkp::SMEM_BYTES, stream>>>(q64, do64, k128, v128, pa, sc);
        LHRS_LAUNCH_CHECK("attn_bwd_dkv_tcp_kernel");
    }
    return LHRS_OK;
}

bool attention_bwd_tcp_ok(const AttnBwdArgs& a) {
    const long long items = static_cast<long long>(a.B) * a.H * ((std::max(a.Sq, a.Skv) + 127) / 128 + 1) + 4096;
    return a.Skv <= dqp::MAX_SKV && items < (1ll << 20) && a.H < (1 << 20);
}

#ifdef LHRS_ATTN_TRACE
extern "C" int lhrs_debug_attn_trace(long long* out, int n) {
    return cudaMemcpyFromSymbol(out, g_attn_trace, sizeof(long long) * (n < 4 * 8192 ? n : 4 * 8192)) == cudaSuccess ? 0 : 1;
}
#endif

int attention_bwd_tcp(const AttnBwdArgs& a, bool causal, cudaStream_t stream) {
    return causal ? launch_bwd_tcp<true>(a, stream) : launch_bwd_tcp<false>(a, stream);
}

}  // namespace lhrs
