// Dense bf16 contraction on the 5th-gen tensor cores (sm_100a):
//   D[M,N] = epilogue(alpha * A[M,K] · B[N,K]^T), fp32 accumulation in TMEM.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      MMA issuer    (one thread issues tcgen05.mma 128xBNx16, commits free the ring slots)
//   warp 2      TMEM allocator (2 accumulator stages of BN fp32 columns)
//   warps 4..7  epilogue      (tcgen05.ld -> registers -> fused epilogue -> global), overlaps the next tile's MMAs
//
// Every nn.Linear of the LHRS-Bot hot path goes through here (see include/lhrs_b200.h for the map to
// the reference call sites); the K-major/MN-major operand switches give the dX and dW backward forms.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "host_common.h"
#include "ptx.cuh"
#include "dropout.cuh"

namespace lhrs {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle atom

struct GemmArgs {
    int M, N, K;
    int num_b, seg_rows;
    int act;
    float alpha;
    const __nv_bfloat16* bias[3];
    const __nv_bfloat16* residual;
    long long ldr;
    void* D;
    long long ldd;
    int d_f32;
    const int* row_map;
    const float* rope_cos;
    const float* rope_sin;
    const int* positions;
    int rope_seq_len;
    __nv_bfloat16* pre_gate;
    __nv_bfloat16* pre_up;
    int kseg;    // MN-major B stacked along K: reduction length per segment (0 = single B)
    int kseg_nshift;  // segment s lands in output columns [s*nshift, s*nshift + its width): block-diagonal B (LoRA dT)
    int coalesce; // epilogue I/O through the per-warp smem transpose (1) or direct per-lane rows (0)
    int group_m; // row-blocks per rasterisation group (L2 reuse)
    int pf_dist; // k-blocks of L2 prefetch distance ahead of the TMA loads (0 = off)
    int splitk;  // >1: the K loop is split over `splitk` CTAs per tile; partial sums are added with fp32 atomics into D
    int ext_k;   // LoRA K-extension: columns of A2 per B segment (0 = none)
    int ext_kb;  // extra 64-wide k-blocks appended after the main K loop
    uint32_t drop_key;  // LINEAR: LoRA dropout mask applied to the accumulator (drop_t > 0), see dropout.cuh
    int drop_t;
};

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) owns a 256 x BN tile; each CTA
// stages its own 128 rows of A and HALF of B (BN/2 rows), the pair's single MMA reads both halves, which cuts the operand
// traffic per flop (L2 -> SM and smem writes) by a third.
template <int BN, int CG = 1>
struct GemmCfg {
    static constexpr int BN_CTA = BN / CG;   // B rows staged by one CTA
    static constexpr int STAGES = (BN_CTA == 256) ? 4 : 6;
    static constexpr uint32_t A_BYTES = BM * BK * 2;
    static constexpr uint32_t B_BYTES = BN_CTA * BK * 2;
    static constexpr uint32_t TMEM_COLS = 2 * BN;
    static constexpr uint32_t EPI_STAGE_BYTES = 32 * 80;   // per epilogue warp: 32 rows x 32 bf16, rows padded to 80 B
    static constexpr uint32_t SMEM_BYTES = STAGES * (A_BYTES + B_BYTES) + 256 /*barriers*/ + 4 * EPI_STAGE_BYTES + 1024 /*align slack*/;
};

__device__ __forceinline__ float act_apply(float v, int act) {
    if (act == LHRS_ACT_GELU_ERF) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    if (act == LHRS_ACT_QUICK_GELU) return v / (1.0f + __expf(-1.702f * v));
    return v;
}

__device__ __forceinline__ void load_bf16x32(const __nv_bfloat16* p, float (&out)[32]) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 u = __ldg(q + i);
        out[i * 8 + 0] = bf16_lo(u.x); out[i * 8 + 1] = bf16_hi(u.x);
        out[i * 8 + 2] = bf16_lo(u.y); out[i * 8 + 3] = bf16_hi(u.y);
        out[i * 8 + 4] = bf16_lo(u.z); out[i * 8 + 5] = bf16_hi(u.z);
        out[i * 8 + 6] = bf16_lo(u.w); out[i * 8 + 7] = bf16_hi(u.w);
    }
}
__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* p, const float (&v)[32]) {
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16(v[i * 8 + 0], v[i * 8 + 1]);
        u.y = pack_bf16(v[i * 8 + 2], v[i * 8 + 3]);
        u.z = pack_bf16(v[i * 8 + 4], v[i * 8 + 5]);
        u.w = pack_bf16(v[i * 8 + 6], v[i * 8 + 7]);
        q[i] = u;
    }
}
__device__ __forceinline__ void store_f32x32(float* p, const float (&v)[32]) {
    float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
}

// ---- coalesced epilogue I/O.  A warp owns 32 rows x 32 columns with lane = row (the TMEM read layout).  Storing that
// directly makes every instruction touch 32 different rows (half-used 32-byte sectors).  Instead the tile is transposed
// through a small per-warp smem slab so that each instruction moves 8 rows x 64 contiguous bytes.
__device__ __forceinline__ void tile_store(uint8_t* slab, const float (&v)[32], __nv_bfloat16* g_row0, long long ld, int nvalid, int lane) {
    uint4* mine = reinterpret_cast<uint4*>(slab + lane * 80);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16(v[i * 8 + 0], v[i * 8 + 1]); u.y = pack_bf16(v[i * 8 + 2], v[i * 8 + 3]);
        u.z = pack_bf16(v[i * 8 + 4], v[i * 8 + 5]); u.w = pack_bf16(v[i * 8 + 6], v[i * 8 + 7]);
        mine[i] = u;
    }
    __syncwarp();
    const int cc = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rr = (lane >> 2) + 8 * i;
        if (rr < nvalid) *reinterpret_cast<uint4*>(g_row0 + rr * ld + cc * 8) = *reinterpret_cast<const uint4*>(slab + rr * 80 + cc * 16);
    }
    __syncwarp();
}
__device__ __forceinline__ void tile_load(uint8_t* slab, float (&out)[32], const __nv_bfloat16* g_row0, long long ld, int nvalid, int lane) {
    const int cc = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rr = (lane >> 2) + 8 * i;
        uint4 u = make_uint4(0, 0, 0, 0);
        if (rr < nvalid) u = __ldg(reinterpret_cast<const uint4*>(g_row0 + rr * ld + cc * 8));
        *reinterpret_cast<uint4*>(slab + rr * 80 + cc * 16) = u;
    }
    __syncwarp();
    const uint4* mine = reinterpret_cast<const uint4*>(slab + lane * 80);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint4 u = mine[i];
        out[i * 8 + 0] = bf16_lo(u.x); out[i * 8 + 1] = bf16_hi(u.x); out[i * 8 + 2] = bf16_lo(u.y); out[i * 8 + 3] = bf16_hi(u.y);
        out[i * 8 + 4] = bf16_lo(u.z); out[i * 8 + 5] = bf16_hi(u.z); out[i * 8 + 6] = bf16_lo(u.w); out[i * 8 + 7] = bf16_hi(u.w);
    }
    __syncwarp();
}

// tile index -> (m_blk, n_blk): groups of GROUP_M row-blocks are walked column by column so that the ~148
// tiles in flight share a small set of A row-blocks and B column-blocks in L2.
__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int& m_blk, int& n_blk, int GROUP_M) {
    const int group_size = GROUP_M * num_n;
    const int g = tile / group_size;
    const int first_m = g * GROUP_M;
    const int gm = min(GROUP_M, num_m - first_m);
    const int r = tile - g * group_size;
    m_blk = first_m + r % gm;
    n_blk = r / gm;
}

template <int BN, int KIND, bool A_MN, bool B_MN, int CG>
__global__ void __launch_bounds__(256, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
                 const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmE0,
                 const __grid_constant__ CUtensorMap tmE1, const __grid_constant__ CUtensorMap tmE2,
                 const GemmArgs args) {
    using Cfg = GemmCfg<BN, CG>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int BN_CTA = Cfg::BN_CTA;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;   // 0 = leader of the pair (issues the MMAs)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (Cfg::A_BYTES + Cfg::B_BYTES));
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tmem_full_bar = bars + 2 * STAGES;
    uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    uint8_t* epi_stage_base = reinterpret_cast<uint8_t*>(bars) + 256;   // 4 x EPI_STAGE_BYTES, one slab per epilogue warp

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int num_m = (args.M + BM * CG - 1) / (BM * CG);   // row-blocks of 128*CG (one per CTA or per CTA pair)
    const int num_n = (args.N + BN - 1) / BN;
    const int splitk = args.splitk;                         // 1 unless the host asked for a K split (skinny problems)
    const int num_tiles = num_m * num_n * splitk;
    const int num_kb_main = (args.K + BK - 1) / BK;
    const int num_kb = num_kb_main + args.ext_kb;
    const int kb_per = (num_kb + splitk - 1) / splitk;     // k-blocks per split (ext_kb == 0 when splitk > 1)
    const int ext_span = (KIND == LHRS_EPI_SWIGLU) ? 2 * args.ext_k : args.ext_k;   // non-zero k columns of the K-extension

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB0);
        if (args.num_b > 1) tma_prefetch_desc(&tmB1);
        if (args.num_b > 2) tma_prefetch_desc(&tmB2);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], CG);      // pair: both CTAs' producers arrive on the leader's barrier
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full_bar[s], 1);
            mbar_init(&tmem_empty_bar[s], 128 * CG);   // pair: both CTAs' epilogue threads arrive on the leader's barrier
        }
        fence_barrier_init();
    }
    if constexpr (CG == 2) cluster_sync_all();   // peer barriers must exist before anything targets them
    if (warp == 2) {
        if constexpr (CG == 2) { tmem_alloc_2sm(tmem_ptr_smem, Cfg::TMEM_COLS); tmem_relinquish_2sm(); }
        else { tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS); tmem_relinquish(); }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================================================================== TMA producer
        {
            // whole warp in the (uniform) control flow, one elected lane issues: keeps the TMA operands in uniform registers
            const bool issuer = elect_one();
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x / CG; tile < num_tiles; tile += gridDim.x / CG) {
                int m_blk, n_blk;
                tile_coords(tile / splitk, num_m, num_n, m_blk, n_blk, args.group_m);
                const int kb_begin = (tile % splitk) * kb_per, kb_end = min(kb_begin + kb_per, num_kb);
#ifdef LHRS_GEMM_DEBUG
                if (args.pf_dist == -3) m_blk = n_blk = 0;   // every cluster streams the same operand tiles (pure L2 hits)
#endif
                const int m0 = (m_blk * CG + static_cast<int>(rank)) * BM;
                const int n0 = n_blk * BN;
                const int nb0 = n0 + static_cast<int>(rank) * BN_CTA;   // first B row staged by this CTA
                // TMA issue: plain for CG=1; for a pair the completion is signalled on the leader's full barrier.
                // In prefetch mode the same coordinates are only pulled into L2 (no smem destination, no barrier).
                bool pf_mode = false;
                auto LD = [&](void* dst, const CUtensorMap* tm, int c0, int c1) {
                    if (pf_mode) { tma_prefetch_2d(tm, c0, c1); return; }
                    if constexpr (CG == 2) tma_load_2d_2sm(dst, tm, &full_bar[stage], c0, c1);
                    else tma_load_2d(dst, tm, &full_bar[stage], c0, c1);
                };
                auto issue = [&](int kb) {
                    uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
                    uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
                    if (kb >= num_kb_main) {
                        // ---- LoRA K-extension: A2 = T (x·A^T, scaled), B2 = lora_B of this tile's segment(s).
                        // Out-of-range coordinates (negative or >= extent) are zero-filled by TMA, which is what
                        // confines each segment to its own ext_k columns of T.
                        const int e0 = (kb - num_kb_main) * BK;
                        if constexpr (KIND == LHRS_EPI_SWIGLU) {
                            const int h0 = n_blk * (BN / 2);
                            LD(sa, &tmA2, e0, m0);  // T = [T_gate | T_up]
                            if constexpr (CG == 2) {   // rank 0 stages the gate half, rank 1 the up half
                                if (rank == 0) LD(sb, &tmE0, e0, h0); else LD(sb, &tmE1, e0 - args.ext_k, h0);
                            } else {
                                LD(sb, &tmE0, e0, h0);
                                LD(sb + (BN / 2) * BK * 2, &tmE1, e0 - args.ext_k, h0);
                            }
                        } else if constexpr (B_MN) {
                            // dX form: A2 = dT [M, ext_k] (K-major), B2 = [ext_k, N] read MN-major in place (LoRA A stack);
                            // rows of B2 beyond ext_k are zero-filled by TMA
                            LD(sa, &tmA2, e0, m0);
#pragma unroll
                            for (int a = 0; a < BN_CTA / 64; ++a) LD(sb + a * 8192, &tmE0, nb0 + a * 64, e0);
                        } else {
                            const int seg = (args.num_b > 1) ? (n0 / args.seg_rows) : 0;
                            const int r0 = nb0 - seg * args.seg_rows;
                            const CUtensorMap* tm = (seg == 0) ? &tmE0 : (seg == 1 ? &tmE1 : &tmE2);
                            LD(sa, &tmA2, seg * args.ext_k + e0, m0);
                            LD(sb, tm, e0, r0);
                        }
                        return;
                    }
                    const int k0 = kb * BK;
#ifdef LHRS_GEMM_DEBUG   // timing experiments only (results are garbage): -1 skip B loads, -2 skip A loads
                    if (args.pf_dist == -2) goto load_b;
#endif
                    if constexpr (!A_MN) {
                        LD(sa, &tmA, k0, m0);  // box {64 k, 128 m}
                    } else {
#pragma unroll
                        for (int a = 0; a < BM / 64; ++a)  // box {64 m, 64 k} per 128B atom column
                            LD(sa + a * 8192, &tmA, m0 + a * 64, k0);
                    }
#ifdef LHRS_GEMM_DEBUG
                load_b:
                    if (args.pf_dist == -1) return;
#endif
                    if constexpr (!B_MN) {
                        if constexpr (KIND == LHRS_EPI_SWIGLU) {
                            // tile = [BN/2 gate rows | BN/2 up rows] of the same hidden units
                            const int h0 = n_blk * (BN / 2);
                            if constexpr (CG == 2) {   // rank 0 stages the gate rows, rank 1 the up rows
                                LD(sb, rank == 0 ? &tmB0 : &tmB1, k0, h0);
                            } else {
                                LD(sb, &tmB0, k0, h0);
                                LD(sb + (BN / 2) * BK * 2, &tmB1, k0, h0);
                            }
                        } else {
                            const int seg = (args.num_b > 1) ? (n0 / args.seg_rows) : 0;
                            const int r0 = nb0 - seg * args.seg_rows;
                            const CUtensorMap* tm = (seg == 0) ? &tmB0 : (seg == 1 ? &tmB1 : &tmB2);
                            LD(sb, tm, k0, r0);  // box {64 k, BN_CTA n}
                        }
                    } else {
                        // MN-major B.  With kseg > 0 the B segments are stacked along K: the dX of a concatenated output,
                        // e.g. [dq|dk|dv] · [Wq;Wk;Wv], runs as ONE contraction with fp32 accumulation across the segments.
                        const int seg = (args.kseg > 0) ? (k0 / args.kseg) : 0;
                        const int kk = k0 - seg * args.kseg;
                        const CUtensorMap* tm = (seg == 0) ? &tmB0 : (seg == 1 ? &tmB1 : &tmB2);
#pragma unroll
                        for (int a = 0; a < BN_CTA / 64; ++a)
                            LD(sb + a * 8192, tm, nb0 + a * 64 - seg * args.kseg_nshift, kk);
                    }
                };
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    if (issuer && args.pf_dist > 0 && kb + args.pf_dist < kb_end) {   // warm L2 a few k-blocks ahead of the smem ring
                        pf_mode = true;
                        issue(kb + args.pf_dist);
                        pf_mode = false;
                    }
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (issuer) {
#ifdef LHRS_GEMM_DEBUG
                        if (args.pf_dist < 0) {
                            const uint32_t by = (args.pf_dist == -2 ? 0 : Cfg::A_BYTES) + (args.pf_dist == -1 ? 0 : Cfg::B_BYTES);
                            if (CG == 2 && rank != 0) mbar_arrive_cluster(&full_bar[stage], 0);
                            else mbar_arrive_expect_tx(&full_bar[stage], CG * by);
                            issue(kb);
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                            continue;
                        }
#endif
                        if constexpr (CG == 2) {
                            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES));
                            else mbar_arrive_cluster(&full_bar[stage], 0);
                        } else {
                            mbar_arrive_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES);
                        }
                        issue(kb);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (pair: the leader CTA only)
        if (rank == 0) {
            // The whole warp walks the (warp-uniform) control flow; one elected lane issues.  The loop body is kept to the
            // minimum instruction count: with 128-cycle MMAs a k-block is only 512 tensor cycles, and a single warp
            // issuing ~100 dependent uniform-datapath instructions per k-block was itself the bound (ncu: pipe 76 %).
            constexpr uint32_t idesc = make_idesc_bf16(BM * CG, BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
            // descriptor words: lo = start>>4 | LBO>>4 << 16 ; hi = SBO>>4 | version 1 << 14 | SW128 << 29 (same for A and B)
            constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (UMMA_LAYOUT_SW128 << 29);
            constexpr uint32_t a_kstep = (A_MN ? 2048u : 32u) >> 4, b_kstep = (B_MN ? 2048u : 32u) >> 4;
            const uint32_t a_lo0 = ((smem_u32(smem_a) >> 4) & 0x3FFFu) | (A_MN ? ((8192u >> 4) << 16) : 0u);
            const uint32_t b_lo0 = ((smem_u32(smem_b) >> 4) & 0x3FFFu) | (B_MN ? ((8192u >> 4) << 16) : 0u);
            const bool issuer = elect_one();
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x / CG; tile < num_tiles; tile += gridDim.x / CG, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * BN;
                const int kb_begin = (tile % splitk) * kb_per, kb_end = min(kb_begin + kb_per, num_kb);
                uint32_t acc = 0;
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (issuer) {
                        const uint32_t a_lo = a_lo0 + stage * (Cfg::A_BYTES >> 4);
                        const uint32_t b_lo = b_lo0 + stage * (Cfg::B_BYTES >> 4);
                        // a LoRA K-extension block only carries ext_span (16 / 32 / 48 at r = 16) non-zero k columns: the rest of the
                        // 64-wide box is TMA zero fill, so its K=16 steps would add exact zeros
                        const int nk = (kb >= num_kb_main) ? min(BK / 16, (ext_span - (kb - num_kb_main) * BK + 15) / 16) : BK / 16;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            // K-major SW128: +32 B per K=16 step inside the 128 B row.  MN-major SW128: 64-wide MN atoms
                            // 8 KB apart (LBO), 8-row K groups 1 KB apart (SBO); +2 KB per K=16.
                            if (k < nk) {
                                if constexpr (CG == 2) umma_bf16_2sm_w(tmem_d, a_lo + k * a_kstep, b_lo + k * b_kstep, desc_hi, idesc, k == 0 ? acc : 1u);
                                else umma_bf16_w(tmem_d, a_lo + k * a_kstep, b_lo + k * b_kstep, desc_hi, idesc, k == 0 ? acc : 1u);
                            }
                        }
                        // slot reusable once these MMAs have read it (pair: released in both CTAs)
                        if constexpr (CG == 2) umma_commit_2sm(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
                        if (kb == kb_end - 1) {
                            if constexpr (CG == 2) umma_commit_2sm(&tmem_full_bar[as]); else umma_commit(&tmem_full_bar[as]);
                        }
                    }
                    acc = 1;
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================================================================== epilogue (warps 4..7)
        const int q = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may read
        int it = 0;
        for (int tile = blockIdx.x / CG; tile < num_tiles; tile += gridDim.x / CG, ++it) {
            int m_blk, n_blk;
            tile_coords(tile / splitk, num_m, num_n, m_blk, n_blk, args.group_m);
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            mbar_wait(&tmem_full_bar[as], aphase);
            tc_fence_after();

            const int row = (m_blk * CG + static_cast<int>(rank)) * BM + q * 32 + lane;
            const bool row_ok = row < args.M;
            long long drow = row;
            if (row_ok && args.row_map != nullptr) drow = args.row_map[row];
            const bool store_ok = row_ok && drow >= 0;
            // coalesced path (all epilogues unless rows are scattered or the output is fp32)
            const bool coal = (args.row_map == nullptr) && !args.d_f32 && args.coalesce;
            const long long row_w0 = static_cast<long long>(m_blk * CG + static_cast<int>(rank)) * BM + q * 32;   // first row of this warp
            const int nvalid = static_cast<int>(min(32LL, max(0LL, static_cast<long long>(args.M) - row_w0)));
            uint8_t* slab = epi_stage_base + q * Cfg::EPI_STAGE_BYTES;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
            const int n_tile0 = n_blk * BN;

            if constexpr (KIND == LHRS_EPI_LINEAR) {
                // per-segment bias: a tile never straddles two B segments (seg_rows % BN == 0)
                const int bseg = (args.num_b > 1) ? (n_tile0 / args.seg_rows) : 0;
                const __nv_bfloat16* bias_seg = (bseg == 0) ? args.bias[0] : (bseg == 1 ? args.bias[1] : args.bias[2]);
                const __nv_bfloat16* bias = (bias_seg != nullptr) ? bias_seg - bseg * args.seg_rows : nullptr;  // index by global n
                if (args.pre_up != nullptr && coal) {
                    // SwiGLU backward fused into the dX GEMM of down_proj (acc = d_act; D = d_gu = [d_gate | d_up], d_act is never
                    // stored).  The stashed pre-activations g, u are 2 x 16 KB per warp and tile: their loads for chunk c+1 are
                    // issued BEFORE chunk c is processed, so the ~1 us global latency sits under a chunk of math and stores.
                    // (Unpipelined — load, wait, compute, store per chunk — this epilogue took longer than the K = 4096 mainloop
                    // and the fused form lost 2 ms per step to the separate swiglu_bwd pass, gpurun_out/s2b_sft_fuse.json.)
                    const int cc = lane & 3;
                    uint4 gq[4], uq[4];
                    auto issue = [&](int n0) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int rr = (lane >> 2) + 8 * i;
                            gq[i] = make_uint4(0, 0, 0, 0); uq[i] = make_uint4(0, 0, 0, 0);
                            if (rr < nvalid && n0 + 32 <= args.N) {
                                gq[i] = __ldg(reinterpret_cast<const uint4*>(args.pre_gate + (row_w0 + rr) * args.N + n0 + cc * 8));
                                uq[i] = __ldg(reinterpret_cast<const uint4*>(args.pre_up + (row_w0 + rr) * args.N + n0 + cc * 8));
                            }
                        }
                    };
                    auto transpose_in = [&](const uint4 (&q4)[4], float (&out)[32]) {   // 8 rows x 64 B pieces -> this lane's row
#pragma unroll
                        for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(slab + ((lane >> 2) + 8 * i) * 80 + cc * 16) = q4[i];
                        __syncwarp();
                        const uint4* mine = reinterpret_cast<const uint4*>(slab + lane * 80);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint4 u4 = mine[i];
                            out[i * 8 + 0] = bf16_lo(u4.x); out[i * 8 + 1] = bf16_hi(u4.x); out[i * 8 + 2] = bf16_lo(u4.y); out[i * 8 + 3] = bf16_hi(u4.y);
                            out[i * 8 + 4] = bf16_lo(u4.z); out[i * 8 + 5] = bf16_hi(u4.z); out[i * 8 + 6] = bf16_lo(u4.w); out[i * 8 + 7] = bf16_hi(u4.w);
                        }
                        __syncwarp();
                    };
                    issue(n_tile0);
#pragma unroll 1
                    for (int c = 0; c < BN / 32; ++c) {
                        const int n0 = n_tile0 + c * 32;
                        if (n0 + 32 > args.N) break;
                        float g[32], u[32];
                        transpose_in(gq, g);
                        transpose_in(uq, u);
                        if (c + 1 < BN / 32) issue(n0 + 32);      // in flight under this chunk's TMEM read, math and stores
                        uint32_t r[32];
                        tmem_ld_32x32(taddr + c * 32, r);
                        tmem_ld_wait();
                        float dg[32], du[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float da = bf16_round(__uint_as_float(r[j]) * args.alpha);
                            const float sg = 1.f / (1.f + __expf(-g[j]));
                            dg[j] = da * u[j] * sg * (1.f + g[j] * (1.f - sg));
                            du[j] = da * g[j] * sg;
                        }
                        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(args.D) + row_w0 * args.ldd + n0;
                        tile_store(slab, dg, d, args.ldd, nvalid, lane);
                        tile_store(slab, du, d + args.N, args.ldd, nvalid, lane);
                    }
                } else
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    const int n0 = n_tile0 + c * 32;
                    if (n0 >= args.N) break;
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + c * 32, r);
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * args.alpha;
                    if (args.pre_up != nullptr) {
                        // SwiGLU backward fused into the dX GEMM of down_proj: acc = d_act; with the stashed pre-activations
                        // (pre_gate = g, pre_up = u) write [d_gate | d_up] into D = d_gu [M, 2N] and never store d_act
                        if (n0 + 32 <= args.N) {
                            float g[32], u[32], dg[32], du[32];
                            if (coal) {
                                tile_load(slab, g, args.pre_gate + row_w0 * args.N + n0, args.N, nvalid, lane);
                                tile_load(slab, u, args.pre_up + row_w0 * args.N + n0, args.N, nvalid, lane);
                            } else if (row_ok) {
                                load_bf16x32(args.pre_gate + static_cast<long long>(row) * args.N + n0, g);
                                load_bf16x32(args.pre_up + static_cast<long long>(row) * args.N + n0, u);
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float da = bf16_round(v[j]);
                                const float sg = 1.f / (1.f + __expf(-g[j]));
                                dg[j] = da * u[j] * sg * (1.f + g[j] * (1.f - sg));
                                du[j] = da * g[j] * sg;
                            }
                            if (coal) {
                                __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(args.D) + row_w0 * args.ldd + n0;
                                tile_store(slab, dg, d, args.ldd, nvalid, lane);
                                tile_store(slab, du, d + args.N, args.ldd, nvalid, lane);
                            } else if (store_ok) {
                                __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(args.D) + drow * args.ldd + n0;
                                store_bf16x32(d, dg);
                                store_bf16x32(d + args.N, du);
                            }
                        }
                        continue;
                    }
                    if (splitk > 1) {   // partial sum of a K split: fp32 atomics into the (zero-initialised) fp32 output
                        if (store_ok) {
                            float* d = reinterpret_cast<float*>(args.D) + drow * args.ldd + n0;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n0 + j < args.N) atomicAdd(d + j, v[j]);
                        }
                        continue;
                    }
                    const bool full_chunk = (n0 + 32 <= args.N);
                    if (full_chunk) {
                        if (bias != nullptr) {
                            float b[32];
                            load_bf16x32(bias + n0, b);
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] += b[j];
                        }
                        if (args.pre_gate != nullptr && args.pre_up == nullptr) {   // keep the pre-activation for the backward pass
                            if (args.coalesce) tile_store(slab, v, args.pre_gate + row_w0 * args.N + n0, args.N, nvalid, lane);
                            else if (row_ok) store_bf16x32(args.pre_gate + static_cast<long long>(row) * args.N + n0, v);
                        }
                        if (args.act != LHRS_ACT_NONE) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = act_apply(bf16_round(v[j]), args.act);
                        }
                        if (args.drop_t > 0) {   // LoRA dropout on the dX correction dT·A: one mask word per 2 x 2 block
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const uint32_t word = drop_word(args.drop_key, static_cast<uint32_t>(row), static_cast<uint32_t>(n0 + j),
                                                                static_cast<uint32_t>(args.N));
                                if (!drop_keep(word, static_cast<uint32_t>(row), 0u, args.drop_t)) v[j] = 0.f;
                                if (!drop_keep(word, static_cast<uint32_t>(row), 1u, args.drop_t)) v[j + 1] = 0.f;
                            }
                        }
                        if (args.residual != nullptr) {
                            float rr[32];
                            if (args.coalesce) tile_load(slab, rr, args.residual + row_w0 * args.ldr + n0, args.ldr, nvalid, lane);
                            else if (row_ok) load_bf16x32(args.residual + static_cast<long long>(row) * args.ldr + n0, rr);
                            if (args.coalesce || row_ok) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]) + rr[j];
                            }
                        }
                        if (coal) {
                            tile_store(slab, v, reinterpret_cast<__nv_bfloat16*>(args.D) + row_w0 * args.ldd + n0, args.ldd, nvalid, lane);
                        } else if (store_ok) {
                            if (args.d_f32)
                                store_f32x32(reinterpret_cast<float*>(args.D) + drow * args.ldd + n0, v);
                            else
                                store_bf16x32(reinterpret_cast<__nv_bfloat16*>(args.D) + drow * args.ldd + n0, v);
                        }
                    } else {
                        // ragged N tail (e.g. LoRA rank): scalar, predicated
                        for (int j = 0; j < 32; ++j) {
                            const int n = n0 + j;
                            if (n < args.N && store_ok) {
                                float x = v[j];
                                if (bias != nullptr) x += __bfloat162float(bias[n]);
                                if (args.act != LHRS_ACT_NONE) x = act_apply(bf16_round(x), args.act);
                                if (args.residual != nullptr)
                                    x = bf16_round(x) + __bfloat162float(args.residual[static_cast<long long>(row) * args.ldr + n]);
                                if (args.d_f32)
                                    reinterpret_cast<float*>(args.D)[drow * args.ldd + n] = x;
                                else
                                    reinterpret_cast<__nv_bfloat16*>(args.D)[drow * args.ldd + n] = __float2bfloat16_rn(x);
                            }
                        }
                    }
                }
            } else if constexpr (KIND == LHRS_EPI_SWIGLU) {
                // accumulator columns [0, BN/2) = gate, [BN/2, BN) = up, same hidden units
                constexpr int HALF = BN / 2;
                const int h_tile0 = n_blk * HALF;
                const int n_hidden = args.N / 2;
#pragma unroll 1
                for (int c = 0; c < HALF / 32; ++c) {
                    const int h0 = h_tile0 + c * 32;
                    if (h0 >= n_hidden) break;
                    uint32_t rg[32], ru[32];
                    tmem_ld_32x32(taddr + c * 32, rg);
                    tmem_ld_32x32(taddr + HALF + c * 32, ru);
                    tmem_ld_wait();
                    float g[32], u[32], o[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        g[j] = bf16_round(__uint_as_float(rg[j]) * args.alpha);
                        u[j] = bf16_round(__uint_as_float(ru[j]) * args.alpha);
                        const float s = bf16_round(g[j] / (1.0f + __expf(-g[j])));
                        o[j] = s * u[j];
                    }
                    if (coal) {
                        tile_store(slab, o, reinterpret_cast<__nv_bfloat16*>(args.D) + row_w0 * args.ldd + h0, args.ldd, nvalid, lane);
                        if (args.pre_gate != nullptr) {
                            tile_store(slab, g, args.pre_gate + row_w0 * n_hidden + h0, n_hidden, nvalid, lane);
                            tile_store(slab, u, args.pre_up + row_w0 * n_hidden + h0, n_hidden, nvalid, lane);
                        }
                    } else if (store_ok) {
                        store_bf16x32(reinterpret_cast<__nv_bfloat16*>(args.D) + drow * args.ldd + h0, o);
                        if (args.pre_gate != nullptr) {
                            store_bf16x32(args.pre_gate + static_cast<long long>(row) * n_hidden + h0, g);
                            store_bf16x32(args.pre_up + static_cast<long long>(row) * n_hidden + h0, u);
                        }
                    }
                }
            } else {  // LHRS_EPI_ROPE: columns are [q | k | v], each seg_rows wide, heads of 128
                int pos = 0;
                if (row_ok) pos = (args.positions != nullptr) ? args.positions[row] : (row % args.rope_seq_len);
                const float* cs = args.rope_cos + static_cast<long long>(pos) * 64;
                const float* sn = args.rope_sin + static_cast<long long>(pos) * 64;
#pragma unroll 1
                for (int h = 0; h < BN / 128; ++h) {
                    const int nh0 = n_tile0 + h * 128;
                    if (nh0 >= args.N) break;
                    const bool rotate = (nh0 / args.seg_rows) < 2;
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        uint32_t r1[32], r2[32];
                        tmem_ld_32x32(taddr + h * 128 + c * 32, r1);
                        tmem_ld_32x32(taddr + h * 128 + 64 + c * 32, r2);
                        tmem_ld_wait();
                        float x1[32], x2[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            x1[j] = bf16_round(__uint_as_float(r1[j]) * args.alpha);
                            x2[j] = bf16_round(__uint_as_float(r2[j]) * args.alpha);
                        }
                        if (rotate && row_ok) {
#pragma unroll
                            for (int j4 = 0; j4 < 8; ++j4) {
                                const float4 c4 = __ldg(reinterpret_cast<const float4*>(cs + c * 32) + j4);
                                const float4 s4 = __ldg(reinterpret_cast<const float4*>(sn + c * 32) + j4);
                                const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
                                const float ss[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int j = j4 * 4 + e;
                                    const float a = x1[j], b = x2[j];
                                    // HF rotate_half: out = x*cos + cat(-x2, x1)*sin, each product rounded to bf16
                                    x1[j] = bf16_round(a * cc[e]) - bf16_round(b * ss[e]);
                                    x2[j] = bf16_round(b * cc[e]) + bf16_round(a * ss[e]);
                                }
                            }
                        }
                        if (coal) {
                            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(args.D) + row_w0 * args.ldd + nh0;
                            tile_store(slab, x1, d + c * 32, args.ldd, nvalid, lane);
                            tile_store(slab, x2, d + 64 + c * 32, args.ldd, nvalid, lane);
                        } else if (store_ok) {
                            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(args.D) + drow * args.ldd + nh0;
                            store_bf16x32(d + c * 32, x1);
                            store_bf16x32(d + 64 + c * 32, x2);
                        }
                    }
                }
            }
            tc_fence_before();
            if constexpr (CG == 2) mbar_arrive_cluster(&tmem_empty_bar[as], 0); else mbar_arrive(&tmem_empty_bar[as]);
        }
    }

    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if constexpr (CG == 2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D bf16 tensor map over a row-major [outer, inner] matrix with `ld` elements between rows, 128B swizzle.
static int make_tmap(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                     uint32_t box_inner, uint32_t box_outer) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return LHRS_ERR_CUDA;
    }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu outer=%llu ld=%llu box=%ux%u", (int)r, ptr,
                  (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner, box_outer);
        return LHRS_ERR_CUDA;
    }
    return LHRS_OK;
}

template <int BN, int KIND, bool A_MN, bool B_MN, int CG = 1>
static int launch(const CUtensorMap& tA, const CUtensorMap (&tB)[3], const CUtensorMap& tA2, const CUtensorMap (&tE)[3],
                  const GemmArgs& a, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, CG>;
    auto kern = gemm_bf16_kernel<BN, KIND, A_MN, B_MN, CG>;
    static bool attr_set = false;
    if (!attr_set) {
        LHRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int num_tiles = ((a.M + BM * CG - 1) / (BM * CG)) * ((a.N + BN - 1) / BN) * a.splitk;   // per CTA (CG=1) or CTA pair
    const int max_units = num_sms() / CG;
    const int grid = (num_tiles < max_units ? num_tiles : max_units) * CG;
    const bool prof = prof_on();
    if (prof) {
        const double kk = (double)a.K + (double)a.ext_k;  // algorithmic reduction length (LoRA rank, not its 64-padding)
        prof_begin(CG == 2 ? PROF_GEMM : PROF_GEMM_SMALL, 2.0 * a.M * (double)a.N * kk, 2.0 * ((double)a.M * a.K + (double)a.N * a.K + (double)a.M * a.N), stream);
    }
    if constexpr (CG == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        LHRS_CUDA(cudaLaunchKernelEx(&cfg, kern, tA, tB[0], tB[1], tB[2], tA2, tE[0], tE[1], tE[2], a));
    } else {
        kern<<<grid, 256, Cfg::SMEM_BYTES, stream>>>(tA, tB[0], tB[1], tB[2], tA2, tE[0], tE[1], tE[2], a);
    }
    if (prof) prof_end(stream);
    LHRS_LAUNCH_CHECK("gemm_bf16_kernel");
    return LHRS_OK;
}

}  // namespace lhrs

using namespace lhrs;

extern "C" int lhrs_gemm_bf16(const LhrsGemm* g, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    LHRS_CHECK_ARG(g != nullptr, "lhrs_gemm_bf16: null descriptor");
    LHRS_CHECK_ARG(g->M > 0 && g->N > 0 && g->K > 0, "lhrs_gemm_bf16: empty problem M=%d N=%d K=%d", g->M, g->N, g->K);
    LHRS_CHECK_ARG(g->A != nullptr && g->B[0] != nullptr && g->D != nullptr, "lhrs_gemm_bf16: null operand");
    LHRS_CHECK_ARG(g->num_b >= 1 && g->num_b <= 3, "lhrs_gemm_bf16: num_b=%d", g->num_b);
    LHRS_CHECK_ARG((g->lda % 8) == 0 && (g->ldb % 8) == 0, "lhrs_gemm_bf16: lda/ldb must be multiples of 8 elements (TMA 16B stride)");
    LHRS_CHECK_ARG((reinterpret_cast<uintptr_t>(g->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(g->B[0]) & 15) == 0,
                   "lhrs_gemm_bf16: operands must be 16-byte aligned");
    LHRS_CHECK_ARG((g->ldd % 8) == 0 && (reinterpret_cast<uintptr_t>(g->D) & 15) == 0, "lhrs_gemm_bf16: D alignment");
    if (g->residual) LHRS_CHECK_ARG((g->ldr % 8) == 0 && (reinterpret_cast<uintptr_t>(g->residual) & 15) == 0, "lhrs_gemm_bf16: residual alignment");

    const int kind = g->epilogue;
    const bool a_mn = g->a_mn_major != 0, b_mn = g->b_mn_major != 0;
    int bn = 256;
    // LHRS_GEMM_CG=1/2 forces single-CTA / CTA-pair tiles (read per call: the parity tests flip it between cases so that the
    // pair kernel and its K-extensions are also exercised on small problems)
    const char* cg_env = getenv("LHRS_GEMM_CG");
    const int forced = cg_env ? atoi(cg_env) : 0;
    if (kind == LHRS_EPI_LINEAR) {
        if (g->N <= 128 || (g->N % 256) != 0) bn = 128;
        // small problems: more, smaller tiles fill the 148 SMs better
        const long long tiles256 = (long long)((g->M + BM - 1) / BM) * ((g->N + 255) / 256);
        if (tiles256 < num_sms() && g->N > 128 && !(forced == 2 && (g->N % 256) == 0)) bn = 128;
        if (g->num_b > 1 && b_mn) {
            // segments stacked along K (each [K/num_b, ldb>=N]); seg_rows is ignored
            LHRS_CHECK_ARG((g->K % g->num_b) == 0 && ((g->K / g->num_b) % BK) == 0 && g->B[1] != nullptr && (g->num_b < 3 || g->B[2] != nullptr),
                           "lhrs_gemm_bf16: K=%d must split into %d segments of a multiple of %d", g->K, g->num_b, BK);
        } else if (g->num_b > 1) {
            LHRS_CHECK_ARG(g->seg_rows > 0 && g->seg_rows * g->num_b == g->N && (g->seg_rows % bn) == 0,
                           "lhrs_gemm_bf16: seg_rows=%d must tile N=%d by %d", g->seg_rows, g->N, bn);
        }
    } else if (kind == LHRS_EPI_SWIGLU) {
        LHRS_CHECK_ARG(g->num_b == 2 && !a_mn && !b_mn && (g->N % 256) == 0 && g->B[1] != nullptr,
                       "lhrs_gemm_bf16: SWIGLU needs 2 K-major segments and N %% 256 == 0 (N=%d)", g->N);
    } else if (kind == LHRS_EPI_ROPE) {
        LHRS_CHECK_ARG(g->num_b == 3 && !a_mn && !b_mn && g->seg_rows > 0 && (g->seg_rows % 256) == 0 &&
                           g->seg_rows * 3 == g->N && g->B[1] && g->B[2] && g->rope_cos && g->rope_sin &&
                           (g->positions != nullptr || g->rope_seq_len > 0),
                       "lhrs_gemm_bf16: ROPE needs q/k/v segments with seg_rows %% 256 == 0 and cos/sin tables");
    } else {
        LHRS_CHECK_ARG(false, "lhrs_gemm_bf16: unknown epilogue %d", kind);
    }

    // CTA pairs (cta_group::2, 256 x 256 tiles) when the problem has at least one pair-tile per SM pair; LHRS_GEMM_CG=1/2 forces
    int cg = 1;
    {
        const long long pair_tiles = (long long)((g->M + 2 * BM - 1) / (2 * BM)) * ((g->N + 255) / 256);
        if (bn == 256 && (g->N % 256) == 0 && pair_tiles >= num_sms() / 2) cg = 2;
        if (forced == 1) cg = 1;
        if (forced == 2 && bn == 256 && (g->N % 256) == 0) cg = 2;
    }
    CUtensorMap tA, tB[3];
    int rc;
    if (!a_mn) rc = make_tmap(&tA, g->A, g->K, g->M, g->lda, BK, BM);
    else       rc = make_tmap(&tA, g->A, g->M, g->K, g->lda, 64, BK);
    if (rc) return rc;
    int kseg = 0;
    if (b_mn) {
        const int nseg = g->num_b;
        if (nseg > 1) kseg = g->K / nseg;
        for (int i = 0; i < 3; ++i) {
            if (i < nseg) {
                // block-diagonal mode: each segment is only b_seg_nshift columns wide (zero-filled elsewhere by TMA)
                const uint64_t seg_n = (nseg > 1 && g->b_seg_nshift > 0) ? (uint64_t)g->b_seg_nshift : (uint64_t)g->N;
                rc = make_tmap(&tB[i], g->B[i], seg_n, nseg > 1 ? kseg : g->K, g->ldb, 64, BK);
                if (rc) return rc;
            } else {
                tB[i] = tB[0];
            }
        }
    } else if (kind == LHRS_EPI_SWIGLU) {
        for (int i = 0; i < 2; ++i) {
            rc = make_tmap(&tB[i], g->B[i], g->K, g->N / 2, g->ldb, BK, bn / 2);
            if (rc) return rc;
        }
        tB[2] = tB[0];
    } else {
        const int rows = (g->num_b > 1) ? g->seg_rows : g->N;
        for (int i = 0; i < 3; ++i) {
            if (i < g->num_b) {
                rc = make_tmap(&tB[i], g->B[i], g->K, rows, g->ldb, BK, bn / cg);
                if (rc) return rc;
            } else {
                tB[i] = tB[0];
            }
        }
    }

    if (kind == LHRS_EPI_LINEAR && g->pre_up != nullptr)
        LHRS_CHECK_ARG(g->pre_gate != nullptr && (g->N % 32) == 0 && !g->d_f32 && !g->bias[0] && !g->residual && g->act == LHRS_ACT_NONE && g->split_k <= 1,
                       "lhrs_gemm_bf16: fused SwiGLU backward needs pre_gate+pre_up, N %% 32 == 0 and a plain bf16 epilogue");
    CUtensorMap tA2 = tA, tE[3] = {tB[0], tB[0], tB[0]};
    int ext_k = 0, ext_kb = 0;
    if (g->A2 != nullptr && g->ext_k > 0) {
        LHRS_CHECK_ARG(!a_mn, "lhrs_gemm_bf16: the K-extension needs a K-major A operand");
        LHRS_CHECK_ARG((g->ext_k % 8) == 0 && (g->lda2 % 8) == 0 && (g->ldb2 % 8) == 0 &&
                           (reinterpret_cast<uintptr_t>(g->A2) & 15) == 0,
                       "lhrs_gemm_bf16: K-extension alignment (ext_k=%d lda2=%lld ldb2=%lld)", g->ext_k, (long long)g->lda2, (long long)g->ldb2);
        ext_k = g->ext_k;
        const int span = (kind == LHRS_EPI_SWIGLU) ? 2 * ext_k : ext_k;
        ext_kb = (span + BK - 1) / BK;
        if (b_mn) {
            // A2 [M, ext_k] K-major; B2[0] = [ext_k, ldb2 >= N] MN-major (one matrix, whatever the K-segmentation of B)
            rc = make_tmap(&tA2, g->A2, (uint64_t)ext_k, g->M, g->lda2, BK, BM);
            if (rc) return rc;
            LHRS_CHECK_ARG(g->B2[0] != nullptr && (reinterpret_cast<uintptr_t>(g->B2[0]) & 15) == 0, "lhrs_gemm_bf16: B2[0] null/unaligned");
            rc = make_tmap(&tE[0], g->B2[0], g->N, ext_k, g->ldb2, 64, BK);
            if (rc) return rc;
            tE[1] = tE[0]; tE[2] = tE[0];
        } else {
        rc = make_tmap(&tA2, g->A2, (uint64_t)g->num_b * ext_k, g->M, g->lda2, BK, BM);
        if (rc) return rc;
        const int rows = (g->num_b > 1) ? g->seg_rows : g->N;
        for (int i = 0; i < g->num_b; ++i) {
            LHRS_CHECK_ARG(g->B2[i] != nullptr && (reinterpret_cast<uintptr_t>(g->B2[i]) & 15) == 0, "lhrs_gemm_bf16: B2[%d] null/unaligned", i);
            rc = make_tmap(&tE[i], g->B2[i], ext_k, rows, g->ldb2, BK, (kind == LHRS_EPI_SWIGLU) ? bn / 2 : bn / cg);
            if (rc) return rc;
        }
        }
    }

    GemmArgs a;
    a.ext_k = ext_k; a.ext_kb = ext_kb;
    {
        static int env_pf = -1000;
        int env_gm = 0;                    // read per call (tools/gemm_raster.py sweeps it)
        { const char* e = getenv("LHRS_GEMM_GROUP_M"); env_gm = e ? atoi(e) : 0; }
        if (env_pf == -1000) { const char* e = getenv("LHRS_GEMM_PF"); env_pf = e ? atoi(e) : -100; }
        static int env_coal = -1;
        if (env_coal < 0) { const char* e = getenv("LHRS_EPI_COAL"); env_coal = e ? atoi(e) : 1; }
        a.coalesce = env_coal;
        a.group_m = env_gm > 0 ? env_gm : 16;
        a.pf_dist = env_pf >= 0 ? env_pf : 0;
#ifdef LHRS_GEMM_DEBUG
        if (env_pf > -100) a.pf_dist = env_pf;
#endif
    }
    a.splitk = 1;
    if (g->split_k > 1) {
        LHRS_CHECK_ARG(kind == LHRS_EPI_LINEAR && g->d_f32 && ext_kb == 0 && !g->bias[0] && !g->residual && g->act == LHRS_ACT_NONE && !g->pre_gate,
                       "lhrs_gemm_bf16: split_k needs a plain LINEAR epilogue with an fp32 (zero-initialised) output");
        const int nkb = (g->K + BK - 1) / BK;
        int sk = g->split_k < nkb ? g->split_k : nkb;
        while (sk > 1 && ((nkb + sk - 1) / sk) * (sk - 1) >= nkb) --sk;   // every split must own at least one k-block
        a.splitk = sk;
    }
    a.M = g->M; a.N = g->N; a.K = g->K;
    a.kseg = kseg;
    a.kseg_nshift = (kseg > 0) ? g->b_seg_nshift : 0;
    a.num_b = (kseg > 0) ? 1 : g->num_b;   // K-stacked segments look like a single B to the epilogue
    a.seg_rows = (a.num_b > 1) ? g->seg_rows : g->N;
    a.act = g->act; a.alpha = g->alpha;
    for (int i = 0; i < 3; ++i) a.bias[i] = reinterpret_cast<const __nv_bfloat16*>(g->bias[i]);
    a.residual = reinterpret_cast<const __nv_bfloat16*>(g->residual);
    a.ldr = g->ldr;
    a.D = g->D; a.ldd = g->ldd; a.d_f32 = g->d_f32;
    a.row_map = g->row_map;
    a.rope_cos = g->rope_cos; a.rope_sin = g->rope_sin; a.positions = g->positions; a.rope_seq_len = g->rope_seq_len;
    a.pre_gate = reinterpret_cast<__nv_bfloat16*>(g->pre_gate);
    a.pre_up = reinterpret_cast<__nv_bfloat16*>(g->pre_up);
    a.drop_key = g->drop_key; a.drop_t = g->drop_t;
    if (g->drop_t > 0)
        LHRS_CHECK_ARG(kind == LHRS_EPI_LINEAR && (g->N % 32) == 0 && g->drop_t < 256 && g->split_k <= 1 && !g->pre_up,
                       "lhrs_gemm_bf16: the dropout mask needs a LINEAR epilogue with N %% 32 == 0");

    if (cg == 2) {
        if (kind == LHRS_EPI_SWIGLU) return launch<256, LHRS_EPI_SWIGLU, false, false, 2>(tA, tB, tA2, tE, a, stream);
        if (kind == LHRS_EPI_ROPE) return launch<256, LHRS_EPI_ROPE, false, false, 2>(tA, tB, tA2, tE, a, stream);
        if (!a_mn && !b_mn) return launch<256, LHRS_EPI_LINEAR, false, false, 2>(tA, tB, tA2, tE, a, stream);
        if (!a_mn && b_mn) return launch<256, LHRS_EPI_LINEAR, false, true, 2>(tA, tB, tA2, tE, a, stream);
        if (a_mn && b_mn) return launch<256, LHRS_EPI_LINEAR, true, true, 2>(tA, tB, tA2, tE, a, stream);
        return launch<256, LHRS_EPI_LINEAR, true, false, 2>(tA, tB, tA2, tE, a, stream);
    }
    if (kind == LHRS_EPI_SWIGLU) return launch<256, LHRS_EPI_SWIGLU, false, false>(tA, tB, tA2, tE, a, stream);
    if (kind == LHRS_EPI_ROPE) return launch<256, LHRS_EPI_ROPE, false, false>(tA, tB, tA2, tE, a, stream);
    if (bn == 256) {
        if (!a_mn && !b_mn) return launch<256, LHRS_EPI_LINEAR, false, false>(tA, tB, tA2, tE, a, stream);
        if (!a_mn && b_mn) return launch<256, LHRS_EPI_LINEAR, false, true>(tA, tB, tA2, tE, a, stream);
        if (a_mn && b_mn) return launch<256, LHRS_EPI_LINEAR, true, true>(tA, tB, tA2, tE, a, stream);
        return launch<256, LHRS_EPI_LINEAR, true, false>(tA, tB, tA2, tE, a, stream);
    } else {
        if (!a_mn && !b_mn) return launch<128, LHRS_EPI_LINEAR, false, false>(tA, tB, tA2, tE, a, stream);
        if (!a_mn && b_mn) return launch<128, LHRS_EPI_LINEAR, false, true>(tA, tB, tA2, tE, a, stream);
        if (a_mn && b_mn) return launch<128, LHRS_EPI_LINEAR, true, true>(tA, tB, tA2, tE, a, stream);
        return launch<128, LHRS_EPI_LINEAR, true, false>(tA, tB, tA2, tE, a, stream);
    }
}
