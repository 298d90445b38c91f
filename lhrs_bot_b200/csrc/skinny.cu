// Streaming kernels for the rank-r side products of LoRA (text_modal.py:133-151: peft wraps q,k,v,o,gate,up,down of every layer).
// With r = 16 these products have 16..48 output columns against 4096..22016 input columns: they are HBM-bound (one pass over an
// activation of 64..360 MB per call), not tensor-bound, so they do not belong on 128-row tcgen05 tiles (which left them at ~1/3 of
// HBM speed and 13 % of the SFT step).  Two kernels, both built from warp-level mma.sync.m16n8k16 tiles fed by a multi-stage cp.async
// ring (tools/lora_bench.py; round 1, M = 8192: 2.4-4.4 TB/s, 21.7 ms per SFT step for all 512 side products against 32 ms on the
// tcgen05 tiles; 128-wide k chunks beat 64-wide ones by 10 %, the L2::256B prefetch hint and 256-column row-reduce blocks change
// nothing).  Round 2 (ncu --set full): with ~1.4 CTAs of 4 warps per SM the panels were bound by instruction latency (42 % issue
// slots busy at 18 % occupancy, stalls `wait` / `short_scoreboard`), not by HBM: 8 warps per CTA, 16-row panels where they pay and
// no integer divisions per chunk brought dT qkv from 66.6 to 47.1 us and dT gate/up from 100.4 to 71.7 us at M = 6740 (3.5 / 4.1
// TB/s); a deeper ring (LHRS_SKINNY_DEEP) does not help.
//   lora_panel_kernel      out[M, n] = alpha * X[M, K] · W          one CTA per 16- or 32-row panel, X streamed once, W from L2
//        W_KN = false:  W = [n, K] K-major (forward  T  = s · x · [A_0;A_1;..]^T)
//        W_KN = true :  W_s = [kseg, 16] per K segment, block diagonal (backward dT_s = s · dy_s · B_s, B_s = lora_B [out, r])
//   lora_rowreduce_kernel  G[C, n] = P[M, C]^T · Q[M, n]            one CTA per 128-column block x row split, P streamed once;
//        per-split fp32 partials are summed, cast and laid out by lora_reduce_store_kernel (no atomics: deterministic)
//        (backward dA = dT^T · x  -> [n, in];   dB_s = dy_s^T · T_s -> [out, r] per segment)
#include <stdlib.h>

#include <cooperative_groups.h>

#include "host_common.h"
#include "ptx.cuh"
#include "dropout.cuh"

namespace lhrs {

typedef __nv_bfloat16 bf16;

constexpr int SK_THREADS = 128;
// panel kernel: PN_BM (32 or 16) rows x BK k per stage, 8 warps per CTA = (PN_BM / 16 row tiles) x (4 or 8 k-shares of every chunk).
// ncu (round 2): 42 % issue-slot utilisation at 18 % warp occupancy, top stalls `wait` and `short_scoreboard` — the kernel is bound by
// instruction latency, not by HBM, as long as a 32-row panel per CTA leaves only ~1.4 CTAs per SM at M = 6.7-8 k rows; 16-row panels
// double the resident warps (launch_panel picks them while the grid stays under 4 CTAs per SM).
constexpr int PN_THREADS = 256;
constexpr int RR_ST = 4;                  // row-reduce kernel: BR rows x BC columns (16 KB) per stage
constexpr int RR_THREADS = 256;           //   4 column groups x 2 halves of every chunk's rows (same reason as the panel kernel's 8 warps)

__device__ __forceinline__ void sk_cp16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
// same, asking L2 to fetch the whole 256-byte pair of lines: fewer, longer DRAM bursts for streams read 128 bytes per row at a time
__device__ __forceinline__ void sk_cp16_l2(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void sk_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void sk_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void sk_ldsm(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void sk_ldsm_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void sk_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// [rows][COLS] bf16 tile, 16-byte chunks XOR-swizzled by the row within each 128-byte group
template <int COLS>
__device__ __forceinline__ uint32_t offsw(int row, int chunk) {
    return static_cast<uint32_t>(row * (COLS * 2) + (((chunk & ~7) | ((chunk & 7) ^ (row & 7))) << 4));
}

struct PanelArgs {
    const bf16* X; long long ldx; int M, K;
    const bf16* W[3]; long long ldw;   // !W_KN: W[0] = [n, ldw];  W_KN: W[s] = [kseg, ldw] with ldw == 16
    int kseg;                          // W_KN: K columns per segment (K = nseg * kseg)
    float alpha;
    bf16* out; long long ldo;
    // !W_KN, drop_t > 0: peft input dropout — the 16 output columns of projection p see X masked by key[p] (dropout.cuh);
    // the survivors' 1/keep is folded into alpha by the caller
    uint32_t key[3];
    int drop_t;
};

// zero the halves of a 2 x bf16 register whose elements (row, col) and (row, col + 1) the mask word drops (col even)
// (the four byte draws of the word are compared at once: __vcmpgeu4 gives 0xff per surviving byte; a byte permute then widens
//  the two bytes of interest to the two bf16 halves — 4 integer instructions per register instead of 10)
__device__ __forceinline__ uint32_t drop_pair_cols(uint32_t v, uint32_t word, uint32_t row, int t) {
    const uint32_t k4 = __vcmpgeu4(word, static_cast<uint32_t>(t) * 0x01010101u);       // byte i: 0xff = keep draw i
    return v & __byte_perm(k4, 0u, (row & 1u) ? 0x3322u : 0x1100u);                     // draws 2r, 2r+1 -> low / high half
}
// same for a register holding (row, col) and (row + 1, col) (row even): the transposed fragments of the row-reduce kernel
__device__ __forceinline__ uint32_t drop_pair_rows(uint32_t v, uint32_t word, uint32_t col, int t) {
    const uint32_t k4 = __vcmpgeu4(word, static_cast<uint32_t>(t) * 0x01010101u);
    return v & __byte_perm(k4, 0u, (col & 1u) ? 0x3311u : 0x2200u);                     // draws c, 2 + c -> low / high half
}

template <int NT, bool W_KN, int PN_BK, int PN_ST, bool L2H, int PN_BM>
__global__ void __launch_bounds__(PN_THREADS)
lora_panel_kernel(const PanelArgs a) {
    constexpr int RT = PN_BM / 16;                                       // 16-row mma tiles of the panel
    constexpr int PN_KW = (PN_THREADS / 32) / RT;                        // warps sharing the k-steps of every chunk of one row tile
    constexpr int X_STAGE = PN_BM * PN_BK * 2;
    constexpr int W_STAGE = W_KN ? PN_BK * 16 * 2 : NT * 8 * PN_BK * 2;
    constexpr int STAGE = X_STAGE + W_STAGE;
    constexpr int CPR = PN_BK / 8;                                       // 16-byte chunks per tile row
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t s0 = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, mat = lane >> 3;
    const int wr = warp % RT, wk = warp / RT;                            // row tile x k share of every chunk
    const int row0 = blockIdx.x * PN_BM;
    const int nchunks = a.K / PN_BK;
    const int cps = W_KN ? a.kseg / PN_BK : nchunks;                    // chunks per K segment
    int seg_c = 0, seg_left = cps;                                       // segment of the chunk being consumed (no division per chunk)
    int seg_l = 0, kin_l = 0;                                            // segment / k offset of the next chunk to be requested

    auto load = [&](int chunk, int stage) {
        const uint32_t sx = s0 + stage * STAGE, sw = sx + X_STAGE;
        const int k0 = chunk * PN_BK;
#pragma unroll
        for (int i = 0; i < (PN_BM * CPR + PN_THREADS - 1) / PN_THREADS; ++i) {
            const int idx = tid + i * PN_THREADS, r = idx / CPR, c = idx % CPR;
            if ((PN_BM * CPR) % PN_THREADS != 0 && idx >= PN_BM * CPR) break;
            const int gr = row0 + r;
            const bool ok = gr < a.M;
            const bf16* src = a.X + static_cast<long long>(ok ? gr : 0) * a.ldx + k0 + c * 8;
            if (L2H) sk_cp16_l2(sx + offsw<PN_BK>(r, c), src, ok); else sk_cp16(sx + offsw<PN_BK>(r, c), src, ok);
        }
        if constexpr (W_KN) {
            const bf16* w = a.W[seg_l] + static_cast<long long>(kin_l) * 16;   // chunks are requested in order: running segment / offset
            kin_l += PN_BK;
            if (kin_l == a.kseg) { kin_l = 0; ++seg_l; }
            for (int idx = tid; idx < PN_BK * 2; idx += PN_THREADS) {       // BK k-rows x 2 chunks
                const int r = idx >> 1, c = idx & 1;
                sk_cp16(sw + r * 32 + c * 16, w + r * 16 + c * 8, true);
            }
        } else {
            for (int idx = tid; idx < NT * 8 * CPR; idx += PN_THREADS) {
                const int r = idx / CPR, c = idx % CPR;
                sk_cp16(sw + offsw<PN_BK>(r, c), a.W[0] + static_cast<long long>(r) * a.ldw + k0 + c * 8, true);
            }
        }
    };

    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

#pragma unroll
    for (int s = 0; s < PN_ST - 1; ++s) {
        if (s < nchunks) load(s, s);
        sk_commit();
    }
    for (int ch = 0; ch < nchunks; ++ch) {
        sk_wait<PN_ST - 2>();
        __syncthreads();
        if (ch + PN_ST - 1 < nchunks) load(ch + PN_ST - 1, (ch + PN_ST - 1) % PN_ST);
        sk_commit();
        const uint32_t sx = s0 + (ch % PN_ST) * STAGE, sw = sx + X_STAGE;
        if (W_KN && seg_left == 0) { ++seg_c; seg_left = cps; }
        --seg_left;
#pragma unroll
        for (int kk = 0; kk < PN_BK / (16 * PN_KW); ++kk) {
            const int ks = wk * (PN_BK / (16 * PN_KW)) + kk;                // this warp's k-steps of the chunk
            uint32_t af[4];
            sk_ldsm(sx + offsw<PN_BK>(wr * 16 + (lane & 15), ks * 2 + (lane >> 4)), af[0], af[1], af[2], af[3]);
            if constexpr (W_KN) {
                uint32_t b0, b1, b2, b3;
                sk_ldsm_t(sw + (ks * 16 + (mat & 1) * 8 + (lane & 7)) * 32 + (mat >> 1) * 16, b0, b1, b2, b3);
                const int s = seg_c;                                        // block-uniform: which segment's two n-tiles
                if (s == 0) { sk_mma(acc[0], af, b0, b1); sk_mma(acc[1], af, b2, b3); }
                if constexpr (NT >= 4) { if (s == 1) { sk_mma(acc[2], af, b0, b1); sk_mma(acc[3], af, b2, b3); } }
                if constexpr (NT >= 6) { if (s == 2) { sk_mma(acc[4], af, b0, b1); sk_mma(acc[5], af, b2, b3); } }
            } else {
#pragma unroll
                for (int np = 0; np < NT / 2; ++np) {
                    uint32_t b0, b1, b2, b3;
                    sk_ldsm(sw + offsw<PN_BK>(np * 16 + (mat >> 1) * 8 + (lane & 7), ks * 2 + (mat & 1)), b0, b1, b2, b3);
                    if (a.drop_t > 0) {
                        // fragment elements: af[0] (r_lo, k..k+1), af[1] (r_hi, k..k+1), af[2] (r_lo, k+8..), af[3] (r_hi, k+8..)
                        const uint32_t r_lo = static_cast<uint32_t>(row0 + wr * 16 + (lane >> 2)), r_hi = r_lo + 8;
                        const uint32_t kc = static_cast<uint32_t>(ch * PN_BK + ks * 16 + (lane & 3) * 2);
                        uint32_t m[4];
                        m[0] = drop_pair_cols(af[0], drop_word(a.key[np], r_lo, kc, static_cast<uint32_t>(a.K)), r_lo, a.drop_t);
                        m[1] = drop_pair_cols(af[1], drop_word(a.key[np], r_hi, kc, static_cast<uint32_t>(a.K)), r_hi, a.drop_t);
                        m[2] = drop_pair_cols(af[2], drop_word(a.key[np], r_lo, kc + 8, static_cast<uint32_t>(a.K)), r_lo, a.drop_t);
                        m[3] = drop_pair_cols(af[3], drop_word(a.key[np], r_hi, kc + 8, static_cast<uint32_t>(a.K)), r_hi, a.drop_t);
                        sk_mma(acc[np * 2], m, b0, b1);
                        sk_mma(acc[np * 2 + 1], m, b2, b3);
                    } else {
                        sk_mma(acc[np * 2], af, b0, b1);
                        sk_mma(acc[np * 2 + 1], af, b2, b3);
                    }
                }
            }
        }
    }
    sk_wait<0>();
    __syncthreads();
    // the PN_KW k-shares of each 16-row tile meet in shared memory; warps with wk == 0 finish and store
    float* red = reinterpret_cast<float*>(smem);                            // [PN_KW - 1][RT row tiles][NT][32 lanes][4]
    if (wk > 0) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            *reinterpret_cast<float4*>(red + ((((wk - 1) * RT + wr) * NT + i) * 32 + lane) * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    __syncthreads();
    if (wk == 0) {
        const int g = lane >> 2, t4 = lane & 3;
        const int r_lo = row0 + wr * 16 + g, r_hi = r_lo + 8;
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < PN_KW - 1; ++w) {                           // fixed order: deterministic
                const float4 v = *reinterpret_cast<const float4*>(red + (((w * RT + wr) * NT + i) * 32 + lane) * 4);
                o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
            }
            const int col = i * 8 + t4 * 2;
            if (r_lo < a.M)
                *reinterpret_cast<uint32_t*>(a.out + static_cast<long long>(r_lo) * a.ldo + col) = pack_bf16((acc[i][0] + o.x) * a.alpha, (acc[i][1] + o.y) * a.alpha);
            if (r_hi < a.M)
                *reinterpret_cast<uint32_t*>(a.out + static_cast<long long>(r_hi) * a.ldo + col) = pack_bf16((acc[i][2] + o.z) * a.alpha, (acc[i][3] + o.w) * a.alpha);
        }
    }
}

struct ReduceArgs {
    const bf16* P; long long ldp; int M, C;
    const bf16* Q; long long ldq;
    int seg_c;             // > 0: block diagonal — column block c0 belongs to segment c0 / seg_c and uses Q columns [seg*NT*8, +NT*8)
    int rows_per_split;    // multiple of RR_BR
    float* partial;        // [nsplit][C][NT*8] (only without the cluster reduction)
    // cluster reduction: the row splits of one column block form a thread-block cluster (1 x nsplit); their partial tiles meet in
    // distributed shared memory and every rank sums, scales, casts and stores an interleaved share of the block's result — no scratch,
    // no second launch (summation order = rank order: deterministic)
    int cluster;
    int transpose;         // 1: dst[0][j, c] (ld = ldd);  0: dst[seg][c - seg*seg_c, j] (ld = ldd)
    bf16* dst[3];
    long long ldd;
    float alpha;
    // drop_t > 0 (dA form, seg_c == 0): the 16 Q columns of projection p contract with P masked by key[p] (ld of the mask = C)
    uint32_t key[3];
    int drop_t;
};

template <int NT, int RR_BC, bool L2H>
__global__ void __launch_bounds__(RR_THREADS)
lora_rowreduce_kernel(const ReduceArgs a) {
    constexpr int RR_BR = 8192 / RR_BC;               // 64 x 128 or 32 x 256
    constexpr int P_STAGE = RR_BR * RR_BC * 2;        // 16 KB
    constexpr int Q_STAGE = RR_BR * NT * 16;          // BR rows x NT chunks
    constexpr int STAGE = P_STAGE + Q_STAGE;
    constexpr int CPR = RR_BC / 8;
    constexpr int MT = RR_BC / 64;                    // 16-column tiles per warp
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t s0 = smem_u32(smem);
    const int tid = threadIdx.x, warp_all = tid >> 5, lane = tid & 31, mat = lane >> 3;
    const int warp = warp_all & 3;                    // column group: 16-column tiles [warp * MT, warp * MT + MT)
    const int kh = warp_all >> 2;                     // which half of every chunk's k-steps (rows): two warps per column group
    const int c0 = blockIdx.x * RR_BC;
    const int split = blockIdx.y;
    const int m_begin = split * a.rows_per_split;
    const int m_end = min(a.M, m_begin + a.rows_per_split);
    const int nchunks = m_end > m_begin ? (m_end - m_begin + RR_BR - 1) / RR_BR : 0;
    const int qcol0 = a.seg_c > 0 ? (c0 / a.seg_c) * NT * 8 : 0;

    auto load = [&](int chunk, int stage) {
        const uint32_t sp = s0 + stage * STAGE, sq = sp + P_STAGE;
        const int m0 = m_begin + chunk * RR_BR;
#pragma unroll
        for (int i = 0; i < (RR_BR * CPR) / RR_THREADS; ++i) {
            const int idx = tid + i * RR_THREADS, r = idx / CPR, c = idx % CPR;
            const int gm = m0 + r;
            const bool ok = gm < m_end && c0 + c * 8 < a.C;
            const bf16* src = a.P + static_cast<long long>(ok ? gm : 0) * a.ldp + (ok ? c0 + c * 8 : 0);
            if (L2H) sk_cp16_l2(sp + offsw<RR_BC>(r, c), src, ok); else sk_cp16(sp + offsw<RR_BC>(r, c), src, ok);
        }
        for (int idx = tid; idx < RR_BR * NT; idx += RR_THREADS) {
            const int r = idx / NT, c = idx - r * NT;
            const int gm = m0 + r;
            const bool ok = gm < m_end;
            sk_cp16(sq + (r * NT + c) * 16, a.Q + static_cast<long long>(ok ? gm : 0) * a.ldq + qcol0 + c * 8, ok);
        }
    };

    float acc[MT][NT][4];
#pragma unroll
    for (int t = 0; t < MT; ++t)
#pragma unroll
        for (int i = 0; i < NT; ++i) { acc[t][i][0] = acc[t][i][1] = acc[t][i][2] = acc[t][i][3] = 0.f; }

#pragma unroll
    for (int s = 0; s < RR_ST - 1; ++s) {
        if (s < nchunks) load(s, s);
        sk_commit();
    }
    for (int ch = 0; ch < nchunks; ++ch) {
        sk_wait<RR_ST - 2>();
        __syncthreads();
        if (ch + RR_ST - 1 < nchunks) load(ch + RR_ST - 1, (ch + RR_ST - 1) % RR_ST);
        sk_commit();
        const uint32_t sp = s0 + (ch % RR_ST) * STAGE, sq = sp + P_STAGE;
#pragma unroll
        for (int kk = 0; kk < RR_BR / 32; ++kk) {
            const int ks = kh * (RR_BR / 32) + kk;
            uint32_t bq[NT / 2][4];
#pragma unroll
            for (int np = 0; np < NT / 2; ++np)
                sk_ldsm_t(sq + ((ks * 16 + (mat & 1) * 8 + (lane & 7)) * NT + np * 2 + (mat >> 1)) * 16, bq[np][0], bq[np][1], bq[np][2], bq[np][3]);
#pragma unroll
            for (int t = 0; t < MT; ++t) {
                // A = P^T: fragment rows are the columns of P, fragment k the rows -> transposed 8x8 loads
                uint32_t af[4];
                sk_ldsm_t(sp + offsw<RR_BC>(ks * 16 + (mat >> 1) * 8 + (lane & 7), warp * (MT * 2) + t * 2 + (mat & 1)), af[0], af[1], af[2], af[3]);
                if (a.drop_t > 0) {
                    // transposed fragments: af[i] holds P[m, c] (low) and P[m + 1, c] (high), m even:
                    //   m = chunk row0 + ks*16 + (i >> 1)*8 + 2*t4,  c = c0 + (warp*MT*2 + t*2 + (i & 1))*8 + g
                    const uint32_t mb = static_cast<uint32_t>(m_begin + ch * RR_BR + ks * 16 + (lane & 3) * 2);
                    const uint32_t cb = static_cast<uint32_t>(c0 + (warp * (MT * 2) + t * 2) * 8 + (lane >> 2));
#pragma unroll
                    for (int np = 0; np < NT / 2; ++np) {
                        uint32_t m[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t mr = mb + (i >> 1) * 8, cc = cb + (i & 1) * 8;
                            m[i] = drop_pair_rows(af[i], drop_word(a.key[np], mr, cc, static_cast<uint32_t>(a.C)), cc, a.drop_t);
                        }
                        sk_mma(acc[t][np * 2], m, bq[np][0], bq[np][1]);
                        sk_mma(acc[t][np * 2 + 1], m, bq[np][2], bq[np][3]);
                    }
                } else {
#pragma unroll
                    for (int np = 0; np < NT / 2; ++np) {
                        sk_mma(acc[t][np * 2], af, bq[np][0], bq[np][1]);
                        sk_mma(acc[t][np * 2 + 1], af, bq[np][2], bq[np][3]);
                    }
                }
            }
        }
    }
    sk_wait<0>();
    const int g = lane >> 2, t4 = lane & 3;
    {   // the two k-halves of every column group meet in shared memory (fixed order); warps 0-3 carry the sums from here on
        __syncthreads();                                   // every warp is done with the pipeline buffers
        float* mrg = reinterpret_cast<float*>(smem);       // [4 groups][MT][NT][32 lanes][4] fp32 (<= 24.6 KB)
        if (kh == 1) {
#pragma unroll
            for (int t = 0; t < MT; ++t)
#pragma unroll
                for (int i = 0; i < NT; ++i)
                    *reinterpret_cast<float4*>(mrg + (((warp * MT + t) * NT + i) * 32 + lane) * 4) = make_float4(acc[t][i][0], acc[t][i][1], acc[t][i][2], acc[t][i][3]);
        }
        __syncthreads();
        if (kh == 0) {
#pragma unroll
            for (int t = 0; t < MT; ++t)
#pragma unroll
                for (int i = 0; i < NT; ++i) {
                    const float4 o = *reinterpret_cast<const float4*>(mrg + (((warp * MT + t) * NT + i) * 32 + lane) * 4);
                    acc[t][i][0] += o.x; acc[t][i][1] += o.y; acc[t][i][2] += o.z; acc[t][i][3] += o.w;
                }
        }
    }
    if (a.cluster) {
        namespace cg = cooperative_groups;
        cg::cluster_group cluster = cg::this_cluster();
        constexpr int NQ = NT * 8;
        __syncthreads();                                   // the merge area is read: it may be overwritten
        float* red = reinterpret_cast<float*>(smem);       // [RR_BC][NQ] fp32 (<= 48 KB, inside the first stages)
#pragma unroll
        for (int t = 0; t < MT && kh == 0; ++t) {
            const int cl = warp * (MT * 16) + t * 16 + g;
#pragma unroll
            for (int i = 0; i < NT; ++i) {
                const int j = i * 8 + t4 * 2;
                *reinterpret_cast<float2*>(red + cl * NQ + j) = make_float2(acc[t][i][0], acc[t][i][1]);
                *reinterpret_cast<float2*>(red + (cl + 8) * NQ + j) = make_float2(acc[t][i][2], acc[t][i][3]);
            }
        }
        cluster.sync();
        {   // every rank finishes an interleaved share of the block's [RR_BC x NQ] result from all ranks' tiles
            const unsigned nranks = cluster.num_blocks(), me = cluster.block_rank();
            const int seg = a.seg_c > 0 ? c0 / a.seg_c : 0;
            bf16* out = a.dst[a.transpose ? 0 : seg];
            const float* tiles[8];
#pragma unroll
            for (unsigned r = 0; r < 8; ++r) tiles[r] = r < nranks ? cluster.map_shared_rank(red, r) : red;
            for (int idx = me * RR_THREADS + tid; idx < RR_BC * NQ; idx += nranks * RR_THREADS) {
                // transpose: consecutive threads walk the columns of one output row; else consecutive elements of [c][j]
                const int cl = a.transpose ? idx % RR_BC : idx / NQ;
                const int j = a.transpose ? idx / RR_BC : idx % NQ;
                float sum = 0.f;
#pragma unroll
                for (unsigned r = 0; r < 8; ++r)
                    if (r < nranks) sum += tiles[r][cl * NQ + j];
                const int c = c0 + cl;
                if (c < a.C && out != nullptr) {
                    if (a.transpose) out[static_cast<long long>(j) * a.ldd + c] = __float2bfloat16_rn(sum * a.alpha);
                    else out[static_cast<long long>(c - seg * a.seg_c) * a.ldd + j] = __float2bfloat16_rn(sum * a.alpha);
                }
            }
        }
        cluster.sync();                                    // shared memory stays alive until every rank has read its share
        return;
    }
    float* dst = a.partial + static_cast<long long>(split) * a.C * (NT * 8);
#pragma unroll
    for (int t = 0; t < MT && kh == 0; ++t) {
        const int c_lo = c0 + warp * (MT * 16) + t * 16 + g, c_hi = c_lo + 8;
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            const int j = i * 8 + t4 * 2;
            if (c_lo < a.C) *reinterpret_cast<float2*>(dst + static_cast<long long>(c_lo) * (NT * 8) + j) = make_float2(acc[t][i][0], acc[t][i][1]);
            if (c_hi < a.C) *reinterpret_cast<float2*>(dst + static_cast<long long>(c_hi) * (NT * 8) + j) = make_float2(acc[t][i][2], acc[t][i][3]);
        }
    }
}

// sum the row-split partials, cast, and lay the result out:
//   transpose = 1:  dst0[j, c]            (ld = ldd)                      — dA = dT^T x as [n, in]
//   transpose = 0:  dst_s[c - s*seg_c, j] (ld = ldd), s = c / seg_c       — dB_s = dy_s^T T_s as [out, r] per segment
__global__ void __launch_bounds__(256)
lora_reduce_store_kernel(const float* __restrict__ partial, int nsplit, int C, int n, int transpose, int seg_c, bf16* d0, bf16* d1, bf16* d2,
                         long long ldd, float alpha) {
    const long long i = blockIdx.x * 256LL + threadIdx.x;
    if (i >= static_cast<long long>(C) * n) return;
    int c, j;
    if (transpose) { j = static_cast<int>(i / C); c = static_cast<int>(i - static_cast<long long>(j) * C); }
    else { c = static_cast<int>(i / n); j = static_cast<int>(i - static_cast<long long>(c) * n); }
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += partial[(static_cast<long long>(k) * C + c) * n + j];
    s *= alpha;
    if (transpose) {
        d0[static_cast<long long>(j) * ldd + c] = __float2bfloat16_rn(s);
    } else {
        const int seg = seg_c > 0 ? c / seg_c : 0;
        bf16* d = seg == 0 ? d0 : (seg == 1 ? d1 : d2);
        if (d != nullptr) d[static_cast<long long>(c - seg * (seg_c > 0 ? seg_c : 0)) * ldd + j] = __float2bfloat16_rn(s);
    }
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <int NT, bool W_KN, int BK, int ST, bool L2H, int BM>
static int launch_panel_cfg(const PanelArgs& a, cudaStream_t st) {
    constexpr int STAGE = BM * BK * 2 + (W_KN ? BK * 16 * 2 : NT * 8 * BK * 2);
    constexpr int RED = (8 / (BM / 16) - 1) * (BM / 16) * NT * 32 * 4 * 4;   // k-share meeting area (aliases the ring after the loop)
    constexpr int SMEM = STAGE * ST > RED ? STAGE * ST : RED;
    auto kern = lora_panel_kernel<NT, W_KN, BK, ST, L2H, BM>;
    static bool attr = false;
    if (!attr) { LHRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); attr = true; }
    const bool prof = prof_on();
    if (prof) prof_begin(PROF_SKINNY, 2.0 * a.M * (double)(NT * 8) * (W_KN ? a.kseg : a.K), 2.0 * (double)a.M * a.K, st);
    kern<<<(a.M + BM - 1) / BM, PN_THREADS, SMEM, st>>>(a);
    if (prof) prof_end(st);
    LHRS_LAUNCH_CHECK("lora_panel_kernel");
    return LHRS_OK;
}
template <int NT, bool W_KN>
static int launch_panel(const PanelArgs& a, cudaStream_t st) {
    static int bk = -1, l2h = -1, bm = -1;
    if (bk < 0) { bk = env_int("LHRS_SKINNY_BK", 128); l2h = env_int("LHRS_SKINNY_L2HINT", 0); bm = env_int("LHRS_SKINNY_BM", 0); }
    // 16-row panels while 32-row ones would leave fewer than two CTAs per SM (LHRS_SKINNY_BM=16 / 32 forces one)
    // (measured at M = 6740: dT qkv 66.6 -> 58.5 us, dT gate/up 100.4 -> 87.0, dT o / down 25.6 -> 23.6, T down 43.1 -> 40.0; the
    //  48-column forward form re-stages a 12 KB factor chunk per 4 KB of activation with 16 rows and loses: 27.6 -> 31.7 us)
    const bool small = bm == 16 || (bm != 32 && (a.M + 31) / 32 < 2 * num_sms() && (W_KN || NT <= 4));
    if (bk == 128 && a.K % 128 == 0 && (!W_KN || a.kseg % 128 == 0)) {
        if (l2h) return launch_panel_cfg<NT, W_KN, 128, 5, true, 32>(a, st);
        return small ? launch_panel_cfg<NT, W_KN, 128, 6, false, 16>(a, st) : launch_panel_cfg<NT, W_KN, 128, 5, false, 32>(a, st);
    }
    return l2h ? launch_panel_cfg<NT, W_KN, 64, 8, true, 32>(a, st) : launch_panel_cfg<NT, W_KN, 64, 8, false, 32>(a, st);
}

template <int NT, int BC, bool L2H>
static int launch_rowreduce_cfg(const ReduceArgs& a, int nsplit, cudaStream_t st) {
    constexpr int SMEM = (16384 + (8192 / BC) * NT * 16) * RR_ST;
    auto kern = lora_rowreduce_kernel<NT, BC, L2H>;
    static bool attr = false;
    if (!attr) { LHRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM)); attr = true; }
    const bool prof = prof_on();
    if (prof) prof_begin(PROF_SKINNY, 2.0 * a.M * (double)a.C * (NT * 8), 2.0 * (double)a.M * a.C, st);
    if (a.cluster) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((a.C + BC - 1) / BC, nsplit); cfg.blockDim = dim3(RR_THREADS); cfg.dynamicSmemBytes = SMEM; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = nsplit; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        LHRS_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    } else {
        kern<<<dim3((a.C + BC - 1) / BC, nsplit), RR_THREADS, SMEM, st>>>(a);
    }
    if (prof) prof_end(st);
    LHRS_LAUNCH_CHECK("lora_rowreduce_kernel");
    return LHRS_OK;
}
static int rowreduce_bc(int C, int seg_c) {
    static int bc = -1;
    if (bc < 0) bc = env_int("LHRS_SKINNY_BC", 128);
    return (bc == 256 && C % 256 == 0 && (seg_c == 0 || seg_c % 256 == 0)) ? 256 : 128;
}
template <int NT>
static int launch_rowreduce(const ReduceArgs& a, int nsplit, int bc, cudaStream_t st) {
    static int l2h = -1;
    if (l2h < 0) l2h = env_int("LHRS_SKINNY_L2HINT", 0);
    if (bc == 256) return l2h ? launch_rowreduce_cfg<NT, 256, true>(a, nsplit, st) : launch_rowreduce_cfg<NT, 256, false>(a, nsplit, st);
    return l2h ? launch_rowreduce_cfg<NT, 128, true>(a, nsplit, st) : launch_rowreduce_cfg<NT, 128, false>(a, nsplit, st);
}

// ------------------------------------------------------------------------------------------------ dX correction under dropout
// dx[M, C] += inv * sum_p mask_p o (dT_p[M, 16] · A_p[16, C]) — the LoRA branch's input gradient when peft's input dropout is on.
// Without dropout this term rides on the dX GEMM as a K-extension; with it every projection's rank-16 product has its own
// elementwise mask, so it is one streaming read-modify-write pass over dx (HBM-bound: 2 bytes in, 2 out per element) with the
// products on mma.sync and the mask bits regenerated from the call seed (dropout.cuh).  CTA = 64 rows x 128 columns.
struct DxDropArgs {
    bf16* dx; long long ldx; int M, C;
    const bf16* dT; long long ldt;
    const bf16* A[3];              // lora_A of the projections, [16, C] each
    uint32_t key[3];
    int t;
    float inv;
};

template <int NP>
__global__ void __launch_bounds__(SK_THREADS)
lora_dx_dropout_kernel(const DxDropArgs a) {
    constexpr int TR = 64, TC = 128;
    // operands first, then (after a barrier) the fp32 result tile in the same bytes
    __shared__ __align__(128) uint8_t s_buf[TR * (TC + 4) * 4];
    uint8_t* s_dt = s_buf;                                          // [64][NP*16] bf16, 16-byte chunks XOR-swizzled by row (<= 6 KB)
    uint8_t* s_a = s_buf + 6144;                                    // [NP][16][128] bf16, offsw<128> (<= 12 KB)
    float* s_acc = reinterpret_cast<float*>(s_buf);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, mat = lane >> 3;
    const int c0 = blockIdx.x * TC, row0 = blockIdx.y * TR;
    const uint32_t sdt = smem_u32(s_dt), sa = smem_u32(s_a);
    // stage dT rows (NP*2 chunks of 16 B per row) and the A tiles (16 rows x 16 chunks per projection)
    for (int idx = tid; idx < TR * NP * 2; idx += SK_THREADS) {
        const int r = idx / (NP * 2), c = idx - r * (NP * 2);
        const bool ok = row0 + r < a.M;
        sk_cp16(sdt + r * (NP * 32) + ((c ^ (r & 1)) << 4), a.dT + static_cast<long long>(ok ? row0 + r : 0) * a.ldt + c * 8, ok);
    }
    for (int idx = tid; idx < NP * 16 * 16; idx += SK_THREADS) {
        const int p = idx >> 8, r = (idx >> 4) & 15, c = idx & 15;
        const bool ok = c0 + c * 8 < a.C;
        sk_cp16(sa + p * (16 * TC * 2) + offsw<TC>(r, c), a.A[p] + static_cast<long long>(r) * a.C + (ok ? c0 + c * 8 : 0), ok);
    }
    sk_commit();
    sk_wait<0>();
    __syncthreads();
    const int g = lane >> 2, t4 = lane & 3;
    const uint32_t r_lo = static_cast<uint32_t>(row0 + warp * 16 + g), r_hi = r_lo + 8;
    float acc[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        uint32_t af[4];
        {   // A fragment of the 16 x 16 dT_p tile: rows warp*16 + (lane & 15), 16-byte chunk p*2 + (lane >> 4)
            const int r = warp * 16 + (lane & 15), c = p * 2 + (lane >> 4);
            sk_ldsm(sdt + r * (NP * 32) + ((c ^ (r & 1)) << 4), af[0], af[1], af[2], af[3]);
        }
#pragma unroll
        for (int np = 0; np < 8; ++np) {                       // two 8-column tiles per transposed x4 load
            uint32_t b0, b1, b2, b3;
            sk_ldsm_t(sa + p * (16 * TC * 2) + offsw<TC>((mat & 1) * 8 + (lane & 7), np * 2 + (mat >> 1)), b0, b1, b2, b3);
            float c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
            sk_mma(c1, af, b0, b1);
            sk_mma(c2, af, b2, b3);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float* cc = h ? c2 : c1;
                const int nt = np * 2 + h;
                const uint32_t col = static_cast<uint32_t>(c0 + nt * 8 + t4 * 2);
                const uint32_t w_lo = drop_word(a.key[p], r_lo, col, static_cast<uint32_t>(a.C));
                const uint32_t w_hi = drop_word(a.key[p], r_hi, col, static_cast<uint32_t>(a.C));
                if (drop_keep(w_lo, r_lo, 0u, a.t)) acc[nt][0] += cc[0];
                if (drop_keep(w_lo, r_lo, 1u, a.t)) acc[nt][1] += cc[1];
                if (drop_keep(w_hi, r_hi, 0u, a.t)) acc[nt][2] += cc[2];
                if (drop_keep(w_hi, r_hi, 1u, a.t)) acc[nt][3] += cc[3];
            }
        }
    }
    // through shared memory so that the read-modify-write of dx moves whole 16-byte pieces
    __syncthreads();                                               // every warp has read its operand fragments
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
        const int col = nt * 8 + t4 * 2;
        *reinterpret_cast<float2*>(s_acc + (warp * 16 + g) * (TC + 4) + col) = make_float2(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<float2*>(s_acc + (warp * 16 + g + 8) * (TC + 4) + col) = make_float2(acc[nt][2], acc[nt][3]);
    }
    __syncthreads();
    for (int idx = tid; idx < TR * (TC / 8); idx += SK_THREADS) {
        const int r = idx / (TC / 8), c = idx - r * (TC / 8);
        if (row0 + r >= a.M || c0 + c * 8 >= a.C) continue;
        bf16* ptr = a.dx + static_cast<long long>(row0 + r) * a.ldx + c0 + c * 8;
        uint4 v = *reinterpret_cast<const uint4*>(ptr);
        const float* s = s_acc + r * (TC + 4) + c * 8;
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            w[j] = pack_bf16(bf16_lo(w[j]) + a.inv * s[2 * j], bf16_hi(w[j]) + a.inv * s[2 * j + 1]);
        *reinterpret_cast<uint4*>(ptr) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

}  // namespace lhrs

using namespace lhrs;

extern "C" int lhrs_lora_dx_dropout(void* dx, int64_t ldx, int64_t M, int32_t C, const void* dT, int64_t ldt, const void* const* lora_a,
                                    int32_t nproj, uint64_t seed, int32_t module0, float p, void* stream) {
    LHRS_CHECK_ARG(dx && dT && lora_a && M > 0 && C > 0 && nproj >= 1 && nproj <= 3, "lhrs_lora_dx_dropout: null/empty");
    LHRS_CHECK_ARG(C % 8 == 0 && ldx % 8 == 0 && ldt % 8 == 0 && p > 0.f && p < 1.f && (reinterpret_cast<uintptr_t>(dx) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(dT) & 15) == 0,
                   "lhrs_lora_dx_dropout: C, ldx, ldt %% 8, 16-byte alignment and 0 < p < 1 required (rank 16)");
    DxDropArgs a;
    memset(&a, 0, sizeof(a));
    a.dx = (bf16*)dx; a.ldx = ldx; a.M = (int)M; a.C = C; a.dT = (const bf16*)dT; a.ldt = ldt;
    a.t = drop_threshold(p); a.inv = drop_inv_keep(a.t);
    for (int i = 0; i < nproj; ++i) {
        LHRS_CHECK_ARG(lora_a[i] != nullptr && (reinterpret_cast<uintptr_t>(lora_a[i]) & 15) == 0, "lhrs_lora_dx_dropout: lora_a[%d] null/unaligned", i);
        a.A[i] = (const bf16*)lora_a[i];
        a.key[i] = drop_key(seed, (uint32_t)(module0 + i));
    }
    if (a.t == 0) return LHRS_OK;
    const dim3 grid((C + 127) / 128, (unsigned)((M + 63) / 64));
    cudaStream_t st = (cudaStream_t)stream;
    const bool prof = prof_on();
    if (prof) prof_begin(PROF_SKINNY, 2.0 * M * (double)C * 16 * nproj, 4.0 * M * (double)C, st);
    if (nproj == 1) lora_dx_dropout_kernel<1><<<grid, SK_THREADS, 0, st>>>(a);
    else if (nproj == 2) lora_dx_dropout_kernel<2><<<grid, SK_THREADS, 0, st>>>(a);
    else lora_dx_dropout_kernel<3><<<grid, SK_THREADS, 0, st>>>(a);
    if (prof) prof_end(st);
    LHRS_LAUNCH_CHECK("lora_dx_dropout_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_lora_panel(const void* x, int64_t ldx, int64_t M, int32_t K, const void* const* w, int32_t nseg, int32_t w_kn,
                               int64_t ldw, int32_t n, float alpha, void* out, int64_t ldo, void* stream) {
    LHRS_CHECK_ARG(x && w && w[0] && out && M > 0 && K > 0, "lhrs_lora_panel: null/empty");
    LHRS_CHECK_ARG(n == 16 || n == 32 || n == 48, "lhrs_lora_panel: n=%d (supported: 16, 32, 48)", n);
    LHRS_CHECK_ARG(K % 64 == 0 && ldx % 8 == 0 && ldo % 2 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "lhrs_lora_panel: K %% 64, ldx %% 8 and 16-byte alignment required");
    PanelArgs a;
    memset(&a, 0, sizeof(a));
    a.X = (const bf16*)x; a.ldx = ldx; a.M = (int)M; a.K = K; a.alpha = alpha; a.out = (bf16*)out; a.ldo = ldo; a.ldw = ldw;
    cudaStream_t st = (cudaStream_t)stream;
    if (w_kn) {
        LHRS_CHECK_ARG(nseg >= 1 && nseg <= 3 && n == nseg * 16 && ldw == 16 && K % nseg == 0 && (K / nseg) % 64 == 0,
                       "lhrs_lora_panel: block-diagonal form needs r == 16, n == 16*nseg and K/nseg %% 64 == 0");
        for (int s = 0; s < nseg; ++s) { LHRS_CHECK_ARG(w[s] != nullptr, "lhrs_lora_panel: null segment"); a.W[s] = (const bf16*)w[s]; }
        a.kseg = K / nseg;
        if (n == 16) return launch_panel<2, true>(a, st);
        if (n == 32) return launch_panel<4, true>(a, st);
        return launch_panel<6, true>(a, st);
    }
    LHRS_CHECK_ARG(nseg == 1 && ldw % 8 == 0 && (reinterpret_cast<uintptr_t>(w[0]) & 15) == 0, "lhrs_lora_panel: K-major W must be one [n, ldw] matrix");
    a.W[0] = (const bf16*)w[0];
    if (n == 16) return launch_panel<2, false>(a, st);
    if (n == 32) return launch_panel<4, false>(a, st);
    return launch_panel<6, false>(a, st);
}

extern "C" int lhrs_lora_panel_dropout(const void* x, int64_t ldx, int64_t M, int32_t K, const void* w, int64_t ldw, int32_t n,
                                       float alpha, uint64_t seed, int32_t module0, float p, void* out, int64_t ldo, void* stream) {
    LHRS_CHECK_ARG(x && w && out && M > 0 && K > 0, "lhrs_lora_panel_dropout: null/empty");
    LHRS_CHECK_ARG(n == 16 || n == 32 || n == 48, "lhrs_lora_panel_dropout: n=%d (supported: 16, 32, 48)", n);
    LHRS_CHECK_ARG(K % 64 == 0 && ldx % 8 == 0 && ldo % 2 == 0 && ldw % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(w) & 15) == 0 && p >= 0.f && p < 1.f,
                   "lhrs_lora_panel_dropout: K %% 64, ldx / ldw %% 8, 16-byte alignment and 0 <= p < 1 required");
    PanelArgs a;
    memset(&a, 0, sizeof(a));
    a.X = (const bf16*)x; a.ldx = ldx; a.M = (int)M; a.K = K; a.out = (bf16*)out; a.ldo = ldo; a.ldw = ldw;
    a.W[0] = (const bf16*)w;
    a.drop_t = drop_threshold(p);
    a.alpha = alpha * drop_inv_keep(a.drop_t);
    for (int i = 0; i < n / 16; ++i) a.key[i] = drop_key(seed, (uint32_t)(module0 + i));
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 16) return launch_panel<2, false>(a, st);
    if (n == 32) return launch_panel<4, false>(a, st);
    return launch_panel<6, false>(a, st);
}

extern "C" size_t lhrs_lora_rowreduce_scratch_bytes(int64_t M, int32_t C, int32_t n) {
    (void)M;
    return (size_t)8 * (size_t)C * (size_t)n * sizeof(float);
}

static int rowreduce_impl(const void* p, int64_t ldp, int64_t M, int32_t C, const void* q, int64_t ldq, int32_t n, int32_t seg_c,
                          int32_t transpose, void* const* dst, int64_t ldd, float alpha, float* scratch, size_t scratch_bytes,
                          void* stream, int drop_t, const uint32_t* keys);

extern "C" int lhrs_lora_rowreduce(const void* p, int64_t ldp, int64_t M, int32_t C, const void* q, int64_t ldq, int32_t n, int32_t seg_c,
                                   int32_t transpose, void* const* dst, int64_t ldd, float alpha, float* scratch, size_t scratch_bytes,
                                   void* stream) {
    return rowreduce_impl(p, ldp, M, C, q, ldq, n, seg_c, transpose, dst, ldd, alpha, scratch, scratch_bytes, stream, 0, nullptr);
}

extern "C" int lhrs_lora_rowreduce_dropout(const void* p, int64_t ldp, int64_t M, int32_t C, const void* q, int64_t ldq, int32_t n,
                                           void* dst, int64_t ldd, uint64_t seed, int32_t module0, float p_drop, float* scratch,
                                           size_t scratch_bytes, void* stream) {
    LHRS_CHECK_ARG(p_drop >= 0.f && p_drop < 1.f && (n == 16 || n == 32 || n == 48), "lhrs_lora_rowreduce_dropout: bad p / n");
    const int t = drop_threshold(p_drop);
    uint32_t keys[3] = {0, 0, 0};
    for (int i = 0; i < n / 16; ++i) keys[i] = drop_key(seed, (uint32_t)(module0 + i));
    void* d[1] = {dst};
    return rowreduce_impl(p, ldp, M, C, q, ldq, n, 0, 1, d, ldd, drop_inv_keep(t), scratch, scratch_bytes, stream, t, keys);
}

static int rowreduce_impl(const void* p, int64_t ldp, int64_t M, int32_t C, const void* q, int64_t ldq, int32_t n, int32_t seg_c,
                          int32_t transpose, void* const* dst, int64_t ldd, float alpha, float* scratch, size_t scratch_bytes,
                          void* stream, int drop_t, const uint32_t* keys) {
    LHRS_CHECK_ARG(p && q && dst && dst[0] && scratch && M > 0 && C > 0, "lhrs_lora_rowreduce: null/empty");
    LHRS_CHECK_ARG(n == 16 || n == 32 || n == 48, "lhrs_lora_rowreduce: n=%d (supported: 16, 32, 48)", n);
    LHRS_CHECK_ARG(C % 8 == 0 && ldp % 8 == 0 && ldq % 8 == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0,
                   "lhrs_lora_rowreduce: C, ldp, ldq %% 8 and 16-byte alignment required");
    LHRS_CHECK_ARG(seg_c == 0 || (seg_c % 128 == 0 && C % seg_c == 0 && C / seg_c <= 3 && !transpose),
                   "lhrs_lora_rowreduce: block-diagonal form needs seg_c %% 128 == 0 and at most 3 segments");
    cudaStream_t st = (cudaStream_t)stream;
    const int bc = rowreduce_bc(C, seg_c);
    const int RR_BR = 8192 / bc;
    const int cblocks = (C + bc - 1) / bc;
    int nsplit = (2 * num_sms() + cblocks - 1) / cblocks;
    if (nsplit > 8) nsplit = 8;
    const long long max_by_rows = (M + RR_BR - 1) / RR_BR;
    if (nsplit > max_by_rows) nsplit = (int)max_by_rows;
    if (nsplit < 1) nsplit = 1;
    // Measured (tools/lora_bench.py, M = 8192): the cluster reduction wins 2 us on the [out, 16] block-diagonal form (dB) and loses
    // 2-10 us on the transposed [n, in] form (dA: 8-block clusters of 90 KB CTAs schedule worse than independent CTAs), so the
    // default takes it for the former only.  LHRS_SKINNY_CLUSTER = 0 / 1 forces it off / on for both.
    static int cluster_env = -2;
    if (cluster_env == -2) cluster_env = env_int("LHRS_SKINNY_CLUSTER", -1);
    const int use_cluster = cluster_env >= 0 ? cluster_env : (transpose ? 0 : 1);
    LHRS_CHECK_ARG(use_cluster || scratch_bytes >= (size_t)nsplit * C * n * sizeof(float), "lhrs_lora_rowreduce: scratch too small");
    ReduceArgs a;
    memset(&a, 0, sizeof(a));
    const int nseg_dst = seg_c > 0 ? C / seg_c : 1;
    a.cluster = use_cluster; a.transpose = transpose; a.ldd = ldd; a.alpha = alpha;
    a.drop_t = drop_t;
    if (drop_t > 0 && keys != nullptr) { a.key[0] = keys[0]; a.key[1] = keys[1]; a.key[2] = keys[2]; }
    for (int sidx = 0; sidx < nseg_dst && sidx < 3; ++sidx) a.dst[sidx] = (bf16*)dst[sidx];
    a.P = (const bf16*)p; a.ldp = ldp; a.M = (int)M; a.C = C; a.Q = (const bf16*)q; a.ldq = ldq; a.seg_c = seg_c; a.partial = scratch;
    long long rps = (M + nsplit - 1) / nsplit;
    rps = ((rps + RR_BR - 1) / RR_BR) * RR_BR;
    a.rows_per_split = (int)rps;
    nsplit = (int)((M + rps - 1) / rps);
    int rc;
    if (n == 16) rc = launch_rowreduce<2>(a, nsplit, bc, st);
    else if (n == 32) rc = launch_rowreduce<4>(a, nsplit, bc, st);
    else rc = launch_rowreduce<6>(a, nsplit, bc, st);
    if (rc) return rc;
    if (use_cluster) return LHRS_OK;
    const long long total = (long long)C * n;
    const int nseg = seg_c > 0 ? C / seg_c : 1;
    lora_reduce_store_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(scratch, nsplit, C, n, transpose, seg_c, (bf16*)dst[0],
                                                                            nseg > 1 ? (bf16*)dst[1] : nullptr, nseg > 2 ? (bf16*)dst[2] : nullptr,
                                                                            ldd, alpha);
    LHRS_LAUNCH_CHECK("lora_reduce_store_kernel");
    return LHRS_OK;
}
