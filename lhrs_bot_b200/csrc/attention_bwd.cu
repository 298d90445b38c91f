// Backward of the fused attention (autograd through HF LlamaAttention / nn.MultiheadAttention).
// Recompute-based ("flash") backward from Q, K, V, O, dO and the saved log-sum-exp; nothing S x S touches HBM.
//   delta  = rowsum(dO * O)
//   dQ     : one CTA per 64-query block, loops over key blocks          (dS K)
//   dK, dV : one CTA per 64-key block, loops over query blocks          (dS^T Q, P^T dO)
// Two passes recompute S twice but need no atomics, so gradients are deterministic.  Warp-level mma.sync tiles as in
// attention.cu (attention is a few % of the step; the dense dX/dW contractions around it are tcgen05).
#include "attention_common.h"
#include <stdlib.h>
#include "host_common.h"
#include "ptx.cuh"

namespace lhrs {


// inverse rotate-half on one thread's fragment: tiles i (dims 8i+2t, +1) and i+8 (dims +64) of the same row pair up
template <int DTILES>
__device__ __forceinline__ void unrope_frag(float (&g)[DTILES][4], const float* cosT, const float* sinT, int row_a, int row_b, int t4,
                                            int nrows) {
    if constexpr (DTILES == 16) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int j = i * 8 + t4 * 2;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = r == 0 ? row_a : row_b;
                if (row >= nrows) continue;
                const float2 c = *reinterpret_cast<const float2*>(cosT + static_cast<long long>(row) * 64 + j);
                const float2 s = *reinterpret_cast<const float2*>(sinT + static_cast<long long>(row) * 64 + j);
                const float y1a = g[i][r * 2], y1b = g[i][r * 2 + 1], y2a = g[i + 8][r * 2], y2b = g[i + 8][r * 2 + 1];
                g[i][r * 2] = y1a * c.x + y2a * s.x;          // dx1 = dy1 c + dy2 s
                g[i][r * 2 + 1] = y1b * c.y + y2b * s.y;
                g[i + 8][r * 2] = y2a * c.x - y1a * s.x;      // dx2 = dy2 c - dy1 s
                g[i + 8][r * 2 + 1] = y2b * c.y - y1b * s.y;
            }
        }
    }
}

__device__ __forceinline__ void cp_async16b(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int HD>
__device__ __forceinline__ uint32_t toff(int row, int chunk) {
    return static_cast<uint32_t>(row * HD * 2 + (((chunk & ~7) | ((chunk & 7) ^ (row & 7))) << 4));
}
template <int HD, int ROWS, int THREADS>
__device__ __forceinline__ void load_rows(uint32_t smem_base, const __nv_bfloat16* gptr, long long row_stride, int row0,
                                          int nrows_total, int tid) {
    constexpr int CHUNKS = HD / 8;
    constexpr int TOTAL = ROWS * CHUNKS;
#pragma unroll
    for (int i = 0; i < (TOTAL + THREADS - 1) / THREADS; ++i) {
        const int idx = tid + i * THREADS;
        if (idx < TOTAL) {
            const int r = idx / CHUNKS, c = idx % CHUNKS;
            const int gr = row0 + r;
            const bool ok = gr < nrows_total;
            cp_async16b(smem_base + toff<HD>(r, c), gptr + static_cast<long long>(ok ? gr : 0) * row_stride + c * 8, ok);
        }
    }
}

// ------------------------------------------------------------------ delta = rowsum(dO * O), layout [B,H,Sq]
// up to three problems that share (B, H) in one launch (the AttnPooler's query groups): block index = group * B + batch
struct AttnBwdArgsG {
    AttnBwdArgs a[3];
    int n;
};

template <int HD>
__global__ void __launch_bounds__(256) attn_delta_kernel(const AttnBwdArgsG G) {
    const int gi = static_cast<int>(blockIdx.y) / G.a[0].B;
    const AttnBwdArgs p = (gi == 0) ? G.a[0] : (gi == 1 ? G.a[1] : G.a[2]);
    if (static_cast<int>(blockIdx.x) >= p.Sq) return;
    const int s = blockIdx.x, b = static_cast<int>(blockIdx.y) - gi * p.B;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long row = s;                                   // ragged batch: row s of sequence b is row seq_off[b] + s (o_bs is 0)
    if (p.seq_off != nullptr) {
        const int rb = p.seq_off[b];
        if (s >= p.seq_off[b + 1] - rb) return;
        row += rb;
    }
    for (int h = warp; h < p.H; h += 8) {
        const __nv_bfloat16* o = p.o + b * p.o_bs + row * p.o_rs + h * p.o_hs;
        const __nv_bfloat16* d = p.d_o + b * p.o_bs + row * p.o_rs + h * p.o_hs;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < HD / 64; ++i) {
            const uint32_t ov = *reinterpret_cast<const uint32_t*>(o + (lane + i * 32) * 2);
            const uint32_t dv = *reinterpret_cast<const uint32_t*>(d + (lane + i * 32) * 2);
            acc += bf16_lo(ov) * bf16_lo(dv) + bf16_hi(ov) * bf16_hi(dv);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) p.delta[(static_cast<long long>(b) * p.H + h) * p.Sq + s] = acc;
    }
    // the same pre-pass packs the key-padding mask to bits for the tcgen05 kernels: block (s, b) writes word s of batch row b
    if (p.kbits != nullptr && s < p.kbits_w && warp == 0) {
        const int key = s * 32 + lane;
        const bool on = key < p.Skv && p.kmask[static_cast<long long>(b) * p.Skv + key] != 0;
        const uint32_t bits = __ballot_sync(0xffffffffu, on);
        if (lane == 0) p.kbits[static_cast<long long>(b) * p.kbits_w + s] = bits;
    }
}

// ------------------------------------------------------------------ dQ
template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const AttnBwdArgsG G) {
    constexpr int BQ = 64, BKV = 64, THREADS = 128, KSTEPS = HD / 16, DTILES = HD / 8;
    const int gi = static_cast<int>(blockIdx.z) / G.a[0].B;
    const AttnBwdArgs p = (gi == 0) ? G.a[0] : (gi == 1 ? G.a[1] : G.a[2]);
    if (static_cast<int>(blockIdx.x) * BQ >= p.Sq) return;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sdO = sQ + BQ * HD * 2;
    const uint32_t sK = sdO + BQ * HD * 2;
    const uint32_t sV = sK + 2 * BKV * HD * 2;
    __shared__ uint8_t sMask[2][BKV];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int qb = blockIdx.x, h = blockIdx.y, b = static_cast<int>(blockIdx.z) - gi * p.B;
    const int q0 = qb * BQ;
    const __nv_bfloat16* qg = p.q + b * p.q_bs + h * p.q_hs;
    const __nv_bfloat16* kg = p.k + b * p.k_bs + h * p.k_hs;
    const __nv_bfloat16* vg = p.v + b * p.v_bs + h * p.v_hs;
    const __nv_bfloat16* dog = p.d_o + b * p.o_bs + h * p.o_hs;
    const uint8_t* km = p.kmask ? p.kmask + static_cast<long long>(b) * p.Skv : nullptr;
    const int shift = p.Skv - p.Sq;
    int kv_end = p.Skv;
    if (CAUSAL) kv_end = min(p.Skv, q0 + BQ + shift);
    const int nblk = (kv_end + BKV - 1) / BKV;

    load_rows<HD, BQ, THREADS>(sQ, qg, p.q_rs, q0, p.Sq, tid);
    load_rows<HD, BQ, THREADS>(sdO, dog, p.o_rs, q0, p.Sq, tid);
    load_rows<HD, BKV, THREADS>(sK, kg, p.k_rs, 0, p.Skv, tid);
    load_rows<HD, BKV, THREADS>(sV, vg, p.v_rs, 0, p.Skv, tid);
    cp_commit();
    if (km != nullptr && tid < BKV) sMask[0][tid] = (tid < p.Skv) ? km[tid] : 0;

    const int qrow0 = q0 + warp * 16 + g;
    float lse2[2], dl[2];
    {
        const long long base = (static_cast<long long>(b) * p.H + h) * p.Sq;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int qr = qrow0 + r * 8;
            const float l = (qr < p.Sq) ? p.lse[base + qr] : INFINITY;
            lse2[r] = (l == -INFINITY) ? INFINITY : l * 1.4426950408889634f;  // fully masked row -> P = 0
            dl[r] = (qr < p.Sq) ? p.delta[base + qr] : 0.f;
        }
    }
    float dq[DTILES][4];
#pragma unroll
    for (int i = 0; i < DTILES; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }

    for (int j = 0; j < nblk; ++j) {
        const int buf = j & 1;
        if (j + 1 < nblk) {
            load_rows<HD, BKV, THREADS>(sK + (buf ^ 1) * BKV * HD * 2, kg, p.k_rs, (j + 1) * BKV, p.Skv, tid);
            load_rows<HD, BKV, THREADS>(sV + (buf ^ 1) * BKV * HD * 2, vg, p.v_rs, (j + 1) * BKV, p.Skv, tid);
            cp_commit();
            if (km != nullptr && tid < BKV) {
                const int key = (j + 1) * BKV + tid;
                sMask[buf ^ 1][tid] = (key < p.Skv) ? km[key] : 0;
            }
            cp_wait<1>();
        } else {
            cp_wait<0>();
        }
        __syncthreads();
        const uint32_t kb = sK + buf * BKV * HD * 2, vb = sV + buf * BKV * HD * 2;

        float s[BKV / 8][4], dp[BKV / 8][4];
#pragma unroll
        for (int i = 0; i < BKV / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            uint32_t qa[4], da[4];
            ldsm4(sQ + toff<HD>(warp * 16 + (lane & 15), ks * 2 + (lane >> 4)), qa[0], qa[1], qa[2], qa[3]);
            ldsm4(sdO + toff<HD>(warp * 16 + (lane & 15), ks * 2 + (lane >> 4)), da[0], da[1], da[2], da[3]);
#pragma unroll
            for (int np = 0; np < BKV / 16; ++np) {
                const int mat = lane >> 3;
                const int key = np * 16 + (mat >> 1) * 8 + (lane & 7);
                const int c = ks * 2 + (mat & 1);
                uint32_t b0, b1, b2, b3;
                ldsm4(kb + toff<HD>(key, c), b0, b1, b2, b3);
                mma16816(s[np * 2], qa, b0, b1);
                mma16816(s[np * 2 + 1], qa, b2, b3);
                ldsm4(vb + toff<HD>(key, c), b0, b1, b2, b3);
                mma16816(dp[np * 2], da, b0, b1);
                mma16816(dp[np * 2 + 1], da, b2, b3);
            }
        }
        // P = exp(S*scale - lse); dS = P * (dP - delta) * scale  -> bf16 A fragments
        uint32_t dsf[BKV / 16][4];
        const int kbase = j * BKV;
#pragma unroll
        for (int i = 0; i < BKV / 8; ++i) {
            float dsv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int kl = i * 8 + t4 * 2 + (e & 1);
                const int key = kbase + kl;
                const int r = e >> 1;
                const int qr = qrow0 + r * 8;
                bool ok = key < p.Skv;
                if (CAUSAL) ok = ok && (key <= qr + shift);
                if (km != nullptr) ok = ok && (sMask[buf][kl] != 0);
                const float pv = ok ? exp2f(s[i][e] * p.scale_log2 - lse2[r]) : 0.f;
                dsv[e] = pv * (dp[i][e] - dl[r]) * p.scale;
            }
            dsf[i >> 1][(i & 1) * 2 + 0] = pack_bf16(dsv[0], dsv[1]);
            dsf[i >> 1][(i & 1) * 2 + 1] = pack_bf16(dsv[2], dsv[3]);
        }
        // dQ += dS K
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
#pragma unroll
            for (int dpi = 0; dpi < DTILES / 2; ++dpi) {
                const int mat = lane >> 3;
                const int key = kk * 16 + (mat & 1) * 8 + (lane & 7);
                const int c = dpi * 2 + (mat >> 1);
                uint32_t b0, b1, b2, b3;
                ldsm4t(kb + toff<HD>(key, c), b0, b1, b2, b3);
                mma16816(dq[dpi * 2], dsf[kk], b0, b1);
                mma16816(dq[dpi * 2 + 1], dsf[kk], b2, b3);
            }
        }
        __syncthreads();
    }
    if (p.rope_cos != nullptr) unrope_frag<DTILES>(dq, p.rope_cos, p.rope_sin, qrow0, qrow0 + 8, t4, p.Sq);
    __nv_bfloat16* dqg = p.dq + b * p.dq_bs + h * p.dq_hs;
#pragma unroll
    for (int i = 0; i < DTILES; ++i) {
        const int d = i * 8 + t4 * 2;
        if (qrow0 < p.Sq) *reinterpret_cast<uint32_t*>(dqg + static_cast<long long>(qrow0) * p.dq_rs + d) = pack_bf16(dq[i][0], dq[i][1]);
        if (qrow0 + 8 < p.Sq) *reinterpret_cast<uint32_t*>(dqg + static_cast<long long>(qrow0 + 8) * p.dq_rs + d) = pack_bf16(dq[i][2], dq[i][3]);
    }
}

// ------------------------------------------------------------------ dK, dV
template <int HD, int BQ, bool CAUSAL>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const AttnBwdArgsG G) {
    constexpr int BKV = 64, THREADS = 128, KSTEPS = HD / 16, DTILES = HD / 8;
    const int gi = static_cast<int>(blockIdx.z) / G.a[0].B;
    const AttnBwdArgs p = (gi == 0) ? G.a[0] : (gi == 1 ? G.a[1] : G.a[2]);
    if (static_cast<int>(blockIdx.x) * BKV >= p.Skv) return;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sK = smem_u32(smem);
    const uint32_t sV = sK + BKV * HD * 2;
    const uint32_t sQ = sV + BKV * HD * 2;            // 2 buffers of BQ rows
    const uint32_t sdO = sQ + 2 * BQ * HD * 2;        // 2 buffers of BQ rows
    float* sStat = reinterpret_cast<float*>(smem + (2 * BKV + 4 * BQ) * HD * 2);  // [2 buf][2 (lse2, delta)][BQ]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int kblk = blockIdx.x, h = blockIdx.y, b = static_cast<int>(blockIdx.z) - gi * p.B;
    const int k0 = kblk * BKV;
    const __nv_bfloat16* qg = p.q + b * p.q_bs + h * p.q_hs;
    const __nv_bfloat16* kg = p.k + b * p.k_bs + h * p.k_hs;
    const __nv_bfloat16* vg = p.v + b * p.v_bs + h * p.v_hs;
    const __nv_bfloat16* dog = p.d_o + b * p.o_bs + h * p.o_hs;
    const long long stat_base = (static_cast<long long>(b) * p.H + h) * p.Sq;
    const int shift = p.Skv - p.Sq;
    // first query block that can see key k0:  q + shift >= k0
    int qstart = 0;
    if (CAUSAL) qstart = max(0, (k0 - shift) / BQ);
    const int nq = (p.Sq + BQ - 1) / BQ;

    auto load_q = [&](int qi, int buf) {
        load_rows<HD, BQ, THREADS>(sQ + buf * BQ * HD * 2, qg, p.q_rs, qi * BQ, p.Sq, tid);
        load_rows<HD, BQ, THREADS>(sdO + buf * BQ * HD * 2, dog, p.o_rs, qi * BQ, p.Sq, tid);
        if (tid < BQ) {
            const int qr = qi * BQ + tid;
            float l = (qr < p.Sq) ? p.lse[stat_base + qr] : INFINITY;
            l = (l == -INFINITY) ? INFINITY : l * 1.4426950408889634f;
            sStat[(buf * 2 + 0) * BQ + tid] = l;
            sStat[(buf * 2 + 1) * BQ + tid] = (qr < p.Sq) ? p.delta[stat_base + qr] : 0.f;
        }
    };

    load_rows<HD, BKV, THREADS>(sK, kg, p.k_rs, k0, p.Skv, tid);
    load_rows<HD, BKV, THREADS>(sV, vg, p.v_rs, k0, p.Skv, tid);
    if (qstart < nq) load_q(qstart, 0);
    cp_commit();

    const int krow0 = k0 + warp * 16 + g;  // this thread's key rows: krow0, krow0 + 8
    bool key_ok[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int key = krow0 + r * 8;
        key_ok[r] = key < p.Skv;
        if (p.kmask != nullptr && key_ok[r]) key_ok[r] = p.kmask[static_cast<long long>(b) * p.Skv + key] != 0;
    }
    float dk[DTILES][4], dv[DTILES][4];
#pragma unroll
    for (int i = 0; i < DTILES; ++i) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }

    for (int qi = qstart, it = 0; qi < nq; ++qi, ++it) {
        const int buf = it & 1;
        if (qi + 1 < nq) {
            load_q(qi + 1, buf ^ 1);
            cp_commit();
            cp_wait<1>();
        } else {
            cp_wait<0>();
        }
        __syncthreads();
        const uint32_t qb = sQ + buf * BQ * HD * 2, dob = sdO + buf * BQ * HD * 2;
        const float* lse2 = sStat + (buf * 2 + 0) * BQ;
        const float* dl = sStat + (buf * 2 + 1) * BQ;

        // S^T = K Q^T and dP^T = V dO^T   (16 keys x BQ queries per warp)
        float st[BQ / 8][4], dpt[BQ / 8][4];
#pragma unroll
        for (int i = 0; i < BQ / 8; ++i) { st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f; dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            uint32_t ka[4], va[4];
            ldsm4(sK + toff<HD>(warp * 16 + (lane & 15), ks * 2 + (lane >> 4)), ka[0], ka[1], ka[2], ka[3]);
            ldsm4(sV + toff<HD>(warp * 16 + (lane & 15), ks * 2 + (lane >> 4)), va[0], va[1], va[2], va[3]);
#pragma unroll
            for (int np = 0; np < BQ / 16; ++np) {
                const int mat = lane >> 3;
                const int qr = np * 16 + (mat >> 1) * 8 + (lane & 7);
                const int c = ks * 2 + (mat & 1);
                uint32_t b0, b1, b2, b3;
                ldsm4(qb + toff<HD>(qr, c), b0, b1, b2, b3);
                mma16816(st[np * 2], ka, b0, b1);
                mma16816(st[np * 2 + 1], ka, b2, b3);
                ldsm4(dob + toff<HD>(qr, c), b0, b1, b2, b3);
                mma16816(dpt[np * 2], va, b0, b1);
                mma16816(dpt[np * 2 + 1], va, b2, b3);
            }
        }
        uint32_t pf[BQ / 16][4], dsf[BQ / 16][4];
        const int qbase = qi * BQ;
#pragma unroll
        for (int i = 0; i < BQ / 8; ++i) {
            float pv[4], dsv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int ql = i * 8 + t4 * 2 + (e & 1);   // query (column) within the block
                const int r = e >> 1;                       // key row g / g+8
                const int key = krow0 + r * 8;
                bool ok = key_ok[r];
                if (CAUSAL) ok = ok && (key <= qbase + ql + shift);
                pv[e] = ok ? exp2f(st[i][e] * p.scale_log2 - lse2[ql]) : 0.f;
                dsv[e] = pv[e] * (dpt[i][e] - dl[ql]) * p.scale;
            }
            pf[i >> 1][(i & 1) * 2 + 0] = pack_bf16(pv[0], pv[1]);
            pf[i >> 1][(i & 1) * 2 + 1] = pack_bf16(pv[2], pv[3]);
            dsf[i >> 1][(i & 1) * 2 + 0] = pack_bf16(dsv[0], dsv[1]);
            dsf[i >> 1][(i & 1) * 2 + 1] = pack_bf16(dsv[2], dsv[3]);
        }
        // dV += P^T dO ; dK += dS^T Q      (reduction over the BQ queries)
#pragma unroll
        for (int kk = 0; kk < BQ / 16; ++kk) {
#pragma unroll
            for (int dpi = 0; dpi < DTILES / 2; ++dpi) {
                const int mat = lane >> 3;
                const int qr = kk * 16 + (mat & 1) * 8 + (lane & 7);
                const int c = dpi * 2 + (mat >> 1);
                uint32_t b0, b1, b2, b3;
                ldsm4t(dob + toff<HD>(qr, c), b0, b1, b2, b3);
                mma16816(dv[dpi * 2], pf[kk], b0, b1);
                mma16816(dv[dpi * 2 + 1], pf[kk], b2, b3);
                ldsm4t(qb + toff<HD>(qr, c), b0, b1, b2, b3);
                mma16816(dk[dpi * 2], dsf[kk], b0, b1);
                mma16816(dk[dpi * 2 + 1], dsf[kk], b2, b3);
            }
        }
        __syncthreads();
    }
    if (p.rope_cos != nullptr) unrope_frag<DTILES>(dk, p.rope_cos, p.rope_sin, krow0, krow0 + 8, t4, p.Skv);
    __nv_bfloat16* dkg = p.dk + b * p.dk_bs + h * p.dk_hs;
    __nv_bfloat16* dvg = p.dv + b * p.dv_bs + h * p.dv_hs;
#pragma unroll
    for (int i = 0; i < DTILES; ++i) {
        const int d = i * 8 + t4 * 2;
        if (krow0 < p.Skv) {
            *reinterpret_cast<uint32_t*>(dkg + static_cast<long long>(krow0) * p.dk_rs + d) = pack_bf16(dk[i][0], dk[i][1]);
            *reinterpret_cast<uint32_t*>(dvg + static_cast<long long>(krow0) * p.dv_rs + d) = pack_bf16(dv[i][0], dv[i][1]);
        }
        if (krow0 + 8 < p.Skv) {
            *reinterpret_cast<uint32_t*>(dkg + static_cast<long long>(krow0 + 8) * p.dk_rs + d) = pack_bf16(dk[i][2], dk[i][3]);
            *reinterpret_cast<uint32_t*>(dvg + static_cast<long long>(krow0 + 8) * p.dv_rs + d) = pack_bf16(dv[i][2], dv[i][3]);
        }
    }
}

template <int HD, bool CAUSAL>
static int launch_bwd(const AttnBwdArgsG& G, cudaStream_t stream) {
    const AttnBwdArgs& a = G.a[0];
    constexpr int BQ2 = (HD == 128) ? 32 : 64;
    constexpr int SMEM_DQ = (2 * 64 + 4 * 64) * HD * 2;
    constexpr int SMEM_DKV = (2 * 64 + 4 * BQ2) * HD * 2 + 4 * BQ2 * 4;
    auto kd = attn_delta_kernel<HD>;
    auto kq = attn_bwd_dq_kernel<HD, CAUSAL>;
    auto kkv = attn_bwd_dkv_kernel<HD, BQ2, CAUSAL>;
    static bool attr_set = false;
    if (!attr_set) {
        LHRS_CUDA(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DQ));
        LHRS_CUDA(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DKV));
        attr_set = true;
    }
    const bool prof = prof_on();
    int max_sq = 0, max_skv = 0;
    double pairs = 0;
    for (int i = 0; i < G.n; ++i) {
        max_sq = G.a[i].Sq > max_sq ? G.a[i].Sq : max_sq;
        max_skv = G.a[i].Skv > max_skv ? G.a[i].Skv : max_skv;
        pairs += CAUSAL ? 0.5 * G.a[i].Sq * (double)G.a[i].Skv : (double)G.a[i].Sq * G.a[i].Skv;
    }
    if (prof) prof_begin(PROF_ATTN, 10.0 * a.B * a.H * pairs * HD, 0.0, stream);  // 5 contractions of 2*pairs*HD (algorithmic)
    kd<<<dim3(max_sq, a.B * G.n), 256, 0, stream>>>(G);
    LHRS_LAUNCH_CHECK("attn_delta_kernel");
    bool tma_ok = true;   // TMA views and 16-byte stores need 8-element strides
    for (long long s : {a.q_bs, a.q_rs, a.q_hs, a.k_bs, a.k_rs, a.k_hs, a.v_bs, a.v_rs, a.v_hs, a.o_bs, a.o_rs, a.o_hs, a.dq_bs, a.dq_rs,
                        a.dq_hs, a.dk_bs, a.dk_rs, a.dk_hs, a.dv_bs, a.dv_rs, a.dv_hs})
        tma_ok = tma_ok && (s % 8) == 0;
    if (G.n == 1 && HD == 128 && a.Sq >= 128 && a.Skv >= 128 && tma_ok && use_tc_attention()) {   // tcgen05 dQ and dK/dV kernels
        const char* pe = getenv("LHRS_ATTN_BWD_PERSIST");   // 0: one CTA per tile (attention_bwd_tc.cu); read per call for A/B runs
        const bool persist = (pe == nullptr || atoi(pe) != 0) && attention_bwd_tcp_ok(a) && (a.kmask == nullptr || a.kbits != nullptr);
        if (a.seq_off != nullptr && !persist) {
            set_error("lhrs_attention_bwd: a ragged batch (seq_off) runs on the persistent tcgen05 kernels only");
            return LHRS_ERR_INVALID;
        }
        const int rc = persist ? attention_bwd_tcp(a, CAUSAL, stream) : attention_bwd_tc(a, CAUSAL, stream);
        if (prof) prof_end(stream);
        return rc;
    }
    if (a.seq_off != nullptr) {
        set_error("lhrs_attention_bwd: a ragged batch (seq_off) needs the tcgen05 path (head_dim 128, Sq >= 128, 8-element strides)");
        return LHRS_ERR_INVALID;
    }
    kq<<<dim3((max_sq + 63) / 64, a.H, a.B * G.n), 128, SMEM_DQ, stream>>>(G);
    LHRS_LAUNCH_CHECK("attn_bwd_dq_kernel");
    kkv<<<dim3((max_skv + 63) / 64, a.H, a.B * G.n), 128, SMEM_DKV, stream>>>(G);
    if (prof) prof_end(stream);
    LHRS_LAUNCH_CHECK("attn_bwd_dkv_kernel");
    return LHRS_OK;
}

}  // namespace lhrs

using namespace lhrs;

extern "C" int64_t lhrs_attention_bwd_scratch_floats(int32_t B, int32_t H, int32_t Sq, int32_t Skv) {
    return static_cast<int64_t>(B) * H * Sq + static_cast<int64_t>(B) * 2 * ((Skv + 63) / 64);
}

static int fill_bwd_args(const LhrsAttentionBwd* d, AttnBwdArgs& a) {
    LHRS_CHECK_ARG(d != nullptr, "lhrs_attention_bwd: null descriptor");
    const LhrsAttention* f = &d->fwd;
    LHRS_CHECK_ARG(f->q && f->k && f->v && f->o && f->lse && d->d_o && d->dq && d->dk && d->dv && d->delta,
                   "lhrs_attention_bwd: null operand (lse and delta scratch are required)");
    LHRS_CHECK_ARG(f->head_dim == 64 || f->head_dim == 128, "lhrs_attention_bwd: head_dim %d", f->head_dim);
    const long long strides[] = {f->q_bs, f->q_rs, f->q_hs, f->k_bs, f->k_rs, f->k_hs, f->v_bs, f->v_rs, f->v_hs, f->o_bs, f->o_rs, f->o_hs,
                                 d->dq_bs, d->dq_rs, d->dq_hs, d->dk_bs, d->dk_rs, d->dk_hs, d->dv_bs, d->dv_rs, d->dv_hs};
    for (long long s : strides) LHRS_CHECK_ARG((s % 2) == 0, "lhrs_attention_bwd: odd stride");
    a.q = (const __nv_bfloat16*)f->q; a.k = (const __nv_bfloat16*)f->k; a.v = (const __nv_bfloat16*)f->v;
    a.o = (const __nv_bfloat16*)f->o; a.d_o = (const __nv_bfloat16*)d->d_o;
    a.dq = (__nv_bfloat16*)d->dq; a.dk = (__nv_bfloat16*)d->dk; a.dv = (__nv_bfloat16*)d->dv;
    a.lse = f->lse; a.delta = d->delta; a.kmask = f->key_mask;
    // mask bits live behind the delta values in the caller's scratch (lhrs_attention_bwd_scratch_floats)
    a.kbits_w = 2 * ((f->Skv + 63) / 64);
    a.kbits = (f->key_mask != nullptr && f->head_dim == 128 && a.kbits_w <= f->Sq)
                  ? reinterpret_cast<uint32_t*>(d->delta + static_cast<long long>(f->B) * f->H * f->Sq) : nullptr;
    a.q_bs = f->q_bs; a.q_rs = f->q_rs; a.q_hs = f->q_hs; a.k_bs = f->k_bs; a.k_rs = f->k_rs; a.k_hs = f->k_hs;
    a.v_bs = f->v_bs; a.v_rs = f->v_rs; a.v_hs = f->v_hs; a.o_bs = f->o_bs; a.o_rs = f->o_rs; a.o_hs = f->o_hs;
    a.dq_bs = d->dq_bs; a.dq_rs = d->dq_rs; a.dq_hs = d->dq_hs; a.dk_bs = d->dk_bs; a.dk_rs = d->dk_rs; a.dk_hs = d->dk_hs;
    a.dv_bs = d->dv_bs; a.dv_rs = d->dv_rs; a.dv_hs = d->dv_hs;
    a.B = f->B; a.H = f->H; a.Sq = f->Sq; a.Skv = f->Skv;
    a.scale = f->scale; a.scale_log2 = f->scale * 1.4426950408889634f;
    a.rope_cos = d->rope_cos; a.rope_sin = d->rope_sin;
    a.seq_off = f->seq_off; a.total_rows = f->total_rows;
    if (a.seq_off != nullptr) {
        LHRS_CHECK_ARG(f->head_dim == 128 && f->causal && f->Sq == f->Skv && f->Sq >= 128 && f->key_mask == nullptr && f->total_rows > 0 &&
                           f->total_rows < (1ll << 31),
                       "lhrs_attention_bwd: seq_off needs head_dim 128, causal, Sq == Skv >= 128, no key_mask, total_rows");
        a.q_bs = a.k_bs = a.v_bs = a.o_bs = a.dq_bs = a.dk_bs = a.dv_bs = 0;     // one row space for the whole batch
    }
    LHRS_CHECK_ARG(a.rope_cos == nullptr || (f->head_dim == 128 && a.rope_sin != nullptr && f->Sq == f->Skv),
                   "lhrs_attention_bwd: fused un-RoPE needs head_dim 128, both tables and Sq == Skv");
    return LHRS_OK;
}

extern "C" int lhrs_attention_bwd(const LhrsAttentionBwd* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    AttnBwdArgsG G;
    G.n = 1;
    const int rc = fill_bwd_args(d, G.a[0]);
    if (rc) return rc;
    G.a[1] = G.a[0]; G.a[2] = G.a[0];
    const LhrsAttention* f = &d->fwd;
    if (f->head_dim == 128) return f->causal ? launch_bwd<128, true>(G, stream) : launch_bwd<128, false>(G, stream);
    return f->causal ? launch_bwd<64, true>(G, stream) : launch_bwd<64, false>(G, stream);
}

// n (<= 3) backward problems with the same B, H, head_dim 64 and causal flag: delta, dQ and dK/dV each as ONE launch
extern "C" int lhrs_attention_bwd_grouped(const LhrsAttentionBwd* d, int32_t n, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    LHRS_CHECK_ARG(d != nullptr && n >= 1 && n <= 3, "lhrs_attention_bwd_grouped: 1..3 problems");
    AttnBwdArgsG G;
    G.n = n;
    bool same = true;
    for (int i = 0; i < n; ++i) {
        const int rc = fill_bwd_args(&d[i], G.a[i]);
        if (rc) return rc;
        same = same && d[i].fwd.B == d[0].fwd.B && d[i].fwd.H == d[0].fwd.H && d[i].fwd.head_dim == 64 && d[i].fwd.causal == d[0].fwd.causal;
    }
    if (!same || n == 1) {
        for (int i = 0; i < n; ++i) { const int rc = lhrs_attention_bwd(&d[i], stream_); if (rc) return rc; }
        return LHRS_OK;
    }
    for (int i = n; i < 3; ++i) G.a[i] = G.a[0];
    return d[0].fwd.causal ? launch_bwd<64, true>(G, stream) : launch_bwd<64, false>(G, stream);
}
