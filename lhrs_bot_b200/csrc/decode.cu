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
// The GEMVs are CTA-cooperative (gemv_coop_kernel: a CTA owns a contiguous range of rows, all 256 threads walk each row, two
// register sets of loads in flight, 3 CTAs per SM; the input vector arrives by cp.async together with the "finished" flag) and
// the attention's four context slices per head are a thread-block cluster (attn_decode2_kernel): 322 -> 392 tok/s at 7B over the
// round-1 chain (warp-per-row gemv_kernel, ticket-merged attn_decode_kernel; kept behind LHRS_GEMV_COOP=0 / LHRS_DECODE_ATTN2=0),
// i.e. 5.3 TB/s of weight streaming = 0.80 of the measured HBM copy rate.
// Every kernel is launched with programmatic dependent launch: before its dependency wait it starts streaming its first weight
// rows (registers) and prefetches more towards L2, so e.g. the o-projection's weights arrive while the tiny attention runs.
// (Measured alternative, round 1: ONE persistent cooperative kernel per token with grid barriers between the phases was
// slower - 5.7 ms vs 3.2 ms per token - because 560 of 592 blocks idle through each attention phase and the row quantisation
// of every phase lands on a barrier; the dependent-launch chain overlaps exactly those tails.)
#include <stdlib.h>

#include <cooperative_groups.h>

#include "host_common.h"
#include "ptx.cuh"
#include "models_common.h"
#include "sampling.cuh"

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
    int pf_bytes_per_warp;        // L2 prefetch budget of each warp's upcoming rows, issued before the dependency wait
    const int* done;              // nullable: device flag (decode state[3]); non-zero = the sequence has finished, skip the work
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
// Programmatic dependent launch: every decode kernel is launched with the stream-serialization attribute, lets its
// successor start early (launch_dependents) and waits for its predecessor's results only where it first needs them.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

constexpr int GV_INFLIGHT = 8;   // 16-byte loads in flight per lane

__device__ __forceinline__ void prefetch_row(const __nv_bfloat16* row, int nchunks, int lane, uint4 (&pre)[GV_INFLIGHT]) {
    const uint4* p = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int i = 0; i < GV_INFLIGHT; ++i) {
        const int c = lane + i * 32;
        pre[i] = (c < nchunks) ? ldg_stream(p + c) : make_uint4(0, 0, 0, 0);
    }
}
// one weight row against the smem-resident input (per-lane partial sum); `pre` holds the first GV_INFLIGHT chunks, already loaded
__device__ __forceinline__ float dot_partial(const __nv_bfloat16* row, const uint4* xs, int nchunks, int lane,
                                             const uint4 (&pre)[GV_INFLIGHT]) {
    const uint4* p = reinterpret_cast<const uint4*>(row);
    float acc = 0.f;
    int c = lane;
#pragma unroll
    for (int i = 0; i < GV_INFLIGHT; ++i)
        if (c + i * 32 < nchunks) acc += dot8(pre[i], xs[c + i * 32]);
    c += GV_INFLIGHT * 32;
    for (; c + (GV_INFLIGHT - 1) * 32 < nchunks; c += GV_INFLIGHT * 32) {
        uint4 w[GV_INFLIGHT];
#pragma unroll
        for (int i = 0; i < GV_INFLIGHT; ++i) w[i] = ldg_stream(p + c + i * 32);
#pragma unroll
        for (int i = 0; i < GV_INFLIGHT; ++i) acc += dot8(w[i], xs[c + i * 32]);
    }
    for (; c < nchunks; c += 32) acc += dot8(ldg_stream(p + c), xs[c]);
    return acc;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int MODE>
__device__ __forceinline__ const __nv_bfloat16* gemv_row(const GemvArgs& a, int n) {
    if constexpr (MODE == GV_QKV) {
        const int s = n / a.rows;
        const __nv_bfloat16* m = s == 0 ? a.w0 : (s == 1 ? a.w1 : a.w2);
        return m + static_cast<long long>(n - s * a.rows) * a.K;
    } else {
        return a.w0 + static_cast<long long>(n) * a.K;
    }
}

// ---- a GEMV phase in two halves, so both the stand-alone kernels and the persistent whole-token kernel can put a dependency
// wait / grid barrier between them.
// gemv_pre: touches only weights (never produced on the device).  Starts streaming the warp's first row into registers and pulls
// its next rows towards L2 (128-byte lines, budgeted) so HBM keeps streaming while the consumer of the previous phase drains.
template <int MODE>
__device__ __forceinline__ void gemv_pre(const GemvArgs& a, int gw, int nw, int lane, bool regs, uint4 (&pre)[GV_INFLIGHT]) {
    const int nchunks = a.K / 8;
    const int total = (MODE == GV_QKV) ? 3 * a.rows : a.rows;
    if (regs && gw < total) prefetch_row(gemv_row<MODE>(a, gw), nchunks, lane, pre);
    const int lines_per_row = a.K / 64;   // 128-byte lines in one weight row
    int budget = a.pf_bytes_per_warp / 128;
    for (int n = gw; n < total && budget > 0; n += nw) {
        const char* r0 = reinterpret_cast<const char*>(gemv_row<MODE>(a, n));
        const int take = min(budget, lines_per_row);
        for (int l = lane; l < take; l += 32) prefetch_l2(r0 + l * 128);
        if constexpr (MODE == GV_GATEUP) {
            const char* r1 = reinterpret_cast<const char*>(a.w1 + static_cast<long long>(n) * a.K);
            for (int l = lane; l < take; l += 32) prefetch_l2(r1 + l * 128);
            budget -= take;
        }
        budget -= take;
    }
}

struct GemvShared {
    float red[8];
    float bval[8];
    int bidx[8];
};

// gemv_main: stage the input vector in smem (HF RMSNorm fused: w * bf16(x * rstd)), then one weight row per warp trip.
template <int MODE>
__device__ __forceinline__ void gemv_main(const GemvArgs& a, __nv_bfloat16* xs, GemvShared& sh, int gw, int nw, int block_id,
                                          bool have_pre, uint4 (&pre)[GV_INFLIGHT]) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunks = a.K / 8;
    const int total = (MODE == GV_QKV) ? 3 * a.rows : a.rows;
    if (!have_pre && gw < total) prefetch_row(gemv_row<MODE>(a, gw), nchunks, lane, pre);
    if (a.norm_w != nullptr) {
        float ss = 0.f;
        for (int i = tid; i < a.K; i += 256) { const float v = __bfloat162float(a.x[i]); ss += v * v; }
        ss = warp_red(ss);
        if (lane == 0) sh.red[warp] = ss;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += sh.red[i];
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
    // rows are software-pipelined: the first chunks of the next row are requested before this row's warp reduction
    for (int n = gw; n < total; n += nw) {
        float acc = dot_partial(gemv_row<MODE>(a, n), xv, nchunks, lane, pre);
        if constexpr (MODE == GV_GATEUP) {
            prefetch_row(a.w1 + static_cast<long long>(n) * a.K, nchunks, lane, pre);
            const float v0 = warp_red(acc);
            acc = dot_partial(a.w1 + static_cast<long long>(n) * a.K, xv, nchunks, lane, pre);
            if (n + nw < total) prefetch_row(gemv_row<MODE>(a, n + nw), nchunks, lane, pre);
            const float u0 = warp_red(acc);
            if (lane == 0) {
                const float g = bf16_round(v0), u = bf16_round(u0);
                a.out[n] = __float2bfloat16_rn(bf16_round(g / (1.f + __expf(-g))) * u);
            }
        } else {
            if (n + nw < total) prefetch_row(gemv_row<MODE>(a, n + nw), nchunks, lane, pre);
            const float v0 = warp_red(acc);
            if (lane == 0) {
                if constexpr (MODE == GV_QKV) {
                    a.out[n] = __float2bfloat16_rn(v0);
                } else if constexpr (MODE == GV_O || MODE == GV_DOWN) {
                    a.out[n] = __float2bfloat16_rn(bf16_round(v0) + __bfloat162float(a.out[n]));
                } else {  // LMHEAD: HF casts the bf16 logits to fp32
                    const float v = bf16_round(v0);
                    a.logits[n] = v;
                    if (v > best || (v == best && n < best_i)) { best = v; best_i = n; }
                }
            }
        }
    }
    if constexpr (MODE == GV_LMHEAD) {
        if (lane == 0) { sh.bval[warp] = best; sh.bidx[warp] = best_i; }
        __syncthreads();
        if (tid == 0) {
            for (int i = 1; i < 8; ++i)
                if (sh.bval[i] > sh.bval[0] || (sh.bval[i] == sh.bval[0] && sh.bidx[i] < sh.bidx[0])) { sh.bval[0] = sh.bval[i]; sh.bidx[0] = sh.bidx[i]; }
            a.part_val[block_id] = sh.bval[0];
            a.part_idx[block_id] = sh.bidx[0];
        }
    }
    __syncthreads();   // xs / sh are reused by whatever runs next in this block
}

template <int MODE>
__global__ void __launch_bounds__(256, 4)
gemv_kernel(const GemvArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ GemvShared sh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
    pdl_launch_dependents();
    uint4 pre[GV_INFLIGHT];
    gemv_pre<MODE>(a, gw, nw, lane, true, pre);
    pdl_wait();
    if (a.done != nullptr && *a.done != 0) return;   // EOS / stop sequence already hit: the remaining enqueued steps cost nothing
    gemv_main<MODE>(a, reinterpret_cast<__nv_bfloat16*>(smem), sh, gw, nw, blockIdx.x, true, pre);
}

// ---- CTA-cooperative form (default): a CTA owns a CONTIGUOUS range of output rows and its 256 threads walk every row together
// (thread t takes 16-byte chunks t, t + 256, ...), the per-warp sums meet in shared memory once at the end.  With one row per
// warp trip (above) the work is quantised per WARP: 12288 q|k|v rows over 4736 warps is 2.59 rows each, so 41 % of the warps run a
// third trip while the rest idle (86 % balance; 77 % for gate/up, 87 % for down, whose 22 KB rows also serialise five dependent
// round trips in one warp).  Here the quantum is a row per CTA: 20.76 rows -> 21 (98.8 %), every row is eight warps wide.
constexpr int GV_MAX_LOCAL = 192;     // rows of one CTA (x 2 for gate/up pairs) the partial-sum table can hold

template <int MODE>
__device__ __forceinline__ const __nv_bfloat16* coop_row(const GemvArgs& a, int u0, int j) {
    if constexpr (MODE == GV_GATEUP) {           // local rows alternate gate_n, up_n
        const int n = u0 + (j >> 1);
        return ((j & 1) ? a.w1 : a.w0) + static_cast<long long>(n) * a.K;
    } else {
        return gemv_row<MODE>(a, u0 + j);
    }
}

// this CTA's unit range: units = output elements (QKV: 3*rows, GATEUP: hidden units, else rows), split as evenly as possible
template <int MODE>
__device__ __forceinline__ void coop_range(const GemvArgs& a, int block_id, int nblocks, int& u0, int& cnt) {
    const int units = (MODE == GV_QKV) ? 3 * a.rows : a.rows;
    const int base = units / nblocks, rem = units - base * nblocks;
    u0 = block_id * base + min(block_id, rem);
    cnt = base + (block_id < rem ? 1 : 0);
}

template <int MODE, int CPR, int RB>
__device__ __forceinline__ void coop_load(const GemvArgs& a, int u0, int rows_local, int j0, int nchunks, uint4 (&w)[RB * CPR]) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int r = 0; r < RB; ++r) {
        const uint4* row = reinterpret_cast<const uint4*>(coop_row<MODE>(a, u0, min(j0 + r, rows_local - 1)));
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
            const int c = tid + 256 * i;
            w[r * CPR + i] = (j0 + r < rows_local && c < nchunks) ? ldg_stream(row + c) : make_uint4(0, 0, 0, 0);
        }
    }
}

// rows [0, rows_local) of the CTA against the smem-resident input; part[j * 8 + warp] receives each warp's share of row j.
// Two register sets of RB rows each, refilled alternately: while one set is being consumed the other one's loads are in flight
// (a single set oscillates between RB * CPR loads in flight and none).  wa / wb hold rows [0, RB) / [RB, 2 RB), loaded before the
// dependency wait.
template <int MODE, int CPR, int RB>
__device__ __forceinline__ void coop_rows(const GemvArgs& a, const uint4* xv, float* part, int u0, int rows_local, int nchunks,
                                          uint4 (&wa)[RB * CPR], uint4 (&wb)[RB * CPR]) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr bool XREG = CPR <= 2;                  // this thread's slice of the input vector is the same for every row: keep it in
    uint4 xr[XREG ? CPR : 1];                        // registers when it is short (wide rows re-read shared memory: no spills)
    if constexpr (XREG) {
#pragma unroll
        for (int i = 0; i < CPR; ++i) xr[i] = (tid + 256 * i < nchunks) ? xv[tid + 256 * i] : make_uint4(0, 0, 0, 0);
    }
    auto consume = [&](uint4 (&w)[RB * CPR], int j0) {
        float acc[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            acc[r] = 0.f;
#pragma unroll
            for (int i = 0; i < CPR; ++i) {
                if constexpr (XREG) acc[r] += dot8(w[r * CPR + i], xr[i]);
                else if (tid + 256 * i < nchunks) acc[r] += dot8(w[r * CPR + i], xv[tid + 256 * i]);
            }
        }
        if (j0 + 2 * RB < rows_local) coop_load<MODE, CPR, RB>(a, u0, rows_local, j0 + 2 * RB, nchunks, w);   // refill this set
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const float v = warp_red(acc[r]);
            if (lane == 0 && j0 + r < rows_local) part[(j0 + r) * 8 + warp] = v;
        }
    };
    for (int j0 = 0; j0 < rows_local; j0 += 2 * RB) {
        consume(wa, j0);
        if (j0 + RB < rows_local) consume(wb, j0 + RB);
    }
}

template <int MODE, int CPR, int RB>
__device__ __forceinline__ void gemv_coop(const GemvArgs& a, uint8_t* smem, GemvShared& sh) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunks = a.K / 8;
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem);
    float* part = reinterpret_cast<float*>(smem + static_cast<size_t>(a.K) * 2);
    int u0, cnt;
    coop_range<MODE>(a, blockIdx.x, gridDim.x, u0, cnt);
    const int rows_local = (MODE == GV_GATEUP) ? 2 * cnt : cnt;
    pdl_launch_dependents();
    uint4 wa[RB * CPR], wb[RB * CPR];
    if (rows_local > 0) coop_load<MODE, CPR, RB>(a, u0, rows_local, 0, nchunks, wa);       // weights only: before the dependency wait
    if (rows_local > RB) coop_load<MODE, CPR, RB>(a, u0, rows_local, RB, nchunks, wb);
    {   // pull the CTA's next rows towards L2 while the producer of x drains (budgeted; 128-byte lines)
        const int lines_per_row = a.K / 64;
        const int lines = min(a.pf_bytes_per_warp * 8 / 128, rows_local * lines_per_row);
        for (int l = 2 * RB * lines_per_row + tid; l < lines; l += 256) {
            const int j = l / lines_per_row;
            prefetch_l2(reinterpret_cast<const char*>(coop_row<MODE>(a, u0, j)) + (l - j * lines_per_row) * 128);
        }
    }
    pdl_wait();
    // the input vector (cp.async straight into shared memory: no registers next to the two weight sets) and the "sequence finished"
    // flag are requested together: a flag test in front of the loads would put a second dependent L2 round trip at the head of
    // every launch of the chain
    {
        const uint32_t xs_u = smem_u32(xs);
        for (int c = tid; c < nchunks; c += 256)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(xs_u + c * 16), "l"(reinterpret_cast<const uint4*>(a.x) + c) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int finished = (a.done != nullptr) ? __ldcg(a.done) : 0;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (finished != 0) return;
    if (a.norm_w != nullptr) {                       // HF RMSNorm fused: w * bf16(x * rstd); every thread works on the chunks it copied
        float ss = 0.f;
        for (int c = tid; c < nchunks; c += 256) {
            const uint4 u = reinterpret_cast<const uint4*>(xs)[c];
            const uint32_t wds[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float lo = bf16_lo(wds[e]), hi = bf16_hi(wds[e]); ss += lo * lo + hi * hi; }
        }
        ss = warp_red(ss);
        if (lane == 0) sh.red[warp] = ss;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += sh.red[i];
        const float rstd = rsqrtf(tot / a.K + a.eps);
        for (int c = tid; c < nchunks; c += 256) {
            const uint4 u = reinterpret_cast<const uint4*>(xs)[c];
            const uint4 wn = reinterpret_cast<const uint4*>(a.norm_w)[c];
            const uint32_t xd[4] = {u.x, u.y, u.z, u.w}, wd[4] = {wn.x, wn.y, wn.z, wn.w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                o[e] = pack_bf16(bf16_lo(wd[e]) * bf16_round(bf16_lo(xd[e]) * rstd), bf16_hi(wd[e]) * bf16_round(bf16_hi(xd[e]) * rstd));
            reinterpret_cast<uint4*>(xs)[c] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
    __syncthreads();
    if (rows_local > 0) coop_rows<MODE, CPR, RB>(a, reinterpret_cast<const uint4*>(xs), part, u0, rows_local, nchunks, wa, wb);
    __syncthreads();
    // ---- finish: one thread per output element sums the eight warp shares in a fixed order
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    for (int j = tid; j < cnt; j += 256) {
        const int n = u0 + j;
        auto total = [&](int row) {
            const float* q = part + row * 8;
            return ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
        };
        if constexpr (MODE == GV_GATEUP) {
            const float g = bf16_round(total(2 * j)), u = bf16_round(total(2 * j + 1));
            a.out[n] = __float2bfloat16_rn(bf16_round(g / (1.f + __expf(-g))) * u);
        } else if constexpr (MODE == GV_QKV) {
            a.out[n] = __float2bfloat16_rn(total(j));
        } else if constexpr (MODE == GV_O || MODE == GV_DOWN) {
            a.out[n] = __float2bfloat16_rn(bf16_round(total(j)) + __bfloat162float(a.out[n]));
        } else {
            const float v = bf16_round(total(j));
            a.logits[n] = v;
            if (v > best || (v == best && n < best_i)) { best = v; best_i = n; }
        }
    }
    if constexpr (MODE == GV_LMHEAD) {               // block argmax (ties -> lowest index)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
        }
        if (lane == 0) { sh.bval[warp] = best; sh.bidx[warp] = best_i; }
        __syncthreads();
        if (tid == 0) {
            for (int i = 1; i < 8; ++i)
                if (sh.bval[i] > sh.bval[0] || (sh.bval[i] == sh.bval[0] && sh.bidx[i] < sh.bidx[0])) { sh.bval[0] = sh.bval[i]; sh.bidx[0] = sh.bidx[i]; }
            a.part_val[blockIdx.x] = sh.bval[0];
            a.part_idx[blockIdx.x] = sh.bidx[0];
        }
    }
}

// CPR = 16-byte chunks of a row per thread (ceil(K / 8 / 256)), RB = rows per register set (2 * RB * CPR loads in flight per thread)
template <int MODE, int CPR, int RB>
__global__ void __launch_bounds__(256, 3)
gemv_coop_kernel(const GemvArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ GemvShared sh;
    gemv_coop<MODE, CPR, RB>(a, smem, sh);
}

// state: [0] token to feed next, [1] ctx_len (positions already in the cache), [2] number of tokens emitted, [3] unused
__device__ __forceinline__ void commit_token_body(const float* __restrict__ part_val, const int* __restrict__ part_idx, int nparts,
                                                  int forced_token, int* __restrict__ state, int* __restrict__ tokens_out, int max_tokens,
                                                  int set_ctx, const __nv_bfloat16* __restrict__ embed, int dim,
                                                  __nv_bfloat16* __restrict__ xbuf) {
    __shared__ int s_tok;
    __shared__ float s_bv[256];
    __shared__ int s_bi[256];
    if (forced_token < 0) {   // argmax over the per-block partials of the lm_head GEMV (ties -> lowest index, like torch.argmax)
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
            const float v = part_val[i];
            const int ix = part_idx[i];
            if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
        }
        s_bv[threadIdx.x] = bv; s_bi[threadIdx.x] = bi;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                const float v = s_bv[threadIdx.x + o];
                const int ix = s_bi[threadIdx.x + o];
                if (v > s_bv[threadIdx.x] || (v == s_bv[threadIdx.x] && ix < s_bi[threadIdx.x])) { s_bv[threadIdx.x] = v; s_bi[threadIdx.x] = ix; }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        int tok = forced_token;
        if (tok < 0) tok = s_bi[0];
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

__global__ void __launch_bounds__(256)
commit_token_kernel(const float* __restrict__ part_val, const int* __restrict__ part_idx, int nparts, int forced_token,
                    int* __restrict__ state, int* __restrict__ tokens_out, int max_tokens, int set_ctx,
                    const __nv_bfloat16* __restrict__ embed, int dim, __nv_bfloat16* __restrict__ xbuf) {
    pdl_launch_dependents();
    pdl_wait();
    commit_token_body(part_val, part_idx, nparts, forced_token, state, tokens_out, max_tokens, set_ctx, embed, dim, xbuf);
}

// ATT_SPLITS blocks per head ("flash decoding"): each takes a contiguous slice of the context, RoPEs q (and, if it owns the new
// position, RoPEs k and appends K/V to the PAGED cache), computes its slice's max / sum / un-normalised P·V and publishes the
// partial; the block that arrives last for a head (atomic ticket) merges the partials and writes the head's output.  One block
// per head was latency-bound: ~240 positions walked 4 rows at a time cost 12 us per layer (0.39 ms of a 3.2 ms token).
// dynamic smem: scores [max_ctx] fp32 | page ids [max_pages] int32
constexpr int ATT_SPLITS = 4;   // 128 blocks = one wave (the kernel needs 168 registers: one 256-thread block per SM); 8 splits measured slower
constexpr int ATT_PART = 132;    // floats per partial: m, l, pad, pad, o[128]
struct AttnDecodeShared {
    float qs[128];
    float red[8];
    float part[8][128];
    int last;
};
__device__ __forceinline__ void attn_decode_body(float* sc, AttnDecodeShared& sh, int h, int split, const __nv_bfloat16* __restrict__ qkv,
                                                 int dim, const KvGeom& kv, int layer, const int* __restrict__ state,
                                                 const float* __restrict__ cosT, const float* __restrict__ sinT,
                                                 __nv_bfloat16* __restrict__ obuf, int max_ctx, float* part_g, int* counter) {
    int* spg = reinterpret_cast<int*>(sc + max_ctx);
    float (&qs)[128] = sh.qs;
    float (&red)[8] = sh.red;
    float (&part)[8][128] = sh.part;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pos = state[1];
    const int n = pos + 1;
    const int per = (n + ATT_SPLITS - 1) / ATT_SPLITS;
    const int p0 = min(n, split * per), p1 = min(n, p0 + per);
    const int cnt = p1 - p0;
    const int npages = (n + kv.page_size - 1) / kv.page_size;
    const float scale = rsqrtf(128.f);
    const __nv_bfloat16* q = qkv + h * 128;
    const __nv_bfloat16* k = qkv + dim + h * 128;
    const __nv_bfloat16* v = qkv + 2 * dim + h * 128;
    for (int i = tid; i < npages; i += 256) spg[i] = kv.block_table[i];
    const bool owner = pos >= p0 && pos < p1;         // exactly one split owns the new position
    const int page_new = kv.block_table[pos / kv.page_size];
    const int slot_new = pos % kv.page_size;
    if (tid < 64) {
        const float c = cosT[static_cast<long long>(pos) * 64 + tid], s = sinT[static_cast<long long>(pos) * 64 + tid];
        const float q1 = __bfloat162float(q[tid]), q2 = __bfloat162float(q[tid + 64]);
        // HF apply_rotary_pos_emb in bf16: each product rounded, then the sum rounded
        qs[tid] = bf16_round(bf16_round(q1 * c) - bf16_round(q2 * s));
        qs[tid + 64] = bf16_round(bf16_round(q2 * c) + bf16_round(q1 * s));
        if (owner) {
            const float k1 = __bfloat162float(k[tid]), k2 = __bfloat162float(k[tid + 64]);
            __nv_bfloat16* kd = kv_ptr(kv, layer, 0, page_new, h, slot_new);
            kd[tid] = __float2bfloat16_rn(bf16_round(k1 * c) - bf16_round(k2 * s));
            kd[tid + 64] = __float2bfloat16_rn(bf16_round(k2 * c) + bf16_round(k1 * s));
        }
    } else if (tid < 128 && owner) {
        __nv_bfloat16* vd = kv_ptr(kv, layer, 1, page_new, h, slot_new);
        const int d = tid - 64;
        vd[d] = v[d];
        vd[d + 64] = v[d + 64];
    }
    __syncthreads();
    // ---- scores: one position per thread, 16 x 16-byte loads of its K row in flight
    float mx = -INFINITY;
    for (int i = tid; i < cnt; i += 256) {
        const int p = p0 + i;
        const uint4* kp = reinterpret_cast<const uint4*>(kv_ptr(kv, layer, 0, spg[p / kv.page_size], h, p % kv.page_size));
        uint4 w[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) w[c] = kp[c];
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float* qq = qs + c * 8;
            acc += bf16_lo(w[c].x) * qq[0] + bf16_hi(w[c].x) * qq[1] + bf16_lo(w[c].y) * qq[2] + bf16_hi(w[c].y) * qq[3] +
                   bf16_lo(w[c].z) * qq[4] + bf16_hi(w[c].z) * qq[5] + bf16_lo(w[c].w) * qq[6] + bf16_hi(w[c].w) * qq[7];
        }
        acc *= scale;                    // scores stay fp32 (as in the prefill flash kernel)
        sc[i] = acc;
        mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
    __syncthreads();
    float sum = 0.f;
    for (int i = tid; i < cnt; i += 256) {
        const float e = __expf(sc[i] - mx);
        sc[i] = e;
        sum += e;
    }
    sum = warp_red(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i];
    // ---- slice of P V: warp w takes positions w, w+8, ...; lane l owns dims 4l..4l+3 (one coalesced 256-byte row per load)
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int i = warp;
    for (; i + 24 < cnt; i += 32) {        // 4 independent row loads in flight
        uint2 r[4];
        float pr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int pp = p0 + i + u * 8;
            r[u] = reinterpret_cast<const uint2*>(kv_ptr(kv, layer, 1, spg[pp / kv.page_size], h, pp % kv.page_size))[lane];
            pr[u] = sc[i + u * 8];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            a0 += pr[u] * bf16_lo(r[u].x); a1 += pr[u] * bf16_hi(r[u].x);
            a2 += pr[u] * bf16_lo(r[u].y); a3 += pr[u] * bf16_hi(r[u].y);
        }
    }
    for (; i < cnt; i += 8) {
        const int pp = p0 + i;
        const uint2 r = reinterpret_cast<const uint2*>(kv_ptr(kv, layer, 1, spg[pp / kv.page_size], h, pp % kv.page_size))[lane];
        const float pr = sc[i];
        a0 += pr * bf16_lo(r.x); a1 += pr * bf16_hi(r.x); a2 += pr * bf16_lo(r.y); a3 += pr * bf16_hi(r.y);
    }
    part[warp][lane * 4 + 0] = a0; part[warp][lane * 4 + 1] = a1; part[warp][lane * 4 + 2] = a2; part[warp][lane * 4 + 3] = a3;
    __syncthreads();
    // ---- publish the partial, take a ticket; the last block of the head merges
    float* mine = part_g + (static_cast<long long>(h) * ATT_SPLITS + split) * ATT_PART;
    if (tid < 128) {
        float o = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) o += part[w][tid];
        mine[4 + tid] = o;
    }
    if (tid == 0) { mine[0] = mx; mine[1] = tot; }
    __threadfence();
    __syncthreads();
    if (tid == 0) sh.last = (atomicAdd(&counter[h], 1) == ATT_SPLITS - 1);
    __syncthreads();
    if (!sh.last) return;
    __threadfence();
    const float* all = part_g + static_cast<long long>(h) * ATT_SPLITS * ATT_PART;
    float M = -INFINITY;
#pragma unroll
    for (int s = 0; s < ATT_SPLITS; ++s) M = fmaxf(M, __ldcg(all + s * ATT_PART));
    if (tid < 128) {
        float L = 0.f, o = 0.f;
#pragma unroll
        for (int s = 0; s < ATT_SPLITS; ++s) {
            const float ms = __ldcg(all + s * ATT_PART);
            const float wgt = (ms == -INFINITY) ? 0.f : __expf(ms - M);
            L += wgt * __ldcg(all + s * ATT_PART + 1);
            o += wgt * __ldcg(all + s * ATT_PART + 4 + tid);
        }
        obuf[h * 128 + tid] = __float2bfloat16_rn(o / L);
    }
    if (tid == 0) counter[h] = 0;          // ready for the next layer / token
}

// ---- cluster form (default): the ATT_SPLITS blocks of a head are one thread-block cluster.  Against the ticket form above:
//   * K rows (one position per thread, 16 loads) and V rows are requested TOGETHER right after the RoPE barrier — V as 16-byte
//     chunks with the block's threads laid out (16 position lanes) x (16 dim chunks), so every V row of a <= 128-position slice
//     is in flight at once instead of 4 rows per warp behind the softmax (three dependent round trips at ctx ~ 300);
//   * the slices' (max, sum, P·V) meet through distributed shared memory and one cluster barrier instead of a global partial,
//     __threadfence, an atomic ticket and a read-back.
struct AttnDecode2Shared {
    float qs[128];
    float red[8];
    float part[16][128];
    float stat[4];      // m, l of this slice
    float o[128];       // un-normalised P·V of this slice (read by rank 0 through DSMEM)
};

__global__ void __launch_bounds__(256)
attn_decode2_kernel(const __nv_bfloat16* __restrict__ qkv, int dim, KvGeom kv, int layer, int* __restrict__ state,
                    const float* __restrict__ cosT, const float* __restrict__ sinT, __nv_bfloat16* __restrict__ obuf, int max_ctx) {
    extern __shared__ float sc[];            // scores [max_ctx] | page ids [max_pages]
    __shared__ AttnDecode2Shared sh;
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    pdl_launch_dependents();
    const int h = blockIdx.x, split = blockIdx.y;
    int* spg = reinterpret_cast<int*>(sc + max_ctx);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // the block table is written once per generation, long before the chain: stage all of it while the q|k|v GEMV drains
    for (int i = tid; i < kv.max_pages; i += 256) spg[i] = kv.block_table[i];
    pdl_wait();
    const int4 stv = __ldcg(reinterpret_cast<const int4*>(state));   // token, position, emitted, finished: one round trip
    if (stv.w != 0) return;                  // (the same for every block of the cluster: nobody is left at a cluster barrier)
    const int pos = stv.y;
    const int n = pos + 1;
    const int per = (n + ATT_SPLITS - 1) / ATT_SPLITS;
    const int p0 = min(n, split * per), p1 = min(n, p0 + per);
    const int cnt = p1 - p0;
    const int npages = (n + kv.page_size - 1) / kv.page_size;
    const float scale = rsqrtf(128.f);
    const __nv_bfloat16* q = qkv + h * 128;
    const __nv_bfloat16* k = qkv + dim + h * 128;
    const __nv_bfloat16* v = qkv + 2 * dim + h * 128;
    (void)npages;
    const bool owner = pos >= p0 && pos < p1;         // exactly one slice owns the new position
    const int page_new = kv.block_table[pos / kv.page_size];
    const int slot_new = pos % kv.page_size;
    if (tid < 64) {
        const float c = cosT[static_cast<long long>(pos) * 64 + tid], s = sinT[static_cast<long long>(pos) * 64 + tid];
        const float q1 = __bfloat162float(q[tid]), q2 = __bfloat162float(q[tid + 64]);
        // HF apply_rotary_pos_emb in bf16: each product rounded, then the sum rounded
        sh.qs[tid] = bf16_round(bf16_round(q1 * c) - bf16_round(q2 * s));
        sh.qs[tid + 64] = bf16_round(bf16_round(q2 * c) + bf16_round(q1 * s));
        if (owner) {
            const float k1 = __bfloat162float(k[tid]), k2 = __bfloat162float(k[tid + 64]);
            __nv_bfloat16* kd = kv_ptr(kv, layer, 0, page_new, h, slot_new);
            kd[tid] = __float2bfloat16_rn(bf16_round(k1 * c) - bf16_round(k2 * s));
            kd[tid + 64] = __float2bfloat16_rn(bf16_round(k2 * c) + bf16_round(k1 * s));
        }
    } else if (tid < 128 && owner) {
        __nv_bfloat16* vd = kv_ptr(kv, layer, 1, page_new, h, slot_new);
        const int d = tid - 64;
        vd[d] = v[d];
        vd[d + 64] = v[d + 64];
    }
    __syncthreads();
    // ---- request this thread's K row (first 256 positions of the slice) and its V chunks (first 128 positions) together
    const int pl = tid >> 4, vc = tid & 15;           // V layout: position lane x 16-byte dim chunk
    auto k_row = [&](int i) { const int p = p0 + i; return reinterpret_cast<const uint4*>(kv_ptr(kv, layer, 0, spg[p / kv.page_size], h, p % kv.page_size)); };
    auto v_chunk = [&](int i) { const int p = p0 + i; return reinterpret_cast<const uint4*>(kv_ptr(kv, layer, 1, spg[p / kv.page_size], h, p % kv.page_size)) + vc; };
    uint4 kw[16], vw[8];
    if (tid < cnt) {
        const uint4* kp = k_row(tid);
#pragma unroll
        for (int c = 0; c < 16; ++c) kw[c] = kp[c];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int i = pl + 16 * j;
        vw[j] = (i < cnt) ? *v_chunk(i) : make_uint4(0, 0, 0, 0);
    }
    // ---- scores (fp32, as in the prefill flash kernel)
    float mx = -INFINITY;
    for (int i = tid; i < cnt; i += 256) {
        if (i >= 256) {
            const uint4* kp = k_row(i);
#pragma unroll
            for (int c = 0; c < 16; ++c) kw[c] = kp[c];
        }
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float* qq = sh.qs + c * 8;
            acc += bf16_lo(kw[c].x) * qq[0] + bf16_hi(kw[c].x) * qq[1] + bf16_lo(kw[c].y) * qq[2] + bf16_hi(kw[c].y) * qq[3] +
                   bf16_lo(kw[c].z) * qq[4] + bf16_hi(kw[c].z) * qq[5] + bf16_lo(kw[c].w) * qq[6] + bf16_hi(kw[c].w) * qq[7];
        }
        acc *= scale;
        sc[i] = acc;
        mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) sh.red[warp] = mx;
    __syncthreads();
    mx = sh.red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, sh.red[i]);
    __syncthreads();
    float sum = 0.f;
    for (int i = tid; i < cnt; i += 256) {
        const float e = __expf(sc[i] - mx);
        sc[i] = e;
        sum += e;
    }
    sum = warp_red(sum);
    if (lane == 0) sh.red[warp] = sum;
    __syncthreads();                                  // also publishes every sc[i]
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += sh.red[i];
    // ---- P·V: thread (pl, vc) accumulates dims [8 vc, 8 vc + 8) over positions pl, pl + 16, ...
    float a[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = 0.f;
    for (int base = 0; base < cnt; base += 128) {
        if (base > 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = base + pl + 16 * j;
                vw[j] = (i < cnt) ? *v_chunk(i) : make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = base + pl + 16 * j;
            const float pr = (i < cnt) ? sc[i] : 0.f;
            a[0] += pr * bf16_lo(vw[j].x); a[1] += pr * bf16_hi(vw[j].x); a[2] += pr * bf16_lo(vw[j].y); a[3] += pr * bf16_hi(vw[j].y);
            a[4] += pr * bf16_lo(vw[j].z); a[5] += pr * bf16_hi(vw[j].z); a[6] += pr * bf16_lo(vw[j].w); a[7] += pr * bf16_hi(vw[j].w);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) sh.part[pl][vc * 8 + e] = a[e];
    __syncthreads();
    if (tid < 128) {
        float o = 0.f;
#pragma unroll
        for (int w = 0; w < 16; ++w) o += sh.part[w][tid];
        sh.o[tid] = o;
    }
    if (tid == 0) { sh.stat[0] = mx; sh.stat[1] = tot; }
    cluster.sync();
    if (split == 0 && tid < 128) {                     // rank 0 merges the slices (fixed order)
        float ms[ATT_SPLITS], M = -INFINITY;
        const AttnDecode2Shared* peer[ATT_SPLITS];
#pragma unroll
        for (int r = 0; r < ATT_SPLITS; ++r) {
            peer[r] = cluster.map_shared_rank(&sh, r);
            ms[r] = peer[r]->stat[0];
            M = fmaxf(M, ms[r]);
        }
        float L = 0.f, o = 0.f;
#pragma unroll
        for (int r = 0; r < ATT_SPLITS; ++r) {
            const float wgt = (ms[r] == -INFINITY) ? 0.f : __expf(ms[r] - M);
            L += wgt * peer[r]->stat[1];
            o += wgt * peer[r]->o[tid];
        }
        obuf[h * 128 + tid] = __float2bfloat16_rn(o / L);
    }
    cluster.sync();                                    // the peers' shared memory stays alive until rank 0 has read it
}

__global__ void __launch_bounds__(256)
attn_decode_kernel(const __nv_bfloat16* __restrict__ qkv, int dim, KvGeom kv, int layer, int* __restrict__ state,
                   const float* __restrict__ cosT, const float* __restrict__ sinT, __nv_bfloat16* __restrict__ obuf, int max_ctx,
                   float* part_g, int* counter) {
    extern __shared__ float sc[];            // scores [max_ctx] | page ids [max_pages]
    __shared__ AttnDecodeShared sh;
    pdl_launch_dependents();
    pdl_wait();
    if (state[3] != 0) return;
    attn_decode_body(sc, sh, blockIdx.x, blockIdx.y, qkv, dim, kv, layer, state, cosT, sinT, obuf, max_ctx, part_g, counter);
}

// Device-side selection + commit of the next token (sampling.cuh).  Decode chain: state != nullptr — the history is the tokens
// generated so far, the draw index is their count, and the token is committed exactly like commit_token_kernel does; state[3] is
// raised on EOS / a stop sequence.  Stand-alone (state == nullptr): history / n_hist / draw are given, the token goes to token_out.
__global__ void __launch_bounds__(SMP_THREADS)
sample_commit_kernel(const float* __restrict__ logits, int V, SampleParams p, float* __restrict__ z, const int* history,
                     int n_hist, unsigned long long draw, int* state, int* tokens_out, int max_tokens,
                     int set_ctx, const __nv_bfloat16* __restrict__ embed, int dim, __nv_bfloat16* __restrict__ xbuf,
                     int* __restrict__ token_out, unsigned long long* __restrict__ dbg, int z_in_smem,
                     const float* __restrict__ part_val, const int* __restrict__ part_idx, int nparts) {
    extern __shared__ unsigned smp_hist[];
    __shared__ SampleShared sh;
    // the working copy of the scores lives in shared memory when the vocabulary fits next to the histograms (LLaMA: 128 KB):
    // the ~12 passes of the selection then run at shared-memory latency instead of L2 latency
    if (z_in_smem) z = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(smp_hist) + SMP_SMEM_BYTES);
    pdl_launch_dependents();
    pdl_wait();
    if (state != nullptr) {
        if (state[3] != 0) return;
        n_hist = min(state[2], max_tokens);
        history = tokens_out;
        draw = static_cast<unsigned long long>(state[2]);
    }
    int tok;
    const bool pen_on = p.penalty > 0.f && p.penalty != 1.f && n_hist > 0;
    if (!p.do_sample && !pen_on && nparts > 0) {
        // plain greedy: the lm_head GEMV already reduced every block's rows to (max, argmax); finish over those partials
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int i = threadIdx.x; i < nparts; i += SMP_THREADS) {
            const float v = part_val[i];
            const int ix = part_idx[i];
            if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { sh.wmax[threadIdx.x >> 5] = bv; sh.widx[threadIdx.x >> 5] = bi; }
        __syncthreads();
        bv = sh.wmax[0]; bi = sh.widx[0];
#pragma unroll
        for (int i = 1; i < SMP_WARPS; ++i)
            if (sh.wmax[i] > bv || (sh.wmax[i] == bv && sh.widx[i] < bi)) { bv = sh.wmax[i]; bi = sh.widx[i]; }
        tok = bi;
    } else {
        tok = smp_choose(logits, V, p, z, history, n_hist, draw, smp_hist, sh, dbg);
    }
    if (state == nullptr) {
        if (threadIdx.x == 0) token_out[0] = tok;
        return;
    }
    if (threadIdx.x == 0) {
        const int k = state[2];
        if (k < max_tokens) tokens_out[k] = tok;
        state[0] = tok;
        state[2] = k + 1;
        if (set_ctx >= 0) state[1] = set_ctx;
        else if (set_ctx == -2) state[1] += 1;
        if (smp_is_stop(p, tokens_out, min(k + 1, max_tokens), tok)) state[3] = 1;
    }
    const uint4* src = reinterpret_cast<const uint4*>(embed + static_cast<long long>(tok) * dim);
    for (int c = threadIdx.x; c < dim / 8; c += blockDim.x) reinterpret_cast<uint4*>(xbuf)[c] = src[c];
}

// launch with programmatic stream serialization (PDL): the kernel may start while its predecessor drains
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace lhrs

using namespace lhrs;
typedef __nv_bfloat16 bf16;

template <int MODE>
static int launch_gemv(const GemvArgs& a_in, cudaStream_t st) {
    const int units = (MODE == GV_QKV) ? 3 * a_in.rows : a_in.rows;   // one weight row (GATEUP: one gate+up row pair) per warp trip
    // CTAs per SM: few enough that the successor grid (programmatic dependent launch) becomes co-resident and prefetches
    static int per_sm = -1, pf_mb = -1;
    if (per_sm < 0) { const char* e = getenv("LHRS_GEMV_CTAS_PER_SM"); per_sm = e ? atoi(e) : 3; if (per_sm < 1 || per_sm > 8) per_sm = 3; }
    if (pf_mb < 0) { const char* e = getenv("LHRS_GEMV_PF_MB"); pf_mb = e ? atoi(e) : 12; }
    int grid = (units + 7) / 8;
    const int cap = num_sms() * per_sm;
    if (grid > cap) grid = cap;
    GemvArgs a = a_in;
    a.pf_bytes_per_warp = (int)(((long long)pf_mb << 20) / ((long long)grid * 8));
    const size_t smem = (size_t)a.K * 2;
    LHRS_CHECK_ARG(smem <= 64 * 1024 && a.K % 8 == 0, "gemv: K=%d unsupported", a.K);
    const bool prof = prof_on();
    const double bytes = 2.0 * (double)a.K * (MODE == GV_QKV ? 3.0 * a.rows : (MODE == GV_GATEUP ? 2.0 * a.rows : (double)a.rows));
    // CTA-cooperative rows (default; LHRS_GEMV_COOP=0: one row per warp trip) when a row is 1, 2 or 5-6 chunks per thread wide
    static int coop = -1;
    if (coop < 0) { const char* e = getenv("LHRS_GEMV_COOP"); coop = e ? atoi(e) : 1; }
    const int cpr = (a.K / 8 + 255) / 256;
    const int cgrid = units < cap ? units : cap;               // GATEUP: `units` counts hidden units (row pairs)
    const int local = ((units + cgrid - 1) / cgrid) * (MODE == GV_GATEUP ? 2 : 1);
    if (coop && (cpr == 1 || cpr == 2 || cpr == 5 || cpr == 6) && local <= GV_MAX_LOCAL && a.K % 64 == 0) {
        const size_t csmem = smem + (size_t)local * 8 * sizeof(float);
        a.pf_bytes_per_warp = (int)(((long long)pf_mb << 20) / ((long long)cgrid * 8));
        auto launch = [&](auto kern) -> int {
            static bool attr_done = false;           // (one flag per instantiation of this generic lambda)
            if (!attr_done) { LHRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024)); attr_done = true; }
            if (prof) prof_begin(PROF_OTHER, 0.0, bytes, st);
            LHRS_CUDA(launch_pdl(kern, dim3(cgrid), dim3(256), csmem, st, a));
            if (prof) prof_end(st);
            LHRS_LAUNCH_CHECK("gemv_coop_kernel");
            return LHRS_OK;
        };
        int rc;
        static int deep = -1;                        // rows per register set (two sets): LHRS_GEMV_DEEP=1 doubles them
        if (deep < 0) { const char* e = getenv("LHRS_GEMV_DEEP"); deep = e ? atoi(e) : 0; }
        if (cpr == 1) rc = launch(gemv_coop_kernel<MODE, 1, 4>);
        else if (cpr == 2) rc = deep == 1 ? launch(gemv_coop_kernel<MODE, 2, 3>) : launch(gemv_coop_kernel<MODE, 2, 2>);
        else rc = launch(gemv_coop_kernel<MODE, 6, 1>);
        return rc ? -1 : cgrid;
    }
    auto kern = gemv_kernel<MODE>;
    static bool attr = false;
    if (!attr) {
        LHRS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr = true;
    }
    if (prof) prof_begin(PROF_OTHER, 0.0, bytes, st);
    LHRS_CUDA(launch_pdl(kern, dim3(grid), dim3(256), smem, st, a));
    if (prof) prof_end(st);
    LHRS_LAUNCH_CHECK("gemv_kernel");
    return grid;
}

static SampleParams sample_params(const LhrsSampling* s) {
    SampleParams p;
    p.do_sample = s->do_sample; p.temperature = s->temperature; p.top_k = s->top_k; p.top_p = s->top_p;
    p.penalty = s->repetition_penalty; p.eos = s->eos_token; p.seed = s->seed;
    p.seed_dev = (const unsigned long long*)s->seed_dev;
    p.stop_seqs = s->stop_seqs; p.n_stop = s->stop_seqs ? s->n_stop : 0; p.stop_len = s->stop_len;
    return p;
}
constexpr size_t SMP_MAX_SMEM = 220 * 1024;   // 227 KB per block minus the static shared variables
static size_t sample_smem(int vocab) {   // histograms + (if it fits) the fp32 working copy of the scores
    const size_t with_z = SMP_SMEM_BYTES + (size_t)vocab * sizeof(float);
    return with_z <= SMP_MAX_SMEM ? with_z : SMP_SMEM_BYTES;
}
static int sample_attr() {
    static bool done = false;
    if (!done) {
        LHRS_CUDA(cudaFuncSetAttribute(sample_commit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMP_MAX_SMEM));
        done = true;
    }
    return LHRS_OK;
}

static int lm_head_and_commit(const LhrsLlamaWeights* w, const LhrsDecodeBuffers* b, const bf16* x, const bf16* norm_w, int greedy,
                              int set_ctx, cudaStream_t st, const LhrsSampling* smp = nullptr, bool first = false) {
    GemvArgs a;
    memset(&a, 0, sizeof(a));
    a.w0 = (const bf16*)w->lm_head; a.rows = w->vocab; a.K = w->dim; a.x = x; a.norm_w = norm_w; a.eps = w->eps;
    a.logits = b->logits; a.part_val = b->part_val; a.part_idx = b->part_idx;
    a.done = first ? nullptr : (const int*)b->state + 3;
    const int grid = launch_gemv<GV_LMHEAD>(a, st);
    if (grid <= 0) return LHRS_ERR_CUDA;
    if (smp != nullptr) {
        if (sample_attr()) return LHRS_ERR_CUDA;
        const size_t smem = sample_smem(w->vocab);
        LHRS_CUDA(launch_pdl(sample_commit_kernel, dim3(1), dim3(SMP_THREADS), smem, st, (const float*)b->logits, (int)w->vocab,
                             sample_params(smp), (float*)smp->work, (const int*)nullptr, 0, 0ull, (int*)b->state, (int*)b->tokens_out,
                             (int)b->max_tokens, set_ctx, (const bf16*)w->embed, (int)w->dim, (bf16*)b->xbuf, (int*)nullptr,
                             (unsigned long long*)nullptr, (int)(smem > SMP_SMEM_BYTES), (const float*)b->part_val, (const int*)b->part_idx, grid));
        LHRS_LAUNCH_CHECK("sample_commit_kernel");
        return LHRS_OK;
    }
    if (greedy) {
        LHRS_CUDA(launch_pdl(commit_token_kernel, dim3(1), dim3(256), 0, st, (const float*)b->part_val, (const int*)b->part_idx, grid, -1,
                             (int*)b->state, (int*)b->tokens_out, (int)b->max_tokens, set_ctx, (const bf16*)w->embed, (int)w->dim, (bf16*)b->xbuf));
        LHRS_LAUNCH_CHECK("commit_token_kernel");
    }
    return LHRS_OK;
}

extern "C" int lhrs_decode_commit_token(const LhrsLlamaWeights* w, const LhrsDecodeBuffers* b, int32_t token, int32_t set_ctx, void* stream) {
    LHRS_CHECK_ARG(w && b && token >= 0 && token < w->vocab, "lhrs_decode_commit_token: bad token %d", token);
    LHRS_CUDA(launch_pdl(commit_token_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, (const float*)b->part_val, (const int*)b->part_idx, 0,
                         (int)token, (int*)b->state, (int*)b->tokens_out, (int)b->max_tokens, (int)set_ctx, (const bf16*)w->embed, (int)w->dim,
                         (bf16*)b->xbuf));
    LHRS_LAUNCH_CHECK("commit_token_kernel");
    return LHRS_OK;
}

static int check_sampling(const LhrsSampling* s, const char* who) {
    LHRS_CHECK_ARG(s && s->work, "%s: LhrsSampling.work (>= vocab floats of device scratch) is required", who);
    LHRS_CHECK_ARG(!s->stop_seqs || (s->n_stop > 0 && s->stop_len > 0), "%s: stop_seqs needs n_stop, stop_len > 0", who);
    return LHRS_OK;
}

extern "C" int lhrs_sample_logits(const float* logits, int32_t vocab, const int32_t* history, int32_t n_history, const LhrsSampling* s,
                                  uint64_t draw, int32_t* token_out, uint64_t* debug4, void* stream) {
    LHRS_CHECK_ARG(logits && vocab > 0 && vocab <= SMP_MAX_VOCAB && token_out && (n_history == 0 || history),
                   "lhrs_sample_logits: bad args (vocab must be in 1..%d)", SMP_MAX_VOCAB);
    if (check_sampling(s, "lhrs_sample_logits")) return LHRS_ERR_INVALID;
    if (sample_attr()) return LHRS_ERR_CUDA;
    const size_t smem = sample_smem(vocab);
    sample_commit_kernel<<<1, SMP_THREADS, smem, (cudaStream_t)stream>>>(
        logits, vocab, sample_params(s), s->work, (const int*)history, n_history, (unsigned long long)draw, nullptr, nullptr, 0, 0, nullptr, 0,
        nullptr, (int*)token_out, (unsigned long long*)debug4, (int)(smem > SMP_SMEM_BYTES), nullptr, nullptr, 0);
    LHRS_LAUNCH_CHECK("sample_commit_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_llama_first_token(const LhrsLlamaWeights* w, const void* hidden_last, int32_t ctx_len, const LhrsDecodeBuffers* b,
                                      int32_t greedy, void* stream) {
    LHRS_CHECK_ARG(w && hidden_last && b && ctx_len > 0, "lhrs_llama_first_token: bad args");
    LHRS_CUDA(cudaMemsetAsync(b->state, 0, 4 * sizeof(int), (cudaStream_t)stream));
    return lm_head_and_commit(w, b, (const bf16*)hidden_last, nullptr, greedy, ctx_len, (cudaStream_t)stream, nullptr, true);
}

extern "C" int lhrs_llama_first_token_sampled(const LhrsLlamaWeights* w, const void* hidden_last, int32_t ctx_len,
                                              const LhrsDecodeBuffers* b, const LhrsSampling* s, void* stream) {
    LHRS_CHECK_ARG(w && hidden_last && b && ctx_len > 0, "lhrs_llama_first_token_sampled: bad args");
    if (check_sampling(s, "lhrs_llama_first_token_sampled")) return LHRS_ERR_INVALID;
    LHRS_CHECK_ARG(w->vocab <= SMP_MAX_VOCAB, "lhrs_llama_first_token_sampled: vocab %d > %d", w->vocab, SMP_MAX_VOCAB);
    LHRS_CUDA(cudaMemsetAsync(b->state, 0, 4 * sizeof(int), (cudaStream_t)stream));
    return lm_head_and_commit(w, b, (const bf16*)hidden_last, nullptr, 0, ctx_len, (cudaStream_t)stream, s, true);
}

static int decode_step_impl(const LhrsLlamaWeights* w, const LhrsKvCache* kv, const LhrsDecodeBuffers* b, int32_t greedy,
                            int32_t max_ctx, const LhrsSampling* smp, void* stream);

extern "C" int lhrs_llama_decode_step(const LhrsLlamaWeights* w, const LhrsKvCache* kv, const LhrsDecodeBuffers* b, int32_t greedy,
                                      int32_t max_ctx, void* stream) {
    return decode_step_impl(w, kv, b, greedy, max_ctx, nullptr, stream);
}

extern "C" int lhrs_llama_decode_step_sampled(const LhrsLlamaWeights* w, const LhrsKvCache* kv, const LhrsDecodeBuffers* b,
                                              const LhrsSampling* s, int32_t max_ctx, void* stream) {
    if (check_sampling(s, "lhrs_llama_decode_step_sampled")) return LHRS_ERR_INVALID;
    return decode_step_impl(w, kv, b, 0, max_ctx, s, stream);
}

static int decode_step_impl(const LhrsLlamaWeights* w, const LhrsKvCache* kv, const LhrsDecodeBuffers* b, int32_t greedy,
                            int32_t max_ctx, const LhrsSampling* smp, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LHRS_CHECK_ARG(w && kv && b, "lhrs_llama_decode_step: null");
    LHRS_CHECK_ARG(b->attn_part && b->attn_count, "lhrs_llama_decode_step: LhrsDecodeBuffers.attn_part / attn_count are required");
    LHRS_CHECK_ARG(w->dim / w->heads == 128 && kv->head_dim == 128 && kv->heads == w->heads, "lhrs_llama_decode_step: head_dim must be 128");
    LHRS_CHECK_ARG(w->lora_r == 0, "lhrs_llama_decode_step: merge the LoRA adapters first (merge_and_unload; the reference does so for eval, UniBind.py:114)");
    LHRS_CHECK_ARG(max_ctx > 0 && max_ctx <= kv->max_pages * kv->page_size && max_ctx <= w->max_pos, "lhrs_llama_decode_step: max_ctx %d", max_ctx);
    const int D = w->dim, F = w->ffn;
    bf16* x = (bf16*)b->xbuf;
    const int* done = (const int*)b->state + 3;
    for (int l = 0; l < w->num_layers; ++l) {
        GemvArgs a;
        memset(&a, 0, sizeof(a));
        a.w0 = (const bf16*)w->q_w[l]; a.w1 = (const bf16*)w->k_w[l]; a.w2 = (const bf16*)w->v_w[l];
        a.rows = D; a.K = D; a.x = x; a.norm_w = (const bf16*)w->ln1_w[l]; a.eps = w->eps; a.out = (bf16*)b->qkv;
        a.done = done;
        if (launch_gemv<GV_QKV>(a, st) <= 0) return LHRS_ERR_CUDA;
        static int skip_attn = -1;   // experiment knob (changes results): how much of a token is the attention launch?
        if (skip_attn < 0) { const char* e = getenv("LHRS_DECODE_SKIP_ATTN"); skip_attn = e ? atoi(e) : 0; }
        static int attn2 = -1;       // cluster form of the decode attention (default); LHRS_DECODE_ATTN2=0: ticket form
        if (attn2 < 0) { const char* e = getenv("LHRS_DECODE_ATTN2"); attn2 = e ? atoi(e) : 1; }
        const size_t att_smem = (size_t)max_ctx * sizeof(float) + (size_t)kv->max_pages * sizeof(int);
        if (!skip_attn && attn2) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(w->heads, ATT_SPLITS); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = att_smem; cfg.stream = st;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            attr[1].id = cudaLaunchAttributeClusterDimension;
            attr[1].val.clusterDim.x = 1; attr[1].val.clusterDim.y = ATT_SPLITS; attr[1].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 2;
            LHRS_CUDA(cudaLaunchKernelEx(&cfg, attn_decode2_kernel, (const bf16*)b->qkv, (int)D, *kv, (int)l, (int*)b->state,
                                         (const float*)w->rope_cos, (const float*)w->rope_sin, (bf16*)b->obuf, (int)max_ctx));
        } else if (!skip_attn)
        LHRS_CUDA(launch_pdl(attn_decode_kernel, dim3(w->heads, ATT_SPLITS), dim3(256), att_smem, st,
                             (const bf16*)b->qkv, D, *kv, l, (int*)b->state, (const float*)w->rope_cos, (const float*)w->rope_sin, (bf16*)b->obuf,
                             (int)max_ctx, (float*)b->attn_part, (int*)b->attn_count));
        LHRS_LAUNCH_CHECK("attn_decode_kernel");
        memset(&a, 0, sizeof(a));
        a.w0 = (const bf16*)w->o_w[l]; a.rows = D; a.K = D; a.x = (const bf16*)b->obuf; a.out = x; a.done = done;
        if (launch_gemv<GV_O>(a, st) <= 0) return LHRS_ERR_CUDA;
        memset(&a, 0, sizeof(a));
        a.w0 = (const bf16*)w->gate_w[l]; a.w1 = (const bf16*)w->up_w[l]; a.rows = F; a.K = D; a.x = x;
        a.norm_w = (const bf16*)w->ln2_w[l]; a.eps = w->eps; a.out = (bf16*)b->act; a.done = done;
        if (launch_gemv<GV_GATEUP>(a, st) <= 0) return LHRS_ERR_CUDA;
        memset(&a, 0, sizeof(a));
        a.w0 = (const bf16*)w->down_w[l]; a.rows = D; a.K = F; a.x = (const bf16*)b->act; a.out = x; a.done = done;
        if (launch_gemv<GV_DOWN>(a, st) <= 0) return LHRS_ERR_CUDA;
    }
    // the new position is now in the cache for every layer
    {
        int rc = lm_head_and_commit(w, b, x, (const bf16*)w->norm_w, greedy, -2, st, smp);
        if (rc) return rc;
    }
    return LHRS_OK;
}
