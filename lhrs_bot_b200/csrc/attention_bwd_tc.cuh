// Device helpers shared by the two tcgen05 attention-backward translation units (attention_bwd_tc.cu: one CTA per tile;
// attention_bwd_tcp.cu: persistent, tile-pipelined).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "attention_common.h"
#include "ptx.cuh"

namespace lhrs {

struct AttnBwdTcArgs {
    AttnBwdArgs a;
    int q_hfirst, k_hfirst, v_hfirst, o_hfirst;
};

namespace abt {
constexpr int HD = 128;
constexpr int T128 = 128 * HD * 2;     // 128-row operand tile: 2 k-blocks of [128 x 128 B]
constexpr int T64 = 64 * HD * 2;       // 64-row operand tile: 2 k-blocks of [64 x 128 B]
constexpr int PS = 128 * 64 * 2;       // a [128 x 64] bf16 probability / score-gradient tile
constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (UMMA_LAYOUT_SW128 << 29);
constexpr uint32_t LBO_8K = (8192u >> 4) << 16;
constexpr int NCOMPUTE = 256;          // 8 compute warps
}  // namespace abt

__device__ __forceinline__ void bar_sync_compute() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// named barrier of one 128-thread compute group (ids 2 and 3; id 1 is the 256-thread one above, 0 is __syncthreads)
__device__ __forceinline__ void bar_sync_group(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 2) : "memory"); }
__device__ __forceinline__ void prefetch_l1(const void* ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

// Write 32 bf16 values (16 packed words) of row `row` into a [128 x 64] SW128 K-major tile; `half` selects columns 32*half..
__device__ __forceinline__ void store_half_row(uint8_t* tile, int row, int half, const uint32_t (&w)[16]) {
    uint8_t* r = tile + row * 128;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
        const int ch = half * 4 + c4;
        *reinterpret_cast<uint4*>(r + ((ch ^ (row & 7)) << 4)) = make_uint4(w[c4 * 4], w[c4 * 4 + 1], w[c4 * 4 + 2], w[c4 * 4 + 3]);
    }
}

// Epilogue helper: this warp owns rotation pairs jj in [32*half, 32*half+32) of a [128 lanes x 128] fp32 TMEM accumulator:
// columns jj and jj+64.  Scales, optionally un-rotates (position = row index), and stores both 64-byte pieces of the row.
__device__ __forceinline__ void store_grad_row(uint32_t taddr, int half, float scale, const float* cosT, const float* sinT, int pos,
                                               __nv_bfloat16* dst, bool ok, bool have) {
    uint32_t lo[32], hi[32];
    if (have) {
        tmem_ld_32x32(taddr + half * 32, lo);
        tmem_ld_32x32(taddr + 64 + half * 32, hi);
        tmem_ld_wait();
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) lo[i] = hi[i] = 0u;
    }
    if (!ok) return;
    float a[32], b[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        a[i] = __uint_as_float(lo[i]) * scale;
        b[i] = __uint_as_float(hi[i]) * scale;
    }
    if (cosT != nullptr) {
        const float4* c4 = reinterpret_cast<const float4*>(cosT + static_cast<long long>(pos) * 64 + half * 32);
        const float4* s4 = reinterpret_cast<const float4*>(sinT + static_cast<long long>(pos) * 64 + half * 32);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const float4 c = c4[v], s = s4[v];
            const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float y1 = a[v * 4 + e], y2 = b[v * 4 + e];
                a[v * 4 + e] = y1 * cc[e] + y2 * ss[e];   // dx1 = dy1 c + dy2 s
                b[v * 4 + e] = y2 * cc[e] - y1 * ss[e];   // dx2 = dy2 c - dy1 s
            }
        }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        uint4 x, y;
        x.x = pack_bf16(a[v * 8 + 0], a[v * 8 + 1]); x.y = pack_bf16(a[v * 8 + 2], a[v * 8 + 3]);
        x.z = pack_bf16(a[v * 8 + 4], a[v * 8 + 5]); x.w = pack_bf16(a[v * 8 + 6], a[v * 8 + 7]);
        y.x = pack_bf16(b[v * 8 + 0], b[v * 8 + 1]); y.y = pack_bf16(b[v * 8 + 2], b[v * 8 + 3]);
        y.z = pack_bf16(b[v * 8 + 4], b[v * 8 + 5]); y.w = pack_bf16(b[v * 8 + 6], b[v * 8 + 7]);
        *reinterpret_cast<uint4*>(dst + half * 32 + v * 8) = x;
        *reinterpret_cast<uint4*>(dst + 64 + half * 32 + v * 8) = y;
    }
}

}  // namespace lhrs
