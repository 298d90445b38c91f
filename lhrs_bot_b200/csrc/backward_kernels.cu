// Row / elementwise kernels of the backward pass and the fused optimizer step (§8f-2).
// All HBM-bound single passes with 16-byte accesses: RMSNorm / LayerNorm backward (+ residual-stream add),
// SwiGLU / GELU / RoPE backward, column sums for bias gradients, global grad-norm and AdamW on flat buffers.
#include "host_common.h"
#include "ptx.cuh"
#include "dropout.cuh"

namespace lhrs {

constexpr int BT = 256;     // threads per block for row kernels
constexpr int MAXC = 4;     // 16-byte chunks per thread held in registers (dim <= 8192)

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float bsum(float v, float* red) {
    v = wsum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < (blockDim.x >> 5)) ? red[l] : 0.f;
    return wsum(t);
}
__device__ __forceinline__ void up8(const uint4& u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pk8(const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]); u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
    return u;
}

// ------------------------------------------------------------------ RMSNorm backward: y = w * (x * rstd)
// dx = rstd * (u - xhat * mean(u * xhat)), u = dy * w, xhat = x * rstd;   out = dres + dx (residual stream)
__global__ void __launch_bounds__(BT)
rmsnorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, const float* __restrict__ rstd,
                   const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ dres,
                   __nv_bfloat16* __restrict__ dx, int dim) {
    __shared__ float red[32];
    const long long row = blockIdx.x;
    const int nch = dim / 8;
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * dim);
    const uint4* dr = reinterpret_cast<const uint4*>(dy + row * dim);
    const uint4* wr = reinterpret_cast<const uint4*>(w);
    const float rs = rstd[row];
    float u[MAXC][8], xh[MAXC][8];
    float dot = 0.f;
    // the residual-stream gradient does not depend on the row reduction: request it together with x and dy so that the kernel has
    // one memory round trip before the reduction instead of one before and one after
    const uint4* rr = dres ? reinterpret_cast<const uint4*>(dres + row * dim) : nullptr;
    uint4 rraw[MAXC];
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = threadIdx.x + i * BT;
        rraw[i] = (rr != nullptr && c < nch) ? rr[c] : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = threadIdx.x + i * BT;
        if (c < nch) {
            float xv[8], dv[8], wv[8];
            up8(xr[c], xv); up8(dr[c], dv); up8(__ldg(wr + c), wv);
#pragma unroll
            for (int j = 0; j < 8; ++j) { u[i][j] = dv[j] * wv[j]; xh[i][j] = xv[j] * rs; dot += u[i][j] * xh[i][j]; }
        }
    }
    dot = bsum(dot, red) / dim;
    uint4* o = reinterpret_cast<uint4*>(dx + row * dim);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = threadIdx.x + i * BT;
        if (c < nch) {
            float r[8], out[8];
            up8(rraw[i], r);
#pragma unroll
            for (int j = 0; j < 8; ++j) out[j] = r[j] + rs * (u[i][j] - xh[i][j] * dot);
            o[c] = pk8(out);
        }
    }
}

// ------------------------------------------------------------------ LayerNorm backward, one WARP per row (dim = 256 * WC <= 1024)
// The block-per-row form below pays four block barriers per row and, at dim 1024, keeps half of its 256 threads idle: 85 us per
// launch on the pooler's 14592 x 1024 key/value rows.  Here every warp walks its own rows (lane l owns 16-byte chunks l, l + 32,
// ...), the two row sums are warp shuffles, and the eight warps' dgamma / dbeta columns meet in shared memory once at the end.
// Same partial layout as the block form: part[2][grid][dim], summed by colsum_final_kernel.
template <int WC>
__global__ void __launch_bounds__(BT)
layernorm_bwd_warp_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ w,
                          const float* __restrict__ mean, const float* __restrict__ rstd, const __nv_bfloat16* __restrict__ dy,
                          const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx, float* __restrict__ part,
                          long long rows, int dim) {
    __shared__ float buf[8][256 * WC];                     // <= 32 KB
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float dg[WC][8], db[WC][8], wv[WC][8];
#pragma unroll
    for (int i = 0; i < WC; ++i) {
        up8(__ldg(reinterpret_cast<const uint4*>(w) + lane + 32 * i), wv[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) { dg[i][j] = 0.f; db[i][j] = 0.f; }
    }
    const float inv_dim = 1.f / dim;
    for (long long row = static_cast<long long>(blockIdx.x) * 8 + warp; row < rows; row += static_cast<long long>(gridDim.x) * 8) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
        const uint4* dr = reinterpret_cast<const uint4*>(dy + row * dim);
        const uint4* rr = dres ? reinterpret_cast<const uint4*>(dres + row * dim) : nullptr;
        uint4 xraw[WC], draw[WC], rraw[WC];
#pragma unroll
        for (int i = 0; i < WC; ++i) {                     // every load of the row is in flight before the first use
            xraw[i] = xr[lane + 32 * i];
            draw[i] = dr[lane + 32 * i];
            rraw[i] = (rr != nullptr && dx != nullptr) ? rr[lane + 32 * i] : make_uint4(0, 0, 0, 0);
        }
        const float mu = mean[row], rs = rstd[row];
        float u[WC][8], xh[WC][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < WC; ++i) {
            float xv[8], dv[8];
            up8(xraw[i], xv); up8(draw[i], dv);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                xh[i][j] = (xv[j] - mu) * rs;
                u[i][j] = dv[j] * wv[i][j];
                s1 += u[i][j];
                s2 += u[i][j] * xh[i][j];
                dg[i][j] += dv[j] * xh[i][j];
                db[i][j] += dv[j];
            }
        }
        s1 = wsum(s1) * inv_dim;
        s2 = wsum(s2) * inv_dim;
        if (dx != nullptr) {
            uint4* o = reinterpret_cast<uint4*>(dx + row * dim);
#pragma unroll
            for (int i = 0; i < WC; ++i) {
                float r[8], out[8];
                up8(rraw[i], r);
#pragma unroll
                for (int j = 0; j < 8; ++j) out[j] = r[j] + rs * (u[i][j] - s1 - xh[i][j] * s2);
                o[lane + 32 * i] = pk8(out);
            }
        }
    }
    // the eight warps' column partials -> one row of part[] per block (dgamma, then dbeta through the same buffer)
    float* pg = part + static_cast<long long>(blockIdx.x) * dim;
    float* pb = part + (static_cast<long long>(gridDim.x) + blockIdx.x) * dim;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int i = 0; i < WC; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) buf[warp][(lane + 32 * i) * 8 + j] = pass == 0 ? dg[i][j] : db[i][j];
        __syncthreads();
        for (int c = threadIdx.x; c < dim; c += BT) {
            float t = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) t += buf[q][c];
            (pass == 0 ? pg : pb)[c] = t;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ LayerNorm backward
// Each block walks rows blockIdx.x, +gridDim.x, ...; per-thread column partials of dgamma/dbeta go to part[2][grid][dim].
__global__ void __launch_bounds__(BT)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ w,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const __nv_bfloat16* __restrict__ dy,
                     const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx, float* __restrict__ part,
                     long long rows, int dim) {
    __shared__ float red[32];
    const int nch = dim / 8;
    float dg[MAXC][8], db[MAXC][8];
#pragma unroll
    for (int i = 0; i < MAXC; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) { dg[i][j] = 0.f; db[i][j] = 0.f; }
    const uint4* wr = reinterpret_cast<const uint4*>(w);
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
        const uint4* dr = reinterpret_cast<const uint4*>(dy + row * dim);
        const float mu = mean[row], rs = rstd[row];
        float u[MAXC][8], xh[MAXC][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
            const int c = threadIdx.x + i * BT;
            if (c < nch) {
                float xv[8], dv[8], wv[8];
                up8(xr[c], xv); up8(dr[c], dv); up8(__ldg(wr + c), wv);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    xh[i][j] = (xv[j] - mu) * rs;
                    u[i][j] = dv[j] * wv[j];
                    s1 += u[i][j];
                    s2 += u[i][j] * xh[i][j];
                    dg[i][j] += dv[j] * xh[i][j];
                    db[i][j] += dv[j];
                }
            }
        }
        s1 = bsum(s1, red) / dim;
        s2 = bsum(s2, red) / dim;
        if (dx != nullptr) {
            const uint4* rr = dres ? reinterpret_cast<const uint4*>(dres + row * dim) : nullptr;
            uint4* o = reinterpret_cast<uint4*>(dx + row * dim);
#pragma unroll
            for (int i = 0; i < MAXC; ++i) {
                const int c = threadIdx.x + i * BT;
                if (c < nch) {
                    float r[8] = {0, 0, 0, 0, 0, 0, 0, 0}, out[8];
                    if (rr) up8(rr[c], r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) out[j] = r[j] + rs * (u[i][j] - s1 - xh[i][j] * s2);
                    o[c] = pk8(out);
                }
            }
        }
    }
    float* pg = part + static_cast<long long>(blockIdx.x) * dim;
    float* pb = part + (static_cast<long long>(gridDim.x) + blockIdx.x) * dim;
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
        const int c = threadIdx.x + i * BT;
        if (c < nch) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { pg[c * 8 + j] = dg[i][j]; pb[c * 8 + j] = db[i][j]; }
        }
    }
}

// column sums of a bf16 matrix [rows, n] (bias gradients): partial[gridDim.y][n] fp32
__global__ void __launch_bounds__(BT)
colsum_partial_kernel(const __nv_bfloat16* __restrict__ a, long long ld, long long rows, int n, float* __restrict__ part) {
    const int c = blockIdx.x * BT + threadIdx.x;
    if (c * 8 >= n) return;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (long long r = blockIdx.y; r < rows; r += gridDim.y) {
        float v[8];
        up8(*reinterpret_cast<const uint4*>(a + r * ld + c * 8), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    float* p = part + static_cast<long long>(blockIdx.y) * n + c * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j] = acc[j];
}
// out[c] (bf16) = sum_p part[p][c]  (+ out[c] if accumulate)
// 256 threads = 32 columns x 8 row lanes, four independent partial chains per thread (a single thread walking all ~296 partial
// rows was 16 us of dependent L2 round trips per launch, 67 launches per step); fixed summation order: deterministic.
// Launch with COLSUM_FINAL_GRID(n) blocks of 256 threads.
#define COLSUM_FINAL_GRID(n) (((n) + 31) / 32)
__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ part, int nparts, int n, __nv_bfloat16* __restrict__ out, int accumulate) {
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < n) {
        int p = ry;
        for (; p + 24 < nparts; p += 32) {
#pragma unroll
            for (int q = 0; q < 4; ++q) s[q] += part[static_cast<long long>(p + 8 * q) * n + c];
        }
        for (int q = 0; p < nparts; p += 8, ++q) s[q & 3] += part[static_cast<long long>(p) * n + c];
    }
    red[ry][cx] = (s[0] + s[1]) + (s[2] + s[3]);
    __syncthreads();
    if (ry == 0 && c < n) {
        float t = accumulate ? __bfloat162float(out[c]) : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += red[q][cx];
        out[c] = __float2bfloat16_rn(t);
    }
}

// ------------------------------------------------------------------ SwiGLU backward: act = silu(g) * u
__global__ void __launch_bounds__(BT)
swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ d_act, const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ u,
                  __nv_bfloat16* __restrict__ d_gu, long long rows, int f) {
    const long long chunks = rows * (f / 8);
    for (long long i = blockIdx.x * static_cast<long long>(BT) + threadIdx.x; i < chunks; i += static_cast<long long>(gridDim.x) * BT) {
        const long long r = i / (f / 8);
        const int c = i % (f / 8);
        float da[8], gv[8], uv[8], dg[8], du[8];
        up8(reinterpret_cast<const uint4*>(d_act)[i], da);
        up8(reinterpret_cast<const uint4*>(g)[i], gv);
        up8(reinterpret_cast<const uint4*>(u)[i], uv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float sg = 1.f / (1.f + __expf(-gv[j]));
            dg[j] = da[j] * uv[j] * sg * (1.f + gv[j] * (1.f - sg));
            du[j] = da[j] * gv[j] * sg;
        }
        reinterpret_cast<uint4*>(d_gu + r * 2 * f)[c] = pk8(dg);
        reinterpret_cast<uint4*>(d_gu + r * 2 * f + f)[c] = pk8(du);
    }
}

// GELU(erf) backward, in place on d:  d *= gelu'(pre)
__global__ void __launch_bounds__(BT)
gelu_bwd_kernel(__nv_bfloat16* __restrict__ d, const __nv_bfloat16* __restrict__ pre, long long chunks) {
    for (long long i = blockIdx.x * static_cast<long long>(BT) + threadIdx.x; i < chunks; i += static_cast<long long>(gridDim.x) * BT) {
        float dv[8], pv[8];
        up8(reinterpret_cast<const uint4*>(d)[i], dv);
        up8(reinterpret_cast<const uint4*>(pre)[i], pv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float x = pv[j];
            const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
            const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
            dv[j] *= cdf + x * pdf;
        }
        reinterpret_cast<uint4*>(d)[i] = pk8(dv);
    }
}

// RoPE backward in place on the q and k blocks of a packed [rows, 3*dim] gradient (inverse rotation), head_dim 128
__global__ void __launch_bounds__(BT)
rope_bwd_kernel(__nv_bfloat16* __restrict__ dqkv, long long ld, int dim, const float* __restrict__ cosT,
                const float* __restrict__ sinT, const int* __restrict__ positions, int seq_len) {
    const long long row = blockIdx.x;
    const int pos = positions ? positions[row] : static_cast<int>(row % seq_len);
    const float* cs = cosT + static_cast<long long>(pos) * 64;
    const float* sn = sinT + static_cast<long long>(pos) * 64;
    const int heads2 = 2 * dim / 128;  // q heads then k heads
    // one thread handles 8 frequency slots of one head: x1 = [j, j+8), x2 = [64+j, 64+j+8)
    for (int t = threadIdx.x; t < heads2 * 8; t += BT) {
        const int hh = t / 8, j0 = (t % 8) * 8;
        __nv_bfloat16* base = dqkv + row * ld + hh * 128;
        float a[8], b[8], oa[8], ob[8];
        up8(*reinterpret_cast<const uint4*>(base + j0), a);
        up8(*reinterpret_cast<const uint4*>(base + 64 + j0), b);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float c = cs[j0 + j], s = sn[j0 + j];
            oa[j] = a[j] * c + b[j] * s;     // y1 = x1 c - x2 s ; y2 = x2 c + x1 s  =>  dx1 = dy1 c + dy2 s
            ob[j] = b[j] * c - a[j] * s;     //                                          dx2 = dy2 c - dy1 s
        }
        *reinterpret_cast<uint4*>(base + j0) = pk8(oa);
        *reinterpret_cast<uint4*>(base + 64 + j0) = pk8(ob);
    }
}

// ------------------------------------------------------------------ flat-buffer optimizer (AdamW) + global grad norm
__global__ void __launch_bounds__(BT)
sumsq_partial_kernel(const __nv_bfloat16* __restrict__ g, long long n, float* __restrict__ part) {
    __shared__ float red[32];
    float acc = 0.f;
    for (long long i = blockIdx.x * static_cast<long long>(BT) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * BT) {
        const float v = __bfloat162float(g[i]);
        acc += v * v;
    }
    acc = bsum(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}
__global__ void sumsq_final_kernel(const float* __restrict__ part, int nparts, float* __restrict__ out) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += part[i];
    acc = bsum(acc, red);
    if (threadIdx.x == 0) out[0] = acc;
}
// p (fp32 master), m, v fp32; g bf16 (already averaged over ranks); writes bf16 params.  clip: scale = min(1, max_norm / ||g||)
__global__ void __launch_bounds__(BT)
adamw_kernel(float* __restrict__ master, float* __restrict__ m, float* __restrict__ v, const __nv_bfloat16* __restrict__ g,
             __nv_bfloat16* __restrict__ p_bf16, const float* __restrict__ decay_mask, long long n, float lr, float beta1,
             float beta2, float eps, float weight_decay, float bc1, float bc2, const float* __restrict__ gnorm_sq,
             float max_norm, float grad_scale) {
    float clip = grad_scale;
    if (max_norm > 0.f && gnorm_sq != nullptr) {
        const float norm = sqrtf(gnorm_sq[0]) * grad_scale;
        if (norm > max_norm) clip *= max_norm / (norm + 1e-6f);
    }
    for (long long i = blockIdx.x * static_cast<long long>(BT) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * BT) {
        const float gi = __bfloat162float(g[i]) * clip;
        float w = master[i];
        const float wd = decay_mask ? decay_mask[i] * weight_decay : weight_decay;
        w *= 1.f - lr * wd;                                  // decoupled weight decay (torch.optim.AdamW)
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        w -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
        master[i] = w;
        p_bf16[i] = __float2bfloat16_rn(w);
    }
}

}  // namespace lhrs

using namespace lhrs;
typedef __nv_bfloat16 bf16;

static inline unsigned grid_for(long long work, int per_block, int cap = 148 * 16) {
    long long g = (work + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return static_cast<unsigned>(g);
}

extern "C" int lhrs_rmsnorm_bwd(const void* x, const void* w, const float* rstd, const void* dy, const void* dres, void* dx,
                                int64_t rows, int32_t dim, void* stream) {
    LHRS_CHECK_ARG(x && w && rstd && dy && dx && rows > 0 && dim % 8 == 0 && dim <= 8 * BT * MAXC, "lhrs_rmsnorm_bwd: bad args");
    rmsnorm_bwd_kernel<<<(unsigned)rows, BT, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)w, rstd, (const bf16*)dy,
                                                                       (const bf16*)dres, (bf16*)dx, dim);
    LHRS_LAUNCH_CHECK("rmsnorm_bwd_kernel");
    return LHRS_OK;
}

extern "C" size_t lhrs_layernorm_bwd_scratch_bytes(int32_t dim) { return (size_t)2 * 296 * dim * sizeof(float); }

extern "C" int lhrs_layernorm_bwd(const void* x, int64_t ldx, const void* w, const float* mean, const float* rstd, const void* dy,
                                  const void* dres, void* dx, void* dw, void* db, int32_t accumulate, float* scratch,
                                  int64_t rows, int32_t dim, void* stream) {
    LHRS_CHECK_ARG(x && w && mean && rstd && dy && scratch && rows > 0 && dim % 8 == 0 && dim <= 8 * BT * MAXC, "lhrs_layernorm_bwd: bad args");
    unsigned nb = (unsigned)(rows < 296 ? rows : 296);
    const bool warp_form = (dim == 256 || dim == 512 || dim == 768 || dim == 1024) && ldx % 8 == 0 &&
                           ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(w)) & 15) == 0;
    if (warp_form) {
        const long long need = (rows + 7) / 8;               // blocks of eight row-warps
        nb = (unsigned)(need < 296 ? need : 296);
        auto launch = [&](auto kern) {
            kern<<<nb, BT, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (const bf16*)w, mean, rstd, (const bf16*)dy, (const bf16*)dres,
                                                      (bf16*)dx, scratch, rows, dim);
        };
        if (dim == 256) launch(layernorm_bwd_warp_kernel<1>);
        else if (dim == 512) launch(layernorm_bwd_warp_kernel<2>);
        else if (dim == 768) launch(layernorm_bwd_warp_kernel<3>);
        else launch(layernorm_bwd_warp_kernel<4>);
    } else {
        layernorm_bwd_kernel<<<nb, BT, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (const bf16*)w, mean, rstd, (const bf16*)dy,
                                                                 (const bf16*)dres, (bf16*)dx, scratch, rows, dim);
    }
    LHRS_LAUNCH_CHECK("layernorm_bwd_kernel");
    if (dw) {
        colsum_final_kernel<<<COLSUM_FINAL_GRID(dim), 256, 0, (cudaStream_t)stream>>>(scratch, nb, dim, (bf16*)dw, accumulate);
        LHRS_LAUNCH_CHECK("colsum_final_kernel");
    }
    if (db) {
        colsum_final_kernel<<<COLSUM_FINAL_GRID(dim), 256, 0, (cudaStream_t)stream>>>(scratch + (size_t)nb * dim, nb, dim, (bf16*)db, accumulate);
        LHRS_LAUNCH_CHECK("colsum_final_kernel");
    }
    return LHRS_OK;
}

extern "C" size_t lhrs_colsum_scratch_bytes(int32_t n) { return (size_t)64 * n * sizeof(float); }

extern "C" int lhrs_colsum(const void* a, int64_t ld, int64_t rows, int32_t n, void* out, int32_t accumulate, float* scratch,
                           void* stream) {
    LHRS_CHECK_ARG(a && out && scratch && rows > 0 && n % 8 == 0 && ld % 8 == 0, "lhrs_colsum: bad args");
    const unsigned ny = (unsigned)(rows < 64 ? rows : 64);
    colsum_partial_kernel<<<dim3((n / 8 + BT - 1) / BT, ny), BT, 0, (cudaStream_t)stream>>>((const bf16*)a, ld, rows, n, scratch);
    LHRS_LAUNCH_CHECK("colsum_partial_kernel");
    colsum_final_kernel<<<COLSUM_FINAL_GRID(n), 256, 0, (cudaStream_t)stream>>>(scratch, ny, n, (bf16*)out, accumulate);
    LHRS_LAUNCH_CHECK("colsum_final_kernel");
    return LHRS_OK;
}

// out = x with the elements dropped by the LoRA dropout mask of `key` zeroed (survivors unscaled: the callers fold 1/keep into
// the alpha of the product that consumes it).  One thread per 8 columns: four 2 x 2 blocks, one mask word each.
__global__ void __launch_bounds__(BT)
lora_dropout_mask_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, long long rows, int cols, uint32_t key, int t,
                         __nv_bfloat16* __restrict__ out, long long ldo) {
    const int cpr = cols / 8;
    const long long total = rows * cpr;
    for (long long i = blockIdx.x * static_cast<long long>(BT) + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * BT) {
        const long long r = i / cpr;
        const int c0 = static_cast<int>(i - r * cpr) * 8;
        uint4 v = *reinterpret_cast<const uint4*>(x + r * ldx + c0);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t word = drop_word(key, static_cast<uint32_t>(r), static_cast<uint32_t>(c0 + 2 * j), static_cast<uint32_t>(cols));
            if (!drop_keep(word, static_cast<uint32_t>(r), 0u, t)) w[j] &= 0xffff0000u;
            if (!drop_keep(word, static_cast<uint32_t>(r), 1u, t)) w[j] &= 0x0000ffffu;
        }
        *reinterpret_cast<uint4*>(out + r * ldo + c0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

extern "C" int lhrs_lora_dropout_mask(const void* x, int64_t ldx, int64_t rows, int32_t cols, uint64_t seed, int32_t module, float p,
                                      void* out, int64_t ldo, void* stream) {
    LHRS_CHECK_ARG(x && out && rows > 0 && cols > 0 && cols % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && p >= 0.f && p < 1.f,
                   "lhrs_lora_dropout_mask: bad args (cols, ldx, ldo must be multiples of 8; 0 <= p < 1)");
    lora_dropout_mask_kernel<<<grid_for(rows * (cols / 8), BT), BT, 0, (cudaStream_t)stream>>>(
        (const bf16*)x, ldx, rows, cols, drop_key(seed, static_cast<uint32_t>(module)), drop_threshold(p), (bf16*)out, ldo);
    LHRS_LAUNCH_CHECK("lora_dropout_mask_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_swiglu_bwd(const void* d_act, const void* pre_gate, const void* pre_up, void* d_gu, int64_t rows, int32_t f,
                               void* stream) {
    LHRS_CHECK_ARG(d_act && pre_gate && pre_up && d_gu && rows > 0 && f % 8 == 0, "lhrs_swiglu_bwd: bad args");
    swiglu_bwd_kernel<<<grid_for(rows * (f / 8), BT), BT, 0, (cudaStream_t)stream>>>((const bf16*)d_act, (const bf16*)pre_gate,
                                                                                  (const bf16*)pre_up, (bf16*)d_gu, rows, f);
    LHRS_LAUNCH_CHECK("swiglu_bwd_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_gelu_bwd(void* d, const void* pre, int64_t n, void* stream) {
    LHRS_CHECK_ARG(d && pre && n > 0 && n % 8 == 0, "lhrs_gelu_bwd: bad args");
    gelu_bwd_kernel<<<grid_for(n / 8, BT), BT, 0, (cudaStream_t)stream>>>((bf16*)d, (const bf16*)pre, n / 8);
    LHRS_LAUNCH_CHECK("gelu_bwd_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_rope_bwd(void* dqkv, int64_t ld, int64_t rows, int32_t dim, const float* cos, const float* sin,
                             const int32_t* positions, int32_t seq_len, void* stream) {
    LHRS_CHECK_ARG(dqkv && cos && sin && rows > 0 && dim % 128 == 0 && (positions || seq_len > 0), "lhrs_rope_bwd: bad args");
    rope_bwd_kernel<<<(unsigned)rows, BT, 0, (cudaStream_t)stream>>>((bf16*)dqkv, ld, dim, cos, sin, positions, seq_len);
    LHRS_LAUNCH_CHECK("rope_bwd_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_grad_sumsq(const void* g, int64_t n, float* out, float* scratch /*>= 1024 floats*/, void* stream) {
    LHRS_CHECK_ARG(g && out && scratch && n > 0, "lhrs_grad_sumsq: bad args");
    const unsigned nb = grid_for(n, BT * 8, 1024);
    sumsq_partial_kernel<<<nb, BT, 0, (cudaStream_t)stream>>>((const bf16*)g, n, scratch);
    LHRS_LAUNCH_CHECK("sumsq_partial_kernel");
    sumsq_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(scratch, nb, out);
    LHRS_LAUNCH_CHECK("sumsq_final_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_adamw_step(float* master, float* m, float* v, const void* grad, void* param_bf16, const float* decay_mask,
                               int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                               const float* gnorm_sq, float max_norm, float grad_scale, void* stream) {
    LHRS_CHECK_ARG(master && m && v && grad && param_bf16 && n > 0 && step >= 1, "lhrs_adamw_step: bad args");
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    adamw_kernel<<<grid_for(n, BT * 4), BT, 0, (cudaStream_t)stream>>>(master, m, v, (const bf16*)grad, (bf16*)param_bf16, decay_mask, n,
                                                                      lr, beta1, beta2, eps, weight_decay, bc1, bc2, gnorm_sq, max_norm,
                                                                      grad_scale);
    LHRS_LAUNCH_CHECK("adamw_kernel");
    return LHRS_OK;
}

