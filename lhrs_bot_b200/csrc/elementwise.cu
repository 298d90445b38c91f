// HBM-bound row kernels of the hot path: RMSNorm / LayerNorm, CLIP patch extraction + embedding assembly,
// the multimodal embed splice (integer-exact restatement of text_modal.py:296-526) and the shifted cross-entropy.
// All of them are one pass over their rows with 16-byte vector accesses; none is worth a tensor core.
#include "host_common.h"
#include "ptx.cuh"

namespace lhrs {

constexpr int NORM_THREADS = 256;
constexpr int NORM_MAX_CHUNKS = 4;  // 16-byte chunks per thread kept in registers -> dim <= 8192

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum for NORM_THREADS threads; result broadcast to all threads
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < (blockDim.x >> 5)) ? red[l] : 0.f;
    t = warp_sum(t);
    return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (l < (blockDim.x >> 5)) ? red[l] : -INFINITY;
    t = warp_max(t);
    return t;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 u;
    u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]);
    u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
    return u;
}

// ------------------------------------------------------------------ RMSNorm (HF LlamaRMSNorm)
__global__ void __launch_bounds__(NORM_THREADS)
rmsnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                   __nv_bfloat16* __restrict__ y, float* __restrict__ rstd_out, int dim, float eps) {
    __shared__ float red[32];
    const long long row = blockIdx.x;
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * dim);
    const uint4* wr = reinterpret_cast<const uint4*>(w);
    uint4* yr = reinterpret_cast<uint4*>(y + row * dim);
    const int nchunks = dim / 8;
    uint4 xv[NORM_MAX_CHUNKS];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_CHUNKS; ++i) {
        const int c = threadIdx.x + i * NORM_THREADS;
        if (c < nchunks) {
            xv[i] = xr[c];
            float f[8];
            unpack8(xv[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
        }
    }
    ss = block_sum(ss, red);
    const float rstd = rsqrtf(ss / dim + eps);
    if (rstd_out != nullptr && threadIdx.x == 0) rstd_out[row] = rstd;
#pragma unroll
    for (int i = 0; i < NORM_MAX_CHUNKS; ++i) {
        const int c = threadIdx.x + i * NORM_THREADS;
        if (c < nchunks) {
            float f[8], g[8];
            unpack8(xv[i], f);
            unpack8(__ldg(wr + c), g);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = g[j] * bf16_round(f[j] * rstd);
            yr[c] = pack8(f);
        }
    }
}

// ------------------------------------------------------------------ LayerNorm (fp32 statistics)
__global__ void __launch_bounds__(NORM_THREADS)
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ w,
                     const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y, long long ldy,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out, int dim, float eps) {
    __shared__ float red[32];
    const long long row = blockIdx.x;
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
    uint4* yr = reinterpret_cast<uint4*>(y + row * ldy);
    const int nchunks = dim / 8;
    uint4 xv[NORM_MAX_CHUNKS];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_CHUNKS; ++i) {
        const int c = threadIdx.x + i * NORM_THREADS;
        if (c < nchunks) {
            xv[i] = xr[c];
            float f[8];
            unpack8(xv[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += f[j];
        }
    }
    const float mean = block_sum(s, red) / dim;
    float vs = 0.f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_CHUNKS; ++i) {
        const int c = threadIdx.x + i * NORM_THREADS;
        if (c < nchunks) {
            float f[8];
            unpack8(xv[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = f[j] - mean; vs += d * d; }
        }
    }
    const float rstd = rsqrtf(block_sum(vs, red) / dim + eps);
    if (threadIdx.x == 0) {
        if (mean_out != nullptr) mean_out[row] = mean;
        if (rstd_out != nullptr) rstd_out[row] = rstd;
    }
    const uint4* wr = reinterpret_cast<const uint4*>(w);
    const uint4* br = reinterpret_cast<const uint4*>(b);
#pragma unroll
    for (int i = 0; i < NORM_MAX_CHUNKS; ++i) {
        const int c = threadIdx.x + i * NORM_THREADS;
        if (c < nchunks) {
            float f[8], g[8], h[8];
            unpack8(xv[i], f);
            unpack8(__ldg(wr + c), g);
            unpack8(__ldg(br + c), h);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * rstd * g[j] + h[j];
            yr[c] = pack8(f);
        }
    }
}

// ------------------------------------------------------------------ CLIP patch extraction (conv-as-GEMM operand)
__global__ void vit_im2col_kernel(const __nv_bfloat16* __restrict__ px, __nv_bfloat16* __restrict__ out, int H, int W,
                                  int P, int kpad) {
    // one block per patch; threads over the kpad output columns (c, kh, kw)
    const int gw = W / P, gh = H / P;
    const long long patch = blockIdx.x;
    const int b = patch / (gh * gw);
    const int pr = patch % (gh * gw);
    const int py = pr / gw, pxi = pr % gw;
    const int kk = 3 * P * P;
    for (int col = threadIdx.x; col < kpad; col += blockDim.x) {
        __nv_bfloat16 v = __float2bfloat16(0.f);
        if (col < kk) {
            const int c = col / (P * P);
            const int r = col % (P * P);
            const int kh = r / P, kw = r % P;
            v = px[((static_cast<long long>(b) * 3 + c) * H + (py * P + kh)) * W + (pxi * P + kw)];
        }
        out[patch * kpad + col] = v;
    }
}

// tokens[b,0] = cls + pos[0]; tokens[b,1+p] = patch[b,p] + pos[1+p]; LayerNorm (pre_layrnorm) in the same pass
__global__ void __launch_bounds__(NORM_THREADS)
vit_embed_ln_kernel(const __nv_bfloat16* __restrict__ patch_emb, const __nv_bfloat16* __restrict__ cls,
                    const __nv_bfloat16* __restrict__ pos, const __nv_bfloat16* __restrict__ w,
                    const __nv_bfloat16* __restrict__ bb, __nv_bfloat16* __restrict__ tokens, int num_patches, int dim,
                    float eps) {
    __shared__ float red[32];
    const int T = num_patches + 1;
    const long long row = blockIdx.x;
    const int b = row / T, t = row % T;
    const uint4* src = (t == 0) ? reinterpret_cast<const uint4*>(cls)
                                : reinterpret_cast<const uint4*>(patch_emb + (static_cast<long long>(b) * num_patches + (t - 1)) * dim);
    const uint4* pr = reinterpret_cast<const uint4*>(pos + static_cast<long long>(t) * dim);
    uint4* yr = reinterpret_cast<uint4*>(tokens + row * dim);
    const int nchunks = dim / 8;
    float xv[NORM_MAX_CHUNKS][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_CHUNKS; ++i) {
        const int c = threadIdx.x + i * NORM_THREADS;
        if (c < nchunks) {
            float f[8], g[8];
            unpack8(src[c], f);
            unpack8(__ldg(pr + c), g);
#pragma unroll
            for (int j = 0; j < 8; ++j) { xv[i][j] = bf16_round(f[j] + g[j]); s += xv[i][j]; }
        }
    }
    const float mean = block_sum(s, red) / dim;
    float vs = 0.f;
#pragma unroll
    for (int i = 0; i < NORM_MAX_CHUNKS; ++i) {
        const int c = threadIdx.x + i * NORM_THREADS;
        if (c < nchunks) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float d = xv[i][j] - mean; vs += d * d; }
        }
    }
    const float rstd = rsqrtf(block_sum(vs, red) / dim + eps);
    const uint4* wr = reinterpret_cast<const uint4*>(w);
    const uint4* br = reinterpret_cast<const uint4*>(bb);
#pragma unroll
    for (int i = 0; i < NORM_MAX_CHUNKS; ++i) {
        const int c = threadIdx.x + i * NORM_THREADS;
        if (c < nchunks) {
            float g[8], h[8], f[8];
            unpack8(__ldg(wr + c), g);
            unpack8(__ldg(br + c), h);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = (xv[i][j] - mean) * rstd * g[j] + h[j];
            yr[c] = pack8(f);
        }
    }
}

// ------------------------------------------------------------------ multimodal splice
constexpr long long IMAGE_TOKEN = -200;
constexpr long long IGNORE_LABEL = -100;

// info layout (int32): [ (B+1) x 4 header | B x T exclusive image-token counts ]
//   header[b] = {n_img, new_len, slot_base, 0};  header[B] = {total_slots, max_len, min_len, 0}
__global__ void splice_scan_kernel(const long long* __restrict__ ids, int B, int T, int nq, int* __restrict__ info) {
    int* header = info;
    int* cnt = info + (B + 1) * 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int b = warp; b < B; b += nwarps) {
        int running = 0;
        for (int t0 = 0; t0 < T; t0 += 32) {
            const int t = t0 + lane;
            const bool is_img = (t < T) && (ids[static_cast<long long>(b) * T + t] == IMAGE_TOKEN);
            const unsigned m = __ballot_sync(0xffffffffu, is_img);
            if (t < T) cnt[b * T + t] = running + __popc(m & ((1u << lane) - 1u));
            running += __popc(m);
        }
        if (lane == 0) {
            header[b * 4 + 0] = running;
            header[b * 4 + 1] = T + running * (nq - 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int slots = 0, mx = 0, mn = 0x7fffffff;
        for (int b = 0; b < B; ++b) {
            header[b * 4 + 2] = slots;
            header[b * 4 + 3] = 0;
            const int n = header[b * 4 + 0];
            slots += (n > 0) ? n : 1;  // a text-only sample still consumes one image slot (text_modal.py:338)
            mx = max(mx, header[b * 4 + 1]);
            mn = min(mn, header[b * 4 + 1]);
        }
        header[B * 4 + 0] = slots;
        header[B * 4 + 1] = mx;
        header[B * 4 + 2] = mn;
        header[B * 4 + 3] = 0;
    }
}

__device__ __forceinline__ void copy_row16(uint4* dst, const uint4* src, int nchunks) {
    for (int c = threadIdx.x; c < nchunks; c += blockDim.x) dst[c] = src[c];
}

// grid (T + S_out + nq, B). Blocks x < T scatter input token x; blocks T <= x < T + S_out own output position s = x - T
// (padding rows + the attention mask, which the reference left-extends rather than aligning to tokens); blocks x >= T + S_out
// copy image-feature row q = x - T - S_out of every image token of the sample (one block per row: a single block walking all
// 144 rows of an image cost 165 us per call).
__global__ void __launch_bounds__(128)
splice_fill_kernel(const long long* __restrict__ ids, const long long* __restrict__ labels,
                   const uint8_t* __restrict__ amask, const int* __restrict__ info,
                   const __nv_bfloat16* __restrict__ table, const __nv_bfloat16* __restrict__ img, int B, int T,
                   int S_out, int nq, int dim, int n_slots, __nv_bfloat16* __restrict__ embeds,
                   long long* __restrict__ labels_out, uint8_t* __restrict__ mask_out, int* __restrict__ row_of_slot) {
    const int b = blockIdx.y;
    const int* header = info;
    const int* cnt = info + (B + 1) * 4;
    const int new_len = header[b * 4 + 1];
    const int slot_base = header[b * 4 + 2];
    const int nchunks = dim / 8;
    if (static_cast<int>(blockIdx.x) < T) {
        const int t = blockIdx.x;
        const long long id = ids[static_cast<long long>(b) * T + t];
        const int c = cnt[b * T + t];
        const int dst = t + c * (nq - 1);
        if (id != IMAGE_TOKEN) {
            copy_row16(reinterpret_cast<uint4*>(embeds + (static_cast<long long>(b) * S_out + dst) * dim),
                       reinterpret_cast<const uint4*>(table + id * dim), nchunks);
            if (labels_out != nullptr && threadIdx.x == 0)
                labels_out[static_cast<long long>(b) * S_out + dst] = labels[static_cast<long long>(b) * T + t];
        } else {
            const int slot = slot_base + c;
            if (slot < n_slots) {
                for (int q = threadIdx.x; q < nq; q += blockDim.x) {
                    if (labels_out != nullptr) labels_out[static_cast<long long>(b) * S_out + dst + q] = IGNORE_LABEL;
                    if (row_of_slot != nullptr) row_of_slot[slot * nq + q] = b * S_out + dst + q;
                }
            }
        }
    } else if (static_cast<int>(blockIdx.x) >= T + S_out) {
        if (img == nullptr) return;
        const int q = blockIdx.x - T - S_out;
        for (int t = 0; t < T; ++t) {                       // block-uniform scan: a sample holds very few image tokens
            if (ids[static_cast<long long>(b) * T + t] != IMAGE_TOKEN) continue;
            const int c = cnt[b * T + t];
            const int slot = slot_base + c;
            if (slot >= n_slots) continue;
            const int dst = t + c * (nq - 1);
            copy_row16(reinterpret_cast<uint4*>(embeds + (static_cast<long long>(b) * S_out + dst + q) * dim),
                       reinterpret_cast<const uint4*>(img + (static_cast<long long>(slot) * nq + q) * dim), nchunks);
        }
    } else {
        const int s = blockIdx.x - T;
        if (s >= new_len) {
            uint4* d = reinterpret_cast<uint4*>(embeds + (static_cast<long long>(b) * S_out + s) * dim);
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (int c = threadIdx.x; c < nchunks; c += blockDim.x) d[c] = z;
            if (labels_out != nullptr && threadIdx.x == 0) labels_out[static_cast<long long>(b) * S_out + s] = IGNORE_LABEL;
        }
        if (mask_out != nullptr && threadIdx.x == 0) {
            const int grown = new_len - T;  // True x (grown) | original mask | False x pad   (text_modal.py:474-497, 511-523)
            uint8_t m;
            if (s < grown) m = 1;
            else if (s < new_len) m = amask[static_cast<long long>(b) * T + (s - grown)];
            else m = 0;
            mask_out[static_cast<long long>(b) * S_out + s] = m;
        }
    }
}

__global__ void __launch_bounds__(128)
splice_bwd_kernel(const __nv_bfloat16* __restrict__ d_embeds, const int* __restrict__ row_of_slot,
                  __nv_bfloat16* __restrict__ d_img, int dim) {
    const long long r = blockIdx.x;
    const int src = row_of_slot[r];
    uint4* d = reinterpret_cast<uint4*>(d_img + r * dim);
    const int nchunks = dim / 8;
    if (src >= 0) {
        copy_row16(d, reinterpret_cast<const uint4*>(d_embeds + static_cast<long long>(src) * dim), nchunks);
    } else {
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int c = threadIdx.x; c < nchunks; c += blockDim.x) d[c] = z;
    }
}

// ------------------------------------------------------------------ shifted cross-entropy on bf16 logits
// row r = b*S + s predicts labels[b, s+1]; rows with s == S-1 or label == -100 contribute nothing.
__global__ void __launch_bounds__(NORM_THREADS)
ce_fwd_kernel(const __nv_bfloat16* __restrict__ logits, long long ld, const long long* __restrict__ labels, int S,
              int V, float* __restrict__ row_lse, float* __restrict__ row_loss) {
    __shared__ float red[32];
    const long long r = blockIdx.x;
    const int s = r % S;
    const long long tgt = (s + 1 < S) ? labels[r + 1] : IGNORE_LABEL;
    if (tgt == IGNORE_LABEL) {
        if (threadIdx.x == 0) { row_lse[r] = 0.f; row_loss[r] = 0.f; }
        return;
    }
    const uint4* lr = reinterpret_cast<const uint4*>(logits + r * ld);
    const int nchunks = V / 8;
    float m = -INFINITY, l = 0.f;
    for (int c = threadIdx.x; c < nchunks; c += NORM_THREADS) {
        float f[8];
        unpack8(__ldg(lr + c), f);
        float cm = f[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) cm = fmaxf(cm, f[j]);
        const float nm = fmaxf(m, cm);
        float add = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) add += __expf(f[j] - nm);
        l = l * __expf(m - nm) + add;
        m = nm;
    }
    for (int v = nchunks * 8 + threadIdx.x; v < V; v += NORM_THREADS) {  // V % 8 tail
        const float f = __bfloat162float(logits[r * ld + v]);
        const float nm = fmaxf(m, f);
        l = l * __expf(m - nm) + __expf(f - nm);
        m = nm;
    }
    const float gm = block_max(m, red);
    const float contrib = (m == -INFINITY) ? 0.f : l * __expf(m - gm);
    const float gl = block_sum(contrib, red);
    if (threadIdx.x == 0) {
        const float lse = gm + logf(gl);
        row_lse[r] = lse;
        row_loss[r] = lse - __bfloat162float(logits[r * ld + tgt]);
    }
}

// deterministic (fixed-order) reduction of the per-row losses
__global__ void __launch_bounds__(1024)
ce_reduce_kernel(const float* __restrict__ row_loss, const long long* __restrict__ labels, long long rows, int S,
                 float* __restrict__ loss_sum, int* __restrict__ count) {
    __shared__ float sl[1024];
    __shared__ int sc[1024];
    float acc = 0.f;
    int cnt = 0;
    for (long long r = threadIdx.x; r < rows; r += 1024) {
        const int s = r % S;
        const bool valid = (s + 1 < S) && (labels[r + 1] != IGNORE_LABEL);
        if (valid) { acc += row_loss[r]; cnt += 1; }
    }
    sl[threadIdx.x] = acc;
    sc[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sl[threadIdx.x] += sl[threadIdx.x + o]; sc[threadIdx.x] += sc[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { loss_sum[0] = sl[0]; count[0] = sc[0]; }
}

__global__ void __launch_bounds__(NORM_THREADS)
ce_bwd_kernel(const __nv_bfloat16* __restrict__ logits, long long ld, const long long* __restrict__ labels, int S,
              int V, const float* __restrict__ row_lse, const int* __restrict__ count, float grad_scale,
              const float* __restrict__ grad_scale_dev,
              __nv_bfloat16* __restrict__ d_logits) {
    const long long r = blockIdx.x;
    const int s = r % S;
    const long long tgt = (s + 1 < S) ? labels[r + 1] : IGNORE_LABEL;
    uint4* dr = reinterpret_cast<uint4*>(d_logits + r * ld);
    const int nchunks = V / 8;
    if (tgt == IGNORE_LABEL) {
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int c = threadIdx.x; c < nchunks; c += NORM_THREADS) dr[c] = z;
        return;
    }
    const uint4* lr = reinterpret_cast<const uint4*>(logits + r * ld);
    const float lse = row_lse[r];
    const int n = count[0];
    const float g = grad_scale * (grad_scale_dev ? grad_scale_dev[0] : 1.f) / static_cast<float>(n > 0 ? n : 1);
    for (int c = threadIdx.x; c < nchunks; c += NORM_THREADS) {
        float f[8];
        unpack8(__ldg(lr + c), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float p = __expf(f[j] - lse);
            if (c * 8 + j == tgt) p -= 1.f;
            f[j] = p * g;
        }
        dr[c] = pack8(f);
    }
}

__global__ void cast_f32_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, long long n4) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float4 v = src[i];
        dst[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
}

}  // namespace lhrs

using namespace lhrs;
typedef __nv_bfloat16 bf16;

extern "C" int lhrs_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream) {
    LHRS_CHECK_ARG(src && dst && n > 0 && n % 4 == 0, "lhrs_cast_f32_bf16: bad args");
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    cast_f32_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)src, (uint2*)dst, n / 4);
    LHRS_LAUNCH_CHECK("cast_f32_bf16_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd_out, int64_t rows, int32_t dim,
                                float eps, void* stream) {
    LHRS_CHECK_ARG(x && w && y && rows > 0, "lhrs_rmsnorm_fwd: null/empty");
    LHRS_CHECK_ARG(dim % 8 == 0 && dim <= 8 * NORM_THREADS * NORM_MAX_CHUNKS, "lhrs_rmsnorm_fwd: dim %d unsupported", dim);
    rmsnorm_fwd_kernel<<<static_cast<unsigned>(rows), NORM_THREADS, 0, (cudaStream_t)stream>>>(
        (const bf16*)x, (const bf16*)w, (bf16*)y, rstd_out, dim, eps);
    LHRS_LAUNCH_CHECK("rmsnorm_fwd_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_layernorm_fwd(const void* x, int64_t ldx, const void* w, const void* b, void* y, int64_t ldy,
                                  float* mean_out, float* rstd_out, int64_t rows, int32_t dim, float eps, void* stream) {
    LHRS_CHECK_ARG(x && w && b && y && rows > 0, "lhrs_layernorm_fwd: null/empty");
    LHRS_CHECK_ARG(dim % 8 == 0 && dim <= 8 * NORM_THREADS * NORM_MAX_CHUNKS && ldx % 8 == 0 && ldy % 8 == 0,
                   "lhrs_layernorm_fwd: dim %d / strides unsupported", dim);
    layernorm_fwd_kernel<<<static_cast<unsigned>(rows), NORM_THREADS, 0, (cudaStream_t)stream>>>(
        (const bf16*)x, ldx, (const bf16*)w, (const bf16*)b, (bf16*)y, ldy, mean_out, rstd_out, dim, eps);
    LHRS_LAUNCH_CHECK("layernorm_fwd_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_vit_im2col(const void* pixels, void* patches, int32_t B, int32_t H, int32_t W, int32_t P,
                               int32_t kpad, void* stream) {
    LHRS_CHECK_ARG(pixels && patches && B > 0, "lhrs_vit_im2col: null/empty");
    LHRS_CHECK_ARG(H % P == 0 && W % P == 0 && kpad >= 3 * P * P && kpad % 8 == 0, "lhrs_vit_im2col: bad geometry");
    const unsigned patches_n = static_cast<unsigned>(B) * (H / P) * (W / P);
    vit_im2col_kernel<<<patches_n, 128, 0, (cudaStream_t)stream>>>((const bf16*)pixels, (bf16*)patches, H, W, P, kpad);
    LHRS_LAUNCH_CHECK("vit_im2col_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_vit_embed_ln(const void* patch_emb, const void* cls, const void* pos, const void* ln_w,
                                 const void* ln_b, void* tokens, int32_t B, int32_t num_patches, int32_t dim, float eps,
                                 void* stream) {
    LHRS_CHECK_ARG(patch_emb && cls && pos && ln_w && ln_b && tokens && B > 0, "lhrs_vit_embed_ln: null/empty");
    LHRS_CHECK_ARG(dim % 8 == 0 && dim <= 8 * NORM_THREADS * NORM_MAX_CHUNKS, "lhrs_vit_embed_ln: dim %d unsupported", dim);
    vit_embed_ln_kernel<<<static_cast<unsigned>(B) * (num_patches + 1), NORM_THREADS, 0, (cudaStream_t)stream>>>(
        (const bf16*)patch_emb, (const bf16*)cls, (const bf16*)pos, (const bf16*)ln_w, (const bf16*)ln_b, (bf16*)tokens,
        num_patches, dim, eps);
    LHRS_LAUNCH_CHECK("vit_embed_ln_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_splice_scan(const int64_t* input_ids, int32_t B, int32_t T, int32_t num_query, int32_t* info,
                                void* stream) {
    LHRS_CHECK_ARG(input_ids && info && B > 0 && T > 0 && num_query > 0, "lhrs_splice_scan: null/empty");
    splice_scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>((const long long*)input_ids, B, T, num_query, info);
    LHRS_LAUNCH_CHECK("splice_scan_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_splice_fill(const int64_t* input_ids, const int64_t* labels, const uint8_t* attn_mask,
                                const int32_t* info, const void* embed_table, const void* image_feats, int32_t B,
                                int32_t T, int32_t S_out, int32_t num_query, int32_t dim, int32_t n_slots,
                                void* embeds_out, int64_t* labels_out, uint8_t* mask_out, int32_t* row_of_slot,
                                void* stream) {
    LHRS_CHECK_ARG(input_ids && info && embed_table && embeds_out, "lhrs_splice_fill: null operand");
    LHRS_CHECK_ARG((labels_out == nullptr) || (labels != nullptr), "lhrs_splice_fill: labels_out without labels");
    LHRS_CHECK_ARG((mask_out == nullptr) || (attn_mask != nullptr), "lhrs_splice_fill: mask_out without attention_mask");
    LHRS_CHECK_ARG(dim % 8 == 0 && S_out >= T, "lhrs_splice_fill: dim %d / S_out %d", dim, S_out);
    if (row_of_slot != nullptr && n_slots > 0)
        LHRS_CUDA(cudaMemsetAsync(row_of_slot, 0xFF, sizeof(int32_t) * (size_t)n_slots * num_query, (cudaStream_t)stream));
    dim3 grid(T + S_out + (image_feats != nullptr ? num_query : 0), B);
    splice_fill_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(
        (const long long*)input_ids, (const long long*)labels, attn_mask, info, (const bf16*)embed_table,
        (const bf16*)image_feats, B, T, S_out, num_query, dim, n_slots, (bf16*)embeds_out, (long long*)labels_out,
        mask_out, row_of_slot);
    LHRS_LAUNCH_CHECK("splice_fill_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_splice_bwd(const void* d_embeds, const int32_t* row_of_slot, void* d_image, int64_t n_rows,
                               int32_t dim, void* stream) {
    LHRS_CHECK_ARG(d_embeds && row_of_slot && d_image && n_rows > 0 && dim % 8 == 0, "lhrs_splice_bwd: bad args");
    splice_bwd_kernel<<<static_cast<unsigned>(n_rows), 128, 0, (cudaStream_t)stream>>>(
        (const bf16*)d_embeds, row_of_slot, (bf16*)d_image, dim);
    LHRS_LAUNCH_CHECK("splice_bwd_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_ce_fwd(const void* logits, int64_t ld, const int64_t* labels, int32_t B, int32_t S, int32_t V,
                           float* row_lse, float* loss_sum, int32_t* count, void* stream) {
    LHRS_CHECK_ARG(logits && labels && row_lse && loss_sum && count && B > 0 && S > 0 && V > 0, "lhrs_ce_fwd: bad args");
    LHRS_CHECK_ARG(ld % 8 == 0, "lhrs_ce_fwd: ld must be a multiple of 8");
    const long long rows = (long long)B * S;
    ce_fwd_kernel<<<static_cast<unsigned>(rows), NORM_THREADS, 0, (cudaStream_t)stream>>>(
        (const bf16*)logits, ld, (const long long*)labels, S, V, row_lse, row_lse + rows);
    LHRS_LAUNCH_CHECK("ce_fwd_kernel");
    ce_reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(row_lse + rows, (const long long*)labels, rows, S, loss_sum, count);
    LHRS_LAUNCH_CHECK("ce_reduce_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_ce_bwd(const void* logits, int64_t ld, const int64_t* labels, int32_t B, int32_t S, int32_t V,
                           const float* row_lse, const int32_t* count, float grad_scale, const float* grad_scale_dev,
                           void* d_logits, void* stream) {
    LHRS_CHECK_ARG(logits && labels && row_lse && count && d_logits, "lhrs_ce_bwd: null operand");
    LHRS_CHECK_ARG(ld % 8 == 0 && V % 8 == 0, "lhrs_ce_bwd: ld and V must be multiples of 8");
    const long long rows = (long long)B * S;
    ce_bwd_kernel<<<static_cast<unsigned>(rows), NORM_THREADS, 0, (cudaStream_t)stream>>>(
        (const bf16*)logits, ld, (const long long*)labels, S, V, row_lse, count, grad_scale, grad_scale_dev, (bf16*)d_logits);
    LHRS_LAUNCH_CHECK("ce_bwd_kernel");
    return LHRS_OK;
}
