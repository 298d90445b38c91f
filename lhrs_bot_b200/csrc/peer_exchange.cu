// Gradient exchange + optimizer step over NVLink peer memory.  The reference runs DeepSpeed ZeRO-2 here
// (main_pretrain_stage1.py:28-85, 215-220; hook/deepspeed_hook.py:5-9): gradients reduce-scattered, optimizer state sharded,
// updated parameters all-gathered — three NCCL collectives around a (CPU-offloaded, in stage 1) optimizer.  On an NVSwitch node
// every rank can load and store every peer's buffers directly, so the same schedule is two kernels over symmetric allocations:
//   p2p_reduce_kernel   rank r sums ITS 1/world slice of the flat bf16 gradient buffers of all ranks (16-byte peer loads, fixed
//                       rank order -> deterministic), keeps the fp32 sum, and publishes the slice's sum of squares to every peer
//   p2p_adamw_kernel    global-norm clip from the published partials, AdamW on the slice's fp32 master / moments (state is
//                       1/world per rank), and the bf16 result stored straight into EVERY rank's flat parameter buffer
// with a cross-rank barrier (caller's, e.g. torch symmetric-memory signal pads) before, between and after.  When the buffers also
// have NVLS multicast mappings the slice is reduced INSIDE the NVSwitch (multimem.ld_reduce: one load instead of one per peer)
// and the updated slice is broadcast with one multimem.st.  Frozen ViT / LLaMA
// weights never move.  Per rank and step at world 8 with 120 M trainable elements: 210 MB in + 210 MB out over NVLink and an
// optimizer pass over 15 M elements instead of 120 M.
#include "host_common.h"
#include "ptx.cuh"

namespace lhrs {

constexpr int PX_THREADS = 256;

struct PeerPtrs {
    const __nv_bfloat16* grads[16];
    __nv_bfloat16* params[16];
    float* norm_slots[16];
    const __nv_bfloat16* mc_grads;   // NVLS multicast mappings of the same buffers (or null): one load reduces in the switch,
    __nv_bfloat16* mc_params;        // one store lands in every rank's copy
    int world, rank;
    long long offset, n;      // this rank's slice, in elements (multiples of 8)
};

// NVSwitch in-network reduction: ONE load from the multicast address returns the sum over all ranks' copies (fp32 accumulation
// inside the switch, bf16 result), so a rank receives its slice once instead of once per peer.
__device__ __forceinline__ uint4 multimem_ld_reduce_bf16x8(const void* mc) {
    uint4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(mc) : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st_bf16x8(void* mc, const uint4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1,%2,%3,%4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

__global__ void __launch_bounds__(PX_THREADS)
p2p_reduce_kernel(const PeerPtrs p, float* __restrict__ gsum, float* __restrict__ partial) {
    __shared__ float red[PX_THREADS / 32];
    float ss = 0.f;
    const long long nvec = p.n / 8;
    for (long long i = blockIdx.x * static_cast<long long>(PX_THREADS) + threadIdx.x; i < nvec; i += static_cast<long long>(gridDim.x) * PX_THREADS) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (p.mc_grads != nullptr) {
            unpack8(multimem_ld_reduce_bf16x8(p.mc_grads + p.offset + i * 8), acc);
        } else {
            uint4 v[16];
#pragma unroll
            for (int r = 0; r < 16; ++r)                     // all peer loads in flight before the first add
                if (r < p.world) v[r] = *reinterpret_cast<const uint4*>(p.grads[r] + p.offset + i * 8);
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                if (r < p.world) {
                    float f[8];
                    unpack8(v[r], f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] += f[e];
                }
            }
        }
        *reinterpret_cast<float4*>(gsum + i * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(gsum + i * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += acc[e] * acc[e];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < PX_THREADS / 32; ++w) t += red[w];
        partial[blockIdx.x] = t;
    }
}

// one block: slice sum of squares -> slot [rank] of every rank's norm table
__global__ void __launch_bounds__(PX_THREADS)
p2p_publish_norm_kernel(const PeerPtrs p, const float* __restrict__ partial, int nparts) {
    __shared__ float red[PX_THREADS / 32];
    float t = 0.f;
    for (int i = threadIdx.x; i < nparts; i += PX_THREADS) t += partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < PX_THREADS / 32; ++w) s += red[w];
        for (int r = 0; r < p.world; ++r) p.norm_slots[r][p.rank] = s;
        __threadfence_system();
    }
}

__global__ void __launch_bounds__(PX_THREADS)
p2p_adamw_kernel(const PeerPtrs p, float* __restrict__ master, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ gsum,
                 const float* __restrict__ decay_mask, float lr, float beta1, float beta2, float eps, float weight_decay, float bc1,
                 float bc2, float max_norm, float grad_scale) {
    float clip = grad_scale;
    if (max_norm > 0.f) {
        float sq = 0.f;
        for (int r = 0; r < p.world; ++r) sq += p.norm_slots[p.rank][r];
        const float norm = sqrtf(sq) * grad_scale;
        if (norm > max_norm) clip *= max_norm / (norm + 1e-6f);
    }
    const long long nvec = p.n / 8;
    for (long long i = blockIdx.x * static_cast<long long>(PX_THREADS) + threadIdx.x; i < nvec; i += static_cast<long long>(gridDim.x) * PX_THREADS) {
        float w8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const long long j = i * 8 + e;
            const float gi = gsum[j] * clip;
            float w = master[j];
            const float wd = decay_mask ? decay_mask[j] * weight_decay : weight_decay;
            w *= 1.f - lr * wd;
            const float mi = beta1 * m[j] + (1.f - beta1) * gi;
            const float vi = beta2 * v[j] + (1.f - beta2) * gi * gi;
            m[j] = mi; v[j] = vi;
            w -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
            master[j] = w;
            w8[e] = w;
        }
        uint4 o;
        o.x = pack_bf16(w8[0], w8[1]); o.y = pack_bf16(w8[2], w8[3]); o.z = pack_bf16(w8[4], w8[5]); o.w = pack_bf16(w8[6], w8[7]);
        if (p.mc_params != nullptr) multimem_st_bf16x8(p.mc_params + p.offset + i * 8, o);
        else for (int r = 0; r < p.world; ++r) *reinterpret_cast<uint4*>(p.params[r] + p.offset + i * 8) = o;
    }
    __threadfence_system();
}

// Adan (stage 1, `adanp` / `adanw`; same update as adan_kernel in optimizers.cu) on the rank's slice, result broadcast to every rank
__global__ void __launch_bounds__(PX_THREADS)
p2p_adan_kernel(const PeerPtrs p, float* __restrict__ master, float* __restrict__ m, float* __restrict__ d, float* __restrict__ nsq,
                float* __restrict__ g_prev, const float* __restrict__ gsum, const float* __restrict__ decay_mask, float lr, float b1,
                float b2, float b3, float eps, float weight_decay, float bc1, float bc2, float sqrt_bc3, int first, int no_prox,
                float max_norm, float grad_scale) {
    float clip = grad_scale;
    if (max_norm > 0.f) {
        float sq = 0.f;
        for (int r = 0; r < p.world; ++r) sq += p.norm_slots[p.rank][r];
        const float norm = sqrtf(sq) * grad_scale;
        if (norm > max_norm) clip *= max_norm / (norm + 1e-6f);
    }
    const long long nvec = p.n / 8;
    for (long long i = blockIdx.x * static_cast<long long>(PX_THREADS) + threadIdx.x; i < nvec; i += static_cast<long long>(gridDim.x) * PX_THREADS) {
        float w8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const long long j = i * 8 + e;
            const float gi = gsum[j] * clip;
            const float diff = first ? 0.f : gi - g_prev[j];
            const float mi = m[j] + (1.f - b1) * (gi - m[j]);
            const float di = d[j] + (1.f - b2) * (diff - d[j]);
            const float u = gi + b2 * diff;
            const float ni = nsq[j] * b3 + (1.f - b3) * u * u;
            m[j] = mi; d[j] = di; nsq[j] = ni; g_prev[j] = gi;
            const float denom = sqrtf(ni) / sqrt_bc3 + eps;
            const float upd = (mi / bc1 + b2 * di / bc2) / denom;
            const float wd = decay_mask ? decay_mask[j] * weight_decay : weight_decay;
            float w = master[j];
            if (no_prox) w = w * (1.f - lr * wd) - lr * upd;
            else w = (w - lr * upd) / (1.f + lr * wd);
            master[j] = w;
            w8[e] = w;
        }
        uint4 o;
        o.x = pack_bf16(w8[0], w8[1]); o.y = pack_bf16(w8[2], w8[3]); o.z = pack_bf16(w8[4], w8[5]); o.w = pack_bf16(w8[6], w8[7]);
        if (p.mc_params != nullptr) multimem_st_bf16x8(p.mc_params + p.offset + i * 8, o);
        else for (int r = 0; r < p.world; ++r) *reinterpret_cast<uint4*>(p.params[r] + p.offset + i * 8) = o;
    }
    __threadfence_system();
}

static int fill(PeerPtrs& p, const LhrsPeerExchange* x, const char* who) {
    LHRS_CHECK_ARG(x && x->world >= 1 && x->world <= 16 && x->rank >= 0 && x->rank < x->world, "%s: bad world/rank", who);
    LHRS_CHECK_ARG(x->slice_n > 0 && x->slice_n % 8 == 0 && x->slice_offset % 8 == 0, "%s: the slice must be a multiple of 8 elements", who);
    memset(&p, 0, sizeof(p));
    for (int r = 0; r < x->world; ++r) {
        LHRS_CHECK_ARG(x->grads[r] && x->params[r] && x->norm_slots[r], "%s: null peer pointer for rank %d", who, r);
        LHRS_CHECK_ARG((reinterpret_cast<uintptr_t>(x->grads[r]) & 15) == 0 && (reinterpret_cast<uintptr_t>(x->params[r]) & 15) == 0, "%s: peer buffers must be 16-byte aligned", who);
        p.grads[r] = (const __nv_bfloat16*)x->grads[r]; p.params[r] = (__nv_bfloat16*)x->params[r]; p.norm_slots[r] = x->norm_slots[r];
    }
    p.mc_grads = (const __nv_bfloat16*)x->mc_grads; p.mc_params = (__nv_bfloat16*)x->mc_params;
    LHRS_CHECK_ARG((reinterpret_cast<uintptr_t>(x->mc_grads) & 15) == 0 && (reinterpret_cast<uintptr_t>(x->mc_params) & 15) == 0, "%s: multicast mappings must be 16-byte aligned", who);
    p.world = x->world; p.rank = x->rank; p.offset = x->slice_offset; p.n = x->slice_n;
    return LHRS_OK;
}

}  // namespace lhrs

using namespace lhrs;

extern "C" int lhrs_p2p_reduce_slice(const LhrsPeerExchange* x, float* grad_sum, float* scratch, void* stream) {
    PeerPtrs p;
    if (fill(p, x, "lhrs_p2p_reduce_slice")) return LHRS_ERR_INVALID;
    LHRS_CHECK_ARG(grad_sum && scratch, "lhrs_p2p_reduce_slice: null buffer");
    long long blocks = (p.n / 8 + PX_THREADS - 1) / PX_THREADS;
    if (blocks > 1024) blocks = 1024;
    p2p_reduce_kernel<<<(unsigned)blocks, PX_THREADS, 0, (cudaStream_t)stream>>>(p, grad_sum, scratch);
    LHRS_LAUNCH_CHECK("p2p_reduce_kernel");
    p2p_publish_norm_kernel<<<1, PX_THREADS, 0, (cudaStream_t)stream>>>(p, scratch, (int)blocks);
    LHRS_LAUNCH_CHECK("p2p_publish_norm_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_p2p_adamw_slice(const LhrsPeerExchange* x, float* master, float* m, float* v, const float* grad_sum,
                                    const float* decay_mask, float lr, float beta1, float beta2, float eps, float weight_decay,
                                    int32_t step, float max_norm, float grad_scale, void* stream) {
    PeerPtrs p;
    if (fill(p, x, "lhrs_p2p_adamw_slice")) return LHRS_ERR_INVALID;
    LHRS_CHECK_ARG(master && m && v && grad_sum && step >= 1, "lhrs_p2p_adamw_slice: bad args");
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    long long blocks = (p.n / 8 + PX_THREADS - 1) / PX_THREADS;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    p2p_adamw_kernel<<<(unsigned)blocks, PX_THREADS, 0, (cudaStream_t)stream>>>(p, master, m, v, grad_sum, decay_mask, lr, beta1, beta2, eps,
                                                                               weight_decay, bc1, bc2, max_norm, grad_scale);
    LHRS_LAUNCH_CHECK("p2p_adamw_kernel");
    return LHRS_OK;
}

extern "C" int lhrs_p2p_adan_slice(const LhrsPeerExchange* x, float* master, float* exp_avg, float* exp_avg_diff, float* exp_avg_sq,
                                   float* pre_grad, const float* grad_sum, const float* decay_mask, float lr, float beta1, float beta2,
                                   float beta3, float eps, float weight_decay, int32_t step, int32_t no_prox, float max_norm,
                                   float grad_scale, void* stream) {
    PeerPtrs p;
    if (fill(p, x, "lhrs_p2p_adan_slice")) return LHRS_ERR_INVALID;
    LHRS_CHECK_ARG(master && exp_avg && exp_avg_diff && exp_avg_sq && pre_grad && grad_sum && step >= 1, "lhrs_p2p_adan_slice: bad args");
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step), bc3 = 1.f - powf(beta3, (float)step);
    long long blocks = (p.n / 8 + PX_THREADS - 1) / PX_THREADS;
    const long long cap = (long long)num_sms() * 8;
    if (blocks > cap) blocks = cap;
    p2p_adan_kernel<<<(unsigned)blocks, PX_THREADS, 0, (cudaStream_t)stream>>>(p, master, exp_avg, exp_avg_diff, exp_avg_sq, pre_grad, grad_sum,
                                                                              decay_mask, lr, beta1, beta2, beta3, eps, weight_decay, bc1, bc2,
                                                                              sqrtf(bc3), step == 1, no_prox, max_norm, grad_scale);
    LHRS_LAUNCH_CHECK("p2p_adan_kernel");
    return LHRS_OK;
}
