// Adan on flat buffers — the stage-1 optimizer of the reference (`optimizer: adanp`, Config/multi_modal_stage1.yaml:89 ->
// timm create_optimizer_v2(opt="adanp") = Adan(no_prox=False), lhrs/optimizer/build_optimizer.py:76-86).  One pass over the
// trainable set: 2 bytes of bf16 gradient in, 5 fp32 state streams read + written, 2 bytes of bf16 parameter out
// (= 44 bytes per parameter, HBM-bound).  Global-norm clipping and the 1/world factor of the summed allreduce are folded in,
// exactly as in adamw_kernel (backward_kernels.cu).
#include "host_common.h"
#include "ptx.cuh"

namespace lhrs {

constexpr int OPT_BT = 256;

// timm/optim/adan.py (timm==0.9.12, pyproject.toml:17; not vendored) restated:
// (betas in timm's convention, default (0.98, 0.92, 0.99); bc_k = 1 - beta_k^step)
//   diff   = g - g_prev                      (0 on the first step: pre_grad is initialised to the first gradient)
//   m     <- lerp(m, g, 1 - b1)              exp_avg
//   d     <- lerp(d, diff, 1 - b2)           exp_avg_diff
//   u      = g + b2 * diff
//   n     <- b3 * n + (1 - b3) * u^2         exp_avg_sq
//   denom  = sqrt(n) / sqrt(bc3) + eps
//   upd    = (m / bc1 + b2 * d / bc2) / denom
//   no_prox:  w <- w * (1 - lr*wd) - lr*upd          else:  w <- (w - lr*upd) / (1 + lr*wd)
__global__ void __launch_bounds__(OPT_BT)
adan_kernel(float* __restrict__ master, float* __restrict__ m, float* __restrict__ d, float* __restrict__ nsq,
            float* __restrict__ g_prev, const __nv_bfloat16* __restrict__ g, __nv_bfloat16* __restrict__ p_bf16,
            const float* __restrict__ decay_mask, long long n, float lr, float b1, float b2, float b3, float eps, float weight_decay,
            float bc1, float bc2, float sqrt_bc3, int first, int no_prox, const float* __restrict__ gnorm_sq, float max_norm,
            float grad_scale) {
    float clip = grad_scale;
    if (max_norm > 0.f && gnorm_sq != nullptr) {
        const float norm = sqrtf(gnorm_sq[0]) * grad_scale;
        if (norm > max_norm) clip *= max_norm / (norm + 1e-6f);
    }
    for (long long i = blockIdx.x * static_cast<long long>(OPT_BT) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * OPT_BT) {
        const float gi = __bfloat162float(g[i]) * clip;
        const float diff = first ? 0.f : gi - g_prev[i];
        const float mi = m[i] + (1.f - b1) * (gi - m[i]);    // torch lerp_: start + weight * (end - start)
        const float di = d[i] + (1.f - b2) * (diff - d[i]);
        const float u = gi + b2 * diff;
        const float ni = nsq[i] * b3 + (1.f - b3) * u * u;
        m[i] = mi; d[i] = di; nsq[i] = ni; g_prev[i] = gi;
        const float denom = sqrtf(ni) / sqrt_bc3 + eps;
        const float upd = (mi / bc1 + b2 * di / bc2) / denom;
        const float wd = decay_mask ? decay_mask[i] * weight_decay : weight_decay;
        float w = master[i];
        if (no_prox) w = w * (1.f - lr * wd) - lr * upd;
        else w = (w - lr * upd) / (1.f + lr * wd);
        master[i] = w;
        p_bf16[i] = __float2bfloat16_rn(w);
    }
}

}  // namespace lhrs

using namespace lhrs;

extern "C" int lhrs_adan_step(float* master, float* exp_avg, float* exp_avg_diff, float* exp_avg_sq, float* pre_grad, const void* grad,
                              void* param_bf16, const float* decay_mask, int64_t n, float lr, float beta1, float beta2, float beta3,
                              float eps, float weight_decay, int32_t step, int32_t no_prox, const float* gnorm_sq, float max_norm,
                              float grad_scale, void* stream) {
    LHRS_CHECK_ARG(master && exp_avg && exp_avg_diff && exp_avg_sq && pre_grad && grad && param_bf16 && n > 0 && step >= 1,
                   "lhrs_adan_step: bad args");
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    const float bc3 = 1.f - powf(beta3, (float)step);
    long long blocks = (n + OPT_BT * 4 - 1) / (OPT_BT * 4);
    const long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    adan_kernel<<<(unsigned)blocks, OPT_BT, 0, (cudaStream_t)stream>>>(master, exp_avg, exp_avg_diff, exp_avg_sq, pre_grad,
                                                                      (const __nv_bfloat16*)grad, (__nv_bfloat16*)param_bf16, decay_mask, n,
                                                                      lr, beta1, beta2, beta3, eps, weight_decay, bc1, bc2, sqrtf(bc3),
                                                                      step == 1, no_prox, gnorm_sq, max_norm, grad_scale);
    LHRS_LAUNCH_CHECK("adan_kernel");
    return LHRS_OK;
}
