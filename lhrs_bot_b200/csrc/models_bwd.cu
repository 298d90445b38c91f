// Backward layer loops, one C call per module:
//   lhrs_llama_bwd   dX-only backward of the frozen LLaMA stack (+ LoRA dA/dB), from the activation stash of lhrs_llama_fwd
//   lhrs_pooler_bwd  full backward of the AttnPooler (dX and every parameter gradient)
// Frozen weights need no dW, so the LLaMA backward is ~1x the forward FLOPs: each Linear's dX is one tcgen05 GEMM with the
// weight read MN-major in place (no transposed copies); concatenated outputs ([dq|dk|dv], [dgate|dup]) contract in ONE GEMM
// with the weight segments stacked along K.  dW of the pooler / LoRA factors are the TN form of the same kernel.
#include <stdlib.h>
#include "host_common.h"
#include "ptx.cuh"
#include "dropout.cuh"
#include "models_common.h"

namespace lhrs {

__device__ __forceinline__ void up8b(const uint4& u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}

// gradient of the learned queries = sum over the batch of (grad of the initial query state + grad of the query rows of kv)
__global__ void __launch_bounds__(128)
pooler_dquery_kernel(const __nv_bfloat16* __restrict__ dxq, const __nv_bfloat16* __restrict__ dkv, PoolerGeom geo, int B, int dim,
                     __nv_bfloat16* __restrict__ dquery) {
    const int qi = blockIdx.x;
    int g = 0;
    while (g + 1 < geo.G && qi >= geo.q_off[g + 1]) ++g;
    const int j = qi - geo.q_off[g];
    const int lkv = geo.stage[g] + geo.split[g];
    for (int c = threadIdx.x; c < dim / 8; c += blockDim.x) {
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int b = 0; b < B; ++b) {
            float a[8], k[8];
            up8b(reinterpret_cast<const uint4*>(dxq + (static_cast<long long>(B) * geo.q_off[g] + static_cast<long long>(b) * geo.stage[g] + j) * dim)[c], a);
            up8b(reinterpret_cast<const uint4*>(dkv + (static_cast<long long>(B) * geo.kv_off[g] + static_cast<long long>(b) * lkv + j) * dim)[c], k);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += a[e] + k[e];
        }
        uint4 u;
        u.x = pack_bf16(acc[0], acc[1]); u.y = pack_bf16(acc[2], acc[3]); u.z = pack_bf16(acc[4], acc[5]); u.w = pack_bf16(acc[6], acc[7]);
        reinterpret_cast<uint4*>(dquery + static_cast<long long>(qi) * dim)[c] = u;
    }
}

// image-token rows of the kv gradient back to the batch-major (B, img_total, dim) layout of image_embs
__global__ void __launch_bounds__(128)
pooler_dimage_kernel(const __nv_bfloat16* __restrict__ dkv, PoolerGeom geo, int B, int dim, __nv_bfloat16* __restrict__ dimg) {
    const long long r = blockIdx.x;
    const int b = r / geo.img_total, t = r % geo.img_total;
    int g = 0;
    while (g + 1 < geo.G && t >= geo.img_off[g + 1]) ++g;
    const int lkv = geo.stage[g] + geo.split[g];
    const long long src = static_cast<long long>(B) * geo.kv_off[g] + static_cast<long long>(b) * lkv + geo.stage[g] + (t - geo.img_off[g]);
    for (int c = threadIdx.x; c < dim / 8; c += blockDim.x)
        reinterpret_cast<uint4*>(dimg + r * dim)[c] = reinterpret_cast<const uint4*>(dkv + src * dim)[c];
}

// batch-major d_out rows -> group-major order (inverse of the forward's scatter epilogue)
__global__ void __launch_bounds__(128)
pooler_dout_gather_kernel(const __nv_bfloat16* __restrict__ dout, long long ldo, PoolerGeom geo, int B, int dim,
                          __nv_bfloat16* __restrict__ dst) {
    const int r = blockIdx.x;
    int g = 0;
    while (g + 1 < geo.G && r >= B * geo.q_off[g + 1]) ++g;
    const int rl = r - B * geo.q_off[g];
    const int b = rl / geo.stage[g], i = rl % geo.stage[g];
    const long long src = static_cast<long long>(b) * geo.nq + geo.q_off[g] + i;
    for (int c = threadIdx.x; c < dim / 8; c += blockDim.x)
        reinterpret_cast<uint4*>(dst + static_cast<long long>(r) * dim)[c] = reinterpret_cast<const uint4*>(dout + src * ldo)[c];
}

// ---- GEMM helpers for the three backward forms
// dX[M,N] = dY[M,K] · W[K,N]   (W is the nn.Linear weight [out=K, in=N], read MN-major in place)
static int gemm_dx(cudaStream_t st, long long M, int N, int K, const void* dY, long long ldy, const void* W, long long ldw, void* dX,
                   long long ldx, const void* residual = nullptr, float alpha = 1.f) {
    LhrsGemm g = gemm_desc(M, N, K, dY, ldy, W, ldw, dX, ldx);
    g.b_mn_major = 1; g.alpha = alpha;
    if (residual) { g.residual = residual; g.ldr = ldx; }
    return lhrs_gemm_bf16(&g, st);
}
// dW[out,in] = dY[rows,out]^T · X[rows,in]   (both operands read MN-major, reduction over rows)
static int gemm_dw(cudaStream_t st, int out, int in, long long rows, const void* dY, long long ldy, const void* X, long long ldx,
                   void* dW, long long ldw, float alpha = 1.f, float* skinny_scratch = nullptr) {
    LhrsGemm g = gemm_desc(out, in, static_cast<int>(rows), dY, ldy, X, ldx, dW, ldw);
    g.a_mn_major = 1; g.b_mn_major = 1; g.alpha = alpha;
    return skinny_gemm(g, skinny_scratch, st);   // plain launch unless the problem is skinny and scratch is given
}

// diagonal blocks of a [n*out, n*r] product -> the n separate [out, r] gradient tensors
__global__ void lora_extract_diag_kernel(const __nv_bfloat16* __restrict__ full, int n, int out_dim, int r,
                                         __nv_bfloat16* g0, __nv_bfloat16* g1, __nv_bfloat16* g2) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long per = static_cast<long long>(out_dim) * r;
    if (i >= per * n) return;
    const int p = i / per;
    const long long e = i % per;
    const int row = e / r, col = e % r;
    __nv_bfloat16* dst = (p == 0) ? g0 : (p == 1 ? g1 : g2);
    if (dst != nullptr) dst[e] = full[(static_cast<long long>(p) * out_dim + row) * (n * r) + p * r + col];
}

// K-major dX GEMMs (transposed weight copies, LhrsLlamaWeights::*_wt) take the LoRA K-extension dT · [A_0;A_1;..] with the A
// stack transposed too: At[in, n*r].  The factors change every step, so all of a backward's stacks are transposed up front in
// one launch per 32 layers (pointers travel as kernel parameters).
struct LoraATransposeArgs {
    const __nv_bfloat16* a[32 * 7];
    __nv_bfloat16* out;        // per layer: At_qkv [D, 3r] | At_o [D, r] | At_gu [D, 2r] | At_down [F, r]
    int layers, r, D, F;
};
__host__ __device__ inline long long lora_at_layer_elems(int r, int D, int F) { return (long long)D * 6 * r + (long long)F * r; }
__host__ __device__ inline long long lora_at_group_offset(int group, int r, int D) {   // group: 0 qkv, 1 o, 2 gate/up, 3 down
    return group == 0 ? 0 : (group == 1 ? (long long)D * 3 * r : (group == 2 ? (long long)D * 4 * r : (long long)D * 6 * r));
}
__global__ void __launch_bounds__(256) lora_a_transpose_kernel(const LoraATransposeArgs p) {
    const int mod = blockIdx.y;                       // layer * 7 + projection
    const int layer = mod / 7, proj = mod - layer * 7;
    const int group = proj < 3 ? 0 : (proj == 3 ? 1 : (proj < 6 ? 2 : 3));
    const int first = group == 0 ? 0 : (group == 1 ? 3 : (group == 2 ? 4 : 6));
    const int n = group == 0 ? 3 : (group == 2 ? 2 : 1);
    const int in_dim = (proj == 6) ? p.F : p.D;
    const __nv_bfloat16* a = p.a[mod];
    __nv_bfloat16* out = p.out + layer * lora_at_layer_elems(p.r, p.D, p.F) + lora_at_group_offset(group, p.r, p.D) + (proj - first) * p.r;
    const int ldo = n * p.r;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < in_dim; i += gridDim.x * blockDim.x)
        for (int j = 0; j < p.r; ++j) out[static_cast<long long>(i) * ldo + j] = a[static_cast<long long>(j) * in_dim + i];
}

// LoRA backward for `nproj` projections that share the input x:  y_p += (s * x A_p^T) B_p^T = T_p B_p^T
//   dT_p = s * dy_p B_p,   dx += dT_p A_p,   dA_p = dT_p^T x,   dB_p = dy_p^T T_p        (T kept from the forward pass)
// Grouped path (A factors and their gradients contiguous, dy_p column blocks of one matrix): dT for all projections is ONE
// block-diagonal K-segmented GEMM, `dx += dT A` rides on the main dX GEMM as a K-extension (no extra pass over dx), and
// dA / dB are one TN GEMM each.  lora_bwd_pre runs before the main dX GEMM `g`, lora_bwd_post after it.
struct LoraBwd {
    bool active = false, grouped = false, stream = false;
    int idx0 = 0, nproj = 0, in_dim = 0, out_dim = 0;
    const void* x = nullptr; long long ldx = 0;
    const void* dy[3] = {nullptr, nullptr, nullptr}; long long ldy = 0;
    const __nv_bfloat16* T = nullptr;   // [M, nproj*r] from the forward stash
    __nv_bfloat16* dt = nullptr;        // [M, nproj*r]
    float* scratch = nullptr;           // fp32 split-K accumulator for the skinny side GEMMs
    __nv_bfloat16* drop_x = nullptr;    // [M, in_dim] masked copy of x (LoRA dropout only)
    int drop_t = 0;
    const __nv_bfloat16* at = nullptr;  // K-major dX GEMM: the transposed A stack [in_dim, nproj*r] of this group
    long long M = 0;
};

static int lora_bwd_pre(cudaStream_t st, const LhrsLlamaWeights* w, void* const* ga, LoraBwd& L, LhrsGemm& g) {
    if (!L.active) return LHRS_OK;
    const int r = w->lora_r, n = L.nproj;
    const long long ldt = (long long)n * r;
    int rc;
    L.grouped = lora_a_adjacent(w->lora_a, L.idx0, n, (long long)r * L.in_dim) &&
                (ga == nullptr || lora_a_adjacent(ga, L.idx0, n, (long long)r * L.in_dim)) && (long long)n * L.out_dim == L.ldy;
    for (int p = 1; p < n && L.grouped; ++p)
        L.grouped = reinterpret_cast<const __nv_bfloat16*>(L.dy[p]) == reinterpret_cast<const __nv_bfloat16*>(L.dy[0]) + (long long)p * L.out_dim;
    if (!L.grouped) return LHRS_OK;
    L.stream = lora_stream_ok(w, L.in_dim, L.out_dim, n) && L.ldy % 8 == 0 && L.ldx % 8 == 0;
    if (L.stream) {   // dT = s * [dy_0|dy_1|..] · blockdiag(B_0, B_1, ..): one streaming pass over dy
        const void* wb[3] = {w->lora_b[L.idx0], n > 1 ? w->lora_b[L.idx0 + 1] : nullptr, n > 2 ? w->lora_b[L.idx0 + 2] : nullptr};
        if ((rc = lhrs_lora_panel(L.dy[0], L.ldy, L.M, n * L.out_dim, wb, n, 1, r, n * r, w->lora_scale, L.dt, ldt, st))) return rc;
    } else {   // dT = s * [dy_0|dy_1|..] · blockdiag(B_0, B_1, ..)
        LhrsGemm d = gemm_desc(L.M, n * r, n * L.out_dim, L.dy[0], L.ldy, w->lora_b[L.idx0], r, L.dt, ldt);
        d.b_mn_major = 1; d.num_b = n; d.alpha = w->lora_scale;
        for (int p = 1; p < n; ++p) d.B[p] = w->lora_b[L.idx0 + p];
        if (n > 1) d.b_seg_nshift = r;
        if ((rc = skinny_gemm(d, L.scratch, st))) return rc;
    }
    // dx += dT · [A_0;A_1;..] accumulates in the main GEMM's TMEM tile — unless dropout masks each projection's term
    // separately (then lora_bwd_post adds mask_p o (dT_p · A_p) per projection)
    if (L.drop_t == 0) {
        g.A2 = L.dt; g.lda2 = ldt; g.ext_k = n * r;
        if (g.b_mn_major) { g.B2[0] = w->lora_a[L.idx0]; g.ldb2 = L.in_dim; }      // [n*r, in] read MN-major in place
        else { LHRS_CHECK_ARG(L.at != nullptr, "lora_bwd_pre: K-major dX needs the transposed A stack"); g.B2[0] = L.at; g.ldb2 = ldt; }
    }
    return LHRS_OK;
}

static int lora_bwd_post(cudaStream_t st, const LhrsLlamaWeights* w, void* const* ga, void* const* gb, LoraBwd& L, void* dx,
                         long long lddx, __nv_bfloat16* diag_buf) {
    if (!L.active) return LHRS_OK;
    const int r = w->lora_r, n = L.nproj;
    const long long ldt = (long long)n * r;
    int rc;
    if (L.drop_t > 0 && L.grouped) {
        // peft input dropout (mask_p per LoRA module, regenerated from the call seed):
        //   dx  += (1/keep) * mask_p o (dT_p · A_p)        a K = r GEMM per projection with the mask in its epilogue
        //   dA_p = (1/keep) * dT_p^T · (mask_p o x)        from a masked copy of x
        //   dB_p = dy_p^T · T_p                            unchanged (T already carries the mask)
        const float inv = drop_inv_keep(L.drop_t);
        LHRS_CHECK_ARG(dx != nullptr && L.drop_x != nullptr, "lora_bwd_post: dropout needs the materialised dX and the masked-input scratch");
        const size_t sb = (size_t)8 * (size_t)(n * L.out_dim > L.in_dim ? n * L.out_dim : L.in_dim) * (size_t)(n * r) * sizeof(float);
        if (r == 16) {   // one streaming pass over dx for all projections of the group
            if ((rc = lhrs_lora_dx_dropout(dx, lddx, L.M, L.in_dim, L.dt, ldt, w->lora_a + L.idx0, n, w->lora_seed, L.idx0, w->lora_dropout, st))) return rc;
        }
        bool da_done = false;
        if (L.stream && ga != nullptr && ga[L.idx0]) {   // [dA_0;dA_1;..] = (1/keep) dT^T (mask_p o x): one pass over x, masks on the fragments
            if ((rc = lhrs_lora_rowreduce_dropout(L.x, L.ldx, L.M, L.in_dim, L.dt, ldt, n * r, ga[L.idx0], L.in_dim, w->lora_seed, L.idx0,
                                                  w->lora_dropout, L.scratch, sb, st))) return rc;
            da_done = true;
        }
        for (int p = 0; p < n; ++p) {
            const int idx = L.idx0 + p;
            if (r != 16) {
                LhrsGemm c = gemm_desc(L.M, L.in_dim, r, L.dt + p * r, ldt, w->lora_a[idx], L.in_dim, dx, lddx);
                c.b_mn_major = 1; c.alpha = inv; c.residual = dx; c.ldr = lddx;
                c.drop_key = drop_key(w->lora_seed, (uint32_t)idx); c.drop_t = L.drop_t;
                if ((rc = lhrs_gemm_bf16(&c, st))) return rc;
            }
            if (!da_done && ga != nullptr && ga[idx]) {
                if ((rc = lhrs_lora_dropout_mask(L.x, L.ldx, L.M, L.in_dim, w->lora_seed, idx, w->lora_dropout, L.drop_x, L.in_dim, st))) return rc;
                if (L.stream) {
                    void* dst[1] = {ga[idx]};
                    if ((rc = lhrs_lora_rowreduce(L.drop_x, L.in_dim, L.M, L.in_dim, L.dt + p * r, ldt, r, 0, 1, dst, L.in_dim, inv, L.scratch, sb, st))) return rc;
                } else {
                    if ((rc = gemm_dw(st, r, L.in_dim, L.M, L.dt + p * r, ldt, L.drop_x, L.in_dim, ga[idx], L.in_dim, inv, L.scratch))) return rc;
                }
            }
        }
        if (gb != nullptr) {
            if (L.stream) {
                void* dst[3] = {gb[L.idx0], n > 1 ? gb[L.idx0 + 1] : nullptr, n > 2 ? gb[L.idx0 + 2] : nullptr};
                if (dst[0] != nullptr)
                    if ((rc = lhrs_lora_rowreduce(L.dy[0], L.ldy, L.M, n * L.out_dim, L.T, ldt, r, L.out_dim, 0, dst, r, 1.f, L.scratch, sb, st))) return rc;
            } else {
                for (int p = 0; p < n; ++p)
                    if (gb[L.idx0 + p]) if ((rc = gemm_dw(st, L.out_dim, r, L.M, L.dy[p], L.ldy, L.T + p * r, ldt, gb[L.idx0 + p], r))) return rc;
            }
        }
        return LHRS_OK;
    }
    LHRS_CHECK_ARG(L.drop_t == 0, "lora backward with dropout needs the grouped (flat, SftStepper) LoRA parameter layout");
    if (L.grouped && L.stream) {
        const size_t sb = (size_t)8 * (size_t)(n * L.out_dim > L.in_dim ? n * L.out_dim : L.in_dim) * (size_t)(n * r) * sizeof(float);
        if (gb != nullptr) {   // dB_p = dy_p^T T_p: one streaming pass over dy, block-diagonal pairing of column blocks and T columns
            void* dst[3] = {gb[L.idx0], n > 1 ? gb[L.idx0 + 1] : nullptr, n > 2 ? gb[L.idx0 + 2] : nullptr};
            if (dst[0] != nullptr)
                if ((rc = lhrs_lora_rowreduce(L.dy[0], L.ldy, L.M, n * L.out_dim, L.T, ldt, r, L.out_dim, 0, dst, r, 1.f, L.scratch, sb, st))) return rc;
        }
        if (ga != nullptr && ga[L.idx0]) {   // [dA_0;dA_1;..] = dT^T x: one streaming pass over x
            void* dst[1] = {ga[L.idx0]};
            if ((rc = lhrs_lora_rowreduce(L.x, L.ldx, L.M, L.in_dim, L.dt, ldt, n * r, 0, 1, dst, L.in_dim, 1.f, L.scratch, sb, st))) return rc;
        }
        return LHRS_OK;
    }
    if (L.grouped) {
        if (gb != nullptr) {   // [dy_0|dy_1|..]^T T -> diagonal blocks are the dB_p
            if (n == 1) {
                if (gb[L.idx0]) if ((rc = gemm_dw(st, L.out_dim, r, L.M, L.dy[0], L.ldy, L.T, ldt, gb[L.idx0], r, 1.f, L.scratch))) return rc;
            } else {
                if ((rc = gemm_dw(st, n * L.out_dim, n * r, L.M, L.dy[0], L.ldy, L.T, ldt, diag_buf, ldt, 1.f, L.scratch))) return rc;
                const long long total = (long long)n * L.out_dim * r;
                lora_extract_diag_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
                    diag_buf, n, L.out_dim, r, (__nv_bfloat16*)gb[L.idx0], (__nv_bfloat16*)gb[L.idx0 + 1],
                    n > 2 ? (__nv_bfloat16*)gb[L.idx0 + 2] : nullptr);
                LHRS_LAUNCH_CHECK("lora_extract_diag_kernel");
            }
        }
        if (ga != nullptr && ga[L.idx0]) if ((rc = gemm_dw(st, n * r, L.in_dim, L.M, L.dt, ldt, L.x, L.ldx, ga[L.idx0], L.in_dim, 1.f, L.scratch))) return rc;
        return LHRS_OK;
    }
    for (int p = 0; p < n; ++p) {   // separate tensors: one projection at a time, dx accumulated in a read-modify-write pass
        const int idx = L.idx0 + p;
        if (gb && gb[idx]) if ((rc = gemm_dw(st, L.out_dim, r, L.M, L.dy[p], L.ldy, L.T + p * r, ldt, gb[idx], r))) return rc;
        if ((rc = gemm_dx(st, L.M, r, L.out_dim, L.dy[p], L.ldy, w->lora_b[idx], r, L.dt + p * r, ldt, nullptr, w->lora_scale))) return rc;
        if (ga && ga[idx]) if ((rc = gemm_dw(st, r, L.in_dim, L.M, L.dt + p * r, ldt, L.x, L.ldx, ga[idx], L.in_dim))) return rc;
        if (dx) if ((rc = gemm_dx(st, L.M, L.in_dim, r, L.dt + p * r, ldt, w->lora_a[idx], L.in_dim, dx, lddx, dx))) return rc;
    }
    return LHRS_OK;
}

static LoraBwd lora_ctx(const LhrsLlamaWeights* w, int layer, int first, int nproj, const void* x, long long ldx, int in_dim,
                        const void* dy0, long long ldy, int out_dim, long long M, const __nv_bfloat16* T, __nv_bfloat16* dt,
                        float* scratch, __nv_bfloat16* drop_x = nullptr) {
    LoraBwd L;
    L.scratch = scratch;
    L.drop_x = drop_x;
    L.drop_t = (w->lora_r > 0) ? drop_threshold(w->lora_dropout) : 0;
    L.active = w->lora_r > 0 && w->lora_a != nullptr && w->lora_b != nullptr;
    L.idx0 = layer * 7 + first; L.nproj = nproj; L.in_dim = in_dim; L.out_dim = out_dim;
    L.x = x; L.ldx = ldx; L.ldy = ldy; L.M = M; L.T = T; L.dt = dt;
    for (int p = 0; p < nproj; ++p) L.dy[p] = reinterpret_cast<const __nv_bfloat16*>(dy0) + (long long)p * out_dim;
    return L;
}

}  // namespace lhrs

using namespace lhrs;
typedef __nv_bfloat16 bf16;

// ================================================================================================ LLaMA backward
namespace {
struct LlamaBwdBufs { bf16 *dxa, *dxb, *dh, *d_act, *d_gu, *dqkv, *d_o, *h, *t, *dt, *diag, *drop_x, *at; float *delta, *skinny; };
LlamaBwdBufs llama_bwd_plan(Arena& a, const LhrsLlamaWeights* w, long long M) {
    LlamaBwdBufs b;
    const int D = w->dim, F = w->ffn;
    b.dxa = a.take<bf16>(M * D); b.dxb = a.take<bf16>(M * D); b.dh = a.take<bf16>(M * D);
    b.d_act = a.take<bf16>(M * F); b.d_gu = a.take<bf16>(M * 2 * F);
    b.dqkv = a.take<bf16>(M * 3 * D); b.d_o = a.take<bf16>(M * D);
    b.delta = a.take<float>(M * w->heads + M);   // + the key mask as bits (lhrs_attention_bwd_scratch_floats; M >= B * 2 * ceil(S/64))
    const bool lora = w->lora_r > 0;
    b.h = nullptr;   // normalised inputs now come from the stash (h1 / h2)
    b.t = lora ? a.take<bf16>(M * 3 * w->lora_r) : nullptr;
    b.dt = lora ? a.take<bf16>(M * 3 * w->lora_r) : nullptr;
    const long long widest = (3LL * w->dim > 2LL * w->ffn) ? 3LL * w->dim : 2LL * w->ffn;
    b.diag = lora ? a.take<bf16>(widest * 3 * w->lora_r) : nullptr;
    b.skinny = lora ? a.take<float>(skinny_scratch_elems(w, M)) : nullptr;
    b.drop_x = (lora && w->lora_dropout > 0.f) ? a.take<bf16>(M * (F > D ? F : D)) : nullptr;
    b.at = (lora && w->qkv_wt != nullptr) ? a.take<bf16>(w->num_layers * lora_at_layer_elems(w->lora_r, D, F)) : nullptr;
    return b;
}
}  // namespace

extern "C" size_t lhrs_llama_bwd_workspace_bytes(const LhrsLlamaWeights* w, int32_t B, int32_t S) {
    Arena a(nullptr, 0);
    llama_bwd_plan(a, w, (long long)B * S);
    return a.used();
}

extern "C" int lhrs_lm_head_bwd(const LhrsLlamaWeights* w, const void* d_logits, int64_t rows, void* d_hidden, void* stream) {
    LHRS_CHECK_ARG(w && d_logits && d_hidden && rows > 0, "lhrs_lm_head_bwd: null/empty");
    if (w->lm_head_wt != nullptr) {      // K-major form on the transposed copy [dim, vocab]
        LhrsGemm g = gemm_desc(rows, w->dim, w->vocab, d_logits, w->vocab, w->lm_head_wt, w->vocab, d_hidden, w->dim);
        return lhrs_gemm_bf16(&g, stream);
    }
    return gemm_dx((cudaStream_t)stream, rows, w->dim, w->vocab, d_logits, w->vocab, w->lm_head, w->dim, d_hidden, w->dim);
}

static int llama_bwd_impl(const LhrsLlamaWeights* w, void* const* lora_a_grads, void* const* lora_b_grads, const void* d_hidden,
                          int32_t B, int32_t S, long long M, const uint8_t* key_mask, const int32_t* seq_off, const void* stash,
                          void* d_inputs_embeds, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    LHRS_CHECK_ARG(w && d_hidden && stash && d_inputs_embeds && B > 0 && S > 0, "lhrs_llama_bwd: null/empty");
    const int D = w->dim, F = w->ffn;
    Arena a(workspace, workspace_bytes);
    LlamaBwdBufs b = llama_bwd_plan(a, w, (long long)B * S);   // (a ragged batch has fewer rows; delta / lse stay [B, heads, S])
    LHRS_CHECK_ARG(a.fits(), "lhrs_llama_bwd: workspace too small (%zu < %zu)", workspace_bytes, a.used());
    Arena sa(const_cast<void*>(stash), (size_t)-1);
    LlamaStash s = llama_stash_plan(sa, w, B, S);
    int rc;
    const bool kmaj = w->qkv_wt != nullptr && w->o_wt != nullptr && w->gu_wt != nullptr && w->down_wt != nullptr;
    const bool lora_on = w->lora_r > 0 && w->lora_a != nullptr && w->lora_b != nullptr;
    const long long at_layer = lora_on ? lora_at_layer_elems(w->lora_r, D, F) : 0;
    if (kmaj && lora_on && w->lora_dropout <= 0.f) {
        LHRS_CHECK_ARG(w->num_layers <= 32 * 8, "lhrs_llama_bwd: too many layers");
        for (int l0 = 0; l0 < w->num_layers; l0 += 32) {
            LoraATransposeArgs ta;
            memset(&ta, 0, sizeof(ta));
            ta.layers = (w->num_layers - l0 < 32) ? w->num_layers - l0 : 32;
            ta.r = w->lora_r; ta.D = D; ta.F = F; ta.out = b.at + l0 * at_layer;
            for (int i = 0; i < ta.layers * 7; ++i) ta.a[i] = reinterpret_cast<const bf16*>(w->lora_a[l0 * 7 + i]);
            lora_a_transpose_kernel<<<dim3((F + 255) / 256, ta.layers * 7), 256, 0, st>>>(ta);
            LHRS_LAUNCH_CHECK("lora_a_transpose_kernel");
        }
    }
    auto at_of = [&](int layer, int group) -> const bf16* {
        return (kmaj && lora_on) ? b.at + layer * at_layer + lora_at_group_offset(group, w->lora_r, D) : nullptr;
    };
    bf16* dx = b.dxa;
    bf16* dx_other = b.dxb;
    if ((rc = lhrs_rmsnorm_bwd(s.x_final, w->norm_w, s.rstd_final, d_hidden, nullptr, dx, M, D, st))) return rc;
    for (int l = w->num_layers - 1; l >= 0; --l) {
        const LlamaLayerStash& t = s.layer[l];
        // ---- MLP: x_out = x_mid + down(silu(gate(h2)) * up(h2)),  h2 = rmsnorm(x_mid)
        {
            LoraBwd L = lora_ctx(w, l, 6, 1, t.act, F, F, dx, D, D, M, t.lora_t[3], b.dt, b.skinny, b.drop_x);
            // d_act = dx · W_down (+ LoRA) with the SwiGLU backward applied in the epilogue: writes d_gu = [d_gate | d_up]
            LhrsGemm g = gemm_desc(M, F, D, dx, D, w->down_w[l], F, b.d_gu, 2 * F);
            g.b_mn_major = 1;
            if (kmaj) { g = gemm_desc(M, F, D, dx, D, w->down_wt[l], D, b.d_gu, 2 * F); L.at = at_of(l, 3); }   // Wdown^T [F, D]: K-major
            const char* fe = getenv("LHRS_FUSE_SWIGLU_BWD");   // read per call (the parity tests run both forms)
            const int fuse = fe ? atoi(fe) : 0;
            g.pre_gate = t.pre_gate; g.pre_up = t.pre_up;
            if ((rc = lora_bwd_pre(st, w, lora_a_grads, L, g))) return rc;
            if ((L.active && !L.grouped) || !fuse || L.drop_t > 0) {   // un-batched LoRA fallback needs d_act materialised for its read-modify-write pass
                g.D = b.d_act; g.ldd = F; g.pre_gate = nullptr; g.pre_up = nullptr;
                if ((rc = lhrs_gemm_bf16(&g, st))) return rc;
                if ((rc = lora_bwd_post(st, w, lora_a_grads, lora_b_grads, L, b.d_act, F, b.diag))) return rc;
                if ((rc = lhrs_swiglu_bwd(b.d_act, t.pre_gate, t.pre_up, b.d_gu, M, F, st))) return rc;
            } else {
                if ((rc = lhrs_gemm_bf16(&g, st))) return rc;
                if ((rc = lora_bwd_post(st, w, lora_a_grads, lora_b_grads, L, nullptr, F, b.diag))) return rc;
            }
        }
        {
            LoraBwd L = lora_ctx(w, l, 4, 2, t.h2, D, D, b.d_gu, 2 * F, F, M, t.lora_t[2], b.dt, b.skinny, b.drop_x);
            LhrsGemm g = gemm_desc(M, D, 2 * F, b.d_gu, 2 * F, w->gate_w[l], D, b.dh, D);
            g.b_mn_major = 1; g.B[1] = w->up_w[l]; g.num_b = 2;
            if (kmaj) { g = gemm_desc(M, D, 2 * F, b.d_gu, 2 * F, w->gu_wt[l], 2 * F, b.dh, D); L.at = at_of(l, 2); }   // [Wgate;Wup]^T [D, 2F]
            if ((rc = lora_bwd_pre(st, w, lora_a_grads, L, g))) return rc;
            if ((rc = lhrs_gemm_bf16(&g, st))) return rc;
            if ((rc = lora_bwd_post(st, w, lora_a_grads, lora_b_grads, L, b.dh, D, b.diag))) return rc;
        }
        if ((rc = lhrs_rmsnorm_bwd(t.x_mid, w->ln2_w[l], t.rstd2, b.dh, dx, dx_other, M, D, st))) return rc;
        { bf16* tmp = dx; dx = dx_other; dx_other = tmp; }   // dx = grad wrt x_mid
        // ---- attention: x_mid = x_in + o_proj(attn(rope(q), rope(k), v)),  q,k,v = proj(h1), h1 = rmsnorm(x_in)
        {
            LoraBwd L = lora_ctx(w, l, 3, 1, t.o, D, D, dx, D, D, M, t.lora_t[1], b.dt, b.skinny, b.drop_x);
            LhrsGemm g = gemm_desc(M, D, D, dx, D, w->o_w[l], D, b.d_o, D);
            g.b_mn_major = 1;
            if (kmaj) { g = gemm_desc(M, D, D, dx, D, w->o_wt[l], D, b.d_o, D); L.at = at_of(l, 1); }
            if ((rc = lora_bwd_pre(st, w, lora_a_grads, L, g))) return rc;
            if ((rc = lhrs_gemm_bf16(&g, st))) return rc;
            if ((rc = lora_bwd_post(st, w, lora_a_grads, lora_b_grads, L, b.d_o, D, b.diag))) return rc;
        }
        {
            LhrsAttentionBwd ab;
            memset(&ab, 0, sizeof(ab));
            ab.fwd = attn_desc(t.qkv, t.qkv + D, t.qkv + 2 * D, 3 * D, (long long)S * 3 * D, t.o, D, (long long)S * D, B, w->heads, S, S, 128, 1);
            ab.fwd.lse = t.lse; ab.fwd.key_mask = key_mask;
            ab.fwd.seq_off = seq_off; ab.fwd.total_rows = seq_off ? M : 0;
            ab.d_o = b.d_o; ab.delta = b.delta;
            ab.dq = b.dqkv; ab.dk = b.dqkv + D; ab.dv = b.dqkv + 2 * D;
            ab.dq_rs = ab.dk_rs = ab.dv_rs = 3 * D; ab.dq_bs = ab.dk_bs = ab.dv_bs = (long long)S * 3 * D;
            ab.dq_hs = ab.dk_hs = ab.dv_hs = 128;
            ab.rope_cos = w->rope_cos; ab.rope_sin = w->rope_sin;   // dq / dk leave the kernel already un-rotated
            if ((rc = lhrs_attention_bwd(&ab, st))) return rc;
        }
        {
            LoraBwd L = lora_ctx(w, l, 0, 3, t.h1, D, D, b.dqkv, 3 * D, D, M, t.lora_t[0], b.dt, b.skinny, b.drop_x);
            LhrsGemm g = gemm_desc(M, D, 3 * D, b.dqkv, 3 * D, w->q_w[l], D, b.dh, D);
            g.b_mn_major = 1; g.B[1] = w->k_w[l]; g.B[2] = w->v_w[l]; g.num_b = 3;
            if (kmaj) { g = gemm_desc(M, D, 3 * D, b.dqkv, 3 * D, w->qkv_wt[l], 3 * D, b.dh, D); L.at = at_of(l, 0); }   // [Wq;Wk;Wv]^T [D, 3D]
            if ((rc = lora_bwd_pre(st, w, lora_a_grads, L, g))) return rc;
            if ((rc = lhrs_gemm_bf16(&g, st))) return rc;
            if ((rc = lora_bwd_post(st, w, lora_a_grads, lora_b_grads, L, b.dh, D, b.diag))) return rc;
        }
        bf16* dst = (l == 0) ? reinterpret_cast<bf16*>(d_inputs_embeds) : dx_other;
        if ((rc = lhrs_rmsnorm_bwd(t.x_in, w->ln1_w[l], t.rstd1, b.dh, dx, dst, M, D, st))) return rc;
        { bf16* tmp = dx; dx = dx_other; dx_other = tmp; }
    }
    return LHRS_OK;
}

extern "C" int lhrs_llama_bwd(const LhrsLlamaWeights* w, void* const* lora_a_grads, void* const* lora_b_grads, const void* d_hidden,
                              int32_t B, int32_t S, const uint8_t* key_mask, const void* stash, void* d_inputs_embeds,
                              void* workspace, size_t workspace_bytes, void* stream_) {
    return llama_bwd_impl(w, lora_a_grads, lora_b_grads, d_hidden, B, S, (long long)B * S, key_mask, nullptr, stash, d_inputs_embeds,
                          workspace, workspace_bytes, stream_);
}

extern "C" int lhrs_llama_bwd_ragged(const LhrsLlamaWeights* w, void* const* lora_a_grads, void* const* lora_b_grads,
                                     const void* d_hidden, int32_t B, int32_t S_max, int64_t rows, const int32_t* seq_off,
                                     const void* stash, void* d_inputs_embeds, void* workspace, size_t workspace_bytes, void* stream_) {
    LHRS_CHECK_ARG(seq_off && rows > 0 && rows <= (long long)B * S_max && S_max >= 128, "lhrs_llama_bwd_ragged: seq_off / rows (<= B * S_max) / S_max >= 128");
    return llama_bwd_impl(w, lora_a_grads, lora_b_grads, d_hidden, B, S_max, rows, nullptr, seq_off, stash, d_inputs_embeds, workspace,
                          workspace_bytes, stream_);
}

// ================================================================================================ AttnPooler backward
namespace {
struct PoolBwdBufs { bf16 *dout_gm, *dx, *dx2, *d_f, *dh, *d_ao, *dqp, *dkvp, *dkvn, *dkv_raw; float *delta, *scratch; };
PoolBwdBufs pool_bwd_plan(Arena& a, const LhrsPoolerWeights* w, int B, const PoolerGeom& geo) {
    PoolBwdBufs p;
    const long long RQ = (long long)B * geo.nq, RKV = (long long)B * geo.kv_total;
    const int D = w->dim, F = w->ffn;
    p.dout_gm = a.take<bf16>(RQ * w->out_dim);
    p.dx = a.take<bf16>(RQ * D); p.dx2 = a.take<bf16>(RQ * D);
    p.d_f = a.take<bf16>(RQ * F); p.dh = a.take<bf16>(RQ * D); p.d_ao = a.take<bf16>(RQ * D); p.dqp = a.take<bf16>(RQ * D);
    p.dkvp = a.take<bf16>(RKV * 2 * D); p.dkvn = a.take<bf16>(RKV * D); p.dkv_raw = a.take<bf16>(RKV * D);
    p.delta = a.take<float>(RQ * w->heads + RQ);
    const int widest = F > w->out_dim ? F : w->out_dim;
    p.scratch = a.take<float>((long long)2 * 296 * (widest > 2 * D ? widest : 2 * D));
    return p;
}
}  // namespace

extern "C" size_t lhrs_pooler_bwd_workspace_bytes(const LhrsPoolerWeights* w, int32_t B) {
    Arena a(nullptr, 0);
    pool_bwd_plan(a, w, B, pooler_geom(w));
    return a.used();
}

extern "C" int lhrs_pooler_bwd(const LhrsPoolerWeights* w, const LhrsPoolerWeights* gr, const void* d_out, int64_t ldo, int32_t B,
                               const void* stash, void* d_image, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    LHRS_CHECK_ARG(w && gr && d_out && stash && B > 0, "lhrs_pooler_bwd: null/empty");
    const PoolerGeom geo = pooler_geom(w);
    Arena a(workspace, workspace_bytes);
    PoolBwdBufs p = pool_bwd_plan(a, w, B, geo);
    LHRS_CHECK_ARG(a.fits(), "lhrs_pooler_bwd: workspace too small (%zu < %zu)", workspace_bytes, a.used());
    Arena sa(const_cast<void*>(stash), (size_t)-1);
    PoolerStash s = pooler_stash_plan(sa, w, B, geo);
    const int D = w->dim, F = w->ffn, O = w->out_dim;
    const long long RQ = (long long)B * geo.nq, RKV = (long long)B * geo.kv_total;
    int rc;
    auto G = [](const void* p) { return const_cast<void*>(p); };

    // ---- out_proj: out = x_final W_out^T + b
    pooler_dout_gather_kernel<<<(unsigned)RQ, 128, 0, st>>>((const bf16*)d_out, ldo, geo, B, O, p.dout_gm);
    LHRS_LAUNCH_CHECK("pooler_dout_gather_kernel");
    if (gr->out_w) if ((rc = gemm_dw(st, O, D, RQ, p.dout_gm, O, s.x_final, D, G(gr->out_w), D))) return rc;
    if (gr->out_b) if ((rc = lhrs_colsum(p.dout_gm, O, RQ, O, G(gr->out_b), 0, p.scratch, st))) return rc;
    bf16* dx = p.dx;
    bf16* dx_other = p.dx2;
    if ((rc = gemm_dx(st, RQ, D, O, p.dout_gm, O, w->out_w, D, dx, D))) return rc;
    LHRS_CUDA(cudaMemsetAsync(p.dkv_raw, 0, RKV * D * 2, st));

    for (int l = w->num_layers - 1; l >= 0; --l) {
        const PoolerLayerStash& t = s.layer[l];
        const bf16* in_w = (const bf16*)w->in_w[l];
        bf16* g_in_w = (bf16*)G(gr->in_w ? gr->in_w[l] : nullptr);
        bf16* g_in_b = (bf16*)G(gr->in_b ? gr->in_b[l] : nullptr);
        // ---- MLP: x_out = x_mid + c_proj(gelu(c_fc(ln_2(x_mid))))
        if (gr->pj_w && gr->pj_w[l]) if ((rc = gemm_dw(st, D, F, RQ, dx, D, t.f_act, F, G(gr->pj_w[l]), F))) return rc;
        if (gr->pj_b && gr->pj_b[l]) if ((rc = lhrs_colsum(dx, D, RQ, D, G(gr->pj_b[l]), 0, p.scratch, st))) return rc;
        if ((rc = gemm_dx(st, RQ, F, D, dx, D, w->pj_w[l], F, p.d_f, F))) return rc;
        if ((rc = lhrs_gelu_bwd(p.d_f, t.f_pre, RQ * F, st))) return rc;
        if (gr->fc_w && gr->fc_w[l]) if ((rc = gemm_dw(st, F, D, RQ, p.d_f, F, t.h2, D, G(gr->fc_w[l]), D))) return rc;
        if (gr->fc_b && gr->fc_b[l]) if ((rc = lhrs_colsum(p.d_f, F, RQ, F, G(gr->fc_b[l]), 0, p.scratch, st))) return rc;
        if ((rc = gemm_dx(st, RQ, D, F, p.d_f, F, w->fc_w[l], D, p.dh, D))) return rc;
        if ((rc = lhrs_layernorm_bwd(t.x_mid, D, w->ln2_w[l], t.m_mean, t.m_rstd, p.dh, dx, dx_other, G(gr->ln2_w ? gr->ln2_w[l] : nullptr),
                                     G(gr->ln2_b ? gr->ln2_b[l] : nullptr), 0, p.scratch, RQ, D, st))) return rc;
        { bf16* tmp = dx; dx = dx_other; dx_other = tmp; }   // dx = grad wrt x_mid
        // ---- attention: x_mid = x_in + out_proj(MHA(q = ln_1(x_in) Wq, k/v = ln_1_kv(kv) Wk/Wv))
        if (gr->ao_w && gr->ao_w[l]) if ((rc = gemm_dw(st, D, D, RQ, dx, D, t.ao, D, G(gr->ao_w[l]), D))) return rc;
        if (gr->ao_b && gr->ao_b[l]) if ((rc = lhrs_colsum(dx, D, RQ, D, G(gr->ao_b[l]), 0, p.scratch, st))) return rc;
        if ((rc = gemm_dx(st, RQ, D, D, dx, D, w->ao_w[l], D, p.d_ao, D))) return rc;
        LhrsAttentionBwd abs_[4];
        for (int g = 0; g < geo.G; ++g) {
            const int Lq = geo.stage[g], Lkv = geo.stage[g] + geo.split[g];
            const long long qo = (long long)B * geo.q_off[g], ko = (long long)B * geo.kv_off[g];
            LhrsAttentionBwd& ab = abs_[g];
            memset(&ab, 0, sizeof(ab));
            ab.fwd = attn_desc(t.qp + qo * D, t.kvp + ko * 2 * D, t.kvp + ko * 2 * D + D, D, (long long)Lq * D, t.ao + qo * D, D,
                               (long long)Lq * D, B, w->heads, Lq, Lkv, 64, 0);
            ab.fwd.k_rs = ab.fwd.v_rs = 2 * D; ab.fwd.k_bs = ab.fwd.v_bs = (long long)Lkv * 2 * D;
            ab.fwd.lse = t.lse + qo * w->heads;
            ab.d_o = p.d_ao + qo * D; ab.delta = p.delta + qo * w->heads;
            ab.dq = p.dqp + qo * D; ab.dq_rs = D; ab.dq_bs = (long long)Lq * D; ab.dq_hs = 64;
            ab.dk = p.dkvp + ko * 2 * D; ab.dv = p.dkvp + ko * 2 * D + D;
            ab.dk_rs = ab.dv_rs = 2 * D; ab.dk_bs = ab.dv_bs = (long long)Lkv * 2 * D; ab.dk_hs = ab.dv_hs = 64;
        }
        if (geo.G <= 3) {      // delta, dQ and dK/dV of all query groups: three launches instead of nine
            if ((rc = lhrs_attention_bwd_grouped(abs_, geo.G, st))) return rc;
        } else {
            for (int g = 0; g < geo.G; ++g) if ((rc = lhrs_attention_bwd(&abs_[g], st))) return rc;
        }
        // q projection (in_proj rows [0, D)) and k/v projection (rows [D, 3D))
        if (g_in_w) if ((rc = gemm_dw(st, D, D, RQ, p.dqp, D, t.hq, D, g_in_w, D))) return rc;
        if (g_in_b) if ((rc = lhrs_colsum(p.dqp, D, RQ, D, g_in_b, 0, p.scratch, st))) return rc;
        if ((rc = gemm_dx(st, RQ, D, D, p.dqp, D, in_w, D, p.dh, D))) return rc;
        if (g_in_w) if ((rc = gemm_dw(st, 2 * D, D, RKV, p.dkvp, 2 * D, t.kvn, D, g_in_w + (long long)D * D, D))) return rc;
        if (g_in_b) if ((rc = lhrs_colsum(p.dkvp, 2 * D, RKV, 2 * D, g_in_b + D, 0, p.scratch, st))) return rc;
        if ((rc = gemm_dx(st, RKV, D, 2 * D, p.dkvp, 2 * D, in_w + (long long)D * D, D, p.dkvn, D))) return rc;
        if ((rc = lhrs_layernorm_bwd(t.x_in, D, w->ln1_w[l], t.q_mean, t.q_rstd, p.dh, dx, dx_other, G(gr->ln1_w ? gr->ln1_w[l] : nullptr),
                                     G(gr->ln1_b ? gr->ln1_b[l] : nullptr), 0, p.scratch, RQ, D, st))) return rc;
        { bf16* tmp = dx; dx = dx_other; dx_other = tmp; }   // dx = grad wrt x_in (the previous layer's output)
        // the kv input is the same tensor for every layer: accumulate its gradient in place
        if ((rc = lhrs_layernorm_bwd(s.kv_raw, D, w->lnkv_w[l], t.kv_mean, t.kv_rstd, p.dkvn, p.dkv_raw, p.dkv_raw,
                                     G(gr->lnkv_w ? gr->lnkv_w[l] : nullptr), G(gr->lnkv_b ? gr->lnkv_b[l] : nullptr), 0, p.scratch, RKV, D, st))) return rc;
    }
    if (gr->query) {
        pooler_dquery_kernel<<<geo.nq, 128, 0, st>>>(dx, p.dkv_raw, geo, B, D, (bf16*)G(gr->query));
        LHRS_LAUNCH_CHECK("pooler_dquery_kernel");
    }
    if (d_image != nullptr) {
        pooler_dimage_kernel<<<(unsigned)((long long)B * geo.img_total), 128, 0, st>>>(p.dkv_raw, geo, B, D, (bf16*)d_image);
        LHRS_LAUNCH_CHECK("pooler_dimage_kernel");
    }
    return LHRS_OK;
}
