// Shared between the attention translation units (mma.sync kernels in attention*.cu, tcgen05 kernels in attention*_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lhrs_b200.h"

namespace lhrs {

struct AttnBwdArgs {
    const __nv_bfloat16 *q, *k, *v, *o, *d_o;
    __nv_bfloat16 *dq, *dk, *dv;
    const float* lse;    // [B,H,Sq]
    float* delta;        // [B,H,Sq]
    const uint8_t* kmask;
    uint32_t* kbits;     // kmask packed to bits by the delta pre-pass: [B, kbits_w] words, word t = keys [32t, 32t+32) (tcgen05 path)
    int kbits_w;         // words per batch row (even, >= ceil(Skv / 32))
    const int* seq_off;  // ragged batch (LhrsAttention::seq_off) or nullptr; then *_bs are 0, Sq == Skv = longest length, kmask == nullptr
    long long total_rows;
    long long q_bs, q_rs, q_hs, k_bs, k_rs, k_hs, v_bs, v_rs, v_hs;
    long long o_bs, o_rs, o_hs;        // O and dO share a layout
    long long dq_bs, dq_rs, dq_hs, dk_bs, dk_rs, dk_hs, dv_bs, dv_rs, dv_hs;
    int B, H, Sq, Skv;
    float scale, scale_log2;
    const float* rope_cos;   // non-null (HD 128 only): inverse RoPE on dQ / dK at store time
    const float* rope_sin;
};

// 4-D bf16 TMA view {head_dim, rows | heads (smaller stride first), batch}; box = 64 dims x box_rows rows of one head, SW128.
// *hfirst tells the kernel which coordinate order the map expects (attention_tc.cu).
int make_tmap_bshd(CUtensorMap* out, int* hfirst, const void* ptr, int hd, int S, int H, int B, long long rs, long long hs,
                   long long bs, int box_rows);

bool use_tc_attention();                                                    // LHRS_ATTN_TC (default on)
int attention_fwd_tc(const LhrsAttention* d, cudaStream_t stream);          // attention_tc.cu
int attention_bwd_tc(const AttnBwdArgs& a, bool causal, cudaStream_t stream);   // attention_bwd_tc.cu (after the delta kernel)
bool attention_bwd_tcp_ok(const AttnBwdArgs& a);                                // attention_bwd_tcp.cu: persistent, tile-pipelined
int attention_bwd_tcp(const AttnBwdArgs& a, bool causal, cudaStream_t stream);  //   form of the same two passes (default)

#ifdef __CUDACC__
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#endif

}  // namespace lhrs
