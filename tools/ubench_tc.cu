// Micro-benchmarks that size the attention kernels' per-step budget on sm_100a (run on the GPU box; numbers go to profiles/):
//   mma      back-to-back tcgen05.mma M=128 x N x K=16 (bf16, both operands in SW128 smem), cycles per instruction for
//            N = 64 / 128 / 256, B K-major and MN-major   -> is an N=64 MMA smem-read bound?
//   ldtm     tcgen05.ld.32x32b.x32 by 4 / 8 warps of one CTA, cycles per load and bytes per cycle per SM
//   ex2      MUFU.EX2 throughput with 8 warps
//   step     the dQ-kernel compute body (2 x LDTM.x32, 32 x (ffma, ex2, select, sub, mul), pack, 4 x STS.128) for 8 warps
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I lhrs_bot_b200/csrc tools/ubench_tc.cu -o tools/ubench_tc
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

using namespace lhrs;

__device__ __forceinline__ float ex2a(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (UMMA_LAYOUT_SW128 << 29);

// ---------------------------------------------------------------- mma: one thread issues REPS x 8 MMAs, then commits
template <int N, bool B_MN>
__global__ void __launch_bounds__(128, 1) mma_kernel(long long* out, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 1) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_ptr;
    if (warp == 0) {
        const bool issuer = elect_one();
        constexpr uint32_t idesc = make_idesc_bf16(128, N, 0u, B_MN ? 1u : 0u);
        const uint32_t a_lo = (smem_u32(smem) >> 4) & 0x3FFFu;                 // A: 128 rows x 128 B (two k-blocks available)
        const uint32_t b_lo = ((smem_u32(smem + 32768) >> 4) & 0x3FFFu) | (B_MN ? ((8192u >> 4) << 16) : 0u);
        long long t0 = 0, t1 = 0, ti = 0, tc = 0;
        for (int pass = 0; pass < 2; ++pass) {                                 // pass 0 warms up
            __syncwarp();
            t0 = clock64();
            if (issuer) {
                for (int r = 0; r < reps; ++r) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint32_t ao = (kk >> 2) * (16384 >> 4) + (kk & 3) * 2;
                        const uint32_t bo = B_MN ? (kk & 3) * (2048 >> 4) : (kk >> 2) * ((N * 128) >> 4) + (kk & 3) * 2;
                        umma_bf16_w(tb + (r & 1) * 256, a_lo + ao, b_lo + bo, DESC_HI, idesc, kk ? 1u : 0u);
                    }
                }
                ti = clock64();
                umma_commit(&bar);
                tc = clock64();
            }
            __syncwarp();
            mbar_wait(&bar, pass & 1);
            t1 = clock64();
        }
        if (issuer && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = ti - t0; out[2] = tc - ti; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

// ---------------------------------------------------------------- ldtm: NW warps read TMEM
__global__ void __launch_bounds__(256, 1) ldtm_kernel(long long* out, int reps, int nwarps, uint32_t* sink) {
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_ptr + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    long long t0 = clock64();
    if (warp < nwarps) {
        for (int r = 0; r < reps; ++r) {
            uint32_t a[32], b[32];
            tmem_ld_32x32(tb + ((r * 64) & 255) + (warp >> 2) * 32, a);
            tmem_ld_32x32(tb + 256 + ((r * 64) & 255) + (warp >> 2) * 32, b);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) acc ^= a[i] + b[i];
        }
    }
    long long t1 = clock64();
    if (acc == 0x12345u) sink[threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_ptr, 512); }
}

// ---------------------------------------------------------------- ex2
__global__ void __launch_bounds__(256, 1) ex2_kernel(long long* out, int reps, float* sink) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = -0.001f * (threadIdx.x + i);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = ex2a(v[i]) - 1.0f;
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
    if (s == 123.f) sink[threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

// ---------------------------------------------------------------- step: the dQ compute body, 8 warps, no MMA
__global__ void __launch_bounds__(256, 1) step_kernel(long long* out, int reps, uint32_t* sink, int with_ld, int with_sts) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int quarter = warp & 3, half = warp >> 2, row = quarter * 32 + lane;
    const uint32_t tb = tmem_ptr + (static_cast<uint32_t>(quarter * 32) << 16);
    const float c = 0.127f, lse2 = 3.f, dl = 0.01f;
    uint32_t acc = 0;
    uint32_t rs[32], rd[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { rs[i] = __float_as_uint(0.01f * i); rd[i] = __float_as_uint(0.02f * i); }
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        const int st = r & 1;
        if (with_ld) {
            tmem_ld_32x32(tb + st * 64 + half * 32, rs);
            tmem_ld_32x32(tb + 128 + st * 64 + half * 32, rd);
            tmem_ld_wait();
        }
        uint32_t wm = 0xffffffffu >> (r & 7);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float p0 = ex2a(fmaf(__uint_as_float(rs[2 * i]), c, -lse2));
            float p1 = ex2a(fmaf(__uint_as_float(rs[2 * i + 1]), c, -lse2));
            if (!((wm >> (2 * i)) & 1u)) p0 = 0.f;
            if (!((wm >> (2 * i + 1)) & 1u)) p1 = 0.f;
            pk[i] = pack_bf16(p0 * (__uint_as_float(rd[2 * i]) - dl), p1 * (__uint_as_float(rd[2 * i + 1]) - dl));
        }
        if (with_sts) {
            uint8_t* rp = smem + st * 16384 + row * 128;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const int ch = half * 4 + c4;
                *reinterpret_cast<uint4*>(rp + ((ch ^ (row & 7)) << 4)) = make_uint4(pk[c4 * 4], pk[c4 * 4 + 1], pk[c4 * 4 + 2], pk[c4 * 4 + 3]);
            }
            fence_proxy_async_smem();
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc ^= pk[i];
        }
        if (!with_ld) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { rs[i] += pk[i & 15] & 1u; }
        }
    }
    long long t1 = clock64();
    if (acc == 0x12345u) sink[threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_ptr, 512); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

// ---------------------------------------------------------------- cost of an mbarrier wait that is already satisfied
__global__ void __launch_bounds__(128, 1) wait_kernel(long long* out, int reps) {
    __shared__ uint64_t bar[2];
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) mbar_arrive(&bar[0]);       // phase 0 of bar[0] completes
    __syncthreads();
    if (threadIdx.x < 32) {
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) mbar_wait(&bar[0], 0);       // satisfied
        long long t1 = clock64();
        for (int r = 0; r < reps; ++r) mbar_wait(&bar[1], 1);       // fresh barrier, "previous phase" parity: satisfied
        long long t2 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; }
    }
}

template <int N, bool B_MN>
static int run_mma(long long* d_out, const char* name, int reps = 256) {
    const int smem = 32768 + 65536 + 1024;
    CK(cudaFuncSetAttribute(mma_kernel<N, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    mma_kernel<N, B_MN><<<148, 128, smem>>>(d_out, reps);
    CK(cudaDeviceSynchronize());
    long long c3[3];
    CK(cudaMemcpy(c3, d_out, 24, cudaMemcpyDeviceToHost));
    const long long cyc = c3[0];
    const double per = (double)cyc / (reps * 8);
    printf("mma M128 N%-3d K16 %s x%4d: %8.1f cycles / instruction (%.0f MAC/clk/SM; floor %d) | issue loop returned after %lld cycles "
           "(%.1f / instr), commit %lld, total %lld\n", N, name, reps * 8, per, 128.0 * N * 16 / per, 128 * N / 256, c3[1],
           (double)c3[1] / (reps * 8), c3[2], cyc);
    return 0;
}

int main() {
    long long* d_out;
    uint32_t* sink;
    CK(cudaMalloc(&d_out, 64));
    CK(cudaMalloc(&sink, 4096));
    {
        wait_kernel<<<1, 128>>>(d_out, 1000);
        CK(cudaDeviceSynchronize());
        long long c2[2];
        CK(cudaMemcpy(c2, d_out, 16, cudaMemcpyDeviceToHost));
        printf("satisfied mbarrier wait (try_wait.parity + branch): %.1f cycles (completed phase), %.1f cycles (fresh barrier, parity 1)\n",
               c2[0] / 1000.0, c2[1] / 1000.0);
    }
    for (int reps : {1, 2, 4, 8}) if (run_mma<64, false>(d_out, "B K-major ", reps)) return 1;
    for (int reps : {1, 2, 4}) if (run_mma<128, false>(d_out, "B K-major ", reps)) return 1;
    if (run_mma<64, false>(d_out, "B K-major ")) return 1;
    if (run_mma<128, false>(d_out, "B K-major ")) return 1;
    if (run_mma<256, false>(d_out, "B K-major ")) return 1;
    if (run_mma<64, true>(d_out, "B MN-major")) return 1;
    if (run_mma<128, true>(d_out, "B MN-major")) return 1;
    long long cyc;
    for (int nw : {1, 4, 8}) {
        const int reps = 2048;
        ldtm_kernel<<<148, 256>>>(d_out, reps, nw, sink);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost));
        printf("ldtm 32x32b.x32, %d warp(s): %7.1f cycles per pair of loads per warp, %6.1f B/clk/SM\n", nw, (double)cyc / reps,
               (double)nw * reps * 2 * 4096 / cyc);
    }
    {
        const int reps = 1024;
        ex2_kernel<<<148, 256>>>(d_out, reps, reinterpret_cast<float*>(sink));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost));
        printf("ex2.approx, 8 warps: %6.2f ops/clk/SM (+1 FADD each)\n", 256.0 * 32 * reps / cyc);
    }
    CK(cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 1024));
    for (int mode = 0; mode < 4; ++mode) {
        const int reps = 1024, with_ld = mode & 1, with_sts = (mode >> 1) & 1;
        step_kernel<<<148, 256, 32768 + 1024>>>(d_out, reps, sink, with_ld, with_sts);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost));
        printf("dQ compute body (8 warps, 128x64 step) ldtm=%d sts=%d: %7.1f cycles per step\n", with_ld, with_sts, (double)cyc / reps);
    }
    return 0;
}
