// Device-side next-token selection for the decode loop: the logits processors the reference's callers enable on
// HF generate — repetition penalty, temperature, top-k, top-p (cli_qa.py:176-186: do_sample, temperature; lhrs_webui.py:206-218:
// top_p 0.95, repetition_penalty 1.05) — followed by the draw, the EOS / stop-sequence test (token-suffix part of
// KeywordsStoppingCriteria, lhrs/utils/eval_utils.py:24-56) and the commit of the token into the device-side decode state.
// One block of 1024 threads; every reduction is over INTEGERS (fixed-point probability masses, counts), so the result does not
// depend on the order in which threads arrive and can be restated bit for bit on the host (oracle/sampling.py):
//   z_i    = pen(l_i) / T                       fp32, IEEE division, penalty before temperature (HF processor order)
//   top-k  : z_i = -inf where z_i < k-th largest z       (radix select over order-preserving 32-bit keys, integer counts)
//   mass_i = trunc(expf(z_i - max z) * 2^40)    uint64
//   top-p  : ascending in z, drop tokens while the cumulative mass stays <= trunc((1 - top_p) * sum mass)   (ties kept together;
//            the largest z is always kept) — HF's `cumulative_probs <= 1 - top_p` rule on exact integer masses
//   draw   : u = Philox4x32-10(seed, counter = draw index) -> 64 bits; target = (u * kept_mass) >> 64; the token is the first one
//            in token-id order whose inclusive cumulative kept mass exceeds target  (inverse CDF == multinomial over softmax)
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace lhrs {

constexpr int SMP_THREADS = 1024;
constexpr int SMP_WARPS = SMP_THREADS / 32;
constexpr size_t SMP_SMEM_BYTES = (size_t)SMP_WARPS * 256 * 2 * sizeof(unsigned);   // per-warp histograms, two 32-bit limbs per bin
constexpr int SMP_MAX_VOCAB = 65536;   // <= 2048 tokens per warp: the 20-bit limbs of <= 2^40 masses cannot overflow 32 bits

struct SampleParams {
    int do_sample;
    float temperature;
    int top_k;
    float top_p;
    float penalty;
    int eos;
    unsigned long long seed;
    const unsigned long long* seed_dev;
    const int* stop_seqs;
    int n_stop, stop_len;
};

__device__ __forceinline__ unsigned smp_key(float z) {   // order-preserving map float -> uint32 (-inf lowest)
    const unsigned u = __float_as_uint(z);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long smp_mass(float z, float m) {
    if (!(z > -INFINITY)) return 0ull;                    // -inf (filtered) and NaN carry no mass
    return static_cast<unsigned long long>(expf(z - m) * 1099511627776.0f);   // * 2^40 is exact; conversion truncates
}

// Philox4x32-10 (Salmon et al. 2011), counter (c0, c1, 0, 0), key (k0, k1); returns the first two output words as 64 bits
__device__ __forceinline__ unsigned long long smp_philox(unsigned long long seed, unsigned long long counter) {
    unsigned c0 = static_cast<unsigned>(counter), c1 = static_cast<unsigned>(counter >> 32), c2 = 0u, c3 = 0u;
    unsigned k0 = static_cast<unsigned>(seed), k1 = static_cast<unsigned>(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return (static_cast<unsigned long long>(c1) << 32) | c0;
}

struct SampleShared {
    unsigned long long tot[256];
    unsigned long long wsum[SMP_WARPS];
    float wmax[SMP_WARPS];
    int widx[SMP_WARPS];
    unsigned long long acc;      // weight strictly below the selected prefix
    unsigned prefix;
    int token;
};

__device__ __forceinline__ unsigned long long smp_block_sum(unsigned long long v, SampleShared& sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh.wsum[warp] = v;
    __syncthreads();
    unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < SMP_WARPS; ++i) t += sh.wsum[i];
    __syncthreads();          // wsum is reused by the caller
    return t;
}

// Ascending radix select over the keys of z[0..V): the smallest key v such that W(key <= v) > thr, with W the sum of per-token
// weights (COUNT: 1 per token; else the fixed-point mass).  Also returns W(key < v).  Requires thr < W(all).
// Histograms are per warp and kept as two 32-bit limbs per bin (low 20 bits | the rest), so the adds are native 32-bit
// shared-memory atomics (a 64-bit shared atomicAdd compiles to a compare-and-swap loop, which crawls when most keys share a digit).
template <bool COUNT>
__device__ __forceinline__ void smp_select(const float* __restrict__ z, int V, float m, unsigned long long thr,
                                           unsigned* hist, SampleShared& sh, unsigned& v_out, unsigned long long& below_out) {
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned* my = hist + warp * 512;
    if (tid == 0) { sh.acc = 0ull; sh.prefix = 0u; }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = tid; i < SMP_WARPS * 512; i += SMP_THREADS) hist[i] = 0u;
        __syncthreads();
        const unsigned prefix = sh.prefix;
        for (int i = tid; i < V; i += SMP_THREADS) {
            const float zi = z[i];
            const unsigned k = smp_key(zi);
            if (pass == 0 || (k >> (shift + 8)) == (prefix >> (shift + 8))) {
                const unsigned bin = ((k >> shift) & 255u) * 2;
                if (COUNT) {
                    atomicAdd(&my[bin], 1u);
                } else {
                    const unsigned long long w = smp_mass(zi, m);
                    if (w) { atomicAdd(&my[bin], static_cast<unsigned>(w & 0xFFFFFu)); atomicAdd(&my[bin + 1], static_cast<unsigned>(w >> 20)); }
                }
            }
        }
        __syncthreads();
        if (tid < 256) {
            unsigned long long t = 0;
#pragma unroll 8
            for (int w = 0; w < SMP_WARPS; ++w)
                t += static_cast<unsigned long long>(hist[w * 512 + tid * 2]) + (static_cast<unsigned long long>(hist[w * 512 + tid * 2 + 1]) << 20);
            sh.tot[tid] = t;
        }
        __syncthreads();
        if (tid < 32) {   // first warp: each lane owns 8 consecutive bins; find the first bin where the running weight exceeds thr
            unsigned long long loc[8], mine = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) { loc[b] = sh.tot[tid * 8 + b]; mine += loc[b]; }
            unsigned long long incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += t;
            }
            const unsigned long long base = sh.acc + incl - mine;          // weight below this lane's first bin
            const bool crosses = base + mine > thr;                          // the crossing bin is in this lane's range or an earlier one
            const unsigned ballot = __ballot_sync(0xffffffffu, crosses);
            const int owner = ballot ? __ffs(ballot) - 1 : 31;               // no crossing: the last bin (cannot happen when thr < total)
            if (tid == owner) {
                unsigned long long acc = base;
                int d = 0;
                for (; d < 7; ++d) {
                    if (acc + loc[d] > thr) break;
                    acc += loc[d];
                }
                sh.acc = acc;
                sh.prefix = prefix | (static_cast<unsigned>(tid * 8 + d) << shift);
            }
        }
        __syncthreads();
    }
    v_out = sh.prefix;
    below_out = sh.acc;
    __syncthreads();
}

// Returns the selected token in every thread.  z: scratch [V] (global).  hist: dynamic smem (SMP_SMEM_BYTES).  dbg (nullable):
// {sum mass, kept mass, selection key, target} for the parity tests.
__device__ __forceinline__ int smp_choose(const float* __restrict__ logits, int V, const SampleParams& p, float* __restrict__ z,
                                          const int* history, int n_hist, unsigned long long draw,
                                          unsigned* hist, SampleShared& sh, unsigned long long* dbg) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool t_on = p.do_sample && p.temperature > 0.f && p.temperature != 1.f;
    const bool pen_on = p.penalty > 0.f && p.penalty != 1.f && n_hist > 0;
    for (int i = tid; i < V; i += SMP_THREADS) z[i] = t_on ? logits[i] / p.temperature : logits[i];
    __syncthreads();
    if (pen_on) {
        // HF RepetitionPenaltyLogitsProcessor: gather the ORIGINAL scores of the seen ids, rescale, scatter (duplicates write the same value)
        for (int j = tid; j < n_hist; j += SMP_THREADS) {
            const int t = history[j];
            if (t >= 0 && t < V) {
                float l = logits[t];
                l = l < 0.f ? l * p.penalty : l / p.penalty;
                z[t] = t_on ? l / p.temperature : l;
            }
        }
        __syncthreads();
    }
    if (p.do_sample && p.top_k > 0 && p.top_k < V) {
        unsigned kth;
        unsigned long long below;
        smp_select<true>(z, V, 0.f, static_cast<unsigned long long>(V - p.top_k), hist, sh, kth, below);
        for (int i = tid; i < V; i += SMP_THREADS)
            if (smp_key(z[i]) < kth) z[i] = -INFINITY;
        __syncthreads();
    }
    // max (and argmax, lowest index on ties — torch.argmax)
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < V; i += SMP_THREADS) {
        const float v = z[i];
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sh.wmax[warp] = bv; sh.widx[warp] = bi; }
    __syncthreads();
    bv = sh.wmax[0]; bi = sh.widx[0];
#pragma unroll
    for (int i = 1; i < SMP_WARPS; ++i)
        if (sh.wmax[i] > bv || (sh.wmax[i] == bv && sh.widx[i] < bi)) { bv = sh.wmax[i]; bi = sh.widx[i]; }
    __syncthreads();
    if (!p.do_sample) return bi;

    const float m = bv;
    unsigned long long local = 0;
    for (int i = tid; i < V; i += SMP_THREADS) local += smp_mass(z[i], m);
    const unsigned long long Z = smp_block_sum(local, sh);
    unsigned vsel = 0u;
    unsigned long long below = 0ull;
    if (p.top_p > 0.f && p.top_p < 1.f) {
        const unsigned long long thr = static_cast<unsigned long long>(static_cast<double>(1.0f - p.top_p) * static_cast<double>(Z));
        smp_select<false>(z, V, m, thr, hist, sh, vsel, below);
    }
    const unsigned long long K = Z - below;
    const unsigned long long target = __umul64hi(smp_philox(p.seed_dev != nullptr ? p.seed_dev[0] : p.seed, draw), K);
    // inverse CDF in token-id order: thread t owns the contiguous ids [t*c, (t+1)*c)
    const int c = (V + SMP_THREADS - 1) / SMP_THREADS;
    const int i0 = min(tid * c, V), i1 = min(i0 + c, V);
    local = 0;
    for (int i = i0; i < i1; ++i) {
        const float zi = z[i];
        if (smp_key(zi) >= vsel) local += smp_mass(zi, m);
    }
    unsigned long long incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sh.wsum[warp] = incl;
    if (tid == 0) sh.token = bi;           // unreachable fallback (K > 0 always: the maximum has mass 2^40)
    __syncthreads();
    unsigned long long base = 0;
    for (int w = 0; w < warp; ++w) base += sh.wsum[w];
    const unsigned long long excl = base + incl - local;
    if (local > 0 && target >= excl && target < excl + local) {
        unsigned long long run = excl;
        for (int i = i0; i < i1; ++i) {
            const float zi = z[i];
            if (smp_key(zi) >= vsel) {
                run += smp_mass(zi, m);
                if (run > target) { sh.token = i; break; }
            }
        }
    }
    __syncthreads();
    if (dbg != nullptr && tid == 0) { dbg[0] = Z; dbg[1] = K; dbg[2] = vsel; dbg[3] = target; }
    return sh.token;
}

// EOS / stop-sequence test on the tokens generated so far (the newest one already stored at tokens[n-1])
__device__ __forceinline__ int smp_is_stop(const SampleParams& p, const int* __restrict__ tokens, int n, int tok) {
    if (p.eos >= 0 && tok == p.eos) return 1;
    for (int s = 0; s < p.n_stop; ++s) {
        const int* q = p.stop_seqs + s * p.stop_len;
        int first = 0;
        while (first < p.stop_len && q[first] < 0) ++first;
        const int len = p.stop_len - first;
        if (len == 0 || len > n) continue;
        bool eq = true;
        for (int j = 0; j < len && eq; ++j) eq = tokens[n - len + j] == q[first + j];
        if (eq) return 1;
    }
    return 0;
}

}  // namespace lhrs
