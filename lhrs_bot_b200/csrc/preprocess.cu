// CLIP image preprocessing on the device: uint8 RGB tiles -> normalised (3, 224, 224) pixel_values.
// Replaces, for the training / inference input path, what the reference does on the host with PIL + numpy per image:
// lhrs/Dataset/build_transform.py:43-45 (CLIPImageProcessor) called from cap_dataset.py:167-175 and cli_qa.py:119-126 —
//   resize (shortest edge -> 224, PIL BICUBIC on uint8) -> center crop 224 -> * 1/255 -> (x - mean) / std -> CHW.
// The resize is Pillow's ImagingResample restated exactly (src/libImaging/Resample.c): per output coordinate a window of
// `support = 2 * max(1, in/out)` source pixels, bicubic (a = -0.5) weights normalised in double and rounded to 22-bit fixed
// point (precompute_coeffs + normalize_coeffs_8bpc, evaluated here on the host in the same double arithmetic), a horizontal
// 8-bit pass followed by a vertical 8-bit pass, each accumulating in int32 from 1 << 21 and clamping (>> 22) to 0..255.
// Integer arithmetic end to end => bit-exact with PIL; the float stage is a 256-entry table per channel evaluated with numpy's
// operation order in IEEE arithmetic, so pixel_values are exact too.  Only the 224x224 crop window is ever computed.
// HBM-bound: each source byte is read once (B*H*W*3), the intermediate is B*rows*224*3 bytes, the output B*3*224*224*2.
#include <math.h>
#include <map>
#include <mutex>
#include <vector>

#include "host_common.h"
#include "ptx.cuh"

namespace lhrs {

constexpr int PRE_BITS = 32 - 8 - 2;   // Resample.c PRECISION_BITS

struct AxisCoeffs {   // one axis, restricted to the crop window
    int ksize = 0, first = 0, count = 0;   // source range [first, first + count) touched by the window's outputs
    std::vector<int> bounds;               // [n_out][2] = (xmin, n)
    std::vector<int> kk;                   // [n_out][ksize]
};

static double bicubic_filter(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

// precompute_coeffs + normalize_coeffs_8bpc for outputs [o0, o0 + n_out) of a full-image resize in_size -> out_size
static AxisCoeffs axis_coeffs(int in_size, int out_size, int o0, int n_out) {
    AxisCoeffs c;
    double scale, filterscale;
    filterscale = scale = (double)in_size / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;
    c.ksize = (int)ceil(support) * 2 + 1;
    c.bounds.assign((size_t)n_out * 2, 0);
    c.kk.assign((size_t)n_out * c.ksize, 0);
    const double ss = 1.0 / filterscale;
    int lo = in_size, hi = 0;
    std::vector<double> k((size_t)c.ksize);
    for (int i = 0; i < n_out; ++i) {
        const int xx = o0 + i;
        const double center = (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) k[x] /= ww;
            c.kk[(size_t)i * c.ksize + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << PRE_BITS)) : (int)(0.5 + k[x] * (1 << PRE_BITS));
        }
        c.bounds[(size_t)i * 2] = xmin;
        c.bounds[(size_t)i * 2 + 1] = xmax;
        if (xmin < lo) lo = xmin;
        if (xmin + xmax > hi) hi = xmin + xmax;
    }
    c.first = lo; c.count = hi - lo;
    return c;
}

struct PrePlan {        // device-resident tables for one (H, W, out) geometry; tiny, library-owned, never freed
    int H, W, out;
    int identity_h, identity_v;          // Pillow skips a pass whose size does not change
    int hk, vk;                          // ksize per axis
    int row0, nrows;                     // source rows the vertical pass reads
    int col_first;                       // crop offset when the horizontal pass is skipped
    int row_first;                       // crop offset when the vertical pass is skipped
    int *h_bounds, *h_kk, *v_bounds, *v_kk;
};

static std::mutex g_plan_mu;
static std::map<long long, PrePlan> g_plans;

static int upload(const std::vector<int>& v, int** dst) {
    LHRS_CUDA(cudaMalloc(dst, v.size() * sizeof(int) + 16));
    LHRS_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
    return LHRS_OK;
}

// transformers get_resize_output_image_size(size={"shortest_edge": s}, default_to_square=False)
static void resize_dims(int H, int W, int s, int* nh, int* nw) {
    const int shortv = W <= H ? W : H, longv = W <= H ? H : W;
    const int new_long = (int)((double)s * longv / shortv);
    if (W <= H) { *nw = s; *nh = new_long; } else { *nh = s; *nw = new_long; }
}

static int get_plan(int H, int W, int out, PrePlan* p) {
    int dev = 0;
    LHRS_CUDA(cudaGetDevice(&dev));                      // the tables live in the memory of the device that was current
    const long long key = ((long long)(dev & 0xF) << 60) | ((long long)H << 40) | ((long long)W << 16) | out;
    std::lock_guard<std::mutex> lk(g_plan_mu);
    auto it = g_plans.find(key);
    if (it != g_plans.end()) { *p = it->second; return LHRS_OK; }
    int nh, nw;
    resize_dims(H, W, out, &nh, &nw);
    const int top = (nh - out) / 2, left = (nw - out) / 2;
    PrePlan q;
    memset(&q, 0, sizeof(q));
    q.H = H; q.W = W; q.out = out;
    q.identity_h = (nw == W); q.identity_v = (nh == H);
    q.col_first = left; q.row_first = top;
    AxisCoeffs hc = axis_coeffs(W, nw, left, out), vc = axis_coeffs(H, nh, top, out);
    q.hk = hc.ksize; q.vk = vc.ksize;
    if (q.identity_v) { q.row0 = top; q.nrows = out; } else { q.row0 = vc.first; q.nrows = vc.count; }
    // vertical bounds are stored relative to row0 (the intermediate only holds the needed rows)
    for (int i = 0; i < out; ++i) vc.bounds[(size_t)i * 2] -= q.row0;
    int rc;
    if ((rc = upload(hc.bounds, &q.h_bounds)) || (rc = upload(hc.kk, &q.h_kk)) || (rc = upload(vc.bounds, &q.v_bounds)) ||
        (rc = upload(vc.kk, &q.v_kk)))
        return rc;
    g_plans[key] = q;
    *p = q;
    return LHRS_OK;
}

__device__ __forceinline__ int clip8(int v) {
    v >>= PRE_BITS;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// Horizontal pass: one block per (needed source row, image).  The row is staged in shared memory with 16-byte loads, then each
// thread produces crop-window bytes (x, c) from <= ksize taps.  tmp: [B][nrows][out][3] uint8.
__global__ void __launch_bounds__(256)
clip_resize_h_kernel(const uint8_t* __restrict__ src, int H, int W, PrePlan p, uint8_t* __restrict__ tmp) {
    extern __shared__ __align__(16) uint8_t row[];
    const int r = blockIdx.x, b = blockIdx.y;
    const uint8_t* s = src + ((static_cast<long long>(b) * H + p.row0 + r) * W) * 3;
    const int nbytes = W * 3;
    // the row start is only byte-aligned in general: copy the unaligned head, then 16-byte chunks, then the tail
    const int head = static_cast<int>((16 - (reinterpret_cast<uintptr_t>(s) & 15)) & 15);
    const int h = head < nbytes ? head : nbytes;
    // smem layout keeps the global alignment phase so that the vector copies line up: row[i + phase] = s[i]
    const int phase = static_cast<int>(reinterpret_cast<uintptr_t>(s) & 15);
    for (int i = threadIdx.x; i < h; i += blockDim.x) row[phase + i] = s[i];
    const int nvec = (nbytes - h) / 16;
    const uint4* sv = reinterpret_cast<const uint4*>(s + h);
    uint4* rv = reinterpret_cast<uint4*>(row + phase + h);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) rv[i] = __ldg(sv + i);
    for (int i = h + nvec * 16 + threadIdx.x; i < nbytes; i += blockDim.x) row[phase + i] = s[i];
    __syncthreads();
    const uint8_t* px = row + phase;
    uint8_t* d = tmp + (static_cast<long long>(b) * p.nrows + r) * p.out * 3;
    for (int o = threadIdx.x; o < p.out * 3; o += blockDim.x) {
        const int x = o / 3, c = o - x * 3;
        int v;
        if (p.identity_h) {
            v = px[(p.col_first + x) * 3 + c];
        } else {
            const int xmin = p.h_bounds[x * 2], n = p.h_bounds[x * 2 + 1];
            const int* k = p.h_kk + x * p.hk;
            int acc = 1 << (PRE_BITS - 1);
            for (int t = 0; t < n; ++t) acc += static_cast<int>(px[(xmin + t) * 3 + c]) * k[t];
            v = clip8(acc);
        }
        d[o] = static_cast<uint8_t>(v);
    }
}

// Vertical pass + crop + 1/255 + normalise + CHW: one block per (output row, image); adjacent threads read adjacent bytes of
// the intermediate rows.  The float stage is a [3][256] table ((float)(u * (1/255)) - mean[c]) / std[c] built per block.
__global__ void __launch_bounds__(256)
clip_resize_v_kernel(const uint8_t* __restrict__ tmp, PrePlan p, float m0, float m1, float m2, float s0, float s1, float s2,
                     void* __restrict__ out, int out_f32, uint8_t* __restrict__ resized_u8) {
    __shared__ float s_lut[3 * 256];
    const int y = blockIdx.x, b = blockIdx.y;
    // numpy order of operations (transformers image_transforms.rescale / normalize): float32(u * (1/255) evaluated in float64),
    // then (x - float32(mean)) / float32(std) in float32 — IEEE operations, identical on the device
    for (int i = threadIdx.x; i < 768; i += blockDim.x) {
        const int c = i >> 8, u = i & 255;
        const float x = static_cast<float>(static_cast<double>(u) * (1.0 / 255.0));
        const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
        s_lut[i] = __fdiv_rn(__fsub_rn(x, mean), sd);
    }
    __syncthreads();
    const int n = p.identity_v ? 1 : p.v_bounds[y * 2 + 1];
    const int ymin = p.identity_v ? y : p.v_bounds[y * 2];
    const int* k = p.v_kk + y * p.vk;
    const int rowbytes = p.out * 3;
    const uint8_t* base = tmp + (static_cast<long long>(b) * p.nrows + ymin) * rowbytes;
    const long long plane = static_cast<long long>(p.out) * p.out;
    for (int o = threadIdx.x; o < rowbytes; o += blockDim.x) {
        int v;
        if (p.identity_v) {
            v = base[o];
        } else {
            int acc = 1 << (PRE_BITS - 1);
            for (int t = 0; t < n; ++t) acc += static_cast<int>(base[static_cast<long long>(t) * rowbytes + o]) * k[t];
            v = clip8(acc);
        }
        const int x = o / 3, c = o - x * 3;
        if (resized_u8 != nullptr) resized_u8[(static_cast<long long>(b) * p.out + y) * rowbytes + o] = static_cast<uint8_t>(v);
        const float f = s_lut[c * 256 + v];
        const long long idx = (static_cast<long long>(b) * 3 + c) * plane + static_cast<long long>(y) * p.out + x;
        if (out_f32) reinterpret_cast<float*>(out)[idx] = f;
        else reinterpret_cast<__nv_bfloat16*>(out)[idx] = __float2bfloat16_rn(f);
    }
}

}  // namespace lhrs

using namespace lhrs;

static size_t pre_tmp_bytes(const PrePlan& p, int B) { return ((size_t)B * p.nrows * p.out * 3 + 255) & ~(size_t)255; }

extern "C" size_t lhrs_clip_preprocess_workspace_bytes(int32_t B, int32_t H, int32_t W, int32_t out_size) {
    if (B <= 0 || H <= 0 || W <= 0 || out_size <= 0) return 0;
    // rows needed by the vertical pass never exceed H
    return (((size_t)B * H * out_size * 3 + 255) & ~(size_t)255) + 256;
}

extern "C" int lhrs_clip_preprocess(const uint8_t* images, int32_t B, int32_t H, int32_t W, int32_t out_size, const float* mean3,
                                    const float* std3, void* out, int32_t out_f32, uint8_t* resized_u8, void* workspace,
                                    size_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    LHRS_CHECK_ARG(images && out && mean3 && std3 && workspace, "lhrs_clip_preprocess: null argument");
    LHRS_CHECK_ARG(B > 0 && H > 0 && W > 0 && out_size > 0 && out_size <= 1024 && H < (1 << 20) && W < (1 << 20),
                   "lhrs_clip_preprocess: bad geometry B=%d H=%d W=%d out=%d", B, H, W, out_size);
    LHRS_CHECK_ARG((size_t)W * 3 + 32 <= 200 * 1024, "lhrs_clip_preprocess: rows wider than %d pixels are not supported", (200 * 1024 - 32) / 3);
    PrePlan p;
    int rc = get_plan(H, W, out_size, &p);
    if (rc) return rc;
    const size_t tmp_bytes = pre_tmp_bytes(p, B);
    LHRS_CHECK_ARG(workspace_bytes >= tmp_bytes, "lhrs_clip_preprocess: workspace too small (%zu < %zu)", workspace_bytes, tmp_bytes);
    uint8_t* tmp = reinterpret_cast<uint8_t*>(workspace);
    const size_t smem = (size_t)W * 3 + 32;
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        LHRS_CUDA(cudaFuncSetAttribute(clip_resize_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        smem_set = 200 * 1024;
    }
    clip_resize_h_kernel<<<dim3(p.nrows, B), 256, smem, st>>>(images, H, W, p, tmp);
    LHRS_LAUNCH_CHECK("clip_resize_h_kernel");
    clip_resize_v_kernel<<<dim3(out_size, B), 256, 0, st>>>(tmp, p, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], out, out_f32,
                                                            resized_u8);
    LHRS_LAUNCH_CHECK("clip_resize_v_kernel");
    return LHRS_OK;
}
