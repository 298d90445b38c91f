// Host-side plumbing shared by every translation unit of liblhrs_b200.so: error slot, launch counter.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/lhrs_b200.h"

namespace lhrs {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms();

#define LHRS_CHECK_ARG(cond, ...)            \
    do {                                     \
        if (!(cond)) {                       \
            ::lhrs::set_error(__VA_ARGS__);  \
            return LHRS_ERR_INVALID;         \
        }                                    \
    } while (0)

#define LHRS_CUDA(expr)                                                                          \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            ::lhrs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return LHRS_ERR_CUDA;                                                                \
        }                                                                                        \
    } while (0)

// After a kernel launch: pick up launch-configuration errors without synchronising.
#define LHRS_LAUNCH_CHECK(name)                                                                  \
    do {                                                                                         \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess) {                                                                 \
            ::lhrs::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));          \
            return LHRS_ERR_CUDA;                                                                \
        }                                                                                        \
        ::lhrs::count_launch();                                                                  \
    } while (0)

}  // namespace lhrs

// ---- optional per-kernel timing (bench.py's roofline leg): CUDA events around every launch of a kernel family
namespace lhrs {
enum ProfKind { PROF_GEMM = 0, PROF_ATTN = 1, PROF_OTHER = 2, PROF_GEMM_SMALL = 3, PROF_SKINNY = 4, PROF_KINDS = 5 };
bool prof_on();
void prof_begin(ProfKind kind, double flops, double bytes, cudaStream_t stream);
void prof_end(cudaStream_t stream);
}  // namespace lhrs
