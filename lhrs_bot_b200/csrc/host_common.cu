#include "host_common.h"

#include <stdarg.h>
#include <string.h>

namespace lhrs {

static thread_local char t_error[1024] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

}  // namespace lhrs

extern "C" const char* lhrs_last_error(void) { return lhrs::t_error; }
extern "C" int lhrs_version(void) { return 100; }
extern "C" uint64_t lhrs_launch_count(void) { return lhrs::g_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------ per-kernel timing with CUDA events
#include <mutex>
#include <vector>
namespace lhrs {
struct ProfRec { cudaEvent_t a, b; double flops, bytes; int kind; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_event_pool;
static bool g_prof_on = false;

bool prof_on() { return g_prof_on; }
static cudaEvent_t take_event() {
    if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void prof_begin(ProfKind kind, double flops, double bytes, cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r{take_event(), take_event(), flops, bytes, (int)kind};
    cudaEventRecord(r.a, stream);
    g_prof.push_back(r);
}
void prof_end(cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof.empty()) cudaEventRecord(g_prof.back().b, stream);
}
}  // namespace lhrs

extern "C" int lhrs_prof_enable(int on) {
    std::lock_guard<std::mutex> lk(lhrs::g_prof_mu);
    for (auto& r : lhrs::g_prof) { lhrs::g_event_pool.push_back(r.a); lhrs::g_event_pool.push_back(r.b); }
    lhrs::g_prof.clear();
    lhrs::g_prof_on = on != 0;
    return LHRS_OK;
}

extern "C" int lhrs_prof_summary(int kind, double* ms, double* flops, double* bytes, int64_t* launches) {
    LHRS_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(lhrs::g_prof_mu);
    double t = 0, f = 0, b = 0;
    int64_t n = 0;
    for (auto& r : lhrs::g_prof) {
        if (r.kind != kind) continue;
        float e = 0.f;
        if (cudaEventElapsedTime(&e, r.a, r.b) == cudaSuccess) { t += e; f += r.flops; b += r.bytes; ++n; }
    }
    if (ms) *ms = t;
    if (flops) *flops = f;
    if (bytes) *bytes = b;
    if (launches) *launches = n;
    return LHRS_OK;
}
