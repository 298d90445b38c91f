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
