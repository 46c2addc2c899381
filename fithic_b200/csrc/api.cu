// Error reporting and bookkeeping shared by every entry point of libfithic_b200.so.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace fhc {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace fhc

extern "C" int fhc_abi_version(void) { return FHC_ABI_VERSION; }
extern "C" const char *fhc_last_error(void) { return fhc::g_err; }
extern "C" int64_t fhc_launch_count(void) { return fhc::g_launches.load(std::memory_order_relaxed); }
