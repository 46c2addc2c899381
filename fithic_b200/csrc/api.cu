// Error reporting and bookkeeping shared by every entry point of libfithic_b200.so.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

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

// ---- per-kernel timing -------------------------------------------------------------------------------------------
// When enabled, every launch is followed by a cudaEventRecord on its stream; an entry point records one more event
// when it is entered (name == nullptr).  All launches of a run share one stream, so the time between two consecutive
// events is the device time of the kernel recorded by the second one (plus any idle gap in front of it, which is zero
// while the host runs ahead of the GPU).
bool g_profile_on = false;
struct Mark {
    cudaEvent_t ev;
    const char *name;
};
static std::vector<Mark> g_marks;
static std::vector<cudaEvent_t> g_pool;
static std::mutex g_prof_mu;

void profile_mark(const char *name, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEvent_t ev;
    if (!g_pool.empty()) {
        ev = g_pool.back();
        g_pool.pop_back();
    } else if (cudaEventCreate(&ev) != cudaSuccess) {
        return;
    }
    cudaEventRecord(ev, st);
    g_marks.push_back({ev, name});
}

}  // namespace fhc

extern "C" int fhc_abi_version(void) { return FHC_ABI_VERSION; }
extern "C" const char *fhc_last_error(void) { return fhc::g_err; }
extern "C" int64_t fhc_launch_count(void) { return fhc::g_launches.load(std::memory_order_relaxed); }

extern "C" int fhc_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(fhc::g_prof_mu);
    fhc::g_profile_on = on != 0;
    return FHC_OK;
}

// Synchronises the device, folds the recorded events into per-kernel totals and writes them as JSON
// ({"kernel": {"ms": total, "launches": n}, ...}) into buf; clears the log.  Returns the JSON length or an error.
extern "C" int fhc_profile_collect(char *buf, size_t buf_bytes) {
    using namespace fhc;
    FHC_REQUIRE(buf != nullptr && buf_bytes > 2, FHC_E_INVALID, "fhc_profile_collect: no buffer");
    FHC_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::map<std::string, std::pair<double, long long>> acc;
    for (size_t i = 1; i < g_marks.size(); ++i) {
        if (g_marks[i].name == nullptr) continue;  // entry marker: starts a new interval
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_marks[i - 1].ev, g_marks[i].ev) != cudaSuccess) continue;
        auto &a = acc[g_marks[i].name];
        a.first += ms;
        a.second += 1;
    }
    for (auto &m : g_marks) g_pool.push_back(m.ev);
    g_marks.clear();
    std::string out = "{";
    bool first = true;
    for (auto &kv : acc) {
        char tmp[256];
        snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"ms\": %.6f, \"launches\": %lld}", first ? "" : ", ", kv.first.c_str(),
                 kv.second.first, kv.second.second);
        out += tmp;
        first = false;
    }
    out += "}";
    FHC_REQUIRE(out.size() + 1 <= buf_bytes, FHC_E_WORKSPACE, "fhc_profile_collect: buffer of %zu bytes, need %zu",
                buf_bytes, out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return (int)out.size();
}

// ---- plumbing for hosts without a CUDA binding of their own ------------------------------------------------------------
// The Python host holds device buffers as torch tensors; these two calls let it move the small per-pass tables between
// pinned host memory and the device on the pass's stream without going through a tensor operation (8 us each in Python).
extern "C" int fhc_copy_async(void *dst, const void *src, size_t bytes, void *stream) {
    FHC_REQUIRE(bytes == 0 || (dst && src), FHC_E_INVALID, "fhc_copy_async: null pointer");
    if (bytes == 0) return FHC_OK;
    FHC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
    return FHC_OK;
}

extern "C" int fhc_stream_synchronize(void *stream) {
    FHC_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
    return FHC_OK;
}

// A CUDA event for hosts without a CUDA binding: the engine waits for the device->host copy of K1's histogram alone while
// the kernels it launched behind that copy (the pre-pass of K3) keep running.
extern "C" int fhc_event_create(void **event_out) {
    FHC_REQUIRE(event_out != nullptr, FHC_E_INVALID, "fhc_event_create: null pointer");
    cudaEvent_t ev;
    FHC_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    *event_out = ev;
    return FHC_OK;
}
extern "C" int fhc_event_record(void *event, void *stream) {
    FHC_CUDA(cudaEventRecord(static_cast<cudaEvent_t>(event), static_cast<cudaStream_t>(stream)));
    return FHC_OK;
}
extern "C" int fhc_event_synchronize(void *event) {
    FHC_CUDA(cudaEventSynchronize(static_cast<cudaEvent_t>(event)));
    return FHC_OK;
}
extern "C" int fhc_event_destroy(void *event) {
    if (event != nullptr) cudaEventDestroy(static_cast<cudaEvent_t>(event));
    return FHC_OK;
}
