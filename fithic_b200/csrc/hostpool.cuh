// A small pool of spinning worker threads for the host stages between the kernels (csrc/hoststage.cu).
#pragma once
#include <time.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace fhc {

// ---- worker pool -------------------------------------------------------------------------------------------------------
// parallel_for(njobs, fn): the calling thread and the workers take job indices from one atomic counter.  Workers spin for a
// while after their last job (and after prewarm()) before they go to sleep on a condition variable: a pass uses the pool
// three times within ~0.3 ms, and waking a sleeping thread costs 30-50 us.
class HostPool {
public:
    static HostPool &get() {
        static HostPool p;
        return p;
    }
    void ensure(int nworkers) {
        std::lock_guard<std::mutex> lk(mu_);
        while ((int)threads_.size() < nworkers && (int)threads_.size() < 63) threads_.emplace_back([this] { worker(); });
    }
    void prewarm() {
        wake_until_.store(now_ns() + kSpinNs);
        if (sleepers_.load() > 0) {
            std::lock_guard<std::mutex> lk(mu_);
            cv_.notify_all();
        }
    }
    void parallel_for(int njobs, int nworkers, const std::function<void(int)> &fn) {
        if (njobs <= 0) return;
        if (nworkers <= 0 || njobs == 1) {
            for (int i = 0; i < njobs; ++i) fn(i);
            return;
        }
        ensure(nworkers);
        std::lock_guard<std::mutex> run_lk(run_mu_);  // one batch at a time
        fn_ = &fn;
        njobs_ = njobs;
        remaining_.store(njobs);
        const unsigned long long e = (ticket_.load() >> 32) + 1;
        wake_until_.store(now_ns() + kSpinNs);
        ticket_.store(e << 32);  // publishes fn_ / njobs_ / remaining_ (sequentially consistent)
        if (sleepers_.load() > 0) {
            std::lock_guard<std::mutex> lk(mu_);
            cv_.notify_all();
        }
        drain();
        while (remaining_.load(std::memory_order_acquire) > 0) cpu_relax();
        // close the batch: no index can be claimed until the next batch has published its own fields
        ticket_.store((e << 32) | 0xffffffffull);
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_.store(true);
            cv_.notify_all();
        }
        for (auto &t : threads_) t.join();
    }

private:
    static constexpr long long kSpinNs = 1500000;  // 1.5 ms
    static long long now_ns() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return (long long)ts.tv_sec * 1000000000ll + ts.tv_nsec;
    }
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    // A job index is claimed by a compare-and-swap on (batch number << 32 | next index): a worker that is late for a
    // finished batch can neither run a job of it twice nor disturb the counter of the next one.
    bool claim(int &i) {
        unsigned long long v = ticket_.load(std::memory_order_acquire);
        for (;;) {
            const unsigned int idx = (unsigned int)(v & 0xffffffffull);
            if (idx >= (unsigned int)njobs_) return false;
            if (ticket_.compare_exchange_weak(v, v + 1, std::memory_order_acq_rel, std::memory_order_acquire)) {
                i = (int)idx;
                return true;
            }
        }
    }
    void drain() {
        int i;
        while (claim(i)) {
            (*fn_)(i);
            remaining_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    void worker() {
        for (;;) {
            // spin while the pool is "warm", then sleep
            unsigned long long v;
            for (;;) {
                if (stop_.load(std::memory_order_relaxed)) return;
                drain();
                v = ticket_.load();
                if (now_ns() > wake_until_.load()) break;
                for (int k = 0; k < 32; ++k) cpu_relax();
            }
            std::unique_lock<std::mutex> lk(mu_);
            sleepers_.fetch_add(1);
            cv_.wait(lk, [&] { return stop_.load() || ticket_.load() != v || now_ns() <= wake_until_.load(); });
            sleepers_.fetch_sub(1);
        }
    }
    std::vector<std::thread> threads_;
    std::mutex mu_, run_mu_;
    std::condition_variable cv_;
    std::atomic<bool> stop_{false};
    std::atomic<unsigned long long> ticket_{0xffffffffull};
    std::atomic<long long> wake_until_{0};
    std::atomic<int> sleepers_{0}, remaining_{0};
    volatile int njobs_ = 0;
    const std::function<void(int)> *volatile fn_ = nullptr;
};

}  // namespace fhc
