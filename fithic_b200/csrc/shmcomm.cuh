// A few hundred bytes exchanged between the ranks of one node through POSIX shared memory: the host stage of a multi-GPU
// pass shares its possible-pair sums among the ranks (every rank would otherwise repeat the same 0.6 M additions) and needs
// the result a few microseconds later, on the HOST -- a device collective would have to queue behind the kernels that run
// during the host stage.  Slots are double buffered by epoch parity (the argument of csrc/comm.cu holds here too).
#pragma once
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <atomic>

namespace fhc {

constexpr int kShmMaxRanks = 64;

struct ShmHeader {
    std::atomic<unsigned long long> flag[kShmMaxRanks * 8];  // flag[r * 8]: one cache line per rank
};

struct ShmComm {
    int rank = 0, world = 1, fd = -1;
    long long slot_bytes = 0;
    size_t bytes = 0;
    unsigned char *base = nullptr;
    unsigned long long epoch = 0;
    char name[128];
    bool creator = false;

    unsigned char *slot(int parity, int r) const {
        return base + sizeof(ShmHeader) + ((size_t)parity * world + r) * (size_t)slot_bytes;
    }
};

static inline double shm_now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

// rank 0 creates the object, the others wait for it to appear (and to have its full size)
static inline int shm_open_comm(ShmComm &c, const char *name, int rank, int world, long long slot_bytes, double timeout_s) {
    if (world < 1 || world > kShmMaxRanks || rank < 0 || rank >= world || slot_bytes <= 0 || strlen(name) >= sizeof(c.name)) return -1;
    c.rank = rank;
    c.world = world;
    c.slot_bytes = (slot_bytes + 63) & ~63ll;
    c.bytes = sizeof(ShmHeader) + 2ull * (size_t)world * (size_t)c.slot_bytes;
    strcpy(c.name, name);
    const double t0 = shm_now_s();
    if (rank == 0) {
        shm_unlink(name);
        c.fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        if (c.fd < 0 || ftruncate(c.fd, (off_t)c.bytes) != 0) return -2;
        c.creator = true;
    } else {
        for (;;) {
            c.fd = shm_open(name, O_RDWR, 0600);
            if (c.fd >= 0) {
                struct stat st;
                if (fstat(c.fd, &st) == 0 && (size_t)st.st_size >= c.bytes) break;
                close(c.fd);
                c.fd = -1;
            }
            if (shm_now_s() - t0 > timeout_s) return -3;
            usleep(1000);
        }
    }
    void *p = mmap(nullptr, c.bytes, PROT_READ | PROT_WRITE, MAP_SHARED, c.fd, 0);
    if (p == MAP_FAILED) return -4;
    c.base = reinterpret_cast<unsigned char *>(p);
    return 0;
}

static inline void shm_close_comm(ShmComm &c) {
    if (c.base) munmap(c.base, c.bytes);
    if (c.fd >= 0) close(c.fd);
    if (c.creator) shm_unlink(c.name);
    c.base = nullptr;
    c.fd = -1;
}

// data[i] (n uint64 words) <- sum over the ranks (wrap-around); returns 0, or -1 when a rank did not arrive in time
static inline int shm_allreduce_u64(ShmComm &c, unsigned long long *data, int n, double timeout_s) {
    if ((long long)n * 8 > c.slot_bytes) return -2;
    ShmHeader *h = reinterpret_cast<ShmHeader *>(c.base);
    c.epoch += 1;
    const int parity = (int)(c.epoch & 1ull);
    memcpy(c.slot(parity, c.rank), data, (size_t)n * 8);
    h->flag[c.rank * 8].store(c.epoch, std::memory_order_release);
    const double t0 = shm_now_s();
    for (int r = 0; r < c.world; ++r) {
        long long spins = 0;
        while (h->flag[r * 8].load(std::memory_order_acquire) < c.epoch) {
            if ((++spins & 0xffff) == 0 && shm_now_s() - t0 > timeout_s) return -1;
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
        }
    }
    for (int i = 0; i < n; ++i) data[i] = 0;
    for (int r = 0; r < c.world; ++r) {
        const unsigned long long *s = reinterpret_cast<const unsigned long long *>(c.slot(parity, r));
        for (int i = 0; i < n; ++i) data[i] += s[i];
    }
    return 0;
}

}  // namespace fhc
