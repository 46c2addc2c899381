// Single-node collectives over NVLink peer memory, written for the two small exchanges of a spline pass:
//   exchange 1   sum over the GPUs of [distance histogram | totals | rank slots]      (<= 400 kB per rank)
//   exchange 2   every rank's value histogram of its p-values, gathered on every rank (256 kB per rank)
// Both are latency bound: the payload crosses NVSwitch in a few microseconds, what counts is the number of launches,
// synchronisations and software layers around it.  NCCL needs ~30-50 us per call at this size (plus the host side of the
// binding in front of it); here a collective is two small kernels and no host synchronisation:
//   push     every rank stores its payload into ITS slot of every peer's window (remote stores are posted: they do not wait
//            for a round trip like remote loads), fences, and the last CTA to finish raises this rank's flag on every peer
//   reduce   waits (bounded) until the flags of all ranks show the current epoch, then sums / copies the slots of the LOCAL
//            window -- no remote access on the critical path of the consumer
// Windows are double buffered by epoch parity.  A rank can only be one collective ahead of a peer: it starts collective e + 2
// after its own reduce of e + 1, which waited for the peer's flag e + 1, which the peer raises after its reduce of e (stream
// order on the peer) -- so the buffer of parity e is free again when a push for e + 2 arrives.
// The windows are plain cudaMalloc memory shared through CUDA IPC handles; the host (torch.distributed, MPI, a file ...)
// only has to carry the handles from rank to rank once.
#define FHC_PROFILE_STREAM st
#include <string.h>

#include <vector>

#include "common.cuh"

namespace fhc {

constexpr int kCommMaxRanks = 16;
constexpr int kCommFlagBytes = 4096;          // 2 parities x kCommMaxRanks x 8 B, padded
constexpr long long kCommSpinLimit = 1ll << 27;  // polls of a flag before a wait is given up (about a minute, not for ever)

struct CommPeers {
    unsigned char *win[kCommMaxRanks];  // base of every rank's window (own window included)
};

struct Comm {
    int rank = 0, world = 1, device = 0;
    long long slot_bytes = 0;  // payload capacity per rank and parity
    unsigned char *window = nullptr;
    unsigned int *ticket = nullptr;  // [0] CTAs of the push kernel that have finished, [1] error flag (a wait timed out)
    CommPeers peers;
    std::vector<void *> opened;
    unsigned long long epoch = 0;
    bool connected = false;
};

__device__ __forceinline__ unsigned long long *flag_of(unsigned char *win, int parity, int r) {
    return reinterpret_cast<unsigned long long *>(win) + parity * kCommMaxRanks + r;
}
__device__ __forceinline__ unsigned char *slot_of(unsigned char *win, long long slot_bytes, int world, int parity, int r) {
    return win + kCommFlagBytes + ((long long)parity * world + r) * slot_bytes;
}

// push: payload -> slot `rank` of every peer's window (16-byte stores), then the flags
__global__ void __launch_bounds__(256) comm_push_kernel(const uint4 *__restrict__ src, long long n16, CommPeers P, int rank,
                                                       int world, long long slot_bytes, int parity, unsigned long long epoch,
                                                       unsigned int *ticket) {
    for (int p = 0; p < world; ++p) {
        uint4 *dst = reinterpret_cast<uint4 *>(slot_of(P.win[p], slot_bytes, world, parity, rank));
        for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n16; i += (long long)gridDim.x * 256) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(ticket, 1u) + 1u;
        if (done == gridDim.x) {  // the payload of every CTA is on its way and fenced: tell everybody
            *ticket = 0;
            __threadfence_system();
            for (int p = 0; p < world; ++p) {
                volatile unsigned long long *f = flag_of(P.win[p], parity, rank);
                *f = epoch;
            }
        }
    }
}

__device__ __forceinline__ void comm_wait(unsigned char *win, int world, int parity, unsigned long long epoch,
                                          unsigned int *err) {
    if (threadIdx.x < world) {
        volatile unsigned long long *f = flag_of(win, parity, threadIdx.x);
        long long spins = 0;
        while (*f != epoch) {
            if (++spins > kCommSpinLimit) {
                atomicExch(err, 1u);
                break;
            }
        }
    }
    __syncthreads();
    __threadfence_system();
}

// reduce: data[i] = sum over ranks of slot r [i]  (uint64, wrap-around), reading the local window around L1; the loads of
// eight ranks are in flight together.  (Also writing the sums into pinned host memory from here, to save the caller's
// device-to-host copy, was measured at 8 GPUs: the 8-byte stores over PCIe cost the kernel more than the copy engine takes.)
__global__ void __launch_bounds__(256) comm_sum_u64_kernel(unsigned char *win, int world, long long slot_bytes, int parity,
                                                          unsigned long long epoch, unsigned long long *__restrict__ data,
                                                          long long n, unsigned int *err) {
    comm_wait(win, world, parity, epoch, err);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        unsigned long long s = 0;
        for (int r0 = 0; r0 < world; r0 += 8) {
            unsigned long long v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int r = r0 + k < world ? r0 + k : r0;  // (a duplicate load instead of a branch; not added)
                v[k] = __ldcg(reinterpret_cast<const unsigned long long *>(slot_of(win, slot_bytes, world, parity, r)) + i);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) s += r0 + k < world ? v[k] : 0ull;
        }
        data[i] = s;
    }
}

// gather: dst[r * n16 + i] = slot r [i]
__global__ void __launch_bounds__(256) comm_gather_kernel(unsigned char *win, int world, long long slot_bytes, int parity,
                                                         unsigned long long epoch, uint4 *__restrict__ dst, long long n16,
                                                         unsigned int *err) {
    comm_wait(win, world, parity, epoch, err);
    for (int r = 0; r < world; ++r) {
        const uint4 *s = reinterpret_cast<const uint4 *>(slot_of(win, slot_bytes, world, parity, r));
        for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n16; i += (long long)gridDim.x * 256)
            dst[(long long)r * n16 + i] = __ldcg(s + i);
    }
}

struct CommHandle {  // what a rank hands to its peers
    cudaIpcMemHandle_t mem;
    int rank, device;
    long long slot_bytes;
};

}  // namespace fhc

struct fhc_comm {
    fhc::Comm c;
};

extern "C" int64_t fhc_comm_handle_bytes(void) { return (int64_t)sizeof(fhc::CommHandle); }

extern "C" int fhc_comm_create(int32_t rank, int32_t world, int64_t slot_bytes, fhc_comm **comm_out, void *handle_out) {
    using namespace fhc;
    FHC_REQUIRE(comm_out && handle_out, FHC_E_INVALID, "fhc_comm_create: null pointer");
    FHC_REQUIRE(world >= 1 && world <= kCommMaxRanks && rank >= 0 && rank < world, FHC_E_INVALID,
                "fhc_comm_create: need 1 <= world <= %d and 0 <= rank < world", kCommMaxRanks);
    FHC_REQUIRE(slot_bytes >= 16, FHC_E_INVALID, "fhc_comm_create: slot_bytes must be at least 16");
    fhc_comm *h = new fhc_comm();
    Comm &c = h->c;
    c.rank = rank;
    c.world = world;
    c.slot_bytes = (slot_bytes + 255) & ~255ll;
    FHC_CUDA(cudaGetDevice(&c.device));
    const size_t bytes = (size_t)kCommFlagBytes + 2ull * (size_t)world * (size_t)c.slot_bytes;
    FHC_CUDA(cudaMalloc(&c.window, bytes));
    FHC_CUDA(cudaMemset(c.window, 0, bytes));
    FHC_CUDA(cudaMalloc(&c.ticket, 256));
    FHC_CUDA(cudaMemset(c.ticket, 0, 256));
    FHC_CUDA(cudaDeviceSynchronize());
    CommHandle hd;
    memset(&hd, 0, sizeof(hd));
    FHC_CUDA(cudaIpcGetMemHandle(&hd.mem, c.window));
    hd.rank = rank;
    hd.device = c.device;
    hd.slot_bytes = c.slot_bytes;
    memcpy(handle_out, &hd, sizeof(hd));
    for (int r = 0; r < kCommMaxRanks; ++r) c.peers.win[r] = nullptr;
    c.peers.win[rank] = c.window;
    *comm_out = h;
    return FHC_OK;
}

extern "C" int fhc_comm_connect(fhc_comm *comm, const void *all_handles) {
    using namespace fhc;
    FHC_REQUIRE(comm && all_handles, FHC_E_INVALID, "fhc_comm_connect: null pointer");
    Comm &c = comm->c;
    const CommHandle *hs = reinterpret_cast<const CommHandle *>(all_handles);
    for (int r = 0; r < c.world; ++r) {
        FHC_REQUIRE(hs[r].rank == r && hs[r].slot_bytes == c.slot_bytes, FHC_E_INVALID,
                    "fhc_comm_connect: handle %d does not belong to rank %d of this communicator", r, r);
        if (r == c.rank) continue;
        void *p = nullptr;
        FHC_CUDA(cudaIpcOpenMemHandle(&p, hs[r].mem, cudaIpcMemLazyEnablePeerAccess));
        c.opened.push_back(p);
        c.peers.win[r] = reinterpret_cast<unsigned char *>(p);
    }
    c.connected = true;
    return FHC_OK;
}

static int comm_push(fhc::Comm &c, const void *src, int64_t bytes, cudaStream_t st, int *parity_out) {
    using namespace fhc;
    FHC_REQUIRE(c.connected || c.world == 1, FHC_E_INVALID, "fhc_comm: not connected");
    FHC_REQUIRE(bytes > 0 && bytes <= c.slot_bytes && (bytes & 15) == 0 && aligned16(src), FHC_E_INVALID,
                "fhc_comm: payload of %lld bytes (a multiple of 16, 16-byte aligned, at most %lld)", (long long)bytes,
                (long long)c.slot_bytes);
    c.epoch += 1;
    const int parity = (int)(c.epoch & 1ull);
    const long long n16 = bytes / 16;
    int blocks = (int)((n16 + 255) / 256);
    if (blocks > 32) blocks = 32;
    comm_push_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4 *>(src), n16, c.peers, c.rank, c.world, c.slot_bytes,
                                             parity, c.epoch, c.ticket);
    FHC_LAUNCH_CHECK("comm_push_kernel");
    *parity_out = parity;
    return FHC_OK;
}

// data[i] (uint64, n of them, n even) <- sum over all ranks, in place; asynchronous on `stream`
extern "C" int fhc_comm_allreduce_u64(fhc_comm *comm, uint64_t *data, int64_t n, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(comm && data && n > 0 && (n & 1) == 0, FHC_E_INVALID, "fhc_comm_allreduce_u64: need an even number of words");
    Comm &c = comm->c;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    int parity = 0;
    const int rc = comm_push(c, data, n * 8, st, &parity);
    if (rc != FHC_OK) return rc;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
    comm_sum_u64_kernel<<<blocks, 256, 0, st>>>(c.window, c.world, c.slot_bytes, parity, c.epoch,
                                                reinterpret_cast<unsigned long long *>(data), n, c.ticket + 1);
    FHC_LAUNCH_CHECK("comm_sum_u64_kernel");
    return FHC_OK;
}

// dst [world x bytes] <- the payload (bytes, a multiple of 16) of every rank in rank order; asynchronous on `stream`
extern "C" int fhc_comm_allgather(fhc_comm *comm, const void *src, void *dst, int64_t bytes, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(comm && src && dst && aligned16(dst), FHC_E_INVALID, "fhc_comm_allgather: null or misaligned pointer");
    Comm &c = comm->c;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    int parity = 0;
    const int rc = comm_push(c, src, bytes, st, &parity);
    if (rc != FHC_OK) return rc;
    const long long n16 = bytes / 16;
    int blocks = (int)((n16 + 255) / 256);
    if (blocks > 64) blocks = 64;
    comm_gather_kernel<<<blocks, 256, 0, st>>>(c.window, c.world, c.slot_bytes, parity, c.epoch, reinterpret_cast<uint4 *>(dst),
                                               n16, c.ticket + 1);
    FHC_LAUNCH_CHECK("comm_gather_kernel");
    return FHC_OK;
}

extern "C" int32_t fhc_comm_world(fhc_comm *comm) { return comm ? comm->c.world : 0; }
extern "C" int32_t fhc_comm_rank(fhc_comm *comm) { return comm ? comm->c.rank : -1; }

// 1 when a wait inside a collective gave up (a peer never arrived): results since then are not to be trusted
extern "C" int fhc_comm_failed(fhc_comm *comm) {
    using namespace fhc;
    FHC_REQUIRE(comm != nullptr, FHC_E_INVALID, "fhc_comm_failed: null communicator");
    unsigned int e = 0;
    FHC_CUDA(cudaMemcpy(&e, comm->c.ticket + 1, sizeof(e), cudaMemcpyDeviceToHost));
    return e ? 1 : 0;
}

extern "C" int fhc_comm_destroy(fhc_comm *comm) {
    using namespace fhc;
    if (comm == nullptr) return FHC_OK;
    Comm &c = comm->c;
    cudaDeviceSynchronize();
    for (void *p : c.opened) cudaIpcCloseMemHandle(p);
    if (c.window) cudaFree(c.window);
    if (c.ticket) cudaFree(c.ticket);
    delete comm;
    return FHC_OK;
}
