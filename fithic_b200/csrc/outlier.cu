// K5 -- outlier bookkeeping for spline pass >= 2.
//
// The reference keeps every line with p < 1/T in two SortedLists (line index, |mid1-mid2|; fithic/fithic.py:1215-1217)
// and, at the start of the next pass, walks the sorted distances forward through the new bins and decrements the
// possible-pair count of the bin that holds each one, clamping past the last bin (:528-548).  Bins are contiguous from
// 0, so "the bin that holds d" is the first bin with ub >= d, or the last bin.  The flagging itself is fused into K3
// (pvalue.cu: outlier_mark); this kernel turns the per-line multiplicities into per-bin decrements.  HBM-bound: 9 B/line.
#define FHC_PROFILE_STREAM st
#include "common.cuh"

namespace fhc {

constexpr int kOutlThreads = 256;
constexpr int kOutlMaxBins = 2048;

// SMEM: bin edges and per-bin counters staged in shared memory (nbins <= kOutlMaxBins, every -b a user is likely to give);
// otherwise both stay in global memory (outliers are few, so the atomics are too).
template <bool SMEM>
__global__ void __launch_bounds__(kOutlThreads)
outlier_bin_kernel(const int *__restrict__ mid1, const int *__restrict__ mid2, const unsigned char *__restrict__ outl,
                   long long n, const long long *__restrict__ bin_ub, int nbins, unsigned long long *dec) {
    __shared__ long long ub_s[SMEM ? kOutlMaxBins : 1];
    __shared__ unsigned int local[SMEM ? kOutlMaxBins : 1];
    const long long *ub = SMEM ? ub_s : bin_ub;
    if (SMEM) {
        for (int b = threadIdx.x; b < nbins; b += kOutlThreads) {
            ub_s[b] = bin_ub[b];
            local[b] = 0;
        }
        __syncthreads();
    }
    const long long ngroups = (n + 3) >> 2;
    for (long long g = (long long)blockIdx.x * kOutlThreads + threadIdx.x; g < ngroups;
         g += (long long)gridDim.x * kOutlThreads) {
        const long long i0 = g << 2;
        unsigned int flags;
        if (i0 + 3 < n) {
            flags = __ldg(reinterpret_cast<const unsigned int *>(outl) + g);
        } else {
            flags = 0;
            for (int k = 0; k < 4 && i0 + k < n; ++k) flags |= (unsigned int)outl[i0 + k] << (8 * k);
        }
        if (flags == 0) continue;  // the common case: one 4-byte load per 4 lines
        for (int k = 0; k < 4; ++k) {
            const unsigned int mult = (flags >> (8 * k)) & 255u;
            if (mult == 0) continue;
            const long long i = i0 + k;
            long long d = (long long)mid1[i] - (long long)mid2[i];
            d = d < 0 ? -d : d;
            int lo = 0, hi = nbins - 1;  // first bin with ub >= d, else the last bin
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (ub[mid] >= d)
                    hi = mid;
                else
                    lo = mid + 1;
            }
            if (SMEM)
                atomicAdd(&local[lo], mult);
            else
                atomicAdd(&dec[lo], (unsigned long long)mult);
        }
    }
    if (SMEM) {
        __syncthreads();
        for (int b = threadIdx.x; b < nbins; b += kOutlThreads)
            if (local[b]) atomicAdd(&dec[b], (unsigned long long)local[b]);
    }
}

}  // namespace fhc

extern "C" int fhc_outlier_bin_decrements(const int32_t *mid1, const int32_t *mid2, const uint8_t *outl, int64_t n,
                                          const int64_t *bin_ub, int32_t nbins, uint64_t *dec, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(n >= 0 && nbins > 0, FHC_E_INVALID, "fhc_outlier_bin_decrements: need n >= 0 and nbins > 0 (got %d)", nbins);
    FHC_REQUIRE(dec && bin_ub, FHC_E_INVALID, "fhc_outlier_bin_decrements: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    FHC_CUDA(cudaMemsetAsync(dec, 0, sizeof(uint64_t) * nbins, st));
    if (n == 0) return FHC_OK;
    FHC_REQUIRE(mid1 && mid2 && outl, FHC_E_INVALID, "fhc_outlier_bin_decrements: null pointer");
    FHC_REQUIRE(aligned16(outl), FHC_E_INVALID, "fhc_outlier_bin_decrements: outl must be 16-byte aligned");
    long long blocks = (((n + 3) >> 2) + kOutlThreads - 1) / kOutlThreads;
    if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
    if (nbins <= kOutlMaxBins)
        outlier_bin_kernel<true><<<(unsigned int)blocks, kOutlThreads, 0, st>>>(
            mid1, mid2, outl, n, reinterpret_cast<const long long *>(bin_ub), nbins, reinterpret_cast<unsigned long long *>(dec));
    else
        outlier_bin_kernel<false><<<(unsigned int)blocks, kOutlThreads, 0, st>>>(
            mid1, mid2, outl, n, reinterpret_cast<const long long *>(bin_ub), nbins, reinterpret_cast<unsigned long long *>(dec));
    FHC_LAUNCH_CHECK("outlier_bin_kernel");
    return FHC_OK;
}

// ---- result digest ------------------------------------------------------------------------------------------------------
// An order-independent 128-bit digest of (file line, p bits, q bits) over this call's lines: two 64-bit sums of hashes, so
// digests of disjoint shards add up (mod 2^64) to the digest of the whole file.  A run on N GPUs and a run on one GPU have
// computed the same p- and q-values for every line of the file if and only if (up to hash collisions) their digests agree:
// bench.py prints it at every N.  Lines come as runs: local lines [run_local[j], run_local[j + 1]) are the file lines
// run_global[j], run_global[j] + 1, ...
namespace fhc {
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {  // splitmix64 finaliser
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

constexpr int kDigestMaxRuns = FHC_MAX_CHR_RUNS;

__global__ void __launch_bounds__(256) digest_kernel(const double *__restrict__ p, const double *__restrict__ q, long long n,
                                                    const long long *__restrict__ run_local,
                                                    const long long *__restrict__ run_global, int nruns,
                                                    unsigned long long *out) {
    __shared__ long long rl[kDigestMaxRuns + 1], rg[kDigestMaxRuns];
    __shared__ unsigned long long acc[2];
    for (int r = threadIdx.x; r <= nruns; r += 256) rl[r] = run_local[r];
    for (int r = threadIdx.x; r < nruns; r += 256) rg[r] = run_global[r];
    if (threadIdx.x < 2) acc[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long s0 = 0, s1 = 0;
    int run = 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        while (rl[run + 1] <= i) ++run;
        const unsigned long long line = (unsigned long long)(rg[run] + (i - rl[run]));
        const unsigned long long pb = (unsigned long long)__double_as_longlong(p[i]);
        const unsigned long long qb = (unsigned long long)__double_as_longlong(q[i]);
        const unsigned long long h = mix64(line * 0x9e3779b97f4a7c15ull + mix64(pb) + 3ull * mix64(qb ^ 0x5555555555555555ull));
        s0 += h;
        s1 += mix64(h);
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&acc[0], s0);
        atomicAdd(&acc[1], s1);
    }
    __syncthreads();
    if (threadIdx.x < 2 && acc[threadIdx.x]) atomicAdd(&out[threadIdx.x], acc[threadIdx.x]);
}
}  // namespace fhc

extern "C" int fhc_digest_lines(const double *p, const double *q, int64_t n, const int64_t *run_local, const int64_t *run_global,
                                int32_t nruns, uint64_t *out, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(n >= 0 && out != nullptr, FHC_E_INVALID, "fhc_digest_lines: n < 0 or null output");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    FHC_CUDA(cudaMemsetAsync(out, 0, 2 * sizeof(uint64_t), st));
    if (n == 0) return FHC_OK;
    FHC_REQUIRE(p && q && run_local && run_global && nruns >= 1 && nruns <= kDigestMaxRuns, FHC_E_INVALID,
                "fhc_digest_lines: null pointer or nruns outside 1 ... %d", kDigestMaxRuns);
    long long blocks = (n + 255) / 256;
    if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
    digest_kernel<<<(unsigned int)blocks, 256, 0, st>>>(p, q, n, reinterpret_cast<const long long *>(run_local),
                                                        reinterpret_cast<const long long *>(run_global), nruns,
                                                        reinterpret_cast<unsigned long long *>(out));
    FHC_LAUNCH_CHECK("digest_kernel");
    return FHC_OK;
}
