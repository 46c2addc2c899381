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

__global__ void __launch_bounds__(kOutlThreads)
outlier_bin_kernel(const int *__restrict__ mid1, const int *__restrict__ mid2, const unsigned char *__restrict__ outl,
                   long long n, const long long *__restrict__ bin_ub, int nbins, unsigned long long *dec) {
    __shared__ long long ub[kOutlMaxBins];
    __shared__ unsigned int local[kOutlMaxBins];
    for (int b = threadIdx.x; b < nbins; b += kOutlThreads) {
        ub[b] = bin_ub[b];
        local[b] = 0;
    }
    __syncthreads();
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
            atomicAdd(&local[lo], mult);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += kOutlThreads)
        if (local[b]) atomicAdd(&dec[b], (unsigned long long)local[b]);
}

}  // namespace fhc

extern "C" int fhc_outlier_bin_decrements(const int32_t *mid1, const int32_t *mid2, const uint8_t *outl, int64_t n,
                                          const int64_t *bin_ub, int32_t nbins, uint64_t *dec, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(n >= 0 && nbins > 0 && nbins <= kOutlMaxBins, FHC_E_INVALID,
                "fhc_outlier_bin_decrements: need n >= 0 and 0 < nbins <= %d (got %d)", kOutlMaxBins, nbins);
    FHC_REQUIRE(dec && bin_ub, FHC_E_INVALID, "fhc_outlier_bin_decrements: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    FHC_CUDA(cudaMemsetAsync(dec, 0, sizeof(uint64_t) * nbins, st));
    if (n == 0) return FHC_OK;
    FHC_REQUIRE(mid1 && mid2 && outl, FHC_E_INVALID, "fhc_outlier_bin_decrements: null pointer");
    FHC_REQUIRE(aligned16(outl), FHC_E_INVALID, "fhc_outlier_bin_decrements: outl must be 16-byte aligned");
    long long blocks = (((n + 3) >> 2) + kOutlThreads - 1) / kOutlThreads;
    if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
    outlier_bin_kernel<<<(unsigned int)blocks, kOutlThreads, 0, st>>>(mid1, mid2, outl, n,
                                                                     reinterpret_cast<const long long *>(bin_ub), nbins,
                                                                     reinterpret_cast<unsigned long long *>(dec));
    FHC_LAUNCH_CHECK("outlier_bin_kernel");
    return FHC_OK;
}
