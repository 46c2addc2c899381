// Value buckets behind the tightened BH cut (bh.cu), shared with K3, which can fill the histogram while it still has
// every p-value in registers (pvalue_lists.cu).
#pragma once
#include <string.h>

#include "common.cuh"

namespace fhc {

constexpr int kCutBuckets = FHC_BH_CUT_BUCKETS;  // 32768 = top 16 bits of a non-negative double below 1.0
static_assert(kCutBuckets == 32768, "bucket = high word >> 15");

__host__ __device__ inline int cut_bucket(double x) {  // x is not NaN
    if (!(x > 0.0)) return 0;
#if defined(__CUDA_ARCH__)
    const unsigned int hi = (unsigned int)__double2hiint(x);
#else
    unsigned long long b;
    memcpy(&b, &x, sizeof(b));
    const unsigned int hi = (unsigned int)(b >> 32);
#endif
    const unsigned int j = hi >> 15;
    return j < (unsigned int)kCutBuckets ? (int)j : kCutBuckets - 1;  // x >= 1: last bucket (its edge is below x)
}
__host__ __device__ inline double cut_edge(int j) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(j << 15, 0);
#else
    const unsigned long long b = (unsigned long long)(unsigned int)(j << 15) << 32;
    double x;
    memcpy(&x, &b, sizeof(x));
    return x;
#endif
}


// does K4 rank this p-value at the first (rank-bound) cut?  (p == 1.0 -> q = 1.0, NaN -> NaN, p >= cut -> q = 1.0)
__host__ __device__ inline bool cut_counts(double x, double p_cut0) { return !(x == 1.0) && !(x != x) && !(x >= p_cut0); }

}  // namespace fhc
