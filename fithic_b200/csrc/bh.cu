// K4 -- q-values: compaction, 64-bit LSD radix sort with index payload, forward prefix-MAX scan, scatter.
//
// Replaces myStats.benjamini_hochberg_correction (reference fithic/myStats.py:24-48).  The reference walks the p-values
// in ascending order and keeps a FORWARD running MAX of min(1, p*T/rank) (not the textbook backward-min BH);
// p == 1.0 short-circuits to 1.0 and NaN sorts last and stays NaN, so only p < 1 needs ranking.
//
//   bh_compact_kernel   p -> (order-preserving uint64 key, line index) for p < 1; q = 1 / NaN written directly
//   radix sort          8 passes of 8 bits: per-tile digit histogram, exclusive scan, stable rank + scatter
//                       (warp-level match_any ranking, tile reordered in shared memory so global writes are runs)
//   bh_tilemax/bh_scan  bh_j = min((p_j*T)/rank_j, 1) with IEEE mul/div in the reference's order (bit exact), running
//                       max as a two-level scan (tile maxima, then warp-shuffle scans inside the tile), q[line] = value
//
// Every kernel reads the number of sorted elements from device memory, so the whole K4 is enqueued without a host sync.
// HBM-bound: 16 algorithmic bytes per contact (read p, write q); the LSD passes move (8+12+12) B per element per pass.
#define FHC_PROFILE_STREAM st
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace fhc {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortIPT = 16;
constexpr int kSortTile = kSortThreads * kSortIPT;  // 4096 keys per CTA
constexpr int kRadix = 256;
constexpr int kScanThreads = 1024;
constexpr int kScanTile = kScanThreads * 4;
constexpr size_t kDownsweepSmem = kSortTile * (sizeof(unsigned long long) + sizeof(unsigned int)) +
                                  (kSortWarps * kRadix + 2 * kRadix + 36) * sizeof(unsigned int);

__device__ __forceinline__ u64 key_of(double p) {  // order preserving for every non-NaN double
    const u64 b = (u64)__double_as_longlong(p);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
// Number of sort tiles that hold data.  The launch grids are sized for the capacity n_max (known on the host), but how
// many keys survive the compaction is only known on the device: every sort / scan kernel derives the live tile count from
// *d_n, and CTAs beyond it leave at once, so the cost of a pass follows the number of RANKED keys, not the lines.
__device__ __forceinline__ u32 live_tiles(const u64 *d_n) {
    const u64 n = *d_n;
    const u64 t = (n + (u64)kSortTile - 1) / (u64)kSortTile;
    return t ? (u32)t : 1u;
}

__device__ __forceinline__ double p_of(u64 k) {
    const u64 b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// ---------------------------------------------------------------------------------------------------------------------
// compaction
// ---------------------------------------------------------------------------------------------------------------------
// p_cut: every p >= p_cut is known to end with q = 1.0 exactly, so it is not ranked (see bh_p_cut below).
// One CTA iteration handles 2048 p-values (8 per thread as two 4-value groups); the (key, line) pairs of the ranked ones
// are appended with ONE global atomic per CTA iteration (same-address atomics serialise in L2: a per-warp atomic costs
// 5 ms at 300 M lines).  The order of the appended pairs is irrelevant, the sort follows.
constexpr int kCompactPerThread = 8;
__global__ void __launch_bounds__(256) bh_compact_kernel(const double *__restrict__ p, long long n,
                                                        const double *__restrict__ d_p_cut, double *__restrict__ q,
                                                        u64 *__restrict__ keys, u32 *__restrict__ vals, u64 *nsel,
                                                        u32 *__restrict__ ghist, int q_prefilled) {
    __shared__ u32 warp_tot[8];
    __shared__ u64 block_base;
    __shared__ u32 dhist[8 * 256];  // digit histograms of the ranked keys for the one-sweep sort (ghist != nullptr)
    if (ghist != nullptr) {
        for (int i = threadIdx.x; i < 8 * 256; i += 256) dhist[i] = 0;
        __syncthreads();
    }
    const double p_cut = *d_p_cut;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long per_iter = 256ll * kCompactPerThread;
    u32 cut_total = 0;
    for (long long b0 = (long long)blockIdx.x * per_iter; b0 < n; b0 += (long long)gridDim.x * per_iter) {
        double v[kCompactPerThread];
        u32 selmask = 0;
        int mine = 0;
#pragma unroll
        for (int h = 0; h < kCompactPerThread / 4; ++h) {
            const long long i0 = b0 + (long long)(h * 256 + threadIdx.x) * 4;
            double qv[4];
            if (i0 + 3 < n) {
                const double2 a = __ldcs(reinterpret_cast<const double2 *>(p + i0));
                const double2 b = __ldcs(reinterpret_cast<const double2 *>(p + i0 + 2));
                v[h * 4 + 0] = a.x; v[h * 4 + 1] = a.y; v[h * 4 + 2] = b.x; v[h * 4 + 3] = b.y;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[h * 4 + k] = (i0 + k < n) ? p[i0 + k] : 1.0;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double x = v[h * 4 + k];
                // p == 1.0 -> 1.0 (fithic/myStats.py:32-33); NaN -> NaN (min(nan, 1) -> nan, NaNs sort last)
                const bool rankable = !(x == 1.0) && !isnan(x) && (i0 + k < n);
                const bool sel = rankable && !(x >= p_cut);
                qv[k] = isnan(x) ? x : 1.0;
                selmask |= sel ? (1u << (h * 4 + k)) : 0u;
                mine += sel ? 1 : 0;
                cut_total += (rankable && !sel) ? 1u : 0u;
            }
            if (q_prefilled) {
                // the caller filled q with 1.0 while the GPU had nothing else to do: only the NaN lines are left to write
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (i0 + k < n && isnan(v[h * 4 + k])) q[i0 + k] = qv[k];
            } else if (i0 + 3 < n) {
                // ranked lines get their q from bh_scatter_kernel later; writing 1.0 first is harmless
                __stcs(reinterpret_cast<double2 *>(q + i0), make_double2(qv[0], qv[1]));
                __stcs(reinterpret_cast<double2 *>(q + i0 + 2), make_double2(qv[2], qv[3]));
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (i0 + k < n) q[i0 + k] = qv[k];
            }
        }
        // nothing of these 2048 values is ranked (most tiles of a sparse map once the cut is tightened): no scan, no atomic
        if (!__syncthreads_or(mine)) continue;
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = (u32)inc;
        __syncthreads();
        if (threadIdx.x == 0) {
            u32 tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const u32 c = warp_tot[w];
                warp_tot[w] = tot;
                tot += c;
            }
            block_base = tot ? atomicAdd(nsel, (u64)tot) : 0;
        }
        __syncthreads();
        u64 dst = block_base + warp_tot[warp] + (u64)(inc - mine);
#pragma unroll
        for (int j = 0; j < kCompactPerThread; ++j) {
            if (selmask & (1u << j)) {
                const u64 k = key_of(v[j]);
                keys[dst] = k;
                vals[dst] = (u32)(b0 + (long long)((j >> 2) * 256 + threadIdx.x) * 4 + (j & 3));
                ++dst;
                if (ghist != nullptr) {
#pragma unroll
                    for (int d = 0; d < 8; ++d) atomicAdd(&dhist[d * 256 + ((u32)(k >> (8 * d)) & 255u)], 1u);
                }
            }
        }
        __syncthreads();
    }
    const u32 cut = (u32)warp_sum((unsigned long long)cut_total);
    if (lane == 0 && cut) atomicAdd(nsel + 1, (u64)cut);  // rankable but known to end at q = 1.0
    if (ghist != nullptr) {
        __syncthreads();
        for (int i = threadIdx.x; i < 8 * 256; i += 256)
            if (dhist[i]) atomicAdd(&ghist[i], dhist[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// tightening the cut: a histogram of the p-values below the rank-bound cut tells where q reaches 1.0 for good
// ---------------------------------------------------------------------------------------------------------------------
// bh_p_cut (below) only uses rank <= lines.  The ranks themselves say more: with the p-values below the first cut counted
// in value buckets (sign/exponent/5 mantissa bits of the double: bucket edges are exact doubles), every p-value of bucket
// j has rank <= rank_offset + c_j (c_j = p-values in buckets <= j) and p >= edge_j.  If rn(edge_j T) >= rank_offset + c_j
// for a non-empty bucket j, then rn(rn(p T) / rank) >= 1 for all of bucket j (rounding is monotone), the reference caps
// those at 1 (fithic/myStats.py:36-37) and its running max (:43) stays 1.0 for every larger p: nothing from edge_j on
// needs a rank.  The smallest such edge replaces the cut; results are bit-identical, and on sparse maps with few
// significant contacts almost nothing is left to sort.
constexpr int kCutBuckets = FHC_BH_CUT_BUCKETS;  // 32768 = top 16 bits of a non-negative double below 1.0
static_assert(kCutBuckets == 32768, "bucket = high word >> 15");
constexpr int kCutHistThreads = 1024;

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

// One CTA per SM with the whole histogram in shared memory (128 KB of uint32); flushed with one global atomic per
// non-empty bucket.  Only p-values that bh_compact_kernel would rank at the first cut are counted.
// q_nan (nullable): q is pre-filled with 1.0; the NaN lines get their q = NaN here, so that a pass with nothing to rank needs no
// second sweep over p.
__global__ void __launch_bounds__(kCutHistThreads) bh_cut_hist_kernel(const double *__restrict__ p, long long n, double p_cut0,
                                                                     u64 *__restrict__ hist, double *__restrict__ q_nan) {
    extern __shared__ __align__(16) unsigned char cut_smem[];
    u32 *sh = reinterpret_cast<u32 *>(cut_smem);
    for (int i = threadIdx.x; i < kCutBuckets; i += kCutHistThreads) sh[i] = 0;
    __syncthreads();
    const long long n2 = n >> 1;
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
    for (long long i = (long long)blockIdx.x * kCutHistThreads + threadIdx.x; i < n2;
         i += (long long)gridDim.x * kCutHistThreads) {
        const double2 v = p2[i];
        if (!(v.x == 1.0) && !isnan(v.x) && !(v.x >= p_cut0)) atomicAdd(sh + cut_bucket(v.x), 1u);
        if (!(v.y == 1.0) && !isnan(v.y) && !(v.y >= p_cut0)) atomicAdd(sh + cut_bucket(v.y), 1u);
        if (q_nan != nullptr) {
            if (isnan(v.x)) q_nan[2 * i] = v.x;
            if (isnan(v.y)) q_nan[2 * i + 1] = v.y;
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const double x = p[n - 1];
        if (!(x == 1.0) && !isnan(x) && !(x >= p_cut0)) atomicAdd(sh + cut_bucket(x), 1u);
        if (q_nan != nullptr && isnan(x)) q_nan[n - 1] = x;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kCutBuckets; i += kCutHistThreads) {
        const u32 c = sh[i];
        if (c) atomicAdd(hist + i, (u64)c);
    }
}

// the smallest edge from which on every q is 1.0, or p_cut0 (same code on host and device)
__host__ __device__ inline bool cut_bucket_closes(int j, u64 c_incl, double T, double rank_offset) {
#if defined(__CUDA_ARCH__)
    const double lhs = __dmul_rn(cut_edge(j), T);
#else
    const double lhs = cut_edge(j) * T;
#endif
    return lhs >= rank_offset + (double)c_incl;  // both sides exact integers below 2^53 or a correctly rounded product
}

// inclusive scan of one u64 per thread over a 1024-thread CTA; returns the inclusive value, *total = sum over the CTA
__device__ __forceinline__ u64 block_inclusive_scan_u64(u64 v, u64 *wsum /*32*/, u64 *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u64 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();  // wsum may still be read from the previous call
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    u64 pre = 0, tot = 0;
#pragma unroll 8
    for (int w = 0; w < 32; ++w) {
        const u64 x = wsum[w];
        if (w < warp) pre += x;
        tot += x;
    }
    *total = tot;
    return inc + pre;
}

// The cut from `nranks` value histograms laid out back to back (one on a single GPU): the smallest bucket edge from which
// on every q is 1.0 (cut_bucket_closes on the summed histogram), or none.  One CTA of 1024 threads with the summed histogram
// in 128 KB of shared memory: coalesced loads (thread = bucket, 32 x nranks independent loads per thread), then every thread
// walks its own 32 CONSECUTIVE buckets behind one block-wide scan of the thread totals.  sh[j + j / 32]: the padding keeps the
// walk free of bank conflicts.  A bucket total that does not fit 32 bits switches the tightening off (always valid).
// Returns the closing bucket (kCutBuckets: none) to every thread.
constexpr int kCutPer = kCutBuckets / 1024;                               // 32 buckets per thread
constexpr size_t kCutFindSmem = (size_t)(kCutBuckets + 1024) * sizeof(u32);

__device__ __forceinline__ int cut_find_block(const u64 *__restrict__ hists, int nranks, double T, double rank_offset,
                                              u32 *sh, u64 *wsum, int *best) {
    __shared__ int overflow;
    if (threadIdx.x == 0) {
        *best = kCutBuckets;
        overflow = 0;
    }
    __syncthreads();
    for (int k = 0; k < kCutPer; ++k) {
        const int j = k * 1024 + threadIdx.x;
        u64 c = 0;
        for (int r = 0; r < nranks; ++r) c += hists[(size_t)r * kCutBuckets + j];
        if (c > 0xffffffffull) overflow = 1;
        sh[j + (j >> 5)] = (u32)c;
    }
    __syncthreads();
    if (overflow) return kCutBuckets;
    u64 mine = 0;
    const int j0 = threadIdx.x * kCutPer;
    for (int k = 0; k < kCutPer; ++k) mine += sh[j0 + k + threadIdx.x];  // (j0 + k) + (j0 + k) / 32 = j0 + k + threadIdx.x
    u64 total;
    u64 pre = block_inclusive_scan_u64(mine, wsum, &total) - mine;
    int found = kCutBuckets;
    for (int k = 0; k < kCutPer; ++k) {
        const u32 c = sh[j0 + k + threadIdx.x];
        pre += c;
        if (c && found == kCutBuckets && cut_bucket_closes(j0 + k, pre, T, rank_offset)) found = j0 + k;
    }
    if (found < kCutBuckets) atomicMin(best, found);
    __syncthreads();
    return *best;
}

// single GPU: tighten == 0 just forwards p_cut0
__global__ void __launch_bounds__(1024) bh_cut_find_kernel(const u64 *__restrict__ hist, double T, double rank_offset,
                                                          double p_cut0, int tighten, double *__restrict__ p_cut_out) {
    extern __shared__ __align__(16) unsigned char cut_find_smem[];
    __shared__ u64 wsum[32];
    __shared__ int best;
    if (!tighten || !(T > 0.0)) {
        if (threadIdx.x == 0) *p_cut_out = p_cut0;
        return;
    }
    const int b = cut_find_block(hist, 1, T, rank_offset, reinterpret_cast<u32 *>(cut_find_smem), wsum, &best);
    if (threadIdx.x == 0) *p_cut_out = b < kCutBuckets ? fmin(p_cut0, cut_edge(b)) : p_cut0;
}

// Multi-GPU: every rank's value histogram (all-gathered, nranks x kCutBuckets) -> the global cut, and how many p-values
// each rank holds below it.  info (8 + nranks words): [0] the cut (double), [1] p-values below it on all ranks, [2] on this
// rank, [3] the largest share of one rank, [8 + r] the share of rank r.
__global__ void __launch_bounds__(1024) bh_cut_from_hists_kernel(const u64 *__restrict__ hists, int nranks, int my_rank,
                                                                double T, double p_cut0, u64 *__restrict__ info) {
    extern __shared__ __align__(16) unsigned char cut_find_smem[];
    __shared__ u64 wsum[32];
    __shared__ int best;
    __shared__ u64 share[64];
    if (threadIdx.x < 64) share[threadIdx.x] = 0;
    int upto = kCutBuckets;
    if (T > 0.0) upto = cut_find_block(hists, nranks, T, 0.0, reinterpret_cast<u32 *>(cut_find_smem), wsum, &best);
    __syncthreads();
    // shares: the buckets below the closing one (all of them when none closes: the histograms only hold p < p_cut0);
    // coalesced loads, one shared-memory atomic per warp and rank
    for (int r = 0; r < nranks; ++r) {
        u64 c = 0;
        for (int k = 0; k < kCutPer; ++k) {
            const int j = k * 1024 + threadIdx.x;
            if (j < upto) c += hists[(size_t)r * kCutBuckets + j];
        }
        c = warp_sum(c);
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(&share[r], c);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double cut = p_cut0;
        if (upto < kCutBuckets) cut = fmin(cut, cut_edge(upto));
        info[0] = (u64)__double_as_longlong(cut);
        u64 tot = 0, mx = 0;
        for (int r = 0; r < nranks; ++r) {
            tot += share[r];
            mx = share[r] > mx ? share[r] : mx;
            info[8 + r] = share[r];
        }
        info[1] = tot;
        info[2] = share[my_rank];
        info[3] = mx;
    }
}

// Multi-GPU, the lean form: the value histograms of all ranks already summed (all-reduce) + this rank's own histogram ->
// the global cut, the number of p-values below it on all ranks and on this one.  info: [0] the cut (double), [1] below it
// on all ranks, [2] on this rank.
__global__ void __launch_bounds__(1024) bh_cut_from_sum_kernel(const u64 *__restrict__ summed, const u64 *__restrict__ local,
                                                              double T, double p_cut0, u64 *__restrict__ info) {
    extern __shared__ __align__(16) unsigned char cut_find_smem[];
    __shared__ u64 wsum[32];
    __shared__ int best;
    __shared__ u64 share[2];
    if (threadIdx.x < 2) share[threadIdx.x] = 0;
    int upto = kCutBuckets;
    if (T > 0.0) upto = cut_find_block(summed, 1, T, 0.0, reinterpret_cast<u32 *>(cut_find_smem), wsum, &best);
    __syncthreads();
    u64 ca = 0, cm = 0;
    for (int k = 0; k < kCutPer; ++k) {
        const int j = k * 1024 + threadIdx.x;
        if (j < upto) {
            ca += summed[j];
            cm += local[j];
        }
    }
    ca = warp_sum(ca);
    cm = warp_sum(cm);
    if ((threadIdx.x & 31) == 0) {
        if (ca) atomicAdd(&share[0], ca);
        if (cm) atomicAdd(&share[1], cm);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double cut = p_cut0;
        if (upto < kCutBuckets) cut = fmin(cut, cut_edge(upto));
        info[0] = (u64)__double_as_longlong(cut);
        info[1] = share[0];
        info[2] = share[1];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// exclusive scan of a uint32 array (three phases)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 block_exclusive_scan_u32(u32 v, u32 *smem_warp /*32*/, u32 &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u32 w = lane < nw ? smem_warp[lane] : 0;
        u32 winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        smem_warp[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) smem_warp[32] = winc;
    }
    __syncthreads();
    const u32 r = smem_warp[warp] + inc - v;
    total = smem_warp[32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const u32 *__restrict__ data, const u64 *d_n,
                                                                  u32 *__restrict__ blocksums) {
    __shared__ u32 sw[33];
    const long long len = (long long)kRadix * live_tiles(d_n);
    if ((long long)blockIdx.x * kScanTile >= len) return;
    const long long i0 = ((long long)blockIdx.x * kScanThreads + threadIdx.x) * 4;
    u32 s = 0;
    if (i0 + 3 < len) {
        const uint4 v = *reinterpret_cast<const uint4 *>(data + i0);
        s = v.x + v.y + v.z + v.w;
    } else {
        for (long long i = i0; i < len; ++i) s += data[i];
    }
    u32 total;
    block_exclusive_scan_u32(s, sw, total);
    if (threadIdx.x == 0) blocksums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_blocksums_kernel(u32 *blocksums, const u64 *d_n) {
    __shared__ u32 sw[33];
    __shared__ u32 carry;
    const int nb = (int)(((long long)kRadix * live_tiles(d_n) + kScanTile - 1) / kScanTile);
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const u32 v = i < nb ? blocksums[i] : 0;
        u32 total;
        const u32 ex = block_exclusive_scan_u32(v, sw, total);
        if (i < nb) blocksums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(u32 *__restrict__ data, const u64 *d_n,
                                                                 const u32 *__restrict__ blocksums) {
    __shared__ u32 sw[33];
    const long long len = (long long)kRadix * live_tiles(d_n);
    if ((long long)blockIdx.x * kScanTile >= len) return;
    const long long i0 = ((long long)blockIdx.x * kScanThreads + threadIdx.x) * 4;
    u32 a = 0, b = 0, c = 0, d = 0;
    const bool full = i0 + 3 < len;
    if (full) {
        const uint4 v = *reinterpret_cast<const uint4 *>(data + i0);
        a = v.x; b = v.y; c = v.z; d = v.w;
    } else {
        if (i0 < len) a = data[i0];
        if (i0 + 1 < len) b = data[i0 + 1];
        if (i0 + 2 < len) c = data[i0 + 2];
    }
    u32 total;
    const u32 ex = block_exclusive_scan_u32(a + b + c + d, sw, total) + blocksums[blockIdx.x];
    const uint4 o = make_uint4(ex, ex + a, ex + a + b, ex + a + b + c);
    if (full) {
        *reinterpret_cast<uint4 *>(data + i0) = o;
    } else {
        if (i0 < len) data[i0] = o.x;
        if (i0 + 1 < len) data[i0 + 1] = o.y;
        if (i0 + 2 < len) data[i0 + 2] = o.z;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// radix sort passes
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads) radix_upsweep_kernel(const u64 *__restrict__ keys, const u64 *d_n,
                                                                    int shift, u32 *__restrict__ counts) {
    __shared__ u32 hist[kSortWarps][kRadix];
    const u32 ntiles = live_tiles(d_n);
    const long long n = (long long)*d_n;
    const int warp = threadIdx.x >> 5;
    for (u32 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {  // persistent CTAs: the grid does not follow n_max
        for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&hist[0][0])[i] = 0;
        __syncthreads();
        const long long base = (long long)tile * kSortTile;
#pragma unroll
        for (int r = 0; r < kSortIPT; ++r) {
            const long long i = base + r * kSortThreads + threadIdx.x;
            if (i < n) atomicAdd(&hist[warp][(u32)(keys[i] >> shift) & 255u], 1u);
        }
        __syncthreads();
        u32 s = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) s += hist[w][threadIdx.x];
        counts[(size_t)threadIdx.x * ntiles + tile] = s;  // digit-major so one flat scan gives global offsets
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kSortThreads, 3)
radix_downsweep_kernel(const u64 *__restrict__ keys_in, const u32 *__restrict__ vals_in, u64 *__restrict__ keys_out,
                       u32 *__restrict__ vals_out, const u64 *d_n, int shift, const u32 *__restrict__ offsets) {
    extern __shared__ __align__(16) unsigned char dsmem[];
    u64 *skeys = reinterpret_cast<u64 *>(dsmem);                                  // [kSortTile]
    u32 *svals = reinterpret_cast<u32 *>(skeys + kSortTile);                      // [kSortTile]
    u32(*whist)[kRadix] = reinterpret_cast<u32(*)[kRadix]>(svals + kSortTile);    // [kSortWarps][kRadix]
    u32 *tile_off = reinterpret_cast<u32 *>(whist + kSortWarps);                  // [kRadix]
    u32 *gbase = tile_off + kRadix;                                               // [kRadix]
    u32 *sw = gbase + kRadix;                                                     // [33]
    const long long n = (long long)*d_n;
    const u32 ntiles = live_tiles(d_n);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u32 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {  // persistent CTAs
    const long long base = (long long)tile * kSortTile;
    if (base >= n) break;
    const int tile_n = (int)((n - base) < kSortTile ? (n - base) : kSortTile);
    for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&whist[0][0])[i] = 0;
    __syncthreads();

    u64 key[kSortIPT];
    unsigned short rank[kSortIPT];
    const u32 lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < kSortIPT; ++r) {  // warp `warp` owns [warp*512, warp*512+512) of the tile, 32 at a time
        const int li = warp * (32 * kSortIPT) + r * 32 + lane;
        const bool valid = li < tile_n;
        key[r] = valid ? keys_in[base + li] : ~0ull;
    }
#pragma unroll
    for (int r = 0; r < kSortIPT; ++r) {
        const int li = warp * (32 * kSortIPT) + r * 32 + lane;
        const bool valid = li < tile_n;
        const u32 d = valid ? ((u32)(key[r] >> shift) & 255u) : 256u;
        const u32 m = __match_any_sync(0xffffffffu, d);
        u32 pre = 0;
        if (valid) pre = whist[warp][d];
        __syncwarp();
        if (valid && lane == (__ffs(m) - 1)) whist[warp][d] = pre + __popc(m);
        __syncwarp();
        rank[r] = (unsigned short)(pre + __popc(m & lt));
    }
    __syncthreads();
    // per digit: exclusive prefix over warps, tile total
    u32 cnt = 0;
    {
        const int d = threadIdx.x;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const u32 c = whist[w][d];
            whist[w][d] = cnt;
            cnt += c;
        }
    }
    u32 total;
    const u32 ex = block_exclusive_scan_u32(cnt, sw, total);
    tile_off[threadIdx.x] = ex;
    gbase[threadIdx.x] = offsets[(size_t)threadIdx.x * ntiles + tile];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortIPT; ++r) {
        const int li = warp * (32 * kSortIPT) + r * 32 + lane;
        if (li < tile_n) {
            const u32 d = (u32)(key[r] >> shift) & 255u;
            const u32 pos = tile_off[d] + whist[warp][d] + rank[r];
            skeys[pos] = key[r];
            svals[pos] = vals_in[base + li];
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortIPT; ++r) {
        const int i = r * kSortThreads + threadIdx.x;
        if (i < tile_n) {
            const u64 k = skeys[i];
            const u32 d = (u32)(k >> shift) & 255u;
            const u32 dst = gbase[d] + ((u32)i - tile_off[d]);
            keys_out[dst] = k;
            vals_out[dst] = svals[i];
        }
    }
    __syncthreads();  // shared memory is reused by the next tile
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// one-sweep passes: the same stable rank + scatter per tile, but the position of a tile's digit groups in the output comes
// from a decoupled look-back over the tiles before it instead of a separate histogram pass and a three-kernel scan
// ---------------------------------------------------------------------------------------------------------------------
// Digit histograms of all eight passes in one read of the keys (ghist[pass * 256 + digit]); the BH path gets them for free
// from the compaction kernel, which holds every key in registers anyway.
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const u64 *__restrict__ keys, const u64 *d_n,
                                                                 u32 *__restrict__ ghist) {
    __shared__ u32 hist[8 * kRadix];
    for (int i = threadIdx.x; i < 8 * kRadix; i += kSortThreads) hist[i] = 0;
    __syncthreads();
    const long long n = (long long)*d_n;
    for (long long i = (long long)blockIdx.x * kSortThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kSortThreads) {
        const u64 k = keys[i];
#pragma unroll
        for (int p = 0; p < 8; ++p) atomicAdd(&hist[p * kRadix + ((u32)(k >> (8 * p)) & 255u)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * kRadix; i += kSortThreads)
        if (hist[i]) atomicAdd(&ghist[i], hist[i]);
}

// gbase[pass * 256 + digit] = keys with a smaller digit in that pass; uniform[pass] = 1 when every key has the same digit
// there (the pass would move nothing).  One CTA of 256 threads.
__global__ void __launch_bounds__(kRadix) radix_digit_scan_kernel(const u32 *__restrict__ ghist, const u64 *d_n,
                                                                 u32 *__restrict__ gbase, u32 *__restrict__ uniform) {
    __shared__ u32 sw[33];
    const u64 n = *d_n;
    for (int p = 0; p < 8; ++p) {
        const u32 c = ghist[p * kRadix + threadIdx.x];
        u32 total;
        const u32 ex = block_exclusive_scan_u32(c, sw, total);
        gbase[p * kRadix + threadIdx.x] = ex;
        if (threadIdx.x == 0) uniform[p] = 0;
        __syncthreads();
        if (c != 0 && (u64)c == n) uniform[p] = 1;
        __syncthreads();
    }
}

constexpr u32 kStatusAggregate = 1u, kStatusPrefix = 2u;  // low two bits of a status word; the count sits above them
constexpr int kLookbackSpinLimit = 1 << 22;

// status[tile * 256 + digit]: 0 = nothing yet, (count << 2) | 1 = this tile's own count, (count << 2) | 2 = count of this
// tile and all tiles before it.  Tiles are handed out by an atomic counter, so every tile before a running one has been
// claimed by a running CTA and publishes its own count without waiting for anybody: the look-back cannot dead-lock.  The
// spin is bounded all the same (err is raised instead of hanging the GPU).
// kT threads per CTA, kSortTile / kT keys per thread
template <int kT>
__global__ void __launch_bounds__(kT, kT == 512 ? 3 : 3)
radix_onesweep_kernel(const u64 *__restrict__ keys_in, const u32 *__restrict__ vals_in, u64 *__restrict__ keys_out,
                      u32 *__restrict__ vals_out, const u64 *d_n, int shift, const u32 *__restrict__ gbase_pass,
                      volatile u32 *status, u32 *tile_counter, u32 *err) {
    extern __shared__ __align__(16) unsigned char dsmem[];
    u64 *skeys = reinterpret_cast<u64 *>(dsmem);                                  // [kSortTile]
    u32 *svals = reinterpret_cast<u32 *>(skeys + kSortTile);                      // [kSortTile]
    u32(*whist)[kRadix] = reinterpret_cast<u32(*)[kRadix]>(svals + kSortTile);    // [kW][kRadix]
    constexpr int kW = kT / 32, kI = kSortTile / kT;
    u32 *tile_off = reinterpret_cast<u32 *>(whist + kW);                          // [kRadix]
    u32 *gpos = tile_off + kRadix;                                                // [kRadix]
    u32 *sw = gpos + kRadix;                                                      // [33] + the tile number
    const long long n = (long long)*d_n;
    const u32 ntiles = n > 0 ? live_tiles(d_n) : 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (;;) {
        if (threadIdx.x == 0) sw[34] = atomicAdd(tile_counter, 1u);
        __syncthreads();
        const u32 tile = sw[34];
        if (tile >= ntiles) break;
        const long long base = (long long)tile * kSortTile;
        const int tile_n = (int)((n - base) < kSortTile ? (n - base) : kSortTile);
        for (int i = threadIdx.x; i < kW * kRadix; i += kT) (&whist[0][0])[i] = 0;
        __syncthreads();

        u64 key[kI];
        unsigned short rank[kI];
        const u32 lt = (1u << lane) - 1u;
#pragma unroll
        for (int r = 0; r < kI; ++r) {  // warp `warp` owns 32 * kI consecutive keys of the tile, 32 at a time
            const int li = warp * (32 * kI) + r * 32 + lane;
            key[r] = li < tile_n ? keys_in[base + li] : ~0ull;
        }
#pragma unroll
        for (int r = 0; r < kI; ++r) {
            const int li = warp * (32 * kI) + r * 32 + lane;
            const bool valid = li < tile_n;
            const u32 d = valid ? ((u32)(key[r] >> shift) & 255u) : 256u;
            const u32 m = __match_any_sync(0xffffffffu, d);
            // the first lane of every group of equal digits advances the warp's own counter and hands the old value to its
            // group: one shared-memory atomic and one shuffle per round, and no load that waits for the previous round's
            // store -- the rounds' atomics are in flight together (same-address atomics of one warp keep program order)
            const int leader = __ffs(m) - 1;
            u32 pre = 0;
            if (valid && lane == leader) pre = atomicAdd(&whist[warp][d], (u32)__popc(m));
            pre = __shfl_sync(0xffffffffu, pre, leader);
            rank[r] = (unsigned short)(pre + __popc(m & lt));
        }
        __syncthreads();
        // per digit (thread = digit): exclusive prefix over warps, tile total
        const bool digit_thread = kT == kRadix || threadIdx.x < kRadix;
        u32 cnt = 0;
        if (digit_thread) {
            const int d = threadIdx.x;
#pragma unroll
            for (int w = 0; w < kW; ++w) {
                const u32 c = whist[w][d];
                whist[w][d] = cnt;
                cnt += c;
            }
        }
        // publish this tile's count of the digit, then add up the tiles before it
        u32 excl = 0;
        if (digit_thread) {
            volatile u32 *mine = status + (size_t)tile * kRadix + threadIdx.x;
            if (tile == 0) {
                *mine = (cnt << 2) | kStatusPrefix;
            } else {
                *mine = (cnt << 2) | kStatusAggregate;
                // The status words of the kLookAhead tiles before the current position are loaded together (their latencies
                // overlap: the walk is a chain of L2 round trips otherwise, and with ~450 tiles in flight it is ~100 steps
                // long) and consumed in order up to the first inclusive prefix; what was loaded beyond it is dropped.
                constexpr int kLookAhead = 8;
                bool done = false;
                for (long long t = (long long)tile - 1; t >= 0 && !done; t -= kLookAhead) {
                    u32 v[kLookAhead];
#pragma unroll
                    for (int k = 0; k < kLookAhead; ++k)
                        v[k] = t - k >= 0 ? (u32)status[(size_t)(t - k) * kRadix + threadIdx.x] : (u32)kStatusPrefix;  // before tile 0: nothing
#pragma unroll
                    for (int k = 0; k < kLookAhead; ++k) {
                        if (done) continue;
                        u32 x = v[k];
                        if ((x & 3u) == 0u) {  // not published yet: wait for this one
                            const volatile u32 *theirs = status + (size_t)(t - k) * kRadix + threadIdx.x;
                            int spins = 0;
                            x = *theirs;
                            while ((x & 3u) == 0u) {
                                if (++spins > kLookbackSpinLimit) {
                                    atomicExch(err, 1u);
                                    x = kStatusPrefix;
                                    break;
                                }
                                x = *theirs;
                            }
                        }
                        excl += x >> 2;
                        if ((x & 3u) == kStatusPrefix) done = true;
                    }
                }
                *mine = ((excl + cnt) << 2) | kStatusPrefix;
            }
        }
        u32 total;
        const u32 ex = block_exclusive_scan_u32(cnt, sw, total);  // (threads beyond the digits add 0 behind them)
        if (digit_thread) {
            tile_off[threadIdx.x] = ex;
            gpos[threadIdx.x] = gbase_pass[threadIdx.x] + excl;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kI; ++r) {
            const int li = warp * (32 * kI) + r * 32 + lane;
            if (li < tile_n) {
                const u32 d = (u32)(key[r] >> shift) & 255u;
                const u32 pos = tile_off[d] + whist[warp][d] + rank[r];
                skeys[pos] = key[r];
                svals[pos] = vals_in[base + li];
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kI; ++r) {
            const int i = r * kT + threadIdx.x;
            if (i < tile_n) {
                const u64 k = skeys[i];
                const u32 d = (u32)(k >> shift) & 255u;
                const u32 dst = gpos[d] + ((u32)i - tile_off[d]);
                keys_out[dst] = k;
                vals_out[dst] = svals[i];
            }
        }
        __syncthreads();  // shared memory is reused by the next tile
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// running max of min((p*T)/rank, 1) over the sorted keys and scatter
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double bh_value(u64 key, double T, long long rank) {
    // bh = pv * T / (i + 1); bh = min(bh, 1)   (fithic/myStats.py:35-37, evaluated left to right)
    const double v = __ddiv_rn(__dmul_rn(p_of(key), T), (double)rank);
    return v > 1.0 ? 1.0 : v;
}

// d_n[3] != 0: the sort left its result in the second pair of buffers (an odd number of passes ran)
__global__ void __launch_bounds__(kSortThreads) bh_tilemax_kernel(const u64 *__restrict__ keys_a, const u64 *__restrict__ keys_b,
                                                                 const u64 *d_n, double T, long long rank_offset,
                                                                 double *__restrict__ tilemax) {
    __shared__ double sm[kSortWarps];
    const u64 *__restrict__ keys = d_n[3] ? keys_b : keys_a;
    const long long n = (long long)*d_n;
    const u32 ntiles = live_tiles(d_n);
    for (u32 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = (long long)tile * kSortTile;
        double m = 0.0;
#pragma unroll 4
        for (int r = 0; r < kSortIPT; ++r) {
            const long long i = base + r * kSortThreads + threadIdx.x;
            if (i < n) m = fmax(m, bh_value(keys[i], T, rank_offset + i + 1));
        }
        m = warp_max(m);
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = sm[0];
#pragma unroll
            for (int w = 1; w < kSortWarps; ++w) t = fmax(t, sm[w]);
            tilemax[tile] = t;
        }
        __syncthreads();
    }
}

// exclusive running max over the tile maxima, seeded with carry_in (single CTA)
__global__ void __launch_bounds__(kScanThreads) bh_tilescan_kernel(double *tilemax, int ntiles, double carry_in,
                                                                  double *carry_out, const u64 *d_n, u64 *n_sorted_out) {
    __shared__ double sw[32];
    __shared__ double carry;
    if (threadIdx.x == 0) carry = carry_in;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (ntiles > 0) ntiles = min(ntiles, (int)live_tiles(d_n));  // tiles beyond hold no keys
    for (int base = 0; base < ntiles; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const double v = i < ntiles ? tilemax[i] : 0.0;
        double inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc = fmax(inc, t);
        }
        double exl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) exl = 0.0;
        if (lane == 31) sw[warp] = inc;
        __syncthreads();
        double wpre = 0.0;
        for (int w = 0; w < warp; ++w) wpre = fmax(wpre, sw[w]);
        const double c = carry;
        if (i < ntiles) tilemax[i] = fmax(c, fmax(wpre, exl));
        __syncthreads();
        if (threadIdx.x == kScanThreads - 1) carry = fmax(c, fmax(wpre, inc));
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // d_n[1] = p-values that were not ranked because they end at exactly 1.0: they follow every ranked one
        const u64 ncut = d_n[1];
        if (carry_out) *carry_out = ncut ? fmax(carry, 1.0) : carry;
        if (n_sorted_out) *n_sorted_out = *d_n + ncut;
    }
}

// Every thread takes kSortIPT CONSECUTIVE sorted keys: a running max in registers, one exclusive max-scan of the thread
// totals per tile (instead of one block scan per 256 keys), then the scattered stores of q.
__global__ void __launch_bounds__(kSortThreads)
bh_scatter_kernel(const u64 *__restrict__ keys_a, const u32 *__restrict__ vals_a, const u64 *__restrict__ keys_b,
                  const u32 *__restrict__ vals_b, const u64 *d_n, double T, long long rank_offset,
                  const double *__restrict__ tilepre, double floor_in, double *__restrict__ q) {
    __shared__ double sw[kSortWarps];
    const u64 *__restrict__ keys = d_n[3] ? keys_b : keys_a;
    const u32 *__restrict__ vals = d_n[3] ? vals_b : vals_a;
    const long long n = (long long)*d_n;
    const u32 ntiles = live_tiles(d_n);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u32 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long base = (long long)tile * kSortTile;
        if (base >= n) break;
        const long long i0 = base + (long long)threadIdx.x * kSortIPT;
        double v[kSortIPT];
        double run = 0.0;
        if (i0 + kSortIPT <= n) {
            const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(keys + i0);  // i0 is a multiple of 16: aligned
#pragma unroll
            for (int k = 0; k < kSortIPT; k += 2) {
                const ulonglong2 kk = k2[k >> 1];
                run = fmax(run, bh_value(kk.x, T, rank_offset + i0 + k + 1));
                v[k] = run;
                run = fmax(run, bh_value(kk.y, T, rank_offset + i0 + k + 2));
                v[k + 1] = run;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kSortIPT; ++k) {
                if (i0 + k < n) run = fmax(run, bh_value(keys[i0 + k], T, rank_offset + i0 + k + 1));
                v[k] = run;
            }
        }
        // exclusive running max over the threads of the tile
        double inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc = fmax(inc, t);
        }
        double excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = 0.0;
        if (lane == 31) sw[warp] = inc;
        __syncthreads();
        double pre = fmax(tilepre[tile], floor_in);
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w)
            if (w < warp) pre = fmax(pre, sw[w]);
        pre = fmax(pre, excl);
        if (i0 + kSortIPT <= n) {
            const uint4 *v4 = reinterpret_cast<const uint4 *>(vals + i0);
#pragma unroll
            for (int k = 0; k < kSortIPT; k += 4) {
                const uint4 ii = v4[k >> 2];
                q[ii.x] = fmax(v[k], pre);  // bh = max(bh, prev)   (fithic/myStats.py:43)
                q[ii.y] = fmax(v[k + 1], pre);
                q[ii.z] = fmax(v[k + 2], pre);
                q[ii.w] = fmax(v[k + 3], pre);
            }
        } else {
#pragma unroll
            for (int k = 0; k < kSortIPT; ++k)
                if (i0 + k < n) q[vals[i0 + k]] = fmax(v[k], pre);
        }
        __syncthreads();  // sw is reused by the next tile
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
struct SortWs {
    u32 *counts;     // LSD passes: per-tile digit counts; one-sweep passes: the look-back status words (same size)
    u32 *blocksums;
    u32 *onesweep;   // [0, 2048) digit histograms of the 8 passes, [2048, 4096) their exclusive scans, [4096, 4104) "every
                     // key has the same digit" per pass, [4104, 4112) tile counters per pass, [4112] look-back error flag
    size_t counts_len;
    int nb;
    u32 ntiles;
};
constexpr int kOsHist = 0, kOsBase = 2048, kOsUniform = 4096, kOsCounter = 4104, kOsErr = 4112, kOsWords = 4128;

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// persistent grids: at most `per_sm` CTAs per SM loop over the tiles that hold keys (known on the device only)
static unsigned int sort_grid(u32 ntiles, int per_sm) {
    const u32 cap = (u32)kNumSMs * (u32)per_sm;
    return ntiles < cap ? (ntiles ? ntiles : 1u) : cap;
}

static size_t sort_ws_layout(int64_t n, char *base, SortWs *ws) {
    const u32 ntiles = (u32)((n + kSortTile - 1) / kSortTile);
    const size_t counts_len = (size_t)kRadix * (ntiles ? ntiles : 1);
    const int nb = (int)((counts_len + kScanTile - 1) / kScanTile);
    size_t off = 0;
    if (ws) ws->counts = reinterpret_cast<u32 *>(base + off);
    off += align_up(counts_len * sizeof(u32) + 16);
    if (ws) ws->blocksums = reinterpret_cast<u32 *>(base + off);
    off += align_up((size_t)nb * sizeof(u32));
    if (ws) ws->onesweep = reinterpret_cast<u32 *>(base + off);
    off += align_up((size_t)kOsWords * sizeof(u32));
    if (ws) {
        ws->counts_len = counts_len;
        ws->nb = nb;
        ws->ntiles = ntiles ? ntiles : 1;
    }
    return off;
}

// sorts n_max-capacity buffers holding *d_n valid pairs; result ends in (keys_b, vals_b) after 8 passes -> we run an
// even number of passes so the result is back in (keys_a, vals_a)
static int sort_pairs_device_n(u64 *keys_a, u32 *vals_a, u64 *keys_b, u32 *vals_b, int64_t n_max, const u64 *d_n,
                               const SortWs &ws_full, cudaStream_t st, long long known_n = -1) {
    u64 *kin = keys_a, *kout = keys_b;
    u32 *vin = vals_a, *vout = vals_b;
    SortWs ws = ws_full;
    if (known_n >= 0) {  // the caller read the key count back: size the grids for it instead of for the capacity
        const u32 t = (u32)((known_n + kSortTile - 1) / kSortTile);
        ws.ntiles = t ? t : 1;
        ws.nb = (int)(((size_t)kRadix * ws.ntiles + kScanTile - 1) / kScanTile);
    }
    FHC_CUDA(cudaFuncSetAttribute(radix_downsweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kDownsweepSmem));
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = pass * 8;
        radix_upsweep_kernel<<<sort_grid(ws.ntiles, 8), kSortThreads, 0, st>>>(kin, d_n, shift, ws.counts);
        FHC_LAUNCH_CHECK("radix_upsweep_kernel");
        scan_reduce_kernel<<<ws.nb, kScanThreads, 0, st>>>(ws.counts, d_n, ws.blocksums);
        FHC_LAUNCH_CHECK("scan_reduce_kernel");
        scan_blocksums_kernel<<<1, kScanThreads, 0, st>>>(ws.blocksums, d_n);
        FHC_LAUNCH_CHECK("scan_blocksums_kernel");
        scan_apply_kernel<<<ws.nb, kScanThreads, 0, st>>>(ws.counts, d_n, ws.blocksums);
        FHC_LAUNCH_CHECK("scan_apply_kernel");
        radix_downsweep_kernel<<<sort_grid(ws.ntiles, 3), kSortThreads, kDownsweepSmem, st>>>(kin, vin, kout, vout, d_n, shift, ws.counts);
        FHC_LAUNCH_CHECK("radix_downsweep_kernel");
        u64 *tk = kin; kin = kout; kout = tk;
        u32 *tv = vin; vin = vout; vout = tv;
    }
    (void)n_max;
    return FHC_OK;
}

// One-sweep LSD sort of n_max-capacity buffers holding *d_n valid pairs (n < 2^30: a status word keeps a count in 30 bits).
// have_hist: the digit histograms are already in ws.onesweep (bh_compact_kernel); otherwise one pass over the keys builds
// them.  scanned: radix_digit_scan_kernel has run on them.  skip[pass] (host, nullable): passes the caller knows to be
// uniform are not launched.  Returns the buffer that
// holds the result in *result_in_a (1: keys_a / vals_a, 0: keys_b / vals_b).
static int sort_pairs_onesweep(u64 *keys_a, u32 *vals_a, u64 *keys_b, u32 *vals_b, const u64 *d_n, const SortWs &ws_full,
                               cudaStream_t st, long long known_n, bool have_hist, bool scanned, const unsigned char *skip,
                               int *result_in_a) {
    SortWs ws = ws_full;
    if (known_n >= 0) {
        const u32 t = (u32)((known_n + kSortTile - 1) / kSortTile);
        ws.ntiles = t ? t : 1;
    }
    u32 *os = ws.onesweep;
    if (!have_hist) {
        FHC_CUDA(cudaMemsetAsync(os + kOsHist, 0, 2048 * sizeof(u32), st));
        radix_hist_kernel<<<sort_grid(ws.ntiles, 8), kSortThreads, 0, st>>>(keys_a, d_n, os + kOsHist);
        FHC_LAUNCH_CHECK("radix_hist_kernel");
    }
    if (!scanned) {
        radix_digit_scan_kernel<<<1, kRadix, 0, st>>>(os + kOsHist, d_n, os + kOsBase, os + kOsUniform);
        FHC_LAUNCH_CHECK("radix_digit_scan_kernel");
    }
    FHC_CUDA(cudaMemsetAsync(os + kOsCounter, 0, (kOsWords - kOsCounter) * sizeof(u32), st));
    // 512 threads x 8 keys (48 warps per SM at 40 registers) measured the same 0.18 ms per pass as 256 x 16, and so did nine
    // ballots in place of MATCH.ANY: a tile is a chain of exposed latencies (claim, key load, look-back, value load) between
    // barriers, and three CTAs per SM do not cover them
    constexpr int kOsThreads = 256;
    constexpr size_t kOsSmem = kSortTile * (sizeof(unsigned long long) + sizeof(unsigned int)) +
                               ((kOsThreads / 32) * kRadix + 2 * kRadix + 36) * sizeof(unsigned int);
    FHC_CUDA(cudaFuncSetAttribute(radix_onesweep_kernel<kOsThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOsSmem));
    u64 *kin = keys_a, *kout = keys_b;
    u32 *vin = vals_a, *vout = vals_b;
    int in_a = 1;
    for (int pass = 0; pass < 8; ++pass) {
        if (skip != nullptr && skip[pass]) continue;
        FHC_CUDA(cudaMemsetAsync(ws.counts, 0, (size_t)ws.ntiles * kRadix * sizeof(u32), st));
        radix_onesweep_kernel<kOsThreads><<<sort_grid(ws.ntiles, 3), kOsThreads, kOsSmem, st>>>(
            kin, vin, kout, vout, d_n, pass * 8, os + kOsBase + pass * kRadix, ws.counts, os + kOsCounter + pass, os + kOsErr);
        FHC_LAUNCH_CHECK("radix_onesweep_kernel");
        u64 *tk = kin; kin = kout; kout = tk;
        u32 *tv = vin; vin = vout; vout = tv;
        in_a ^= 1;
    }
    *result_in_a = in_a;
    return FHC_OK;
}

static bool use_onesweep(int64_t n) {
    const char *e = getenv("FHC_SORT");  // lsd: the histogram / scan / scatter passes of round 1 (also taken for n >= 2^30)
    if (e && e[0] == 'l') return false;
    return n < (1ll << 30);
}

}  // namespace fhc

extern "C" size_t fhc_sort_workspace_bytes(int64_t n) {
    if (n < 0) n = 0;
    return fhc::sort_ws_layout(n, nullptr, nullptr) + 256 /* device copy of n */;
}

extern "C" int fhc_sort_pairs_u64(uint64_t *keys_in, uint32_t *vals_in, uint64_t *keys_out, uint32_t *vals_out,
                                  int64_t n, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(n >= 0 && n < (1ll << 32), FHC_E_INVALID, "fhc_sort_pairs_u64: need 0 <= n < 2^32 (got %lld)", (long long)n);
    if (n == 0) return FHC_OK;
    FHC_REQUIRE(keys_in && vals_in && keys_out && vals_out && workspace, FHC_E_INVALID, "fhc_sort_pairs_u64: null pointer");
    FHC_REQUIRE(workspace_bytes >= fhc_sort_workspace_bytes(n), FHC_E_WORKSPACE,
                "fhc_sort_pairs_u64: workspace of %zu bytes, need %zu", workspace_bytes, fhc_sort_workspace_bytes(n));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    char *base = reinterpret_cast<char *>(workspace);
    u64 *d_n = reinterpret_cast<u64 *>(base);
    SortWs ws;
    sort_ws_layout(n, base + 256, &ws);
    const u64 hn = (u64)n;
    FHC_CUDA(cudaMemcpyAsync(d_n, &hn, sizeof(u64), cudaMemcpyHostToDevice, st));
    int in_a = 1;
    int rc;
    if (use_onesweep(n))
        rc = sort_pairs_onesweep(reinterpret_cast<u64 *>(keys_in), vals_in, reinterpret_cast<u64 *>(keys_out), vals_out, d_n,
                                 ws, st, n, false, false, nullptr, &in_a);
    else
        rc = sort_pairs_device_n(reinterpret_cast<u64 *>(keys_in), vals_in, reinterpret_cast<u64 *>(keys_out), vals_out, n,
                                 d_n, ws, st);
    if (rc != FHC_OK) return rc;
    if (in_a) {  // an even number of passes: the result is back in keys_in / vals_in; move it where the caller asked for it
        FHC_CUDA(cudaMemcpyAsync(keys_out, keys_in, sizeof(u64) * n, cudaMemcpyDeviceToDevice, st));
        FHC_CUDA(cudaMemcpyAsync(vals_out, vals_in, sizeof(u32) * n, cudaMemcpyDeviceToDevice, st));
    }
    return FHC_OK;
}

namespace fhc {
struct BhWs {
    u64 *d_n;        // [0] ranked keys, [1] rankable but cut, then (as double) [2] the cut in force
    u64 *cut_hist;   // [kCutBuckets]
    u64 *keys_a, *keys_b;
    u32 *vals_a, *vals_b;
    double *tilemax;
    SortWs sort;
};
static size_t bh_ws_layout(int64_t n, char *base, BhWs *ws) {
    size_t off = 0;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    const size_t ntiles = (nn + kSortTile - 1) / kSortTile;
    if (ws) ws->d_n = reinterpret_cast<u64 *>(base + off);
    off += 256;
    if (ws) ws->cut_hist = reinterpret_cast<u64 *>(base + off);
    off += align_up((size_t)kCutBuckets * sizeof(u64));
    if (ws) ws->keys_a = reinterpret_cast<u64 *>(base + off);
    off += align_up(nn * sizeof(u64));
    if (ws) ws->keys_b = reinterpret_cast<u64 *>(base + off);
    off += align_up(nn * sizeof(u64));
    if (ws) ws->vals_a = reinterpret_cast<u32 *>(base + off);
    off += align_up(nn * sizeof(u32));
    if (ws) ws->vals_b = reinterpret_cast<u32 *>(base + off);
    off += align_up(nn * sizeof(u32));
    if (ws) ws->tilemax = reinterpret_cast<double *>(base + off);
    off += align_up(ntiles * sizeof(double));
    off += sort_ws_layout(n, base + off, ws ? &ws->sort : nullptr);
    return off;
}
}  // namespace fhc

extern "C" size_t fhc_bh_workspace_bytes(int64_t n) { return fhc::bh_ws_layout(n, nullptr, nullptr); }

// compaction + sort + tile maxima + tile scan: everything of K4 that does not need the running max of smaller keys held
// by other GPUs.  carry_out receives max(carry_in, every bh value of this call).
// Every p-value of this call has rank <= rank_bound, so p >= rank_bound / T gives (p*T)/rank >= 1: bh is capped at 1
// and the running max is exactly 1.0 from the first such p on (they all sort after the smaller ones, whose ranks they do
// not influence).  Those lines get q = 1.0 without being ranked.  With T (possible pairs) >> lines, as in every
// sparse high-resolution map, this removes ~98 % of the sort.  The 1e-9 margin keeps the claim safe under rounding.
static double bh_p_cut(double T, double rank_bound) {
    if (!(T > 0.0) || !(rank_bound >= 0.0)) return INFINITY;
    return (rank_bound / T) * (1.0 + 1e-9);
}

static int cut_hist_launch(const double *p, int64_t n, double p_cut0, fhc::u64 *hist, cudaStream_t st,
                           double *q_nan = nullptr) {
    using namespace fhc;
    const size_t smem = (size_t)kCutBuckets * sizeof(u32);
    FHC_CUDA(cudaFuncSetAttribute(bh_cut_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = (n / 2 + kCutHistThreads - 1) / kCutHistThreads;
    if (blocks > kNumSMs) blocks = kNumSMs;
    if (blocks < 1) blocks = 1;
    bh_cut_hist_kernel<<<(unsigned int)blocks, kCutHistThreads, smem, st>>>(p, n, p_cut0, hist, q_nan);
    FHC_LAUNCH_CHECK("bh_cut_hist_kernel");
    return FHC_OK;
}

// tighten: derive the cut from the value histogram (single-GPU path); otherwise p_cut is used as given (the multi-GPU
// path tightens globally before the exchange, fhc_bh_cut_hist + fhc_host_bh_cut_find)
// host_ns (nullable): synchronise the stream once after the compaction, read the number of ranked keys into *host_ns and
// launch only what that number needs (nothing when no key is ranked: 40 near-empty launches cost 0.35 ms).
static int bh_prepare(const double *p, int64_t n, double T, int64_t rank_offset, double carry_in, double p_cut, bool tighten,
                      double *q, double *carry_out, int64_t *n_sorted_out, const fhc::BhWs &ws, cudaStream_t st,
                      long long *host_ns = nullptr, int q_prefilled = 0) {
    using namespace fhc;
    FHC_CUDA(cudaMemsetAsync(ws.d_n, 0, 4 * sizeof(u64), st));
    int ntiles = (int)ws.sort.ntiles;
    if (host_ns) *host_ns = -1;
    if (n > 0) {
        const char *tenv = getenv("FHC_BH_TIGHTEN");  // =0: rank-bound cut only (experiments, worst-case timing)
        if (tenv && tenv[0] == '0') tighten = false;
        double *d_p_cut = reinterpret_cast<double *>(ws.d_n + 2);
        if (tighten) {
            FHC_CUDA(cudaMemsetAsync(ws.cut_hist, 0, (size_t)kCutBuckets * sizeof(u64), st));
            const int rc = cut_hist_launch(p, n, p_cut, ws.cut_hist, st);
            if (rc != FHC_OK) return rc;
        }
        FHC_CUDA(cudaFuncSetAttribute(bh_cut_find_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCutFindSmem));
        bh_cut_find_kernel<<<1, 1024, kCutFindSmem, st>>>(ws.cut_hist, T, (double)rank_offset, p_cut, tighten ? 1 : 0, d_p_cut);
        FHC_LAUNCH_CHECK("bh_cut_find_kernel");
        long long blocks = (n + 256 * kCompactPerThread - 1) / (256 * kCompactPerThread);
        if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
        const bool onesweep = use_onesweep(n);
        u32 *os = ws.sort.onesweep;
        if (onesweep) FHC_CUDA(cudaMemsetAsync(os + kOsHist, 0, 2048 * sizeof(u32), st));
        bh_compact_kernel<<<(unsigned int)blocks, 256, 0, st>>>(p, n, d_p_cut, q, ws.keys_a, ws.vals_a, ws.d_n,
                                                                onesweep ? os + kOsHist : nullptr, q_prefilled);
        FHC_LAUNCH_CHECK("bh_compact_kernel");
        if (onesweep) {  // digit offsets of all eight passes, and which passes would move nothing
            radix_digit_scan_kernel<<<1, kRadix, 0, st>>>(os + kOsHist, ws.d_n, os + kOsBase, os + kOsUniform);
            FHC_LAUNCH_CHECK("radix_digit_scan_kernel");
        }
        long long known = -1;
        unsigned char skip[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (host_ns) {
            u64 h = 0;
            u32 uni[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            FHC_CUDA(cudaMemcpyAsync(&h, ws.d_n, sizeof(u64), cudaMemcpyDeviceToHost, st));
            if (onesweep) FHC_CUDA(cudaMemcpyAsync(uni, os + kOsUniform, sizeof(uni), cudaMemcpyDeviceToHost, st));
            FHC_CUDA(cudaStreamSynchronize(st));
            known = (long long)h;
            *host_ns = known;
            ntiles = (int)((known + kSortTile - 1) / kSortTile);
            for (int k = 0; k < 8; ++k) skip[k] = uni[k] ? 1 : 0;
        }
        if (known != 0) {
            int in_a = 1;
            const int rc = onesweep ? sort_pairs_onesweep(ws.keys_a, ws.vals_a, ws.keys_b, ws.vals_b, ws.d_n, ws.sort, st, known,
                                                          true, true, skip, &in_a)
                                    : sort_pairs_device_n(ws.keys_a, ws.vals_a, ws.keys_b, ws.vals_b, n, ws.d_n, ws.sort, st, known);
            if (rc != FHC_OK) return rc;
            if (!in_a) FHC_CUDA(cudaMemsetAsync(ws.d_n + 3, 0xff, sizeof(u64), st));  // the result sits in keys_b / vals_b
            bh_tilemax_kernel<<<sort_grid((u32)ntiles, 8), kSortThreads, 0, st>>>(ws.keys_a, ws.keys_b, ws.d_n, T, rank_offset,
                                                                                  ws.tilemax);
            FHC_LAUNCH_CHECK("bh_tilemax_kernel");
        }
    }
    bh_tilescan_kernel<<<1, kScanThreads, 0, st>>>(ws.tilemax, n > 0 ? ntiles : 0, carry_in, carry_out, ws.d_n,
                                                   reinterpret_cast<u64 *>(n_sorted_out));
    FHC_LAUNCH_CHECK("bh_tilescan_kernel");
    return FHC_OK;
}

static int bh_finish(int64_t n, double T, int64_t rank_offset, double floor_in, double *q, const fhc::BhWs &ws,
                     cudaStream_t st, long long known_n = -1) {
    using namespace fhc;
    if (n > 0 && known_n != 0) {
        const u32 tiles = known_n > 0 ? (u32)((known_n + kSortTile - 1) / kSortTile) : ws.sort.ntiles;
        bh_scatter_kernel<<<sort_grid(tiles, 8), kSortThreads, 0, st>>>(ws.keys_a, ws.vals_a, ws.keys_b, ws.vals_b, ws.d_n, T,
                                                                        rank_offset, ws.tilemax, floor_in, q);
        FHC_LAUNCH_CHECK("bh_scatter_kernel");
    }
    return FHC_OK;
}

static int bh_check_args(const char *who, const double *p, int64_t n, double *q, void *workspace, size_t workspace_bytes) {
    FHC_REQUIRE(n >= 0 && n < (1ll << 32), FHC_E_INVALID, "%s: need 0 <= n < 2^32 (got %lld)", who, (long long)n);
    FHC_REQUIRE(workspace != nullptr, FHC_E_INVALID, "%s: null workspace", who);
    FHC_REQUIRE(workspace_bytes >= fhc_bh_workspace_bytes(n), FHC_E_WORKSPACE, "%s: workspace of %zu bytes, need %zu", who,
                workspace_bytes, fhc_bh_workspace_bytes(n));
    FHC_REQUIRE(n == 0 || (p && q && p != q), FHC_E_INVALID, "%s: p and q must be distinct non-null arrays", who);
    return FHC_OK;
}

extern "C" int fhc_bh_qvalues(const double *p, int64_t n, double T, int64_t rank_offset, double carry_in, double *q,
                              double *carry_out, int64_t *n_sorted_out, void *workspace, size_t workspace_bytes,
                              void *stream) {
    using namespace fhc;
    int rc = bh_check_args("fhc_bh_qvalues", p, n, q, workspace, workspace_bytes);
    if (rc != FHC_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    BhWs ws;
    bh_ws_layout(n, reinterpret_cast<char *>(workspace), &ws);
    rc = bh_prepare(p, n, T, rank_offset, carry_in, bh_p_cut(T, (double)rank_offset + (double)n), true, q, carry_out,
                    n_sorted_out, ws, st);
    if (rc != FHC_OK) return rc;
    return bh_finish(n, T, rank_offset, 0.0, q, ws, st);  // carry_in is already folded into the tile prefixes
}

extern "C" int fhc_bh_qvalues_hostcount(const double *p, int64_t n, double T, int64_t rank_offset, double carry_in, double *q,
                                        double *carry_out, int64_t *n_sorted_out, int64_t *n_ranked_host, void *workspace,
                                        size_t workspace_bytes, int32_t q_prefilled, void *stream) {
    using namespace fhc;
    int rc = bh_check_args("fhc_bh_qvalues_hostcount", p, n, q, workspace, workspace_bytes);
    if (rc != FHC_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    BhWs ws;
    bh_ws_layout(n, reinterpret_cast<char *>(workspace), &ws);
    long long ns = -1;
    rc = bh_prepare(p, n, T, rank_offset, carry_in, bh_p_cut(T, (double)rank_offset + (double)n), true, q, carry_out,
                    n_sorted_out, ws, st, &ns, q_prefilled ? 1 : 0);
    if (rc != FHC_OK) return rc;
    if (n_ranked_host) *n_ranked_host = ns < 0 ? 0 : ns;
    return bh_finish(n, T, rank_offset, 0.0, q, ws, st, ns);
}

extern "C" int fhc_bh_prepare(const double *p, int64_t n, double T, int64_t rank_offset, double p_cut, double *q,
                              double *local_max_out, int64_t *n_sorted_out, void *workspace, size_t workspace_bytes,
                              void *stream) {
    using namespace fhc;
    int rc = bh_check_args("fhc_bh_prepare", p, n, q, workspace, workspace_bytes);
    if (rc != FHC_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    BhWs ws;
    bh_ws_layout(n, reinterpret_cast<char *>(workspace), &ws);
    return bh_prepare(p, n, T, rank_offset, 0.0, p_cut, false, q, local_max_out, n_sorted_out, ws, st);
}

extern "C" int fhc_bh_finish(int64_t n, double T, int64_t rank_offset, double floor_in, double *q, void *workspace,
                             size_t workspace_bytes, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(n >= 0 && n < (1ll << 32) && workspace != nullptr && workspace_bytes >= fhc_bh_workspace_bytes(n),
                FHC_E_INVALID, "fhc_bh_finish: bad n / workspace");
    FHC_REQUIRE(n == 0 || q != nullptr, FHC_E_INVALID, "fhc_bh_finish: null q");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    BhWs ws;
    bh_ws_layout(n, reinterpret_cast<char *>(workspace), &ws);
    return bh_finish(n, T, rank_offset, floor_in, q, ws, st);
}

// ---------------------------------------------------------------------------------------------------------------------
// range partitioning of p-values over GPUs (multi-GPU BH): sample, count, scatter, gather back
// ---------------------------------------------------------------------------------------------------------------------
namespace fhc {

constexpr int kMaxParts = 64;

struct Splitters {
    u64 key[kMaxParts];  // part r holds keys in [key[r-1], key[r]); key[nparts-1] unused
    int nparts;
};

__device__ __forceinline__ int part_of(u64 k, const Splitters &sp) {
    int r = 0;
    for (int i = 0; i < sp.nparts - 1; ++i) r += (k >= sp.key[i]) ? 1 : 0;
    return r;
}

__global__ void bh_sample_kernel(const double *__restrict__ p, long long n, long long stride, long long nsamples,
                                 double p_cut, u64 *__restrict__ keys) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nsamples) return;
    const long long i = s * stride;
    u64 k = ~0ull;  // "no sample": sorts last
    if (i < n) {
        const double v = p[i];
        if (!(v >= p_cut) && !(v == 1.0) && !isnan(v)) k = key_of(v);
    }
    keys[s] = k;
}

// part of a value for the two kernels below: -1 when the value is not ranked
__device__ __forceinline__ int part_or_none(double v, double p_cut, const Splitters &sp) {
    if (v >= p_cut || v == 1.0 || isnan(v)) return -1;
    return part_of(key_of(v), sp);
}

// Both kernels read 8 values per thread (two 4-value groups, 128-bit loads) and aggregate per warp before touching a
// shared counter: lanes with the same destination are found with match_any, one of them adds the group's size.
__global__ void __launch_bounds__(256) bh_part_count_kernel(const double *__restrict__ p, long long n, const Splitters sp,
                                                           double p_cut, u64 *__restrict__ counts) {
    __shared__ u32 local[kMaxParts];
    if (threadIdx.x < kMaxParts) local[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long ngroups = (n + 3) >> 2;
    for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g < ((ngroups + 31) & ~31ll);
         g += (long long)gridDim.x * 256) {
        const long long i0 = g << 2;
        double v[4];
        if (i0 + 3 < n) {
            const double2 a = __ldcs(reinterpret_cast<const double2 *>(p + i0));
            const double2 b = __ldcs(reinterpret_cast<const double2 *>(p + i0 + 2));
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? p[i0 + k] : 1.0;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int part = part_or_none(v[k], p_cut, sp);
            const u32 m = __match_any_sync(0xffffffffu, part);
            if (part >= 0 && lane == __ffs(m) - 1) atomicAdd(&local[part], (u32)__popc(m));
        }
    }
    __syncthreads();
    if (threadIdx.x < sp.nparts && local[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (u64)local[threadIdx.x]);
}

// p -> send buffer grouped by destination part (order inside a part is arbitrary: the receiver sorts); idx[j] = source line
// of send[j]; q gets 1.0 / NaN for the p-values that are not ranked.  cursors[r] must start at the first slot of part r.
// Per CTA iteration (1024 values): count per part in shared memory, reserve the ranges with one global atomic per part,
// then write.
__global__ void __launch_bounds__(256) bh_part_scatter_kernel(const double *__restrict__ p, long long n, const Splitters sp,
                                                             double p_cut, u64 *cursors, double *__restrict__ send,
                                                             u32 *__restrict__ idx, double *__restrict__ q, int q_prefilled) {
    __shared__ u32 cnt[kMaxParts];
    __shared__ u64 base[kMaxParts];
    const int lane = threadIdx.x & 31;
    for (long long b0 = (long long)blockIdx.x * 1024; b0 < n; b0 += (long long)gridDim.x * 1024) {
        if (threadIdx.x < kMaxParts) cnt[threadIdx.x] = 0;
        __syncthreads();
        const long long i0 = b0 + (long long)threadIdx.x * 4;
        double v[4];
        if (i0 + 3 < n) {
            const double2 a = __ldcs(reinterpret_cast<const double2 *>(p + i0));
            const double2 b = __ldcs(reinterpret_cast<const double2 *>(p + i0 + 2));
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? p[i0 + k] : 1.0;
        }
        int part[4];
        u32 slot[4];
        double qv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            part[k] = (i0 + k < n) ? part_or_none(v[k], p_cut, sp) : -1;
            qv[k] = isnan(v[k]) ? v[k] : 1.0;
            const u32 m = __match_any_sync(0xffffffffu, part[k]);
            const int leader = __ffs(m) - 1;
            u32 first = 0;
            if (part[k] >= 0 && lane == leader) first = atomicAdd(&cnt[part[k]], (u32)__popc(m));
            first = __shfl_sync(0xffffffffu, first, leader);
            slot[k] = first + __popc(m & ((1u << lane) - 1u));
        }
        if (q_prefilled) {  // q holds 1.0 already: only the NaN lines are written
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i0 + k < n && isnan(v[k])) q[i0 + k] = qv[k];
        } else if (i0 + 3 < n) {
            // ranked lines get their q back from the owning GPU later; writing 1.0 first is harmless
            __stcs(reinterpret_cast<double2 *>(q + i0), make_double2(qv[0], qv[1]));
            __stcs(reinterpret_cast<double2 *>(q + i0 + 2), make_double2(qv[2], qv[3]));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (i0 + k < n) q[i0 + k] = qv[k];
        }
        __syncthreads();
        if (threadIdx.x < sp.nparts)
            base[threadIdx.x] = cnt[threadIdx.x] ? atomicAdd(&cursors[threadIdx.x], (u64)cnt[threadIdx.x]) : 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (part[k] >= 0) {
                const u64 dst = base[part[k]] + slot[k];
                send[dst] = v[k];
                idx[dst] = (u32)(i0 + k);
            }
        }
        __syncthreads();
    }
}

__global__ void scatter_f64_kernel(const double *__restrict__ src, const u32 *__restrict__ idx, long long n,
                                   double *__restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[idx[i]] = src[i];
}

static int make_splitters(const uint64_t *splitter_keys, int nparts, Splitters *sp) {
    FHC_REQUIRE(nparts >= 1 && nparts <= kMaxParts, FHC_E_INVALID, "need 1 <= nparts <= %d (got %d)", kMaxParts, nparts);
    FHC_REQUIRE(nparts == 1 || splitter_keys != nullptr, FHC_E_INVALID, "null splitters");
    sp->nparts = nparts;
    for (int i = 0; i < nparts - 1; ++i) {
        sp->key[i] = splitter_keys[i];
        FHC_REQUIRE(i == 0 || sp->key[i] >= sp->key[i - 1], FHC_E_INVALID, "splitters must be ascending");
    }
    return FHC_OK;
}

}  // namespace fhc

extern "C" int fhc_bh_sample_keys(const double *p, int64_t n, int64_t nsamples, double p_cut, uint64_t *keys_out,
                                  void *stream) {
    using namespace fhc;
    FHC_REQUIRE(n >= 0 && nsamples > 0 && keys_out != nullptr, FHC_E_INVALID, "fhc_bh_sample_keys: bad arguments");
    FHC_REQUIRE(n == 0 || p != nullptr, FHC_E_INVALID, "fhc_bh_sample_keys: null p");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    const long long stride = n > nsamples ? n / nsamples : 1;
    bh_sample_kernel<<<(unsigned int)((nsamples + 255) / 256), 256, 0, st>>>(p, n, stride, nsamples, p_cut,
                                                                            reinterpret_cast<u64 *>(keys_out));
    FHC_LAUNCH_CHECK("bh_sample_kernel");
    return FHC_OK;
}

// smallest p that is certain to end with q = 1.0 when no p-value has a rank above rank_bound (host helper)
extern "C" double fhc_bh_p_cut(double T, double rank_bound) { return bh_p_cut(T, rank_bound); }

extern "C" int fhc_bh_cut_hist(const double *p, int64_t n, double p_cut0, uint64_t *hist, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(n >= 0 && hist != nullptr, FHC_E_INVALID, "fhc_bh_cut_hist: n < 0 or null histogram");
    if (n == 0) return FHC_OK;
    FHC_REQUIRE(p != nullptr && aligned16(p), FHC_E_INVALID, "fhc_bh_cut_hist: p must be a 16-byte aligned device array");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    return cut_hist_launch(p, n, p_cut0, reinterpret_cast<u64 *>(hist), st);
}

extern "C" int fhc_bh_cut_from_hists(const uint64_t *hists, int32_t nranks, int32_t my_rank, double T, double p_cut0,
                                     uint64_t *info, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(hists && info && nranks >= 1 && nranks <= 64 && my_rank >= 0 && my_rank < nranks, FHC_E_INVALID,
                "fhc_bh_cut_from_hists: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    FHC_CUDA(cudaFuncSetAttribute(bh_cut_from_hists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCutFindSmem));
    bh_cut_from_hists_kernel<<<1, 1024, kCutFindSmem, st>>>(reinterpret_cast<const u64 *>(hists), nranks, my_rank, T, p_cut0,
                                                 reinterpret_cast<u64 *>(info));
    FHC_LAUNCH_CHECK("bh_cut_from_hists_kernel");
    return FHC_OK;
}

// The first half of the multi-GPU correction in one call: value histogram of this rank's p-values below p_cut0, summed
// over the ranks by the library's own all-reduce, global cut + how many p-values lie below it (on all ranks, on this one),
// read back.  work [dev]: two histograms + 8 words; info_host [pinned host]: 8 words ([0] cut as a double, [1] below the
// cut on all ranks, [2] on this rank), valid on return (the call synchronises the stream once).  q_nan (nullable): q, filled
// with 1.0 by the caller; the lines whose p is NaN get q = NaN in the same sweep, so that q is final when nothing lies below
// the cut.
extern "C" int fhc_bh_dist_cut(fhc_comm *comm, const double *p, int64_t n, double T, double p_cut0, uint64_t *work,
                               uint64_t *info_host, double *q_nan, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(comm && work && info_host && n >= 0 && (n == 0 || p != nullptr), FHC_E_INVALID, "fhc_bh_dist_cut: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    u64 *local = reinterpret_cast<u64 *>(work);
    u64 *summed = local + kCutBuckets;
    u64 *info = summed + kCutBuckets;
    FHC_CUDA(cudaMemsetAsync(local, 0, (size_t)kCutBuckets * sizeof(u64), st));
    if (n > 0) {
        const int rc = cut_hist_launch(p, n, p_cut0, local, st, q_nan);
        if (rc != FHC_OK) return rc;
    }
    FHC_CUDA(cudaMemcpyAsync(summed, local, (size_t)kCutBuckets * sizeof(u64), cudaMemcpyDeviceToDevice, st));
    const int rc = fhc_comm_allreduce_u64(comm, reinterpret_cast<uint64_t *>(summed), kCutBuckets, stream);
    if (rc != FHC_OK) return rc;
    FHC_CUDA(cudaFuncSetAttribute(bh_cut_from_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCutFindSmem));
    // pinned host memory is device accessible (unified addressing): the kernel stores its three words there itself and the
    // device-to-host copy of 64 bytes, a copy-engine launch of its own, is saved.  Anything else gets the copy.
    bool direct = false;
    {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, info_host) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
            attr.devicePointer == info_host)
            direct = true;
        else
            cudaGetLastError();
    }
    bh_cut_from_sum_kernel<<<1, 1024, kCutFindSmem, st>>>(summed, local, T, p_cut0, direct ? reinterpret_cast<u64 *>(info_host) : info);
    FHC_LAUNCH_CHECK("bh_cut_from_sum_kernel");
    if (!direct) FHC_CUDA(cudaMemcpyAsync(info_host, info, 8 * sizeof(u64), cudaMemcpyDeviceToHost, st));
    FHC_CUDA(cudaStreamSynchronize(st));
    return FHC_OK;
}

extern "C" int32_t fhc_host_bh_cut_bucket(double p) { return fhc::cut_bucket(p); }

extern "C" double fhc_host_bh_cut_find(const uint64_t *hist, double T, double rank_offset, double p_cut0) {
    using namespace fhc;
    if (hist == nullptr || !(T > 0.0)) return p_cut0;
    u64 c = 0;
    for (int j = 0; j < kCutBuckets; ++j) {
        if (hist[j] == 0) continue;
        c += hist[j];
        if (cut_bucket_closes(j, c, T, rank_offset)) {
            const double e = cut_edge(j);
            return e < p_cut0 ? e : p_cut0;
        }
    }
    return p_cut0;
}

extern "C" uint64_t fhc_bh_key_of(double p) {  // host copy of the order-preserving key (for choosing splitters)
    uint64_t b;
    memcpy(&b, &p, sizeof(b));
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

extern "C" int fhc_bh_partition_count(const double *p, int64_t n, const uint64_t *splitter_keys, int32_t nparts,
                                      double p_cut, uint64_t *counts, void *stream) {
    using namespace fhc;
    Splitters sp;
    const int rc = make_splitters(splitter_keys, nparts, &sp);
    if (rc != FHC_OK) return rc;
    FHC_REQUIRE(n >= 0 && counts != nullptr && (n == 0 || p != nullptr), FHC_E_INVALID, "fhc_bh_partition_count: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    FHC_CUDA(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * nparts, st));
    if (n == 0) return FHC_OK;
    long long blocks = (((n + 3) >> 2) + 255) / 256;
    if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
    bh_part_count_kernel<<<(unsigned int)blocks, 256, 0, st>>>(p, n, sp, p_cut, reinterpret_cast<u64 *>(counts));
    FHC_LAUNCH_CHECK("bh_part_count_kernel");
    return FHC_OK;
}

extern "C" int fhc_bh_partition_scatter(const double *p, int64_t n, const uint64_t *splitter_keys, int32_t nparts,
                                        double p_cut, uint64_t *cursors, double *send, uint32_t *idx, double *q,
                                        int32_t q_prefilled, void *stream) {
    using namespace fhc;
    Splitters sp;
    const int rc = make_splitters(splitter_keys, nparts, &sp);
    if (rc != FHC_OK) return rc;
    FHC_REQUIRE(n >= 0 && n < (1ll << 32) && cursors != nullptr, FHC_E_INVALID, "fhc_bh_partition_scatter: bad arguments");
    if (n == 0) return FHC_OK;
    FHC_REQUIRE(p && send && idx && q, FHC_E_INVALID, "fhc_bh_partition_scatter: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    long long blocks = (n + 1023) / 1024;
    if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
    bh_part_scatter_kernel<<<(unsigned int)blocks, 256, 0, st>>>(p, n, sp, p_cut, reinterpret_cast<u64 *>(cursors), send,
                                                                 idx, q, q_prefilled ? 1 : 0);
    FHC_LAUNCH_CHECK("bh_part_scatter_kernel");
    return FHC_OK;
}

// q-values that are not exactly 1.0 (ranked lines and NaN), as (line, value) pairs: on a sparse map that is all a caller
// has to copy to the host -- the rest of the 8 B per line is the constant 1.0.
namespace fhc {
__global__ void __launch_bounds__(256) gather_ne_one_kernel(const double *__restrict__ q, long long n, u64 capacity,
                                                           u32 *__restrict__ idx, double *__restrict__ val, u64 *count) {
    __shared__ u32 warp_tot[8];
    __shared__ u64 block_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long per_iter = 256ll * 4;
    for (long long b0 = (long long)blockIdx.x * per_iter; b0 < n; b0 += (long long)gridDim.x * per_iter) {
        const long long i0 = b0 + (long long)threadIdx.x * 4;
        double v[4];
        if (i0 + 3 < n) {
            const double2 a = __ldcs(reinterpret_cast<const double2 *>(q + i0));
            const double2 b = __ldcs(reinterpret_cast<const double2 *>(q + i0 + 2));
            v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? q[i0 + k] : 1.0;
        }
        int mine = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) mine += (v[k] == 1.0) ? 0 : 1;  // NaN != 1.0
        const u32 any = __ballot_sync(0xffffffffu, mine != 0);
        int inc = mine;
        if (any) {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
        }
        if (lane == 31) warp_tot[warp] = (u32)inc;
        __syncthreads();
        if (threadIdx.x == 0) {
            u32 tot = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const u32 c = warp_tot[w];
                warp_tot[w] = tot;
                tot += c;
            }
            block_base = tot ? atomicAdd(count, (u64)tot) : 0;
        }
        __syncthreads();
        if (mine) {
            u64 dst = block_base + warp_tot[warp] + (u64)(inc - mine);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!(v[k] == 1.0)) {
                    if (dst < capacity) {
                        idx[dst] = (u32)(i0 + k);
                        val[dst] = v[k];
                    }
                    ++dst;
                }
            }
        }
        __syncthreads();
    }
}
}  // namespace fhc

extern "C" int fhc_gather_ne_one(const double *q, int64_t n, int64_t capacity, uint32_t *idx, double *val, uint64_t *count,
                                 void *stream) {
    using namespace fhc;
    FHC_REQUIRE(n >= 0 && n < (1ll << 32) && capacity >= 0, FHC_E_INVALID, "fhc_gather_ne_one: need 0 <= n < 2^32, capacity >= 0");
    FHC_REQUIRE(count != nullptr, FHC_E_INVALID, "fhc_gather_ne_one: null count");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    FHC_CUDA(cudaMemsetAsync(count, 0, sizeof(u64), st));
    if (n == 0) return FHC_OK;
    FHC_REQUIRE(q != nullptr && aligned16(q) && (capacity == 0 || (idx && val)), FHC_E_INVALID,
                "fhc_gather_ne_one: null or misaligned array");
    long long blocks = (n + 1023) / 1024;
    if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
    gather_ne_one_kernel<<<(unsigned int)blocks, 256, 0, st>>>(q, n, (u64)capacity, idx, val, reinterpret_cast<u64 *>(count));
    FHC_LAUNCH_CHECK("gather_ne_one_kernel");
    return FHC_OK;
}

extern "C" int fhc_scatter_f64(const double *src, const uint32_t *idx, int64_t n, double *dst, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(n >= 0, FHC_E_INVALID, "fhc_scatter_f64: n < 0");
    if (n == 0) return FHC_OK;
    FHC_REQUIRE(src && idx && dst, FHC_E_INVALID, "fhc_scatter_f64: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    scatter_f64_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, st>>>(src, idx, n, dst);
    FHC_LAUNCH_CHECK("scatter_f64_kernel");
    return FHC_OK;
}

// dst[i] = v: the engine fills q with 1.0 while the host bins and fits (the GPU has nothing else to do then), after which
// the kernels of K4 only write the lines whose q is not 1.0 (q_prefilled)
namespace fhc {
__global__ void __launch_bounds__(256) fill_f64_kernel(double *__restrict__ dst, long long n, double v) {
    const long long n2 = n >> 1;
    double2 *d2 = reinterpret_cast<double2 *>(dst);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n2; i += (long long)gridDim.x * 256)
        __stcs(d2 + i, make_double2(v, v));
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) dst[n - 1] = v;
}
}  // namespace fhc

extern "C" int fhc_fill_f64(double *dst, int64_t n, double v, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(n >= 0 && (n == 0 || (dst != nullptr && aligned16(dst))), FHC_E_INVALID, "fhc_fill_f64: null or misaligned array");
    if (n == 0) return FHC_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    long long blocks = ((n >> 1) + 255) / 256;
    if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
    if (blocks < 1) blocks = 1;
    fill_f64_kernel<<<(unsigned int)blocks, 256, 0, st>>>(dst, n, v);
    FHC_LAUNCH_CHECK("fill_f64_kernel");
    return FHC_OK;
}
