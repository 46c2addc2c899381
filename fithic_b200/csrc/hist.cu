// K1 -- distance histogram + observed totals.
//
// Replaces the accumulation loop of read_Interactions (reference fithic/fithic.py:406-441) and the classification of
// myUtils.Interaction.getType (fithic/myUtils.py:135-148).
//
// Design (B200): HBM-bound streaming pass, 16 algorithmic bytes per contact (4 x int32 SoA, 128-bit loads, 4 contacts
// per thread per tile).  One persistent 1024-thread CTA per SM keeps a private uint32 histogram of the whole distance
// axis in shared memory (D <= 53,248 slots = 208 KB covers 5 kb whole-genome: chr1 = 49,792 slots), so the hot
// short-distance slots never leave the SM; the CTA flushes to the global uint64 histogram every 2^20 contacts (bounds
// every 32-bit partial sum below 2^32 because only counts < 4096 take the shared path) and at the end.  Totals are kept
// in registers and reduced warp -> CTA -> one global atomic per CTA.
#define FHC_PROFILE_STREAM st
#include "common.cuh"

namespace fhc {

constexpr int kHistThreads = 1024;
constexpr int kHistPairsPerTile = kHistThreads * 4;
constexpr int kHistSmemSlotsMax = 53248;    // 208 KB of uint32 slots (227 KB usable per CTA)
constexpr int kHistSmallCount = 4096;       // counts below this go through shared memory
constexpr int kHistFlushTiles = 256;        // 256 tiles x 4096 contacts x 4095 < 2^32
constexpr int kHistMaxRuns = FHC_MAX_CHR_RUNS;

struct HistAcc {
    unsigned long long inrange_sum = 0, intra_sum = 0, inter_sum = 0;
    unsigned int inter_n = 0, inrange_n = 0, intra_n = 0, offgrid = 0, nonpos = 0;
    int maxc = 0;
};

// what a contact needs besides its own fields, in 32-bit form (mid points are int32, so every distance fits 32 bits)
struct HistConst {
    unsigned int Llo, Uhi;   // in-range window clamped to 32 bits
    unsigned int none;       // 1: L beyond 2^32 - 1 (nothing is in range)
    unsigned int res;
    unsigned int div_m, div_sh;  // d / res for d < 2^31 by one multiply-high (see pvalue_lists.cu: FrontConst)
    unsigned int D32;        // number of slots clamped to 32 bits
    unsigned int S;          // slots kept in shared memory
};

// INTRA: the caller knows the line is intra-chromosomal (chromosome runs), so no chromosome ids are looked at.
// Written without early exits: every total takes a select-and-add, so that the four lines of a thread run as one straight
// instruction stream (the version with returns spent a third of its instructions on branches and reconvergence).
template <bool INTRA>
__device__ __forceinline__ void hist_one(int m1, int m2, int c, unsigned int ch, bool skipped, const HistConst &K,
                                         unsigned int *sh, unsigned long long *hist, unsigned int *present, HistAcc &a) {
    a.maxc = max(a.maxc, c);
    const unsigned long long cs = (unsigned long long)(long long)c;
    const bool kept = !skipped;
    const bool intra = INTRA ? kept : (kept && (ch & 0xffffu) == (ch >> 16));
    if (!INTRA) {  // inter (fithic/fithic.py:420-422)
        const bool inter = kept && !intra;
        a.inter_sum += inter ? cs : 0ull;
        a.inter_n += inter ? 1u : 0u;
    }
    a.intra_sum += intra ? cs : 0ull;  // any type of intra (:423-425)
    a.intra_n += intra ? 1u : 0u;
    const unsigned int d = m1 > m2 ? (unsigned int)m1 - (unsigned int)m2 : (unsigned int)m2 - (unsigned int)m1;
    const bool inr = intra && d >= K.Llo && d <= K.Uhi;  // (an unreachable window comes as Llo > Uhi); else intraShort / Long
    a.inrange_sum += inr ? cs : 0ull;  // :439
    a.inrange_n += inr ? 1u : 0u;
    const unsigned int slot = K.res == 1 ? d : (d < 0x80000000u ? (__umulhi(d, K.div_m) >> K.div_sh) : d / K.res);
    const bool on_grid = slot * K.res == d && slot < K.D32;
    a.offgrid += (inr && !on_grid) ? 1u : 0u;
    const bool counted = inr && on_grid;
    const bool in_smem = (unsigned int)(c - 1) < (unsigned int)(kHistSmallCount - 1) && slot < K.S;  // 0 < c < kHistSmallCount
    if (counted && in_smem) {
        atomicAdd(&sh[slot], (unsigned int)c);
    } else if (counted) {
        if (c != 0) atomicAdd(&hist[slot], cs);
        if (c <= 0) {
            atomicOr(&present[slot >> 5], 1u << (slot & 31));
            a.nonpos += 1;
        }
    }
}

__device__ __forceinline__ void hist_flush(unsigned int *sh, int S, unsigned long long *hist) {
    __syncthreads();
    for (int s = threadIdx.x; s < S; s += kHistThreads) {
        const unsigned int v = sh[s];
        if (v) {
            atomicAdd(&hist[s], (unsigned long long)v);
            sh[s] = 0;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kHistThreads, 1)
hist_distance_kernel(const int4 *__restrict__ mid1, const int4 *__restrict__ mid2, const int4 *__restrict__ cnt,
                     const int4 *__restrict__ chrs, const long long *__restrict__ run_start,
                     const unsigned int *__restrict__ run_val, int nruns, const unsigned int *__restrict__ skip,
                     long long skip_limit, long long n, const HistConst K, unsigned long long *hist,
                     unsigned int *present, unsigned long long *scalars, int max_slot) {
    const int S = (int)K.S;
    extern __shared__ unsigned int sh[];
    __shared__ unsigned long long red[FHC_N_SCALARS];
    // chromosome ids as runs (contact files are grouped by chromosome): run r covers lines [rs[r], rs[r + 1])
    __shared__ long long rs[kHistMaxRuns + 1];
    __shared__ unsigned int rv[kHistMaxRuns];
    if (chrs == nullptr) {
        for (int r = threadIdx.x; r <= nruns; r += kHistThreads) rs[r] = run_start[r];
        for (int r = threadIdx.x; r < nruns; r += kHistThreads) rv[r] = run_val[r];
    }
    int run = 0;  // this thread's lines only move forward, so does its position in the run table
    for (int s = threadIdx.x; s < S; s += kHistThreads) sh[s] = 0;
    if (threadIdx.x < FHC_N_SCALARS) red[threadIdx.x] = 0;
    __syncthreads();

    HistAcc a;
    const long long ntiles = n / kHistPairsPerTile;  // full tiles; the tail is handled below
    int since_flush = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long g = t * kHistThreads + threadIdx.x;  // index of this thread's group of 4 contacts
        const int4 a1 = ldg_stream(mid1 + g), a2 = ldg_stream(mid2 + g), ac = ldg_stream(cnt + g);
        const long long i0 = g * 4;
        int4 ah = make_int4(0, 0, 0, 0);
        bool all_intra = false;  // the four lines lie in one intra-chromosomal run: no chromosome ids to look at
        if (chrs != nullptr) {
            ah = ldg_stream(chrs + g);
        } else {
            while (rs[run + 1] <= i0) ++run;
            const int v = (int)rv[run];
            ah = make_int4(v, v, v, v);
            if (rs[run + 1] < i0 + 4) {  // the group of four straddles a run boundary
                int r2 = run;
                while (rs[r2 + 1] <= i0 + 1) ++r2;
                ah.y = (int)rv[r2];
                while (rs[r2 + 1] <= i0 + 2) ++r2;
                ah.z = (int)rv[r2];
                while (rs[r2 + 1] <= i0 + 3) ++r2;
                ah.w = (int)rv[r2];
            } else {
                all_intra = ((unsigned int)v & 0xffffu) == ((unsigned int)v >> 16);
            }
        }
        unsigned int sk = 0;
        if (skip != nullptr) {
            sk = __ldg(skip + g);
            // a line is dropped when flagged and not past the reference's stalled pointer (:408-412)
            if (i0 + 0 > skip_limit) sk &= ~0x000000ffu;
            if (i0 + 1 > skip_limit) sk &= ~0x0000ff00u;
            if (i0 + 2 > skip_limit) sk &= ~0x00ff0000u;
            if (i0 + 3 > skip_limit) sk &= ~0xff000000u;
        }
        if (all_intra) {
            hist_one<true>(a1.x, a2.x, ac.x, 0u, (sk & 0x000000ffu) != 0, K, sh, hist, present, a);
            hist_one<true>(a1.y, a2.y, ac.y, 0u, (sk & 0x0000ff00u) != 0, K, sh, hist, present, a);
            hist_one<true>(a1.z, a2.z, ac.z, 0u, (sk & 0x00ff0000u) != 0, K, sh, hist, present, a);
            hist_one<true>(a1.w, a2.w, ac.w, 0u, (sk & 0xff000000u) != 0, K, sh, hist, present, a);
        } else {
            hist_one<false>(a1.x, a2.x, ac.x, (unsigned int)ah.x, (sk & 0x000000ffu) != 0, K, sh, hist, present, a);
            hist_one<false>(a1.y, a2.y, ac.y, (unsigned int)ah.y, (sk & 0x0000ff00u) != 0, K, sh, hist, present, a);
            hist_one<false>(a1.z, a2.z, ac.z, (unsigned int)ah.z, (sk & 0x00ff0000u) != 0, K, sh, hist, present, a);
            hist_one<false>(a1.w, a2.w, ac.w, (unsigned int)ah.w, (sk & 0xff000000u) != 0, K, sh, hist, present, a);
        }
        if (++since_flush == kHistFlushTiles) {
            hist_flush(sh, S, hist);
            since_flush = 0;
        }
    }
    // tail: fewer than one tile of contacts, scalar loads, spread over the first CTA
    if (blockIdx.x == 0) {
        const int *m1 = reinterpret_cast<const int *>(mid1), *m2 = reinterpret_cast<const int *>(mid2);
        const int *cc = reinterpret_cast<const int *>(cnt);
        const unsigned int *hh = reinterpret_cast<const unsigned int *>(chrs);
        const unsigned char *sb = reinterpret_cast<const unsigned char *>(skip);
        for (long long i = ntiles * kHistPairsPerTile + threadIdx.x; i < n; i += kHistThreads) {
            const bool skipped = sb != nullptr && sb[i] != 0 && i <= skip_limit;
            unsigned int ch;
            if (chrs != nullptr) {
                ch = hh[i];
            } else {
                int lo = 0, hi = nruns - 1;  // last run that starts at or before i
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (rs[mid] <= i)
                        lo = mid;
                    else
                        hi = mid - 1;
                }
                ch = rv[lo];
            }
            hist_one<false>(m1[i], m2[i], cc[i], ch, skipped, K, sh, hist, present, a);
        }
    }
    hist_flush(sh, S, hist);

    // totals: warp shuffle -> shared atomics -> one global atomic per CTA and scalar
    unsigned long long v[FHC_N_SCALARS];
    v[FHC_S_INTRA_INRANGE_SUM] = warp_sum(a.inrange_sum);
    v[FHC_S_INTRA_ALL_SUM] = warp_sum(a.intra_sum);
    v[FHC_S_INTER_ALL_SUM] = warp_sum(a.inter_sum);
    v[FHC_S_INTER_ALL_COUNT] = warp_sum((unsigned long long)a.inter_n);
    v[FHC_S_OFFGRID] = warp_sum((unsigned long long)a.offgrid);
    v[FHC_S_INTRA_INRANGE_LINES] = warp_sum((unsigned long long)a.inrange_n);
    v[FHC_S_INTRA_ALL_LINES] = warp_sum((unsigned long long)a.intra_n);
    v[FHC_S_NONPOS_LINES] = warp_sum((unsigned long long)a.nonpos);
    int mc = a.maxc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mc = max(mc, __shfl_xor_sync(0xffffffffu, mc, o));
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < FHC_N_SCALARS; ++k)
            if (k != FHC_S_MAX_COUNT && v[k]) atomicAdd(&red[k], v[k]);
        atomicMax(&red[FHC_S_MAX_COUNT], (unsigned long long)mc);
    }
    __syncthreads();
    if (threadIdx.x < FHC_N_SCALARS) {
        const unsigned long long r = red[threadIdx.x];
        if (threadIdx.x == FHC_S_MAX_COUNT)
            atomicMax(&scalars[max_slot], r);
        else if (r)
            atomicAdd(&scalars[threadIdx.x], r);
    }
}

}  // namespace fhc

extern "C" int fhc_hist_distance(const int32_t *mid1, const int32_t *mid2, const int32_t *cnt, const uint32_t *chrs,
                                 const int64_t *run_start, const uint32_t *run_val, int32_t nruns, const uint8_t *skip, int64_t skip_limit, int64_t n, int64_t L, int64_t U, int32_t res,
                                 uint64_t *hist, uint32_t *present, int64_t D, uint64_t *scalars, int32_t n_rank_slots,
                                 int32_t my_slot, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(n >= 0 && D > 0 && res > 0, FHC_E_INVALID, "fhc_hist_distance: need n >= 0, D > 0, res > 0 (got %lld, %lld, %d)",
                (long long)n, (long long)D, res);
    FHC_REQUIRE(hist && present && scalars, FHC_E_INVALID, "fhc_hist_distance: null output pointer");
    FHC_REQUIRE(n == 0 || (mid1 && mid2 && cnt), FHC_E_INVALID, "fhc_hist_distance: null input pointer");
    FHC_REQUIRE(n == 0 || chrs != nullptr || (run_start && run_val && nruns >= 1 && nruns <= FHC_MAX_CHR_RUNS), FHC_E_INVALID,
                "fhc_hist_distance: need chrs or 1 <= nruns <= %d chromosome runs", FHC_MAX_CHR_RUNS);
    FHC_REQUIRE(aligned16(mid1) && aligned16(mid2) && aligned16(cnt) && aligned16(chrs) && aligned16(skip), FHC_E_INVALID,
                "fhc_hist_distance: input arrays must be 16-byte aligned");
    FHC_REQUIRE(L >= -1 && U >= -1, FHC_E_INVALID, "fhc_hist_distance: L and U must be >= -1");
    FHC_REQUIRE(n_rank_slots >= 0 && (n_rank_slots == 0 || (my_slot >= 0 && my_slot < n_rank_slots)), FHC_E_INVALID,
                "fhc_hist_distance: bad rank slots (%d of %d)", my_slot, n_rank_slots);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    FHC_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint64_t) * D, st));
    FHC_CUDA(cudaMemsetAsync(present, 0, sizeof(uint32_t) * ((D + 31) / 32), st));
    FHC_CUDA(cudaMemsetAsync(scalars, 0, sizeof(uint64_t) * (FHC_N_SCALARS + n_rank_slots), st));
    if (n == 0) return FHC_OK;
    const int S = (int)(D < kHistSmemSlotsMax ? D : kHistSmemSlotsMax);
    const size_t smem = sizeof(unsigned int) * (size_t)S;
    FHC_CUDA(cudaFuncSetAttribute(hist_distance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(sizeof(unsigned int) * kHistSmemSlotsMax)));
    const long long ntiles = n / kHistPairsPerTile;
    int grid = (int)(ntiles < kNumSMs ? (ntiles > 0 ? ntiles : 1) : kNumSMs);
    const long long Llo = L < 0 ? 0 : L;
    const long long Uhi = U < 0 ? INT64_MAX : U;
    HistConst K;
    K.none = Llo > 0xffffffffll ? 1u : 0u;
    K.Llo = (unsigned int)(Llo > 0xffffffffll ? 0xffffffffll : Llo);
    K.Uhi = (unsigned int)(Uhi > 0xffffffffll ? 0xffffffffll : Uhi);
    if (K.none) {  // nothing can be in range: an empty window
        K.Llo = 1u;
        K.Uhi = 0u;
    }
    K.res = (unsigned int)res;
    {
        unsigned int l = 0;
        while ((1ull << l) < (unsigned long long)res) ++l;  // ceil(log2 res)
        K.div_sh = l ? l - 1 : 0;
        K.div_m = res > 1 ? (unsigned int)(((1ull << (31 + l)) + (unsigned long long)res - 1) / (unsigned long long)res) : 0u;
    }
    K.D32 = (unsigned int)(D > 0xffffffffll ? 0xffffffffll : D);
    K.S = (unsigned int)S;
    hist_distance_kernel<<<grid, kHistThreads, smem, st>>>(
        reinterpret_cast<const int4 *>(mid1), reinterpret_cast<const int4 *>(mid2), reinterpret_cast<const int4 *>(cnt),
        reinterpret_cast<const int4 *>(chrs), reinterpret_cast<const long long *>(run_start), run_val, nruns,
        reinterpret_cast<const unsigned int *>(skip), skip_limit, n, K, reinterpret_cast<unsigned long long *>(hist), present,
        reinterpret_cast<unsigned long long *>(scalars), n_rank_slots > 0 ? FHC_N_SCALARS + my_slot : FHC_S_MAX_COUNT);
    FHC_LAUNCH_CHECK("hist_distance_kernel");
    return FHC_OK;
}

// ---- span of the mid points -----------------------------------------------------------------------------------------------
// The length of the distance axis (D) follows from the largest |mid1 - mid2| an intra line can have, which is bounded by
// max(mid) - min(mid): one streaming pass over the two mid-point arrays when contacts arrive (8 B per line), instead of
// four library reductions.  out[0] = min over both arrays, out[1] = max (int64; INT64_MAX / INT64_MIN when n == 0).
namespace fhc {
__global__ void __launch_bounds__(256) mid_range_kernel(const int4 *__restrict__ mid1, const int4 *__restrict__ mid2, long long n,
                                                       long long *out) {
    int lo = INT32_MAX, hi = INT32_MIN;
    const long long ng = n >> 2;
    for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g < ng; g += (long long)gridDim.x * 256) {
        const int4 a = ldg_stream(mid1 + g), b = ldg_stream(mid2 + g);
        lo = min(min(min(a.x, a.y), min(a.z, a.w)), min(lo, min(min(b.x, b.y), min(b.z, b.w))));
        hi = max(max(max(a.x, a.y), max(a.z, a.w)), max(hi, max(max(b.x, b.y), max(b.z, b.w))));
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
        const long long i = (ng << 2) + threadIdx.x;
        const int a = reinterpret_cast<const int *>(mid1)[i], b = reinterpret_cast<const int *>(mid2)[i];
        lo = min(lo, min(a, b));
        hi = max(hi, max(a, b));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
        atomicMin(out, (long long)lo);
        atomicMax(out + 1, (long long)hi);
    }
}
}  // namespace fhc

extern "C" int fhc_mid_range(const int32_t *mid1, const int32_t *mid2, int64_t n, int64_t *out, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(n >= 0 && out != nullptr, FHC_E_INVALID, "fhc_mid_range: n < 0 or null output");
    FHC_REQUIRE(n == 0 || (mid1 && mid2 && aligned16(mid1) && aligned16(mid2)), FHC_E_INVALID,
                "fhc_mid_range: null or misaligned array");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    FHC_PROFILE_ENTRY(st);
    const long long init[2] = {INT64_MAX, INT64_MIN};
    FHC_CUDA(cudaMemcpyAsync(out, init, sizeof(init), cudaMemcpyHostToDevice, st));
    if (n == 0) return FHC_OK;
    long long blocks = ((n >> 2) + 255) / 256;
    if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
    if (blocks < 1) blocks = 1;
    mid_range_kernel<<<(unsigned int)blocks, 256, 0, st>>>(reinterpret_cast<const int4 *>(mid1),
                                                           reinterpret_cast<const int4 *>(mid2), n,
                                                           reinterpret_cast<long long *>(out));
    FHC_LAUNCH_CHECK("mid_range_kernel");
    return FHC_OK;
}
