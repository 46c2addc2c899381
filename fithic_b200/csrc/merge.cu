// Merge-filter step after a Fit-Hi-C run (reference fithic/utils/CombineNearbyInteraction.py):
// connected components of the significant bin pairs under 8- or 4-connectivity, per-component box statistics and the greedy
// choice of representative loops.  The reference tests every pair of nodes for adjacency (O(n^2), :313-333) and walks the
// bounding box of every component cell by cell (:389-394); here the bin pairs are sorted once (the radix sort of K4), a
// neighbour is a binary search, the components come from a lock-free union-find, and a box row is one range query.
//
// Entries: the n input lines sorted by key = chromosome << 48 | bin1 << 24 | bin2 (stable, so a repeated bin pair keeps its
// FIRST line in front, :306).  The first entry of a run of equal keys is a NODE; the others carry label -1.  A component is
// named by its ROOT, the smallest entry index in it.
//
// Every per-entry step is a __host__ __device__ function: the kernels call it with atomics, fhc_host_merge_* (tests only)
// call the same code serially on host arrays.
#include <algorithm>
#include <limits.h>
#include <string.h>
#include <vector>

#include "common.cuh"

#define FHC_PROFILE_STREAM st

namespace fhc {

#define MHD __host__ __device__ inline

constexpr int kBinBits = 24;
constexpr unsigned int kBinMask = (1u << kBinBits) - 1u;
constexpr int kMergeThreads = 256;
constexpr int kFindGuard = 1 << 22;

MHD uint64_t merge_key(uint32_t chr, uint32_t b1, uint32_t b2) { return ((uint64_t)chr << 48) | ((uint64_t)b1 << kBinBits) | b2; }
MHD int key_b1(uint64_t k) { return (int)((k >> kBinBits) & kBinMask); }
MHD int key_b2(uint64_t k) { return (int)(k & kBinMask); }

MHD int64_t lower_bound_u64(const uint64_t *a, int64_t n, uint64_t x) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t m = (lo + hi) >> 1;
        if (a[m] < x) lo = m + 1; else hi = m;
    }
    return lo;
}

// ---- atomics: the real thing on the device, plain statements in the serial host drivers --------------------------------
MHD int cas_i32(int *p, int cmp, int val) {
#ifdef __CUDA_ARCH__
    return atomicCAS(p, cmp, val);
#else
    const int old = *p;
    if (old == cmp) *p = val;
    return old;
#endif
}
MHD int load_i32(const int *p) {
#ifdef __CUDA_ARCH__
    return *reinterpret_cast<const volatile int *>(p);  // never from a stale L1 line: another CTA may have hooked this root
#else
    return *p;
#endif
}
MHD void store_i32(int *p, int v) {
#ifdef __CUDA_ARCH__
    *reinterpret_cast<volatile int *>(p) = v;
#else
    *p = v;
#endif
}
MHD void add_i32(int *p, int v) {
#ifdef __CUDA_ARCH__
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
MHD void min_i32(int *p, int v) {
#ifdef __CUDA_ARCH__
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}
MHD void max_i32(int *p, int v) {
#ifdef __CUDA_ARCH__
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
MHD void min_u32(unsigned int *p, unsigned int v) {
#ifdef __CUDA_ARCH__
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}
MHD void add_i64(int64_t *p, int64_t v) {
#ifdef __CUDA_ARCH__
    atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v);  // two's complement: fine for v < 0
#else
    *p += v;
#endif
}

// ---- union-find over entry indices; parent[i] <= i always (a larger root is hooked under a smaller one, path halving
// only moves a pointer further down), so there are no cycles and the root of a component is its smallest entry ----------
MHD int uf_find(int *parent, int i, int *err) {
    for (int it = 0; it < kFindGuard; ++it) {
        const int p = load_i32(parent + i);
        if (p == i) return i;
        const int gp = load_i32(parent + p);
        if (gp == p) return p;
        store_i32(parent + i, gp);  // path halving; a concurrent writer can only have stored another ancestor
        i = gp;
    }
    *err = 1;
    return i;
}

MHD void uf_union(int *parent, int a, int b, int *err) {
    for (int it = 0; it < kFindGuard; ++it) {
        a = uf_find(parent, a, err);
        b = uf_find(parent, b, err);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        if (cas_i32(parent + a, a, b) == a) return;  // a was still a root: hooked
    }
    *err = 1;
}

MHD bool is_node(const uint64_t *keys, int64_t i) { return i == 0 || keys[i - 1] != keys[i]; }

// the edges of CombineNearbyInteraction.py:313-333 that lead from node i to a LARGER key (the other half is found from the
// other end): |d bin1| <= 1 and |d bin2| <= 1 (8), |d bin1| + |d bin2| <= 1 (4); any other value of -c adds no edge at all
MHD void merge_link(const uint64_t *keys, int64_t n, int64_t i, int conn, int *parent, int *err) {
    if (!is_node(keys, i)) return;
    const uint64_t k = keys[i];
    const int b1 = key_b1(k), b2 = key_b2(k);
    const int d1[4] = {0, 1, 1, 1}, d2[4] = {1, 0, -1, 1};
    const int nd = conn == 8 ? 4 : (conn == 4 ? 2 : 0);
    for (int t = 0; t < nd; ++t) {
        const int nb1 = b1 + d1[t], nb2 = b2 + d2[t];
        if (nb1 > (int)kBinMask || nb2 < 0 || nb2 > (int)kBinMask) continue;
        const uint64_t nk = merge_key((uint32_t)(k >> 48), (uint32_t)nb1, (uint32_t)nb2);
        const int64_t j = lower_bound_u64(keys, n, nk);
        if (j < n && keys[j] == nk) uf_union(parent, (int)i, (int)j, err);
    }
}

// per component (indexed by root entry): number of nodes, first line, bounding box, sum of counts (:362-385)
struct MergeComp {
    int32_t *size;
    uint32_t *first_line;
    int32_t *box;  // [4 n]: min bin1, max bin1, min bin2, max bin2
    int64_t *sum_cc;
    int64_t *have;  // nodes of ANY component inside the box (:389-394)
};

MHD void merge_comp_init(const MergeComp &C, int64_t i) {
    C.size[i] = 0;
    C.first_line[i] = 0xffffffffu;
    C.box[4 * i + 0] = INT_MAX;
    C.box[4 * i + 1] = INT_MIN;
    C.box[4 * i + 2] = INT_MAX;
    C.box[4 * i + 3] = INT_MIN;
    C.sum_cc[i] = 0;
    C.have[i] = 0;
}

MHD void merge_stat(const uint64_t *keys, const uint32_t *order, const int32_t *label, const int64_t *cc, int64_t i,
                    const MergeComp &C) {
    const int r = label[i];
    if (r < 0) return;
    const uint64_t k = keys[i];
    add_i32(C.size + r, 1);
    min_u32(C.first_line + r, order[i]);
    min_i32(C.box + 4 * (int64_t)r + 0, key_b1(k));
    max_i32(C.box + 4 * (int64_t)r + 1, key_b1(k));
    min_i32(C.box + 4 * (int64_t)r + 2, key_b2(k));
    max_i32(C.box + 4 * (int64_t)r + 3, key_b2(k));
    add_i64(C.sum_cc + r, cc[order[i]]);
}

// nodes with bin1 == a and lo2 <= bin2 <= hi2 on the chromosome of `chr_key`
MHD int64_t merge_box_row(const uint64_t *keys, int64_t n, uint64_t chr_key, int a, int lo2, int hi2) {
    const uint32_t chr = (uint32_t)(chr_key >> 48);
    const uint64_t first = merge_key(chr, (uint32_t)a, (uint32_t)lo2), last = merge_key(chr, (uint32_t)a, (uint32_t)hi2);
    int64_t count = 0;
    for (int64_t j = lower_bound_u64(keys, n, first); j < n && keys[j] <= last; ++j) count += is_node(keys, j) ? 1 : 0;
    return count;
}

// ---- the order in which the reference's heap hands out the nodes of a component (:611-625): q (or -q with -s 1), then the
// larger count, then bin1, bin2.  Entries are in (bin1, bin2) order already; three stable sorts by the keys below finish it.
MHD uint64_t sortkey_count(int64_t cc) { return (uint64_t)(0x7fffffffll - cc); }  // 0 <= cc < 2^31 (checked by the caller)
MHD uint64_t sortkey_q(double q, int sort_order) {
    double hv = sort_order == 0 ? q : -q;
    hv += 0.0;  // -0.0 -> +0.0: the heap compares them equal
#ifdef __CUDA_ARCH__
    const uint64_t b = (uint64_t)__double_as_longlong(hv);
#else
    uint64_t b;
    memcpy(&b, &hv, sizeof b);
#endif
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
MHD uint64_t sortkey_label(int32_t label) { return label < 0 ? 0xffffffffull : (uint64_t)label; }

// top-K % rule (custom_percent, :38-52, applied to the q-values of the component in heap order): position of the cut
MHD int64_t merge_cut_pos(int64_t m, int top_pct) {
    const int64_t idx = (int64_t)((double)(m * top_pct) / 100.0);  // int((len * K) / 100) with Python's true division
    return idx <= 1 ? m - 1 : idx;
}
MHD bool merge_past_cut(double q, double cut, int sort_order) {
    const double hv = sort_order == 0 ? q : -q;
    // :517 -- with -s 1 the heap value is -q and the cut is a q-value: the reference compares them as they are
    return (sort_order == 0 && hv > cut) || (sort_order == 1 && hv < cut);
}
MHD bool merge_near(uint64_t kept, int b1, int b2, int neigh) {
    const int k1 = (int)(kept >> 32), k2 = (int)(kept & 0xffffffffu);
    const int a = b1 > k1 ? b1 - k1 : k1 - b1, b = b2 > k2 ? b2 - k2 : k2 - b2;
    return a <= neigh && b <= neigh;  // |d bin| * res <= Neigh * res on both ends (:540, :676)
}

// ---- kernels ------------------------------------------------------------------------------------------------------------
__global__ void merge_keys_kernel(const int32_t *chr, const int32_t *b1, const int32_t *b2, int64_t n, uint64_t *keys,
                                  uint32_t *vals) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        keys[i] = merge_key((uint32_t)chr[i], (uint32_t)b1[i], (uint32_t)b2[i]);
        vals[i] = (uint32_t)i;
    }
}

__global__ void merge_init_kernel(int *parent, MergeComp C, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        parent[i] = (int)i;
        merge_comp_init(C, i);
    }
}

__global__ void merge_link_kernel(const uint64_t *keys, int64_t n, int conn, int *parent, int *err) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        merge_link(keys, n, i, conn, parent, err);
}

__global__ void merge_label_kernel(const uint64_t *keys, int64_t n, int *parent, int32_t *label, int *err) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        label[i] = is_node(keys, i) ? uf_find(parent, (int)i, err) : -1;
}

__global__ void merge_stat_kernel(const uint64_t *keys, const uint32_t *order, const int32_t *label, const int64_t *cc, int64_t n,
                                  MergeComp C) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        merge_stat(keys, order, label, cc, i, C);
}

// one warp per root: the rows of its box go round the lanes
__global__ void merge_box_kernel(const uint64_t *keys, const int32_t *label, int64_t n, MergeComp C) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; i < n; i += nwarps) {
        if (label[i] != (int)i) continue;
        const int lo1 = C.box[4 * i + 0], hi1 = C.box[4 * i + 1], lo2 = C.box[4 * i + 2], hi2 = C.box[4 * i + 3];
        unsigned long long count = 0;
        for (int a = lo1 + lane; a <= hi1; a += 32) count += (unsigned long long)merge_box_row(keys, n, keys[i], a, lo2, hi2);
        count = warp_sum(count);
        if (lane == 0) C.have[i] = (int64_t)count;
    }
}

__global__ void merge_sortkey_kernel(int which, const uint32_t *entries, const uint32_t *order, const int32_t *label,
                                     const int64_t *cc, const double *q, int sort_order, int64_t n, uint64_t *keys_out) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t e = entries[p];
        keys_out[p] = which == 0 ? sortkey_count(cc[order[e]]) : (which == 1 ? sortkey_q(q[order[e]], sort_order) : sortkey_label(label[e]));
    }
}

__global__ void merge_iota_kernel(uint32_t *v, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = (uint32_t)i;
}

// one warp per component (the warp of the position where its segment of `ranked` starts): the nodes in heap order, each
// checked against the nodes kept so far by all 32 lanes
__global__ void merge_select_kernel(const uint64_t *keys, const uint32_t *order, const int32_t *label, const int32_t *size,
                                    const double *q, int64_t n, int top_pct, int neigh, int sort_order, const uint32_t *ranked,
                                    uint64_t *kept, uint8_t *keep) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (w >= n) return;
    const int r = label[ranked[w]];
    if (r < 0) return;
    if (w > 0 && label[ranked[w - 1]] == r) return;
    const int64_t m = size[r];
    const bool cut_on = top_pct < 100;
    double cut = 0.0;
    if (cut_on) cut = q[order[ranked[w + merge_cut_pos(m, top_pct)]]];
    int64_t nk = 0;
    for (int64_t t = 0; t < m; ++t) {
        const uint32_t e = ranked[w + t];
        if (cut_on && merge_past_cut(q[order[e]], cut, sort_order)) break;
        const uint64_t k = keys[e];
        const int b1 = key_b1(k), b2 = key_b2(k);
        bool near = false;
        for (int64_t j = lane; j < nk && !near; j += 32) near = merge_near(kept[w + j], b1, b2, neigh);
        near = __any_sync(0xffffffffu, near);
        if (!near) {
            if (lane == 0) {
                kept[w + nk] = ((uint64_t)(uint32_t)b1 << 32) | (uint32_t)b2;
                keep[w + t] = 1;
            }
            ++nk;
            __syncwarp();
        }
    }
}

inline unsigned int merge_blocks(int64_t n, int per_thread = 1) {
    long long b = (n / per_thread + kMergeThreads - 1) / kMergeThreads;
    if (b < 1) b = 1;
    if (b > (long long)kNumSMs * 16) b = (long long)kNumSMs * 16;
    return (unsigned int)b;
}

inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace fhc

using namespace fhc;

extern "C" size_t fhc_merge_workspace_bytes(int64_t n) {
    if (n < 0) n = 0;
    const size_t un = (size_t)n;
    // components: error flag, unsorted keys and lines, parent; select: two key and two entry buffers, the kept list
    const size_t a = 256 + up256(8 * un) + up256(4 * un) + up256(4 * un);
    const size_t b = 2 * up256(8 * un) + 2 * up256(4 * un) + up256(8 * un);
    return (a > b ? a : b) + up256(fhc_sort_workspace_bytes(n));
}

#define MERGE_COMMON_CHECKS(name)                                                                                      \
    FHC_REQUIRE(n >= 0 && n < (1ll << 31), FHC_E_INVALID, name ": need 0 <= n < 2^31 (got %lld)", (long long)n);       \
    cudaStream_t st = static_cast<cudaStream_t>(stream);                                                               \
    FHC_PROFILE_ENTRY(st);                                                                                             \
    if (n == 0) return FHC_OK;                                                                                         \
    FHC_REQUIRE(workspace && workspace_bytes >= fhc_merge_workspace_bytes(n), FHC_E_WORKSPACE,                         \
                name ": workspace of %zu bytes, need %zu", workspace_bytes, fhc_merge_workspace_bytes(n))

extern "C" int fhc_merge_components(const int32_t *chr, const int32_t *b1, const int32_t *b2, const int64_t *cc, int64_t n,
                                    int32_t conn, uint64_t *keys, uint32_t *order, int32_t *label, int32_t *size,
                                    uint32_t *first_line, int32_t *box, int64_t *sum_cc, int64_t *have, void *workspace,
                                    size_t workspace_bytes, void *stream) {
    MERGE_COMMON_CHECKS("fhc_merge_components");
    FHC_REQUIRE(chr && b1 && b2 && cc && keys && order && label && size && first_line && box && sum_cc && have, FHC_E_INVALID,
                "fhc_merge_components: null pointer");
    char *base = reinterpret_cast<char *>(workspace);
    int *err = reinterpret_cast<int *>(base);
    uint64_t *keys_in = reinterpret_cast<uint64_t *>(base + 256);
    uint32_t *vals_in = reinterpret_cast<uint32_t *>(base + 256 + up256(8 * (size_t)n));
    int *parent = reinterpret_cast<int *>(base + 256 + up256(8 * (size_t)n) + up256(4 * (size_t)n));
    char *sort_ws = base + 256 + up256(8 * (size_t)n) + 2 * up256(4 * (size_t)n);
    const MergeComp C{size, first_line, box, sum_cc, have};
    const unsigned int blocks = merge_blocks(n);
    FHC_CUDA(cudaMemsetAsync(err, 0, 256, st));
    merge_keys_kernel<<<blocks, kMergeThreads, 0, st>>>(chr, b1, b2, n, keys_in, vals_in);
    FHC_LAUNCH_CHECK("merge_keys_kernel");
    const int rc = fhc_sort_pairs_u64(keys_in, vals_in, keys, order, n, sort_ws, fhc_sort_workspace_bytes(n), stream);
    if (rc != FHC_OK) return rc;
    merge_init_kernel<<<blocks, kMergeThreads, 0, st>>>(parent, C, n);
    FHC_LAUNCH_CHECK("merge_init_kernel");
    merge_link_kernel<<<blocks, kMergeThreads, 0, st>>>(keys, n, conn, parent, err);
    FHC_LAUNCH_CHECK("merge_link_kernel");
    merge_label_kernel<<<blocks, kMergeThreads, 0, st>>>(keys, n, parent, label, err);
    FHC_LAUNCH_CHECK("merge_label_kernel");
    merge_stat_kernel<<<blocks, kMergeThreads, 0, st>>>(keys, order, label, cc, n, C);
    FHC_LAUNCH_CHECK("merge_stat_kernel");
    merge_box_kernel<<<merge_blocks(n * 32), kMergeThreads, 0, st>>>(keys, label, n, C);
    FHC_LAUNCH_CHECK("merge_box_kernel");
    int herr = 0;
    FHC_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    FHC_CUDA(cudaStreamSynchronize(st));
    FHC_REQUIRE(herr == 0, FHC_E_RANGE, "fhc_merge_components: union-find did not settle within its iteration bound");
    return FHC_OK;
}

extern "C" int fhc_merge_select(const uint64_t *keys, const uint32_t *order, const int32_t *label, const int32_t *size,
                                const int64_t *cc, const double *q, int64_t n, int32_t top_pct, int32_t neigh,
                                int32_t sort_order, uint32_t *ranked, uint8_t *keep, void *workspace, size_t workspace_bytes,
                                void *stream) {
    MERGE_COMMON_CHECKS("fhc_merge_select");
    FHC_REQUIRE(keys && order && label && size && cc && q && ranked && keep, FHC_E_INVALID, "fhc_merge_select: null pointer");
    FHC_REQUIRE(top_pct > 0 && top_pct <= 100, FHC_E_INVALID, "fhc_merge_select: top_pct must lie in (0, 100] (got %d)", top_pct);
    FHC_REQUIRE(sort_order == 0 || sort_order == 1, FHC_E_INVALID, "fhc_merge_select: sort_order must be 0 or 1");
    char *base = reinterpret_cast<char *>(workspace);
    const size_t k8 = up256(8 * (size_t)n), k4 = up256(4 * (size_t)n);
    uint64_t *K0 = reinterpret_cast<uint64_t *>(base), *K1 = reinterpret_cast<uint64_t *>(base + k8);
    uint32_t *V[2] = {reinterpret_cast<uint32_t *>(base + 2 * k8), reinterpret_cast<uint32_t *>(base + 2 * k8 + k4)};
    uint64_t *kept = reinterpret_cast<uint64_t *>(base + 2 * k8 + 2 * k4);
    char *sort_ws = base + 3 * k8 + 2 * k4;
    const unsigned int blocks = merge_blocks(n);
    merge_iota_kernel<<<blocks, kMergeThreads, 0, st>>>(V[0], n);
    FHC_LAUNCH_CHECK("merge_iota_kernel");
    int cur = 0;
    for (int which = 0; which < 3; ++which) {
        merge_sortkey_kernel<<<blocks, kMergeThreads, 0, st>>>(which, V[cur], order, label, cc, q, sort_order, n, K0);
        FHC_LAUNCH_CHECK("merge_sortkey_kernel");
        const int rc = fhc_sort_pairs_u64(K0, V[cur], K1, V[cur ^ 1], n, sort_ws, fhc_sort_workspace_bytes(n), stream);
        if (rc != FHC_OK) return rc;
        cur ^= 1;
    }
    FHC_CUDA(cudaMemcpyAsync(ranked, V[cur], sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    FHC_CUDA(cudaMemsetAsync(keep, 0, (size_t)n, st));
    const long long sel_blocks = (n * 32 + kMergeThreads - 1) / kMergeThreads;
    merge_select_kernel<<<(unsigned int)sel_blocks, kMergeThreads, 0, st>>>(keys, order, label, size, q, n, top_pct, neigh,
                                                                           sort_order, ranked, kept, keep);
    FHC_LAUNCH_CHECK("merge_select_kernel");
    return FHC_OK;
}

// ---- serial host drivers of the same per-entry code (tests only; the product path above needs a GPU) -------------------
extern "C" int fhc_host_merge_components(const int32_t *chr, const int32_t *b1, const int32_t *b2, const int64_t *cc,
                                         int64_t n, int32_t conn, uint64_t *keys, uint32_t *order, int32_t *label,
                                         int32_t *size, uint32_t *first_line, int32_t *box, int64_t *sum_cc, int64_t *have) {
    FHC_REQUIRE(n >= 0 && n < (1ll << 31), FHC_E_INVALID, "fhc_host_merge_components: need 0 <= n < 2^31");
    if (n == 0) return FHC_OK;
    std::vector<uint64_t> k((size_t)n);
    std::vector<uint32_t> idx((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        k[i] = merge_key((uint32_t)chr[i], (uint32_t)b1[i], (uint32_t)b2[i]);
        idx[i] = (uint32_t)i;
    }
    std::stable_sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) { return k[a] < k[b]; });
    for (int64_t i = 0; i < n; ++i) {
        keys[i] = k[idx[i]];
        order[i] = idx[i];
    }
    std::vector<int> parent((size_t)n);
    const MergeComp C{size, first_line, box, sum_cc, have};
    int err = 0;
    for (int64_t i = 0; i < n; ++i) {
        parent[i] = (int)i;
        merge_comp_init(C, i);
    }
    for (int64_t i = 0; i < n; ++i) merge_link(keys, n, i, conn, parent.data(), &err);
    for (int64_t i = 0; i < n; ++i) label[i] = is_node(keys, i) ? uf_find(parent.data(), (int)i, &err) : -1;
    for (int64_t i = 0; i < n; ++i) merge_stat(keys, order, label, cc, i, C);
    for (int64_t i = 0; i < n; ++i) {
        if (label[i] != (int)i) continue;
        int64_t count = 0;
        for (int a = box[4 * i + 0]; a <= box[4 * i + 1]; ++a) count += merge_box_row(keys, n, keys[i], a, box[4 * i + 2], box[4 * i + 3]);
        have[i] = count;
    }
    FHC_REQUIRE(err == 0, FHC_E_RANGE, "fhc_host_merge_components: union-find did not settle within its iteration bound");
    return FHC_OK;
}

extern "C" int fhc_host_merge_select(const uint64_t *keys, const uint32_t *order, const int32_t *label, const int32_t *size,
                                     const int64_t *cc, const double *q, int64_t n, int32_t top_pct, int32_t neigh,
                                     int32_t sort_order, uint32_t *ranked, uint8_t *keep) {
    FHC_REQUIRE(n >= 0 && n < (1ll << 31), FHC_E_INVALID, "fhc_host_merge_select: need 0 <= n < 2^31");
    FHC_REQUIRE(top_pct > 0 && top_pct <= 100, FHC_E_INVALID, "fhc_host_merge_select: top_pct must lie in (0, 100]");
    FHC_REQUIRE(sort_order == 0 || sort_order == 1, FHC_E_INVALID, "fhc_host_merge_select: sort_order must be 0 or 1");
    if (n == 0) return FHC_OK;
    std::vector<uint32_t> v((size_t)n);
    std::vector<uint64_t> k((size_t)n);
    for (int64_t i = 0; i < n; ++i) v[i] = (uint32_t)i;
    for (int which = 0; which < 3; ++which) {
        for (int64_t e = 0; e < n; ++e)
            k[e] = which == 0 ? sortkey_count(cc[order[e]]) : (which == 1 ? sortkey_q(q[order[e]], sort_order) : sortkey_label(label[e]));
        std::stable_sort(v.begin(), v.end(), [&](uint32_t a, uint32_t b) { return k[a] < k[b]; });
    }
    std::vector<uint64_t> kept((size_t)n);
    for (int64_t w = 0; w < n; ++w) {
        ranked[w] = v[w];
        keep[w] = 0;
    }
    for (int64_t w = 0; w < n; ++w) {
        const int r = label[ranked[w]];
        if (r < 0 || (w > 0 && label[ranked[w - 1]] == r)) continue;
        const int64_t m = size[r];
        const bool cut_on = top_pct < 100;
        const double cut = cut_on ? q[order[ranked[w + merge_cut_pos(m, top_pct)]]] : 0.0;
        int64_t nk = 0;
        for (int64_t t = 0; t < m; ++t) {
            const uint32_t e = ranked[w + t];
            if (cut_on && merge_past_cut(q[order[e]], cut, sort_order)) break;
            const int b1 = key_b1(keys[e]), b2 = key_b2(keys[e]);
            bool near = false;
            for (int64_t j = 0; j < nk && !near; ++j) near = merge_near(kept[w + j], b1, b2, neigh);
            if (!near) {
                kept[w + nk++] = ((uint64_t)(uint32_t)b1 << 32) | (uint32_t)b2;
                keep[w + t] = 1;
            }
        }
    }
    return FHC_OK;
}
