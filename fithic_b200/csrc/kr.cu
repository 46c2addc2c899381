// Knight-Ruiz matrix balancing on the contact matrix: the kernels behind fithic_b200/hickry.py, the B200 replacement of the
// reference's bias generator fithic/utils/HiCKRy.py (the producer of Fit-Hi-C's `-t` file).
//
// The matrix never exists as a matrix.  HiCKRy builds coo(z, (x, y)) + its transpose (HiCKRy.py:46-50) and drops the
// sparsest rows and columns (:76-96); here the contact lines stay as they were read -- (row locus, column locus, count)
// in file order, 16 bytes per line -- and a per-locus remap (new index, or -1 for a dropped locus) does the dropping.
//   kr_spmv_kernel       y = (M + M^T) x over the kept loci: y[r] += z x[c] and y[c] += z x[r] per line (a diagonal line
//                        therefore counts twice, as in the reference).  HBM-bound: 16 B per line; x and y (8 B per locus)
//                        stay in L2.  Contact files list all partners of a locus on consecutive lines, so the row side is
//                        reduced inside the warp by segments of equal row before it touches memory (one atomic per
//                        segment); the column side goes out as one atomic per line to distinct addresses.
//   kr_* vector kernels  the element-wise steps and dot products of knightRuizAlg (HiCKRy.py:140-232), fused per
//                        statement group, every product and sum rounded separately like numpy's (no FMA contraction), so
//                        that only the ORDER of the sums (the products with the matrix and the dot products) can differ
//                        from the reference; every reduction leaves one partial per CTA (fixed grid) for the host to add.
#define FHC_PROFILE_STREAM st
#include "common.cuh"

namespace fhc {

constexpr int kKrThreads = 256;
constexpr int kKrBlocks = kNumSMs * 4;  // fixed grid: partial sums land in partial[kKrBlocks]

__device__ __forceinline__ void atomic_add_f64(double *p, double v) { atomicAdd(p, v); }

__global__ void __launch_bounds__(kKrThreads) kr_spmv_kernel(const int *__restrict__ rows, const int *__restrict__ cols,
                                                            const double *__restrict__ vals, long long nnz,
                                                            const int *__restrict__ remap, const double *__restrict__ x,
                                                            double *__restrict__ y) {
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * kKrThreads;
    // every lane of a warp runs the same number of iterations (the shuffles below need the whole warp)
    const long long first = (long long)blockIdx.x * kKrThreads + (threadIdx.x & ~31);
    for (long long w0 = first; w0 < nnz; w0 += stride) {
        const long long e = w0 + lane;
        int r = -1, c = -1;
        double v = 0.0;
        if (e < nnz) {
            r = __ldg(remap + rows[e]);
            c = __ldg(remap + cols[e]);
            v = vals[e];
        }
        const bool live = r >= 0 && c >= 0;
        double to_r = 0.0;
        if (live) {
            to_r = __dmul_rn(v, __ldg(x + c));
            atomic_add_f64(y + c, __dmul_rn(v, __ldg(x + r)));  // column side: distinct addresses in a sorted file
        }
        const int key = live ? r : -1 - lane;  // dead lanes never merge with a neighbour
        // segmented inclusive scan over runs of equal row (contiguous lanes only), then the last lane of a run adds it
        const int prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = lane == 0 || prev != key;
        const unsigned int heads = __ballot_sync(0xffffffffu, head);
        const int seg_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
        double acc = to_r;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, acc, o);
            if (lane - o >= seg_start) acc += t;
        }
        const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
        if (live && tail) atomic_add_f64(y + r, acc);
    }
}

// ---- reductions: one partial per CTA ---------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double *sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kKrThreads / 32; ++w) t += sm[w];
    __syncthreads();
    return t;  // valid in thread 0
}
__device__ __forceinline__ double block_min(double v, double *sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = INFINITY;
    if (threadIdx.x == 0)
        for (int w = 0; w < kKrThreads / 32; ++w) t = fmin(t, sm[w]);
    __syncthreads();
    return t;
}

#define KR_LOOP(i, n) for (long long i = (long long)blockIdx.x * kKrThreads + threadIdx.x; i < (n); i += (long long)gridDim.x * kKrThreads)

// out = a * b
__global__ void __launch_bounds__(kKrThreads) kr_mul_kernel(const double *a, const double *b, double *out, long long n) {
    KR_LOOP(i, n) out[i] = __dmul_rn(a[i], b[i]);
}
// v = x * Ax; rk = 1 - v; partial = sum rk^2      (HiCKRy.py:155-157, :208-211)
__global__ void __launch_bounds__(kKrThreads) kr_residual_kernel(const double *x, const double *Ax, double *v, double *rk,
                                                               long long n, double *partial) {
    __shared__ double sm[kKrThreads / 32];
    double s = 0.0;
    KR_LOOP(i, n) {
        const double vi = __dmul_rn(x[i], Ax[i]);
        const double r = __dsub_rn(1.0, vi);
        v[i] = vi;
        rk[i] = r;
        s = __dadd_rn(s, __dmul_rn(r, r));
    }
    s = block_sum(s, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// Z = rk / v; p = Z; partial = sum rk * Z          (:178-182)
__global__ void __launch_bounds__(kKrThreads) kr_first_kernel(const double *rk, const double *v, double *Z, double *p,
                                                            long long n, double *partial) {
    __shared__ double sm[kKrThreads / 32];
    double s = 0.0;
    KR_LOOP(i, n) {
        const double z = rk[i] / v[i];
        Z[i] = z;
        p[i] = z;
        s = __dadd_rn(s, __dmul_rn(rk[i], z));
    }
    s = block_sum(s, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// p = Z + beta * p; xp = x * p                      (:184-185 and the argument of A.dot in :191)
__global__ void __launch_bounds__(kKrThreads) kr_direction_kernel(const double *Z, double beta, int first, double *p,
                                                                const double *x, double *xp, long long n) {
    KR_LOOP(i, n) {
        const double pi = first ? p[i] : __dadd_rn(Z[i], __dmul_rn(beta, p[i]));
        p[i] = pi;
        xp[i] = __dmul_rn(x[i], pi);
    }
}
// w = x * Axp + v * p; partial = sum p * w          (:191-193)
__global__ void __launch_bounds__(kKrThreads) kr_w_kernel(const double *x, const double *Axp, const double *v, const double *p,
                                                        double *w, long long n, double *partial) {
    __shared__ double sm[kKrThreads / 32];
    double s = 0.0;
    KR_LOOP(i, n) {
        const double wi = __dadd_rn(__dmul_rn(x[i], Axp[i]), __dmul_rn(v[i], p[i]));
        w[i] = wi;
        s = __dadd_rn(s, __dmul_rn(p[i], wi));
    }
    s = block_sum(s, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// partial[b] = min(y + alpha p), partial[grid + b] = -max(y + alpha p)      (:196-197, :206)
__global__ void __launch_bounds__(kKrThreads) kr_ynew_minmax_kernel(const double *y, double alpha, const double *p, long long n,
                                                                  double *partial) {
    __shared__ double sm[kKrThreads / 32];
    double mn = INFINITY, mx = INFINITY;  // mx holds -max
    KR_LOOP(i, n) {
        const double yn = __dadd_rn(y[i], __dmul_rn(alpha, p[i]));
        mn = fmin(mn, yn);
        mx = fmin(mx, -yn);
    }
    mn = block_min(mn, sm);
    mx = block_min(mx, sm);
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = mn;
        partial[gridDim.x + blockIdx.x] = mx;
    }
}
// gamma = min over the selected i of (bound - y) / (alpha p):  mode 0: alpha p < 0 (:201-203), mode 1: y + alpha p > bound (:207-209)
__global__ void __launch_bounds__(kKrThreads) kr_gamma_kernel(const double *y, double alpha, const double *p, double bound,
                                                            int mode, long long n, double *partial) {
    __shared__ double sm[kKrThreads / 32];
    double g = INFINITY;
    KR_LOOP(i, n) {
        const double ap = __dmul_rn(alpha, p[i]);
        const bool sel = mode == 0 ? (ap < 0.0) : (__dadd_rn(y[i], ap) > bound);
        if (sel) g = fmin(g, __ddiv_rn(__dsub_rn(bound, y[i]), ap));
    }
    g = block_min(g, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = g;
}
// y += gamma * (alpha * p)                          (:204, :210)
__global__ void __launch_bounds__(kKrThreads) kr_axpy_kernel(double *y, double gamma, double alpha, const double *p, long long n) {
    KR_LOOP(i, n) y[i] = __dadd_rn(y[i], __dmul_rn(gamma, __dmul_rn(alpha, p[i])));
}
// y += alpha p; rk -= alpha w; Z = rk / v; partial = sum rk * Z        (:213-218)
__global__ void __launch_bounds__(kKrThreads) kr_update_kernel(double *y, double alpha, const double *p, double *rk,
                                                             const double *w, const double *v, double *Z, long long n,
                                                             double *partial) {
    __shared__ double sm[kKrThreads / 32];
    double s = 0.0;
    KR_LOOP(i, n) {
        y[i] = __dadd_rn(y[i], __dmul_rn(alpha, p[i]));
        const double r = __dsub_rn(rk[i], __dmul_rn(alpha, w[i]));
        rk[i] = r;
        const double z = r / v[i];
        Z[i] = z;
        s = __dadd_rn(s, __dmul_rn(r, z));
    }
    s = block_sum(s, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

}  // namespace fhc

using namespace fhc;

#define KR_ENTRY(name, cond)                                                       \
    FHC_REQUIRE(n >= 0, FHC_E_INVALID, name ": n < 0");                            \
    cudaStream_t st = static_cast<cudaStream_t>(stream);                           \
    FHC_PROFILE_ENTRY(st);                                                         \
    if (n == 0) return FHC_OK;                                                     \
    FHC_REQUIRE(cond, FHC_E_INVALID, name ": null pointer")

extern "C" int32_t fhc_kr_partials(void) { return kKrBlocks; }

extern "C" int fhc_kr_spmv(const int32_t *rows, const int32_t *cols, const double *vals, int64_t nnz, const int32_t *remap,
                           const double *x, double *y, int64_t n, void *stream) {
    FHC_REQUIRE(nnz >= 0, FHC_E_INVALID, "fhc_kr_spmv: nnz < 0");
    KR_ENTRY("fhc_kr_spmv", x && y && (nnz == 0 || (rows && cols && vals && remap)));
    FHC_CUDA(cudaMemsetAsync(y, 0, sizeof(double) * (size_t)n, st));
    if (nnz == 0) return FHC_OK;
    long long blocks = (nnz + kKrThreads - 1) / kKrThreads;
    if (blocks > (long long)kNumSMs * 8) blocks = (long long)kNumSMs * 8;
    kr_spmv_kernel<<<(unsigned int)blocks, kKrThreads, 0, st>>>(rows, cols, vals, nnz, remap, x, y);
    FHC_LAUNCH_CHECK("kr_spmv_kernel");
    return FHC_OK;
}

extern "C" int fhc_kr_mul(const double *a, const double *b, double *out, int64_t n, void *stream) {
    KR_ENTRY("fhc_kr_mul", a && b && out);
    kr_mul_kernel<<<kKrBlocks, kKrThreads, 0, st>>>(a, b, out, n);
    FHC_LAUNCH_CHECK("kr_mul_kernel");
    return FHC_OK;
}
extern "C" int fhc_kr_residual(const double *x, const double *Ax, double *v, double *rk, int64_t n, double *partial,
                               void *stream) {
    KR_ENTRY("fhc_kr_residual", x && Ax && v && rk && partial);
    kr_residual_kernel<<<kKrBlocks, kKrThreads, 0, st>>>(x, Ax, v, rk, n, partial);
    FHC_LAUNCH_CHECK("kr_residual_kernel");
    return FHC_OK;
}
extern "C" int fhc_kr_first(const double *rk, const double *v, double *Z, double *p, int64_t n, double *partial, void *stream) {
    KR_ENTRY("fhc_kr_first", rk && v && Z && p && partial);
    kr_first_kernel<<<kKrBlocks, kKrThreads, 0, st>>>(rk, v, Z, p, n, partial);
    FHC_LAUNCH_CHECK("kr_first_kernel");
    return FHC_OK;
}
extern "C" int fhc_kr_direction(const double *Z, double beta, int32_t first, double *p, const double *x, double *xp, int64_t n,
                                void *stream) {
    KR_ENTRY("fhc_kr_direction", Z && p && x && xp);
    kr_direction_kernel<<<kKrBlocks, kKrThreads, 0, st>>>(Z, beta, first, p, x, xp, n);
    FHC_LAUNCH_CHECK("kr_direction_kernel");
    return FHC_OK;
}
extern "C" int fhc_kr_w(const double *x, const double *Axp, const double *v, const double *p, double *w, int64_t n,
                        double *partial, void *stream) {
    KR_ENTRY("fhc_kr_w", x && Axp && v && p && w && partial);
    kr_w_kernel<<<kKrBlocks, kKrThreads, 0, st>>>(x, Axp, v, p, w, n, partial);
    FHC_LAUNCH_CHECK("kr_w_kernel");
    return FHC_OK;
}
extern "C" int fhc_kr_ynew_minmax(const double *y, double alpha, const double *p, int64_t n, double *partial, void *stream) {
    KR_ENTRY("fhc_kr_ynew_minmax", y && p && partial);
    kr_ynew_minmax_kernel<<<kKrBlocks, kKrThreads, 0, st>>>(y, alpha, p, n, partial);
    FHC_LAUNCH_CHECK("kr_ynew_minmax_kernel");
    return FHC_OK;
}
extern "C" int fhc_kr_gamma(const double *y, double alpha, const double *p, double bound, int32_t mode, int64_t n,
                            double *partial, void *stream) {
    KR_ENTRY("fhc_kr_gamma", y && p && partial);
    kr_gamma_kernel<<<kKrBlocks, kKrThreads, 0, st>>>(y, alpha, p, bound, mode, n, partial);
    FHC_LAUNCH_CHECK("kr_gamma_kernel");
    return FHC_OK;
}
extern "C" int fhc_kr_axpy(double *y, double gamma, double alpha, const double *p, int64_t n, void *stream) {
    KR_ENTRY("fhc_kr_axpy", y && p);
    kr_axpy_kernel<<<kKrBlocks, kKrThreads, 0, st>>>(y, gamma, alpha, p, n);
    FHC_LAUNCH_CHECK("kr_axpy_kernel");
    return FHC_OK;
}
extern "C" int fhc_kr_update(double *y, double alpha, const double *p, double *rk, const double *w, const double *v, double *Z,
                             int64_t n, double *partial, void *stream) {
    KR_ENTRY("fhc_kr_update", y && p && rk && w && v && Z && partial);
    kr_update_kernel<<<kKrBlocks, kKrThreads, 0, st>>>(y, alpha, p, rk, w, v, Z, n, partial);
    FHC_LAUNCH_CHECK("kr_update_kernel");
    return FHC_OK;
}
