/*
 * fithic_b200 -- C ABI of the B200-native Fit-Hi-C significance path.
 *
 * The reference (ay-lab/fithic) is pure Python and has no FFI seam of its own; the seams it does have are the Python
 * functions of fithic/fithic.py and fithic/myStats.py.  Each entry point below replaces the per-contact (or per-bin)
 * loop inside one of those functions, cited as <file>:<lines> relative to the reference root.  The Python host
 * (fithic_b200/fithic.py) keeps the reference's function names and argument meaning and binds these symbols through
 * ctypes (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / C++ types.  `stream` is a cudaStream_t passed as void* (NULL = default).
 *   - every pointer marked [dev] is device memory owned by the caller (16-byte aligned), [host] is host memory.
 *   - every call returns 0 on success or a negative FHC_E_* code; fhc_last_error() gives the message (thread local).
 *   - device entry points are asynchronous on `stream`; they allocate nothing except where a *_workspace_bytes
 *     companion exists, in which case the caller passes the workspace.
 *   - contacts ("pairs") are a structure of arrays, one element per line of the contact-counts file, in file order:
 *        mid1[i], mid2[i]  fragment mid points (int32)
 *        cnt[i]            contact count, already truncated toward zero (fithic/fithic.py:415, myUtils.py:123-124)
 *        chrs[i]           chromosome ids, chr1 | chr2 << 16 (uint32); intra <=> both halves equal
 */
#ifndef FITHIC_B200_H
#define FITHIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FHC_OK 0
#define FHC_E_INVALID (-1)   /* bad argument (null pointer, misaligned pointer, negative size ...) */
#define FHC_E_CUDA (-2)      /* a CUDA runtime call or kernel launch failed */
#define FHC_E_RANGE (-3)     /* value outside what the reference itself defines (e.g. N >= 2^31) */
#define FHC_E_WORKSPACE (-4) /* workspace too small */

#define FHC_ABI_VERSION 9

/* fhc_hist_distance scalars[] layout (uint64 each, two's complement where signed) */
#define FHC_S_INTRA_INRANGE_SUM 0 /* observedIntraInRangeSum  fithic/fithic.py:439 */
#define FHC_S_INTRA_ALL_SUM 1     /* observedIntraAllSum      fithic/fithic.py:424 */
#define FHC_S_INTER_ALL_SUM 2     /* observedInterAllSum      fithic/fithic.py:421 */
#define FHC_S_INTER_ALL_COUNT 3   /* observedInterAllCount    fithic/fithic.py:422 */
#define FHC_S_MAX_COUNT 4         /* largest cnt[i] over kept lines (sizes the lbeta table) */
#define FHC_S_OFFGRID 5           /* in-range intra lines whose distance is not k*res with k < D (must be 0) */
#define FHC_S_INTRA_INRANGE_LINES 6 /* observedIntraInRangeCount fithic/fithic.py:440 */
#define FHC_S_INTRA_ALL_LINES 7   /* observedIntraAllCount    fithic/fithic.py:425 */
#define FHC_S_NONPOS_LINES 8     /* in-range intra lines with cnt <= 0: only then the `present` bitmap says more than hist */
#define FHC_N_SCALARS 9
#define FHC_MAX_CHR_RUNS 1024      /* chromosome ids as runs: at most this many (fhc_hist_distance, fhc_pvalues) */

/* fhc_pvalues mode bits (fithic/fithic.py:232-248) */
#define FHC_MODE_INTRA_ONLY 0
#define FHC_MODE_INTER_ONLY 1
#define FHC_MODE_ALL 2

int fhc_abi_version(void);
const char *fhc_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
int64_t fhc_launch_count(void);

/* Per-kernel device timing for bench.py's roofline: when enabled every kernel launch of this library is followed by a
 * cudaEventRecord on its stream.  fhc_profile_collect synchronises the device, writes
 * {"<kernel>": {"ms": total, "launches": n}, ...} into buf and clears the log; returns the string length. */
int fhc_profile_enable(int on);
int fhc_profile_collect(char *buf, size_t buf_bytes);

/* Plumbing for hosts that hold device memory but have no CUDA binding of their own (the Python host keeps torch tensors
 * as buffers): cudaMemcpyAsync(cudaMemcpyDefault) on `stream` -- host memory should be pinned -- and
 * cudaStreamSynchronize. */
int fhc_copy_async(void *dst, const void *src, size_t bytes, void *stream);
int fhc_stream_synchronize(void *stream);
/* ... and a CUDA event (no timing): record it behind a copy, go on launching, wait for the copy alone. */
int fhc_event_create(void **event_out);
int fhc_event_record(void *event, void *stream);
int fhc_event_synchronize(void *event);
int fhc_event_destroy(void *event);

/* FP64 FMA throughput of the current device in TFLOP/s (2 flops per FMA, 16 independent chains per thread): the compute
 * roofline bench.py reports K3 against (K3 is bound by FP64 latency and instruction issue, not by HBM).  seconds <= 0: best
 * of five bursts of ~10 ms; otherwise back-to-back launches for about that long (sustained clocks).  scratch [dev]: one
 * double.  Synchronises the stream. */
int fhc_peak_fp64(double seconds, double *scratch, double *tflops_out, void *stream);

/* ---- single-node collectives over NVLink peer memory (csrc/comm.cu) ---------------------------------------------------
 * The two small exchanges of a multi-GPU spline pass (the sum of [histogram | totals] over the GPUs, and the
 * gathering of every rank's value histogram) without a communication library: one process per GPU, every rank owns a window
 * in device memory that its peers map through CUDA IPC; a collective is two small kernels on the caller's stream (push the
 * payload into every peer's window, then wait for everybody's flag and sum / copy locally) and never synchronises the host.
 *   fhc_comm_create    allocates this rank's window (slot_bytes = largest payload of one rank) and writes its handle
 *                      (fhc_comm_handle_bytes() bytes) for the host to pass to every other rank (any transport)
 *   fhc_comm_connect   all_handles = the handles of ranks 0 .. world-1 back to back
 *   fhc_comm_allreduce_u64  data[i] <- sum over ranks, in place (n even: the payload moves in 16-byte words)
 *   fhc_comm_allgather dst[r * bytes ...] <- src of rank r (bytes a multiple of 16)
 *   fhc_comm_failed    1 when a bounded wait inside a collective gave up (a peer never arrived)
 * Every rank must issue the same collectives in the same order.  world <= 16, one node. */
typedef struct fhc_comm fhc_comm;
int64_t fhc_comm_handle_bytes(void);
int fhc_comm_create(int32_t rank, int32_t world, int64_t slot_bytes, fhc_comm **comm_out, void *handle_out);
int fhc_comm_connect(fhc_comm *comm, const void *all_handles);
int fhc_comm_allreduce_u64(fhc_comm *comm, uint64_t *data, int64_t n, void *stream);
int fhc_comm_allgather(fhc_comm *comm, const void *src, void *dst, int64_t bytes, void *stream);
int32_t fhc_comm_world(fhc_comm *comm);
int32_t fhc_comm_rank(fhc_comm *comm);
int fhc_comm_failed(fhc_comm *comm);
int fhc_comm_destroy(fhc_comm *comm);

/* ---- K1: distance histogram + totals ------------------------------------------------------------------------
 * Replaces the accumulation loop of read_Interactions (fithic/fithic.py:406-441) with the classification of
 * myUtils.Interaction.getType (fithic/myUtils.py:135-148).
 *   hist[k]    += cnt[i] for every kept intra line with L <= d <= U, d = |mid1-mid2| = k*res  (mainDic[d][1])
 *   present bit k is set when such a line has cnt <= 0 (so "distance seen" survives a zero sum, :434-436)
 *   skip       nullable per-line outlier multiplicity (pass >= 2, :408-412); line i is dropped when
 *              skip[i] != 0 and i <= skip_limit (the reference's pointer walk stalls at the first duplicate)
 *   L, U       distLowThres / distUpThres; -1 = unbounded
 * hist, present and scalars are zeroed by the callee.  present has (D+31)/32 words.
 * chrs may be NULL when the chromosome ids come as runs instead (contact files are grouped by chromosome; 4 B per line
 * less to read): run r of nruns <= FHC_MAX_CHR_RUNS covers lines [run_start[r], run_start[r+1]) (run_start[0] = 0,
 * run_start[nruns] = n) and all of them have chrs = run_val[r]  [dev].
 * n_rank_slots / my_slot: every total is a sum and survives a sum over GPUs, the largest count does not.  With
 * n_rank_slots > 0, scalars has FHC_N_SCALARS + n_rank_slots entries and the largest count goes to
 * scalars[FHC_N_SCALARS + my_slot] (scalars[FHC_S_MAX_COUNT] stays 0): after ONE all-reduce(sum) of [hist | scalars]
 * every rank holds every rank's maximum.  n_rank_slots = 0: scalars[FHC_S_MAX_COUNT] as before. */
int fhc_hist_distance(const int32_t *mid1, const int32_t *mid2, const int32_t *cnt, const uint32_t *chrs,
                      const int64_t *run_start, const uint32_t *run_val, int32_t nruns, const uint8_t *skip, int64_t skip_limit, int64_t n, int64_t L, int64_t U, int32_t res,
                      uint64_t *hist, uint32_t *present, int64_t D, uint64_t *scalars, int32_t n_rank_slots,
                      int32_t my_slot, void *stream);

/* Smallest and largest mid point over both arrays: out[0] = min, out[1] = max (int64 [dev]; INT64_MAX / INT64_MIN when
 * n == 0).  Every |mid1 - mid2| is at most out[1] - out[0], which sizes the distance axis D of fhc_hist_distance. */
int fhc_mid_range(const int32_t *mid1, const int32_t *mid2, int64_t n, int64_t *out, void *stream);

/* ---- host helpers for the O(D) sequential stages (bit-exact integer/float bookkeeping) ------------------------
 * makeBinsFromInteractions, fithic/fithic.py:463-553.  dists/sums [host]: the m distinct in-range distances
 * ascending with their count sums.  outl_dec [host, nullable]: not used here (see fhc_host_frag_pairs).
 * Writes bin_lb/bin_ub/bin_sumcc [host, capacity noOfBins]; returns the number of bins (>= 0) or an error. */
int fhc_host_make_bins(const int64_t *dists, const int64_t *sums, int64_t m, int32_t noOfBins, int64_t N,
                       int64_t *bin_lb, int64_t *bin_ub, int64_t *bin_sumcc);

/* generate_FragPairs, fixed-size branch, fithic/fithic.py:596-689.  Per chromosome (ALREADY in the reference's
 * sorted-name order): chr_n[c] mappable loci and chr_maxmid[c] = max(mid) over them (maxFrag = max(mid) - res/2,
 * :602).  bin_pairs (in/out, int64) carries the pass>=2 outlier decrements on entry ([1] and [7] move together).
 * Outputs: bin_pairs, bin_sumdist, totals[0]=possibleIntraInRangeCount (x2 quirk kept, :618+:642),
 * totals[1]=possibleIntraAllCount*2, totals[2]=sum n*(noOfFrags-n) (=2*possibleInterAllCount), totals[3]=noOfFrags */
int fhc_host_frag_pairs(const int64_t *chr_n, const int64_t *chr_maxmid, int32_t nchr, int32_t res, int64_t L,
                        int64_t U, const int64_t *bin_lb, const int64_t *bin_ub, int32_t nbins, int64_t *bin_pairs,
                        double *bin_sumdist, int64_t *totals);

/* generate_FragPairs, restriction-fragment branch (-r 0), fithic/fithic.py:691-778.  mids [host]: ascending mid points of
 * the mappable fragments, chromosome after chromosome in sorted-name order (chr_off[nchr+1]).  bin_pairs1 (`[1]`: one per
 * pair in range) and bin_pairs7 (`[7]`: npairs = templen - d per pair, the reference's quirk) carry the pass>=2 outlier
 * decrements on entry.  totals[5]: possibleIntraInRangeCount, possibleIntraAllCount, sum n*(noOfFrags-n), noOfFrags,
 * maxPossibleGenomicDist.  Same accumulation order as the reference: bit exact. */
int fhc_host_frag_pairs_varsize(const int64_t *mids, const int64_t *chr_off, int32_t nchr, int64_t L, int64_t U,
                                const int64_t *bin_lb, const int64_t *bin_ub, int32_t nbins, int64_t *bin_pairs1,
                                int64_t *bin_pairs7, double *bin_sumdist, int64_t *totals);

/* The same on the GPU from prefix sums over the sorted mid points (csrc/fragpairs.cu): O(log n) per (fragment, bin) instead
 * of one step per pair in range -- what a genome-wide restriction map without -U needs (1e10 pairs on chr1 alone).  Host
 * arrays in and out like fhc_host_frag_pairs_varsize; copies and kernel run on `stream`, which is synchronised.  `[1]`, `[7]`
 * and the totals are exact; `[3]` is the correctly rounded exact sum (128-bit integers, one division by 1e6), which differs
 * from the reference's term-by-term double accumulation by that accumulation's rounding (~1e-13 relative). */
int fhc_frag_pairs_varsize(const int64_t *mids, const int64_t *chr_off, int32_t nchr, int64_t L, int64_t U,
                           const int64_t *bin_lb, const int64_t *bin_ub, int32_t nbins, int64_t *bin_pairs1,
                           int64_t *bin_pairs7, double *bin_sumdist, int64_t *totals, void *stream);
/* the kernel's per-cell code run serially on the host (CPU tests) */
int fhc_host_frag_pairs_varsize_prefix(const int64_t *mids, const int64_t *chr_off, int32_t nchr, int64_t L, int64_t U,
                                       const int64_t *bin_lb, const int64_t *bin_ub, int32_t nbins, int64_t *bin_pairs1,
                                       int64_t *bin_pairs7, double *bin_sumdist, int64_t *totals);

/* ---- the whole host stage between K1 and K3 in one call ----------------------------------------------------------
 * k1buf [host]: fhc_hist_distance's outputs back to back as the engine keeps them, [hist (D) | scalars (FHC_N_SCALARS +
 * n_rank_slots) | present words]; the present words are only read when scalars[FHC_S_NONPOS_LINES] != 0; with
 * n_rank_slots > 0 the largest count is the largest of the rank slots (see fhc_hist_distance).  phases (bit mask):
 *   1  observed distances (dists, sums: capacity D) and makeBinsFromInteractions (bin_lb / bin_ub / bin_sumcc: noOfBins)
 *   2  generate_FragPairs fixed-size branch (bin_pairs, bin_sumdist, totals as fhc_host_frag_pairs; `dec` [nullable] holds
 *      the pass >= 2 outlier decrements per bin) and calculateProbabilities (x_bins, y_bins in bin order); the lbeta tables
 *      (fhc_host_lbeta_table; [0]: N = scalars[INTRA_INRANGE_SUM], [1]: N = scalars[INTER_ALL_SUM]; nullable; ntab from
 *      scalars[MAX_COUNT] as fhc_pvalues wants it) are built by the same worker threads
 *   4  fit_Spline's fit stage when want_spline != 0: (xs, ys) sorted by x, fhc_host_curfit (t, c: capacity noOfBins + 4),
 *      splineX / table (capacity D), lut (capacity D): what fhc_spline_table produces on the device
 *   8  the lbeta tables alone (status 3 of an earlier call: lbeta_cap was below lbeta_ntab)
 *  16  calculateProbabilities alone.  With pairs_world > 1 phase 2 only sums the bins that rank pairs_rank owns (0.0 in
 *      bin_sumdist elsewhere) and stops before the probabilities: the caller adds bin_sumdist over the ranks (every entry is
 *      one rank's sum plus zeros, i.e. unchanged bits) and calls phase 16 (| 4)
 * status: 0 fine; 1 x of the bins not strictly increasing at bad_index (the reference prints an error and exits 2,
 * fithic/fithic.py:940-945); 2 no observed distance inside [min x, max x]; 3 an lbeta table is too small; 4 fewer than 4
 * bins (scipy refuses the fit).  timings [ms]: bins, pairs + lbeta, fit, evaluation, antitonic, lut, -, total.
 * nthreads: the calling thread + (nthreads - 1) pooled workers (fhc_host_pool_prewarm wakes them ahead of the call). */
typedef struct fhc_stage_io {
    const uint64_t *k1buf;
    int64_t D;
    int32_t grid, noOfBins;
    int64_t L, U;
    const int64_t *chr_n, *chr_maxmid;
    int32_t nchr, want_spline, nthreads, n_rank_slots;
    const int64_t *dec;
    double *lbeta_tab[2];
    int64_t lbeta_cap[2];
    /* outputs */
    int64_t *dists, *sums;
    int64_t nseen;
    int64_t *bin_lb, *bin_ub, *bin_sumcc, *bin_pairs;
    double *bin_sumdist, *x_bins, *y_bins, *xs, *ys, *t, *c;
    int64_t *splineX;
    double *table, *lut;
    int64_t m;
    int64_t totals[4];
    int64_t lbeta_ntab[2];
    int32_t nb, nt, ier, calls, status, bad_index;
    double fp;
    double timings[8];
    int32_t pairs_rank, pairs_world; /* in: > 1 ranks share the possible-pair sums of phase 2 (see phases) */
    void *shm;                       /* in, nullable: fhc_shm_open handle -- phase 2 then adds the sums of all ranks itself */
} fhc_stage_io;
int fhc_host_stage(fhc_stage_io *io, int32_t phases);
/* A small all-reduce between the ranks of one node through POSIX shared memory, for data that lives on the HOST (the
 * possible-pair sums of the host stage): rank 0 creates `name`, the others open it.  slot_bytes: largest payload. */
int fhc_shm_open(const char *name, int32_t rank, int32_t world, int64_t slot_bytes, void **handle_out);
int fhc_shm_allreduce_u64(void *handle, uint64_t *data, int32_t n);
int fhc_shm_close(void *handle);
int fhc_host_pool_prewarm(int32_t nthreads);
/* diagnostic: njobs jobs spinning job_us microseconds each on nthreads threads -> elapsed ms */
double fhc_host_pool_selftest(int32_t nthreads, int32_t njobs, int32_t job_us, int32_t *distinct_threads);

/* dst[i] = v for i < n with nthreads host threads (the end-to-end call fills its pinned q array with 1.0 while the GPU
 * works, see fhc_gather_ne_one). */
int fhc_host_fill_f64(double *dst, int64_t n, double v, int32_t nthreads);

/* ---- K2: spline table ------------------------------------------------------------------------------------------
 * Replaces ius(splineX) + IsotonicRegression(increasing=False) of fit_Spline (fithic/fithic.py:952-966) and bakes
 * the clamp + bisect lookup of :1066-1068 into a dense table:
 *   table[j] = antitonic(splev(t, c, 3, splineX[j]))                      j < m
 *   lut[k]   = table[min(bisect_left(splineX, clamp(k*res, xmin, xmax)), m-1)]   k < D
 * t, c: FITPACK knots/coefficients (nt each, c zero padded) [dev]; splineX [dev] int64 ascending.
 * workspace: fhc_spline_workspace_bytes(m) bytes [dev]. */
size_t fhc_spline_workspace_bytes(int64_t m);
int fhc_spline_table(const double *t, const double *c, int32_t nt, const int64_t *splineX, int64_t m, double xmin,
                     double xmax, int32_t res, double *table, double *lut, int64_t D, void *workspace,
                     size_t workspace_bytes, void *stream);

/* The three stages of fhc_spline_table one by one, for callers that pool on the host (the PAVA cascade is a chain of
 * dependent steps; a CPU core runs it from L1 faster than one GPU thread from L2 when the table is large):
 *   fhc_spline_eval      y[j] = splev(t, c, 3, splineX[j])                      [dev]
 *   fhc_host_antitonic   y <- IsotonicRegression(increasing=False)(y), in place [host]
 *   fhc_spline_lut       lut[k] = table[min(bisect_left(splineX, clamp(k*res, xmin, xmax)), m-1)]   [dev] */
int fhc_spline_eval(const double *t, const double *c, int32_t nt, const int64_t *splineX, int64_t m, double *y,
                    void *stream);
int fhc_host_antitonic(double *y, int64_t m);
int fhc_spline_lut(const int64_t *splineX, const double *table, int64_t m, double xmin, double xmax, int32_t res,
                   double *lut, int64_t D, void *stream);

/* ---- lbeta table -----------------------------------------------------------------------------------------------
 * tab[c] = cephes lbeta(c, N - c + 1) for 1 <= c < ntab, the only third-party quantity whose ROUNDING matters at
 * 1e-6: scipy.special.bdtrc -> xsf::cephes::incbet -> lbeta, reached from fithic/fithic.py:1070,:1101. */
int fhc_lbeta_table(int64_t N, double *tab, int64_t ntab, void *stream);

/* Host builds of the same source the device table kernel runs (log evaluated in double-double and rounded once, cephes
 * lgam/lbeta with explicit round-to-nearest steps): lets CPU-only tests pin the table arithmetic.  Not a product path. */
double fhc_host_log_cr(double x);
/* fhc_lbeta_table on the HOST with the C library's log -- the one scipy's cephes calls on this machine -- instead of the
 * correctly rounded log of the device kernel: identical except where the library's log is misrounded (about one argument in
 * 15,000 at some magnitudes), where it follows scipy by one ulp of lgam(N) (4e-6 ... 8e-6 in p).  A caller that wants
 * scipy's value there too uploads this table instead of running fhc_lbeta_table (engine: FHC_LBETA_TABLE=host). */
int fhc_host_lbeta_table(int64_t N, double *tab, int64_t ntab, int32_t nthreads);

/* ---- smoothing-spline fit (host) -------------------------------------------------------------------------------
 * `ius = UnivariateSpline(x, y, s=min(y)**2)` of fit_Spline, fithic/fithic.py:951 (scipy FITPACK curfit, third party):
 * cubic, unit weights, the same sequence of IEEE operations as Dierckx's fpcurf, so t / c equal `ius._eval_args` bit for
 * bit.  x, y [host]: m > 3 points, x increasing.  t, c [host]: room for m + 4 doubles; *n_out knots are returned.
 * *ier_out: FITPACK's ier (<= 0 fine; 1, 2, 3 = the warnings scipy prints); *calls_out: 1, or 2 when the first run hit
 * its storage limit nest = max(m / 2, 8) and was continued with nest = m + 4 (UnivariateSpline._reset_nest). */
int fhc_host_curfit(const double *x, const double *y, int32_t m, double s, double *t, double *c, int32_t *n_out,
                    double *fp_out, int32_t *ier_out, int32_t *calls_out);
double fhc_host_lbeta(double a, double b);
/* scipy.special.bdtrc(count - 1, N, prior) evaluated on the host by the source the work-list kernels of K3 run
 * (classification, division-free continued fraction / tail sum, prefactor with folded divisions), and its 1 - exp(y). */
double fhc_host_bdtrc_lists(int32_t count, int64_t N, double prior);
double fhc_host_one_minus_exp(double y);
/* Numerator and denominator of the lower tail sum K3 uses where the count lies below its expectation (cephes incbet's
 * swapped branch, scipy.special.bdtrc at fithic/fithic.py:1070,:1101): in_place != 0 the form the finish kernel runs for
 * sums of up to 32 terms, else the form of the iterate kernel's queue.  Both must give the same bits (tests). */
void fhc_host_tail_sum(int32_t count, int64_t N, double prior, int32_t in_place, double *num, double *den);

/* ---- K3: per-contact p-value ---------------------------------------------------------------------------------
 * Replaces the per-line loop of fit_Spline (fithic/fithic.py:1017-1123) including scipy.special.bdtrc (:1070,:1101).
 *   bias       nullable dense per-locus bias (-1 = discarded by read_biases, :818-832) with bias_mid the mid point
 *              each slot was read for and chr_off[nchr+1] the first slot of each chromosome id; the slot of
 *              (chr, mid) is chr_off[chr] + mid / res; a slot outside the chromosome or with another mid is "missing"
 *              (-1, :1026-1054)
 *   lut        K2's table (intra in-range prior by distance slot), D entries; may be NULL in inter-only mode
 *   N_intra    observedIntraInRangeSum, N_inter observedInterAllSum (both < 2^31, else FHC_E_RANGE: SURVEY F5)
 *   lbeta_*    fhc_lbeta_table outputs for N_intra / N_inter (nullable => computed per contact)
 *   outl       nullable per-line outlier multiplicity, incremented where p < outl_thres (:1215-1217);
 *              outl_stats[0] += lines flagged now, outl_stats[1] = min(itself, line index whose multiplicity reached
 *              >= 2) -- the caller initialises outl_stats to {0, UINT64_MAX} before the first pass
 *   line_base  index in the whole file of the first contact passed (a caller may score the file slice by slice; the
 *              index is only used for outl_stats[1])
 *   p, expcc   outputs, one double per line (:1119-1122)
 *   bias_sparse  1 = no grid (restriction-fragment mode, -r 0): bias / bias_mid hold, per chromosome, the loci of the bias
 *              file in ascending mid order (chr_off[c] .. chr_off[c+1]) and a locus is found by binary search
 *   bias_mid   may be NULL when every slot s of chromosome c holds the locus at mid = (s - chr_off[c]) * res + res / 2
 *              (fixed-size bins on the regular grid): the mid point is then checked arithmetically, one gather less
 *   chrs       may be NULL (with a workspace) when the chromosome ids come as runs, as in fhc_hist_distance: run r covers
 *              lines [run_start[r], run_start[r+1]) COUNTED LIKE line_base (the first contact passed is line line_base of
 *              the run table's numbering) and all of them have chrs = run_val[r].  4 B per line less to read, and a tile of
 *              contacts inside one intra run is scored without any per-contact chromosome logic.
 *   pre_code, pre_b12  nullable (both or neither): the output of fhc_pvalues_prepass for the same contacts
 *   workspace  [dev, nullable] fhc_pvalues_workspace_bytes(n, ntab) bytes, ntab = max(ntab_intra, ntab_inter).  With a
 *              workspace the contacts that need an iterative evaluation (continued fraction / tail sum) are compacted
 *              into work lists in HBM and the call runs as three kernels with full warps (pvalue_lists.cu); without one
 *              a single tile-phased kernel does everything (pvalue.cu; also forced by FHC_PVAL_IMPL=tile).  Both give
 *              the same p-values to ~1e-13 relative. */
size_t fhc_pvalues_workspace_bytes(int64_t n, int64_t ntab);
int fhc_pvalues(int32_t mode, const int32_t *mid1, const int32_t *mid2, const int32_t *cnt, const uint32_t *chrs,
                const int64_t *run_start, const uint32_t *run_val, int32_t nruns, int64_t n, const double *bias, const int32_t *bias_mid, const int64_t *chr_off, int32_t nchr,
                int32_t bias_sparse, int32_t res, int64_t L, int64_t U, const double *lut, int64_t D, int64_t N_intra, int64_t N_inter,
                double interChrProb, double tL, double tU, const double *lbeta_intra, int64_t ntab_intra,
                const double *lbeta_inter, int64_t ntab_inter, uint8_t *outl, int64_t line_base, double outl_thres,
                uint64_t *outl_stats, double *p, double *expcc, const uint32_t *pre_code, const double *pre_b12,
                void *workspace, size_t workspace_bytes, void *stream);

/* The part of fhc_pvalues that does not need the spline table, for every contact: the two bias lookups, their product and
 * window test, the class of the line (the branch order of fithic/fithic.py:1057-1115) and its distance slot.  Launched right
 * after fhc_hist_distance it runs while the host bins and fits; fhc_pvalues (work-list pipeline) then takes the two arrays
 * as pre_code / pre_b12 and reads neither mid points, chromosome ids nor the bias table (mid1 / mid2 / chrs may be NULL
 * there).  The output depends on the contacts, the bias table, res, L, U, tL, tU and the mode only: later spline passes of
 * a run reuse it.  code [dev, uint32 n]: 0xffffffff = not scored (p = 1, ExpCC = 0), else bit 31 = both biases inside
 * [tL, tU], bit 30 = scored against the inter-chromosomal prior, low bits = distance slot; b12 [dev, double n] = rn(bias1 *
 * bias2).  Arguments as in fhc_pvalues. */
int fhc_pvalues_prepass(int32_t mode, const int32_t *mid1, const int32_t *mid2, const uint32_t *chrs,
                        const int64_t *run_start, const uint32_t *run_val, int32_t nruns, int64_t n, const double *bias,
                        const int32_t *bias_mid, const int64_t *chr_off, int32_t nchr, int32_t bias_sparse, int32_t res,
                        int64_t L, int64_t U, double tL, double tU, int64_t line_base, uint32_t *code, double *b12,
                        void *stream);

/* scipy.special.bdtrc(k, n, prior) element-wise on device arrays (the arithmetic core of K3, exposed for parity
 * tests against the oracle; call sites fithic/fithic.py:1070,:1101).  lbeta nullable. */
int fhc_bdtrc(const int32_t *cnt_minus_1, int64_t N, const double *prior, int64_t n, const double *lbeta,
              int64_t ntab, double *out, void *stream);

/* ---- K4: q-values -----------------------------------------------------------------------------------------------
 * Replaces myStats.benjamini_hochberg_correction (fithic/myStats.py:24-48): ascending order, bh = p*T/rank capped
 * at 1, FORWARD running max, p == 1.0 -> 1.0, NaN -> NaN (sorted last).  rank_offset / carry_in allow a caller that
 * range-partitions p-values over several GPUs to chain the scan (single GPU: 0 and 0.0); carry_out [dev, nullable]
 * receives {running max after the last rankable element} and n_sorted_out [dev, nullable] the number of rankable
 * p-values (p != 1, not NaN).  p-values that are certain to end at q = 1.0 are not sorted at all (see fhc_bh_p_cut).
 *   p [dev] n doubles, q [dev] n doubles (may not alias p). */
size_t fhc_bh_workspace_bytes(int64_t n);
int fhc_bh_qvalues(const double *p, int64_t n, double T, int64_t rank_offset, double carry_in, double *q,
                   double *carry_out, int64_t *n_sorted_out, void *workspace, size_t workspace_bytes, void *stream);

/* fhc_bh_qvalues with ONE host synchronisation of `stream` in the middle: the number of ranked keys is read back after the
 * compaction (*n_ranked_host [host, nullable]) and only the passes that number needs are launched -- none at all when no
 * p-value is below the cut, the usual case on a sparse map, where the 40 launches of 8 radix passes over nothing cost
 * 0.35 ms.  Same results; for callers that are not capturing a CUDA graph. */
int fhc_bh_qvalues_hostcount(const double *p, int64_t n, double T, int64_t rank_offset, double carry_in, double *q,
                             double *carry_out, int64_t *n_sorted_out, int64_t *n_ranked_host, void *workspace,
                             size_t workspace_bytes, int32_t q_prefilled, void *stream);
/* q_prefilled (here and in fhc_bh_partition_scatter): the caller has set every q[i] to 1.0 already (fhc_fill_f64, e.g. while
 * the GPU waited for the host's spline fit), so only the lines whose q is not 1.0 are written: 8 B per line less. */
int fhc_fill_f64(double *dst, int64_t n, double v, void *stream);

/* The same in two halves, for a caller that has to fetch the running max of smaller keys from other GPUs in between:
 * prepare = compaction + sort + per-tile maxima (local_max_out [dev] = max bh value of this call, 0 if none);
 * finish  = scan + scatter with floor_in = max over the key ranges below this one.  Same workspace for both calls. */
int fhc_bh_prepare(const double *p, int64_t n, double T, int64_t rank_offset, double p_cut, double *q,
                   double *local_max_out, int64_t *n_sorted_out, void *workspace, size_t workspace_bytes, void *stream);
int fhc_bh_finish(int64_t n, double T, int64_t rank_offset, double floor_in, double *q, void *workspace,
                  size_t workspace_bytes, void *stream);

/* Range partitioning of p-values over `nparts` GPUs for the global correction: part r receives the
 * rankable p-values (p != 1, not NaN) whose order-preserving key lies in [splitter[r-1], splitter[r]).
 * p_cut (every entry point below and fhc_bh_prepare): p-values >= p_cut are not ranked and get q = 1.0 directly.
 * fhc_bh_p_cut(T, rank_bound) returns the smallest safe value when no rank exceeds rank_bound (the total number of lines):
 * (p*T)/rank >= 1 there, which the reference caps at 1 and the forward running max then keeps at exactly 1.0
 * (fithic/myStats.py:36-43).  Pass INFINITY to rank everything.  fhc_bh_qvalues applies the bound internally.
 *   fhc_bh_sample_keys      keys of nsamples evenly strided p-values (UINT64_MAX where the sample is not rankable)
 *   fhc_bh_key_of           the key of one p-value (host)
 *   fhc_bh_partition_count  counts[r] = rankable p-values of part r
 *   fhc_bh_partition_scatter send[] = p-values grouped by part (cursors[r] [dev] = first slot of part r on entry),
 *                           idx[j] = line of send[j]; q[i] = 1.0 / NaN is written for the p-values that are not ranked
 *   fhc_scatter_f64         dst[idx[j]] = src[j]  (q-values coming back from the owning GPU) */
double fhc_bh_p_cut(double T, double rank_bound);

/* Tightening the cut with the ranks themselves (exact; fhc_bh_qvalues does this internally on one GPU).  Bucket j of
 * the value histogram holds the rankable p-values below p_cut0 whose high 16 bits (sign, exponent, 5 mantissa bits) are
 * j, so every bucket edge is an exact double.  If rn(edge_j * T) >= rank_offset + (p-values in buckets <= j) for a
 * non-empty bucket j, every p-value from edge_j on has (p*T)/rank >= 1: the reference caps it at 1 and its forward
 * running max stays 1.0 (fithic/myStats.py:36-43), so the smallest such edge is a valid, usually far smaller, cut.
 *   fhc_bh_cut_hist        hist[j] += count  (hist [dev] FHC_BH_CUT_BUCKETS uint64, zeroed by the caller; the
 *                          multi-GPU host sums the histograms of all ranks before looking for the cut)
 *   fhc_host_bh_cut_find   min(p_cut0, smallest closing edge) from a HOST copy of the (summed) histogram
 *   fhc_host_bh_cut_bucket the bucket of one p-value (host; for tests) */
#define FHC_BH_CUT_BUCKETS 32768
int fhc_bh_cut_hist(const double *p, int64_t n, double p_cut0, uint64_t *hist, void *stream);
double fhc_host_bh_cut_find(const uint64_t *hist, double T, double rank_offset, double p_cut0);
int32_t fhc_host_bh_cut_bucket(double p);
/* Multi-GPU: hists [dev] = every rank's fhc_bh_cut_hist histogram back to back (all-gathered, nranks x
 * FHC_BH_CUT_BUCKETS).  Writes info [dev, 8 + nranks words]: [0] the global cut (double: the rule of fhc_host_bh_cut_find
 * on the summed histogram), [1] p-values below it on all ranks, [2] on rank my_rank, [3] the largest share of one rank,
 * [8 + r] the share of rank r -- what a rank needs to size the exchange of its survivors, in ONE small read-back. */
int fhc_bh_cut_from_hists(const uint64_t *hists, int32_t nranks, int32_t my_rank, double T, double p_cut0, uint64_t *info,
                          void *stream);
/* fhc_bh_cut_hist + all-reduce over `comm` + cut + read-back in one call: info_host [pinned host, 8 words: [0] the global
 * cut (double), [1] p-values below it on all ranks, [2] on this rank] is valid on return (one stream synchronisation).
 * work [dev]: 2 * FHC_BH_CUT_BUCKETS + 8 words of scratch.  q_nan (nullable): q pre-filled with 1.0 by the caller; lines with a
 * NaN p-value get q = NaN in the same sweep (q is then final whenever nothing lies below the cut). */
int fhc_bh_dist_cut(fhc_comm *comm, const double *p, int64_t n, double T, double p_cut0, uint64_t *work, uint64_t *info_host,
                    double *q_nan, void *stream);
int fhc_bh_sample_keys(const double *p, int64_t n, int64_t nsamples, double p_cut, uint64_t *keys_out, void *stream);
uint64_t fhc_bh_key_of(double p);
int fhc_bh_partition_count(const double *p, int64_t n, const uint64_t *splitter_keys, int32_t nparts, double p_cut,
                           uint64_t *counts, void *stream);
int fhc_bh_partition_scatter(const double *p, int64_t n, const uint64_t *splitter_keys, int32_t nparts, double p_cut,
                             uint64_t *cursors, double *send, uint32_t *idx, double *q, int32_t q_prefilled, void *stream);
int fhc_scatter_f64(const double *src, const uint32_t *idx, int64_t n, double *dst, void *stream);

/* The q-values that are not exactly 1.0 (ranked lines and NaN) as (line, value) pairs, in no particular order:
 * *count [dev] receives how many there are (it may exceed `capacity`, in which case only the first `capacity` slots of
 * idx / val [dev] were written and the caller should fall back to copying q whole).  On a sparse map almost every
 * line of the reference's q column (fithic/fithic.py:1134-1163) is the constant 1.0, so a caller that needs q on the host
 * fills its array with 1.0 and copies only these pairs. */
int fhc_gather_ne_one(const double *q, int64_t n, int64_t capacity, uint32_t *idx, double *val, uint64_t *count,
                      void *stream);

/* Device radix sort of 64-bit keys with 32-bit payloads (ascending, stable), the sort inside K4, exposed for tests
 * and for the multi-GPU range-partitioned BH.  Sorted data ends in keys_out/vals_out; *_in are clobbered.
 * workspace: fhc_sort_workspace_bytes(n). */
size_t fhc_sort_workspace_bytes(int64_t n);
int fhc_sort_pairs_u64(uint64_t *keys_in, uint32_t *vals_in, uint64_t *keys_out, uint32_t *vals_out, int64_t n,
                       void *workspace, size_t workspace_bytes, void *stream);

/* ---- K5: outlier bookkeeping for pass >= 2 ---------------------------------------------------------------------
 * Per-bin decrements of makeBinsFromInteractions (fithic/fithic.py:528-548): for every line, dec[b] += outl[i] where
 * b is the bin whose [lb,ub] holds |mid1-mid2| (clamped to the last bin), exactly the forward scan of the reference
 * over its sorted outlier distances (taken for inter lines too, :1217). */
int fhc_outlier_bin_decrements(const int32_t *mid1, const int32_t *mid2, const uint8_t *outl, int64_t n,
                               const int64_t *bin_ub, int32_t nbins, uint64_t *dec, void *stream);

/* Order-independent digest of (file line, p bits, q bits) over n lines: out[0], out[1] [dev] = two 64-bit sums of hashes,
 * so the digests of disjoint shards add up (mod 2^64) to the digest of the whole file -- a run on N GPUs has computed the
 * same p and q for every line as a run on one GPU iff the digests agree.  Local lines [run_local[j], run_local[j+1]) are
 * the file lines run_global[j], run_global[j] + 1, ... (nruns <= FHC_MAX_CHR_RUNS; run_local[nruns] = n)  [dev]. */
int fhc_digest_lines(const double *p, const double *q, int64_t n, const int64_t *run_local, const int64_t *run_global,
                     int32_t nruns, uint64_t *out, void *stream);

/* ---- KR bias computation ---------------------------------------------------------------------------
 * The kernels behind fithic_b200/hickry.py, the replacement of the reference's bias generator fithic/utils/HiCKRy.py (the
 * producer of the -t file).  The contact lines are the matrix: (rows[e], cols[e], vals[e]) in file order stand for
 * coo(z, (x, y)) + its transpose (HiCKRy.py:46-50); remap[locus] is the locus' index after the sparsest rows were dropped
 * (:76-96), or -1.  All arrays [dev]; every reduction writes fhc_kr_partials() partial results that the caller adds up.
 *   fhc_kr_spmv         y = (M + M^T) x over the kept loci (y is zeroed by the callee)
 *   fhc_kr_residual     v = x * Ax; rk = 1 - v; partial sums of rk^2                       (knightRuizAlg :155-157, :208-211)
 *   fhc_kr_first        Z = rk / v; p = Z; partial sums of rk * Z                          (:178-182)
 *   fhc_kr_direction    p = first ? p : Z + beta p; xp = x * p                            (:184-185, argument of :191)
 *   fhc_kr_w            w = x * Axp + v * p; partial sums of p * w                         (:191-193)
 *   fhc_kr_ynew_minmax  partial[0..P) = min(y + alpha p), partial[P..2P) = -max(y + alpha p)   (:196-197, :206)
 *   fhc_kr_gamma        partial minima of (bound - y) / (alpha p) over alpha p < 0 (mode 0, :201-203) or over
 *                       y + alpha p > bound (mode 1, :207-209)
 *   fhc_kr_axpy         y += gamma (alpha p)                                               (:204, :210)
 *   fhc_kr_update       y += alpha p; rk -= alpha w; Z = rk / v; partial sums of rk * Z     (:213-218)
 *   fhc_kr_mul          out = a * b                                                        (x *= y, :219) */
int32_t fhc_kr_partials(void);
int fhc_kr_spmv(const int32_t *rows, const int32_t *cols, const double *vals, int64_t nnz, const int32_t *remap,
                const double *x, double *y, int64_t n, void *stream);
int fhc_kr_mul(const double *a, const double *b, double *out, int64_t n, void *stream);
int fhc_kr_residual(const double *x, const double *Ax, double *v, double *rk, int64_t n, double *partial, void *stream);
int fhc_kr_first(const double *rk, const double *v, double *Z, double *p, int64_t n, double *partial, void *stream);
int fhc_kr_direction(const double *Z, double beta, int32_t first, double *p, const double *x, double *xp, int64_t n,
                     void *stream);
int fhc_kr_w(const double *x, const double *Axp, const double *v, const double *p, double *w, int64_t n, double *partial,
             void *stream);
int fhc_kr_ynew_minmax(const double *y, double alpha, const double *p, int64_t n, double *partial, void *stream);
int fhc_kr_gamma(const double *y, double alpha, const double *p, double bound, int32_t mode, int64_t n, double *partial,
                 void *stream);
int fhc_kr_axpy(double *y, double gamma, double alpha, const double *p, int64_t n, void *stream);
int fhc_kr_update(double *y, double alpha, const double *p, double *rk, const double *w, const double *v, double *Z, int64_t n,
                  double *partial, void *stream);

/* ---- merge-filter step ------------------------------------------------------------------------------
 * fithic/utils/CombineNearbyInteraction.py (driven by fithic/utils/merge-filter.sh:22-23) groups the significant bin pairs
 * of a chromosome into connected components (-c 8: both bins differ by <= 1; -c 4: by <= 1 in total; :313-333), reports the
 * bounding box, the sum of counts and the share of box cells that hold a significant pair (:362-406), and keeps the pairs of
 * a component in (q, -count, bin1, bin2) order unless both bins lie within `Neigh` bins of a pair already kept (:585-712).
 * The reference tests all pairs of nodes (O(n^2)) and walks every box cell by cell; here the pairs are sorted once.
 *
 * Inputs [dev], one element per input LINE (intra-chromosomal lines only, in file order): chr = rank of the chromosome in
 * the output order (< 65536), b1 <= b2 the bin numbers int(mid + res/2) / res (1 <= bin < 2^24 - 1), cc the count
 * (0 <= cc < 2^31), q the q-value.  "Entry" i below is the i-th line in (chr, b1, b2) order; a repeated bin pair keeps the
 * values of its first line (:306) and its later lines get label -1.
 *   fhc_merge_components   keys[i] = chr << 48 | b1 << 24 | b2, order[i] = line of entry i, label[i] = ROOT entry of the
 *                          component (its smallest entry) or -1; indexed by root entry: size (nodes), first_line (smallest
 *                          line), box[4 i ..] = min b1, max b1, min b2, max b2, sum_cc, have = nodes of any component inside
 *                          the box.  Synchronises the stream (it reads an overflow flag back).
 *   fhc_merge_select       ranked[w] = entries grouped by component and, inside one, in the order of the reference's heap
 *                          (:611-625; sort_order 1 = `-s 1`: descending q); keep[w] = 1 for the pairs that survive the
 *                          neighbourhood rule, with the top-K % cut of :458-577 when 0 < top_pct < 100 (top_pct 100: :585-712)
 *   fhc_host_merge_*       the same per-entry code run serially on HOST arrays (tests; no GPU involved)
 * workspace: fhc_merge_workspace_bytes(n) for either call. */
size_t fhc_merge_workspace_bytes(int64_t n);
int fhc_merge_components(const int32_t *chr, const int32_t *b1, const int32_t *b2, const int64_t *cc, int64_t n, int32_t conn,
                         uint64_t *keys, uint32_t *order, int32_t *label, int32_t *size, uint32_t *first_line, int32_t *box,
                         int64_t *sum_cc, int64_t *have, void *workspace, size_t workspace_bytes, void *stream);
int fhc_merge_select(const uint64_t *keys, const uint32_t *order, const int32_t *label, const int32_t *size, const int64_t *cc,
                     const double *q, int64_t n, int32_t top_pct, int32_t neigh, int32_t sort_order, uint32_t *ranked,
                     uint8_t *keep, void *workspace, size_t workspace_bytes, void *stream);
int fhc_host_merge_components(const int32_t *chr, const int32_t *b1, const int32_t *b2, const int64_t *cc, int64_t n,
                              int32_t conn, uint64_t *keys, uint32_t *order, int32_t *label, int32_t *size,
                              uint32_t *first_line, int32_t *box, int64_t *sum_cc, int64_t *have);
int fhc_host_merge_select(const uint64_t *keys, const uint32_t *order, const int32_t *label, const int32_t *size,
                          const int64_t *cc, const double *q, int64_t n, int32_t top_pct, int32_t neigh, int32_t sort_order,
                          uint32_t *ranked, uint8_t *keep);

/* ---- text boundary (host) ---------------------------------------------------------------------------------------------
 * Native replacements of the two text loops that dominate the reference's wall time once the kernels are fast:
 * `for lines in gzip.open(contactCountsFile)` + split/int/float (fithic/fithic.py:406-417, :1017-1023) and the
 * `.significances.txt.gz` writer (:1166-1212; header :1178, row format :1202/:1212).
 *   fhc_io_read_contacts   parses `chr1 mid1 chr2 mid2 count` lines (whitespace separated, count = int(float(text)));
 *                          returns a handle (NULL on error); chromosome ids are assigned in order of first appearance
 *   fhc_io_contacts_*      size / chromosome names / copy into caller arrays; fhc_io_free releases the handle
 *   fhc_io_write_significances  formats and gzips the reported rows with `nthreads` threads (multi-member gzip, `level`
 *                          0-9); header = 0 leaves the column header out (a rank of a multi-GPU run writes the rows of its
 *                          lines, and the parts are concatenated in file order: concatenated gzip members are one gzip
 *                          file); returns the number of rows written or a negative error code */
int fhc_io_format_double(double v, int kind /* 'e' or 'f' */, char *out /* >= 400 bytes */); /* the writer's "%e" / "%f" */
void *fhc_io_read_contacts(const char *path);
int64_t fhc_io_contacts_n(void *handle);
int32_t fhc_io_contacts_nchrom(void *handle);
const char *fhc_io_contacts_chrom(void *handle, int32_t i);
int fhc_io_contacts_copy(void *handle, int32_t *mid1, int32_t *mid2, int32_t *cnt, uint32_t *chrs);
void fhc_io_free(void *handle);
int64_t fhc_io_write_significances(const char *path, const char *const *chrom_names, int32_t nchrom, const int32_t *mid1,
                                   const int32_t *mid2, const int32_t *cnt, const uint32_t *chrs, const double *p,
                                   const double *q, const double *expcc, int64_t n, int32_t mode, int64_t L, int64_t U,
                                   const double *bias, const int32_t *bias_mid, const int64_t *chr_off,
                                   int32_t nbias_chr, int32_t res, int32_t nthreads, int32_t level, int32_t header);

#ifdef __cplusplus
}
#endif
#endif /* FITHIC_B200_H */
