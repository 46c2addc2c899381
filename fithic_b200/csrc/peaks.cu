// Measured denominators for the rooflines bench.py reports: the FP64 FMA rate of this GPU (K3 is bound by FP64 latency and
// instruction issue, not by HBM; MEASURED_PEAKS.json carries no FP64 figure) and a plain streaming-read bandwidth.
#define FHC_PROFILE_STREAM st
#include "common.cuh"

namespace fhc {

// 16 independent FMA chains per thread: enough to cover the FP64 pipe's latency with 8 warps per scheduler
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    double y0 = x0 * 0.5, y1 = x1 * 0.5, y2 = x2 * 0.5, y3 = x3 * 0.5, y4 = x4 * 0.5, y5 = x5 * 0.5, y6 = x6 * 0.5, y7 = x7 * 0.5;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        y0 = fma(y0, a, b); y1 = fma(y1, a, b); y2 = fma(y2, a, b); y3 = fma(y3, a, b);
        y4 = fma(y4, a, b); y5 = fma(y5, a, b); y6 = fma(y6, a, b); y7 = fma(y7, a, b);
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7)) + ((y0 + y1) + (y2 + y3)) + ((y4 + y5) + (y6 + y7));
    if (s == 12345.678) out[0] = s;  // never true: keeps the chains alive
}

}  // namespace fhc

// FP64 FMA throughput in TFLOP/s (2 flops per FMA).  seconds <= 0: one burst of ~10 ms; otherwise launches back to back for
// about that long (sustained clocks).  scratch [dev]: one double.  Synchronises the stream.
extern "C" int fhc_peak_fp64(double seconds, double *scratch, double *tflops_out, void *stream) {
    using namespace fhc;
    FHC_REQUIRE(scratch && tflops_out, FHC_E_INVALID, "fhc_peak_fp64: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int blocks = kNumSMs * 8, threads = 256, iters = 4096;
    const double flops_per_launch = 2.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
    cudaEvent_t e0, e1;
    FHC_CUDA(cudaEventCreate(&e0));
    FHC_CUDA(cudaEventCreate(&e1));
    fp64_peak_kernel<<<blocks, threads, 0, st>>>(scratch, iters, 1.0000001, 1e-9);  // warm-up
    FHC_LAUNCH_CHECK("fp64_peak_kernel");
    FHC_CUDA(cudaStreamSynchronize(st));
    double best = 0.0, total_ms = 0.0, total_flops = 0.0;
    const int rounds = seconds > 0 ? 1000000 : 5;
    for (int r = 0; r < rounds; ++r) {
        FHC_CUDA(cudaEventRecord(e0, st));
        for (int k = 0; k < 4; ++k) {
            fp64_peak_kernel<<<blocks, threads, 0, st>>>(scratch, iters, 1.0000001, 1e-9);
            FHC_LAUNCH_CHECK("fp64_peak_kernel");
        }
        FHC_CUDA(cudaEventRecord(e1, st));
        FHC_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        FHC_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 4.0 * flops_per_launch / ((double)ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
        total_ms += ms;
        total_flops += 4.0 * flops_per_launch;
        if (seconds > 0 && total_ms >= seconds * 1e3) break;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops_out = seconds > 0 ? total_flops / (total_ms * 1e-3) / 1e12 : best;
    return FHC_OK;
}
