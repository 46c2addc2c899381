// Shared helpers for the fithic_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fithic_b200.h"

namespace fhc {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
// optional per-kernel timing (fhc_profile_*): one CUDA event after every launch on the launching stream
void profile_mark(const char *name, cudaStream_t st);
extern bool g_profile_on;

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

#define FHC_REQUIRE(cond, code, ...)  \
    do {                              \
        if (!(cond)) {                \
            fhc::set_error(__VA_ARGS__); \
            return (code);            \
        }                             \
    } while (0)

#define FHC_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            fhc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return FHC_E_CUDA;                                                                \
        }                                                                                     \
    } while (0)

#define FHC_LAUNCH_CHECK(name)                                                                \
    do {                                                                                      \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess) {                                                             \
            fhc::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));         \
            return FHC_E_CUDA;                                                                \
        }                                                                                     \
        fhc::count_launch();                                                                  \
        if (fhc::g_profile_on) fhc::profile_mark(name, FHC_PROFILE_STREAM);                   \
    } while (0)

// entry points define this to the stream they launch on before using FHC_LAUNCH_CHECK
#define FHC_PROFILE_ENTRY(st) \
    do {                      \
        if (fhc::g_profile_on) fhc::profile_mark(nullptr, st); \
    } while (0)

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 128-bit read-only streaming load (data touched once per kernel: keep it out of L1)
__device__ __forceinline__ int4 ldg_stream(const int4 *p) {
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

__device__ __forceinline__ int2 ldg_stream2(const int2 *p) {
    int2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

__device__ __forceinline__ double2 ldg_stream_d2(const double2 *p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

}  // namespace fhc
