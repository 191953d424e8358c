// fl_common.cuh -- shared helpers of libfluidgrid (sm_100a only; no CPU fallback anywhere).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/fluidgrid.h"

#define FL_SM_COUNT 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void fl_set_error(const char* fmt, ...);

#define FL_REQUIRE(cond, code, ...)          \
    do {                                     \
        if (!(cond)) {                       \
            fl_set_error(__VA_ARGS__);       \
            return (code);                   \
        }                                    \
    } while (0)

#define FL_CUDA(expr)                                                                  \
    do {                                                                               \
        cudaError_t e__ = (expr);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            fl_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return (int)e__;                                                           \
        }                                                                              \
    } while (0)

#define FL_LAUNCH_CHECK()                                                              \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            fl_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return (int)e__;                                                           \
        }                                                                              \
    } while (0)

// cudaFuncSetAttribute is per device: one of these per call site remembers which devices already have it
// (setting it twice is harmless, so a race between two threads on first use only costs a repeated call)
struct FlOncePerDevice {
    unsigned long long done[4] = {0, 0, 0, 0};     // up to 256 devices
    bool first_use() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 256) return true;
        const unsigned long long bit = 1ull << (d & 63);
        if (done[d >> 6] & bit) return false;
        done[d >> 6] |= bit;
        return true;
    }
};

static inline size_t fl_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// streaming (read-once / write-once) 128-bit accesses that do not allocate in L1
__device__ __forceinline__ float4 fl_ldg_stream4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void fl_stg_stream4(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void fl_stg_stream2(float2* p, float2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" :: "l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void fl_stg_stream1(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}
