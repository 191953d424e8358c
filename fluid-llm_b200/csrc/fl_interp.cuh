// fl_interp.cuh -- pieces shared by the per-step kernels (fl_interp.cu: staged / gather, fl_tiled.cu: tiled):
// the normalisation constants and the correctly rounded (x - mean) / std without a division.
#pragma once
#include "fl_geom.cuh"

namespace fli {

struct StagedConst {
    float mean[3], stdv[3], rcp[3];
    int fast_div;        // 1: Markstein division valid for these constants
};

// (x - mean) / std, correctly rounded: q = d*r, then one Markstein step with the exact remainder
__device__ __forceinline__ float norm_fast(float x, float mean, float stdv, float rcp) {
    float d = __fsub_rn(x, mean);
    float q = __fmul_rn(d, rcp);
    float e = __fmaf_rn(-q, stdv, d);
    return __fmaf_rn(e, rcp, q);
}
// two pixels at once on the packed fp32 pipe (FADD2 / FMUL2 / FFMA2): same roundings as norm_fast
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    return ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(a);
}
__device__ __forceinline__ void norm_fast2(float& x0, float& x1, unsigned long long nm, unsigned long long ns, unsigned long long rc) {
    const unsigned long long x = pack2(x0, x1);
    unsigned long long d, q, e, y;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x), "l"(nm));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(q) : "l"(d), "l"(rc));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(e) : "l"(q), "l"(ns), "l"(d));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(y) : "l"(e), "l"(rc), "l"(q));
    x0 = __uint_as_float((unsigned)y);
    x1 = __uint_as_float((unsigned)(y >> 32));
}

// host: is (x - mean) / std safe for the reciprocal + Markstein path?
bool fast_div_ok(const float* mean, const float* stdv);

// fl_tiled.cu: the tiled per-step kernel; returns FL_OK, an error, or 1 ("not applicable": the caller picks another kernel)
int launch_tiled(const FlTraj* d_trajs, const FlTraj* h_trajs, int n_traj, int max_frames, int n_patches, int px, int py,
                 const StagedConst& sc, unsigned flags, cudaStream_t st);

// fl_ring.cu: the whole-frame ring kernel for meshes whose frames fit shared memory a few times over; same return convention
int launch_ring(const FlTraj* d_trajs, const FlTraj* h_trajs, int n_traj, int max_frames, int n_patches, int px, int py,
                const StagedConst& sc, unsigned flags, cudaStream_t st);

}  // namespace fli
