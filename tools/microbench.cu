// microbench.cu -- instruction-throughput probes that decide the per-step kernel's arithmetic mix
// on B200 (sm_100a): fp32<->fp64 conversions, DFMA, integer-built fp64, IEEE fp32 division.
// Prints thread-level results per clock per SM.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define CH 8

template <int MODE>
__global__ void probe(float* out, long long* cycles, float seed) {
    float f[CH];
    double d[CH];
    for (int i = 0; i < CH; ++i) { f[i] = seed + threadIdx.x * 1e-3f + i; d[i] = f[i]; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) {          // cvt.f64.f32 + cvt.rn.f32.f64 round trip
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(f[i]));
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(d[i]));
            } else if (MODE == 1) {   // DFMA
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(1.0000001));
            } else if (MODE == 2) {   // integer-built fp64 from fp32 bits (normal numbers) + DADD to consume
                unsigned x = __float_as_uint(f[i]);
                unsigned hi = (((int)x >> 3) & 0x8fffffffu) + 0x38000000u;
                unsigned lo = x << 29;
                double v = __hiloint2double(hi, lo);
                asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(v));
                f[i] = __uint_as_float(x + 1);
            } else if (MODE == 3) {   // cvt.f64.f32 only, consumed by DADD
                double v;
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(v) : "f"(f[i]));
                asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(v));
                f[i] = __uint_as_float(__float_as_uint(f[i]) + 1);
            } else if (MODE == 4) {   // IEEE fp32 division
                f[i] = __fdiv_rn(f[i], 1.0000001f);
            } else if (MODE == 5) {   // reciprocal-multiply + Markstein correction
                float q = f[i] * 0.9999999f;
                float r = fmaf(-q, 1.0000001f, f[i]);
                f[i] = fmaf(r, 0.9999999f, q);
            } else if (MODE == 6) {   // cvt.rn.f32.f64 only (fed by DADD)
                asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(1.0));
                float v;
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(v) : "d"(d[i]));
                f[i] += v;
            }
        }
    }
    long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < CH; ++i) acc += f[i] + (float)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double ops_per_inner) {
    int blocks = 148 * 4, threads = 512;
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    probe<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
    probe<MODE><<<blocks, threads>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    long long h[148 * 4];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    // 4 blocks of 512 threads resident per SM run concurrently
    double per_clk_sm = ops_per_inner * ITERS * CH * threads * 4.0 / avg;
    printf("%-44s %8.1f per clk per SM  (avg %.0f cycles/block)  %s\n", name, per_clk_sm, avg, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("cvt f32->f64 + cvt f64->f32 (pairs)", 1);
    run<3>("cvt f32->f64 (+DADD)", 1);
    run<6>("cvt f64->f32 (+DADD+FADD)", 1);
    run<1>("DFMA", 1);
    run<2>("int-built f64 (+DADD)", 1);
    run<4>("__fdiv_rn", 1);
    run<5>("rcp-mul + Markstein (3 FMA-pipe ops)", 1);
    return 0;
}
