// membench.cu -- what write bandwidth does B200 give for (a) a flat streaming fill, (b) the per-step
// kernel's output pattern (one CTA walks TF frames of one trajectory, each warp store = 512 B, three
// channel segments per quad, frames 12*P bytes apart), (c) a flat copy.  Sizes mirror bench.py airfoil.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void stg4(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__global__ void fill_flat(float4* out, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += step) stg4(out + i, make_float4(1, 2, 3, 4));
}
__global__ void copy_flat(const float4* in, float4* out, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, step = (size_t)gridDim.x * blockDim.x;
    for (; i < n4; i += step) stg4(out + i, __ldg(in + i));
}
// pattern of k_interp_patchify_staged: item = TF frames; thread walks quads, inner loop over frames
__global__ void fill_pattern(float* out, int n_items, int TF, int L, int ppx, int frames_outer) {
    const size_t frame_out = (size_t)L * 3 * ppx;
    const int nquads = L * ppx / 4;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        float* base = out + (size_t)item * TF * frame_out;
        if (frames_outer) {
            for (int f = 0; f < TF; ++f)
                for (int q = threadIdx.x; q < nquads; q += blockDim.x) {
                    int o = 4 * q, l = o / ppx, k = o - l * ppx;
                    float* dst = base + (size_t)f * frame_out + (size_t)l * 3 * ppx + k;
                    for (int c = 0; c < 3; ++c) stg4((float4*)(dst + c * ppx), make_float4(1, 2, 3, 4));
                }
        } else {
            for (int q = threadIdx.x; q < nquads; q += blockDim.x) {
                int o = 4 * q, l = o / ppx, k = o - l * ppx;
                float* dst = base + (size_t)l * 3 * ppx + k;
                for (int f = 0; f < TF; ++f, dst += frame_out)
                    for (int c = 0; c < 3; ++c) stg4((float4*)(dst + c * ppx), make_float4(1, 2, 3, 4));
            }
        }
    }
}

template <typename F> float timeit(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}

int main() {
    const int L = 91, ppx = 256, TF = 6, frames = 9600;
    const size_t frame_out = (size_t)L * 3 * ppx, total = frame_out * frames;   // floats
    float *out, *in;
    cudaMalloc(&out, total * 4); cudaMalloc(&in, total * 4);
    cudaMemset(in, 0, total * 4);
    double gb = total * 4 / 1e9;
    float ms = timeit([&] { fill_flat<<<148 * 8, 512>>>((float4*)out, total / 4); }, 10);
    printf("fill flat                : %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { copy_flat<<<148 * 8, 512>>>((const float4*)in, (float4*)out, total / 4); }, 10);
    printf("copy flat (r+w)          : %.3f ms  %.0f GB/s\n", ms, 2 * gb / ms * 1e3);
    for (int thr : {512, 1024}) for (int fo : {0, 1}) {
        ms = timeit([&] { fill_pattern<<<148, thr>>>(out, frames / TF, TF, L, ppx, fo); }, 10);
        printf("fill pattern thr=%4d %s: %.3f ms  %.0f GB/s\n", thr, fo ? "frames-outer" : "quad-outer  ", ms, gb / ms * 1e3);
    }
    ms = timeit([&] { fill_pattern<<<296, 512>>>(out, frames / 3, 3, L, ppx, 0); }, 10);
    printf("fill pattern 2x512 TF=3  : %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
