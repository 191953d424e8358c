// fl_stats.cu -- dataset statistics: per-channel (n, mean, M2) of states and diffs over unmasked
// pixels (max/compute_ds_stats.py:20-34,52-62), reduced with warp shuffles and merged in a fixed
// order (Chan et al.) so every run and every rank produces bit-identical aggregates.
//
// Each thread owns one pixel position (l, k) and walks the frames of its chunk once, keeping the
// previous frame's three channel values in registers, so every state element is read from HBM
// once.  Accumulation is in fp64 on centred sums; diffs are formed in fp32 exactly as
// simple_dataloader.py:93 does.
#include "fl_common.cuh"

namespace {

constexpr int NACC = 6;          // state ch0..2, diff ch0..2
constexpr int CHUNK_T = 32;      // frames per CTA chunk

struct Agg { double n, mean, m2; };

__device__ __forceinline__ Agg chan(Agg a, Agg b) {
    if (b.n == 0.0) return a;
    if (a.n == 0.0) return b;
    double n = a.n + b.n, d = b.mean - a.mean;
    Agg r;
    r.n = n;
    r.mean = a.mean + d * (b.n / n);
    r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / n);
    return r;
}

__device__ __forceinline__ Agg warp_merge(Agg a) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {   // butterfly in a fixed order: lane i merges (lower, upper)
        Agg b;
        b.n = __shfl_xor_sync(0xffffffffu, a.n, o);
        b.mean = __shfl_xor_sync(0xffffffffu, a.mean, o);
        b.m2 = __shfl_xor_sync(0xffffffffu, a.m2, o);
        bool upper = (threadIdx.x & o) != 0;
        a = upper ? chan(b, a) : chan(a, b);
    }
    return a;
}

// grid: x = pixel tiles (256 pixels each over L*ppx), y = frame chunks
__global__ void __launch_bounds__(256) k_stats_partial(const float* __restrict__ states, const uint8_t* __restrict__ mask,
                                                       int T, int L, int ppx, double* __restrict__ partials) {
    const long npix = (long)L * ppx;
    const long pix = (long)blockIdx.x * 256 + threadIdx.x;
    const int t_begin = blockIdx.y * CHUNK_T;              // pairs (t, t+1) for t in [t_begin, t_end)
    const int t_end = min(T - 1, t_begin + CHUNK_T);
    // per-thread running sums around a per-thread shift (first accepted sample), fp64
    double cnt[2] = {0.0, 0.0};
    double shift[NACC], s1[NACC], s2[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) { shift[a] = 0.0; s1[a] = 0.0; s2[a] = 0.0; }
    if (pix < npix && t_begin < t_end) {
        const long l = pix / ppx, k = pix - l * ppx;
        auto addr = [&](int t, int c) { return (((long)t * L + l) * 3 + c) * ppx + k; };
        float prev[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) prev[c] = __ldg(states + addr(t_begin, c));
        bool first = true;
        for (int t = t_begin; t < t_end; ++t) {
            float cur[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) cur[c] = __ldg(states + addr(t + 1, c));
            if (!mask[((long)(t + 1) * L + l) * ppx + k]) {   // masks[1:] (simple_dataloader.py:100)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    double sv = (double)prev[c], dv = (double)__fsub_rn(cur[c], prev[c]);
                    if (first) { shift[c] = sv; shift[3 + c] = dv; }
                    double a = sv - shift[c], b = dv - shift[3 + c];
                    s1[c] += a; s2[c] += a * a;
                    s1[3 + c] += b; s2[3 + c] += b * b;
                }
                first = false;
                cnt[0] += 1.0;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) prev[c] = cur[c];
        }
    }
    __shared__ Agg sm[NACC][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        Agg g;
        g.n = cnt[0];
        g.mean = g.n > 0.0 ? shift[a] + s1[a] / g.n : 0.0;
        g.m2 = g.n > 0.0 ? s2[a] - s1[a] * s1[a] / g.n : 0.0;
        g = warp_merge(g);
        if (lane == 0) sm[a][warp] = g;
    }
    __syncthreads();
    if (threadIdx.x < NACC) {
        Agg g = sm[threadIdx.x][0];
        for (int w = 1; w < 8; ++w) g = chan(g, sm[threadIdx.x][w]);
        double* o = partials + ((long)(blockIdx.y * gridDim.x + blockIdx.x) * NACC + threadIdx.x) * 3;
        o[0] = g.n; o[1] = g.mean; o[2] = g.m2;
    }
}

// fixed-shape pairwise tree over the partials: deterministic, one warp per accumulator
__global__ void k_stats_merge(const double* __restrict__ parts, int n_parts, double* __restrict__ out) {
    const int a = threadIdx.x >> 5, lane = threadIdx.x & 31;   // blockDim = NACC*32
    Agg g{0.0, 0.0, 0.0};
    const int per = (n_parts + 31) / 32;                       // contiguous slice per lane, in order
    for (int i = lane * per; i < min(n_parts, (lane + 1) * per); ++i) {
        const double* p = parts + ((long)i * NACC + a) * 3;
        g = chan(g, Agg{p[0], p[1], p[2]});
    }
    g = warp_merge(g);
    if (lane == 0) { out[a * 3] = g.n; out[a * 3 + 1] = g.mean; out[a * 3 + 2] = g.m2; }
}

constexpr size_t MAX_PARTS = 1 << 16;

}  // namespace

extern "C" size_t fl_stats_workspace_bytes(void) { return MAX_PARTS * NACC * 3 * sizeof(double); }

extern "C" int fl_ds_stats(const float* d_states, const uint8_t* d_mask, int T, int L, int px, int py, double* d_agg,
                           void* d_workspace, size_t workspace_bytes, void* stream) {
    FL_REQUIRE(d_states && d_mask && d_agg && d_workspace, FL_E_ARG, "fl_ds_stats: null pointer");
    FL_REQUIRE(T >= 2 && L > 0 && px > 0 && py > 0, FL_E_ARG, "fl_ds_stats: need T >= 2 and positive sizes");
    const int ppx = px * py;
    long npix = (long)L * ppx;
    unsigned gx = (unsigned)((npix + 255) / 256), gy = (unsigned)((T - 1 + CHUNK_T - 1) / CHUNK_T);
    size_t parts = (size_t)gx * gy;
    FL_REQUIRE(parts <= MAX_PARTS && workspace_bytes >= parts * NACC * 3 * sizeof(double), FL_E_WORKSPACE,
               "fl_ds_stats: workspace too small for %zu partial aggregates", parts);
    cudaStream_t st = (cudaStream_t)stream;
    k_stats_partial<<<dim3(gx, gy), 256, 0, st>>>(d_states, d_mask, T, L, ppx, (double*)d_workspace);
    FL_LAUNCH_CHECK();
    k_stats_merge<<<1, NACC * 32, 0, st>>>((const double*)d_workspace, (int)parts, d_agg);
    FL_LAUNCH_CHECK();
    return FL_OK;
}

extern "C" int fl_stats_merge(const double* d_parts, int n_parts, double* d_out, void* stream) {
    FL_REQUIRE(d_parts && d_out && n_parts > 0, FL_E_ARG, "fl_stats_merge: bad arguments");
    k_stats_merge<<<1, NACC * 32, 0, (cudaStream_t)stream>>>(d_parts, n_parts, d_out);
    FL_LAUNCH_CHECK();
    return FL_OK;
}
