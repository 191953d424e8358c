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
constexpr int DEPTH = 2;         // frames ahead of the one being accumulated

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

// Per-CTA tail shared by the two partial kernels: every lane re-bases its sums to one shift of its warp, a fixed butterfly adds
// the lanes, the eight warp aggregates are merged with Chan's formula and written as partial (blockIdx.y, blockIdx.x).
__device__ __forceinline__ void stats_tail(const double (&cnt)[2], const float (&shift_f)[NACC], const double (&s1)[NACC],
                                           const double (&s2)[NACC], double* __restrict__ partials) {
    double shift[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) shift[a] = (double)shift_f[a];
    // Per-warp reduction without a division per lane: every lane re-bases its sums to one shift of the warp (the shift of
    // the first lane that accepted a sample), then the three sums are added across the lanes in a fixed butterfly order.
    //   sum(x - K) = s1 + n (shift - K),   sum (x - K)^2 = s2 + 2 (shift - K) s1 + n (shift - K)^2
    // One lane per warp turns the totals into (n, mean, M2); the eight warp aggregates are merged with Chan's formula.
    __shared__ Agg sm[NACC][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned have = __ballot_sync(0xffffffffu, cnt[0] > 0.0);
    const int src = have ? __ffs(have) - 1 : 0;
    double n_w = cnt[0];
#pragma unroll
    for (int o = 16; o; o >>= 1) n_w += __shfl_xor_sync(0xffffffffu, n_w, o);
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        const double K = __shfl_sync(0xffffffffu, shift[a], src);
        const double d = shift[a] - K;
        double t1 = cnt[0] > 0.0 ? s1[a] + cnt[0] * d : 0.0;
        double t2 = cnt[0] > 0.0 ? s2[a] + 2.0 * d * s1[a] + cnt[0] * d * d : 0.0;
#pragma unroll
        for (int o = 16; o; o >>= 1) { t1 += __shfl_xor_sync(0xffffffffu, t1, o); t2 += __shfl_xor_sync(0xffffffffu, t2, o); }
        if (lane == 0) {
            Agg g;
            g.n = n_w;
            g.mean = n_w > 0.0 ? K + t1 / n_w : 0.0;
            g.m2 = n_w > 0.0 ? fmax(t2 - t1 * t1 / n_w, 0.0) : 0.0;
            sm[a][warp] = g;
        }
    }
    __syncthreads();
    if (threadIdx.x < NACC) {
        Agg g = sm[threadIdx.x][0];
        for (int w = 1; w < 8; ++w) g = chan(g, sm[threadIdx.x][w]);
        double* o = partials + ((long)(blockIdx.y * gridDim.x + blockIdx.x) * NACC + threadIdx.x) * 3;
        o[0] = g.n; o[1] = g.mean; o[2] = g.m2;
    }
}

// grid: x = pixel tiles (256 pixels each over L*ppx), y = frame chunks.  The loop is bound by memory latency, not by
// arithmetic: three CTAs per SM (the shifts are kept as the fp32 numbers they are) and the loads of DEPTH frames are in
// flight while the current pair is accumulated.
__global__ void __launch_bounds__(256, 4) k_stats_partial(const float* __restrict__ states, const uint8_t* __restrict__ mask,
                                                          int T, int L, int ppx, int chunk_t, double* __restrict__ partials) {
    const long npix = (long)L * ppx;
    const long pix = (long)blockIdx.x * 256 + threadIdx.x;
    const int t_begin = blockIdx.y * chunk_t;              // pairs (t, t+1) for t in [t_begin, t_end)
    const int t_end = min(T - 1, t_begin + chunk_t);
    // per-thread running sums in fp64; the state sums are taken around a per-thread shift (the first frame's value)
    double cnt[2] = {0.0, 0.0};
    float shift_f[NACC];
    double s1[NACC], s2[NACC], ks[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int a = 0; a < NACC; ++a) { shift_f[a] = 0.f; s1[a] = 0.0; s2[a] = 0.0; }
    if (pix < npix && t_begin < t_end) {
        const long l = pix / ppx, k = pix - l * ppx;
        const size_t fstride = (size_t)L * 3 * ppx, mstride = (size_t)L * ppx;
        const float* sp = states + ((size_t)t_begin * L + l) * 3 * ppx + k;          // channel c at sp[c * ppx]
        const uint8_t* mp = mask + ((size_t)t_begin * L + l) * ppx + k;
        // DEPTH + 1 frame buffers rotate through the roles (previous, current, DEPTH - 1 frames in flight) with the loop
        // unrolled DEPTH + 1 times, so the buffer indices are compile-time (registers) and no move waits for a load:
        // (a deeper ring was measured: no gain, the loop is bound by issue slots and the conversion unit, not by latency)
        float fb[DEPTH + 1][3];
        uint8_t mb[DEPTH + 1];
#pragma unroll
        for (int i = 0; i <= DEPTH; ++i) { fb[i][0] = fb[i][1] = fb[i][2] = 0.f; mb[i] = 1; }
#pragma unroll
        for (int i = 0; i < DEPTH; ++i) {                  // frames t_begin .. t_begin + DEPTH - 1
            if (t_begin + i <= t_end) {
#pragma unroll
                for (int c = 0; c < 3; ++c) fb[i][c] = __ldg(sp + (size_t)i * fstride + (size_t)c * ppx);
                mb[i] = __ldg(mp + (size_t)i * mstride);
            }
        }
        // shift of the state sums: this thread's first frame (masked or not, it is a finite number near the data);
        // the differences are centred on zero and need none
#pragma unroll
        for (int c = 0; c < 3; ++c) { ks[c] = (double)fb[0][c]; shift_f[c] = fb[0][c]; }
        int t = t_begin;
        while (t < t_end) {
#pragma unroll
            for (int s = 0; s <= DEPTH; ++s) {
                if (t < t_end) {
                    const int ld = (s + DEPTH) % (DEPTH + 1), cu = (s + 1) % (DEPTH + 1);
                    if (t + DEPTH <= t_end) {               // frame t + DEPTH goes where frame t - 1 was
#pragma unroll
                        for (int c = 0; c < 3; ++c) fb[ld][c] = __ldg(sp + (size_t)DEPTH * fstride + (size_t)c * ppx);
                        mb[ld] = __ldg(mp + (size_t)DEPTH * mstride);
                    }
                    if (!mb[cu]) {                          // masks[1:] (simple_dataloader.py:100)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const double a = (double)fb[s][c] - ks[c], b = (double)__fsub_rn(fb[cu][c], fb[s][c]);
                            s1[c] += a; s2[c] += a * a;
                            s1[3 + c] += b; s2[3 + c] += b * b;
                        }
                        cnt[0] += 1.0;
                    }
                    sp += fstride; mp += mstride;
                    ++t;
                }
            }
        }
    }
    stats_tail(cnt, shift_f, s1, s2, partials);
}

// The same with four consecutive pixels per thread (128-bit loads of the three channels, one 32-bit load of the four mask bytes):
// four times the bytes in flight per thread for the same loop, which is what the scalar form lacked (ncu: long_scoreboard 9 per
// issue, 1 KB pieces 184 KB apart).  A thread's four pixels share its accumulators.  grid: x = tiles of 1024 pixels, y = frame chunks.
template <int DEPTH>
__global__ void __launch_bounds__(256, 2) k_stats_partial4(const float* __restrict__ states, const uint8_t* __restrict__ mask,
                                                           int T, int L, int ppx, int chunk_t, double* __restrict__ partials) {
    const long npix = (long)L * ppx;
    const long pix = ((long)blockIdx.x * 256 + threadIdx.x) * 4;
    const int t_begin = blockIdx.y * chunk_t;
    const int t_end = min(T - 1, t_begin + chunk_t);
    double cnt[2] = {0.0, 0.0};
    float shift_f[NACC];
    double s1[NACC], s2[NACC], ks[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int a = 0; a < NACC; ++a) { shift_f[a] = 0.f; s1[a] = 0.0; s2[a] = 0.0; }
    if (pix < npix && t_begin < t_end) {
        const long l = pix / ppx, k = pix - l * ppx;           // ppx is a multiple of 4: the four pixels share the patch
        const size_t fstride = (size_t)L * 3 * ppx, mstride = (size_t)L * ppx;
        const float* sp = states + ((size_t)t_begin * L + l) * 3 * ppx + k;
        const uint8_t* mp = mask + ((size_t)t_begin * L + l) * ppx + k;
        float4 fb[DEPTH + 1][3];
        unsigned mb[DEPTH + 1];
#pragma unroll
        for (int i = 0; i <= DEPTH; ++i) { fb[i][0] = fb[i][1] = fb[i][2] = make_float4(0.f, 0.f, 0.f, 0.f); mb[i] = 0x01010101u; }
#pragma unroll
        for (int i = 0; i < DEPTH; ++i) {
            if (t_begin + i <= t_end) {
#pragma unroll
                for (int c = 0; c < 3; ++c) fb[i][c] = fl_ldg_stream4((const float4*)(sp + (size_t)i * fstride + (size_t)c * ppx));
                mb[i] = __ldg((const unsigned*)(mp + (size_t)i * mstride));
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) { ks[c] = (double)fb[0][c].x; shift_f[c] = fb[0][c].x; }
        int t = t_begin;
        while (t < t_end) {
#pragma unroll
            for (int s = 0; s <= DEPTH; ++s) {
                if (t < t_end) {
                    const int ld = (s + DEPTH) % (DEPTH + 1), cu = (s + 1) % (DEPTH + 1);
                    if (t + DEPTH <= t_end) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) fb[ld][c] = fl_ldg_stream4((const float4*)(sp + (size_t)DEPTH * fstride + (size_t)c * ppx));
                        mb[ld] = __ldg((const unsigned*)(mp + (size_t)DEPTH * mstride));
                    }
                    const unsigned m = mb[cu];                  // masks[1:] (simple_dataloader.py:100), one byte per pixel
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (!((m >> (8 * j)) & 0xffu)) {
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                const float x0 = j == 0 ? fb[s][c].x : j == 1 ? fb[s][c].y : j == 2 ? fb[s][c].z : fb[s][c].w;
                                const float x1 = j == 0 ? fb[cu][c].x : j == 1 ? fb[cu][c].y : j == 2 ? fb[cu][c].z : fb[cu][c].w;
                                const double a = (double)x0 - ks[c], b = (double)__fsub_rn(x1, x0);
                                s1[c] += a; s2[c] += a * a;
                                s1[3 + c] += b; s2[3 + c] += b * b;
                            }
                            cnt[0] += 1.0;
                        }
                    }
                    sp += fstride; mp += mstride;
                    ++t;
                }
            }
        }
    }
    stats_tail(cnt, shift_f, s1, s2, partials);
}

// fixed-shape tree over the partials: deterministic.  MERGE_WARPS warps per accumulator; every lane folds a contiguous
// slice in order, a butterfly merges the lanes, and the warps' results are folded in warp order.
constexpr int MERGE_WARPS = 5;
__global__ void __launch_bounds__(NACC * MERGE_WARPS * 32) k_stats_merge(const double* __restrict__ parts, int n_parts,
                                                                         double* __restrict__ out) {
    __shared__ Agg sm[NACC][MERGE_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a = warp / MERGE_WARPS, sub = warp - a * MERGE_WARPS;
    const int slot = sub * 32 + lane, slots = MERGE_WARPS * 32;
    Agg g{0.0, 0.0, 0.0};
    const int per = (n_parts + slots - 1) / slots;             // contiguous slice per lane, in order
    for (int i = slot * per; i < min(n_parts, (slot + 1) * per); ++i) {
        const double* p = parts + ((long)i * NACC + a) * 3;
        g = chan(g, Agg{p[0], p[1], p[2]});
    }
    g = warp_merge(g);
    if (lane == 0) sm[a][sub] = g;
    __syncthreads();
    if (threadIdx.x < NACC) {
        Agg r = sm[threadIdx.x][0];
        for (int w = 1; w < MERGE_WARPS; ++w) r = chan(r, sm[threadIdx.x][w]);
        out[threadIdx.x * 3] = r.n; out[threadIdx.x * 3 + 1] = r.mean; out[threadIdx.x * 3 + 2] = r.m2;
    }
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
    // four pixels per thread where the layout allows 128-bit loads
    const bool wide = ppx % 4 == 0 && (uintptr_t)d_states % 16 == 0 && (uintptr_t)d_mask % 4 == 0;
    const unsigned gx = (unsigned)((npix + (wide ? 1023 : 255)) / (wide ? 1024 : 256));
    // frames per CTA: one wave of CTAs over the resident slots (148 SMs x 2 or 4 CTAs) where the data allows it -- 1.6 waves
    // run at the speed of 2, and every CTA ends with the same reduction tail, so fewer and longer CTAs are cheaper
    const long slots = (long)FL_SM_COUNT * (wide ? 2 : 4);
    const long gy_fit = slots / gx > 0 ? slots / gx : 1;
    int chunk_t = (int)((T - 1 + gy_fit - 1) / gy_fit);
    if (chunk_t < 16) chunk_t = 16;
    const unsigned gy = (unsigned)((T - 1 + chunk_t - 1) / chunk_t);
    size_t parts = (size_t)gx * gy;
    FL_REQUIRE(parts <= MAX_PARTS && workspace_bytes >= parts * NACC * 3 * sizeof(double), FL_E_WORKSPACE,
               "fl_ds_stats: workspace too small for %zu partial aggregates", parts);
    cudaStream_t st = (cudaStream_t)stream;
    if (wide) k_stats_partial4<2><<<dim3(gx, gy), 256, 0, st>>>(d_states, d_mask, T, L, ppx, chunk_t, (double*)d_workspace);
    else k_stats_partial<<<dim3(gx, gy), 256, 0, st>>>(d_states, d_mask, T, L, ppx, chunk_t, (double*)d_workspace);
    FL_LAUNCH_CHECK();
    k_stats_merge<<<1, NACC * MERGE_WARPS * 32, 0, st>>>((const double*)d_workspace, (int)parts, d_agg);
    FL_LAUNCH_CHECK();
    return FL_OK;
}

extern "C" int fl_stats_merge(const double* d_parts, int n_parts, double* d_out, void* stream) {
    FL_REQUIRE(d_parts && d_out && n_parts > 0, FL_E_ARG, "fl_stats_merge: bad arguments");
    k_stats_merge<<<1, NACC * MERGE_WARPS * 32, 0, (cudaStream_t)stream>>>(d_parts, n_parts, d_out);
    FL_LAUNCH_CHECK();
    return FL_OK;
}
