// fl_interp.cu -- per-step hot path: gather -> fp64 FMA -> fp32 -> mask -> normalise -> patchify.
//
// Replaces, per frame, three mesh_utils.to_grid calls (src/dataloader/mesh_utils.py:82-91: a
// full calculate_plane_coefficients pass over every triangle plus a NumPy gather, per channel),
// _pad, the mask concat, F.unfold, the permute and _normalize
// (src/dataloader/simple_dataloader.py:104-152,166-216; airfoil_ds.py:71-139,216-244).
//
// Numerics: the reference evaluates the fp64 plane a*x + b*y + c and rounds once
// to fp32.  Here the same point value is the fp64 barycentric sum w0*z0 + w1*z1 + w2*z2 (weights
// fp64, from the static table), rounded once to fp32, then (v - mean) and / std in fp32 with
// IEEE division, i.e. the same two roundings as simple_dataloader.py:213-214.
#include "fl_common.cuh"

namespace {

struct NormConst { float mean[3]; float stdv[3]; };

// one output pixel, three channels, fp64 barycentric sum rounded once to fp32
__device__ __forceinline__ void interp3(const float* __restrict__ vel, const float* __restrict__ prs, FlCellIdx id,
                                        double w0, double w1, double w2, float out[3]) {
    const float2 a0 = __ldg((const float2*)vel + id.v0);
    const float2 a1 = __ldg((const float2*)vel + id.v1);
    const float2 a2 = __ldg((const float2*)vel + id.v2);
    const float p0 = __ldg(prs + id.v0), p1 = __ldg(prs + id.v1), p2 = __ldg(prs + id.v2);
    out[0] = (float)fma(w2, (double)a2.x, fma(w1, (double)a1.x, w0 * (double)a0.x));
    out[1] = (float)fma(w2, (double)a2.y, fma(w1, (double)a1.y, w0 * (double)a0.y));
    out[2] = (float)fma(w2, (double)p2, fma(w1, (double)p1, w0 * (double)p0));
}

__device__ __forceinline__ bool finite_f(float v) { return (__float_as_uint(v) & 0x7f800000u) != 0x7f800000u; }

// Generic kernel: one CTA = one patch (px*py threads) x a chunk of frames of one trajectory.
// Each thread keeps its cell record in registers and walks the frames; node values are gathered
// straight from global memory (L1/L2 hits: a patch touches ~100 nodes per frame).
// Works for any mesh size; the staged kernel below is the fast path for meshes that fit in smem.
__global__ void __launch_bounds__(256) k_interp_patchify_gather(const FlTraj* __restrict__ trajs, int n_patches,
                                                                int ppx, int frames_per_cta, NormConst nc,
                                                                unsigned flags) {
    const FlTraj tr = trajs[blockIdx.z];
    const int f0 = blockIdx.y * frames_per_cta;
    if (f0 >= tr.n_frames) return;
    const int f1 = min(tr.n_frames, f0 + frames_per_cta);
    const int l = blockIdx.x;
    const int k = threadIdx.x;
    const size_t o = (size_t)l * ppx + k;
    const FlCellIdx id = tr.d_idx[o];
    const FlCellW w = tr.d_w[o];
    const double w0 = 1.0 - w.w1 - w.w2;
    const bool outside = id.tri < 0;
    const bool mask_aware = flags & FL_MASK_AWARE_NORM, no_norm = flags & FL_NO_NORM;
    for (int f = f0; f < f1; ++f) {
        const size_t t = (size_t)tr.t0 + (size_t)f * tr.interval;
        float v[3] = {0.f, 0.f, 0.f};
        bool masked = outside;
        if (!outside) {
            interp3(tr.d_velocity + t * tr.n_nodes * 2, tr.d_pressure + t * tr.n_nodes, id, w0, w.w1, w.w2, v);
            masked = !finite_f(v[2]);                 // only the pressure mask is kept (simple_dataloader.py:114,119)
#pragma unroll
            for (int c = 0; c < 3; ++c) if (!finite_f(v[c])) v[c] = 0.f;   // mesh_utils.py:89, per channel
        }
        float* dst = tr.d_states + (((size_t)f * n_patches + l) * 3) * ppx + k;
        const bool do_norm = !no_norm && !(mask_aware && masked);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float x = v[c];
            if (do_norm) x = __fdiv_rn(__fsub_rn(x, nc.mean[c]), nc.stdv[c]);
            fl_stg_stream1(dst + (size_t)c * ppx, x);
        }
        if (tr.d_mask) tr.d_mask[((size_t)f * n_patches + l) * ppx + k] = masked ? 1 : 0;
    }
}

// mesh_utils.to_grid for n_fields scalar fields: no pad / patch order, [ix, iy] layout.
__global__ void k_to_grid(const FlCellIdx* __restrict__ idx, const FlCellW* __restrict__ wt, int n_cells_grid,
                          const float* __restrict__ val, int n_fields, int n_nodes, float* __restrict__ data,
                          uint8_t* __restrict__ mask) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells_grid) return;
    const FlCellIdx id = idx[c];
    const FlCellW w = wt[c];
    const double w0 = 1.0 - w.w1 - w.w2;
    for (int f = 0; f < n_fields; ++f) {
        float r = 0.f;
        bool m = true;
        if (id.tri >= 0) {
            const float* z = val + (size_t)f * n_nodes;
            r = (float)fma(w.w2, (double)__ldg(z + id.v2), fma(w.w1, (double)__ldg(z + id.v1), w0 * (double)__ldg(z + id.v0)));
            m = !finite_f(r);
            if (m) r = 0.f;
        }
        data[(size_t)f * n_cells_grid + c] = r;
        if (mask) mask[(size_t)f * n_cells_grid + c] = m ? 1 : 0;
    }
}

int launch_interp(const FlTraj* d_trajs, int n_traj, int max_frames, int n_patches, int px, int py, const float* h_mean,
                  const float* h_std, unsigned flags, cudaStream_t st) {
    NormConst nc;
    for (int c = 0; c < 3; ++c) { nc.mean[c] = h_mean ? h_mean[c] : 0.f; nc.stdv[c] = h_std ? h_std[c] : 1.f; }
    const int ppx = px * py;
    // enough CTAs to fill 148 SMs x 8 resident CTAs a few times over, but keep the per-thread
    // record reuse: at least 4 frames per CTA when there is that much work
    long ctas_per_frame = (long)n_patches * n_traj;
    int fpc = 16;
    while (fpc > 1 && ctas_per_frame * ((max_frames + fpc - 1) / fpc) < 4L * FL_SM_COUNT * 8) fpc >>= 1;
    dim3 grid(n_patches, (max_frames + fpc - 1) / fpc, n_traj);
    k_interp_patchify_gather<<<grid, ppx, 0, st>>>(d_trajs, n_patches, ppx, fpc, nc, flags);
    FL_LAUNCH_CHECK();
    return FL_OK;
}

}  // namespace

extern "C" int fl_interp_patchify_dev(const FlTraj* d_trajs, int n_traj, int max_frames, int n_patches, int px, int py,
                                      const float* h_mean, const float* h_std, unsigned flags, void* stream) {
    FL_REQUIRE(d_trajs, FL_E_ARG, "fl_interp_patchify_dev: null descriptor array");
    FL_REQUIRE(n_traj > 0 && n_traj <= 65535 && max_frames > 0 && n_patches > 0, FL_E_ARG,
               "fl_interp_patchify_dev: bad sizes (n_traj=%d max_frames=%d n_patches=%d)", n_traj, max_frames, n_patches);
    FL_REQUIRE(px > 0 && py > 0 && px * py <= 256 && (px * py) % 32 == 0, FL_E_ARG,
               "fl_interp_patchify_dev: patch of %dx%d pixels unsupported (need px*py <= 256, multiple of 32)", px, py);
    FL_REQUIRE((h_mean && h_std) || (flags & FL_NO_NORM), FL_E_ARG, "fl_interp_patchify_dev: mean/std missing");
    return launch_interp(d_trajs, n_traj, max_frames, n_patches, px, py, h_mean, h_std, flags, (cudaStream_t)stream);
}

extern "C" int fl_interp_patchify(const FlTraj* h_trajs, int n_traj, int n_patches, int px, int py, const float* h_mean,
                                  const float* h_std, unsigned flags, void* stream) {
    FL_REQUIRE(h_trajs, FL_E_ARG, "fl_interp_patchify: null descriptor array");
    FL_REQUIRE(n_traj > 0 && n_traj <= 65535, FL_E_ARG, "fl_interp_patchify: n_traj=%d out of range", n_traj);
    int max_frames = 0;
    for (int i = 0; i < n_traj; ++i) {
        const FlTraj& t = h_trajs[i];
        FL_REQUIRE(t.d_velocity && t.d_pressure && t.d_idx && t.d_w && t.d_states, FL_E_ARG,
                   "fl_interp_patchify: trajectory %d has a null pointer", i);
        FL_REQUIRE(t.n_nodes > 0 && t.n_frames > 0 && t.t0 >= 0 && t.interval > 0, FL_E_ARG,
                   "fl_interp_patchify: trajectory %d has bad sizes", i);
        max_frames = t.n_frames > max_frames ? t.n_frames : max_frames;
    }
    cudaStream_t st = (cudaStream_t)stream;
    FlTraj* d_trajs = nullptr;
    FL_CUDA(cudaMallocAsync((void**)&d_trajs, sizeof(FlTraj) * n_traj, st));
    FL_CUDA(cudaMemcpyAsync(d_trajs, h_trajs, sizeof(FlTraj) * n_traj, cudaMemcpyHostToDevice, st));
    int rc = fl_interp_patchify_dev(d_trajs, n_traj, max_frames, n_patches, px, py, h_mean, h_std, flags, stream);
    cudaFreeAsync(d_trajs, st);
    return rc;
}

extern "C" int fl_to_grid(const FlCellIdx* d_cell_idx, const FlCellW* d_cell_w, int nx, int ny, const float* d_val,
                          int n_fields, int n_nodes, float* d_data, uint8_t* d_mask, void* stream) {
    FL_REQUIRE(d_cell_idx && d_cell_w && d_val && d_data, FL_E_ARG, "fl_to_grid: null pointer");
    FL_REQUIRE(nx > 0 && ny > 0 && n_fields > 0 && n_nodes > 0, FL_E_ARG, "fl_to_grid: sizes must be positive");
    int n = nx * ny;
    k_to_grid<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_cell_idx, d_cell_w, n, d_val, n_fields, n_nodes, d_data, d_mask);
    FL_LAUNCH_CHECK();
    return FL_OK;
}
