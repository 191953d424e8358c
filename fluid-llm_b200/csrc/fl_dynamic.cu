// fl_dynamic.cu -- per-frame dynamic meshes: point location moves from one-off to the per-frame hot loop.
//
// True EAGLE trajectories (/root/reference/max/ds_download/eagle.py:123-144: pointcloud[T,N,2], triangles[T,F,3],
// VX/VY/PS per frame) have a different mesh in every frame, so the reference's per-sample chain
// get_mesh_interpolation -> 3 x to_grid -> _pad -> _patch -> _normalize (src/dataloader/mesh_utils.py:82-106,
// simple_dataloader.py:104-152,193-216) runs once per FRAME.  Here one call handles a whole window of frames with no
// host synchronisation:
//   bin_frames (fl_locate.cu)   the three binning kernels over all frames of a chunk at once (blockIdx.y = frame)
//   k_dyn_locate_interp         one CTA per (patch, frame), one thread per output pixel: grid cell of the pixel (pad,
//                               ring crop, y-flip), triangle id by the tie-break rule over the cell's bin, barycentric
//                               weights, gather of that frame's node values, fp64 sum -> fp32 -> mask -> normalise,
//                               written straight into the patchified layout.  No cell table is materialised.
// Frames are processed in chunks so the binning workspace stays small (fl_dyn_workspace_bytes).
// Triangle ids and values are, frame by frame, the ones fl_locate + fl_interp_patchify produce for that frame's mesh.
#include "fl_geom.cuh"

using namespace flg;

namespace {

constexpr int DYN_CHUNK = 128;       // frames binned per round of kernels
constexpr int DYN_ITEMS_PER_TRI = 8; // default room in the bin-item store, per triangle (plus 2 per bin)

struct DynGeom {
    int nx, ny, px, py, n_bx, n_by, crop, pad_x0, pad_y0, padded_ny, flip_y;
};

__global__ void __launch_bounds__(256) k_dyn_locate_interp(
    const float* __restrict__ pos_all, const float* __restrict__ vel_all, const float* __restrict__ prs_all, int n_nodes,
    int n_cells, const int* __restrict__ tri_v_all, const int* __restrict__ bin_start_all, const int* __restrict__ items_all,
    int nby, int nbins, int capacity, const float* __restrict__ ax, const float* __restrict__ ay, DynGeom g, NormConst nc,
    unsigned flags, int frame0, float* __restrict__ states, uint8_t* __restrict__ mask, int32_t* __restrict__ tri_out) {
    const int fl = blockIdx.y, f = frame0 + fl;          // frame inside the chunk / inside the call
    const int l = blockIdx.x, k = threadIdx.x, ppx = g.px * g.py;
    const float* pos = pos_all + (size_t)f * 2 * n_nodes;
    const int* tri_v = tri_v_all + (size_t)fl * 3 * n_cells;
    const int* bin_start = bin_start_all + (size_t)fl * (nbins + 1);
    const int* items = items_all + (size_t)fl * capacity;
    // output pixel -> grid cell: the same map as k_plan_patch_table (fl_locate.cu)
    const int j = k % g.py, i = k / g.py;
    const int bx = l / g.n_by, by = l - bx * g.n_by;
    int X = (bx + g.crop) * g.px + i, Y = (by + g.crop) * g.py + j;     // padded image coordinates
    if (g.flip_y) Y = g.padded_ny - 1 - Y;                               // airfoil_ds.py:80
    const int ix = X - g.pad_x0, iy = Y - g.pad_y0;
    FlCellIdx id{0, 0, 0, -1};
    double w1 = 0.0, w2 = 0.0;
    if (ix >= 0 && ix < g.nx && iy >= 0 && iy < g.ny) {
        const double qx = (double)ax[ix], qy = (double)ay[iy];
        const int b = (ix / BIN) * nby + iy / BIN;
        const int kbeg = bin_start[b], kend = min(bin_start[b + 1], capacity);    // overflow is reported, never read
        const int tri = locate_in_bin(pos, tri_v, items, kbeg, kend, qx, qy);
        if (tri >= 0) {
            id.v0 = tri_v[3 * tri]; id.v1 = tri_v[3 * tri + 1]; id.v2 = tri_v[3 * tri + 2]; id.tri = tri;
            cell_weights(pos, id.v0, id.v1, id.v2, qx, qy, w1, w2);
        }
    }
    float v[3] = {0.f, 0.f, 0.f};
    bool masked = id.tri < 0;
    if (!masked) {
        interp3(vel_all + (size_t)f * 2 * n_nodes, prs_all + (size_t)f * n_nodes, id, 1.0 - w1 - w2, w1, w2, v);
        masked = !finite_f(v[2]);                 // only the pressure mask is kept (simple_dataloader.py:114,119)
#pragma unroll
        for (int c = 0; c < 3; ++c) if (!finite_f(v[c])) v[c] = 0.f;   // mesh_utils.py:89, per channel
    }
    const size_t n_patches = (size_t)g.n_bx * g.n_by;
    float* dst = states + (((size_t)f * n_patches + l) * 3) * ppx + k;
    const bool do_norm = !(flags & FL_NO_NORM) && !((flags & FL_MASK_AWARE_NORM) && masked);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float x = v[c];
        if (do_norm) x = __fdiv_rn(__fsub_rn(x, nc.mean[c]), nc.stdv[c]);
        fl_stg_stream1(dst + (size_t)c * ppx, x);
    }
    if (mask) mask[((size_t)f * n_patches + l) * ppx + k] = masked ? 1 : 0;
    if (tri_out) tri_out[((size_t)f * n_patches + l) * ppx + k] = id.tri;
}

}  // namespace

extern "C" size_t fl_dyn_workspace_bytes(int n_frames, int n_cells, int nx, int ny) {
    if (n_frames <= 0 || n_cells <= 0 || nx <= 0 || ny <= 0) return 0;
    const int chunk = n_frames < DYN_CHUNK ? n_frames : DYN_CHUNK;
    const size_t nbins = (size_t)((nx + BIN - 1) / BIN) * ((ny + BIN - 1) / BIN);
    const size_t cap = (size_t)DYN_ITEMS_PER_TRI * n_cells + 2 * nbins + 1024;
    return bin_ws_fixed_bytes(chunk, n_cells, (int)nbins) + sizeof(int) * cap * chunk + 256;
}

extern "C" int fl_dyn_interp_patchify(const float* d_pos, const int32_t* d_cells, const float* d_velocity,
                                      const float* d_pressure, int n_frames, int n_nodes, int n_cells,
                                      const float* d_grid_ax, const float* d_grid_ay, int nx, int ny, int px, int py,
                                      int crop_patches, const float* h_mean, const float* h_std, unsigned flags,
                                      float* d_states, uint8_t* d_mask, int32_t* d_tri, int32_t* d_status,
                                      void* d_workspace, size_t workspace_bytes, void* stream) {
    FL_REQUIRE(d_pos && d_cells && d_velocity && d_pressure && d_grid_ax && d_grid_ay && d_states && d_status && d_workspace,
               FL_E_ARG, "fl_dyn_interp_patchify: null pointer");
    FL_REQUIRE(n_frames > 0 && n_nodes > 0 && n_cells > 0 && nx > 0 && ny > 0, FL_E_ARG,
               "fl_dyn_interp_patchify: sizes must be positive");
    FL_REQUIRE(nx <= 32767 * BIN && ny <= 32767 * BIN, FL_E_ARG, "fl_dyn_interp_patchify: grid too large");
    FL_REQUIRE(px > 0 && py > 0 && px * py <= 256 && (px * py) % 32 == 0 && crop_patches >= 0, FL_E_ARG,
               "fl_dyn_interp_patchify: patch of %dx%d pixels unsupported (need px*py <= 256, multiple of 32)", px, py);
    FL_REQUIRE((h_mean && h_std) || (flags & FL_NO_NORM), FL_E_ARG, "fl_dyn_interp_patchify: mean/std missing");
    FL_REQUIRE(((uintptr_t)d_workspace & 255) == 0 && ((uintptr_t)d_velocity & 7) == 0, FL_E_ALIGN,
               "fl_dyn_interp_patchify: workspace must be 256-byte aligned, velocity 8-byte aligned");
    DynGeom g;
    const int pad_x = ((-nx) % px + px) % px, pad_y = ((-ny) % py + py) % py;   // simple_dataloader.py:140-141
    g.nx = nx; g.ny = ny; g.px = px; g.py = py; g.crop = crop_patches;
    g.n_bx = (nx + pad_x) / px - 2 * crop_patches; g.n_by = (ny + pad_y) / py - 2 * crop_patches;
    g.pad_x0 = pad_x / 2; g.pad_y0 = pad_y / 2; g.padded_ny = ny + pad_y; g.flip_y = (flags & FL_FLIP_Y) ? 1 : 0;
    FL_REQUIRE(g.n_bx > 0 && g.n_by > 0, FL_E_ARG, "fl_dyn_interp_patchify: no patches left after cropping");
    NormConst nc;
    for (int c = 0; c < 3; ++c) { nc.mean[c] = h_mean ? h_mean[c] : 0.f; nc.stdv[c] = h_std ? h_std[c] : 1.f; }
    cudaStream_t st = (cudaStream_t)stream;
    const int chunk = n_frames < DYN_CHUNK ? n_frames : DYN_CHUNK;
    BinWs w;
    FL_REQUIRE(bin_ws_carve(d_workspace, workspace_bytes, chunk, n_cells, nx, ny, &w), FL_E_WORKSPACE,
               "fl_dyn_interp_patchify: workspace too small (%zu bytes, see fl_dyn_workspace_bytes)", workspace_bytes);
    FL_CUDA(cudaMemsetAsync(d_status, 0, 2 * sizeof(int32_t), st));
    w.flags = d_status;                 // bad-id count and the largest per-frame item count go to the caller's status words
    for (int f0 = 0; f0 < n_frames; f0 += chunk) {
        w.n_frames = n_frames - f0 < chunk ? n_frames - f0 : chunk;
        int rc = bin_frames(d_pos + (size_t)f0 * 2 * n_nodes, 2 * (size_t)n_nodes, d_cells + (size_t)f0 * 3 * n_cells, n_nodes,
                            d_grid_ax, d_grid_ay, nx, ny, w, true, st);
        if (rc) return rc;
        dim3 grid(g.n_bx * g.n_by, w.n_frames);
        k_dyn_locate_interp<<<grid, px * py, 0, st>>>(d_pos, d_velocity, d_pressure, n_nodes, n_cells, w.tri_v, w.bin_start,
                                                      w.items, w.nby, w.nbins, w.capacity, d_grid_ax, d_grid_ay, g, nc, flags,
                                                      f0, d_states, d_mask, d_tri);
        FL_LAUNCH_CHECK();
    }
    return FL_OK;
}

extern "C" int fl_dyn_capacity(int n_frames, int n_cells, int nx, int ny, size_t workspace_bytes) {
    BinWs w;
    const int chunk = n_frames < DYN_CHUNK ? n_frames : DYN_CHUNK;
    if (n_frames <= 0 || n_cells <= 0 || nx <= 0 || ny <= 0) return 0;
    if (!bin_ws_carve((void*)256, workspace_bytes, chunk, n_cells, nx, ny, &w)) return 0;
    return w.capacity;
}
