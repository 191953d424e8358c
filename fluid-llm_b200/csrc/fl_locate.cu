// fl_locate.cu -- one-off per mesh: triangle binning + point location on the regular grid.
//
// Replaces /root/reference/src/dataloader/mesh_utils.py:103-104 (matplotlib Triangulation +
// TrapezoidMapTriFinder.find_many).  matplotlib builds a randomised trapezoidal map on one CPU
// thread and walks it per query; here every grid cell tests the few triangles whose bounding
// box covers its block of cells, with the same fp64 orientation expression and the tie-break
// rule stated in include/fluidgrid.h, so triangle ids come out bit-identical.
//
// Pipeline (all on `stream`):
//   k_tri_prepare   per triangle: index check, counter-clockwise fix (matplotlib
//                   correct_triangles), bounding box -> range of grid-cell blocks ("bins"),
//                   per-bin counts
//   k_scan          exclusive scan of the bin counts (one CTA; the bin grid is small)
//   k_fill          per triangle: scatter its id into the bins it overlaps
//   k_locate        per grid cell: evaluate the rule over its bin, emit tri id + table record
// Bins are blocks of BIN x BIN grid cells, so the cost follows the number of grid cells and a
// triangle that covers no grid point is dropped in the first kernel.
#include "fl_geom.cuh"

using namespace flg;

namespace {

// blockIdx.y = frame (mesh); the single-mesh entry point runs with one frame
__global__ void k_tri_prepare(const float* __restrict__ pos_all, size_t pos_stride, const int* __restrict__ cells_all,
                              int n_nodes, int n_cells, const float* __restrict__ ax, const float* __restrict__ ay, int nx,
                              int ny, int nby, int nbins, int* __restrict__ tri_v_all, TriRange* __restrict__ tri_range_all,
                              int* __restrict__ bin_count_all, int* __restrict__ flags) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cells) return;
    const int f = blockIdx.y;
    const float* pos = pos_all + (size_t)f * pos_stride;
    const int* cells = cells_all + (size_t)f * 3 * n_cells;
    int* tri_v = tri_v_all + (size_t)f * 3 * n_cells;
    TriRange* tri_range = tri_range_all + (size_t)f * n_cells;
    int* bin_count = bin_count_all + (size_t)f * (nbins + 1);
    int v0 = cells[3 * t], v1 = cells[3 * t + 1], v2 = cells[3 * t + 2];
    TriRange r{1, 0, 1, 0};
    if ((unsigned)v0 >= (unsigned)n_nodes || (unsigned)v1 >= (unsigned)n_nodes || (unsigned)v2 >= (unsigned)n_nodes) {
        atomicAdd(&flags[0], 1);
        tri_v[3 * t] = tri_v[3 * t + 1] = tri_v[3 * t + 2] = 0;
        tri_range[t] = r;
        return;
    }
    double x0 = pos[2 * v0], y0 = pos[2 * v0 + 1];
    double x1 = pos[2 * v1], y1 = pos[2 * v1 + 1];
    double x2 = pos[2 * v2], y2 = pos[2 * v2 + 1];
    // matplotlib Triangulation::correct_triangles: clockwise -> swap vertices 1 and 2
    double cz = __dsub_rn(__dmul_rn(__dsub_rn(x1, x0), __dsub_rn(y2, y0)), __dmul_rn(__dsub_rn(y1, y0), __dsub_rn(x2, x0)));
    if (cz < 0.0) { int s = v1; v1 = v2; v2 = s; }
    tri_v[3 * t] = v0; tri_v[3 * t + 1] = v1; tri_v[3 * t + 2] = v2;
    double xmin = fmin(x0, fmin(x1, x2)), xmax = fmax(x0, fmax(x1, x2));
    double ymin = fmin(y0, fmin(y1, y2)), ymax = fmax(y0, fmax(y1, y2));
    int ix0 = lower_bound_f(ax, nx, xmin), ix1 = upper_bound_f(ax, nx, xmax) - 1;
    int iy0 = lower_bound_f(ay, ny, ymin), iy1 = upper_bound_f(ay, ny, ymax) - 1;
    if (ix0 <= ix1 && iy0 <= iy1) {
        r.bx0 = (short)(ix0 / BIN); r.bx1 = (short)(ix1 / BIN);
        r.by0 = (short)(iy0 / BIN); r.by1 = (short)(iy1 / BIN);
        for (int bx = r.bx0; bx <= r.bx1; ++bx)
            for (int by = r.by0; by <= r.by1; ++by) atomicAdd(&bin_count[bx * nby + by], 1);
    }
    tri_range[t] = r;
}

// one CTA per frame: exclusive scan of count[0..n) into start[0..n], total in start[n]; flags[1] = max total over the frames
__global__ void k_scan(const int* __restrict__ count_all, int* __restrict__ start_all, int n, int* __restrict__ flags) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    const int* count = count_all + (size_t)blockIdx.x * (n + 1);
    int* start = start_all + (size_t)blockIdx.x * (n + 1);
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = i < n ? count[i] : 0;
        int s = v;
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += u; }
        if (lane == 31) warp_sum[warp] = s;
        __syncthreads();
        if (warp == 0) {
            int w = lane < (blockDim.x >> 5) ? warp_sum[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
            warp_sum[lane] = w;
        }
        __syncthreads();
        int prefix = carry + (warp ? warp_sum[warp - 1] : 0) + s - v;
        if (i < n) start[i] = prefix;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) { start[n] = carry; atomicMax(&flags[1], carry); }
}

__global__ void k_fill(const TriRange* __restrict__ tri_range_all, int n_cells, int nby, int nbins,
                       const int* __restrict__ bin_start_all, int* __restrict__ cursor_all, int* __restrict__ items_all,
                       int capacity) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cells) return;
    const int f = blockIdx.y;
    const int* bin_start = bin_start_all + (size_t)f * (nbins + 1);
    int* cursor = cursor_all + (size_t)f * (nbins + 1);
    int* items = items_all + (size_t)f * capacity;
    TriRange r = tri_range_all[(size_t)f * n_cells + t];
    for (int bx = r.bx0; bx <= r.bx1; ++bx)
        for (int by = r.by0; by <= r.by1; ++by) {
            int b = bx * nby + by;
            int p = bin_start[b] + atomicAdd(&cursor[b], 1);
            if (p < capacity) items[p] = t;
        }
}

__global__ void k_locate(const float* __restrict__ pos, const int* __restrict__ tri_v, const float* __restrict__ ax,
                         const float* __restrict__ ay, int nx, int ny, int nby, const int* __restrict__ bin_start,
                         const int* __restrict__ items, int capacity, int* __restrict__ tri_index, FlCellIdx* __restrict__ cell_idx,
                         FlCellW* __restrict__ cell_w) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nx * ny) return;
    int ix = c / ny, iy = c - ix * ny;
    double qx = (double)ax[ix], qy = (double)ay[iy];
    int b = (ix / BIN) * nby + iy / BIN;
    // (an item store that overflowed holds the first `capacity` entries only: the asynchronous entry point reports it)
    int tri = locate_in_bin(pos, tri_v, items, min(bin_start[b], capacity), min(bin_start[b + 1], capacity), qx, qy);
    if (tri_index) tri_index[c] = tri;
    if (cell_idx) {
        FlCellIdx rec{0, 0, 0, -1};
        FlCellW w{0.0, 0.0};
        if (tri >= 0) {
            rec.v0 = tri_v[3 * tri]; rec.v1 = tri_v[3 * tri + 1]; rec.v2 = tri_v[3 * tri + 2]; rec.tri = tri;
            cell_weights(pos, rec.v0, rec.v1, rec.v2, qx, qy, w.w1, w.w2);
        }
        cell_idx[c] = rec;
        cell_w[c] = w;
    }
}

__global__ void k_plan_patch_table(const FlCellIdx* __restrict__ cell_idx, const FlCellW* __restrict__ cell_w, int nx,
                                   int ny, int px, int py, int sx, int sy, int n_bx, int n_by, int crop, int pad_x0, int pad_y0,
                                   int padded_ny, int flip_y, FlCellIdx* __restrict__ out_idx,
                                   FlCellW* __restrict__ out_w, const int* __restrict__ node_slot, FlCellIdx* __restrict__ out_idx_slot) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    int total = n_bx * n_by * px * py;
    if (o >= total) return;
    int j = o % py, i = (o / py) % px, l = o / (px * py);
    int bx = l / n_by, by = l - bx * n_by;
    int X = crop * px + bx * sx + i, Y = crop * py + by * sy + j; // padded image coordinates: `crop` patch sizes of pixels cut, then unfold
    if (flip_y) Y = padded_ny - 1 - Y;                             // airfoil_ds.py:80
    int ix = X - pad_x0, iy = Y - pad_y0;
    FlCellIdx rec{0, 0, 0, -1};
    FlCellW w{0.0, 0.0};
    if (ix >= 0 && ix < nx && iy >= 0 && iy < ny) { rec = cell_idx[ix * ny + iy]; w = cell_w[ix * ny + iy]; }
    out_idx[o] = rec;
    out_w[o] = w;
    if (out_idx_slot) {          // the same record with node ids replaced by their shared-memory slots (cells inside the mesh only)
        if (rec.tri >= 0) { rec.v0 = node_slot[rec.v0]; rec.v1 = node_slot[rec.v1]; rec.v2 = node_slot[rec.v2]; }
        out_idx_slot[o] = rec;
    }
}

__global__ void k_locate_status(const int* __restrict__ flags, int capacity, int* __restrict__ status) {
    status[0] = flags[0];
    status[1] = flags[1] > capacity ? flags[1] - capacity : 0;
}

}  // namespace

int flg::bin_frames(const float* d_pos, size_t pos_stride, const int* d_cells, int n_nodes, const float* d_ax, const float* d_ay,
                    int nx, int ny, const BinWs& w, bool keep_flags, cudaStream_t st) {
    // bin_count, bin_start, cursor (and the flags behind them) start at zero
    FL_CUDA(cudaMemsetAsync(w.bin_count, 0, w.zero_bytes, st));
    if (!keep_flags) FL_CUDA(cudaMemsetAsync(w.flags, 0, 2 * sizeof(int), st));
    const int tb = 128;
    dim3 grid((w.n_cells + tb - 1) / tb, w.n_frames);
    k_tri_prepare<<<grid, tb, 0, st>>>(d_pos, pos_stride, d_cells, n_nodes, w.n_cells, d_ax, d_ay, nx, ny, w.nby, w.nbins,
                                       w.tri_v, w.tri_range, w.bin_count, w.flags);
    FL_LAUNCH_CHECK();
    k_scan<<<w.n_frames, 1024, 0, st>>>(w.bin_count, w.bin_start, w.nbins, w.flags);
    FL_LAUNCH_CHECK();
    k_fill<<<grid, tb, 0, st>>>(w.tri_range, w.n_cells, w.nby, w.nbins, w.bin_start, w.cursor, w.items, w.capacity);
    FL_LAUNCH_CHECK();
    return FL_OK;
}

extern "C" size_t fl_locate_workspace_bytes(int n_nodes, int n_cells) {
    (void)n_nodes;
    if (n_cells < 0) return 0;
    // fixed part for grids up to 4096 x 4096 cells plus room for 16 bin entries per triangle
    size_t nbins_max = (size_t)(4096 / BIN) * (4096 / BIN);
    return bin_ws_fixed_bytes(1, n_cells, (int)nbins_max) + sizeof(int) * (16 * (size_t)n_cells + 4 * nbins_max) + 256;
}

static int locate_impl(const float* d_pos, const int32_t* d_cells, int n_nodes, int n_cells, const float* d_grid_ax,
                       const float* d_grid_ay, int nx, int ny, int32_t* d_tri_index, FlCellIdx* d_cell_idx, FlCellW* d_cell_w,
                       void* d_workspace, size_t workspace_bytes, int32_t* d_status, const char* who, cudaStream_t st) {
    FL_REQUIRE(d_pos && d_cells && d_grid_ax && d_grid_ay && d_workspace, FL_E_ARG, "%s: null pointer", who);
    FL_REQUIRE(n_nodes > 0 && n_cells > 0 && nx > 0 && ny > 0, FL_E_ARG, "%s: sizes must be positive", who);
    FL_REQUIRE(nx <= 32767 * BIN && ny <= 32767 * BIN, FL_E_ARG, "%s: grid too large", who);
    FL_REQUIRE((d_cell_idx == nullptr) == (d_cell_w == nullptr), FL_E_ARG, "%s: d_cell_idx and d_cell_w go together", who);
    FL_REQUIRE(((uintptr_t)d_workspace & 255) == 0, FL_E_ALIGN, "%s: workspace must be 256-byte aligned", who);
    BinWs w;
    FL_REQUIRE(bin_ws_carve(d_workspace, workspace_bytes, 1, n_cells, nx, ny, &w), FL_E_WORKSPACE, "%s: workspace too small (%zu bytes)", who,
               workspace_bytes);
    int rc = bin_frames(d_pos, 0, d_cells, n_nodes, d_grid_ax, d_grid_ay, nx, ny, w, false, st);
    if (rc) return rc;
    if (d_status) {
        k_locate_status<<<1, 1, 0, st>>>(w.flags, w.capacity, d_status);      // no host synchronisation: the caller reads it later
        FL_LAUNCH_CHECK();
    } else {
        int h_flags[2];
        FL_CUDA(cudaMemcpyAsync(h_flags, w.flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
        FL_CUDA(cudaStreamSynchronize(st));  // one-off per mesh: the item count decides whether the workspace fits
        FL_REQUIRE(h_flags[0] == 0, FL_E_RANGE, "%s: %d triangles index nodes outside 0 <= i < %d", who, h_flags[0], n_nodes);
        FL_REQUIRE(h_flags[1] <= w.capacity, FL_E_WORKSPACE, "%s: workspace too small, need %zu more bytes", who,
                   sizeof(int) * ((size_t)h_flags[1] - (size_t)w.capacity));
    }
    int n = nx * ny;
    k_locate<<<(n + 127) / 128, 128, 0, st>>>(d_pos, w.tri_v, d_grid_ax, d_grid_ay, nx, ny, w.nby, w.bin_start, w.items, w.capacity,
                                              d_tri_index, d_cell_idx, d_cell_w);
    FL_LAUNCH_CHECK();
    return FL_OK;
}

extern "C" int fl_locate(const float* d_pos, const int32_t* d_cells, int n_nodes, int n_cells, const float* d_grid_ax,
                         const float* d_grid_ay, int nx, int ny, int32_t* d_tri_index, FlCellIdx* d_cell_idx,
                         FlCellW* d_cell_w, void* d_workspace, size_t workspace_bytes, void* stream) {
    return locate_impl(d_pos, d_cells, n_nodes, n_cells, d_grid_ax, d_grid_ay, nx, ny, d_tri_index, d_cell_idx, d_cell_w, d_workspace,
                       workspace_bytes, nullptr, "fl_locate", (cudaStream_t)stream);
}

extern "C" int fl_locate_async(const float* d_pos, const int32_t* d_cells, int n_nodes, int n_cells, const float* d_grid_ax,
                               const float* d_grid_ay, int nx, int ny, int32_t* d_tri_index, FlCellIdx* d_cell_idx,
                               FlCellW* d_cell_w, void* d_workspace, size_t workspace_bytes, int32_t* d_status, void* stream) {
    FL_REQUIRE(d_status, FL_E_ARG, "fl_locate_async: null status pointer");
    return locate_impl(d_pos, d_cells, n_nodes, n_cells, d_grid_ax, d_grid_ay, nx, ny, d_tri_index, d_cell_idx, d_cell_w, d_workspace,
                       workspace_bytes, d_status, "fl_locate_async", (cudaStream_t)stream);
}

extern "C" int fl_plan_patch_table(const FlCellIdx* d_cell_idx, const FlCellW* d_cell_w, int nx, int ny, int px, int py,
                                   int sx, int sy, int crop_patches, unsigned flags, FlCellIdx* d_out_idx, FlCellW* d_out_w,
                                   int* h_n_bx, int* h_n_by, const int32_t* d_node_slot, FlCellIdx* d_out_idx_slot, void* stream) {
    FL_REQUIRE(nx > 0 && ny > 0 && px > 0 && py > 0 && sx >= 0 && sy >= 0 && crop_patches >= 0, FL_E_ARG, "fl_plan_patch_table: bad sizes");
    if (sx == 0) sx = px;
    if (sy == 0) sy = py;
    int pad_x = 0, pad_y = 0;
    if (!(flags & FL_NO_PAD)) { pad_x = ((-nx) % px + px) % px; pad_y = ((-ny) % py + py) % py; }   // simple_dataloader.py:140-141
    // F.unfold over what is left after `crop_patches` patch sizes of pixels are cut from every side (airfoil_ds.py:132-135):
    // floor((extent - patch) / stride) + 1 windows, none when the extent is shorter than a patch
    int ext_x = nx + pad_x - 2 * crop_patches * px, ext_y = ny + pad_y - 2 * crop_patches * py;
    int n_bx = ext_x >= px ? (ext_x - px) / sx + 1 : 0, n_by = ext_y >= py ? (ext_y - py) / sy + 1 : 0;
    if (h_n_bx) *h_n_bx = n_bx;
    if (h_n_by) *h_n_by = n_by;
    if (!d_out_idx && !d_out_w) return FL_OK;   // size query
    FL_REQUIRE(d_cell_idx && d_cell_w && d_out_idx && d_out_w, FL_E_ARG, "fl_plan_patch_table: null pointer");
    FL_REQUIRE((d_node_slot == nullptr) == (d_out_idx_slot == nullptr), FL_E_ARG, "fl_plan_patch_table: d_node_slot and d_out_idx_slot go together");
    FL_REQUIRE(n_bx > 0 && n_by > 0, FL_E_ARG, "fl_plan_patch_table: no patches left after cropping");
    long total = (long)n_bx * n_by * px * py;
    FL_REQUIRE(total < 0x7fffffffL, FL_E_ARG, "fl_plan_patch_table: too many pixels");
    k_plan_patch_table<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_cell_idx, d_cell_w, nx, ny, px, py, sx, sy, n_bx, n_by, crop_patches, pad_x / 2, pad_y / 2, ny + pad_y,
        (flags & FL_FLIP_Y) ? 1 : 0, d_out_idx, d_out_w, d_node_slot, d_out_idx_slot);
    FL_LAUNCH_CHECK();
    return FL_OK;
}
