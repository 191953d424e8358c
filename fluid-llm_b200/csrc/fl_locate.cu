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
#include "fl_common.cuh"

namespace {

constexpr int BIN = 4;  // grid cells per bin side

struct TriRange { short bx0, bx1, by0, by1; };  // inclusive bin range, bx0 > bx1 = empty

__device__ __forceinline__ int lower_bound_f(const float* a, int n, double v) {  // first i: a[i] >= v
    int lo = 0, hi = n;
    while (lo < hi) { int m = (lo + hi) >> 1; if ((double)a[m] < v) lo = m + 1; else hi = m; }
    return lo;
}
__device__ __forceinline__ int upper_bound_f(const float* a, int n, double v) {  // first i: a[i] > v
    int lo = 0, hi = n;
    while (lo < hi) { int m = (lo + hi) >> 1; if ((double)a[m] <= v) lo = m + 1; else hi = m; }
    return lo;
}

// (p - l) x (r - l) with every product and difference rounded separately (no FMA contraction):
// the expression matplotlib's Edge::get_point_orientation evaluates on x86-64.
__device__ __forceinline__ double orient(double px, double py, double lx, double ly, double rx, double ry) {
    double a = __dmul_rn(__dsub_rn(px, lx), __dsub_rn(ry, ly));
    double b = __dmul_rn(__dsub_rn(py, ly), __dsub_rn(rx, lx));
    return __dsub_rn(a, b);
}

__global__ void k_tri_prepare(const float* __restrict__ pos, const int* __restrict__ cells, int n_nodes, int n_cells,
                              const float* __restrict__ ax, const float* __restrict__ ay, int nx, int ny, int nbx,
                              int nby, int* __restrict__ tri_v, TriRange* __restrict__ tri_range,
                              int* __restrict__ bin_count, int* __restrict__ flags) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cells) return;
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

// exclusive scan of count[0..n) into start[0..n], total in start[n] and flags[1]
__global__ void k_scan(const int* __restrict__ count, int* __restrict__ start, int n, int* __restrict__ flags) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
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
    if (threadIdx.x == 0) { start[n] = carry; flags[1] = carry; }
}

__global__ void k_fill(const TriRange* __restrict__ tri_range, int n_cells, int nby, const int* __restrict__ bin_start,
                       int* __restrict__ cursor, int* __restrict__ items, int capacity) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cells) return;
    TriRange r = tri_range[t];
    for (int bx = r.bx0; bx <= r.bx1; ++bx)
        for (int by = r.by0; by <= r.by1; ++by) {
            int b = bx * nby + by;
            int p = bin_start[b] + atomicAdd(&cursor[b], 1);
            if (p < capacity) items[p] = t;
        }
}

// The tie-break rule (include/fluidgrid.h), evaluated for one (cell, triangle) pair.
// Returns -1 = rejected, 0 = accepted with full priority (strictly inside, on a vertex, or on
// an edge the triangle lies above), 1 = accepted only if nothing lies above that edge.
__device__ __forceinline__ int rule_eval(double qx, double qy, const double* vx, const double* vy) {
    if ((qx == vx[0] && qy == vy[0]) || (qx == vx[1] && qy == vy[1]) || (qx == vx[2] && qy == vy[2])) return 0;
    int prio = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int k1 = (k + 1) % 3;
        bool end_right = (vx[k1] == vx[k]) ? (vy[k1] > vy[k]) : (vx[k1] > vx[k]);
        double s = end_right ? orient(qx, qy, vx[k], vy[k], vx[k1], vy[k1])
                             : orient(qx, qy, vx[k1], vy[k1], vx[k], vy[k]);
        if (end_right) { if (!(s <= 0.0)) return -1; }         // triangle is above this edge
        else { if (!(s >= 0.0)) return -1; if (s == 0.0) prio = 1; }  // triangle is below it
    }
    return prio;
}

__global__ void k_locate(const float* __restrict__ pos, const int* __restrict__ tri_v, const float* __restrict__ ax,
                         const float* __restrict__ ay, int nx, int ny, int nby, const int* __restrict__ bin_start,
                         const int* __restrict__ items, int* __restrict__ tri_index, FlCellIdx* __restrict__ cell_idx,
                         FlCellW* __restrict__ cell_w) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nx * ny) return;
    int ix = c / ny, iy = c - ix * ny;
    double qx = (double)ax[ix], qy = (double)ay[iy];
    int b = (ix / BIN) * nby + iy / BIN;
    unsigned best = 0xffffffffu;  // (prio << 31) | tri
    for (int k = bin_start[b]; k < bin_start[b + 1]; ++k) {
        int t = items[k];
        double vx[3], vy[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { int v = tri_v[3 * t + j]; vx[j] = pos[2 * v]; vy[j] = pos[2 * v + 1]; }
        int p = rule_eval(qx, qy, vx, vy);
        if (p >= 0) best = min(best, ((unsigned)p << 31) | (unsigned)t);
    }
    int tri = best == 0xffffffffu ? -1 : (int)(best & 0x7fffffffu);
    if (tri_index) tri_index[c] = tri;
    if (cell_idx) {
        FlCellIdx rec{0, 0, 0, -1};
        FlCellW w{0.0, 0.0};
        if (tri >= 0) {
            rec.v0 = tri_v[3 * tri]; rec.v1 = tri_v[3 * tri + 1]; rec.v2 = tri_v[3 * tri + 2]; rec.tri = tri;
            double x0 = pos[2 * rec.v0], y0 = pos[2 * rec.v0 + 1];
            double e1x = (double)pos[2 * rec.v1] - x0, e1y = (double)pos[2 * rec.v1 + 1] - y0;
            double e2x = (double)pos[2 * rec.v2] - x0, e2y = (double)pos[2 * rec.v2 + 1] - y0;
            double dx = qx - x0, dy = qy - y0;
            double d = e1x * e2y - e2x * e1y;
            if (d != 0.0) {
                w.w1 = (dx * e2y - e2x * dy) / d;
                w.w2 = (e1x * dy - dx * e1y) / d;
            }
        }
        cell_idx[c] = rec;
        cell_w[c] = w;
    }
}

__global__ void k_plan_patch_table(const FlCellIdx* __restrict__ cell_idx, const FlCellW* __restrict__ cell_w, int nx,
                                   int ny, int px, int py, int n_bx, int n_by, int crop, int pad_x0, int pad_y0,
                                   int padded_ny, int flip_y, FlCellIdx* __restrict__ out_idx,
                                   FlCellW* __restrict__ out_w) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    int total = n_bx * n_by * px * py;
    if (o >= total) return;
    int j = o % py, i = (o / py) % px, l = o / (px * py);
    int bx = l / n_by, by = l - bx * n_by;
    int X = (bx + crop) * px + i, Y = (by + crop) * py + j;       // padded image coordinates
    if (flip_y) Y = padded_ny - 1 - Y;                             // airfoil_ds.py:80
    int ix = X - pad_x0, iy = Y - pad_y0;
    FlCellIdx rec{0, 0, 0, -1};
    FlCellW w{0.0, 0.0};
    if (ix >= 0 && ix < nx && iy >= 0 && iy < ny) { rec = cell_idx[ix * ny + iy]; w = cell_w[ix * ny + iy]; }
    out_idx[o] = rec;
    out_w[o] = w;
}

struct LocateWs {
    int* tri_v; TriRange* tri_range; int* bin_count; int* bin_start; int* cursor; int* flags; int* items;
    int nbx, nby, nbins, capacity;
};

size_t locate_fixed_bytes(int n_cells, int nbins) {
    size_t b = 0;
    b += fl_align_up(sizeof(int) * 3 * (size_t)n_cells, 256);
    b += fl_align_up(sizeof(TriRange) * (size_t)n_cells, 256);
    b += 3 * fl_align_up(sizeof(int) * ((size_t)nbins + 1), 256);
    b += 256;  // flags
    return b;
}

}  // namespace

extern "C" size_t fl_locate_workspace_bytes(int n_nodes, int n_cells) {
    (void)n_nodes;
    if (n_cells < 0) return 0;
    // fixed part for grids up to 4096 x 4096 cells plus room for 16 bin entries per triangle
    size_t nbins_max = (size_t)(4096 / BIN) * (4096 / BIN);
    return locate_fixed_bytes(n_cells, (int)nbins_max) + sizeof(int) * (16 * (size_t)n_cells + 4 * nbins_max) + 256;
}

extern "C" int fl_locate(const float* d_pos, const int32_t* d_cells, int n_nodes, int n_cells, const float* d_grid_ax,
                         const float* d_grid_ay, int nx, int ny, int32_t* d_tri_index, FlCellIdx* d_cell_idx,
                         FlCellW* d_cell_w, void* d_workspace, size_t workspace_bytes, void* stream) {
    FL_REQUIRE(d_pos && d_cells && d_grid_ax && d_grid_ay && d_workspace, FL_E_ARG, "fl_locate: null pointer");
    FL_REQUIRE(n_nodes > 0 && n_cells > 0 && nx > 0 && ny > 0, FL_E_ARG, "fl_locate: sizes must be positive");
    FL_REQUIRE(nx <= 32767 * BIN && ny <= 32767 * BIN, FL_E_ARG, "fl_locate: grid too large");
    FL_REQUIRE((d_cell_idx == nullptr) == (d_cell_w == nullptr), FL_E_ARG, "fl_locate: d_cell_idx and d_cell_w go together");
    FL_REQUIRE(((uintptr_t)d_workspace & 255) == 0, FL_E_ALIGN, "fl_locate: workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    LocateWs w;
    w.nbx = (nx + BIN - 1) / BIN; w.nby = (ny + BIN - 1) / BIN; w.nbins = w.nbx * w.nby;
    size_t fixed = locate_fixed_bytes(n_cells, w.nbins);
    FL_REQUIRE(workspace_bytes >= fixed + 1024, FL_E_WORKSPACE, "fl_locate: workspace too small (%zu < %zu)",
               workspace_bytes, fixed + 1024);
    char* p = (char*)d_workspace;
    w.tri_v = (int*)p; p += fl_align_up(sizeof(int) * 3 * (size_t)n_cells, 256);
    w.tri_range = (TriRange*)p; p += fl_align_up(sizeof(TriRange) * (size_t)n_cells, 256);
    w.bin_count = (int*)p; p += fl_align_up(sizeof(int) * ((size_t)w.nbins + 1), 256);
    w.bin_start = (int*)p; p += fl_align_up(sizeof(int) * ((size_t)w.nbins + 1), 256);
    w.cursor = (int*)p; p += fl_align_up(sizeof(int) * ((size_t)w.nbins + 1), 256);
    w.flags = (int*)p; p += 256;
    w.items = (int*)p;
    size_t cap = (workspace_bytes - (size_t)(p - (char*)d_workspace)) / sizeof(int);
    w.capacity = cap > 0x7fffffff ? 0x7fffffff : (int)cap;

    FL_CUDA(cudaMemsetAsync(w.bin_count, 0, (size_t)((char*)w.items - (char*)w.bin_count), st));
    int tb = 128;
    k_tri_prepare<<<(n_cells + tb - 1) / tb, tb, 0, st>>>(d_pos, d_cells, n_nodes, n_cells, d_grid_ax, d_grid_ay, nx, ny,
                                                          w.nbx, w.nby, w.tri_v, w.tri_range, w.bin_count, w.flags);
    FL_LAUNCH_CHECK();
    k_scan<<<1, 1024, 0, st>>>(w.bin_count, w.bin_start, w.nbins, w.flags);
    FL_LAUNCH_CHECK();
    int h_flags[2];
    FL_CUDA(cudaMemcpyAsync(h_flags, w.flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
    FL_CUDA(cudaStreamSynchronize(st));  // one-off per mesh: the item count decides whether the workspace fits
    FL_REQUIRE(h_flags[0] == 0, FL_E_RANGE, "fl_locate: %d triangles index nodes outside 0 <= i < %d", h_flags[0], n_nodes);
    FL_REQUIRE(h_flags[1] <= w.capacity, FL_E_WORKSPACE, "fl_locate: workspace too small, need %zu more bytes",
               sizeof(int) * ((size_t)h_flags[1] - (size_t)w.capacity));
    k_fill<<<(n_cells + tb - 1) / tb, tb, 0, st>>>(w.tri_range, n_cells, w.nby, w.bin_start, w.cursor, w.items, w.capacity);
    FL_LAUNCH_CHECK();
    int n = nx * ny;
    k_locate<<<(n + 127) / 128, 128, 0, st>>>(d_pos, w.tri_v, d_grid_ax, d_grid_ay, nx, ny, w.nby, w.bin_start, w.items,
                                              d_tri_index, d_cell_idx, d_cell_w);
    FL_LAUNCH_CHECK();
    return FL_OK;
}

extern "C" int fl_plan_patch_table(const FlCellIdx* d_cell_idx, const FlCellW* d_cell_w, int nx, int ny, int px, int py,
                                   int crop_patches, unsigned flags, FlCellIdx* d_out_idx, FlCellW* d_out_w,
                                   int* h_n_bx, int* h_n_by, void* stream) {
    FL_REQUIRE(nx > 0 && ny > 0 && px > 0 && py > 0 && crop_patches >= 0, FL_E_ARG, "fl_plan_patch_table: bad sizes");
    int pad_x = ((-nx) % px + px) % px, pad_y = ((-ny) % py + py) % py;   // simple_dataloader.py:140-141
    int n_bx = (nx + pad_x) / px - 2 * crop_patches, n_by = (ny + pad_y) / py - 2 * crop_patches;
    if (h_n_bx) *h_n_bx = n_bx;
    if (h_n_by) *h_n_by = n_by;
    if (!d_out_idx && !d_out_w) return FL_OK;   // size query
    FL_REQUIRE(d_cell_idx && d_cell_w && d_out_idx && d_out_w, FL_E_ARG, "fl_plan_patch_table: null pointer");
    FL_REQUIRE(n_bx > 0 && n_by > 0, FL_E_ARG, "fl_plan_patch_table: no patches left after cropping");
    long total = (long)n_bx * n_by * px * py;
    FL_REQUIRE(total < 0x7fffffffL, FL_E_ARG, "fl_plan_patch_table: too many pixels");
    k_plan_patch_table<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_cell_idx, d_cell_w, nx, ny, px, py, n_bx, n_by, crop_patches, pad_x / 2, pad_y / 2, ny + pad_y,
        (flags & FL_FLIP_Y) ? 1 : 0, d_out_idx, d_out_w);
    FL_LAUNCH_CHECK();
    return FL_OK;
}
