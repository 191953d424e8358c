// fl_dynamic.cu -- per-frame dynamic meshes: point location moves from one-off to the per-frame hot loop.
//
// True EAGLE trajectories (/root/reference/max/ds_download/eagle.py:123-144: pointcloud[T,N,2], triangles[T,F,3],
// VX/VY/PS per frame) have a different mesh in every frame, so the reference's per-sample chain
// get_mesh_interpolation -> 3 x to_grid -> _pad -> _patch -> _normalize (src/dataloader/mesh_utils.py:82-106,
// simple_dataloader.py:104-152,193-216) runs once per FRAME.  Here one call handles a whole window of frames with no
// host synchronisation:
// Two forms.  When the grid's triangle-id array and one frame's nodes fit in shared memory (every BASELINE config but
// the 2048 x 1024 grid) k_dyn_raster_interp does everything in one persistent kernel, see below.  Otherwise:
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
    int ppx_shift, py_shift;     // log2 of px*py and py when they are powers of two, else -1
};

__global__ void __launch_bounds__(256) k_dyn_locate_interp(
    const float* __restrict__ pos_all, const float* __restrict__ vel_all, const float* __restrict__ prs_all, int n_nodes,
    int n_cells, const int* __restrict__ tri_v_all, const int* __restrict__ bin_start_all, const int* __restrict__ items_all,
    int nby, int nbins, int capacity, const float* __restrict__ ax, const float* __restrict__ ay, DynGeom g, NormConst nc,
    unsigned flags, int frame0, float* __restrict__ states, uint8_t* __restrict__ mask, int32_t* __restrict__ tri_out) {
    const int fl = blockIdx.y, f = frame0 + fl;          // frame inside the chunk / inside the call
    const int l = blockIdx.x, ppx = g.px * g.py;
    const float* pos = pos_all + (size_t)f * 2 * n_nodes;
    const int* tri_v = tri_v_all + (size_t)fl * 3 * n_cells;
    const int* bin_start = bin_start_all + (size_t)fl * (nbins + 1);
    const int* items = items_all + (size_t)fl * capacity;
    for (int k = threadIdx.x; k < ppx; k += blockDim.x) {      // (one pixel per thread for patches of up to 256 pixels)
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
}


// ---- shared-memory rasterising form: one persistent CTA per SM, one frame at a time -------------------------------
// The whole grid's triangle-id array (4 B per cell), the frame's node positions and (when they fit) its node fields
// live in shared memory.  Phase 1 walks the TRIANGLES: each thread takes a triangle, fixes its winding, and evaluates the
// tie-break rule only at the grid points inside the triangle's bounding box (~12 instead of every candidate of a bin
// for every cell), publishing (prio << 31 | tri) with a shared-memory atomicMin -- the same minimum the per-cell loop
// of locate_in_bin takes, so the ids are identical.  Phase 2 walks the OUTPUT PIXELS: id -> vertices -> weights ->
// fp64 sum -> fp32 -> mask -> normalise -> patchified store (128 B per warp instruction).  No workspace, no binning
// kernels; HBM traffic is the compulsory 8 N + 12 F + 12 N in, 12 P (+ P mask) out per frame.
constexpr int RS_THREADS = 1024;

__global__ void __launch_bounds__(RS_THREADS, 1) k_dyn_raster_interp(
    const float* __restrict__ pos_all, const int* __restrict__ cells_all, const float* __restrict__ vel_all,
    const float* __restrict__ prs_all, int n_frames, int n_nodes, int n_cells, const float* __restrict__ ax_g,
    const float* __restrict__ ay_g, DynGeom g, NormConst nc, unsigned flags, int stage_fields, float* __restrict__ states,
    uint8_t* __restrict__ mask, int32_t* __restrict__ tri_out, int32_t* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    const int C = g.nx * g.ny, n_pad = (n_nodes + 1) & ~1;
    unsigned* s_cell = (unsigned*)dyn_smem;                       // [C]
    float2* s_pos = (float2*)(s_cell + ((C + 3) & ~3));           // [n_pad]
    float* s_ax = (float*)(s_pos + n_pad);                        // [nx]
    float* s_ay = s_ax + g.nx;                                    // [ny]
    float2* s_vel = (float2*)(s_ay + ((g.ny + g.nx + 1) & ~1) - g.nx);   // [n_pad] (8-byte aligned), only if stage_fields
    float* s_prs = (float*)(s_vel + n_pad);                       // [n_nodes]
    const int tid = threadIdx.x, nth = blockDim.x;
    const int ppx = g.px * g.py, n_patches = g.n_bx * g.n_by, n_out = n_patches * ppx;
    for (int i = tid; i < g.nx; i += nth) s_ax[i] = ax_g[i];
    for (int i = tid; i < g.ny; i += nth) s_ay[i] = ay_g[i];
    int bad = 0;
    for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
        const float2* pos = (const float2*)(pos_all + (size_t)f * 2 * n_nodes);
        const int* cells = cells_all + (size_t)f * 3 * n_cells;
        const float2* vel = (const float2*)(vel_all + (size_t)f * 2 * n_nodes);
        const float* prs = prs_all + (size_t)f * n_nodes;
        __syncthreads();                                          // previous frame's readers are done
        for (int i = tid; i < C; i += nth) s_cell[i] = 0xffffffffu;
        for (int i = tid; i < n_nodes; i += nth) s_pos[i] = pos[i];
        if (stage_fields) {
            for (int i = tid; i < n_nodes; i += nth) { s_vel[i] = vel[i]; s_prs[i] = prs[i]; }
        }
        __syncthreads();
        // ---- phase 1: triangles -> cells.  A warp takes 32 triangles: every lane sets one up (winding fix, bounding box
        // in grid indices), then the (triangle, grid point) pairs of all 32 are spread evenly over the lanes -- lane work
        // is uniform whatever the triangles' sizes.  A pair's owner is found by a 5-step search over the lanes' running
        // counts; its corner coordinates come over by shuffle.
        for (int base = (tid >> 5) * 32; base < n_cells; base += (nth >> 5) * 32) {
            const int lane = tid & 31, t = base + lane;
            float fx0 = 0.f, fy0 = 0.f, fx1 = 0.f, fy1 = 0.f, fx2 = 0.f, fy2 = 0.f;
            int ix0 = 0, iy0 = 0, h = 1, cnt = 0;
            unsigned hmagic = 0;            // ceil(2^32 / h): r / h == umulhi(r, hmagic) for r * h < 2^32 (0 when h == 1)
            if (t < n_cells) {
                const int v0 = cells[3 * t], v1 = cells[3 * t + 1], v2 = cells[3 * t + 2];
                if ((unsigned)v0 >= (unsigned)n_nodes || (unsigned)v1 >= (unsigned)n_nodes || (unsigned)v2 >= (unsigned)n_nodes) {
                    ++bad;
                } else {
                    const float2 a = s_pos[v0];
                    float2 b = s_pos[v1], c = s_pos[v2];
                    // matplotlib Triangulation::correct_triangles: clockwise -> swap vertices 1 and 2
                    const double x0 = a.x, y0 = a.y;
                    const double cz = __dsub_rn(__dmul_rn(__dsub_rn((double)b.x, x0), __dsub_rn((double)c.y, y0)),
                                                __dmul_rn(__dsub_rn((double)b.y, y0), __dsub_rn((double)c.x, x0)));
                    if (cz < 0.0) { const float2 q = b; b = c; c = q; }
                    fx0 = a.x; fy0 = a.y; fx1 = b.x; fy1 = b.y; fx2 = c.x; fy2 = c.y;
                    // grid points inside the closed bounding box (axes and corners are fp32: compared as such, exactly)
                    const float xmin = fminf(fx0, fminf(fx1, fx2)), xmax = fmaxf(fx0, fmaxf(fx1, fx2));
                    const float ymin = fminf(fy0, fminf(fy1, fy2)), ymax = fmaxf(fy0, fmaxf(fy1, fy2));
                    int lo = 0, hi = g.nx;
                    while (lo < hi) { const int m = (lo + hi) >> 1; if (s_ax[m] < xmin) lo = m + 1; else hi = m; }
                    ix0 = lo; hi = g.nx;
                    while (lo < hi) { const int m = (lo + hi) >> 1; if (s_ax[m] <= xmax) lo = m + 1; else hi = m; }
                    const int w = lo - ix0;
                    lo = 0; hi = g.ny;
                    while (lo < hi) { const int m = (lo + hi) >> 1; if (s_ay[m] < ymin) lo = m + 1; else hi = m; }
                    iy0 = lo; hi = g.ny;
                    while (lo < hi) { const int m = (lo + hi) >> 1; if (s_ay[m] <= ymax) lo = m + 1; else hi = m; }
                    h = lo - iy0;
                    cnt = (w > 0 && h > 0) ? w * h : 0;
                    if (h < 1) h = 1;
                    hmagic = h == 1 ? 0u : 0xffffffffu / (unsigned)h + 1u;
                }
            }
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            for (int j0 = 0; j0 < total; j0 += 32) {
                const int j = j0 + lane;
                int k = 0;                                        // owner: number of lanes whose running count is <= j
#pragma unroll
                for (int step = 16; step; step >>= 1) { const int v = __shfl_sync(0xffffffffu, incl, k + step - 1); if (v <= j) k += step; }
                const bool act = j < total;
                k = act ? k : 31;
                const int r = j - (__shfl_sync(0xffffffffu, incl, k) - __shfl_sync(0xffffffffu, cnt, k));
                const unsigned omagic = __shfl_sync(0xffffffffu, hmagic, k);
                const int oh = __shfl_sync(0xffffffffu, h, k), oix0 = __shfl_sync(0xffffffffu, ix0, k), oiy0 = __shfl_sync(0xffffffffu, iy0, k);
                const float ax0 = __shfl_sync(0xffffffffu, fx0, k), ay0 = __shfl_sync(0xffffffffu, fy0, k);
                const float ax1 = __shfl_sync(0xffffffffu, fx1, k), ay1 = __shfl_sync(0xffffffffu, fy1, k);
                const float ax2 = __shfl_sync(0xffffffffu, fx2, k), ay2 = __shfl_sync(0xffffffffu, fy2, k);
                if (act) {
                    const int rx = oh == 1 ? r : (int)__umulhi((unsigned)r, omagic), ry = r - rx * oh;
                    const int ix = oix0 + rx, iy = oiy0 + ry;
                    const float qx = s_ax[ix], qy = s_ay[iy];
                    // fp32 screen: cross(edge, point - start) for the three edges of the counter-clockwise triangle.  Every
                    // input is an fp32 number, so the computed value is within 2^-22 (|a| + |b|) of the exact one; with a
                    // 4x margin a value beyond the bound has the exact sign, which is all the tie-break rule looks at when
                    // the point is on no edge.  Clearly outside one edge -> rejected; clearly inside all three -> accepted
                    // with full priority; anything else (a point on or next to an edge or vertex) -> the exact fp64 rule.
                    float e_min, m_in;      // smallest (value - bound) and smallest (value + bound) over the edges
                    {
                        const float a0 = (ax1 - ax0) * (qy - ay0), b0 = (ay1 - ay0) * (qx - ax0);
                        const float a1 = (ax2 - ax1) * (qy - ay1), b1 = (ay2 - ay1) * (qx - ax1);
                        const float a2 = (ax0 - ax2) * (qy - ay2), b2 = (ay0 - ay2) * (qx - ax2);
                        const float c = 9.5367431640625e-07f;         // 2^-20
                        const float t0 = c * (fabsf(a0) + fabsf(b0)), t1 = c * (fabsf(a1) + fabsf(b1)), t2 = c * (fabsf(a2) + fabsf(b2));
                        const float e0 = a0 - b0, e1 = a1 - b1, e2 = a2 - b2;
                        e_min = fminf(fminf(e0 - t0, e1 - t1), e2 - t2);
                        m_in = fminf(fminf(e0 + t0, e1 + t1), e2 + t2);
                    }
                    int p = -1;
                    if (e_min > 0.f) p = 0;                          // strictly inside, beyond rounding doubt
                    else if (!(m_in < 0.f)) {                        // not clearly outside: exact evaluation
                        const double vx[3] = {(double)ax0, (double)ax1, (double)ax2}, vy[3] = {(double)ay0, (double)ay1, (double)ay2};
                        p = rule_eval_flat((double)qx, (double)qy, vx, vy);
                    }
                    if (p >= 0) atomicMin(&s_cell[ix * g.ny + iy], ((unsigned)p << 31) | (unsigned)(base + k));
                }
            }
        }
        __syncthreads();
        // ---- phase 2: output pixels
        for (int o = tid; o < n_out; o += nth) {
            const int l = g.ppx_shift >= 0 ? (o >> g.ppx_shift) : o / ppx, k = o - l * ppx;
            const int i = g.py_shift >= 0 ? (k >> g.py_shift) : k / g.py, j = k - i * g.py;
            const int bx = l / g.n_by, by = l - bx * g.n_by;
            int X = (bx + g.crop) * g.px + i, Y = (by + g.crop) * g.py + j;     // padded image coordinates
            if (g.flip_y) Y = g.padded_ny - 1 - Y;                               // airfoil_ds.py:80
            const int ix = X - g.pad_x0, iy = Y - g.pad_y0;
            int tri = -1;
            float v[3] = {0.f, 0.f, 0.f};
            bool masked = true;
            if (ix >= 0 && ix < g.nx && iy >= 0 && iy < g.ny) {
                const unsigned best = s_cell[ix * g.ny + iy];
                if (best != 0xffffffffu) {
                    tri = (int)(best & 0x7fffffffu);
                    int v0 = __ldg(cells + 3 * tri), v1 = __ldg(cells + 3 * tri + 1), v2 = __ldg(cells + 3 * tri + 2);
                    const float2 pa = s_pos[v0];
                    float2 pb = s_pos[v1], pc = s_pos[v2];
                    const double x0 = pa.x, y0 = pa.y;
                    const double cz = __dsub_rn(__dmul_rn(__dsub_rn((double)pb.x, x0), __dsub_rn((double)pc.y, y0)),
                                                __dmul_rn(__dsub_rn((double)pb.y, y0), __dsub_rn((double)pc.x, x0)));
                    if (cz < 0.0) { int s = v1; v1 = v2; v2 = s; const float2 q = pb; pb = pc; pc = q; }
                    // barycentric weights: the expressions of flg::cell_weights
                    const double qx = (double)s_ax[ix], qy = (double)s_ay[iy];
                    const double e1x = (double)pb.x - x0, e1y = (double)pb.y - y0;
                    const double e2x = (double)pc.x - x0, e2y = (double)pc.y - y0;
                    const double dx = qx - x0, dy = qy - y0;
                    const double d = e1x * e2y - e2x * e1y;
                    double w1 = 0.0, w2 = 0.0;
                    if (d != 0.0) {
                        w1 = (dx * e2y - e2x * dy) / d;
                        w2 = (e1x * dy - dx * e1y) / d;
                    }
                    const double w0 = 1.0 - w1 - w2;
                    float2 a0, a1, a2;
                    float p0, p1, p2;
                    if (stage_fields) { a0 = s_vel[v0]; a1 = s_vel[v1]; a2 = s_vel[v2]; p0 = s_prs[v0]; p1 = s_prs[v1]; p2 = s_prs[v2]; }
                    else { a0 = __ldg(vel + v0); a1 = __ldg(vel + v1); a2 = __ldg(vel + v2); p0 = __ldg(prs + v0); p1 = __ldg(prs + v1); p2 = __ldg(prs + v2); }
                    v[0] = (float)fma(w2, (double)a2.x, fma(w1, (double)a1.x, w0 * (double)a0.x));
                    v[1] = (float)fma(w2, (double)a2.y, fma(w1, (double)a1.y, w0 * (double)a0.y));
                    v[2] = (float)fma(w2, (double)p2, fma(w1, (double)p1, w0 * (double)p0));
                    masked = !finite_f(v[2]);                 // only the pressure mask is kept (simple_dataloader.py:114,119)
#pragma unroll
                    for (int c = 0; c < 3; ++c) if (!finite_f(v[c])) v[c] = 0.f;   // mesh_utils.py:89, per channel
                }
            }
            float* dst = states + (((size_t)f * n_patches + l) * 3) * ppx + k;
            const bool do_norm = !(flags & FL_NO_NORM) && !((flags & FL_MASK_AWARE_NORM) && masked);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float x = v[c];
                if (do_norm) x = __fdiv_rn(__fsub_rn(x, nc.mean[c]), nc.stdv[c]);
                fl_stg_stream1(dst + (size_t)c * ppx, x);
            }
            if (mask) mask[((size_t)f * n_patches + l) * ppx + k] = masked ? 1 : 0;
            if (tri_out) tri_out[((size_t)f * n_patches + l) * ppx + k] = tri;
        }
    }
    if (bad) atomicAdd(&status[0], bad);
}

// shared memory the rasterising form needs (bytes); *stage_fields = whether the node fields fit as well
size_t raster_smem_bytes(int nx, int ny, int n_nodes, int* stage_fields) {
    const size_t C = (size_t)nx * ny, n_pad = ((size_t)n_nodes + 1) & ~(size_t)1;
    size_t b = 4 * ((C + 3) & ~(size_t)3) + 8 * n_pad + 4 * (((size_t)nx + ny + 1) & ~(size_t)1);
    const size_t with_fields = b + 8 * n_pad + 4 * (size_t)n_nodes;
    *stage_fields = with_fields <= 227 * 1024 ? 1 : 0;
    return *stage_fields ? with_fields : b;
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
    FL_REQUIRE(px > 0 && py > 0 && (long)px * py <= (1L << 20) && crop_patches >= 0, FL_E_ARG,
               "fl_dyn_interp_patchify: patch of %dx%d pixels unsupported", px, py);
    FL_REQUIRE((h_mean && h_std) || (flags & FL_NO_NORM), FL_E_ARG, "fl_dyn_interp_patchify: mean/std missing");
    FL_REQUIRE(((uintptr_t)d_workspace & 255) == 0 && ((uintptr_t)d_velocity & 7) == 0 && ((uintptr_t)d_pos & 7) == 0, FL_E_ALIGN,
               "fl_dyn_interp_patchify: workspace must be 256-byte aligned, positions and velocity 8-byte aligned");
    DynGeom g;
    const int pad_x = ((-nx) % px + px) % px, pad_y = ((-ny) % py + py) % py;   // simple_dataloader.py:140-141
    g.nx = nx; g.ny = ny; g.px = px; g.py = py; g.crop = crop_patches;
    g.n_bx = (nx + pad_x) / px - 2 * crop_patches; g.n_by = (ny + pad_y) / py - 2 * crop_patches;
    g.pad_x0 = pad_x / 2; g.pad_y0 = pad_y / 2; g.padded_ny = ny + pad_y; g.flip_y = (flags & FL_FLIP_Y) ? 1 : 0;
    FL_REQUIRE(g.n_bx > 0 && g.n_by > 0, FL_E_ARG, "fl_dyn_interp_patchify: no patches left after cropping");
    g.ppx_shift = g.py_shift = -1;
    for (int b = 0; b < 16; ++b) { if ((1 << b) == px * py) g.ppx_shift = b; if ((1 << b) == py) g.py_shift = b; }
    NormConst nc;
    for (int c = 0; c < 3; ++c) { nc.mean[c] = h_mean ? h_mean[c] : 0.f; nc.stdv[c] = h_std ? h_std[c] : 1.f; }
    cudaStream_t st = (cudaStream_t)stream;
    // ---- rasterising form: the grid's triangle-id array and the frame's nodes fit in shared memory ----
    int stage_fields = 0;
    const size_t smem = raster_smem_bytes(nx, ny, n_nodes, &stage_fields);
    if (smem <= 227 * 1024 && n_cells < 0x7fffffff && !(flags & FL_FORCE_GATHER)) {
        static FlOncePerDevice attr;
        if (attr.first_use()) FL_CUDA(cudaFuncSetAttribute(k_dyn_raster_interp, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        FL_CUDA(cudaMemsetAsync(d_status, 0, 2 * sizeof(int32_t), st));
        const int grid = n_frames < FL_SM_COUNT ? n_frames : FL_SM_COUNT;      // one persistent CTA per SM
        k_dyn_raster_interp<<<grid, RS_THREADS, smem, st>>>(d_pos, d_cells, d_velocity, d_pressure, n_frames, n_nodes, n_cells,
                                                            d_grid_ax, d_grid_ay, g, nc, flags, stage_fields, d_states, d_mask,
                                                            d_tri, d_status);
        FL_LAUNCH_CHECK();
        return FL_OK;
    }
    // ---- binned form (any grid size) ----
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
        k_dyn_locate_interp<<<grid, px * py >= 256 ? 256 : (px * py + 31) / 32 * 32, 0, st>>>(d_pos, d_velocity, d_pressure, n_nodes, n_cells, w.tri_v, w.bin_start,
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
