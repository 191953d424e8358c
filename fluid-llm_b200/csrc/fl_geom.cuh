// fl_geom.cuh -- device helpers shared by the point-location and interpolation kernels, and the layout of the
// triangle-binning workspace (fl_locate.cu builds it, fl_locate.cu / fl_dynamic.cu read it).
#pragma once
#include "fl_common.cuh"

namespace flg {

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

// rule_eval without early exits (same result): for warps whose lanes hold unrelated (point, triangle) pairs
__device__ __forceinline__ int rule_eval_flat(double qx, double qy, const double* vx, const double* vy) {
    const bool on_vertex = (qx == vx[0] && qy == vy[0]) || (qx == vx[1] && qy == vy[1]) || (qx == vx[2] && qy == vy[2]);
    bool reject = false, below_on_edge = false;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int k1 = (k + 1) % 3;
        const bool end_right = (vx[k1] == vx[k]) ? (vy[k1] > vy[k]) : (vx[k1] > vx[k]);
        const double lx = end_right ? vx[k] : vx[k1], ly = end_right ? vy[k] : vy[k1];
        const double rx = end_right ? vx[k1] : vx[k], ry = end_right ? vy[k1] : vy[k];
        const double s = orient(qx, qy, lx, ly, rx, ry);
        reject |= end_right ? !(s <= 0.0) : !(s >= 0.0);
        below_on_edge |= !end_right && s == 0.0;
    }
    return on_vertex ? 0 : (reject ? -1 : (below_on_edge ? 1 : 0));
}

// triangle id of the grid point (qx, qy): the rule over the candidates of its bin; -1 = outside the mesh
__device__ __forceinline__ int locate_in_bin(const float* __restrict__ pos, const int* __restrict__ tri_v,
                                             const int* __restrict__ items, int kbeg, int kend, double qx, double qy) {
    unsigned best = 0xffffffffu;  // (prio << 31) | tri
    for (int k = kbeg; k < kend; ++k) {
        int t = items[k];
        double vx[3], vy[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { int v = tri_v[3 * t + j]; vx[j] = pos[2 * v]; vy[j] = pos[2 * v + 1]; }
        int p = rule_eval(qx, qy, vx, vy);
        if (p >= 0) best = min(best, ((unsigned)p << 31) | (unsigned)t);
    }
    return best == 0xffffffffu ? -1 : (int)(best & 0x7fffffffu);
}

// barycentric weights of vertices 1 and 2 (fp64); a degenerate triangle gets (0, 0)
__device__ __forceinline__ void cell_weights(const float* __restrict__ pos, int v0, int v1, int v2, double qx, double qy,
                                             double& w1, double& w2) {
    double x0 = pos[2 * v0], y0 = pos[2 * v0 + 1];
    double e1x = (double)pos[2 * v1] - x0, e1y = (double)pos[2 * v1 + 1] - y0;
    double e2x = (double)pos[2 * v2] - x0, e2y = (double)pos[2 * v2 + 1] - y0;
    double dx = qx - x0, dy = qy - y0;
    double d = e1x * e2y - e2x * e1y;
    w1 = 0.0; w2 = 0.0;
    if (d != 0.0) {
        w1 = (dx * e2y - e2x * dy) / d;
        w2 = (e1x * dy - dx * e1y) / d;
    }
}

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

// ---- triangle-binning workspace: per frame (= per mesh) arrays, frame-major ----------------------------------------
struct BinWs {
    int* tri_v;             // [frames][3 * n_cells]  vertex ids after the counter-clockwise fix
    TriRange* tri_range;    // [frames][n_cells]
    int* bin_count;         // [frames][nbins + 1]
    int* bin_start;         // [frames][nbins + 1]    exclusive scan, total in [nbins]
    int* cursor;            // [frames][nbins + 1]
    int* flags;             // [0] triangles with a bad node id (sum), [1] largest per-frame item count
    int* items;             // [frames][capacity]     triangle ids, bin after bin
    int nbx, nby, nbins, capacity, n_frames, n_cells;
    size_t zero_bytes;      // bin_count, bin_start and cursor, contiguous: cleared before every binning round
};

static inline size_t bin_ws_fixed_bytes(int n_frames, int n_cells, int nbins) {
    size_t b = 0;
    b += fl_align_up(sizeof(int) * 3 * (size_t)n_cells * n_frames, 256);
    b += fl_align_up(sizeof(TriRange) * (size_t)n_cells * n_frames, 256);
    b += 3 * fl_align_up(sizeof(int) * ((size_t)nbins + 1) * n_frames, 256);
    b += 256;  // flags
    return b;
}

// carve `bytes` at `ws` into the arrays above; what is left after the fixed part is split evenly into the frames' item stores
static inline bool bin_ws_carve(void* ws, size_t bytes, int n_frames, int n_cells, int nx, int ny, BinWs* w) {
    w->nbx = (nx + BIN - 1) / BIN; w->nby = (ny + BIN - 1) / BIN; w->nbins = w->nbx * w->nby;
    w->n_frames = n_frames; w->n_cells = n_cells;
    const size_t fixed = bin_ws_fixed_bytes(n_frames, n_cells, w->nbins);
    if (bytes < fixed + 1024 * (size_t)n_frames) return false;
    char* p = (char*)ws;
    w->tri_v = (int*)p; p += fl_align_up(sizeof(int) * 3 * (size_t)n_cells * n_frames, 256);
    w->tri_range = (TriRange*)p; p += fl_align_up(sizeof(TriRange) * (size_t)n_cells * n_frames, 256);
    w->bin_count = (int*)p; p += fl_align_up(sizeof(int) * ((size_t)w->nbins + 1) * n_frames, 256);
    w->bin_start = (int*)p; p += fl_align_up(sizeof(int) * ((size_t)w->nbins + 1) * n_frames, 256);
    w->cursor = (int*)p; p += fl_align_up(sizeof(int) * ((size_t)w->nbins + 1) * n_frames, 256);
    w->zero_bytes = (size_t)(p - (char*)w->bin_count);
    w->flags = (int*)p; p += 256;
    w->items = (int*)p;
    size_t cap = (bytes - (size_t)(p - (char*)ws)) / sizeof(int) / (size_t)n_frames;
    w->capacity = cap > 0x7fffffff ? 0x7fffffff : (int)cap;
    return true;
}

// fl_locate.cu: bin the triangles of n_frames meshes (frame f: pos + f * pos_stride floats, cells + f * 3 * n_cells ints) on
// `st`; no host synchronisation.  w.flags is zeroed first unless keep_flags.
int bin_frames(const float* d_pos, size_t pos_stride, const int* d_cells, int n_nodes, const float* d_ax, const float* d_ay,
               int nx, int ny, const BinWs& w, bool keep_flags, cudaStream_t st);

}  // namespace flg
