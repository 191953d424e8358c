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
#include "fl_interp.cuh"
#include <string.h>
#include <stdlib.h>
#include <math.h>

using flg::NormConst;
using flg::interp3;
using flg::finite_f;
using fli::StagedConst;
using fli::norm_fast;
using fli::norm_fast2;
using fli::pack2;

static thread_local const char* g_last_kernel = "";     // what fl_interp_patchify last launched on this thread

namespace {

// Generic kernel: one CTA = one patch x a chunk of frames of one trajectory; a thread takes pixels k, k + blockDim.x, ... of
// the patch (one pixel each for patches of up to 256 pixels), so ANY patch size works.
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
    const bool mask_aware = flags & FL_MASK_AWARE_NORM, no_norm = flags & FL_NO_NORM;
    for (int k = threadIdx.x; k < ppx; k += blockDim.x) {
    const size_t o = (size_t)l * ppx + k;
    const FlCellIdx id = tr.d_idx[o];
    const FlCellW w = tr.d_w[o];
    const double w0 = 1.0 - w.w1 - w.w2;
    const bool outside = id.tri < 0;
    for (int f = f0; f < f1; ++f) {
        const size_t t = (size_t)tr.t0 + (size_t)f * tr.interval;
        float v[3] = {0.f, 0.f, 0.f};
        bool masked = outside;
        if (!outside) {
            interp3(tr.d_velocity + t * tr.vel_stride, tr.d_pressure + t * tr.prs_stride, id, w0, w.w1, w.w2, v);
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


// ------------------------------------------------------------------------------------------
// Staged kernel (the fast path).  One persistent 512-thread CTA per SM; a work item is TF
// consecutive selected frames of one trajectory.
//   staging   coalesced 128-bit streaming loads of 4 nodes' (u,v) pairs and pressures are written to
//             shared memory as four 16-byte node records with all three values ALREADY widened to fp64 (their
//             high words + one word holding the top byte of each low word: a widened float has no other low
//             bits): conversions happen once per NODE instead of nine times per PIXEL, one 128-bit gather
//             fetches a vertex and three PRMTs rebuild its doubles.
//             Records sit at Morton-ordered slots (d_node_slot); the same pass scans for non-finite /
//             huge values.  (A TMA bulk-copy staging of the frames in their global layout was measured
//             slower: twice the gather instructions and three more conversions per pixel.)
//   compute   a warp owns 128 consecutive output pixels, lane i handles pixels p(i), p(i)+32, +64, +96
//             (p = a 2 x 4-block permutation inside each 32-pixel group), so every gather instruction
//             covers 32 adjacent pixels.  The four table records of a thread are read once (coalesced,
//             L2-resident) and reused for all TF frames.
//   output    128 B per warp store (streaming, no L1 allocate); each (frame, patch, channel) block of
//             1 KB is written whole by two warps.
// HBM traffic: 12 N + 12 P bytes per frame (compulsory); L2 -> SM adds 24 P / TF (table: compact 8-byte index records +
// two fp64 weights per pixel; 32 P / TF with FlCellIdx records).
// Arithmetic per pixel-frame: 9 PRMT, 3 DMUL + 6 DFMA, 3 f64->f32, then (v - mean) / std as
// a reciprocal multiply with one Markstein correction step on the packed fp32 pipe (bit-identical to
// IEEE division for the ranges checked on the host and in the staging scan; anything else takes the
// checked path).  What bounds it today (profiles/README.md): the L1/shared LSU data pipe at ~80 %.
// ------------------------------------------------------------------------------------------
constexpr int ST_THREADS = 512;
constexpr int NP = 4;        // pixels per thread

// wide node record {bits, hi(u), hi(p), hi(v)}: the three values widened to fp64 at staging; byte k of `bits` = the top byte of the
// k-th value's low word (u, v, p), whose other 24 bits are zero for a widened float (normal, denormal, inf or NaN alike)
__device__ __forceinline__ float4 make_wide(float u, float v, float p) {
    const double du = (double)u, dv = (double)v, dp = (double)p;
    const uint32_t bits = ((uint32_t)__double2loint(du) >> 24) | (((uint32_t)__double2loint(dv) >> 24) << 8) |
                          (((uint32_t)__double2loint(dp) >> 24) << 16);
    return make_float4(__uint_as_float(bits), __int_as_float(__double2hiint(du)), __int_as_float(__double2hiint(dp)),
                       __int_as_float(__double2hiint(dv)));
}
// (32-bit shared address: with a warp-uniform base the add folds into LDS.128 [R + UR])
__device__ __forceinline__ void load_wide(uint32_t addr, double& u, double& v, double& p) {
    unsigned long long A, B;        // A = {bits, hi(u)}, B = {hi(p), hi(v)}
    asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(A), "=l"(B) : "r"(addr));
    const uint32_t bits = (uint32_t)A;
    p = __hiloint2double((int)(uint32_t)B, (int)__byte_perm(bits, 0u, 0x2444));
    v = __longlong_as_double((long long)((B & 0xffffffff00000000ull) | __byte_perm(bits, 0u, 0x1444)));
    u = __longlong_as_double((long long)((A & 0xffffffff00000000ull) | __byte_perm(bits, 0u, 0x0444)));
}

template <bool CHECKED, int WIDE, bool IDX16 = false>
__device__ __forceinline__ void staged_item(const FlTraj& tr, const unsigned char* __restrict__ s_nodes, uint32_t s_base, int slot_b,
                                            int fbeg, int nf, int n_patches, int ppx, int ppx_shift, const StagedConst& sc,
                                            unsigned flags) {
    const bool mask_aware = flags & FL_MASK_AWARE_NORM, no_norm = flags & FL_NO_NORM;
    const int nchunks = n_patches * ppx / (32 * NP);
    const int lane_id = threadIdx.x & 31, warps = blockDim.x >> 5;
    // pixel of this lane inside a group of 32 adjacent pixels (2 patch rows x 16): the 8 lanes of a quarter-warp
    // (one LDS.128 wavefront) take a compact 2 x 4 block, which touches fewer distinct nodes than a 1 x 8 strip;
    // any permutation inside the group keeps every warp store one contiguous 128-byte line
    const bool permuted = ppx_shift >= 0 && (ppx & 15) == 0;
    const int lane = permuted ? ((lane_id >> 2) & 1) * 16 + (lane_id >> 3) * 4 + (lane_id & 3) : lane_id;
    const FlCellIdx* idx_tab = tr.d_idx_slot ? tr.d_idx_slot : tr.d_idx;   // node ids as shared-memory slots
    const size_t frame_out = (size_t)n_patches * 3 * ppx;
    unsigned long long nm[3], ns[3], rc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { nm[c] = pack2(-sc.mean[c], -sc.mean[c]); ns[c] = pack2(-sc.stdv[c], -sc.stdv[c]); rc[c] = pack2(sc.rcp[c], sc.rcp[c]); }
    for (int ch = threadIdx.x >> 5; ch < nchunks; ch += warps) {
        const int o = ch * 32 * NP + lane;          // first pixel of this lane; the others are +32*r
        uint32_t ov[NP][3];     // byte offset of each vertex's 16-byte node record {u f32, v f32, p f64} in a frame slot
        double w0[NP], w1[NP], w2[NP];
        unsigned mbits = 0;     // byte r = 1 if pixel r is outside the mesh
#pragma unroll
        for (int r = 0; r < NP; ++r) {
            int4 id;
            if (IDX16) {       // compact table: {v0 | v1 << 16, v2 | outside << 16}, node slots below 65536
                const uint2 q = __ldg((const uint2*)idx_tab + o + 32 * r);
                id = make_int4((int)(q.x & 0xffffu), (int)(q.x >> 16), (int)(q.y & 0xffffu), (q.y >> 16) ? -1 : 0);
            } else {
                id = __ldg((const int4*)idx_tab + o + 32 * r);
            }
            const double2 ww = __ldg((const double2*)tr.d_w + o + 32 * r);
            const bool out = id.w < 0;
            mbits |= out ? (1u << (8 * r)) : 0u;
            w1[r] = out ? 0.0 : ww.x;
            w2[r] = out ? 0.0 : ww.y;
            w0[r] = out ? 0.0 : 1.0 - ww.x - ww.y;
            ov[r][0] = out ? 0u : (uint32_t)id.x * 16u;
            ov[r][1] = out ? 0u : (uint32_t)id.y * 16u;
            ov[r][2] = out ? 0u : (uint32_t)id.z * 16u;
        }
        // mask bytes of the chunk in pixel order: lane j will store the four bytes of pixels 4j..4j+3 with one 32-bit
        // store per frame (the mask is static on the unchecked path); pixel 32r+q lives in byte r of lane inv(q)'s mbits
        unsigned mword = 0;
        if (!CHECKED && NP == 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int q = 4 * (lane_id & 7) + i;                                  // position inside the 32-pixel group
                const int src = !permuted ? q : (((q >> 2) & 3) << 3) | (((q >> 4) & 1) << 2) | (q & 3);   // inverse of the lane permutation
                const unsigned m = __shfl_sync(0xffffffffu, mbits, src);
                mword |= ((m >> (8 * (lane_id >> 3))) & 1u) << (8 * i);
            }
        }
        const int l = ppx_shift >= 0 ? (o >> ppx_shift) : o / ppx, k = o - l * ppx;   // the chunk never straddles a patch
        float* dst = tr.d_states + ((size_t)fbeg * n_patches + l) * 3 * ppx + k;
        uint8_t* mdst = tr.d_mask ? tr.d_mask + ((size_t)fbeg * n_patches + l) * ppx + k : nullptr;
        unsigned* mdst4 = tr.d_mask ? (unsigned*)(tr.d_mask + ((size_t)fbeg * n_patches + l) * ppx + (k - lane)) + lane_id : nullptr;
        const unsigned char* nb = s_nodes;    // loop-carried uniform base: folds into LDS.128 [R + UR]
#pragma unroll 1
        for (int f = 0; f < nf; ++f, nb += slot_b) {
            const uint32_t nb32 = s_base + (uint32_t)f * (uint32_t)slot_b;      // warp-uniform: stays in a uniform register
            float res[3][NP];
            unsigned fm = mbits;
#pragma unroll
            for (int r = 0; r < NP; ++r) {
                double u0, v0, p0, u1, v1, p1, u2, v2, p2;
                if (WIDE == 1) {     // records {bits, hi(u), hi(p), hi(v)}: the doubles are rebuilt with PRMTs, no conversion
                    load_wide(nb32 + ov[r][0], u0, v0, p0);
                    load_wide(nb32 + ov[r][1], u1, v1, p1);
                    load_wide(nb32 + ov[r][2], u2, v2, p2);
                } else {             // records {u f32, v f32, p f64}
                    const float4 a0 = *reinterpret_cast<const float4*>(nb + ov[r][0]);   // one 128-bit gather per vertex
                    const float4 a1 = *reinterpret_cast<const float4*>(nb + ov[r][1]);
                    const float4 a2 = *reinterpret_cast<const float4*>(nb + ov[r][2]);
                    p0 = __hiloint2double(__float_as_int(a0.w), __float_as_int(a0.z));
                    p1 = __hiloint2double(__float_as_int(a1.w), __float_as_int(a1.z));
                    p2 = __hiloint2double(__float_as_int(a2.w), __float_as_int(a2.z));
                    u0 = (double)a0.x; v0 = (double)a0.y; u1 = (double)a1.x; v1 = (double)a1.y; u2 = (double)a2.x; v2 = (double)a2.y;
                }
                res[0][r] = (float)fma(w2[r], u2, fma(w1[r], u1, w0[r] * u0));
                res[1][r] = (float)fma(w2[r], v2, fma(w1[r], v1, w0[r] * v0));
                res[2][r] = (float)fma(w2[r], p2, fma(w1[r], p1, w0[r] * p0));
                if (CHECKED) {
                    if (!finite_f(res[2][r])) fm |= 1u << (8 * r);           // pressure mask only
#pragma unroll
                    for (int c = 0; c < 3; ++c) if (!finite_f(res[c][r])) res[c][r] = 0.f;
                }
            }
            if (mbits) {       // outside the mesh the weights are 0 and the sum is +-0: the reference stores +0.0 (mesh_utils.py:89)
#pragma unroll
                for (int r = 0; r < NP; ++r)
                    if ((mbits >> (8 * r)) & 1u) { res[0][r] = 0.f; res[1][r] = 0.f; res[2][r] = 0.f; }
            }
            if (!no_norm) {
                if (CHECKED) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int r = 0; r < NP; ++r) {
                            const float x = res[c][r];
                            const float y = __fdiv_rn(__fsub_rn(x, sc.mean[c]), sc.stdv[c]);
                            res[c][r] = (mask_aware && ((fm >> (8 * r)) & 1u)) ? x : y;   // airfoil_ds.py:241-242
                        }
                } else if (mask_aware && fm) {      // rare: pixels on the mesh boundary / padding stay raw
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int r = 0; r < NP; ++r)
                            if (!((fm >> (8 * r)) & 1u)) res[c][r] = norm_fast(res[c][r], sc.mean[c], sc.stdv[c], sc.rcp[c]);
                } else {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int r = 0; r < NP; r += 2) norm_fast2(res[c][r], res[c][r + 1], nm[c], ns[c], rc[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int r = 0; r < NP; ++r) fl_stg_stream1(dst + (size_t)c * ppx + 32 * r, res[c][r]);   // 128 B per warp store
            dst += frame_out;
            if (mdst) {
                if (!CHECKED && NP == 4) {
                    *mdst4 = mword;                                   // 128 contiguous bytes per warp
                    mdst4 += (size_t)n_patches * ppx / 4;
                } else {
#pragma unroll
                    for (int r = 0; r < NP; ++r) mdst[32 * r] = (uint8_t)((fm >> (8 * r)) & 1u);
                    mdst += (size_t)n_patches * ppx;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(ST_THREADS, 1)
k_interp_patchify_staged(const FlTraj* __restrict__ trajs, int n_items, int groups_per_traj, int TF, int n_patches,
                         int ppx, int ppx_shift, int slot_nodes, StagedConst sc, unsigned flags, int wide) {
    extern __shared__ __align__(128) unsigned char fl_smem[];
    float4* s_nodes = (float4*)fl_smem;          // [TF][slot_nodes] records {u, v, p as fp64}
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int j = item / groups_per_traj, g = item - j * groups_per_traj;
        const FlTraj tr = trajs[j];
        const int fbeg = g * TF;
        if (fbeg >= tr.n_frames) continue;                     // uniform across the CTA
        const int nf = min(TF, tr.n_frames - fbeg);
        // staging: coalesced 128-bit loads of 4 nodes' (u,v) pairs and pressures -> four 16-byte node
        // records, pressure converted to fp64 once per node; scanned on the way
        float nanacc = 0.f, amax = 0.f;
        auto scan4 = [&](const float4 v) {     // a non-finite or huge value sends the whole item down the checked path
            nanacc = fmaf(v.x, 0.f, fmaf(v.y, 0.f, fmaf(v.z, 0.f, fmaf(v.w, 0.f, nanacc))));
            amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fmaxf(fabsf(v.z), fabsf(v.w)), amax));
        };
        const int nq = tr.prs_stride / 4;      // groups of 4 nodes per frame (pad nodes are zero-filled)
        for (int i = threadIdx.x; i < nf * nq; i += blockDim.x) {
            const int f = i / nq, k4 = i - f * nq;
            const size_t t = (size_t)tr.t0 + (size_t)(fbeg + f) * tr.interval;
            const float4* vsrc = (const float4*)(tr.d_velocity + t * tr.vel_stride) + 2 * k4;
            const float4 va = fl_ldg_stream4(vsrc), vb = fl_ldg_stream4(vsrc + 1);
            const float4 pp = fl_ldg_stream4((const float4*)(tr.d_pressure + t * tr.prs_stride) + k4);
            scan4(va); scan4(vb); scan4(pp);
            float4* d = s_nodes + (size_t)f * slot_nodes;
            int4 sl = make_int4(4 * k4, 4 * k4 + 1, 4 * k4 + 2, 4 * k4 + 3);
            if (tr.d_node_slot) sl = __ldg((const int4*)tr.d_node_slot + k4);      // spatially sorted slots
            if (wide) {
                d[sl.x] = make_wide(va.x, va.y, pp.x);
                d[sl.y] = make_wide(va.z, va.w, pp.y);
                d[sl.z] = make_wide(vb.x, vb.y, pp.z);
                d[sl.w] = make_wide(vb.z, vb.w, pp.w);
            } else {
                const double p0 = (double)pp.x, p1 = (double)pp.y, p2 = (double)pp.z, p3 = (double)pp.w;
                d[sl.x] = make_float4(va.x, va.y, __int_as_float(__double2loint(p0)), __int_as_float(__double2hiint(p0)));
                d[sl.y] = make_float4(va.z, va.w, __int_as_float(__double2loint(p1)), __int_as_float(__double2hiint(p1)));
                d[sl.z] = make_float4(vb.x, vb.y, __int_as_float(__double2loint(p2)), __int_as_float(__double2hiint(p2)));
                d[sl.w] = make_float4(vb.z, vb.w, __int_as_float(__double2loint(p3)), __int_as_float(__double2hiint(p3)));
            }
        }
        const int bad = __syncthreads_or(!(nanacc == 0.f) || amax > 1.0e30f || !sc.fast_div);
        const unsigned char* snb = (const unsigned char*)s_nodes;
        const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(fl_smem);
        if (tr.idx_slot_format == 1) {            // d_idx_slot holds the compact 8-byte records (wide node records only)
            if (bad) staged_item<true, 1, true>(tr, snb, s_base, slot_nodes * 16, fbeg, nf, n_patches, ppx, ppx_shift, sc, flags);
            else staged_item<false, 1, true>(tr, snb, s_base, slot_nodes * 16, fbeg, nf, n_patches, ppx, ppx_shift, sc, flags);
        } else if (wide) {
            if (bad) staged_item<true, 1>(tr, snb, s_base, slot_nodes * 16, fbeg, nf, n_patches, ppx, ppx_shift, sc, flags);
            else staged_item<false, 1>(tr, snb, s_base, slot_nodes * 16, fbeg, nf, n_patches, ppx, ppx_shift, sc, flags);
        } else {
            if (bad) staged_item<true, 0>(tr, snb, s_base, slot_nodes * 16, fbeg, nf, n_patches, ppx, ppx_shift, sc, flags);
            else staged_item<false, 0>(tr, snb, s_base, slot_nodes * 16, fbeg, nf, n_patches, ppx, ppx_shift, sc, flags);
        }
        __syncthreads();   // every gather of this item is done before the next item's staging lands
    }
}

int launch_interp(const FlTraj* d_trajs, const FlTraj* h_trajs, int n_traj, int max_frames, int n_patches, int px, int py,
                  const float* h_mean, const float* h_std, unsigned flags, cudaStream_t st) {
    NormConst nc;
    for (int c = 0; c < 3; ++c) { nc.mean[c] = h_mean ? h_mean[c] : 0.f; nc.stdv[c] = h_std ? h_std[c] : 1.f; }
    const int ppx = px * py;
    // ---- tiled path (fl_tiled.cu): the descriptors carry a tile plan ----
    if (h_trajs && !(flags & (FL_FORCE_GATHER | FL_FORCE_STAGED))) {
        StagedConst sc;
        for (int c = 0; c < 3; ++c) { sc.mean[c] = nc.mean[c]; sc.stdv[c] = nc.stdv[c]; sc.rcp[c] = 1.0f / nc.stdv[c]; }
        sc.fast_div = (flags & FL_NO_NORM) ? 1 : (fli::fast_div_ok(nc.mean, nc.stdv) ? 1 : 0);
        // the frame-ring kernel (fl_ring.cu) is an experiment that measured slower than the staged kernel on B200
        // (profiles/README.md): it runs only when asked for (FL_FORCE_RING, or FLUIDGRID_RING=1 in the environment)
        const char* ring_env = getenv("FLUIDGRID_RING");
        const bool want_ring = (flags & FL_FORCE_RING) || (ring_env && atoi(ring_env) == 1);
        int rc = want_ring && !(flags & FL_FORCE_TILED) ? fli::launch_ring(d_trajs, h_trajs, n_traj, max_frames, n_patches, px, py, sc, flags, st) : 1;
        if (rc == FL_OK) g_last_kernel = "k_interp_patchify_ring";
        if (rc != 1) return rc;          // 1 = not asked for / no tile plan / frames too large for the ring
        rc = fli::launch_tiled(d_trajs, h_trajs, n_traj, max_frames, n_patches, px, py, sc, flags, st);
        if (rc == FL_OK) g_last_kernel = "k_interp_patchify_tiled";
        if (rc != 1) return rc;          // 1 = no tile plan / does not fit: fall through
    }
    // ---- staged path: needs the host copy of the descriptors to check strides / alignment ----
    if (h_trajs && ppx % 128 == 0 && !(flags & FL_FORCE_GATHER)) {
        bool ok = true;
        int slot_vel = 0, slot_prs = 0;
        for (int i = 0; i < n_traj && ok; ++i) {
            const FlTraj& t = h_trajs[i];
            ok = t.vel_stride % 4 == 0 && t.prs_stride % 4 == 0 && t.vel_stride >= 2 * t.prs_stride && t.prs_stride >= t.n_nodes &&
                 ((uintptr_t)t.d_velocity % 16 == 0) && ((uintptr_t)t.d_pressure % 16 == 0) && ((uintptr_t)t.d_states % 16 == 0) &&
                 (t.d_mask == nullptr || (uintptr_t)t.d_mask % 4 == 0) && ((uintptr_t)t.d_node_slot % 16 == 0) &&
                 ((uintptr_t)t.d_idx_slot % 16 == 0);
            slot_vel = t.vel_stride > slot_vel ? t.vel_stride : slot_vel;
            slot_prs = t.prs_stride > slot_prs ? t.prs_stride : slot_prs;
        }
        const size_t frame_bytes = 16 * (size_t)slot_prs;           // one 16-byte record per node (incl. pad nodes)
        const size_t budget = 227 * 1024;
        int TF = ok && frame_bytes ? (int)(budget / frame_bytes) : 0;
        if (TF > 16) TF = 16;
        if (TF > max_frames) TF = max_frames;
        if (TF >= 1) {
            StagedConst sc;
            for (int c = 0; c < 3; ++c) { sc.mean[c] = nc.mean[c]; sc.stdv[c] = nc.stdv[c]; sc.rcp[c] = 1.0f / nc.stdv[c]; }
            sc.fast_div = (flags & FL_NO_NORM) ? 1 : (fli::fast_div_ok(nc.mean, nc.stdv) ? 1 : 0);
            const int gpt = (max_frames + TF - 1) / TF;
            const long n_items = (long)gpt * n_traj;
            const size_t smem = (size_t)TF * frame_bytes;
            static FlOncePerDevice attr;
            if (attr.first_use()) {
                FL_CUDA(cudaFuncSetAttribute(k_interp_patchify_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            }
            int ppx_shift = -1;
            for (int b = 0; b < 16; ++b) if ((1 << b) == ppx) ppx_shift = b;
            const int grid = n_items < FL_SM_COUNT ? (int)n_items : FL_SM_COUNT;      // one persistent CTA per SM
            int wide = 1;        // fully pre-widened records (round 2: +4 % on all three workloads); FLUIDGRID_WIDE=0 keeps {u, v f32, p f64}
            if (const char* e = getenv("FLUIDGRID_WIDE")) wide = atoi(e);
            for (int i = 0; i < n_traj; ++i) if (h_trajs[i].idx_slot_format == 1) wide = 1;      // compact table records go with the wide node records
            k_interp_patchify_staged<<<grid, ST_THREADS, smem, st>>>(d_trajs, (int)n_items, gpt, TF, n_patches, ppx, ppx_shift, slot_prs,
                                                                     sc, flags, wide);
            FL_LAUNCH_CHECK();
            g_last_kernel = "k_interp_patchify_staged";
            return FL_OK;
        }
    }
    // ---- gather path ----
    // enough CTAs to fill 148 SMs x 8 resident CTAs a few times over, but keep the per-thread
    // record reuse: at least 4 frames per CTA when there is that much work
    long ctas_per_frame = (long)n_patches * n_traj;
    int fpc = 16;
    while (fpc > 1 && ctas_per_frame * ((max_frames + fpc - 1) / fpc) < 4L * FL_SM_COUNT * 8) fpc >>= 1;
    dim3 grid(n_patches, (max_frames + fpc - 1) / fpc, n_traj);
    const int threads = ppx >= 256 ? 256 : (ppx + 31) / 32 * 32;
    k_interp_patchify_gather<<<grid, threads, 0, st>>>(d_trajs, n_patches, ppx, fpc, nc, flags);
    FL_LAUNCH_CHECK();
    g_last_kernel = "k_interp_patchify_gather";
    return FL_OK;
}

}  // namespace

// host: is (x - mean) / std safe for the reciprocal + Markstein path?
bool fli::fast_div_ok(const float* mean, const float* stdv) {
    for (int c = 0; c < 3; ++c) {
        float s = fabsf(stdv[c]), m = fabsf(mean[c]);
        uint32_t bits;
        memcpy(&bits, &s, 4);
        if (!(s >= 9.5367431640625e-07f && s <= 1048576.f)) return false;        // 2^-20 .. 2^20
        if ((bits & 0x007fffffu) == 0x007fffffu) return false;                     // Markstein's excluded significand
        if (!(m >= 9.094947017729282e-13f && m <= 1.152921504606847e18f)) return false;  // 2^-40 .. 2^60, non-zero
    }
    return true;
}

static int check_trajs(const char* who, const FlTraj* h_trajs, int n_traj, int* max_frames) {
    FL_REQUIRE(h_trajs, FL_E_ARG, "%s: null descriptor array", who);
    FL_REQUIRE(n_traj > 0 && n_traj <= 65535, FL_E_ARG, "%s: n_traj=%d out of range", who, n_traj);
    int mf = 0;
    for (int i = 0; i < n_traj; ++i) {
        const FlTraj& t = h_trajs[i];
        FL_REQUIRE(t.d_velocity && t.d_pressure && t.d_idx && t.d_w && t.d_states, FL_E_ARG,
                   "%s: trajectory %d has a null pointer", who, i);
        FL_REQUIRE(t.n_nodes > 0 && t.n_frames > 0 && t.t0 >= 0 && t.interval > 0, FL_E_ARG,
                   "%s: trajectory %d has bad sizes", who, i);
        FL_REQUIRE(t.vel_stride >= 2 * t.n_nodes && t.prs_stride >= t.n_nodes, FL_E_ARG,
                   "%s: trajectory %d: frame strides (%d, %d) smaller than the frame (%d nodes)", who, i, t.vel_stride,
                   t.prs_stride, t.n_nodes);
        FL_REQUIRE((uintptr_t)t.d_velocity % 8 == 0 && (uintptr_t)t.d_idx % 16 == 0 && (uintptr_t)t.d_w % 16 == 0 &&
                       t.vel_stride % 2 == 0,
                   FL_E_ALIGN, "%s: trajectory %d: velocity must be 8-byte aligned with an even stride, tables 16-byte aligned", who, i);
        FL_REQUIRE((long long)t.interval * t.vel_stride < 0x7fffffffLL, FL_E_ARG, "%s: trajectory %d: interval * vel_stride overflows", who, i);
        FL_REQUIRE((t.d_idx_slot == nullptr) == (t.d_node_slot == nullptr), FL_E_ARG,
                   "%s: trajectory %d: d_idx_slot and d_node_slot must be given together", who, i);
        FL_REQUIRE(t.idx_slot_format == 0 || (t.idx_slot_format == 1 && t.d_idx_slot && t.prs_stride <= 65536 && (uintptr_t)t.d_idx_slot % 8 == 0),
                   FL_E_ARG, "%s: trajectory %d: idx_slot_format must be 0, or 1 with compact d_idx_slot records and at most 65536 slots", who, i);
        mf = t.n_frames > mf ? t.n_frames : mf;
    }
    *max_frames = mf;
    return FL_OK;
}

extern "C" const char* fl_last_interp_kernel(void) { return g_last_kernel; }

__global__ void k_pack_idx16(const FlCellIdx* __restrict__ in, uint2* __restrict__ out, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const FlCellIdx r = in[i];
    const bool outside = r.tri < 0;
    out[i] = outside ? make_uint2(0u, 1u << 16)
                     : make_uint2((unsigned)r.v0 | ((unsigned)r.v1 << 16), (unsigned)r.v2);
}

extern "C" int fl_pack_idx16(const FlCellIdx* d_idx_slot, long n, int n_slots, void* d_out, void* stream) {
    FL_REQUIRE(d_idx_slot && d_out && n > 0, FL_E_ARG, "fl_pack_idx16: null pointer or n <= 0");
    FL_REQUIRE(n_slots > 0 && n_slots <= 65536, FL_E_ARG, "fl_pack_idx16: %d slots do not fit 16 bits", n_slots);
    FL_REQUIRE((uintptr_t)d_idx_slot % 16 == 0 && (uintptr_t)d_out % 8 == 0, FL_E_ALIGN, "fl_pack_idx16: unaligned buffer");
    FL_REQUIRE((n + 255) / 256 < 0x7fffffffL, FL_E_ARG, "fl_pack_idx16: too many records");
    k_pack_idx16<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_idx_slot, (uint2*)d_out, n);
    FL_LAUNCH_CHECK();
    return FL_OK;
}

extern "C" int fl_interp_patchify_dev(const FlTraj* d_trajs, const FlTraj* h_trajs, int n_traj, int n_patches, int px, int py,
                                      const float* h_mean, const float* h_std, unsigned flags, void* stream) {
    FL_REQUIRE(d_trajs, FL_E_ARG, "fl_interp_patchify_dev: null device descriptor array");
    int max_frames = 0;
    int rc = check_trajs("fl_interp_patchify_dev", h_trajs, n_traj, &max_frames);
    if (rc) return rc;
    FL_REQUIRE(n_patches > 0, FL_E_ARG, "fl_interp_patchify_dev: n_patches=%d", n_patches);
    FL_REQUIRE(px > 0 && py > 0 && (long)px * py <= (1L << 20), FL_E_ARG, "fl_interp_patchify_dev: patch of %dx%d pixels unsupported", px, py);
    FL_REQUIRE((h_mean && h_std) || (flags & FL_NO_NORM), FL_E_ARG, "fl_interp_patchify_dev: mean/std missing");
    return launch_interp(d_trajs, h_trajs, n_traj, max_frames, n_patches, px, py, h_mean, h_std, flags, (cudaStream_t)stream);
}

extern "C" int fl_interp_patchify(const FlTraj* h_trajs, int n_traj, int n_patches, int px, int py, const float* h_mean,
                                  const float* h_std, unsigned flags, void* stream) {
    int max_frames = 0;
    int rc = check_trajs("fl_interp_patchify", h_trajs, n_traj, &max_frames);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    FlTraj* d_trajs = nullptr;
    FL_CUDA(cudaMallocAsync((void**)&d_trajs, sizeof(FlTraj) * n_traj, st));
    FL_CUDA(cudaMemcpyAsync(d_trajs, h_trajs, sizeof(FlTraj) * n_traj, cudaMemcpyHostToDevice, st));
    rc = fl_interp_patchify_dev(d_trajs, h_trajs, n_traj, n_patches, px, py, h_mean, h_std, flags, stream);
    cudaFreeAsync(d_trajs, st);
    return rc;
}

// eagle/Dataloader/IMG_MGN.py:78-157 (DilResNet image loader): per frame three to_grid calls, optional crop of
// `crop` pixels per side, (v - mean) / std on EVERY pixel, frames stored channel-last [T][H][W][3].
__global__ void k_interp_frames(const FlCellIdx* __restrict__ idx, const FlCellW* __restrict__ wt, int nx, int ny, int crop,
                                const float* __restrict__ vel, const float* __restrict__ prs, int vel_stride, int prs_stride,
                                int t0, int interval, int n_frames, NormConst nc, int no_norm, float* __restrict__ states,
                                uint8_t* __restrict__ mask) {
    const int H = nx - 2 * crop, W = ny - 2 * crop;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= H * W) return;
    const int ix = p / W + crop, iy = p % W + crop;
    const FlCellIdx id = idx[ix * ny + iy];
    const FlCellW w = wt[ix * ny + iy];
    const double w0 = 1.0 - w.w1 - w.w2;
    for (int f = blockIdx.y; f < n_frames; f += gridDim.y) {
        const size_t t = (size_t)t0 + (size_t)f * interval;
        float v[3] = {0.f, 0.f, 0.f};
        bool masked = id.tri < 0;
        if (!masked) {
            interp3(vel + t * vel_stride, prs + t * prs_stride, id, w0, w.w1, w.w2, v);
            masked = !finite_f(v[2]);
#pragma unroll
            for (int c = 0; c < 3; ++c) if (!finite_f(v[c])) v[c] = 0.f;
        }
        float* dst = states + ((size_t)f * H * W + p) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) dst[c] = no_norm ? v[c] : __fdiv_rn(__fsub_rn(v[c], nc.mean[c]), nc.stdv[c]);
        if (mask) mask[(size_t)f * H * W + p] = masked ? 1 : 0;
    }
}

extern "C" int fl_interp_frames(const FlCellIdx* d_cell_idx, const FlCellW* d_cell_w, int nx, int ny, int crop,
                                const float* d_velocity, const float* d_pressure, int n_nodes, int vel_stride, int prs_stride,
                                int t0, int interval, int n_frames, const float* h_mean, const float* h_std, unsigned flags,
                                float* d_states, uint8_t* d_mask, void* stream) {
    FL_REQUIRE(d_cell_idx && d_cell_w && d_velocity && d_pressure && d_states, FL_E_ARG, "fl_interp_frames: null pointer");
    FL_REQUIRE(nx > 0 && ny > 0 && crop >= 0 && nx > 2 * crop && ny > 2 * crop && n_frames > 0 && t0 >= 0 && interval > 0 && n_nodes > 0,
               FL_E_ARG, "fl_interp_frames: bad sizes");
    FL_REQUIRE(vel_stride >= 2 * n_nodes && prs_stride >= n_nodes && vel_stride % 2 == 0 && (uintptr_t)d_velocity % 8 == 0, FL_E_ARG,
               "fl_interp_frames: bad frame strides / alignment");
    FL_REQUIRE((h_mean && h_std) || (flags & FL_NO_NORM), FL_E_ARG, "fl_interp_frames: mean/std missing");
    NormConst nc;
    for (int c = 0; c < 3; ++c) { nc.mean[c] = h_mean ? h_mean[c] : 0.f; nc.stdv[c] = h_std ? h_std[c] : 1.f; }
    const int HW = (nx - 2 * crop) * (ny - 2 * crop);
    dim3 grid((HW + 255) / 256, n_frames < 64 ? n_frames : 64);
    k_interp_frames<<<grid, 256, 0, (cudaStream_t)stream>>>(d_cell_idx, d_cell_w, nx, ny, crop, d_velocity, d_pressure, vel_stride,
                                                          prs_stride, t0, interval, n_frames, nc, (flags & FL_NO_NORM) ? 1 : 0, d_states,
                                                          d_mask);
    FL_LAUNCH_CHECK();
    return FL_OK;
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
