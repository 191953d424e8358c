// fl_tiled.cu -- the per-step kernel, tiled form: gather -> fp64 FMA -> fp32 -> mask -> normalise -> patchify for meshes
// of ANY size out of shared memory.
//
// Same work and the same arithmetic as k_interp_patchify_staged (fl_interp.cu; replaces simple_dataloader.py:104-152,
// 166-216 / airfoil_ds.py:71-139,216-244 of the reference), reorganised around what ncu showed to bound that kernel
// (profiles/README.md): the L1/shared data pipe carried the gathers AND 12 scalar global stores per 4 pixel-frames, the
// conversion unit 9 fp32->fp64 conversions per pixel-frame, and staging was serialised behind a CTA-wide barrier.
//
//   tiles     the output patches are split into tiles (runs of patches along a serpentine through the patch grid); a
//             tile's node list holds exactly the mesh nodes its pixels touch (FlTraj::d_tile_*), and the table carries
//             tile-local slots.  A work item = one tile x TF consecutive selected frames of one trajectory, so shared
//             memory holds TF x (nodes of ONE tile) records whatever the size of the mesh.
//   records   16 bytes per node and frame, ALREADY in fp64 form: the high words of (double)u, (double)v, (double)p and
//             one word with the three non-zero bits of each low word (a float widened to double has 29 zero bits at the
//             bottom).  A vertex costs one LDS.128 + three PRMT instead of one LDS.128 + two F2F.F64.F32: the conversion
//             unit only sees the three results per pixel-frame.
//   staging   double buffered and spread over the compute loop: while a warp works through the TF frames of a chunk,
//             each of its threads loads one node of the NEXT item per frame iteration (issued before the gathers,
//             written to the other buffer after the stores), so the global-load latency hides behind the arithmetic and
//             the only CTA-wide barrier is the buffer swap.
//   output    results go to a per-patch shared-memory tile ([3][px*py] floats + px*py mask bytes, conflict-free
//             STS.32) and leave as ONE bulk copy (cp.async.bulk shared -> global, 3 KB) per frame and patch, issued by
//             one thread of the warps that share the patch; two tiles per patch group, so the copy of frame f overlaps
//             the arithmetic of frame f+1.  No global store goes through the LSU.
// HBM traffic per frame: 12 P (+ P mask) written, 12 x (sum of the tiles' node counts) read (= 12 N plus the halo
// nodes shared by neighbouring tiles, most of which hit L2 because the tiles of a frame group run at the same time).
#include "fl_interp.cuh"
#include <stdlib.h>

using flg::finite_f;
using fli::StagedConst;
using fli::norm_fast;
using fli::norm_fast2;
using fli::pack2;

namespace {

constexpr int TL_THREADS = 512;
constexpr int TL_WARPS = TL_THREADS / 32;
constexpr int NP = 4;                        // pixels per thread and chunk; a chunk = 128 consecutive output pixels
constexpr int SMEM_TOTAL = 227 * 1024;
constexpr int ITEM_WORDS = 32;               // decoded work item kept in shared memory (two of them)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void group_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts32u(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 ldg_stream2(const float* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// node record: {hi(p), hi(u), low-word bits, hi(v)}; the three bytes of `bits` hold the top byte of the low words
__device__ __forceinline__ uint4 make_record(float u, float v, float p) {
    const double du = (double)u, dv = (double)v, dp = (double)p;
    const uint32_t bits = ((uint32_t)__double2loint(du) >> 24) | (((uint32_t)__double2loint(dv) >> 24) << 8) |
                          (((uint32_t)__double2loint(dp) >> 24) << 16);
    return make_uint4((uint32_t)__double2hiint(dp), (uint32_t)__double2hiint(du), bits, (uint32_t)__double2hiint(dv));
}
__device__ __forceinline__ double rec_u(const uint4& a) { return __hiloint2double((int)a.y, (int)__byte_perm(a.z, 0u, 0x0444)); }
__device__ __forceinline__ double rec_v(const uint4& a) { return __hiloint2double((int)a.w, (int)__byte_perm(a.z, 0u, 0x1444)); }
__device__ __forceinline__ double rec_p(const uint4& a) { return __hiloint2double((int)a.x, (int)__byte_perm(a.z, 0u, 0x2444)); }

// A decoded work item (tile x frame group of one trajectory), written to shared memory by one warp.
struct Item {
    const float* vel;        // node fields at the item's first selected frame
    const float* prs;
    int vstep, pstep;        // floats between selected frames (interval * stride)
    const int* nodes;        // the tile's node list
    const int* patches;      // the tile's patch ids
    const int4* idx;         // table with tile-local slots * 16
    const double2* w;
    float* states;           // output at the item's first frame
    uint8_t* mask;           // or NULL
    int n_nodes, n_patches, nf, valid;
};
static_assert(sizeof(Item) <= ITEM_WORDS * 4, "Item does not fit its shared-memory slot");

struct TiledArgs {
    const FlTraj* trajs;
    int n_items, groups, n_tiles, TF, n_patches, slot_rec;   // slot_rec: records per staged frame
    StagedConst sc;
    unsigned flags;
};

__device__ __forceinline__ void decode_item(const TiledArgs& a, int item, int ppx, Item* out) {
    Item it;
    it.valid = item < a.n_items;
    if (!it.valid) { it.nf = 0; it.n_nodes = 0; it.n_patches = 0; *out = it; return; }
    const int tile = item % a.n_tiles;
    const int jg = item / a.n_tiles;
    const int g = jg % a.groups, j = jg / a.groups;
    const FlTraj tr = a.trajs[j];
    const int fbeg = g * a.TF;
    it.nf = max(0, min(a.TF, tr.n_frames - fbeg));
    const int4 d = __ldg((const int4*)tr.d_tile_desc + tile);
    it.nodes = tr.d_tile_nodes + d.x;
    it.n_nodes = it.nf > 0 ? d.y : 0;
    it.patches = tr.d_tile_patches + d.z;
    it.n_patches = it.nf > 0 ? d.w : 0;
    const long long t = (long long)tr.t0 + (long long)fbeg * tr.interval;
    it.vel = tr.d_velocity + t * tr.vel_stride;
    it.prs = tr.d_pressure + t * tr.prs_stride;
    it.vstep = tr.interval * tr.vel_stride;      // < 2^31: checked by the host
    it.pstep = tr.interval * tr.prs_stride;
    it.idx = (const int4*)tr.d_idx_tile;
    it.w = (const double2*)tr.d_w;
    it.states = tr.d_states + (size_t)fbeg * a.n_patches * 3 * ppx;
    it.mask = tr.d_mask ? tr.d_mask + (size_t)fbeg * a.n_patches * ppx : nullptr;
    *out = it;
}

// running scan of the staged values: a non-finite or huge value sends the whole item down the checked path
struct Scan {
    float nanacc = 0.f, amax = 0.f;
    __device__ __forceinline__ void add(float a, float b, float c) {
        nanacc = fmaf(a, 0.f, fmaf(b, 0.f, fmaf(c, 0.f, nanacc)));
        amax = fmaxf(fmaxf(fabsf(a), fabsf(b)), fmaxf(fabsf(c), amax));
    }
    __device__ __forceinline__ int bad() const { return !(nanacc == 0.f) || amax > 1.0e30f; }
};

// stage node `s` of item `nx` for all its frames (not overlapped: prologue, warps without a chunk, left-over nodes)
__device__ __forceinline__ void stage_node_all_frames(const Item* nxs, int s, uint32_t sbuf, int slot_rec, int f0, Scan& sc) {
    const int n = __ldg(nxs->nodes + s);
    const int vstep = nxs->vstep, pstep = nxs->pstep, nf = nxs->nf;
    const float* vp = nxs->vel + 2 * (size_t)n + (size_t)f0 * vstep;
    const float* pp = nxs->prs + (size_t)n + (size_t)f0 * pstep;
    uint32_t dst = sbuf + ((uint32_t)f0 * slot_rec + s) * 16u;
    for (int f = f0; f < nf; ++f, vp += vstep, pp += pstep, dst += slot_rec * 16u) {
        const float2 va = ldg_stream2(vp);
        const float pa = ldg_stream1(pp);
        sc.add(va.x, va.y, pa);
        sts128(dst, make_record(va.x, va.y, pa));
    }
}

// One chunk (128 output pixels of one patch) x the item's frames, with the staging of one node of the next item folded into
// the frame loop.  WPP = warps per patch (ppx / 128); the WPP warps of a patch group share two output tiles.
template <bool CHECKED, int WPP>
__device__ __forceinline__ void chunk_frames(const Item* cs, const Item* nxs, int chunk, uint32_t stage_cur, uint32_t stage_nxt,
                                             int job, uint32_t ring, int& parity, const TiledArgs& a, Scan& scan) {
    constexpr int ppx = 128 * WPP;
    constexpr uint32_t TILE_BYTES = ppx * 13;       // [3][ppx] floats + ppx mask bytes
    const int lane_id = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = WPP == 1 ? 0 : (chunk % WPP);
    const bool leader = sub == 0 && lane_id == 0;
    const int bar_id = 1 + warp / WPP;
    const bool mask_aware = a.flags & FL_MASK_AWARE_NORM, no_norm = a.flags & FL_NO_NORM;
    // pixel of this lane inside a group of 32 adjacent pixels (2 patch rows x 16): the 8 lanes of a quarter-warp (one
    // LDS.128 wavefront) take a compact 2 x 4 block, which touches fewer distinct nodes than a 1 x 8 strip
    const int lane = ((lane_id >> 2) & 1) * 16 + (lane_id >> 3) * 4 + (lane_id & 3);
    // the items live in shared memory (uniform reads); only what the frame loop needs is kept in registers
    const int patch = __ldg(cs->patches + chunk / WPP);
    const int o = patch * ppx + sub * 128 + lane;
    const int4* idx_tab = cs->idx;
    const double2* w_tab = cs->w;
    uint32_t ov[NP][3];
    double w0[NP], w1[NP], w2[NP];
    unsigned mbits = 0;     // byte r = 1 if pixel r is outside the mesh
#pragma unroll
    for (int r = 0; r < NP; ++r) {
        const int4 id = __ldg(idx_tab + o + 32 * r);
        const double2 ww = __ldg(w_tab + o + 32 * r);
        const bool out = id.w < 0;
        mbits |= out ? (1u << (8 * r)) : 0u;
        w1[r] = out ? 0.0 : ww.x;
        w2[r] = out ? 0.0 : ww.y;
        w0[r] = out ? 0.0 : 1.0 - ww.x - ww.y;
        ov[r][0] = out ? 0u : (uint32_t)id.x;        // already 16 * slot
        ov[r][1] = out ? 0u : (uint32_t)id.y;
        ov[r][2] = out ? 0u : (uint32_t)id.z;
    }
    // mask bytes in pixel order: lane j holds the four bytes of pixels 4j..4j+3 of the chunk (static on the unchecked path)
    auto mask_word = [&](unsigned bits) {
        unsigned word = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int q = 4 * (lane_id & 7) + i;                                         // position inside the 32-pixel group
            const int src = (((q >> 2) & 3) << 3) | (((q >> 4) & 1) << 2) | (q & 3);     // inverse of the lane permutation
            const unsigned m = __shfl_sync(0xffffffffu, bits, src);
            word |= ((m >> (8 * (lane_id >> 3))) & 1u) << (8 * i);
        }
        return word;
    };
    const bool want_mask = cs->mask != nullptr;
    const int nf = cs->nf;
    // the patch group's two output tiles; element (c, pixel k of the patch) at c * ppx + k, mask bytes behind the floats
    const uint32_t tile0 = ring + (uint32_t)(warp / WPP) * 2u * TILE_BYTES;
    const uint32_t my_f = (uint32_t)(sub * 128 + lane) * 4u, my_m = 3u * ppx * 4u + (uint32_t)(sub * 128 + 4 * lane_id);
    if (leader) bulk_wait_read0();                    // both tiles are free (this thread issued every copy that read them)
    if (WPP > 1) group_sync(bar_id, 32 * WPP); else __syncwarp();
    if (!CHECKED && want_mask) {
        const unsigned mw = mask_word(mbits);
        sts32u(tile0 + my_m, mw);
        sts32u(tile0 + TILE_BYTES + my_m, mw);
    }
    unsigned long long nm[3], ns[3], rc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        nm[c] = pack2(-a.sc.mean[c], -a.sc.mean[c]);
        ns[c] = pack2(-a.sc.stdv[c], -a.sc.stdv[c]);
        rc[c] = pack2(a.sc.rcp[c], a.sc.rcp[c]);
    }
    // staging job of this thread: node `job` of the next item, one frame per iteration
    const bool job_ok = job < nxs->n_nodes;
    const float* vp = nullptr;
    const float* pp = nullptr;
    uint32_t sdst = stage_nxt + (uint32_t)job * 16u;
    if (job_ok) {
        const int n = __ldg(nxs->nodes + job);
        vp = nxs->vel + 2 * (size_t)n;
        pp = nxs->prs + (size_t)n;
    }
    const int vstep = nxs->vstep, pstep = nxs->pstep;
    const int nf_stage = job_ok ? min(nf, nxs->nf) : 0;
    uint32_t nb = stage_cur;
#pragma unroll 1
    for (int f = 0; f < nf; ++f, nb += a.slot_rec * 16u) {
        float2 sva = make_float2(0.f, 0.f);
        float spa = 0.f;
        const bool do_stage = f < nf_stage;
        if (do_stage) { sva = ldg_stream2(vp); spa = ldg_stream1(pp); vp += vstep; pp += pstep; }
        float res[3][NP];
        unsigned fm = mbits;
#pragma unroll
        for (int r = 0; r < NP; ++r) {
            const uint4 a0 = lds128(nb + ov[r][0]);   // one 128-bit gather per vertex
            const uint4 a1 = lds128(nb + ov[r][1]);
            const uint4 a2 = lds128(nb + ov[r][2]);
            res[0][r] = (float)fma(w2[r], rec_u(a2), fma(w1[r], rec_u(a1), w0[r] * rec_u(a0)));
            res[1][r] = (float)fma(w2[r], rec_v(a2), fma(w1[r], rec_v(a1), w0[r] * rec_v(a0)));
            res[2][r] = (float)fma(w2[r], rec_p(a2), fma(w1[r], rec_p(a1), w0[r] * rec_p(a0)));
            if (CHECKED) {
                if (!finite_f(res[2][r])) fm |= 1u << (8 * r);           // pressure mask only (simple_dataloader.py:114,119)
#pragma unroll
                for (int c = 0; c < 3; ++c) if (!finite_f(res[c][r])) res[c][r] = 0.f;   // mesh_utils.py:89, per channel
            }
        }
        if (!no_norm) {
            if (CHECKED) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int r = 0; r < NP; ++r) {
                        const float x = res[c][r];
                        const float y = __fdiv_rn(__fsub_rn(x, a.sc.mean[c]), a.sc.stdv[c]);
                        res[c][r] = (mask_aware && ((fm >> (8 * r)) & 1u)) ? x : y;   // airfoil_ds.py:241-242
                    }
            } else if (mask_aware && fm) {      // rare: pixels on the mesh boundary / padding stay raw
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int r = 0; r < NP; ++r)
                        if (!((fm >> (8 * r)) & 1u)) res[c][r] = norm_fast(res[c][r], a.sc.mean[c], a.sc.stdv[c], a.sc.rcp[c]);
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int r = 0; r < NP; r += 2) norm_fast2(res[c][r], res[c][r + 1], nm[c], ns[c], rc[c]);
            }
        }
        // this frame's tile: the copy that last read it was confirmed complete before the previous group barrier
        const uint32_t tile = tile0 + (uint32_t)(parity & 1) * TILE_BYTES;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int r = 0; r < NP; ++r) sts32(tile + my_f + (uint32_t)(c * ppx + 32 * r) * 4u, res[c][r]);
        if (CHECKED && want_mask) sts32u(tile + my_m, mask_word(fm));
        fence_async_smem();                            // generic-proxy writes -> visible to the bulk copy
        if (leader) bulk_wait_read0();                 // the previous frame's copy has left the other tile
        if (WPP > 1) group_sync(bar_id, 32 * WPP); else __syncwarp();
        if (leader) {
            const size_t fp = (size_t)f * a.n_patches + patch;       // (frame, patch) block of the item's output
            bulk_store(cs->states + fp * (3 * ppx), tile, 3u * ppx * 4u);
            if (want_mask) bulk_store(cs->mask + fp * ppx, tile + 3u * ppx * 4u, ppx);
            bulk_commit();
        }
        parity ^= 1;
        if (do_stage) {
            scan.add(sva.x, sva.y, spa);
            sts128(sdst, make_record(sva.x, sva.y, spa));
            sdst += a.slot_rec * 16u;
        }
    }
    // frames of the next item beyond this item's frame count (a short last group followed by a full one)
    if (job_ok && nxs->nf > nf_stage) stage_node_all_frames(nxs, job, stage_nxt, a.slot_rec, nf_stage, scan);
}

template <int WPP>
__global__ void __launch_bounds__(TL_THREADS, 1) k_interp_patchify_tiled(TiledArgs a) {
    constexpr int ppx = 128 * WPP;
    constexpr uint32_t TILE_BYTES = ppx * 13;
    constexpr uint32_t RING_BYTES = (TL_WARPS / WPP) * 2 * TILE_BYTES;
    extern __shared__ __align__(128) unsigned char fl_smem[];
    Item* s_items = (Item*)fl_smem;                                   // [2]
    const uint32_t ring = smem_u32(fl_smem) + 2 * ITEM_WORDS * 4;
    const uint32_t stage0 = ring + RING_BYTES;
    const uint32_t stage_bytes = (uint32_t)a.TF * a.slot_rec * 16u;
    const int warp = threadIdx.x >> 5;
    int cur = 0, parity = 0;
    Scan scan;
    // prologue: decode and stage the first item (not overlapped)
    if (threadIdx.x == 0) decode_item(a, blockIdx.x, ppx, &s_items[0]);
    __syncthreads();
    for (int s = threadIdx.x; s < s_items[0].n_nodes; s += TL_THREADS) stage_node_all_frames(&s_items[0], s, stage0, a.slot_rec, 0, scan);
    if (threadIdx.x == 0) decode_item(a, blockIdx.x + gridDim.x, ppx, &s_items[1]);
    int bad = __syncthreads_or(scan.bad() || !a.sc.fast_div);
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const Item* it = &s_items[cur];
        const Item* nx = &s_items[cur ^ 1];
        const uint32_t stage_cur = stage0 + (uint32_t)cur * stage_bytes, stage_nxt = stage0 + (uint32_t)(cur ^ 1) * stage_bytes;
        const int n_chunks = it->n_patches * WPP;
        const int rounds = (n_chunks + TL_WARPS - 1) / TL_WARPS;
        scan = Scan();
        for (int round = 0; round < rounds; ++round) {
            const int chunk = round * TL_WARPS + warp;
            const int job = round * TL_THREADS + threadIdx.x;
            if (chunk < n_chunks) {
                if (bad) chunk_frames<true, WPP>(it, nx, chunk, stage_cur, stage_nxt, job, ring, parity, a, scan);
                else chunk_frames<false, WPP>(it, nx, chunk, stage_cur, stage_nxt, job, ring, parity, a, scan);
            } else if (job < nx->n_nodes) {
                stage_node_all_frames(nx, job, stage_nxt, a.slot_rec, 0, scan);
            }
        }
        for (int s = rounds * TL_THREADS + threadIdx.x; s < nx->n_nodes; s += TL_THREADS)
            stage_node_all_frames(nx, s, stage_nxt, a.slot_rec, 0, scan);
        __syncthreads();       // every warp is done with s_items[cur] and with the gathers out of stage_cur
        if (threadIdx.x == 0) decode_item(a, item + 2 * gridDim.x, ppx, &s_items[cur]);
        bad = __syncthreads_or(scan.bad() || !a.sc.fast_div);
        cur ^= 1;
    }
    bulk_wait_read0();         // no bulk copy may still be reading shared memory when the CTA exits
}

}  // namespace

int fli::launch_tiled(const FlTraj* d_trajs, const FlTraj* h_trajs, int n_traj, int max_frames, int n_patches, int px, int py,
                      const StagedConst& sc, unsigned flags, cudaStream_t st) {
    const int ppx = px * py;
    if (!h_trajs || (ppx != 128 && ppx != 256)) return 1;
    const int wpp = ppx / 128;
    int n_tiles = h_trajs[0].n_tiles, max_nodes = 0;
    for (int i = 0; i < n_traj; ++i) {
        const FlTraj& t = h_trajs[i];
        if (!t.d_idx_tile || !t.d_tile_nodes || !t.d_tile_desc || !t.d_tile_patches || t.n_tiles < 1) return 1;
        FL_REQUIRE(t.n_tiles == n_tiles, FL_E_ARG, "fl_interp_patchify: trajectory %d has %d tiles, trajectory 0 has %d", i, t.n_tiles, n_tiles);
        FL_REQUIRE(t.max_tile_nodes >= 0, FL_E_ARG, "fl_interp_patchify: trajectory %d: negative max_tile_nodes", i);
        FL_REQUIRE((uintptr_t)t.d_idx_tile % 16 == 0 && (uintptr_t)t.d_tile_desc % 16 == 0 && (uintptr_t)t.d_states % 16 == 0 &&
                       (t.d_mask == nullptr || (uintptr_t)t.d_mask % 16 == 0),
                   FL_E_ALIGN, "fl_interp_patchify: trajectory %d: tile tables and outputs must be 16-byte aligned", i);
        max_nodes = t.max_tile_nodes > max_nodes ? t.max_tile_nodes : max_nodes;
    }
    const int slot_rec = max_nodes > 0 ? max_nodes : 1;
    const size_t ring = (size_t)(TL_WARPS / wpp) * 2 * ppx * 13;
    const size_t fixed = 2 * ITEM_WORDS * 4 + ring;
    long TF = ((long)SMEM_TOTAL - (long)fixed) / 2 / (16L * slot_rec);
    if (TF > 16) TF = 16;
    if (TF > max_frames) TF = max_frames;
    if (const char* e = getenv("FLUIDGRID_TF")) { long v = atol(e); if (v >= 1 && v < TF) TF = v; }
    if (TF < 1) return 1;        // a tile's nodes do not fit: the caller falls back (plan smaller tiles)
    TiledArgs a;
    a.trajs = d_trajs;
    a.groups = (max_frames + (int)TF - 1) / (int)TF;
    a.n_tiles = n_tiles;
    const long n_items = (long)a.groups * n_tiles * n_traj;
    FL_REQUIRE(n_items < 0x7fffffffL - 2 * FL_SM_COUNT, FL_E_ARG, "fl_interp_patchify: too many work items (%ld)", n_items);
    a.n_items = (int)n_items;
    a.TF = (int)TF;
    a.n_patches = n_patches;
    a.slot_rec = slot_rec;
    a.sc = sc;
    a.flags = flags;
    const size_t smem = fixed + 2 * (size_t)TF * slot_rec * 16;
    const int grid = n_items < FL_SM_COUNT ? (int)n_items : FL_SM_COUNT;      // one persistent CTA per SM
    if (wpp == 2) {
        static FlOncePerDevice attr;
        if (attr.first_use()) FL_CUDA(cudaFuncSetAttribute(k_interp_patchify_tiled<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
        k_interp_patchify_tiled<2><<<grid, TL_THREADS, smem, st>>>(a);
    } else {
        static FlOncePerDevice attr;
        if (attr.first_use()) FL_CUDA(cudaFuncSetAttribute(k_interp_patchify_tiled<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
        k_interp_patchify_tiled<1><<<grid, TL_THREADS, smem, st>>>(a);
    }
    FL_LAUNCH_CHECK();
    return FL_OK;
}
