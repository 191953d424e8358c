// fl_tiled.cu -- the per-step kernel, tiled and warp-specialised: gather -> fp64 FMA -> fp32 -> mask -> normalise ->
// patchify for meshes of ANY size out of shared memory.
//
// Same work and the same arithmetic as k_interp_patchify_staged (fl_interp.cu; replaces simple_dataloader.py:104-152,
// 166-216 / airfoil_ds.py:71-139,216-244 of the reference), reorganised around what ncu showed to bound that kernel
// (profiles/README.md): the L1/shared data pipe carried the gathers AND 12 scalar global stores per 4 pixel-frames, the
// conversion unit 9 fp32->fp64 conversions per pixel-frame, and staging was serialised behind a CTA-wide barrier.
//
//   tiles      the output patches are split into tiles (runs of patches along a serpentine through the patch grid); a
//              tile's node list holds exactly the mesh nodes its pixels touch (FlTraj::d_tile_*), and the table carries
//              tile-local slots.  A work item = one tile x TF consecutive selected frames of one trajectory, so shared
//              memory holds TF x (nodes of ONE tile) records whatever the size of the mesh.
//   records    16 bytes per node and frame with u, v and p ALREADY in fp64 form: their high words plus one word with the
//              top byte of each low word (a float widened to double has 29 zero bits at the bottom).  A vertex costs
//              one LDS.128 + three PRMT and no conversion.
//   producers  the last warps of the CTA only stage: they walk the CTA's items, load the tile's nodes frame by frame
//              (several frames in flight per thread), convert, write the records of item k into buffer k & 1 and arrive
//              on its `full` mbarrier; they wait on `empty` before overwriting a buffer.
//   consumers  the other warps never meet at a CTA-wide barrier.  A tile has as many patches as the CTA has patch groups
//              (the warps that share a patch), and a CTA works through (trajectory, tile, run of frame groups) units:
//              a group reads the table records of ITS patch once per unit and keeps them in registers for every frame of
//              the run, waits on `full` for each frame group and arrives on `empty` when it is done with it.
//   output     results go to a per-group shared-memory tile ([3][px*py] floats + px*py mask bytes, conflict-free
//              STS.32) and leave as ONE bulk copy (cp.async.bulk shared -> global, 3 KB) per frame and patch, issued by
//              one thread of the group; two tiles per group, so the copy of frame f overlaps the arithmetic of frame
//              f+1.  No global store goes through the LSU.
// HBM traffic per frame: 12 P (+ P mask) written, 12 x (sum of the tiles' node counts) read (= 12 N plus the halo
// nodes shared by neighbouring tiles, part of which hit L2 because the tiles of a frame group run at the same time).
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
constexpr int ITEM_BYTES = 128;              // decoded work item kept in shared memory (two of them)
constexpr int HEAD_BYTES = 2 * ITEM_BYTES + 128;   // items + mbarriers
constexpr int STAGE_BATCH = 8;               // frames a producer thread keeps in flight

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mb_init(uint32_t bar, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mb_wait(uint32_t bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nTL_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TL_DONE;\nbra TL_WAIT;\nTL_DONE:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts32u(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
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

// node record {bits, hi(u), hi(p), hi(v)}: the three values widened to fp64; byte k of `bits` = the top byte of the k-th value's
// low word (u, v, p), whose other 24 bits are zero for a widened float (normal, denormal, inf or NaN alike)
__device__ __forceinline__ uint4 make_record(float u, float v, float p) {
    const double du = (double)u, dv = (double)v, dp = (double)p;
    const uint32_t bits = ((uint32_t)__double2loint(du) >> 24) | (((uint32_t)__double2loint(dv) >> 24) << 8) |
                          (((uint32_t)__double2loint(dp) >> 24) << 16);
    return make_uint4(bits, (uint32_t)__double2hiint(du), (uint32_t)__double2hiint(dp), (uint32_t)__double2hiint(dv));
}
// one 128-bit gather -> the three doubles, no conversion: u and v are completed IN PLACE (a PRMT writes the low word next to
// the high word the load put there), p takes a PRMT and a move
__device__ __forceinline__ void load_record(uint32_t addr, double& u, double& v, double& p) {
    unsigned long long A, B;        // A = {bits, hi(u)}, B = {hi(p), hi(v)}
    asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(A), "=l"(B) : "r"(addr));
    const uint32_t bits = (uint32_t)A;
    p = __hiloint2double((int)(uint32_t)B, (int)__byte_perm(bits, 0u, 0x2444));
    v = __longlong_as_double((long long)((B & 0xffffffff00000000ull) | __byte_perm(bits, 0u, 0x1444)));
    u = __longlong_as_double((long long)((A & 0xffffffff00000000ull) | __byte_perm(bits, 0u, 0x0444)));
}

// A decoded work item (tile x frame group of one trajectory), written to shared memory by a producer thread.
struct alignas(128) Item {
    const float* vel;        // node fields at the item's first selected frame
    const float* prs;
    const int* nodes;        // the tile's node list
    const int* patches;      // the tile's patch ids
    const int4* idx;         // table with tile-local slots * 16
    const double2* w;
    float* states;           // output at the item's first frame
    uint8_t* mask;           // or NULL
    const int4* qslots;      // per quad (4 consecutive node ids) the tile touches: the slots of its 4 nodes (-1: not a node of this tile)
    const int* quads;        // the quads' ids (node id / 4), ascending
    int vstep, pstep;        // floats between selected frames (interval * stride)
    int n_nodes, n_patches, nf;
    int q_cnt, q_max;        // quads of the tile (0: stage node by node); q_max = the largest quad id + 1
    int bad;                 // 1: non-finite / huge node values (or constants unfit for the fast division): checked path
};
static_assert(sizeof(Item) == ITEM_BYTES, "Item must fill its shared-memory slot");

struct TiledArgs {
    const FlTraj* trajs;
    int n_units, fchunks, gpc, groups, n_tiles, TF, n_patches, slot_rec;
    // unit = (trajectory, run `c` of gpc frame groups, tile); groups = frame groups per trajectory; slot_rec = records per staged frame
    StagedConst sc;
    unsigned flags;
    unsigned dbg;     // development ablations (FLUIDGRID_DBG): 1 no proxy fence, 2 no bulk copies, 4 producers stage nothing, 8 no group barrier, 16 no L2 prefetch, 32 node-by-node staging, 64 consumers compute nothing
};

__device__ __forceinline__ void decode_item(const TiledArgs& a, int j, int tile, int g, int ppx, Item* out) {
    Item it;
    const FlTraj tr = a.trajs[j];
    const int fbeg = g * a.TF;
    it.nf = max(0, min(a.TF, tr.n_frames - fbeg));
    const int4 d = __ldg((const int4*)tr.d_tile_desc + 2 * tile), e = __ldg((const int4*)tr.d_tile_desc + 2 * tile + 1);
    it.nodes = tr.d_tile_nodes + d.x;
    it.n_nodes = it.nf > 0 ? d.y : 0;
    it.patches = tr.d_tile_patches + d.z;
    it.n_patches = it.nf > 0 ? d.w : 0;
    // coalesced quad staging needs 16-byte aligned frames whose pitch covers whole quads (the padded layout of the host package)
    const bool quads = tr.d_tile_qslots && tr.vel_stride % 4 == 0 && tr.prs_stride % 4 == 0 && ((uintptr_t)tr.d_velocity % 16 == 0) &&
                       ((uintptr_t)tr.d_pressure % 16 == 0) && 4 * e.z <= tr.prs_stride && 8 * e.z <= tr.vel_stride;
    it.q_max = e.z;
    it.q_cnt = quads && it.nf > 0 ? e.y : 0;
    it.qslots = (const int4*)tr.d_tile_qslots + e.x;
    it.quads = tr.d_tile_quads + e.x;
    const long long t = (long long)tr.t0 + (long long)fbeg * tr.interval;
    it.vel = tr.d_velocity + t * tr.vel_stride;
    it.prs = tr.d_pressure + t * tr.prs_stride;
    it.vstep = tr.interval * tr.vel_stride;      // < 2^31: checked by the host
    it.pstep = tr.interval * tr.prs_stride;
    it.idx = (const int4*)tr.d_idx_tile;
    it.w = (const double2*)tr.d_w;
    it.states = tr.d_states + (size_t)fbeg * a.n_patches * 3 * ppx;
    it.mask = tr.d_mask ? tr.d_mask + (size_t)fbeg * a.n_patches * ppx : nullptr;
    it.bad = a.sc.fast_div ? 0 : 1;
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

// ---- producer warps -------------------------------------------------------------------------------------------------
// nodes `s` and `s2` of the item (s2 < 0: none), all their frames, the loads of STAGE_BATCH frames of both in flight at once
__device__ __forceinline__ void stage_nodes(const Item* it, int s, int s2, uint32_t sbuf, int slot_rec, Scan& sc) {
    const bool two = s2 >= 0;
    const int n = __ldg(it->nodes + s), n2 = two ? __ldg(it->nodes + s2) : 0;
    const int vstep = it->vstep, pstep = it->pstep, nf = it->nf;
    const float* vp = it->vel + 2 * (size_t)n;
    const float* pp = it->prs + (size_t)n;
    const float* vq = it->vel + 2 * (size_t)n2;
    const float* pq = it->prs + (size_t)n2;
    uint32_t dst = sbuf + (uint32_t)s * 16u, dst2 = sbuf + (uint32_t)(two ? s2 : 0) * 16u;
    for (int f0 = 0; f0 < nf; f0 += STAGE_BATCH) {
        float2 v[STAGE_BATCH], v2[STAGE_BATCH];
        float p[STAGE_BATCH], p2[STAGE_BATCH];
#pragma unroll
        for (int i = 0; i < STAGE_BATCH; ++i)
            if (f0 + i < nf) {
                v[i] = ldg_stream2(vp + (size_t)i * vstep);
                p[i] = ldg_stream1(pp + (size_t)i * pstep);
                if (two) { v2[i] = ldg_stream2(vq + (size_t)i * vstep); p2[i] = ldg_stream1(pq + (size_t)i * pstep); }
            }
        vp += (size_t)STAGE_BATCH * vstep; pp += (size_t)STAGE_BATCH * pstep;
        vq += (size_t)STAGE_BATCH * vstep; pq += (size_t)STAGE_BATCH * pstep;
#pragma unroll
        for (int i = 0; i < STAGE_BATCH; ++i)
            if (f0 + i < nf) {
                sc.add(v[i].x, v[i].y, p[i]);
                sts128(dst, make_record(v[i].x, v[i].y, p[i]));
                dst += (uint32_t)slot_rec * 16u;
                if (two) {
                    sc.add(v2[i].x, v2[i].y, p2[i]);
                    sts128(dst2, make_record(v2[i].x, v2[i].y, p2[i]));
                    dst2 += (uint32_t)slot_rec * 16u;
                }
            }
    }
}

// Coalesced form: the tile's nodes lie in q_cnt quads of 4 consecutive node ids (listed ascending, so runs of nodes are runs
// of quads); a thread takes one quad x 4 frames (two 128-bit loads of velocities and one of pressures per frame,
// consecutive threads on consecutive list entries) and writes the records of the nodes the tile uses to their slots.
__device__ __forceinline__ float4 ldg_stream4f(const float* p) { return fl_ldg_stream4((const float4*)p); }

// one quad x FB frames in registers
template <int FB>
struct QuadBatch {
    float4 va[FB], vb[FB], pq[FB];
    int4 sl;
    int f0;
};
template <int FB>
__device__ __forceinline__ void quad_load(QuadBatch<FB>& qb, const Item* it, int u, int qc) {
    const int fb = u / qc, e = u - fb * qc;
    qb.f0 = fb * FB;
    qb.sl = __ldg(it->qslots + e);
    const int q = __ldg(it->quads + e);
    const int vstep = it->vstep, pstep = it->pstep, nf = it->nf;
    const float* vp = it->vel + (size_t)qb.f0 * vstep + 8 * (size_t)q;
    const float* pp = it->prs + (size_t)qb.f0 * pstep + 4 * (size_t)q;
#pragma unroll
    for (int i = 0; i < FB; ++i)
        if (qb.f0 + i < nf) {
            qb.va[i] = ldg_stream4f(vp + (size_t)i * vstep);
            qb.vb[i] = ldg_stream4f(vp + (size_t)i * vstep + 4);
            qb.pq[i] = ldg_stream4f(pp + (size_t)i * pstep);
        }
}
template <int FB>
__device__ __forceinline__ void quad_store(const QuadBatch<FB>& qb, int nf, uint32_t sbuf, int slot_rec, Scan& sc) {
    uint32_t dst = sbuf + (uint32_t)qb.f0 * slot_rec * 16u;
    const int4 sl = qb.sl;
#pragma unroll
    for (int i = 0; i < FB; ++i)
        if (qb.f0 + i < nf) {
            if (sl.x >= 0) { sc.add(qb.va[i].x, qb.va[i].y, qb.pq[i].x); sts128(dst + 16u * sl.x, make_record(qb.va[i].x, qb.va[i].y, qb.pq[i].x)); }
            if (sl.y >= 0) { sc.add(qb.va[i].z, qb.va[i].w, qb.pq[i].y); sts128(dst + 16u * sl.y, make_record(qb.va[i].z, qb.va[i].w, qb.pq[i].y)); }
            if (sl.z >= 0) { sc.add(qb.vb[i].x, qb.vb[i].y, qb.pq[i].z); sts128(dst + 16u * sl.z, make_record(qb.vb[i].x, qb.vb[i].y, qb.pq[i].z)); }
            if (sl.w >= 0) { sc.add(qb.vb[i].z, qb.vb[i].w, qb.pq[i].w); sts128(dst + 16u * sl.w, make_record(qb.vb[i].z, qb.vb[i].w, qb.pq[i].w)); }
            dst += (uint32_t)slot_rec * 16u;
        }
}
// Coalesced form: the tile's nodes lie in q_cnt quads of 4 consecutive node ids (listed ascending, so runs of nodes are runs
// of quads); a thread takes one quad x FB frames at a time (two 128-bit loads of velocities and one of pressures per frame,
// consecutive threads on consecutive list entries) and writes the records of the nodes the tile uses to their slots.
// (Two batches in flight per thread -- loads of the next issued before the current is converted -- spill at 128 registers.)
__device__ __forceinline__ void stage_quads(const Item* it, int ptid, int PT, uint32_t sbuf, int slot_rec, Scan& sc) {
    constexpr int FB = 4;
    const int qc = it->q_cnt, nf = it->nf;
    const int total = qc * ((nf + FB - 1) / FB);
    for (int u = ptid; u < total; u += PT) {
        QuadBatch<FB> qb;
        quad_load(qb, it, u, qc);
        quad_store(qb, nf, sbuf, slot_rec, sc);
    }
}

template <int PROD_WARPS>
__device__ __forceinline__ void producer_loop(const TiledArgs& a, Item* s_items, uint32_t bar_full, uint32_t bar_empty, uint32_t stage0,
                                              uint32_t stage_bytes, int ppx) {
    constexpr int PT = PROD_WARPS * 32;
    const int ptid = threadIdx.x;          // the producers are the first warps of the CTA
    int k = 0;
    for (int unit = blockIdx.x; unit < a.n_units; unit += gridDim.x) {
        const int tile = unit % a.n_tiles, jc = unit / a.n_tiles;
        const int c = jc % a.fchunks, j = jc / a.fchunks;
        const int g1 = min(a.groups, (c + 1) * a.gpc);
        for (int g = c * a.gpc; g < g1; ++g, ++k) {
            const int b = k & 1;
            if (k >= 2) mb_wait(bar_empty + 8u * b, ((k >> 1) - 1) & 1);       // every consumer warp is done with item k - 2
            if (ptid == 0) decode_item(a, j, tile, g, ppx, &s_items[b]);
            named_sync(15, PT);
            const Item* it = &s_items[b];
            Scan scan;
            const int S = it->n_nodes;
            if (!(a.dbg & 4u) || k < 2) {
                if (it->q_cnt > 0 && !(a.dbg & 32u)) stage_quads(it, ptid, PT, stage0 + (uint32_t)b * stage_bytes, a.slot_rec, scan);
                else for (int s = ptid; s < S; s += PT) stage_nodes(it, s, -1, stage0 + (uint32_t)b * stage_bytes, a.slot_rec, scan);
            }
            if (__any_sync(0xffffffffu, scan.bad()) && (threadIdx.x & 31) == 0) atomicOr(&s_items[b].bad, 1);
            mb_arrive(bar_full + 8u * b);           // release: the records (and the bad flag) are visible to whoever waits
        }
    }
}

// ---- consumer warps -------------------------------------------------------------------------------------------------
// What a thread keeps in registers for its 4 pixels of its group's patch, for all frames of a unit
struct PixelRegs {
    uint32_t ov[NP][3];        // byte offsets of the three vertices' records inside a staged frame
    double w0[NP], w1[NP], w2[NP];
    unsigned mbits;            // byte r = 1 if pixel r is outside the mesh
};

// mask bytes in pixel order: lane j returns the four bytes of pixels 4j..4j+3 of the warp's 128 pixels
__device__ __forceinline__ unsigned mask_word(unsigned bits) {
    const int lane_id = threadIdx.x & 31;
    unsigned word = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int q = 4 * (lane_id & 7) + i;                                         // position inside the 32-pixel group
        const int src = (((q >> 2) & 3) << 3) | (((q >> 4) & 1) << 2) | (q & 3);     // inverse of the lane permutation
        const unsigned m = __shfl_sync(0xffffffffu, bits, src);
        word |= ((m >> (8 * (lane_id >> 3))) & 1u) << (8 * i);
    }
    return word;
}

// The frames of one item for the group's patch.  WPP = warps per patch (ppx / 128); the WPP warps of a patch group share
// two output tiles and the named barrier 1 + group.
template <bool CHECKED, int WPP>
__device__ __forceinline__ void run_frames(const PixelRegs& px, const Item* cs, int patch, uint32_t stage_cur, uint32_t tile0, int group,
                                           int& parity, const TiledArgs& a) {
    constexpr int ppx = 128 * WPP;
    constexpr uint32_t TILE_BYTES = ppx * 13;       // [3][ppx] floats + ppx mask bytes
    const int lane_id = threadIdx.x & 31;
    const int sub = WPP == 1 ? 0 : ((threadIdx.x >> 5) % WPP);       // the producer warps come in pairs, so parity is kept
    const bool leader = sub == 0 && lane_id == 0;
    const int bar_id = 1 + group;
    const bool mask_aware = a.flags & FL_MASK_AWARE_NORM, no_norm = a.flags & FL_NO_NORM;
    const int lane = ((lane_id >> 2) & 1) * 16 + (lane_id >> 3) * 4 + (lane_id & 3);
    const bool want_mask = cs->mask != nullptr;
    const int nf = cs->nf;
    // element (c, pixel k of the patch) of a tile at c * ppx + k, mask bytes behind the floats
    const uint32_t my_f = (uint32_t)(sub * 128 + lane) * 4u, my_m = 3u * ppx * 4u + (uint32_t)(sub * 128 + 4 * lane_id);
    unsigned long long nm[3], ns[3], rc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        nm[c] = pack2(-a.sc.mean[c], -a.sc.mean[c]);
        ns[c] = pack2(-a.sc.stdv[c], -a.sc.stdv[c]);
        rc[c] = pack2(a.sc.rcp[c], a.sc.rcp[c]);
    }
    // leader only: where this patch's block of the item's first frame goes
    float* gst = cs->states + (size_t)patch * (3 * ppx);
    uint8_t* gmk = want_mask ? cs->mask + (size_t)patch * ppx : nullptr;
    const size_t gst_step = (size_t)a.n_patches * (3 * ppx), gmk_step = (size_t)a.n_patches * ppx;
    uint32_t nb = stage_cur;
#pragma unroll 1
    for (int f = 0; f < nf; ++f, nb += a.slot_rec * 16u) {
        float res[3][NP];
        unsigned fm = px.mbits;
#pragma unroll
        for (int r = 0; r < NP; ++r) {
            double u0, v0, p0, u1, v1, p1, u2, v2, p2;
            load_record(nb + px.ov[r][0], u0, v0, p0);   // one 128-bit gather per vertex
            load_record(nb + px.ov[r][1], u1, v1, p1);
            load_record(nb + px.ov[r][2], u2, v2, p2);
            res[0][r] = (float)fma(px.w2[r], u2, fma(px.w1[r], u1, px.w0[r] * u0));
            res[1][r] = (float)fma(px.w2[r], v2, fma(px.w1[r], v1, px.w0[r] * v0));
            res[2][r] = (float)fma(px.w2[r], p2, fma(px.w1[r], p1, px.w0[r] * p0));
            if (CHECKED) {
                if (!finite_f(res[2][r])) fm |= 1u << (8 * r);           // pressure mask only (simple_dataloader.py:114,119)
#pragma unroll
                for (int c = 0; c < 3; ++c) if (!finite_f(res[c][r])) res[c][r] = 0.f;   // mesh_utils.py:89, per channel
            }
        }
        if (px.mbits) {        // outside the mesh the weights are 0 and the sum is +-0: the reference stores +0.0 (mesh_utils.py:89)
#pragma unroll
            for (int r = 0; r < NP; ++r)
                if ((px.mbits >> (8 * r)) & 1u) { res[0][r] = 0.f; res[1][r] = 0.f; res[2][r] = 0.f; }
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
        if (!(a.dbg & 1u)) fence_async_smem();         // generic-proxy writes -> visible to the bulk copy
        if (leader) bulk_wait_read0();                 // the previous frame's copy has left the other tile
        if (!(a.dbg & 8u)) { if (WPP > 1) named_sync(bar_id, 32 * WPP); else __syncwarp(); }
        if (leader && !(a.dbg & 2u)) {
            bulk_store(gst, tile, 3u * ppx * 4u);
            if (want_mask) bulk_store(gmk, tile + 3u * ppx * 4u, ppx);
            bulk_commit();
        }
        gst += gst_step;
        gmk += gmk_step;
        parity ^= 1;
    }
}

template <int WPP, int PROD_WARPS>
__device__ __forceinline__ void consumer_loop(const TiledArgs& a, const Item* s_items, uint32_t bar_full, uint32_t bar_empty, uint32_t ring,
                                              uint32_t stage0, uint32_t stage_bytes) {
    constexpr int ppx = 128 * WPP;
    constexpr uint32_t TILE_BYTES = ppx * 13;
    const int lane_id = threadIdx.x & 31;
    const int warp = (threadIdx.x >> 5) - PROD_WARPS, group = warp / WPP;
    const int sub = WPP == 1 ? 0 : (warp % WPP);
    const bool leader = sub == 0 && lane_id == 0;
    const int lane = ((lane_id >> 2) & 1) * 16 + (lane_id >> 3) * 4 + (lane_id & 3);
    const uint32_t tile0 = ring + (uint32_t)group * 2u * TILE_BYTES;
    const uint32_t my_m = 3u * ppx * 4u + (uint32_t)(sub * 128 + 4 * lane_id);
    int parity = 0, k = 0;
    for (int unit = blockIdx.x; unit < a.n_units; unit += gridDim.x) {
        const int tile = unit % a.n_tiles, jc = unit / a.n_tiles;
        const int c = jc % a.fchunks, j = jc / a.fchunks;
        const int g1 = min(a.groups, (c + 1) * a.gpc);
        // ---- once per unit: the table records of this group's patch (the tile has at most one patch per group) ----
        const FlTraj* tr = a.trajs + j;
        const int4 d = __ldg((const int4*)tr->d_tile_desc + 2 * tile);
        const bool has = group < d.w;
        const int patch = has ? __ldg(tr->d_tile_patches + d.z + group) : 0;
        const bool want_mask = tr->d_mask != nullptr;
        PixelRegs px;
        px.mbits = 0;
        if (has) {
            const int4* idx_tab = (const int4*)tr->d_idx_tile;
            const double2* w_tab = (const double2*)tr->d_w;
            const int o = patch * ppx + sub * 128 + lane;
#pragma unroll
            for (int r = 0; r < NP; ++r) {
                const int4 id = __ldg(idx_tab + o + 32 * r);
                const double2 ww = __ldg(w_tab + o + 32 * r);
                const bool out = id.w < 0;
                px.mbits |= out ? (1u << (8 * r)) : 0u;
                px.w1[r] = out ? 0.0 : ww.x;
                px.w2[r] = out ? 0.0 : ww.y;
                px.w0[r] = out ? 0.0 : 1.0 - ww.x - ww.y;
                px.ov[r][0] = out ? 0u : (uint32_t)id.x;        // already 16 * slot
                px.ov[r][1] = out ? 0u : (uint32_t)id.y;
                px.ov[r][2] = out ? 0u : (uint32_t)id.z;
            }
        }
        bool mask_stale = true;       // the tiles' mask bytes are not this patch's static mask (new patch, or a checked item ran)
        for (int g = c * a.gpc; g < g1; ++g, ++k) {
            const int b = k & 1;
            mb_wait(bar_full + 8u * b, (k >> 1) & 1);                          // the producers have staged item k
            const Item* it = &s_items[b];
            if (has && it->nf > 0 && !(a.dbg & 64u)) {
                const uint32_t stage_cur = stage0 + (uint32_t)b * stage_bytes;
                if (it->bad) {
                    run_frames<true, WPP>(px, it, patch, stage_cur, tile0, group, parity, a);
                    mask_stale = true;
                } else {
                    if (mask_stale && want_mask) {
                        if (leader) bulk_wait_read0();            // both tiles are free (this thread issued every copy that read them)
                        if (WPP > 1) named_sync(1 + group, 32 * WPP); else __syncwarp();
                        const unsigned mw = mask_word(px.mbits);
                        sts32u(tile0 + my_m, mw);
                        sts32u(tile0 + TILE_BYTES + my_m, mw);
                        mask_stale = false;
                    }
                    run_frames<false, WPP>(px, it, patch, stage_cur, tile0, group, parity, a);
                }
            }
            __syncwarp();
            if (lane_id == 0) mb_arrive(bar_empty + 8u * b);          // this warp no longer reads buffer b / item slot b
        }
    }
    bulk_wait_read0();         // no bulk copy may still be reading shared memory when the CTA exits
}

template <int WPP, int PROD_WARPS>
__global__ void __launch_bounds__(TL_THREADS, 1) k_interp_patchify_tiled(TiledArgs a) {
    constexpr int ppx = 128 * WPP;
    constexpr int CONS_WARPS = TL_WARPS - PROD_WARPS;
    constexpr uint32_t TILE_BYTES = ppx * 13;
    constexpr uint32_t RING_BYTES = (CONS_WARPS / WPP) * 2 * TILE_BYTES;
    extern __shared__ __align__(128) unsigned char fl_smem[];
    Item* s_items = (Item*)fl_smem;                                   // [2]
    const uint32_t base = smem_u32(fl_smem);
    const uint32_t bar_full = base + 2 * ITEM_BYTES, bar_empty = bar_full + 16;     // [2] each
    const uint32_t ring = base + HEAD_BYTES;
    const uint32_t stage0 = ring + ((RING_BYTES + 127u) & ~127u);
    const uint32_t stage_bytes = (uint32_t)a.TF * a.slot_rec * 16u;
    if (threadIdx.x == 0) {
        mb_init(bar_full, PROD_WARPS * 32);
        mb_init(bar_full + 8, PROD_WARPS * 32);
        mb_init(bar_empty, CONS_WARPS);
        mb_init(bar_empty + 8, CONS_WARPS);
    }
    __syncthreads();
    if ((threadIdx.x >> 5) < PROD_WARPS) producer_loop<PROD_WARPS>(a, s_items, bar_full, bar_empty, stage0, stage_bytes, ppx);
    else consumer_loop<WPP, PROD_WARPS>(a, s_items, bar_full, bar_empty, ring, stage0, stage_bytes);
}

}  // namespace

int fli::launch_tiled(const FlTraj* d_trajs, const FlTraj* h_trajs, int n_traj, int max_frames, int n_patches, int px, int py,
                      const StagedConst& sc, unsigned flags, cudaStream_t st) {
    const int ppx = px * py;
    if (!h_trajs || (ppx != 128 && ppx != 256)) return 1;
    const int wpp = ppx / 128;
    int n_tiles = h_trajs[0].n_tiles, max_nodes = 0, max_tile_patches = 0;
    for (int i = 0; i < n_traj; ++i) {
        const FlTraj& t = h_trajs[i];
        if (!t.d_idx_tile || !t.d_tile_nodes || !t.d_tile_desc || !t.d_tile_patches || t.n_tiles < 1) return 1;
        FL_REQUIRE(t.n_tiles == n_tiles, FL_E_ARG, "fl_interp_patchify: trajectory %d has %d tiles, trajectory 0 has %d", i, t.n_tiles, n_tiles);
        FL_REQUIRE(t.max_tile_nodes >= 0 && t.max_tile_patches >= 1, FL_E_ARG, "fl_interp_patchify: trajectory %d: bad tile sizes", i);
        FL_REQUIRE((uintptr_t)t.d_idx_tile % 16 == 0 && (uintptr_t)t.d_tile_desc % 16 == 0, FL_E_ALIGN,
                   "fl_interp_patchify: trajectory %d: tile tables must be 16-byte aligned", i);
        // the bulk copies write 16-byte aligned blocks; outputs placed otherwise go to the other kernels
        if ((uintptr_t)t.d_states % 16 != 0 || (t.d_mask != nullptr && (uintptr_t)t.d_mask % 16 != 0)) return 1;
        max_nodes = t.max_tile_nodes > max_nodes ? t.max_tile_nodes : max_nodes;
        max_tile_patches = t.max_tile_patches > max_tile_patches ? t.max_tile_patches : max_tile_patches;
    }
    const int slot_rec = max_nodes > 0 ? max_nodes : 1;
    // a tile has one patch per patch group of consumer warps; whatever warps that leaves are producers (the host sizes the
    // tiles: 7 patches of 256 pixels -> 14 consumer + 2 producer warps, 6 -> 12 + 4 for meshes with many nodes per pixel)
    int prod_warps = 0;
    for (int pw = 2; pw <= 6; pw += 2)
        if ((TL_WARPS - pw) / wpp >= max_tile_patches) prod_warps = pw;      // the most producers the tile size leaves room for
    if (!prod_warps) return 1;     // more patches per tile than patch groups: not a plan for this kernel
    const int cons_warps = TL_WARPS - prod_warps;
    const size_t ring = fl_align_up((size_t)(cons_warps / wpp) * 2 * ppx * 13, 128);
    const size_t fixed = HEAD_BYTES + ring;
    // Frames per item: the two record buffers take what is left of a shared-memory budget that deliberately stops short of
    // the 227 KB an SM has -- the rest stays L1 for the producers' list / node loads and the consumers' table reads
    // (measured on the 1 M-triangle mesh, 4 x 64 frames: 11 frames per item with 225 KB of shared memory 2.99 ms, 8 frames
    // with 171 KB 2.47 ms) -- and the frames of a trajectory are cut into groups of equal length (no short last item).
    long budget = 176 * 1024;
    if (const char* e = getenv("FLUIDGRID_TILED_SMEM_KB")) { long v = atol(e) * 1024; if (v >= 64 * 1024 && v <= SMEM_TOTAL) budget = v; }
    long TF = (budget - (long)fixed) / 2 / (16L * slot_rec);
    if (TF < 1) TF = ((long)SMEM_TOTAL - (long)fixed) / 2 / (16L * slot_rec);      // big tiles: whatever fits at all
    if (TF > 16) TF = 16;
    if (TF > max_frames) TF = max_frames;
    if (TF >= 1) { const long groups = (max_frames + TF - 1) / TF; TF = (max_frames + groups - 1) / groups; }
    if (const char* e = getenv("FLUIDGRID_TF")) { long v = atol(e); if (v >= 1 && v < TF) TF = v; }
    if (TF < 1) return 1;        // a tile's nodes do not fit: the caller falls back (plan smaller tiles)
    TiledArgs a;
    a.trajs = d_trajs;
    a.groups = (max_frames + (int)TF - 1) / (int)TF;
    a.n_tiles = n_tiles;
    // a unit = (trajectory, run of gpc frame groups, tile): long runs amortise the per-unit table reads, enough units keep every SM busy
    long fchunks = 1;
    while (fchunks < a.groups && (long)n_traj * n_tiles * fchunks < 8L * FL_SM_COUNT && (a.groups + fchunks) / (fchunks + 1) >= 4) ++fchunks;
    a.gpc = (int)((a.groups + fchunks - 1) / fchunks);
    a.fchunks = (a.groups + a.gpc - 1) / a.gpc;
    const long n_units = (long)n_traj * a.fchunks * n_tiles;
    FL_REQUIRE(n_units < 0x7fffffffL - 2 * FL_SM_COUNT, FL_E_ARG, "fl_interp_patchify: too many work units (%ld)", n_units);
    a.n_units = (int)n_units;
    a.TF = (int)TF;
    a.n_patches = n_patches;
    a.slot_rec = slot_rec;
    a.sc = sc;
    a.flags = flags;
    a.dbg = 0;
    if (const char* e = getenv("FLUIDGRID_DBG")) a.dbg = (unsigned)atol(e);
    const size_t smem = fixed + 2 * (size_t)TF * slot_rec * 16;
    const int grid = n_units < FL_SM_COUNT ? (int)n_units : FL_SM_COUNT;      // one persistent CTA per SM
#define FL_TILED_LAUNCH(W, P)                                                                                                   \
    do {                                                                                                                        \
        static FlOncePerDevice attr;                                                                                            \
        if (attr.first_use())                                                                                                   \
            FL_CUDA(cudaFuncSetAttribute(k_interp_patchify_tiled<W, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL)); \
        k_interp_patchify_tiled<W, P><<<grid, TL_THREADS, smem, st>>>(a);                                                        \
    } while (0)
    if (wpp == 2 && prod_warps == 2) FL_TILED_LAUNCH(2, 2);
    else if (wpp == 2 && prod_warps == 4) FL_TILED_LAUNCH(2, 4);
    else if (wpp == 2) FL_TILED_LAUNCH(2, 6);
    else if (prod_warps == 2) FL_TILED_LAUNCH(1, 2);
    else if (prod_warps == 4) FL_TILED_LAUNCH(1, 4);
    else FL_TILED_LAUNCH(1, 6);
#undef FL_TILED_LAUNCH
    FL_LAUNCH_CHECK();
    return FL_OK;
}
