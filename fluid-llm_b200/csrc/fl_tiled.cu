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
//   records    16 bytes per node and frame with u and v ALREADY in fp64 form: their high words plus one word with the
//              three non-zero bits of each low word (a float widened to double has 29 zero bits at the bottom), and p
//              as fp32.  A vertex costs one LDS.128 + two PRMT (written in place next to the high words) + one
//              F2F.F64.F32 instead of one LDS.128 + two F2F: half the load on the conversion unit per vertex.
//   producers  the last warps of the CTA only stage: they walk the CTA's items, load the tile's nodes frame by frame
//              (several frames in flight per thread), convert, write the records of item k into buffer k & 1 and arrive
//              on its `full` mbarrier; they wait on `empty` before overwriting a buffer.
//   consumers  the other warps never meet at a CTA-wide barrier.  The (item, patch) units of the CTA form one sequence
//              and are dealt round-robin to the patch groups (the warps that share a patch), so a group that is ahead
//              simply starts on the next item as soon as its buffer is full; every consumer warp arrives on `empty`
//              when it is done with an item.
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
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
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

// node record {bits, hi(u), p, hi(v)}: u and v already widened to fp64 (high words; byte 0 / 1 of `bits` = the top byte of
// their low words, whose other 29 bits are zero for a widened float), p still fp32
__device__ __forceinline__ uint4 make_record(float u, float v, float p) {
    const double du = (double)u, dv = (double)v;
    const uint32_t bits = ((uint32_t)__double2loint(du) >> 24) | (((uint32_t)__double2loint(dv) >> 24) << 8);
    return make_uint4(bits, (uint32_t)__double2hiint(du), __float_as_uint(p), (uint32_t)__double2hiint(dv));
}
// one 128-bit gather -> the three doubles.  Written on the 64-bit halves so that u and v are completed IN PLACE (one PRMT
// each writes the low word next to the high word that is already there); p takes the one F2F.F64.F32.
__device__ __forceinline__ void load_record(uint32_t addr, double& u, double& v, double& p) {
    unsigned long long A, B;        // A = {bits, hi(u)}, B = {p as float, hi(v)}
    asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(A), "=l"(B) : "r"(addr));
    const uint32_t bits = (uint32_t)A;
    p = (double)__uint_as_float((uint32_t)B);
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
    int n_items, groups, n_tiles, TF, n_patches, slot_rec;   // slot_rec: records per staged frame
    StagedConst sc;
    unsigned flags;
    unsigned dbg;     // development ablations (FLUIDGRID_DBG): 1 no proxy fence, 2 no bulk copies, 4 producers stage nothing, 8 no group barrier, 16 no L2 prefetch, 32 node-by-node staging, 64 consumers compute nothing
};

__device__ __forceinline__ void decode_item(const TiledArgs& a, int item, int ppx, Item* out) {
    Item it;
    const int tile = item % a.n_tiles;
    const int jg = item / a.n_tiles;
    const int g = jg % a.groups, j = jg / a.groups;
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
__device__ __forceinline__ void stage_quads(const Item* it, int ptid, int PT, uint32_t sbuf, int slot_rec, Scan& sc) {
    constexpr int FB = 4;
    const int qc = it->q_cnt, nf = it->nf, vstep = it->vstep, pstep = it->pstep;
    const int nfb = (nf + FB - 1) / FB;
    const float* vel = it->vel;
    const float* prs = it->prs;
    const int4* qs = it->qslots;
    const int* qid = it->quads;
    int u = ptid;
    int4 sl_n = make_int4(-1, -1, -1, -1);
    int q_n = 0;
    if (u < qc * nfb) { const int e = u % qc; sl_n = __ldg(qs + e); q_n = __ldg(qid + e); }
    for (; u < qc * nfb; u += PT) {
        const int fb = u / qc;
        const int f0 = fb * FB;
        const int4 sl = sl_n;
        const int q = q_n;
        if (u + PT < qc * nfb) { const int e = (u + PT) % qc; sl_n = __ldg(qs + e); q_n = __ldg(qid + e); }   // next entry: off the critical path
        const float* vp = vel + (size_t)f0 * vstep + 8 * (size_t)q;
        const float* pp = prs + (size_t)f0 * pstep + 4 * (size_t)q;
        float4 va[FB], vb[FB], pq[FB];
#pragma unroll
        for (int i = 0; i < FB; ++i)
            if (f0 + i < nf) {
                va[i] = ldg_stream4f(vp + (size_t)i * vstep);
                vb[i] = ldg_stream4f(vp + (size_t)i * vstep + 4);
                pq[i] = ldg_stream4f(pp + (size_t)i * pstep);
            }
        uint32_t dst = sbuf + (uint32_t)f0 * slot_rec * 16u;
#pragma unroll
        for (int i = 0; i < FB; ++i)
            if (f0 + i < nf) {
                if (sl.x >= 0) { sc.add(va[i].x, va[i].y, pq[i].x); sts128(dst + 16u * sl.x, make_record(va[i].x, va[i].y, pq[i].x)); }
                if (sl.y >= 0) { sc.add(va[i].z, va[i].w, pq[i].y); sts128(dst + 16u * sl.y, make_record(va[i].z, va[i].w, pq[i].y)); }
                if (sl.z >= 0) { sc.add(vb[i].x, vb[i].y, pq[i].z); sts128(dst + 16u * sl.z, make_record(vb[i].x, vb[i].y, pq[i].z)); }
                if (sl.w >= 0) { sc.add(vb[i].z, vb[i].w, pq[i].w); sts128(dst + 16u * sl.w, make_record(vb[i].z, vb[i].w, pq[i].w)); }
                dst += (uint32_t)slot_rec * 16u;
            }
    }
}

// Ask L2 for the node fields item `item` will stage: per frame the span of the tile's node ids (the list is ascending), one
// bulk prefetch for the velocities and one for the pressures.  Thread i of the producers takes frame i.  The spans are
// clipped to whole 16-byte units inside the frame, so nothing outside the caller's arrays is touched.
__device__ __forceinline__ void prefetch_item(const TiledArgs& a, int item, int ptid) {
    if (item >= a.n_items || ptid >= a.TF) return;
    const int tile = item % a.n_tiles;
    const int jg = item / a.n_tiles;
    const int g = jg % a.groups, j = jg / a.groups;
    const FlTraj* tr = a.trajs + j;
    const int f = g * a.TF + ptid;
    if (f >= tr->n_frames) return;
    const int4 d = __ldg((const int4*)tr->d_tile_desc + 2 * tile);
    if (d.y < 1) return;
    const int n0 = __ldg(tr->d_tile_nodes + d.x), n1 = __ldg(tr->d_tile_nodes + d.x + d.y - 1) + 1;     // [n0, n1)
    const long long t = (long long)tr->t0 + (long long)f * tr->interval;
    const uintptr_t v0 = ((uintptr_t)(tr->d_velocity + t * tr->vel_stride + 2 * (size_t)n0) + 15) & ~(uintptr_t)15;
    const uintptr_t v1 = (uintptr_t)(tr->d_velocity + t * tr->vel_stride + 2 * (size_t)n1) & ~(uintptr_t)15;
    const uintptr_t p0 = ((uintptr_t)(tr->d_pressure + t * tr->prs_stride + (size_t)n0) + 15) & ~(uintptr_t)15;
    const uintptr_t p1 = (uintptr_t)(tr->d_pressure + t * tr->prs_stride + (size_t)n1) & ~(uintptr_t)15;
    if (v1 > v0) l2_prefetch_bulk((const void*)v0, (uint32_t)(v1 - v0));
    if (p1 > p0) l2_prefetch_bulk((const void*)p0, (uint32_t)(p1 - p0));
}

template <int PROD_WARPS>
__device__ __forceinline__ void producer_loop(const TiledArgs& a, Item* s_items, uint32_t bar_full, uint32_t bar_empty, uint32_t stage0,
                                              uint32_t stage_bytes, int ppx) {
    constexpr int PT = PROD_WARPS * 32;
    const int ptid = threadIdx.x;          // the producers are the first warps of the CTA
    int k = 0;
    if (!(a.dbg & 16u)) prefetch_item(a, blockIdx.x, ptid);
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++k) {
        const int b = k & 1;
        if (!(a.dbg & 16u)) prefetch_item(a, item + gridDim.x, ptid);      // the next item's fields are in L2 by the time its buffer is free
        if (k >= 2) mb_wait(bar_empty + 8u * b, ((k >> 1) - 1) & 1);       // every consumer warp is done with item k - 2
        if (ptid == 0) decode_item(a, item, ppx, &s_items[b]);
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

// ---- consumer warps -------------------------------------------------------------------------------------------------
// One chunk (128 output pixels of one patch) x the item's frames.  WPP = warps per patch (ppx / 128); the WPP warps of a
// patch group share two output tiles and the named barrier 1 + group.
template <bool CHECKED, int WPP>
__device__ __forceinline__ void chunk_frames(const Item* cs, int unit, uint32_t stage_cur, uint32_t tile0, int group, int& parity,
                                             const TiledArgs& a) {
    constexpr int ppx = 128 * WPP;
    constexpr uint32_t TILE_BYTES = ppx * 13;       // [3][ppx] floats + ppx mask bytes
    const int lane_id = threadIdx.x & 31;
    const int sub = WPP == 1 ? 0 : ((threadIdx.x >> 5) % WPP);       // the producer warps come in pairs, so parity is kept
    const bool leader = sub == 0 && lane_id == 0;
    const int bar_id = 1 + group;
    const bool mask_aware = a.flags & FL_MASK_AWARE_NORM, no_norm = a.flags & FL_NO_NORM;
    // pixel of this lane inside a group of 32 adjacent pixels (2 patch rows x 16): the 8 lanes of a quarter-warp (one
    // LDS.128 wavefront) take a compact 2 x 4 block, which touches fewer distinct nodes than a 1 x 8 strip
    const int lane = ((lane_id >> 2) & 1) * 16 + (lane_id >> 3) * 4 + (lane_id & 3);
    // the item lives in shared memory (uniform reads); only what the frame loop needs is kept in registers
    const int patch = __ldg(cs->patches + unit);
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
    // element (c, pixel k of the patch) of a tile at c * ppx + k, mask bytes behind the floats
    const uint32_t my_f = (uint32_t)(sub * 128 + lane) * 4u, my_m = 3u * ppx * 4u + (uint32_t)(sub * 128 + 4 * lane_id);
    if (leader) bulk_wait_read0();                    // both tiles are free (this thread issued every copy that read them)
    if (WPP > 1) named_sync(bar_id, 32 * WPP); else __syncwarp();
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
    // leader only: where this patch's block of the item's first frame goes
    float* gst = cs->states + (size_t)patch * (3 * ppx);
    uint8_t* gmk = want_mask ? cs->mask + (size_t)patch * ppx : nullptr;
    const size_t gst_step = (size_t)a.n_patches * (3 * ppx), gmk_step = (size_t)a.n_patches * ppx;
    uint32_t nb = stage_cur;
#pragma unroll 1
    for (int f = 0; f < nf; ++f, nb += a.slot_rec * 16u) {
        float res[3][NP];
        unsigned fm = mbits;
#pragma unroll
        for (int r = 0; r < NP; ++r) {
            double u0, v0, p0, u1, v1, p1, u2, v2, p2;
            load_record(nb + ov[r][0], u0, v0, p0);   // one 128-bit gather per vertex
            load_record(nb + ov[r][1], u1, v1, p1);
            load_record(nb + ov[r][2], u2, v2, p2);
            res[0][r] = (float)fma(w2[r], u2, fma(w1[r], u1, w0[r] * u0));
            res[1][r] = (float)fma(w2[r], v2, fma(w1[r], v1, w0[r] * v0));
            res[2][r] = (float)fma(w2[r], p2, fma(w1[r], p1, w0[r] * p0));
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
    constexpr int CONS_WARPS = TL_WARPS - PROD_WARPS;
    constexpr int N_GROUPS = CONS_WARPS / WPP;
    constexpr uint32_t TILE_BYTES = 128 * WPP * 13;
    const int warp = (threadIdx.x >> 5) - PROD_WARPS, group = warp / WPP;
    const uint32_t tile0 = ring + (uint32_t)group * 2u * TILE_BYTES;
    int next_unit = group, unit_base = 0, parity = 0, k = 0;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++k) {
        const int b = k & 1;
        mb_wait(bar_full + 8u * b, (k >> 1) & 1);                          // the producers have staged item k
        const Item* it = &s_items[b];
        const int np = it->n_patches, bad = it->bad;
        const uint32_t stage_cur = stage0 + (uint32_t)b * stage_bytes;
        for (; next_unit < unit_base + np; next_unit += N_GROUPS) {
            if (a.dbg & 64u) continue;
            if (bad) chunk_frames<true, WPP>(it, next_unit - unit_base, stage_cur, tile0, group, parity, a);
            else chunk_frames<false, WPP>(it, next_unit - unit_base, stage_cur, tile0, group, parity, a);
        }
        unit_base += np;
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mb_arrive(bar_empty + 8u * b);          // this warp no longer reads buffer b / item slot b
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
    // producer warps: two keep up with the meshes of the reference's data sets (a node per ~9 pixels); denser meshes get four
    const double nodes_per_pixel = (double)slot_rec * n_tiles / ((double)n_patches * ppx);
    int prod_warps = nodes_per_pixel > 0.2 ? 4 : 2;
    if (const char* e = getenv("FLUIDGRID_PROD_WARPS")) { const int v = atoi(e); if (v == 2 || v == 4 || v == 6) prod_warps = v; }
    const int cons_warps = TL_WARPS - prod_warps;
    const size_t ring = fl_align_up((size_t)(cons_warps / wpp) * 2 * ppx * 13, 128);
    const size_t fixed = HEAD_BYTES + ring;
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
    a.dbg = 0;
    if (const char* e = getenv("FLUIDGRID_DBG")) a.dbg = (unsigned)atol(e);
    const size_t smem = fixed + 2 * (size_t)TF * slot_rec * 16;
    const int grid = n_items < FL_SM_COUNT ? (int)n_items : FL_SM_COUNT;      // one persistent CTA per SM
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
