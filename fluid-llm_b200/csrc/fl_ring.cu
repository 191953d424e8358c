// fl_ring.cu -- the per-step kernel for meshes whose node fields of a frame fit shared memory several times over
// (Cylinder / Airfoil / EAGLE-shaped: a few thousand nodes): gather -> fp64 FMA -> fp32 -> mask -> normalise -> patchify.
//
// Same work and the same arithmetic as k_interp_patchify_staged / _tiled (replaces simple_dataloader.py:104-152,166-216 and
// airfoil_ds.py:71-139,216-244 of the reference).  What is different is how the node fields reach the gathers and how the
// results leave, because that is what bounded the other two kernels (profiles/README.md): register-staged loads behind a
// CTA barrier and scalar global stores through the LSU (staged), or a handful of producer threads with too few bytes in
// flight and a fence + barrier + bulk-copy issue inside every consumer warp's frame loop (tiled).
//
//   tiles      as in fl_tiled.cu: a CTA owns one tile (<= 6 patches of 256 pixels, one per pair of consumer warps) for a
//              run of frames and keeps the table records of its pixels in registers for the whole run.
//   frame ring whole frames of the trajectory, exactly as they lie in HBM (velocity [N][2], pressure [N]), are bulk-copied
//              (cp.async.bulk, two copies per frame, completion on an mbarrier) into a ring of D raw slots: no registers,
//              no LSU instructions, D - 1 frames (tens of KB) in flight per SM.  The tiles of a frame run on neighbouring
//              SMs at the same time, so the frame leaves HBM once and the other tiles' copies hit L2.
//   converters two service warps turn the tile's nodes (a few hundred of the frame's thousands; their raw offsets live in
//              registers for the whole run) into 16-byte records with u, v AND p already widened to fp64 -- the three high
//              words plus one word holding the top byte of each low word (a widened float has no other low bits) -- in a ring
//              of record slots, and scan them for non-finite / huge values (such a frame is redone on the checked path).
//   consumers  12 warps, three per scheduler; per frame a lane gathers 3 vertices x 4 pixels with one LDS.128 each, rebuilds
//              the doubles with PRMT (no conversion instruction on the input side), 3 DMUL + 6 DFMA + 3 F2F per pixel,
//              normalises on the packed fp32 pipe and writes its results into the CTA's output tile (conflict-free STS.32).
//              A consumer warp's whole synchronisation per frame is one mbarrier wait (`go`) and one arrive (`done`).
//   load/store one service thread keeps D - 1 frame loads in flight; another waits for `done`, makes the tile visible to the
//              async proxy and sends it off as bulk copies (one per run of consecutive patch ids: 3 KB of states and 256 B of
//              mask per patch): no global store goes through the LSU, and fence / copy issue / drain are off the consumers'
//              critical path.  It arrives on `go` of the frame that will reuse the tile once the copies have read it.
// STATUS: experimental, off by default (FL_FORCE_RING / FLUIDGRID_RING=1).  Bit-identical to the staged kernel, but measured
// SLOWER on B200 (airfoil bench: 0.95 - 1.03 ms per launch against 0.77): profiles/README.md has the ablations that say why
// (12 free-running consumer warps alone need 0.66 ms, of which the gathers 0.55; hand-overs and the service warps add the rest).
// HBM traffic per frame: 12 N read (+ L2 hits for the other tiles), 12 P (+ P mask) written: the algorithmic bytes.
#include "fl_interp.cuh"
#include <stdlib.h>

using flg::finite_f;
using fli::StagedConst;
using fli::norm_fast;
using fli::norm_fast2;
using fli::pack2;

namespace {

constexpr int RG_THREADS = 512;
constexpr int RG_WARPS = RG_THREADS / 32;
constexpr int RG_SVC = 4;                     // service warps, one per scheduler: warp 0 = store thread, warp 1 = load thread, 2..3 = converters
constexpr int RG_CONV = RG_SVC - 2;
constexpr int RG_CONS = RG_WARPS - RG_SVC;    // consumer warps: 12, three per scheduler (the schedulers take warp % 4)
constexpr int NP = 4;                         // pixels per consumer lane
constexpr int SMEM_TOTAL = 227 * 1024;
constexpr int HEAD_BYTES = 512;               // mbarriers + flags
constexpr int MAX_D = 4;                      // raw frame slots
constexpr int MAX_NS = 8;                     // record slots (and `go` / `done` barriers): the converters run up to NS - 1 frames ahead
constexpr int MAX_NT = 8;                     // output tiles: frame k goes to tile k % NT
constexpr int CONV_NPT = 7;                   // tile nodes a converter thread keeps in registers (64 threads: 448 nodes)
constexpr int CONV_T = RG_CONV * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst), "l"(gsrc),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mb_init(uint32_t bar, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mb_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint32_t bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nRG_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra RG_DONE;\nbra RG_WAIT;\nRG_DONE:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts32u(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32u(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// node record {bits, hi(u), hi(p), hi(v)}: the three values widened to fp64; byte k of `bits` = the top byte of the k-th
// value's low word (u, v, p), whose other 24 bits are zero for a widened float (normal, denormal, inf or NaN alike)
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

struct RingArgs {
    const FlTraj* trajs;
    int n_units, n_tiles, fblocks, FB;     // unit = (trajectory, block of FB frames, tile), tile fastest
    int n_patches, rec_bytes;              // bytes of one record slot (16 x the largest tile's node count, rounded up to 128)
    int raw_bytes, raw_prs_off;            // bytes of one raw slot, and where the pressures start inside it
    int out_bytes;                         // bytes of one output tile: 13 * ppx per patch of the largest tile
    int max_tile_patches;                  // patches of the largest tile: the mask bytes of a tile start at 12 * ppx * this
    int D, NS, NT;                         // raw ring depth, record slots (= go / done barriers), output tiles
    StagedConst sc;
    unsigned flags;
    unsigned dbg;     // development ablations (FLUIDGRID_DBG): 2 no bulk stores, 4 nothing is loaded or converted, 8 (with 4) nobody waits for anybody, 64 consumers compute nothing
};

struct Unit { int j, tile, f0, nf; };
__device__ __forceinline__ Unit decode_unit(const RingArgs& a, int unit) {
    Unit u;
    u.tile = unit % a.n_tiles;
    const int jc = unit / a.n_tiles;
    const int fb = jc % a.fblocks;
    u.j = jc / a.fblocks;
    u.f0 = fb * a.FB;
    u.nf = max(0, min(a.FB, __ldg(&a.trajs[u.j].n_frames) - u.f0));
    return u;
}

// shared addresses of the barriers; [slot] = + 8 * slot (bad: + 4 * slot)
//   raw_full[D]  tx-count: the frame's two bulk loads have landed            raw_empty[D]  3 converter warps are done reading it
//   go[NS]       3 converter warps wrote the records of frame k + the store thread freed output tile k % NS
//   done[NS]     12 consumer warps finished frame k: the record slot is free, the output tile is full
struct Bars { uint32_t raw_full, raw_empty, go, done, bad; };

// running scan of the converted values: a non-finite or huge value sends the frame down the checked path
struct Scan {
    float nanacc = 0.f, amax = 0.f;
    __device__ __forceinline__ void add(float a, float b, float c) {
        nanacc = fmaf(a, 0.f, fmaf(b, 0.f, fmaf(c, 0.f, nanacc)));
        amax = fmaxf(fmaxf(fabsf(a), fabsf(b)), fmaxf(fabsf(c), amax));
    }
    __device__ __forceinline__ int bad() const { return !(nanacc == 0.f) || amax > 1.0e30f; }
};

// ---- service warp 0, lane 0: the store thread ---------------------------------------------------------------------------
struct FrameWalk {      // walks the frames of the CTA's units in order
    int unit, f, nf;
    bool more;
    Unit u;
    __device__ __forceinline__ void open(const RingArgs& a) {     // position on the first frame of `unit` or of the next unit that has one
        more = false;
        while (unit < a.n_units) {
            u = decode_unit(a, unit);
            if (u.nf > 0) { f = 0; nf = u.nf; more = true; return; }
            unit += gridDim.x;
        }
    }
    __device__ __forceinline__ bool step(const RingArgs& a) {     // -> true if a new unit was opened
        if (++f < nf) return false;
        unit += gridDim.x;
        open(a);
        return true;
    }
};

// An output tile holds the states of the tile's patches one after the other ([patch][3][ppx] floats), then their mask bytes
// ([patch][ppx]); the patches of a tile are listed with ascending ids, so a run of consecutive ids is one contiguous block of
// a frame's states (and of its mask) in HBM and leaves as ONE bulk copy.
template <int WPP>
__device__ __forceinline__ void store_loop(const RingArgs& a, const Bars& b, uint32_t out0) {
    constexpr int ppx = 128 * WPP;
    constexpr int MAXP = RG_CONS / WPP;
    FrameWalk sw;
    sw.unit = blockIdx.x;
    sw.open(a);
    if (a.dbg & 8u) return;          // ablation: consumers run free
    float* gst = nullptr;
    uint8_t* gmk = nullptr;
    int run_first[MAXP], run_len[MAXP], run_pos[MAXP];    // runs of consecutive patch ids: first id, length, position in the tile
    int n_runs = 0;
    const size_t gst_step = (size_t)a.n_patches * (3 * ppx), gmk_step = (size_t)a.n_patches * ppx;
    const uint32_t mask_off = (uint32_t)a.max_tile_patches * (12u * ppx);
    auto s_setup = [&]() {
        const FlTraj* tr = a.trajs + sw.u.j;
        const int4 desc = __ldg((const int4*)tr->d_tile_desc + 2 * sw.u.tile);
        n_runs = 0;
        int prev = -2;
#pragma unroll
        for (int g = 0; g < MAXP; ++g) {
            run_first[g] = 0; run_len[g] = 0; run_pos[g] = 0;
        }
#pragma unroll
        for (int g = 0; g < MAXP; ++g)
            if (g < desc.w) {
                const int l = __ldg(tr->d_tile_patches + desc.z + g);
                if (l == prev + 1) {
#pragma unroll
                    for (int q = 0; q < MAXP; ++q) if (q == n_runs - 1) run_len[q] += 1;
                } else {
#pragma unroll
                    for (int q = 0; q < MAXP; ++q) if (q == n_runs) { run_first[q] = l; run_len[q] = 1; run_pos[q] = g; }
                    ++n_runs;
                }
                prev = l;
            }
        gst = tr->d_states + (size_t)sw.u.f0 * a.n_patches * (3 * ppx);
        gmk = tr->d_mask ? tr->d_mask + (size_t)sw.u.f0 * a.n_patches * ppx : nullptr;
    };
    if (sw.more) s_setup();
    int s_s = 0, ph_s = 0, t_s = 0;      // `done` slot of the next frame to store + its parity; the frame's output tile
    const int NS = a.NS, NT = a.NT;
    int s_go = (NT - 1) % NS;            // `go` slot the store thread arrives on next (frames 0 .. NT-2 find their tiles free)
    for (int i = 0; i < NT - 1; ++i) mb_arrive(b.go + 8u * i);
    while (sw.more) {
        mb_wait(b.done + 8u * s_s, (unsigned)ph_s);
        // every consumer warp has written frame k into tile s_s: generic-proxy writes -> visible to the bulk copies
        fence_async_smem();
        if (!(a.dbg & 2u)) {
            const uint32_t tile = out0 + (uint32_t)t_s * a.out_bytes;
#pragma unroll
            for (int q = 0; q < MAXP; ++q)
                if (q < n_runs) {
                    bulk_store(gst + (size_t)run_first[q] * (3 * ppx), tile + (uint32_t)run_pos[q] * (12u * ppx), (uint32_t)run_len[q] * (12u * ppx));
                    if (gmk) bulk_store(gmk + (size_t)run_first[q] * ppx, tile + mask_off + (uint32_t)run_pos[q] * ppx, (uint32_t)run_len[q] * ppx);
                }
            bulk_commit();
        }
        bulk_wait_read<1>();             // the copies of frame k - 1 have read their tile: frame k - 1 + NT may be written into it
        mb_arrive(b.go + 8u * s_go);
        if (++s_go == NS) s_go = 0;
        gst += gst_step;
        if (gmk) gmk += gmk_step;
        if (++s_s == NS) { s_s = 0; ph_s ^= 1; }
        if (++t_s == NT) t_s = 0;
        if (sw.step(a) && sw.more) s_setup();
    }
    bulk_wait_read<0>();         // no bulk copy may still be reading shared memory when the CTA exits
}

// ---- service warps 1..3: converters; lane 0 of warp 1 also issues the frame loads ------------------------------------------
struct Loader {
    FrameWalk w;
    const float* v;
    const float* p;
    long long vstep, pstep;
    uint32_t vbytes, pbytes;
    int k, d, ph;                 // frames issued; slot of the next one; parity of the `empty` phase that frees it (from k >= D on)
    __device__ __forceinline__ void setup(const RingArgs& a) {
        const FlTraj tr = a.trajs[w.u.j];
        const long long t = (long long)tr.t0 + (long long)w.u.f0 * tr.interval;
        v = tr.d_velocity + t * tr.vel_stride;
        p = tr.d_pressure + t * tr.prs_stride;
        vstep = (long long)tr.interval * tr.vel_stride;
        pstep = (long long)tr.interval * tr.prs_stride;
        vbytes = 8u * (uint32_t)tr.prs_stride;
        pbytes = 4u * (uint32_t)tr.prs_stride;
    }
    __device__ __forceinline__ void issue(const RingArgs& a, const Bars& b, uint32_t raw0) {
        const uint32_t bar = b.raw_full + 8u * d, dst = raw0 + (uint32_t)d * a.raw_bytes;
        mb_expect_tx(bar, vbytes + pbytes);
        bulk_load(dst, v, vbytes, bar);
        bulk_load(dst + a.raw_prs_off, p, pbytes, bar);
        v += vstep; p += pstep;
        ++k;
        if (++d == a.D) { d = 0; if (k > a.D) ph ^= 1; }
        if (w.step(a) && w.more) setup(a);
    }
};

// service warp 1, lane 0: keeps D - 1 frames of the CTA's frame sequence on their way
__device__ __forceinline__ void load_loop(const RingArgs& a, const Bars& b, uint32_t raw0) {
    if (a.dbg & 4u) return;
    Loader ld;
    ld.w.unit = blockIdx.x; ld.k = 0; ld.d = 0; ld.ph = 0;
    ld.w.open(a);
    if (ld.w.more) ld.setup(a);
    while (ld.w.more) {
        if (ld.k >= a.D) mb_wait(b.raw_empty + 8u * ld.d, (unsigned)ld.ph);      // the converters are done with the frame D back
        ld.issue(a, b, raw0);
    }
}

__device__ __forceinline__ void convert_loop(const RingArgs& a, const Bars& b, uint32_t raw0, uint32_t rec0) {
    const int ptid = threadIdx.x - 64, lane_id = threadIdx.x & 31, cw = (threadIdx.x >> 5) - 2;
    const bool idle = a.dbg & 4u;
    const int NS = a.NS;
    int k = 0, d = 0, dph = 0, s = 0, sph = 0;      // frame counter; raw slot + its `full` parity; record slot + the parity of its `go` phase
    for (int unit = blockIdx.x; unit < a.n_units; unit += gridDim.x) {
        const Unit u = decode_unit(a, unit);
        if (u.nf <= 0) continue;
        const FlTraj* tr = a.trajs + u.j;
        const int4 desc = __ldg((const int4*)tr->d_tile_desc + 2 * u.tile);
        const int* nodes = tr->d_tile_nodes + desc.x;
        const int S = desc.y;
        // raw offsets of this thread's nodes, for the whole run of frames
        uint32_t voff[CONV_NPT], poff[CONV_NPT];
#pragma unroll
        for (int i = 0; i < CONV_NPT; ++i) {
            const int q = ptid + CONV_T * i;
            const int n = q < S ? __ldg(nodes + q) : 0;
            voff[i] = 8u * (uint32_t)n;
            poff[i] = (uint32_t)a.raw_prs_off + 4u * (uint32_t)n;
        }
        for (int f = 0; f < u.nf; ++f, ++k) {
            const uint32_t raw = raw0 + (uint32_t)d * a.raw_bytes, rec = rec0 + (uint32_t)s * a.rec_bytes;
            if (!idle) mb_wait(b.raw_full + 8u * d, (unsigned)dph);
            if (k >= NS && !(a.dbg & 8u)) mb_wait(b.done + 8u * s, (unsigned)(sph ^ 1));        // every consumer warp is done with the frame k - NS
            Scan sc;
            if (!idle) {
                float2 uv[CONV_NPT];
                float pp[CONV_NPT];
#pragma unroll
                for (int i = 0; i < CONV_NPT; ++i)
                    if (ptid + CONV_T * i < S) {
                        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(uv[i].x), "=f"(uv[i].y) : "r"(raw + voff[i]));
                        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pp[i]) : "r"(raw + poff[i]));
                    }
#pragma unroll
                for (int i = 0; i < CONV_NPT; ++i)
                    if (ptid + CONV_T * i < S) {
                        sc.add(uv[i].x, uv[i].y, pp[i]);
                        sts128(rec + 16u * (uint32_t)(ptid + CONV_T * i), make_record(uv[i].x, uv[i].y, pp[i]));
                    }
                for (int q = ptid + CONV_T * CONV_NPT; q < S; q += CONV_T) {          // tiles with more nodes than the register lists hold
                    const int n = __ldg(nodes + q);
                    float2 w; float pq;
                    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(w.x), "=f"(w.y) : "r"(raw + 8u * (uint32_t)n));
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pq) : "r"(raw + (uint32_t)a.raw_prs_off + 4u * (uint32_t)n));
                    sc.add(w.x, w.y, pq);
                    sts128(rec + 16u * (uint32_t)q, make_record(w.x, w.y, pq));
                }
            }
            const int bad = __any_sync(0xffffffffu, sc.bad()) || !a.sc.fast_div;
            if (lane_id == 0) {
                sts8(b.bad + 4u * s + cw, bad ? 1u : 0u);
                mb_arrive(b.go + 8u * s);               // release: this warp's records and its flag are visible to whoever waits
                if (!idle) mb_arrive(b.raw_empty + 8u * d);
            }
            if (++d == a.D) { d = 0; dph ^= 1; }
            if (++s == NS) { s = 0; sph ^= 1; }
        }
    }
}

// ---- consumer warps --------------------------------------------------------------------------------------------------
struct PixelRegs {
    uint32_t ov[NP][3];        // byte offsets of the three vertices' records inside a record slot
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

struct NormRegs { unsigned long long nm[3], ns[3], rc[3]; };

// one frame of this warp's 128 pixels: gathers from the record slot `rec`, results to `dst` (this lane's first pixel of
// channel 0 in the output tile; channels CH_STRIDE bytes apart), mask word to `mdst`
template <bool CHECKED, int CH_STRIDE>
__device__ __forceinline__ void frame_compute(const PixelRegs& px, uint32_t rec, uint32_t dst, uint32_t mdst, unsigned mword,
                                              const NormRegs& nr, const StagedConst& sc, bool mask_aware, bool no_norm) {
    float res[3][NP];
    unsigned fm = px.mbits;
#pragma unroll
    for (int r = 0; r < NP; ++r) {
        double u0, v0, p0, u1, v1, p1, u2, v2, p2;
        load_record(rec + px.ov[r][0], u0, v0, p0);
        load_record(rec + px.ov[r][1], u1, v1, p1);
        load_record(rec + px.ov[r][2], u2, v2, p2);
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
                for (int r = 0; r < NP; r += 2) norm_fast2(res[c][r], res[c][r + 1], nr.nm[c], nr.ns[c], nr.rc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < NP; ++r) sts32(dst + (uint32_t)(c * CH_STRIDE + 128 * r), res[c][r]);
    sts32u(mdst, CHECKED ? mask_word(fm) : mword);
}

template <int WPP>
__device__ __forceinline__ void consumer_loop(const RingArgs& a, const Bars& b, uint32_t out0, uint32_t rec0) {
    constexpr int ppx = 128 * WPP;
    const int lane_id = threadIdx.x & 31;
    const int warp = (threadIdx.x >> 5) - RG_SVC, group = warp / WPP;
    const int sub = WPP == 1 ? 0 : (warp % WPP);
    const int lane = ((lane_id >> 2) & 1) * 16 + (lane_id >> 3) * 4 + (lane_id & 3);     // 2 x 4 pixel block per quarter-warp
    // this lane's places in an output tile: [patch][3][ppx] floats, then [patch][ppx] mask bytes
    const uint32_t my_f = (uint32_t)group * (12u * ppx) + (uint32_t)(sub * 128 + lane) * 4u;
    const uint32_t my_m = (uint32_t)a.max_tile_patches * (12u * ppx) + (uint32_t)group * ppx + (uint32_t)(sub * 128 + 4 * lane_id);
    const bool mask_aware = a.flags & FL_MASK_AWARE_NORM, no_norm = a.flags & FL_NO_NORM;
    const bool skip = a.dbg & 64u;
    const StagedConst sc = a.sc;
    NormRegs nr;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        nr.nm[c] = pack2(-sc.mean[c], -sc.mean[c]);
        nr.ns[c] = pack2(-sc.stdv[c], -sc.stdv[c]);
        nr.rc[c] = pack2(sc.rcp[c], sc.rcp[c]);
    }
    const int NS = a.NS, NT = a.NT;
    int s = 0, sph = 0, t = 0;
    for (int unit = blockIdx.x; unit < a.n_units; unit += gridDim.x) {
        const Unit u = decode_unit(a, unit);
        if (u.nf <= 0) continue;
        // ---- once per unit: the table records of this warp's 128 pixels (the tile has at most one patch per group) ----
        const FlTraj* tr = a.trajs + u.j;
        const int4 desc = __ldg((const int4*)tr->d_tile_desc + 2 * u.tile);
        const bool has = group < desc.w && !skip;
        const int patch = has ? __ldg(tr->d_tile_patches + desc.z + group) : 0;
        PixelRegs px;
        px.mbits = 0;
        if (has) {
            const int4* idx_tab = (const int4*)tr->d_idx_tile;
            const double2* w_tab = (const double2*)tr->d_w;
            const int o = patch * ppx + sub * 128 + lane;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int4 id = __ldg(idx_tab + o + 32 * q);
                const double2 ww = __ldg(w_tab + o + 32 * q);
                const bool out = id.w < 0;
                px.mbits |= out ? (1u << (8 * q)) : 0u;
                px.w1[q] = out ? 0.0 : ww.x;
                px.w2[q] = out ? 0.0 : ww.y;
                px.w0[q] = out ? 0.0 : 1.0 - ww.x - ww.y;
                px.ov[q][0] = out ? 0u : (uint32_t)id.x;        // already 16 * slot
                px.ov[q][1] = out ? 0u : (uint32_t)id.y;
                px.ov[q][2] = out ? 0u : (uint32_t)id.z;
            }
        }
        const unsigned mword = mask_word(px.mbits);
#pragma unroll 1
        for (int f = 0; f < u.nf; ++f) {
            if (!(a.dbg & 8u)) mb_wait(b.go + 8u * s, (unsigned)sph);           // records of frame k are in slot s and output tile s is free
            if (has) {
                const uint32_t rec = rec0 + (uint32_t)s * a.rec_bytes, tile = out0 + (uint32_t)t * a.out_bytes;
                const uint32_t bad = lds32u(b.bad + 4u * s);         // needed only after the frame: the load's latency is hidden
                frame_compute<false, ppx * 4>(px, rec, tile + my_f, tile + my_m, mword, nr, sc, mask_aware, no_norm);
                if (bad) frame_compute<true, ppx * 4>(px, rec, tile + my_f, tile + my_m, mword, nr, sc, mask_aware, no_norm);   // rare: redo
            }
            __syncwarp();
            if (lane_id == 0 && !(a.dbg & 8u)) mb_arrive(b.done + 8u * s);    // release: the warp's tile writes; its gathers of slot s have returned
            if (++s == NS) { s = 0; sph ^= 1; }
            if (++t == NT) t = 0;
        }
    }
}

template <int WPP>
__global__ void __launch_bounds__(RG_THREADS, 1) k_interp_patchify_ring(RingArgs a) {
    extern __shared__ __align__(128) unsigned char fl_smem[];
    const uint32_t base = smem_u32(fl_smem);
    Bars b;
    b.raw_full = base;
    b.raw_empty = base + 8u * MAX_D;
    b.go = base + 16u * MAX_D;
    b.done = b.go + 8u * MAX_NS;
    b.bad = b.done + 8u * MAX_NS;
    const uint32_t out0 = base + HEAD_BYTES;
    const uint32_t rec0 = out0 + (uint32_t)a.NT * a.out_bytes;
    const uint32_t raw0 = rec0 + (uint32_t)a.NS * a.rec_bytes;
    if (threadIdx.x == 0) {
        for (int i = 0; i < MAX_D; ++i) { mb_init(b.raw_full + 8u * i, 1); mb_init(b.raw_empty + 8u * i, RG_CONV); }
        for (int i = 0; i < MAX_NS; ++i) {
            mb_init(b.go + 8u * i, RG_CONV + 1);
            mb_init(b.done + 8u * i, RG_CONS);
            sts32u(b.bad + 4u * i, 0u);         // byte w of word s: converter warp w found the frame in slot s unfit for the fast path
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { if (threadIdx.x == 0) store_loop<WPP>(a, b, out0); }
    else if (warp == 1) { if (threadIdx.x == 32) load_loop(a, b, raw0); }
    else if (warp < RG_SVC) convert_loop(a, b, raw0, rec0);
    else consumer_loop<WPP>(a, b, out0, rec0);
}

}  // namespace

// returns FL_OK, an error, or 1 ("not applicable": the caller picks another kernel)
int fli::launch_ring(const FlTraj* d_trajs, const FlTraj* h_trajs, int n_traj, int max_frames, int n_patches, int px, int py,
                     const StagedConst& sc, unsigned flags, cudaStream_t st) {
    const int ppx = px * py;
    if (!h_trajs || (ppx != 128 && ppx != 256)) return 1;
    const int wpp = ppx / 128;
    const int n_tiles = h_trajs[0].n_tiles;
    int max_nodes = 0, max_tile_patches = 0, ps_max = 0;
    for (int i = 0; i < n_traj; ++i) {
        const FlTraj& t = h_trajs[i];
        if (!t.d_idx_tile || !t.d_tile_nodes || !t.d_tile_desc || !t.d_tile_patches || t.n_tiles < 1) return 1;
        if (t.n_tiles != n_tiles || t.max_tile_nodes < 0 || t.max_tile_patches < 1) return 1;      // launch_tiled reports these
        // whole frames are bulk-copied: 16-byte aligned frames whose pitch covers the padded node count
        if (t.vel_stride % 4 || t.prs_stride % 4 || t.vel_stride < 2 * t.prs_stride || t.prs_stride < t.n_nodes) return 1;
        if ((uintptr_t)t.d_velocity % 16 || (uintptr_t)t.d_pressure % 16 || (uintptr_t)t.d_idx_tile % 16 || (uintptr_t)t.d_tile_desc % 16) return 1;
        if ((uintptr_t)t.d_states % 16 != 0 || (t.d_mask != nullptr && (uintptr_t)t.d_mask % 16 != 0)) return 1;
        max_nodes = t.max_tile_nodes > max_nodes ? t.max_tile_nodes : max_nodes;
        max_tile_patches = t.max_tile_patches > max_tile_patches ? t.max_tile_patches : max_tile_patches;
        ps_max = t.prs_stride > ps_max ? t.prs_stride : ps_max;
    }
    if (max_tile_patches > RG_CONS / wpp) return 1;      // a plan for the node-list kernel (7 patches per tile): that kernel takes it
    RingArgs a;
    a.trajs = d_trajs;
    a.rec_bytes = (int)fl_align_up(16 * (size_t)(max_nodes > 0 ? max_nodes : 1), 128);
    a.out_bytes = (int)fl_align_up((size_t)max_tile_patches * 13 * ppx, 128);
    a.max_tile_patches = max_tile_patches;
    a.raw_bytes = 12 * ps_max;              // a multiple of 16 (ps_max is a multiple of 4)
    a.raw_prs_off = 8 * ps_max;
    a.NS = 4; a.NT = 4;
    if (const char* e = getenv("FLUIDGRID_RING_NS")) { int v = atoi(e); if (v >= 2 && v <= MAX_NS) a.NS = v; }
    if (const char* e = getenv("FLUIDGRID_RING_NT")) { int v = atoi(e); if (v >= 2 && v <= MAX_NT) a.NT = v; }
    if (a.NT > a.NS) a.NS = a.NT;           // a frame's `go` slot must not come round again before its tile does
    const long fixed = HEAD_BYTES + (long)a.NT * a.out_bytes + (long)a.NS * a.rec_bytes;
    long D = (SMEM_TOTAL - fixed) / a.raw_bytes;
    if (D > MAX_D) D = MAX_D;
    if (const char* e = getenv("FLUIDGRID_RING_D")) { long v = atol(e); if (v >= 2 && v < D) D = v; }
    if (D < 2) return 1;                    // the frames of this mesh are too large for a ring: the tiled kernel stages node lists instead
    a.D = (int)D;
    a.n_tiles = n_tiles;
    a.n_patches = n_patches;
    // frame blocks: runs long enough to amortise the per-unit table reads (>= 24 frames), and a unit count that fills the
    // 148 persistent CTAs evenly (each takes units b, b + 148, ...)
    long best_m = 1;
    double best_eff = -1.0;
    const long per_block = (long)n_traj * n_tiles;
    for (long m = 1; m <= max_frames; ++m) {
        const long FB = (max_frames + m - 1) / m;
        if (m > 1 && FB < 24) break;
        const long units = per_block * ((max_frames + FB - 1) / FB);
        const long waves = (units + FL_SM_COUNT - 1) / FL_SM_COUNT;
        const double eff = (double)units / (double)(waves * FL_SM_COUNT) - 0.02 * (24.0 / FB);      // fill, less the per-unit overhead
        if (eff > best_eff + 1e-9) { best_eff = eff; best_m = m; }
    }
    a.FB = (int)((max_frames + best_m - 1) / best_m);
    if (const char* e = getenv("FLUIDGRID_RING_FB")) { long v = atol(e); if (v >= 1) a.FB = (int)v; }
    a.fblocks = (max_frames + a.FB - 1) / a.FB;
    const long n_units = per_block * a.fblocks;
    FL_REQUIRE(n_units < 0x7fffffffL - 2 * FL_SM_COUNT, FL_E_ARG, "fl_interp_patchify: too many work units (%ld)", n_units);
    a.n_units = (int)n_units;
    a.sc = sc;
    a.flags = flags;
    a.dbg = 0;
    if (const char* e = getenv("FLUIDGRID_DBG")) a.dbg = (unsigned)atol(e);
    const size_t smem = (size_t)fixed + (size_t)a.D * a.raw_bytes;
    const int grid = n_units < FL_SM_COUNT ? (int)n_units : FL_SM_COUNT;      // one persistent CTA per SM
#define FL_RING_LAUNCH(W)                                                                                                    \
    do {                                                                                                                     \
        static FlOncePerDevice attr;                                                                                         \
        if (attr.first_use())                                                                                                \
            FL_CUDA(cudaFuncSetAttribute(k_interp_patchify_ring<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL)); \
        k_interp_patchify_ring<W><<<grid, RG_THREADS, smem, st>>>(a);                                                         \
    } while (0)
    if (wpp == 2) FL_RING_LAUNCH(2);
    else FL_RING_LAUNCH(1);
#undef FL_RING_LAUNCH
    FL_LAUNCH_CHECK();
    return FL_OK;
}
