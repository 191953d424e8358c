// fl_embed.cu -- the path's only dense contraction: the patch-embedding projection on the 5th-gen
// tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
// Replaces PatchEmbeddings.forward + MLP.forward + the positional-embedding add of InputEmbeddings.forward
// (src/models/layers/patch_encoder.py:23-30, MLP.py:48-54, input_embeddings.py:36-52,
// positional_encodings/positional_embeddings.py:32-37) as the reference runs them under bf16 autocast:
//   h   = LeakyReLU_0.01( bf16( x_bf16 @ W1^T + b1 ) )                 tokens x 768 -> 512
//   out = bf16( h @ W2^T + b2 ) + (x_emb[p0] + y_emb[p1] + t_emb[p2])  512 -> llm_dim, fp32 result
// One kernel, C[M,N] = epilogue(A[M,K] . B[N,K]^T), used twice; nn.Linear weights are [N,K] K-major already.
//
// Kernel anatomy (persistent: one 320-thread CTA per SM walks 128 x BLOCK_N output tiles):
//   warp 0     TMA producer: cp.async.bulk.tensor.2d of the A (128 x 64) and B (BLOCK_N x 64) bf16 tiles,
//              128-byte swizzle, into a 4-stage shared-memory ring that keeps running across tiles (full/empty mbarriers)
//   warp 1     MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = BLOCK_N,
//              K = 16) four times per stage into one of two TMEM accumulator buffers; tcgen05.commit releases the
//              stage / hands the finished accumulator to the epilogue (tmem_full), which hands it back (tmem_empty)
//   warps 2-9  epilogue (two warps per TMEM lane quarter, half of the columns each): tcgen05.ld (32 lanes x 32
//              columns per call) -> bias, bf16 rounding, LeakyReLU or positional-embedding add -> global; runs under
//              the main loop of the next tile
#include "fl_common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>

namespace {

constexpr int BM = 128, BK = 64, GEMM_THREADS = 320, STAGES = 4, EPI_WARPS = 8, EPI_SCRATCH = 4096 + 192;
// 4 stages x 48 KB operand ring + 8 x 4.2 KB epilogue scratch + barriers = 226 KB of shared memory: one CTA per SM

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t* b, unsigned n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mb_expect_tx(uint64_t* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint64_t* b, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nFLG_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra FLG_DONE;\nbra FLG_WAIT;\nFLG_DONE:\n}\n" ::"r"(s_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* map, uint64_t* bar, int c_inner, int c_outer) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s_u32(smem)), "l"(map), "r"(s_u32(bar)), "r"(c_inner), "r"(c_outer) : "memory");
}
// shared-memory matrix descriptor (sm_100): K-major tile of 64 bf16 (= one 128-byte swizzle atom) per row,
// rows densely packed, 8-row groups 1024 bytes apart (SBO), version 1, SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc(const void* smem) {
    uint64_t d = 0;
    d |= (uint64_t)((s_u32(smem) & 0x3ffff) >> 4);           // start address  [0,14)
    d |= (uint64_t)0 << 16;                                    // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                          // stride byte offset [32,46)
    d |= (uint64_t)1 << 46;                                    // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                                    // layout type: SWIZZLE_128B
    return d;
}
// instruction descriptor: D = f32, A = B = bf16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
        "%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct Epilogue {
    const float* bias;        // [N] fp32 (already rounded to bf16 values, as autocast casts the bias)
    int leaky;                // 1: LeakyReLU(0.01) on the bf16-rounded value, store bf16
    void* out;                // leaky ? bf16 [M,N] : fp32 [M,N]
    const float* x_emb;       // optional positional tables [*, N] fp32 (NULL: none)
    const float* y_emb;
    const float* t_emb;
    const long long* pos_ids; // [M,3] int64
    int max_x, max_y, max_t;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_gemm_tcgen05(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int M, int N, int K, Epilogue ep) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    unsigned char* sa = smem;                                   // [STAGES][128 x 64 bf16], 1024-byte aligned tiles
    unsigned char* sb = smem + STAGES * A_BYTES;                // [STAGES][BN x 64 bf16]
    unsigned char* scratch = smem + STAGES * (A_BYTES + B_BYTES);   // epilogue: 8 warps x (4 KB transpose tile + 192 B position ids)
    uint64_t* full = (uint64_t*)(scratch + EPI_WARPS * EPI_SCRATCH);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;                       // [2] accumulator buffer complete
    uint64_t* tmem_empty = tmem_full + 2;                       // [2] accumulator buffer drained by the epilogue
    uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = K / BK, n_tiles = N / BN, tiles = n_tiles * ((M + BM - 1) / BM);
    const uint32_t tmem_cols = tiles > (int)gridDim.x ? 2 * BN : BN;      // a second accumulator buffer only if this CTA gets a second tile
    // persistent: CTA b takes tiles b, b + grid, ...; the N tiles of one row block are consecutive tile ids, so they run on
    // neighbouring SMs at the same time and the row block's A tile is fetched from HBM once

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], 1); }
        for (int i = 0; i < 2; ++i) { mb_init(&tmem_full[i], 1); mb_init(&tmem_empty[i], EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 2) {   // TMEM allocation: two accumulator buffers of BN fp32 columns each, by one warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                        // ===== TMA producer =====
            int it = 0;                                         // k-block counter across tiles: the ring never drains between tiles
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % STAGES, round = it / STAGES;
                    mb_wait(&empty[s], (round & 1) ^ 1);        // slot free (first round passes immediately)
                    mb_expect_tx(&full[s], A_BYTES + B_BYTES);
                    tma_load_2d(sa + s * A_BYTES, &map_a, &full[s], kb * BK, m0);
                    tma_load_2d(sb + s * B_BYTES, &map_b, &full[s], kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                        // ===== MMA issuer =====
            const uint32_t idesc = umma_idesc(BN);
            int it = 0, j = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++j) {
                const int buf = j & 1;
                mb_wait(&tmem_empty[buf], ((j >> 1) & 1) ^ 1);  // the epilogue has drained this buffer (first two tiles pass)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % STAGES, round = it / STAGES;
                    mb_wait(&full[s], round & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = umma_desc(sa + s * A_BYTES), db = umma_desc(sb + s * B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)           // 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the address field
                        umma_f16(acc, da + 2 * k, db + 2 * k, idesc, (kb | k) ? 1u : 0u);
                    umma_commit(&empty[s]);                     // frees the stage once these MMAs have read it
                }
                umma_commit(&tmem_full[buf]);                   // accumulator complete
            }
        }
    } else {                                                    // ===== epilogue warps 2..9 =====
        const int q = warp & 3;                                 // a warp may only touch TMEM lanes 32*(warp%4) .. +31
        const int ew = warp - 2, chalf = ew >> 2;               // two warps per lane quarter: each takes half of the columns
        // The accumulator arrives one ROW per lane (32 columns per tcgen05.ld).  Written out like that, every
        // store instruction would touch 32 different rows; instead each 32 x 32 block goes through a per-warp,
        // XOR-swizzled shared tile and leaves with 8 lanes per row: 128-bit accesses, 4 full rows per instruction,
        // bias / activation / positional add applied there.
        float4* tile_s = (float4*)(scratch + ew * EPI_SCRATCH); // [32 rows][8 float4], index r*8 + (c4 ^ (r & 7))
        short* s_pos = (short*)(scratch + ew * EPI_SCRATCH + 4096);   // [32 rows][3] clamped position ids
        const int c4 = lane & 7, rsub = lane >> 3;              // after the transpose: this lane's float4 column and row phase
        int j = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++j) {
            const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * BN, buf = j & 1;
            if (ep.pos_ids) {
                const int row = m0 + q * 32 + lane;
                long long p0 = 0, p1 = 0, p2 = 0;
                if (row < M) {
                    p0 = ep.pos_ids[3 * (size_t)row]; p1 = ep.pos_ids[3 * (size_t)row + 1]; p2 = ep.pos_ids[3 * (size_t)row + 2];
                    p0 = p0 < 0 ? 0 : (p0 >= ep.max_x ? ep.max_x - 1 : p0);
                    p1 = p1 < 0 ? 0 : (p1 >= ep.max_y ? ep.max_y - 1 : p1);
                    p2 = p2 < 0 ? 0 : (p2 >= ep.max_t ? ep.max_t - 1 : p2);
                }
                __syncwarp();                                   // the previous tile's rows have been consumed
                s_pos[3 * lane] = (short)p0; s_pos[3 * lane + 1] = (short)p1; s_pos[3 * lane + 2] = (short)p2;
                __syncwarp();                                   // every lane reads other lanes' rows below
            }
            mb_wait(&tmem_full[buf], (j >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem_base + (uint32_t)(buf * BN) + ((uint32_t)(q * 32) << 16);
            for (int c = chalf * (BN / 2); c < (chalf + 1) * (BN / 2); c += 32) {
                // the positional-table rows of this block's 8 output rows per lane: all 24 loads leave before the accumulator
                // read below, so one L2 round trip covers them and the TMEM read together
                const int col = n0 + c + 4 * c4;
                float4 e[8];
                if (!ep.leaky && ep.pos_ids) {
                    float4 e0[8], e1[8], e2[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = 4 * i + rsub;
                        e0[i] = __ldg((const float4*)(ep.x_emb + (size_t)s_pos[3 * r] * N + col));
                        e1[i] = __ldg((const float4*)(ep.y_emb + (size_t)s_pos[3 * r + 1] * N + col));
                        e2[i] = __ldg((const float4*)(ep.t_emb + (size_t)s_pos[3 * r + 2] * N + col));
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        e[i] = make_float4((e0[i].x + e1[i].x) + e2[i].x, (e0[i].y + e1[i].y) + e2[i].y, (e0[i].z + e1[i].z) + e2[i].z,
                                           (e0[i].w + e1[i].w) + e2[i].w);
                }
                const float4 b4 = __ldg((const float4*)(ep.bias + col));
                float v[32];
                tmem_ld32(acc + (uint32_t)c, v);
                if (c + 32 >= (chalf + 1) * (BN / 2)) {         // last read of this buffer by this warp: hand it back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(&tmem_empty[buf])) : "memory");
                }
                __syncwarp();                                   // the previous block has been read out of the tile
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) tile_s[lane * 8 + (jj ^ (lane & 7))] = make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = 4 * i + rsub, row = m0 + q * 32 + r;
                    const float4 a = tile_s[r * 8 + (c4 ^ (r & 7))];
                    if (row >= M) continue;
                    float x[4] = {a.x + b4.x, a.y + b4.y, a.z + b4.z, a.w + b4.w};
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) x[jj] = __bfloat162float(__float2bfloat16_rn(x[jj]));      // the Linear's bf16 output
                    if (ep.leaky) {
                        __nv_bfloat162 lo = __floats2bfloat162_rn(x[0] > 0.f ? x[0] : 0.01f * x[0], x[1] > 0.f ? x[1] : 0.01f * x[1]);
                        __nv_bfloat162 hi = __floats2bfloat162_rn(x[2] > 0.f ? x[2] : 0.01f * x[2], x[3] > 0.f ? x[3] : 0.01f * x[3]);
                        *(uint2*)((__nv_bfloat16*)ep.out + (size_t)row * N + col) = make_uint2(*(unsigned*)&lo, *(unsigned*)&hi);
                    } else {
                        if (ep.pos_ids) { x[0] += e[i].x; x[1] += e[i].y; x[2] += e[i].z; x[3] += e[i].w; }
                        fl_stg_stream4((float4*)((float*)ep.out + (size_t)row * N + col), make_float4(x[0], x[1], x[2], x[3]));
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}

__global__ void k_cast_bf16(const float4* __restrict__ in, uint2* __restrict__ out, long n4) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = fl_ldg_stream4(in + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    out[i] = make_uint2(*(unsigned*)&a, *(unsigned*)&b);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(CUtensorMap* map, const void* base, int rows, int cols, int box_rows) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        FL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        FL_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, FL_E_ARG, "cuTensorMapEncodeTiled is not available in this driver");
        encode = (EncodeTiledFn)fn;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};          // innermost first
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};                       // bytes between rows
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FL_REQUIRE(r == CUDA_SUCCESS, FL_E_ARG, "cuTensorMapEncodeTiled failed (%d) for a %d x %d bf16 matrix", (int)r, rows, cols);
    return FL_OK;
}

static size_t gemm_smem_bytes(int bn) { return (size_t)STAGES * (BM * BK * 2 + bn * BK * 2) + EPI_WARPS * EPI_SCRATCH + (2 * STAGES + 4) * 8 + 16; }

template <int BN>
int launch_gemm(const void* A, const void* B, int M, int N, int K, const Epilogue& ep, cudaStream_t st) {
    CUtensorMap ma, mb;
    int rc = make_map(&ma, A, M, K, BM);
    if (rc) return rc;
    rc = make_map(&mb, B, N, K, BN);
    if (rc) return rc;
    const size_t smem = gemm_smem_bytes(BN);
    static FlOncePerDevice attr;
    if (attr.first_use()) FL_CUDA(cudaFuncSetAttribute(k_gemm_tcgen05<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long tiles = (long)(N / BN) * ((M + BM - 1) / BM);
    const int grid = tiles < FL_SM_COUNT ? (int)tiles : FL_SM_COUNT;          // persistent: one CTA per SM
    k_gemm_tcgen05<BN><<<grid, GEMM_THREADS, smem, st>>>(ma, mb, M, N, K, ep);
    FL_LAUNCH_CHECK();
    return FL_OK;
}


// out[r] = pre[src(r)] + ((x_emb[p0] + y_emb[p1]) + t_emb[p2]): the positional add of the second GEMM's epilogue on its own, over a
// ring of cached pre-positional embeddings.  Row r = (b * c + j) * L + l of the output takes state (start + j) % ctx of the ring
// [ctx][B][L][N].  One thread = 4 columns.
__global__ void k_pos_add_ring(const float4* __restrict__ pre, const float* __restrict__ x_emb, const float* __restrict__ y_emb,
                               const float* __restrict__ t_emb, const long long* __restrict__ ids, int max_x, int max_y, int max_t,
                               float4* __restrict__ out, long rows, int n4, int B, int c, int L, int ctx, int start) {
    const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= rows * n4) return;
    const long r = g / n4;
    const int q = (int)(g - r * n4);
    const int l = (int)(r % L);
    const long bj = r / L;
    const int j = (int)(bj % c), b = (int)(bj / c);
    const long src = ((long)((start + j) % ctx) * B + b) * L + l;
    long long p0 = ids[3 * r], p1 = ids[3 * r + 1], p2 = ids[3 * r + 2];
    p0 = p0 < 0 ? 0 : (p0 >= max_x ? max_x - 1 : p0);
    p1 = p1 < 0 ? 0 : (p1 >= max_y ? max_y - 1 : p1);
    p2 = p2 < 0 ? 0 : (p2 >= max_t ? max_t - 1 : p2);
    const size_t N = (size_t)n4 * 4;
    const float4 a = __ldg(pre + src * n4 + q);
    const float4 e0 = __ldg((const float4*)(x_emb + (size_t)p0 * N) + q), e1 = __ldg((const float4*)(y_emb + (size_t)p1 * N) + q),
                 e2 = __ldg((const float4*)(t_emb + (size_t)p2 * N) + q);
    fl_stg_stream4(out + g, make_float4(a.x + ((e0.x + e1.x) + e2.x), a.y + ((e0.y + e1.y) + e2.y), a.z + ((e0.z + e1.z) + e2.z),
                                        a.w + ((e0.w + e1.w) + e2.w)));
}

}  // namespace

extern "C" int fl_pos_add_ring(const float* d_pre, const float* d_x_emb, const float* d_y_emb, const float* d_t_emb,
                               const long long* d_pos_ids, int max_x, int max_y, int max_t, float* d_out, int B, int c, int L, int ctx,
                               int start, int out_dim, void* stream) {
    FL_REQUIRE(d_pre && d_x_emb && d_y_emb && d_t_emb && d_pos_ids && d_out, FL_E_ARG, "fl_pos_add_ring: null pointer");
    FL_REQUIRE(B > 0 && L > 0 && ctx > 0 && c > 0 && c <= ctx && start >= 0 && start < ctx && out_dim > 0 && out_dim % 4 == 0 &&
                   max_x > 0 && max_y > 0 && max_t > 0,
               FL_E_ARG, "fl_pos_add_ring: bad sizes");
    FL_REQUIRE(((uintptr_t)d_pre | (uintptr_t)d_out | (uintptr_t)d_x_emb | (uintptr_t)d_y_emb | (uintptr_t)d_t_emb) % 16 == 0, FL_E_ALIGN,
               "fl_pos_add_ring: buffers must be 16-byte aligned");
    const long rows = (long)B * c * L, n = rows * (out_dim / 4);
    k_pos_add_ring<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)d_pre, d_x_emb, d_y_emb, d_t_emb, d_pos_ids,
                                                                               max_x, max_y, max_t, (float4*)d_out, rows, out_dim / 4, B,
                                                                               c, L, ctx, start);
    FL_LAUNCH_CHECK();
    return FL_OK;
}

extern "C" int fl_cast_bf16(const float* d_in, void* d_out_bf16, long n, void* stream) {
    FL_REQUIRE(d_in && d_out_bf16 && n > 0 && n % 4 == 0, FL_E_ARG, "fl_cast_bf16: need non-null buffers and n %% 4 == 0");
    FL_REQUIRE(((uintptr_t)d_in % 16 == 0) && ((uintptr_t)d_out_bf16 % 8 == 0), FL_E_ALIGN, "fl_cast_bf16: unaligned buffer");
    long n4 = n / 4;
    k_cast_bf16<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float4*)d_in, (uint2*)d_out_bf16, n4);
    FL_LAUNCH_CHECK();
    return FL_OK;
}

// rows of `cols` floats -> rows of `out_cols` >= cols bf16 values, the tail of every row zero: tokens whose length is not a multiple
// of the GEMM's K step (patches other than 16 x 16) meet weights padded with zero columns the same way
__global__ void k_cast_bf16_rows(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long rows, int cols, int out_cols) {
    const long u = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= rows * out_cols) return;
    const long r = u / out_cols;
    const int c = (int)(u - r * out_cols);
    out[u] = c < cols ? __float2bfloat16_rn(__ldg(in + r * cols + c)) : __float2bfloat16_rn(0.f);
}

extern "C" int fl_cast_bf16_rows(const float* d_in, void* d_out_bf16, long rows, int cols, int out_cols, void* stream) {
    FL_REQUIRE(d_in && d_out_bf16 && rows > 0 && cols > 0 && out_cols >= cols, FL_E_ARG, "fl_cast_bf16_rows: need non-null buffers, rows > 0, 0 < cols <= out_cols");
    FL_REQUIRE(((uintptr_t)d_in % 4 == 0) && ((uintptr_t)d_out_bf16 % 2 == 0), FL_E_ALIGN, "fl_cast_bf16_rows: unaligned buffer");
    const long n = rows * out_cols;
    FL_REQUIRE((n + 255) / 256 < 0x7fffffffL, FL_E_ARG, "fl_cast_bf16_rows: too many values");
    k_cast_bf16_rows<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_in, (__nv_bfloat16*)d_out_bf16, rows, cols, out_cols);
    FL_LAUNCH_CHECK();
    return FL_OK;
}

extern "C" int fl_patch_embed(const void* d_x_bf16, const void* d_w1_bf16, const float* d_b1, const void* d_w2_bf16,
                              const float* d_b2, const float* d_x_emb, const float* d_y_emb, const float* d_t_emb,
                              const long long* d_pos_ids, int max_x, int max_y, int max_t, void* d_hidden_bf16, float* d_out,
                              int n_tokens, int in_dim, int hid_dim, int out_dim, void* stream) {
    FL_REQUIRE(d_x_bf16 && d_w1_bf16 && d_b1 && d_w2_bf16 && d_b2 && d_hidden_bf16 && d_out, FL_E_ARG, "fl_patch_embed: null pointer");
    FL_REQUIRE(n_tokens <= 65535 * BM, FL_E_ARG, "fl_patch_embed: at most %d tokens per call", 65535 * BM);
    FL_REQUIRE(n_tokens > 0 && in_dim % BK == 0 && hid_dim % BK == 0 && hid_dim % 256 == 0 && out_dim % 256 == 0, FL_E_ARG,
               "fl_patch_embed: dims (%d -> %d -> %d) must be multiples of 64 (K) and 256 (N)", in_dim, hid_dim, out_dim);
    FL_REQUIRE((d_pos_ids == nullptr) || (d_x_emb && d_y_emb && d_t_emb && max_x > 0 && max_y > 0 && max_t > 0), FL_E_ARG,
               "fl_patch_embed: position ids given without embedding tables");
    FL_REQUIRE(d_pos_ids == nullptr || (max_x <= 32768 && max_y <= 32768 && max_t <= 32768), FL_E_ARG,
               "fl_patch_embed: positional tables of more than 32768 rows are not supported (ids are staged as 16-bit)");
    FL_REQUIRE(((uintptr_t)d_x_bf16 | (uintptr_t)d_w1_bf16 | (uintptr_t)d_w2_bf16 | (uintptr_t)d_hidden_bf16 | (uintptr_t)d_out) % 16 == 0,
               FL_E_ALIGN, "fl_patch_embed: buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    Epilogue e1{d_b1, 1, d_hidden_bf16, nullptr, nullptr, nullptr, nullptr, 0, 0, 0};
    int rc = launch_gemm<256>(d_x_bf16, d_w1_bf16, n_tokens, hid_dim, in_dim, e1, st);
    if (rc) return rc;
    Epilogue e2{d_b2, 0, d_out, d_x_emb, d_y_emb, d_t_emb, d_pos_ids, max_x, max_y, max_t};
    return launch_gemm<256>(d_hidden_bf16, d_w2_bf16, n_tokens, out_dim, hid_dim, e2, st);
}
