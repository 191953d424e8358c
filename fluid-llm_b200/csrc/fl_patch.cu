// fl_patch.cu -- inverse path: unpatchify / patchify permutations, the fused rollout step, and the
// grid -> node nearest-cell resample.
//
//   fl_patch_to_img   src/utils_model.py:77-92   (reshape + transpose + F.fold, non-overlapping)
//   fl_img_to_patch   src/utils_model.py:95-109  (F.unfold + view + permute)
//   fl_rollout_step   src/models/model.py:164,206,210 (img_to_patch; diffs[mask] = 0; last + diffs)
//   fl_grid2mesh      eagle/Dataloader/IMG_Eagle.py:93-123
//
// The two permutations move 16-byte units: a patch row (py elements) is contiguous on both
// sides, so every thread copies one 128-bit piece of a row; the destination side is fully
// coalesced and the source side is read in whole 32-byte sectors.
#include "fl_common.cuh"
#include <cuda_bf16.h>

namespace {

// unit = 16 bytes; upr = units per patch row (py * elem_size / 16).  IDX = unsigned when every unit index fits 32 bits
// (a 64-bit division costs an order of magnitude more instructions than the copy it addresses).
template <bool TO_IMG, typename IDX>
__global__ void __launch_bounds__(256) k_permute16(const uint4* __restrict__ src, uint4* __restrict__ dst, long total_units,
                                                   int n_bx, int n_by, int C, int px, int upr) {
    const long u0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u0 >= total_units) return;
    const IDX u = (IDX)u0;
    const IDX row_units = (IDX)n_by * (IDX)upr;
    // decompose the IMAGE-side unit index: [b][c][X][Yu] with Yu over n_by*upr units
    IDX img_u, pat_u;
    if (TO_IMG) {
        img_u = u;
        IDX r = u / row_units;
        const IDX yu = u - r * row_units;
        const IDX X = r % (IDX)(n_bx * px);
        r /= (IDX)(n_bx * px);
        const IDX c = r % (IDX)C, b = r / (IDX)C;
        const IDX by = yu / (IDX)upr, ju = yu - by * (IDX)upr, bx = X / (IDX)px, i = X - bx * (IDX)px;
        pat_u = ((((b * n_bx + bx) * n_by + by) * C + c) * px + i) * upr + ju;
    } else {
        pat_u = u;
        IDX r = u / (IDX)upr;
        const IDX ju = u - r * (IDX)upr;
        const IDX i = r % (IDX)px; r /= (IDX)px;
        const IDX c = r % (IDX)C; r /= (IDX)C;
        const IDX by = r % (IDX)n_by; r /= (IDX)n_by;
        const IDX bx = r % (IDX)n_bx, b = r / (IDX)n_bx;
        img_u = (((b * C + c) * n_bx + bx) * px + i) * row_units + by * upr + ju;
    }
    // read-once / write-once: keep both streams out of L1
    const float4 v = fl_ldg_stream4((const float4*)src + (TO_IMG ? pat_u : img_u));
    fl_stg_stream4((float4*)dst + (TO_IMG ? img_u : pat_u), v);
}

// scalar fallback (row bytes not a multiple of 16, or unaligned pointers)
template <typename T, bool TO_IMG>
__global__ void k_permute_elem(const T* __restrict__ src, T* __restrict__ dst, long total, int n_bx, int n_by, int C,
                               int px, int py) {
    long u = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= total) return;
    int j = (int)(u % py);
    long r = u / py;
    int i = (int)(r % px); r /= px;
    int c = (int)(r % C); r /= C;
    int by = (int)(r % n_by); r /= n_by;
    int bx = (int)(r % n_bx);
    long b = r / n_bx;
    long img = (((b * C + c) * n_bx + bx) * px + i) * ((long)n_by * py) + (long)by * py + j;
    if (TO_IMG) dst[img] = src[u]; else dst[u] = src[img];
}

// patch-side unit of 4 floats: diffs = img gather, zero where mask, next = last + diffs
__global__ void k_rollout_step(const float4* __restrict__ img, const uchar4* __restrict__ mask,
                               const float4* __restrict__ last, float4* __restrict__ diffs, float4* __restrict__ next,
                               uint2* __restrict__ next_bf16, long total_units, int n_bx, int n_by, int C, int px, int upr) {
    long u = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= total_units) return;
    int ju = (int)(u % upr);
    long r = u / upr;
    int i = (int)(r % px); r /= px;
    int c = (int)(r % C); r /= C;
    int by = (int)(r % n_by); r /= n_by;
    int bx = (int)(r % n_bx);
    long b = r / n_bx;
    long img_u = (((b * C + c) * n_bx + bx) * px + i) * ((long)n_by * upr) + (long)by * upr + ju;
    float4 d = fl_ldg_stream4(img + img_u);
    uchar4 m = mask[u];
    if (m.x) d.x = 0.f;
    if (m.y) d.y = 0.f;
    if (m.z) d.z = 0.f;
    if (m.w) d.w = 0.f;
    float4 l = fl_ldg_stream4(last + u);
    fl_stg_stream4(diffs + u, d);
    const float4 n = make_float4(l.x + d.x, l.y + d.y, l.z + d.z, l.w + d.w);
    fl_stg_stream4(next + u, n);
    if (next_bf16) {     // the same values as tokens for the patch embedding (what `.to(bfloat16)` under autocast would produce)
        const __nv_bfloat162 lo = __floats2bfloat162_rn(n.x, n.y), hi = __floats2bfloat162_rn(n.z, n.w);
        next_bf16[u] = make_uint2(*(const unsigned*)&lo, *(const unsigned*)&hi);
    }
}

// The same, one value per unit: patch widths that are not a multiple of 4, or unaligned buffers.
__global__ void k_rollout_step1(const float* __restrict__ img, const uint8_t* __restrict__ mask, const float* __restrict__ last,
                                float* __restrict__ diffs, float* __restrict__ next, __nv_bfloat16* __restrict__ next_bf16,
                                long total, int n_bx, int n_by, int C, int px, int py) {
    long u = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= total) return;
    int j = (int)(u % py);
    long r = u / py;
    int i = (int)(r % px); r /= px;
    int c = (int)(r % C); r /= C;
    int by = (int)(r % n_by); r /= n_by;
    int bx = (int)(r % n_bx);
    long b = r / n_bx;
    long img_u = (((b * C + c) * n_bx + bx) * px + i) * ((long)n_by * py) + (long)by * py + j;
    const float d = mask[u] ? 0.f : __ldg(img + img_u);
    const float n = __fadd_rn(__ldg(last + u), d);
    diffs[u] = d;
    next[u] = n;
    if (next_bf16) next_bf16[u] = __float2bfloat16_rn(n);
}

// simple_dataloader.py:93,100: diffs = states[1:] - states[:-1]; masks = mask[1:] repeated over the 3 channels, as bool bytes.
// One unit = 4 pixels of one (frame, patch): 3 x float4 of two frames in, 3 x float4 + 3 x uchar4 out.
__global__ void k_sample_assemble(const float4* __restrict__ states, const uchar4* __restrict__ mask, float4* __restrict__ diffs,
                                  uchar4* __restrict__ mask3, long units, long frame_units /* L * ppx / 4 */, int upp /* ppx / 4 */,
                                  long frames_out_per_sample, long frames_in_per_sample) {
    const long u = (long)blockIdx.x * blockDim.x + threadIdx.x;       // over [B][T-1][L][ppx/4]
    if (u >= units) return;
    const long fo = u / frame_units, r = u - fo * frame_units;          // output frame (over all samples), unit inside the frame
    const long b = fo / frames_out_per_sample, t = fo - b * frames_out_per_sample;
    const long fi = b * frames_in_per_sample + t;                       // states[b, t]; the next frame is fi + 1
    const long l = r / upp, k = r - l * upp;
    const uchar4 m = mask[(fi + 1) * frame_units + r];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const long o = ((l * 3 + c) * upp + k);
        const float4 a = __ldg(states + fi * 3 * frame_units + o), n = __ldg(states + (fi + 1) * 3 * frame_units + o);
        diffs[fo * 3 * frame_units + o] = make_float4(n.x - a.x, n.y - a.y, n.z - a.z, n.w - a.w);
        mask3[fo * 3 * frame_units + o] = m;
    }
}

// The same, one pixel per unit: patches whose pixel count is not a multiple of 4 (5 x 5, 10 x 7, ...) or unaligned buffers.
__global__ void k_sample_assemble1(const float* __restrict__ states, const uint8_t* __restrict__ mask, float* __restrict__ diffs,
                                   uint8_t* __restrict__ mask3, long units, long frame_px /* L * ppx */, int ppx,
                                   long frames_out_per_sample, long frames_in_per_sample) {
    const long u = (long)blockIdx.x * blockDim.x + threadIdx.x;       // over [B][T-1][L][ppx]
    if (u >= units) return;
    const long fo = u / frame_px, r = u - fo * frame_px;
    const long b = fo / frames_out_per_sample, t = fo - b * frames_out_per_sample;
    const long fi = b * frames_in_per_sample + t;
    const long l = r / ppx, k = r - l * ppx;
    const uint8_t m = mask[(fi + 1) * frame_px + r];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const long o = ((l * 3 + c) * ppx + k);
        diffs[fo * 3 * frame_px + o] = __fsub_rn(__ldg(states + (fi + 1) * 3 * frame_px + o), __ldg(states + fi * 3 * frame_px + o));
        mask3[fo * 3 * frame_px + o] = m;
    }
}

// NumPy's npy_floor_dividef (numpy/core/src/npymath/npy_math_internal.h.src), float32 throughout
__device__ __forceinline__ float np_floor_divide_f(float a, float b) {
    if (b == 0.f) return __fdiv_rn(a, b);
    float mod = fmodf(a, b);
    float div = __fdiv_rn(__fsub_rn(a, mod), b);
    if (mod != 0.f && ((b < 0.f) != (mod < 0.f))) div = __fsub_rn(div, 1.0f);
    if (div != 0.f) {
        float fl = floorf(div);
        if (__fsub_rn(div, fl) > 0.5f) fl = __fadd_rn(fl, 1.0f);
        return fl;
    }
    return copysignf(0.f, __fdiv_rn(a, b));
}

__global__ void k_grid2mesh(const float* __restrict__ grid, const float* __restrict__ mesh_pos, float* __restrict__ out,
                            int T, int N, int H, int W, int C, float x_min, float y_min, float half_sx, float half_sy,
                            float sx, float neg_sy, int* __restrict__ oob) {
    long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long)T * N) return;
    long t = g / N;
    const float2 p = __ldg((const float2*)mesh_pos + g);
    // IMG_Eagle.py:111-112 in float32 (NumPy 1.26 keeps float32 arrays float32 against scalars)
    float fx = np_floor_divide_f(__fadd_rn(__fsub_rn(p.x, x_min), half_sx), sx);
    float fy = np_floor_divide_f(__fadd_rn(__fsub_rn(p.y, y_min), half_sy), neg_sy);
    long ix = (long)fx, iy = (long)fy;
    if (ix < 0) ix += W;                       // NumPy negative indices wrap once
    if (iy < 0) iy += H;
    if (oob && (ix < 0 || ix >= W || iy < 0 || iy >= H)) atomicAdd(oob, 1);   // the reference raises IndexError here: counted for the host
    ix = ix < 0 ? 0 : (ix >= W ? W - 1 : ix);  // (and clamped, so that the kernel itself stays inside the grid)
    iy = iy < 0 ? 0 : (iy >= H ? H - 1 : iy);
    long row = H - 1 - iy;                     // np.flip(grid, axis=1), IMG_Eagle.py:105-106
    const float* s = grid + ((t * H + row) * W + ix) * C;
    float* d = out + g * C;
    for (int c = 0; c < C; ++c) d[c] = __ldg(s + c);
}

template <bool TO_IMG>
int permute(const void* src, void* dst, int B, int n_bx, int n_by, int C, int px, int py, int es, cudaStream_t st) {
    long total = (long)B * n_bx * n_by * C * px * py;
    bool vec = ((py * es) % 16 == 0) && (((uintptr_t)src | (uintptr_t)dst) % 16 == 0);
    if (vec) {
        int upr = py * es / 16;
        long units = total * es / 16;
        if (units < 0x7fffffffL)
            k_permute16<TO_IMG, unsigned><<<(unsigned)((units + 255) / 256), 256, 0, st>>>((const uint4*)src, (uint4*)dst, units,
                                                                                           n_bx, n_by, C, px, upr);
        else
            k_permute16<TO_IMG, unsigned long long><<<(unsigned)((units + 255) / 256), 256, 0, st>>>((const uint4*)src, (uint4*)dst,
                                                                                                     units, n_bx, n_by, C, px, upr);
    } else if (es == 4) {
        k_permute_elem<uint32_t, TO_IMG><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const uint32_t*)src, (uint32_t*)dst,
                                                                                       total, n_bx, n_by, C, px, py);
    } else {
        k_permute_elem<uint16_t, TO_IMG><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const uint16_t*)src, (uint16_t*)dst,
                                                                                       total, n_bx, n_by, C, px, py);
    }
    FL_LAUNCH_CHECK();
    return FL_OK;
}

// channel-last affine map of the pre-gridded EAGLE states (eagle/Dataloader/IMG_Eagle.py:72-90): mode 0 = (x - mean) / std,
// mode 1 = x * std + mean, each as two separately rounded fp32 operations like the reference's torch expressions
struct AffineC { float mean[8], stdv[8]; };
__global__ void k_affine_channels(const float* __restrict__ in, float* __restrict__ out, long n, int C, AffineC k, int mode) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % C);
    const float x = __ldg(in + i);
    out[i] = mode == 0 ? __fdiv_rn(__fsub_rn(x, k.mean[c]), k.stdv[c]) : __fadd_rn(__fmul_rn(x, k.stdv[c]), k.mean[c]);
}
__global__ void k_affine_channels4(const float4* __restrict__ in, float4* __restrict__ out, long n4, AffineC k, int mode) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 x = fl_ldg_stream4(in + i);
    float4 y;
    if (mode == 0) {
        y = make_float4(__fdiv_rn(__fsub_rn(x.x, k.mean[0]), k.stdv[0]), __fdiv_rn(__fsub_rn(x.y, k.mean[1]), k.stdv[1]),
                        __fdiv_rn(__fsub_rn(x.z, k.mean[2]), k.stdv[2]), __fdiv_rn(__fsub_rn(x.w, k.mean[3]), k.stdv[3]));
    } else {
        y = make_float4(__fadd_rn(__fmul_rn(x.x, k.stdv[0]), k.mean[0]), __fadd_rn(__fmul_rn(x.y, k.stdv[1]), k.mean[1]),
                        __fadd_rn(__fmul_rn(x.z, k.stdv[2]), k.mean[2]), __fadd_rn(__fmul_rn(x.w, k.stdv[3]), k.mean[3]));
    }
    fl_stg_stream4(out + i, y);
}

int check_perm_args(const char* who, const void* a, const void* b, int B, int n_bx, int n_by, int C, int px, int py, int es) {
    FL_REQUIRE(a && b, FL_E_ARG, "%s: null pointer", who);
    FL_REQUIRE(B > 0 && n_bx > 0 && n_by > 0 && C > 0 && px > 0 && py > 0, FL_E_ARG, "%s: sizes must be positive", who);
    FL_REQUIRE(es == 2 || es == 4, FL_E_ARG, "%s: elem_size must be 2 or 4, got %d", who, es);
    FL_REQUIRE((long)B * n_bx * n_by * C * px * py < (1L << 40), FL_E_ARG, "%s: tensor too large", who);
    return FL_OK;
}

}  // namespace

extern "C" int fl_patch_to_img(const void* d_patches, void* d_img, int B, int n_bx, int n_by, int C, int px, int py,
                               int elem_size, void* stream) {
    int rc = check_perm_args("fl_patch_to_img", d_patches, d_img, B, n_bx, n_by, C, px, py, elem_size);
    if (rc) return rc;
    return permute<true>(d_patches, d_img, B, n_bx, n_by, C, px, py, elem_size, (cudaStream_t)stream);
}

extern "C" int fl_img_to_patch(const void* d_img, void* d_patches, int B, int n_bx, int n_by, int C, int px, int py,
                               int elem_size, void* stream) {
    int rc = check_perm_args("fl_img_to_patch", d_img, d_patches, B, n_bx, n_by, C, px, py, elem_size);
    if (rc) return rc;
    return permute<false>(d_img, d_patches, B, n_bx, n_by, C, px, py, elem_size, (cudaStream_t)stream);
}

extern "C" int fl_rollout_step(const float* d_pred_img, const uint8_t* d_mask, const float* d_last, float* d_diffs,
                               float* d_next, void* d_next_bf16, int B, int n_bx, int n_by, int C, int px, int py, void* stream) {
    FL_REQUIRE(d_pred_img && d_mask && d_last && d_diffs && d_next, FL_E_ARG, "fl_rollout_step: null pointer");
    FL_REQUIRE(B > 0 && n_bx > 0 && n_by > 0 && C > 0 && px > 0 && py > 0, FL_E_ARG, "fl_rollout_step: sizes must be positive");
    FL_REQUIRE(((uintptr_t)d_pred_img | (uintptr_t)d_last | (uintptr_t)d_diffs | (uintptr_t)d_next) % 4 == 0 && (uintptr_t)d_next_bf16 % 2 == 0,
               FL_E_ALIGN, "fl_rollout_step: float buffers must be 4-byte aligned, the bf16 token buffer 2-byte aligned");
    const bool vec = py % 4 == 0 && (((uintptr_t)d_pred_img | (uintptr_t)d_last | (uintptr_t)d_diffs | (uintptr_t)d_next) % 16 == 0) &&
                     ((uintptr_t)d_mask % 4 == 0) && ((uintptr_t)d_next_bf16 % 8 == 0);
    if (!vec) {       // any patch width / alignment: one value per thread
        const long total = (long)B * n_bx * n_by * C * px * py;
        FL_REQUIRE((total + 255) / 256 < 0x7fffffffL, FL_E_ARG, "fl_rollout_step: tensor too large");
        k_rollout_step1<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            d_pred_img, d_mask, d_last, d_diffs, d_next, (__nv_bfloat16*)d_next_bf16, total, n_bx, n_by, C, px, py);
        FL_LAUNCH_CHECK();
        return FL_OK;
    }
    long units = (long)B * n_bx * n_by * C * px * py / 4;
    k_rollout_step<<<(unsigned)((units + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)d_pred_img, (const uchar4*)d_mask, (const float4*)d_last, (float4*)d_diffs, (float4*)d_next,
        (uint2*)d_next_bf16, units, n_bx, n_by, C, px, py / 4);
    FL_LAUNCH_CHECK();
    return FL_OK;
}

extern "C" int fl_sample_assemble(const float* d_states, const uint8_t* d_mask, int B, int T, int L, int px, int py, float* d_diffs,
                                  uint8_t* d_mask3, void* stream) {
    FL_REQUIRE(d_states && d_mask && d_diffs && d_mask3, FL_E_ARG, "fl_sample_assemble: null pointer");
    FL_REQUIRE(B > 0 && T > 1 && L > 0 && px > 0 && py > 0, FL_E_ARG, "fl_sample_assemble: need B > 0, T > 1, L > 0, px > 0, py > 0");
    FL_REQUIRE(((uintptr_t)d_states | (uintptr_t)d_diffs) % 4 == 0, FL_E_ALIGN, "fl_sample_assemble: float buffers must be 4-byte aligned");
    const bool vec = (px * py) % 4 == 0 && (((uintptr_t)d_states | (uintptr_t)d_diffs) % 16 == 0) &&
                     (((uintptr_t)d_mask | (uintptr_t)d_mask3) % 4 == 0);
    if (!vec) {
        const int ppx = px * py;
        const long frame_px = (long)L * ppx, units = (long)B * (T - 1) * frame_px;
        FL_REQUIRE((units + 255) / 256 < 0x7fffffffL, FL_E_ARG, "fl_sample_assemble: too many pixels");
        k_sample_assemble1<<<(unsigned)((units + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_states, d_mask, d_diffs, d_mask3, units,
                                                                                               frame_px, ppx, T - 1, T);
        FL_LAUNCH_CHECK();
        return FL_OK;
    }
    const int upp = px * py / 4;
    const long frame_units = (long)L * upp, units = (long)B * (T - 1) * frame_units;
    k_sample_assemble<<<(unsigned)((units + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const float4*)d_states, (const uchar4*)d_mask, (float4*)d_diffs, (uchar4*)d_mask3, units, frame_units, upp, T - 1, T);
    FL_LAUNCH_CHECK();
    return FL_OK;
}

extern "C" int fl_affine_channels(const float* d_in, float* d_out, long n_values, int C, const float* h_mean, const float* h_std,
                                  int denormalize, void* stream) {
    FL_REQUIRE(d_in && d_out && h_mean && h_std, FL_E_ARG, "fl_affine_channels: null pointer");
    FL_REQUIRE(n_values > 0 && C >= 1 && C <= 8 && n_values % C == 0, FL_E_ARG, "fl_affine_channels: need 1 <= C <= 8 and n_values a multiple of C");
    AffineC k;
    for (int c = 0; c < 8; ++c) { k.mean[c] = c < C ? h_mean[c] : 0.f; k.stdv[c] = c < C ? h_std[c] : 1.f; }
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 4 && ((uintptr_t)d_in | (uintptr_t)d_out) % 16 == 0) {
        const long n4 = n_values / 4;
        k_affine_channels4<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>((const float4*)d_in, (float4*)d_out, n4, k, denormalize ? 1 : 0);
    } else {
        k_affine_channels<<<(unsigned)((n_values + 255) / 256), 256, 0, st>>>(d_in, d_out, n_values, C, k, denormalize ? 1 : 0);
    }
    FL_LAUNCH_CHECK();
    return FL_OK;
}

extern "C" int fl_grid2mesh(const float* d_grid, const float* d_mesh_pos, float* d_out, int T, int N, int H, int W, int C,
                            float x_min, float y_min, double step_x, double step_y, int32_t* d_out_of_range, void* stream) {
    FL_REQUIRE(d_grid && d_mesh_pos && d_out, FL_E_ARG, "fl_grid2mesh: null pointer");
    FL_REQUIRE(T > 0 && N > 0 && H > 0 && W > 0 && C > 0, FL_E_ARG, "fl_grid2mesh: sizes must be positive");
    FL_REQUIRE(step_x != 0.0 && step_y != 0.0, FL_E_ARG, "fl_grid2mesh: zero grid step");
    FL_REQUIRE((uintptr_t)d_mesh_pos % 8 == 0, FL_E_ALIGN, "fl_grid2mesh: mesh_pos must be 8-byte aligned");
    long n = (long)T * N;
    k_grid2mesh<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_grid, d_mesh_pos, d_out, T, N, H, W, C, x_min, y_min, (float)(step_x / 2), (float)(step_y / 2), (float)step_x,
        (float)(-step_y), d_out_of_range);
    FL_LAUNCH_CHECK();
    return FL_OK;
}
