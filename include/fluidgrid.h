/* fluidgrid.h -- C ABI of libfluidgrid.so: FLUID-LLM's per-timestep field data path on B200.
 *
 * The reference (dewan1988/FLUID-LLM) has no FFI: the path sits behind plain Python call
 * signatures whose arithmetic lives in matplotlib's C++ `_tri` module.  Each entry point below
 * names the reference interface it replaces (paths relative to the reference root); the Python
 * host package (fluid-llm_b200/) keeps the reference's own names on top of these calls and
 * INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  Pointers named d_* are DEVICE pointers owned by the
 *     caller (torch owns the memory; the library never frees or keeps them past the call);
 *     h_* are host pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is
 *     enqueued asynchronously on it; nothing synchronises unless stated.
 *   - return 0 = ok, < 0 = argument error (FL_E_*), > 0 = a cudaError_t.  fl_last_error()
 *     returns a thread-local message for the last non-zero return.
 *   - there is NO CPU fallback: without a CUDA device every compute call returns an error.
 */
#ifndef FLUIDGRID_H
#define FLUIDGRID_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FL_OK 0
#define FL_E_ARG (-1)      /* null pointer / non-positive size / bad flag */
#define FL_E_RANGE (-2)    /* index out of range (e.g. a cell vertex >= n_nodes) */
#define FL_E_WORKSPACE (-3)/* workspace too small */
#define FL_E_ALIGN (-4)    /* pointer not aligned as documented */

#define FL_ABI_VERSION 11

/* personality flags of fl_plan_patch_table / fl_interp_patchify */
#define FL_FLIP_Y 1u          /* airfoil_ds.py:80  -- y axis reversed after padding */
#define FL_MASK_AWARE_NORM 2u /* airfoil_ds.py:236-242 -- masked pixels stay 0, others normalised */
#define FL_NO_NORM 4u         /* normalize=False */
#define FL_FORCE_GATHER 8u    /* testing: always the gather-from-global kernel */
#define FL_FORCE_STAGED 16u   /* testing: ignore the tile plan (whole-mesh staged kernel, or the gather kernel) */
#define FL_FORCE_TILED 32u    /* testing: with a tile plan, always the node-list kernel (fl_tiled.cu) */
#define FL_NO_PAD 128u        /* fl_plan_patch_table: pad=False (simple_dataloader.py:118-119) */
#define FL_FORCE_RING 64u     /* testing: with a tile plan of <= 6 patches (px*py = 256) per tile and frames that fit shared memory a few
                                 times over, the experimental frame-ring kernel (fl_ring.cu); otherwise ignored */

/* One grid cell of the static per-mesh table (the product's own intermediate; the reference
 * recomputes the plane coefficients of every triangle per channel per frame instead,
 * src/_triinterpolate.py:262-263).  v[k] = node ids of the containing triangle in its
 * counter-clockwise-corrected order, tri = its index (-1: outside the mesh / in a hole / pad),
 * w1, w2 = barycentric weights of v[1], v[2] in fp64 (weight of v[0] is 1 - w1 - w2). */
typedef struct FlCellIdx { int32_t v0, v1, v2, tri; } FlCellIdx;   /* 16 B */
typedef struct FlCellW { double w1, w2; } FlCellW;                 /* 16 B */

int fl_abi_version(void);
const char* fl_last_error(void);
/* number of CUDA devices visible (0 if none / driver missing); never fails */
int fl_device_count(void);

/* ---- one-off per mesh: point location --------------------------------------------------------
 * Replaces  src/dataloader/mesh_utils.py:103-104  (matplotlib Triangulation + TrapezoidMapTriFinder
 * .find_many over the grid) and the per-triangle set-up of calculate_plane_coefficients.
 * d_pos f32[n_nodes,2], d_cells i32[n_cells,3] (any winding), d_grid_ax f32[nx], d_grid_ay f32[ny]
 * (the axes of grid_pos, mesh_utils.py:64-79; cell (ix,iy) is at (ax[ix], ay[iy])).
 * Outputs, all [nx*ny] in the reference's [ix, iy] order (iy fastest):
 *   d_tri_index i32 -- the stated tie-break rule, which is what matplotlib 3.8.2's TrapezoidMapTriFinder computes as
 *       restated in oracle/tri_oracle.cpp (matplotlib layer: restated, never executed -- it is not installed where this
 *       was built): a query on a mesh vertex -> lowest-index triangle listing it; on an edge -> the triangle above the
 *       edge (left of the edge directed from its lexicographically smaller to larger end point),
 *       else the one below; else the containing triangle; -1 if none.  Bit-exact against that restatement, against an
 *       independent brute-force evaluation of the rule and against an exact-rational one (oracle/exact_locator.py).
 *   d_cell_idx / d_cell_w -- the static table (may be NULL to skip).
 * Workspace: fl_locate_workspace_bytes(n_nodes, n_cells) bytes, 256-B aligned. */
size_t fl_locate_workspace_bytes(int n_nodes, int n_cells);
int fl_locate(const float* d_pos, const int32_t* d_cells, int n_nodes, int n_cells,
              const float* d_grid_ax, const float* d_grid_ay, int nx, int ny,
              int32_t* d_tri_index, FlCellIdx* d_cell_idx, FlCellW* d_cell_w,
              void* d_workspace, size_t workspace_bytes, void* stream);

/* The same without host synchronisation: d_status i32[2] is written on the stream -- [0] = triangles with a node id outside
 * 0 <= i < n_nodes (they are left out), [1] = bin entries that did not fit the workspace (0 = it fitted).  The results are valid
 * iff both are 0; otherwise call again with fl_locate (bad ids raise there) or a workspace larger by 4 * status[1] bytes. */
int fl_locate_async(const float* d_pos, const int32_t* d_cells, int n_nodes, int n_cells,
                    const float* d_grid_ax, const float* d_grid_ay, int nx, int ny,
                    int32_t* d_tri_index, FlCellIdx* d_cell_idx, FlCellW* d_cell_w,
                    void* d_workspace, size_t workspace_bytes, int32_t* d_status, void* stream);

/* Re-order the static table into output-pixel order for one dataset personality:
 * pad to multiples of the patch (simple_dataloader.py:137-152), optional y flip
 * (airfoil_ds.py:80), optional removal of `crop` outer rings of patches (airfoil_ds.py:132-133).
 * Output index = (l*px + i)*py + j with l = bx*n_by + by (F.unfold order,
 * simple_dataloader.py:131).  Padded pixels get tri = -1.
 * sx, sy: the unfold stride (simple_dataloader.py:131 `stride=self.stride`; 0 = the patch size): patch (bx, by) starts at
 * pixel (bx*sx, by*sy) of the padded (and cropped) frame, n_bx = floor((extent - px) / sx) + 1.  FL_NO_PAD (pad=False,
 * simple_dataloader.py:118): no padding; unfold then drops the remainder columns / rows.
 * d_out_idx/d_out_w: [n_bx*n_by*px*py].  n_bx/n_by are returned through out pointers (host).
 * d_node_slot i32[n_nodes] + d_out_idx_slot (optional, both or neither): also emit the table with node ids replaced by
 * d_node_slot[id] (FlTraj::d_idx_slot). */
int fl_plan_patch_table(const FlCellIdx* d_cell_idx, const FlCellW* d_cell_w, int nx, int ny,
                        int px, int py, int sx, int sy, int crop_patches, unsigned flags,
                        FlCellIdx* d_out_idx, FlCellW* d_out_w, int* h_n_bx, int* h_n_by,
                        const int32_t* d_node_slot, FlCellIdx* d_out_idx_slot, void* stream);

/* ---- per step: gather -> fp64 FMA -> fp32 -> mask -> normalise -> patchify --------------------
 * Replaces the body of  simple_dataloader.py:72-121,166-216 / airfoil_ds.py:71-122,189-244
 * (3x to_grid per frame, pad, concat mask, F.unfold, permute, _normalize).
 * One trajectory: d_velocity f32[T,n_nodes,2], d_pressure f32[T,n_nodes,1]; frames
 * t0, t0+interval, ... (n_frames of them) are processed.
 * d_states f32[n_frames, L, 3, px, py]; d_mask u8[n_frames, L, px, py] (NULL to skip; 1 = masked).
 * mean[3], stds[3] are host floats.  flags: FL_MASK_AWARE_NORM, FL_NO_NORM. */
typedef struct FlTraj {
    const float* d_velocity;   /* [T, n_nodes, 2] */
    const float* d_pressure;   /* [T, n_nodes, 1] */
    const FlCellIdx* d_idx;    /* patch-ordered table of this trajectory's mesh */
    const FlCellW* d_w;
    const FlCellIdx* d_idx_slot; /* optional (NULL = d_idx): same table with node ids replaced by shared-memory slots */
    const int32_t* d_node_slot;  /* optional (NULL = identity): slot of every node, [prs_stride]; a spatially sorted
                                    (Morton) order keeps the gathers of neighbouring pixels in neighbouring banks */
    float* d_states;           /* [n_frames, L, 3, px, py] */
    uint8_t* d_mask;           /* [n_frames, L, px, py] or NULL */
    int32_t n_nodes, t0, interval, n_frames;
    int32_t vel_stride;        /* floats between consecutive frames of d_velocity (>= 2*n_nodes) */
    int32_t prs_stride;        /* floats between consecutive frames of d_pressure (>= n_nodes)   */
    /* optional tile plan of the patch table (all NULL / 0 = none): the output patches are split into tiles, each with the
     * list of mesh nodes its pixels touch, so that only a tile's nodes are staged in shared memory -- meshes of any size
     * run the shared-memory kernel.  Every trajectory of a call must use the same split of the patches into tiles. */
    const FlCellIdx* d_idx_tile;   /* [L*px*py] d_idx with node ids replaced by 16 * (slot of the node in its tile's list) */
    const int32_t* d_tile_nodes;   /* the tiles' node lists, one after the other (node ids ascending inside a tile) */
    const int32_t* d_tile_desc;    /* [n_tiles][8] = {first entry in d_tile_nodes, nodes, first entry in d_tile_patches, patches,
                                      first entry in d_tile_quads / d_tile_qslots, quads, largest quad id + 1, 0} */
    const int32_t* d_tile_patches; /* the tiles' patch ids l, one after the other */
    const int32_t* d_tile_quads;   /* optional: per tile the quads (node id / 4, ascending) that hold its nodes, one list after the other */
    const int32_t* d_tile_qslots;  /* with d_tile_quads: [quads][4] the slots of each quad's 4 nodes (-1: not used by the tile); lets the
                                      kernel stage with coalesced 128-bit loads when the frames are 16-byte aligned and padded to quads */
    int32_t n_tiles, max_tile_nodes;
    int32_t max_tile_patches;      /* patches of the largest tile: at most 7 (px*py = 256) or 14 (128), one per patch group */
    int32_t idx_slot_format;       /* 0: d_idx_slot holds FlCellIdx records (16 B per output pixel); 1: the compact records
                                      fl_pack_idx16 makes of them (8 B: v0 | v1 << 16, v2 | outside << 16; slots below 65536).
                                      The staged kernel reads its table records once per work item: a third fewer bytes there
                                      are 2 - 5 % of the launch (profiles/README.md) */
} FlTraj;
/* prs_stride a multiple of 4, vel_stride >= 2 * prs_stride, 16-byte aligned bases and pad floats
 * that are readable and finite select the staged kernel: whole frames are staged in shared memory
 * as 16-byte node records and gathered there (patches of a multiple of 128 pixels).  Anything else runs
 * the gather-from-global kernel -- same results, slower.  px, py: ANY patch size F.unfold takes
 * (simple_dataloader.py:131), e.g. 5 x 5 or 32 x 32; the reference's configs use 16 x 16. */

/* FlCellIdx records with node SLOTS (fl_plan_patch_table's d_out_idx_slot) -> the compact 8-byte form of
 * FlTraj::idx_slot_format = 1.  d_out: uint32[n][2].  n_slots: slots are 0 <= s < n_slots <= 65536. */
int fl_pack_idx16(const FlCellIdx* d_idx_slot, long n, int n_slots, void* d_out, void* stream);

int fl_interp_patchify(const FlTraj* h_trajs, int n_traj, int n_patches, int px, int py,
                       const float* h_mean, const float* h_std, unsigned flags, void* stream);
/* same, with a copy of the descriptors already on the device (d_trajs; nothing is copied inside);
 * h_trajs is the identical host copy, read for validation and kernel selection */
int fl_interp_patchify_dev(const FlTraj* d_trajs, const FlTraj* h_trajs, int n_traj, int n_patches, int px, int py,
                           const float* h_mean, const float* h_std, unsigned flags, void* stream);

/* which kernel the last fl_interp_patchify[_dev] call of this thread launched: "k_interp_patchify_tiled" (tile plan given),
 * "k_interp_patchify_staged" (whole frames fit shared memory) or "k_interp_patchify_gather"; "" before the first call */
const char* fl_last_interp_kernel(void);

/* Plain-grid variant (no pad/patchify): replaces mesh_utils.to_grid (src/dataloader/mesh_utils.py:82-91)
 * for n_fields scalar node fields at once.  d_val f32[n_fields, n_nodes] -> d_data f32[n_fields, nx*ny],
 * d_mask u8[n_fields, nx*ny] (1 = outside mesh or non-finite). */
int fl_to_grid(const FlCellIdx* d_cell_idx, const FlCellW* d_cell_w, int nx, int ny,
               const float* d_val, int n_fields, int n_nodes, float* d_data, uint8_t* d_mask, void* stream);

/* Channel-last frame variant without pad/patchify: replaces eagle/Dataloader/IMG_MGN.py:78-157 (_get_step x T +
 * normalize + permute).  Frames t0, t0+interval, ...; `crop` pixels removed per side (IMG_MGN.py:91-95);
 * every pixel normalised (masked ones become (0 - mean) / std).  d_states f32 [n_frames, nx-2c, ny-2c, 3],
 * d_mask u8 [n_frames, nx-2c, ny-2c] (NULL to skip). */
int fl_interp_frames(const FlCellIdx* d_cell_idx, const FlCellW* d_cell_w, int nx, int ny, int crop,
                     const float* d_velocity, const float* d_pressure, int n_nodes, int vel_stride, int prs_stride,
                     int t0, int interval, int n_frames, const float* h_mean, const float* h_std, unsigned flags,
                     float* d_states, uint8_t* d_mask, void* stream);

/* ---- inverse path --------------------------------------------------------------------------- */
/* src/utils_model.py:77-92 patch_to_img: patches [B, L, C, px, py] -> img [B, C, n_bx*px, n_by*py].
 * elem_size 2 or 4 (bf16/fp16 or fp32; pure permutation). */
int fl_patch_to_img(const void* d_patches, void* d_img, int B, int n_bx, int n_by, int C, int px, int py,
                    int elem_size, void* stream);
/* src/utils_model.py:95-109 img_to_patch: the inverse permutation. */
int fl_img_to_patch(const void* d_img, void* d_patches, int B, int n_bx, int n_by, int C, int px, int py,
                    int elem_size, void* stream);
/* src/models/model.py:164,206,210 fused: diffs = img_to_patch(pred_img); diffs[mask] = 0;
 * next = last + diffs.  fp32.  d_pred_img [B, C, X, Y], d_mask u8 [B, L, C, px, py],
 * d_last [B, L, C, px, py] -> d_diffs, d_next same shape. */
int fl_rollout_step(const float* d_pred_img, const uint8_t* d_mask, const float* d_last,
                    float* d_diffs, float* d_next, void* d_next_bf16, int B, int n_bx, int n_by, int C, int px, int py,
                    void* stream);
/* d_next_bf16 (optional, NULL to skip): d_next rounded to bf16, same layout -- the patch tokens fl_patch_embed takes, so the
 * rollout needs no separate cast of the new state (src/models/model.py:196-204 re-embeds it under bf16 autocast). */
/* src/dataloader/simple_dataloader.py:93,100 (airfoil_ds.py:97-101): the sample's derived tensors in one launch.
 * d_states f32[B, T, L, 3, px, py], d_mask u8[B, T, L, px, py] (the outputs of fl_interp_patchify for B samples of T frames) ->
 * d_diffs f32[B, T-1, L, 3, px, py] = states[:, 1:] - states[:, :-1] and d_mask3 u8[B, T-1, L, 3, px, py] = mask[:, 1:]
 * repeated over the three channels (0 / 1 bytes: a torch.bool tensor's storage). */
int fl_sample_assemble(const float* d_states, const uint8_t* d_mask, int B, int T, int L, int px, int py,
                       float* d_diffs, uint8_t* d_mask3, void* stream);
/* eagle/Dataloader/IMG_Eagle.py:93-123 grid2mesh: nearest-cell grid -> node resample.
 * d_grid f32[T, H, W, C] (row 0 = Ymin, flipped inside as the reference does), d_mesh_pos f32[T, N, 2]
 * -> d_out f32[T, N, C].  Extents/steps are the reference's constants unless overridden;
 * index arithmetic in float32 exactly as NumPy 1.26 evaluates it.  Negative indices wrap once, as NumPy's do.
 * d_out_of_range (optional i32[1], ADDED to, zero it first): nodes whose cell index lies outside the grid after that wrap --
 * the reference's fancy indexing raises IndexError for them; the kernel clamps them and counts. */
int fl_grid2mesh(const float* d_grid, const float* d_mesh_pos, float* d_out, int T, int N, int H, int W, int C,
                 float x_min, float y_min, double step_x, double step_y, int32_t* d_out_of_range, void* stream);

/* eagle/Dataloader/IMG_Eagle.py:72-90 (EagleDataset.normalize / denormalize over the pre-gridded states.npy, 4 channels):
 * channel-last values [..., C], C <= 8; denormalize = 0: (x - mean[c]) / std[c], 1: x * std[c] + mean[c]; every operation
 * rounded separately in fp32 like the reference's torch expressions.  In place (d_out == d_in) is allowed. */
int fl_affine_channels(const float* d_in, float* d_out, long n_values, int C, const float* h_mean, const float* h_std,
                       int denormalize, void* stream);

/* ---- per-frame dynamic meshes --------------------------------------------------------------------
 * True EAGLE trajectories (max/ds_download/eagle.py:123-144: pointcloud[T,N,2], triangles[T,F,3], VX/VY/PS per frame)
 * have a different mesh in every frame, so the chain get_mesh_interpolation -> 3 x to_grid -> _pad -> _patch ->
 * _normalize (src/dataloader/mesh_utils.py:82-106, simple_dataloader.py:104-152,193-216) runs once per FRAME: point
 * location becomes part of the per-frame loop.  One call handles a window of frames without host synchronisation.
 *   d_pos f32[T, n_nodes, 2], d_cells i32[T, n_cells, 3], d_velocity f32[T, n_nodes, 2], d_pressure f32[T, n_nodes]
 *   d_grid_ax/d_grid_ay: ONE grid for all frames (the reference's per-frame grid is the same whenever the frames share
 *       their bounding box, as EAGLE's fixed domain does); px, py, crop_patches, flags (FL_FLIP_Y, FL_MASK_AWARE_NORM,
 *       FL_NO_NORM), h_mean/h_std as for fl_plan_patch_table + fl_interp_patchify; FL_FORCE_GATHER (testing) selects
 *       the binned form even when the shared-memory rasterising form would fit
 *   d_states f32[T, L, 3, px, py]; d_mask u8[T, L, px, py] or NULL; d_tri i32[T, L, px, py] or NULL (triangle id of
 *       every output pixel in that frame's mesh, -1 = outside / padding; same tie-break rule as fl_locate)
 *   d_status i32[2], written on the stream: [0] = triangles with a node id outside 0 <= i < n_nodes (their frames are
 *       located without them), [1] = the largest per-frame bin-item count; results are valid iff [0] == 0 and
 *       [1] <= fl_dyn_capacity(...) for the workspace passed (retry with a larger workspace otherwise)
 * Workspace: fl_dyn_workspace_bytes(n_frames, n_cells, nx, ny) bytes or more, 256-B aligned. */
size_t fl_dyn_workspace_bytes(int n_frames, int n_cells, int nx, int ny);
int fl_dyn_capacity(int n_frames, int n_cells, int nx, int ny, size_t workspace_bytes);
int fl_dyn_interp_patchify(const float* d_pos, const int32_t* d_cells, const float* d_velocity, const float* d_pressure,
                           int n_frames, int n_nodes, int n_cells, const float* d_grid_ax, const float* d_grid_ay,
                           int nx, int ny, int px, int py, int crop_patches, const float* h_mean, const float* h_std,
                           unsigned flags, float* d_states, uint8_t* d_mask, int32_t* d_tri, int32_t* d_status,
                           void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- dataset statistics (max/compute_ds_stats.py:20-34,52-62) ---------------------------------
 * Per-channel (n, mean, M2) of states and of diffs (states[t+1]-states[t]) over unmasked pixels.
 * d_states f32[T, L, 3, px, py], d_mask u8[T, L, px, py]; d_agg f64[18] =
 * {state ch0..2, diff ch0..2} x {n, mean, M2}; overwritten.  Masked diff pixels use mask[t+1]. */
size_t fl_stats_workspace_bytes(void);
int fl_ds_stats(const float* d_states, const uint8_t* d_mask, int T, int L, int px, int py,
                double* d_agg, void* d_workspace, size_t workspace_bytes, void* stream);
/* fixed-order Chan merge of n_parts aggregates (each f64[18]) -> d_out f64[18] */
int fl_stats_merge(const double* d_parts, int n_parts, double* d_out, void* stream);

/* ---- patch embedding on the tensor cores (tcgen05) ------------------------------------------------
 * src/models/layers/patch_encoder.py:23-30 + MLP.py:48-54 + input_embeddings.py:36-52 as run under bf16 autocast:
 *   h = LeakyReLU_0.01(bf16(x @ W1^T + b1));  out = bf16(h @ W2^T + b2) + x_emb[p0] + y_emb[p1] + t_emb[p2]  (fp32)
 * d_x_bf16 [n_tokens, in_dim] bf16 (fl_cast_bf16 of the patchified states: a token is one patch, c*256+i*16+j),
 * d_w1_bf16 [hid_dim, in_dim], d_w2_bf16 [out_dim, hid_dim] bf16 (nn.Linear layout), biases fp32,
 * embedding tables fp32 [max_*, out_dim], d_pos_ids int64 [n_tokens, 3] (NULL: no positional add),
 * d_hidden_bf16 [n_tokens, hid_dim] workspace, d_out fp32 [n_tokens, out_dim].
 * in_dim, hid_dim multiples of 64; hid_dim, out_dim multiples of 256. */
int fl_cast_bf16(const float* d_in, void* d_out_bf16, long n, void* stream);
/* rows of `cols` floats -> rows of `out_cols` >= cols bf16 values with a zero tail: tokens of patches whose 3*px*py is not a
 * multiple of 64 (fl_patch_embed's K step) against weights padded with zero columns in the same way */
int fl_cast_bf16_rows(const float* d_in, void* d_out_bf16, long rows, int cols, int out_cols, void* stream);
/* The positional add of fl_patch_embed on its own, over a ring of cached pre-positional embeddings (fl_patch_embed with
 * d_pos_ids = NULL): src/models/model.py:196-199 re-bases the time ids of the whole context at every rollout step, but only ONE
 * state of the context is new -- the others' bf16(h W2^T + b2) rows are unchanged and only their positional term moves.
 * d_pre f32[ctx, B, L, out_dim] ring of per-state embeddings; the context is states start, start+1, ... (mod ctx), c of them;
 * d_pos_ids int64[B, c, L, 3] -> d_out f32[B, c, L, out_dim] = pre + ((x_emb[p0] + y_emb[p1]) + t_emb[p2]), the same
 * association as the fused epilogue (bit-identical results). */
int fl_pos_add_ring(const float* d_pre, const float* d_x_emb, const float* d_y_emb, const float* d_t_emb,
                    const long long* d_pos_ids, int max_x, int max_y, int max_t, float* d_out, int B, int c, int L, int ctx,
                    int start, int out_dim, void* stream);
int fl_patch_embed(const void* d_x_bf16, const void* d_w1_bf16, const float* d_b1, const void* d_w2_bf16, const float* d_b2,
                   const float* d_x_emb, const float* d_y_emb, const float* d_t_emb, const long long* d_pos_ids,
                   int max_x, int max_y, int max_t, void* d_hidden_bf16, float* d_out,
                   int n_tokens, int in_dim, int hid_dim, int out_dim, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLUIDGRID_H */
