"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU restatement (NumPy) of the Python side of FLUID-LLM's per-timestep field data path.
Every function cites the reference lines it follows.  The matplotlib C++ arithmetic is
restated in oracle/tri_oracle.cpp (front: oracle/mpl_tri.py).

PARITY STATUS.  The reference holds no golden vectors for this path (it has no tests).  This
restatement is pinned the only way available here: tests/test_oracle.py (test_oracle_vs_reference_*_live) imports
the reference's UNMODIFIED Python from /root/reference behind stub modules for the three
packages that are not installed (oracle/stubs: matplotlib -> oracle/mpl_tri.py, cprint,
natsort) and compares every stage bit-for-bit; oracle/make_golden.py freezes those outputs in
tests/golden/.  The matplotlib C++ layer itself (trapezoid-map tie semantics, plane
coefficients) stays "parity unpinned": matplotlib 3.8.2 is absent, so it is restated from its
published algorithm and cross-checked only against an independent brute-force rule.

numpy_semantics: the reference environment pins NumPy 1.26.3 (environemnt.yml:177) whose
value-based scalar promotion differs from the NumPy 2.x installed here (NEP 50).  "1.26" (the
default, the parity target) hard-codes the pinned behaviour; "2.x" reproduces what the
reference's code does when imported under this container's NumPy (used only by the
cross-check against the imported reference).
"""
from __future__ import annotations

import numpy as np

from . import mpl_tri

F32 = np.float32
F64 = np.float64


# --------------------------------------------------------------------------------------------
# mesh -> grid  (src/dataloader/mesh_utils.py)
# --------------------------------------------------------------------------------------------
def grid_shape(x_min, x_max, y_min, y_max, grid_res, numpy_semantics="1.26"):
    """mesh_utils.py:66-76 -- long axis gets grid_res points, short axis int(grid_res*ratio)."""
    x_min, x_max, y_min, y_max = F32(x_min), F32(x_max), F32(y_min), F32(y_max)
    dx, dy = F32(x_max - x_min), F32(y_max - y_min)
    long_axis, short_axis = max(dx, dy), min(dx, dy)
    ratio = F32(short_axis / long_axis)
    if numpy_semantics == "1.26":      # int * float32 scalar -> float64 (legacy promotion)
        n_short = int(F64(grid_res) * F64(ratio))
    else:                              # NEP 50: python int is weak -> float32
        n_short = int(F32(F32(grid_res) * ratio))
    if dx > dy:
        return int(grid_res), n_short
    return n_short, int(grid_res)


def grid_axis(start, stop, n, numpy_semantics="1.26"):
    """np.mgrid[start:stop:n*1j] for float32 scalar bounds, one axis (mesh_utils.py:78).

    numpy/lib/index_tricks.py nd_grid.__getitem__: values = indices * step + start with
    step = (stop - start) / float(n - 1).  NumPy 1.26: float32 - float32 -> float32 scalar,
    then everything promotes to float64; NumPy 2.x: everything stays float32."""
    start, stop = F32(start), F32(stop)
    i = np.arange(n)
    if numpy_semantics == "1.26":
        step = F64(F32(stop - start)) / F64(n - 1) if n != 1 else F64(n)
        return (i.astype(F64) * step + F64(start)).astype(F32)
    step = F32(F32(stop - start) / F32(n - 1)) if n != 1 else F32(n)
    return (i.astype(F32) * step + start).astype(F32)


def grid_pos(x_min, x_max, y_min, y_max, grid_res, numpy_semantics="1.26"):
    """mesh_utils.py:64-79 -> grid_x, grid_y float32 (nx, ny), index order [ix, iy]."""
    nx, ny = grid_shape(x_min, x_max, y_min, y_max, grid_res, numpy_semantics)
    ax = grid_axis(x_min, x_max, nx, numpy_semantics)
    ay = grid_axis(y_min, y_max, ny, numpy_semantics)
    gx = np.ascontiguousarray(np.broadcast_to(ax[:, None], (nx, ny)))
    gy = np.ascontiguousarray(np.broadcast_to(ay[None, :], (nx, ny)))
    return gx, gy


def get_mesh_interpolation(pos, faces, grid_res=238, numpy_semantics="1.26"):
    """mesh_utils.py:94-106."""
    pos = np.asarray(pos)
    x_min, y_min = np.min(pos, axis=0)
    x_max, y_max = np.max(pos, axis=0)
    grid_x, grid_y = grid_pos(x_min, x_max, y_min, y_max, grid_res, numpy_semantics)
    triang = mpl_tri.Triangulation(pos[:, 0], pos[:, 1], triangles=faces)
    tri_index = triang.get_trifinder()(grid_x, grid_y)
    return triang, tri_index, grid_x, grid_y


def to_grid(val, grid_x, grid_y, triang, tri_index):
    """mesh_utils.py:82-91 + _triinterpolate.py:181-206,262-263,278-280.

    fp64 plane a*x + b*y + c (separate mul, mul, add, add), stored into a float32 array (one
    rounding), cells outside the mesh NaN -> masked_invalid -> masked cells zeroed."""
    val = np.asarray(val)
    if val.shape != triang.x.shape:
        raise ValueError("z array must have same length as triangulation x and y arrays")
    if grid_x.shape != grid_y.shape:
        raise ValueError(f"x and y shall have same shapes. Given: {grid_x.shape} and {grid_y.shape}")
    if tri_index.shape != grid_x.shape:
        raise ValueError("tri_index array is provided and shall have same shape as x and y. "
                         f"Given: {tri_index.shape} and {grid_x.shape}")
    plane = triang.calculate_plane_coefficients(val)
    x, y, ti = np.ravel(grid_x), np.ravel(grid_y), np.ravel(tri_index)
    inside = ti != -1
    tv = ti[inside]
    ret = np.empty(x.size, dtype=grid_x.dtype)
    ret[~inside] = np.nan
    ret[inside] = plane[tv, 0] * x[inside] + plane[tv, 1] * y[inside] + plane[tv, 2]
    ret = ret.reshape(grid_x.shape)
    mask = ~np.isfinite(ret)
    ret[mask] = 0.0
    return ret, mask


# --------------------------------------------------------------------------------------------
# dataset item  (src/dataloader/simple_dataloader.py, airfoil_ds.py)
# --------------------------------------------------------------------------------------------
def num_patches(dim_size, kern_size, stride, padding=0):
    """simple_dataloader.py:16-20."""
    return (dim_size + 2 * padding - kern_size) // stride + 1


def pad_amounts(w, h, patch_size):
    """simple_dataloader.py:139-147 -> ((before_x, after_x), (before_y, after_y))."""
    pw, ph = (-w) % patch_size[0], (-h) % patch_size[1]
    return (pw // 2, pw - pw // 2), (ph // 2, ph - ph // 2)


def pad_state(state, mask, patch_size):
    """simple_dataloader.py:137-152: state padded with 0, mask with 1."""
    _, w, h = state.shape
    px, py = pad_amounts(w, h, patch_size)
    return (np.pad(state, ((0, 0), px, py), mode="constant", constant_values=0),
            np.pad(mask, (px, py), mode="constant", constant_values=1))


def get_step(triang, tri_index, grid_x, grid_y, velocity, pressure, step_num, patch_size, pad=True):
    """simple_dataloader.py:104-121: three to_grid calls, only the pressure mask is kept."""
    vx, _ = to_grid(velocity[step_num][:, 0], grid_x, grid_y, triang, tri_index)
    vy, _ = to_grid(velocity[step_num][:, 1], grid_x, grid_y, triang, tri_index)
    p, p_mask = to_grid(pressure[step_num][:, 0], grid_x, grid_y, triang, tri_index)
    state = np.stack([vx, vy, p], axis=0)
    if pad:
        state, p_mask = pad_state(state, p_mask, patch_size)
    return state, p_mask


def airfoil_crop(pos, faces):
    """airfoil_ds.py:164-183 -> (node_mask, cropped pos, renumbered faces)."""
    pos = np.asarray(pos)
    faces = np.asarray(faces)
    mask = (pos[:, 0] > -.5) & (pos[:, 0] < 2) & (pos[:, 1] > -.75) & (pos[:, 1] < 0.75)
    wanted = np.nonzero(mask)[0]
    all_nodes = np.zeros(len(mask), dtype=np.int64)
    all_nodes[mask] = np.arange(len(wanted), dtype=np.int64)
    face_mask = np.isin(faces, wanted).all(axis=1)
    return mask, pos[mask], all_nodes[faces[face_mask]]


def full_seq(traj, step_num, seq_len, seq_interval, resolution, patch_size, personality,
             pad=True, numpy_semantics="1.26"):
    """simple_dataloader.py:166-191 / airfoil_ds.py:158-214 -> float32 (T, 4, Hp, Wp)."""
    pos, faces = traj["mesh_pos"], traj["cells"]
    vel, prs = traj["velocity"], traj["pressure"]
    if personality == "airfoil":
        nmask, pos, faces = airfoil_crop(pos, faces)
        vel, prs = vel[:, nmask], prs[:, nmask]
    triang, tri_index, gx, gy = get_mesh_interpolation(pos, faces, resolution, numpy_semantics)
    out = []
    for i in range(step_num, step_num + seq_len * seq_interval, seq_interval):
        state, mask = get_step(triang, tri_index, gx, gy, vel, prs, i, patch_size, pad)
        out.append(np.concatenate([state, mask[None].astype(state.dtype)], axis=0))
    return np.stack(out).astype(F32), tri_index


def unfold_patches(img, patch_size, stride=None):
    """F.unfold(kernel=patch, stride) + view, simple_dataloader.py:123-135.

    img (B, C, X, Y) -> (B, C, px, py, L) with L = nbx*nby, l = bx*nby + by, nb* = floor((extent - patch) / stride) + 1
    (F.unfold drops what is left over); patch (bx, by) starts at pixel (bx*sx, by*sy)."""
    B, C, X, Y = img.shape
    px, py = patch_size
    sx, sy = patch_size if stride is None else stride
    nbx, nby = max((X - px) // sx + 1, 0), max((Y - py) // sy + 1, 0)
    out = np.empty((B, C, px, py, nbx * nby), dtype=img.dtype)
    for bx in range(nbx):
        for by in range(nby):
            out[..., bx * nby + by] = img[:, :, bx * sx:bx * sx + px, by * sy:by * sy + py]
    return out


CYL_MEANS = np.array([0.823, 0.0005865, 0.04763], dtype=F32)       # simple_dataloader.py:205-209
CYL_STDS = np.array([0.275, 0.275, 0.275], dtype=F32)              # simple_dataloader.py:210
AIR_MEANS = np.array([170.1, -1.183, 9.935e+04], dtype=F32)        # airfoil_ds.py:228-232
AIR_STDS = np.array([50, 50, 6197], dtype=F32)                     # airfoil_ds.py:233 (int64 -> promoted)


def normalize(states, masks, personality, means=None, stds=None):
    """simple_dataloader.py:193-216 (all pixels) / airfoil_ds.py:216-244 (unmasked pixels only).

    float32 subtract, then float32 divide (two roundings)."""
    if means is None:
        means = CYL_MEANS if personality == "cylinder" else AIR_MEANS
    if stds is None:
        stds = CYL_STDS if personality == "cylinder" else AIR_STDS
    m = np.asarray(means, dtype=F32).reshape(1, 1, 3, 1, 1)
    s = np.asarray(stds, dtype=F32).reshape(1, 1, 3, 1, 1)
    states = states.astype(F32)
    normed = ((states - m).astype(F32) / s).astype(F32)
    if personality == "cylinder":
        return normed
    keep = np.broadcast_to(masks[:, :, None].astype(bool), states.shape)
    return np.where(keep, states, normed)


def get_pos_id(seq_len, n_x_patch, n_y_patch):
    """simple_dataloader.py:218-226 (labelling quirk reproduced as is) -> int64 (T-1, L, 3)."""
    n_patch = n_x_patch * n_y_patch
    a = np.arange((seq_len - 1) * n_patch)
    ids = np.stack([a % n_x_patch, (a // n_x_patch) % n_y_patch, a // n_patch], axis=1)
    return ids.reshape(seq_len - 1, n_patch, 3).astype(np.int64)


def ds_get(traj, step_num, seq_len, seq_interval=1, resolution=238, patch_size=(16, 16),
           personality="cylinder", normalize_ds=True, pad=True, numpy_semantics="1.26",
           means=None, stds=None, return_all=False, stride=None, n_patch=None):
    """simple_dataloader.py:72-102 / airfoil_ds.py:71-103 -> the 5-tuple (NumPy arrays)."""
    seq, tri_index = full_seq(traj, step_num, seq_len, seq_interval, resolution, patch_size,
                              personality, pad, numpy_semantics)
    stride = patch_size if stride is None else stride
    # simple_dataloader.py:52-56 / airfoil_ds.py:52-55: the patch-count attributes come from the uncropped frame (minus 2 for
    # the airfoil), which is what unfold yields only for stride == patch_size; the position ids are made from the attributes
    ring = 2 if personality == "airfoil" else 0
    nxp = num_patches(seq.shape[2], patch_size[0], stride[0]) - ring
    nyp = num_patches(seq.shape[3], patch_size[1], stride[1]) - ring
    if n_patch is not None:         # the data set's attributes come from ITS probe file (save_files[1], :46-47), whose unpadded
        nxp, nyp = n_patch          # grid may differ from this trajectory's: the caller passes them
    if personality == "airfoil":
        seq = np.ascontiguousarray(seq[:, :, :, ::-1])                       # airfoil_ds.py:80
        seq = seq[:, :, patch_size[0]:-patch_size[0], patch_size[1]:-patch_size[1]]  # :132-133
    patches = unfold_patches(seq, patch_size, stride)                          # (T, 4, px, py, L)
    states = np.ascontiguousarray(patches[:, :-1].transpose(0, 4, 1, 2, 3))    # (T, L, 3, px, py)
    masks = np.ascontiguousarray(patches[:, -1].transpose(0, 3, 1, 2))         # (T, L, px, py)
    if normalize_ds:
        states = normalize(states, masks, personality, means, stds)
    diffs = (states[1:] - states[:-1]).astype(F32)
    bc = np.repeat(masks[1:, :, None], 3, axis=2).astype(bool)
    out = (states[:-1], states[1:], diffs, bc, get_pos_id(seq_len, nxp, nyp))
    if return_all:
        return out, dict(states=states, masks=masks, tri_index=tri_index, N_x_patch=nxp, N_y_patch=nyp)
    return out


def dynamic_ds_get(mesh_pos, cells, velocity, pressure, step_num, seq_len, seq_interval=1, resolution=238,
                   patch_size=(16, 16), personality="cylinder", normalize_ds=True, extents=None,
                   numpy_semantics="1.26", means=None, stds=None):
    """Per-frame dynamic meshes (max/ds_download/eagle.py:123-144 data fed to the static path): for every frame
    mesh_utils.py:94-106 on THAT frame's mesh, then simple_dataloader.py:104-121 and the rest of ds_get.  The grid is
    the one of `extents` (default: the bounding box of the first selected frame, which the reference would use when all
    frames share it).  -> dict(states (T,L,3,px,py), masks (T,L,px,py), tri_index (T,nx,ny), N_x_patch, N_y_patch)."""
    frames = list(range(step_num, step_num + seq_len * seq_interval, seq_interval))
    if extents is None:
        p0 = np.asarray(mesh_pos[frames[0]])
        (x_min, y_min), (x_max, y_max) = np.min(p0, axis=0), np.max(p0, axis=0)
    else:
        x_min, x_max, y_min, y_max = (F32(e) for e in extents)
    gx, gy = grid_pos(x_min, x_max, y_min, y_max, resolution, numpy_semantics)
    out, tris = [], []
    for t in frames:
        pos = np.asarray(mesh_pos[t])
        triang = mpl_tri.Triangulation(pos[:, 0], pos[:, 1], triangles=np.asarray(cells[t]))
        tri_index = triang.get_trifinder()(gx, gy)
        prs = np.asarray(pressure)
        prs = prs if prs.ndim == 3 else prs[:, :, None]
        state, mask = get_step(triang, tri_index, gx, gy, velocity, prs, t, patch_size, True)
        out.append(np.concatenate([state, mask[None].astype(state.dtype)], axis=0))
        tris.append(tri_index)
    seq = np.stack(out).astype(F32)
    if personality == "airfoil":
        seq = np.ascontiguousarray(seq[:, :, :, ::-1])
        seq = seq[:, :, patch_size[0]:-patch_size[0], patch_size[1]:-patch_size[1]]
    patches = unfold_patches(seq, patch_size)
    states = np.ascontiguousarray(patches[:, :-1].transpose(0, 4, 1, 2, 3))
    masks = np.ascontiguousarray(patches[:, -1].transpose(0, 3, 1, 2))
    if normalize_ds:
        states = normalize(states, masks, personality, means, stds)
    X, Y = seq.shape[2:]
    return dict(states=states, masks=masks, tri_index=np.stack(tris), N_x_patch=X // patch_size[0],
                N_y_patch=Y // patch_size[1], grid_x=gx, grid_y=gy)


def img_mgn_item(traj, t, window_length, kind, numpy_semantics="1.26"):
    """eagle/Dataloader/IMG_MGN.py:46-128,141-157 -> states float32 (T, H, W, 3) normalised, mask bool (T, H, W)."""
    pos, faces, vel, prs = traj["mesh_pos"], traj["cells"], traj["velocity"], traj["pressure"]
    if kind == "airfoil":
        nmask, pos, faces = airfoil_crop(pos, faces)
        vel, prs = vel[:, nmask], prs[:, nmask]
    triang, tri_index, gx, gy = get_mesh_interpolation(pos, faces, 238, numpy_semantics)
    states, masks = [], []
    for i in range(t, t + window_length):
        st, m = get_step(triang, tri_index, gx, gy, vel, prs, i, None, pad=False)
        if kind == "airfoil":
            st, m = st[:, 16:-16, 16:-16], m[16:-16, 16:-16]
        states.append(st)
        masks.append(m)
    states = np.stack(states).astype(F32)
    if kind == "airfoil":
        means, stds = (170.1, -1.183, 9.935e+04), (71.06, 46.73, 8964)
    else:
        means, stds = (0.823, 0.0005865, 0.04763), (0.275, 0.275, 0.275)
    m = np.asarray(means, dtype=F32).reshape(1, 3, 1, 1)
    s = np.asarray(stds, dtype=F32).reshape(1, 3, 1, 1)
    states = ((states - m).astype(F32) / s).astype(F32)
    return np.ascontiguousarray(states.transpose(0, 2, 3, 1)), np.stack(masks)


# --------------------------------------------------------------------------------------------
# inverse path  (src/utils_model.py, src/models/model.py, eagle/Dataloader/IMG_Eagle.py)
# --------------------------------------------------------------------------------------------
def patch_to_img(patches, nx_patch, ny_patch):
    """utils_model.py:77-92 (F.fold, non-overlapping): (bs,T,L,C,px,py) -> (bs,T,C,X,Y)."""
    bs, T, L, C, px, py = patches.shape
    v = patches.reshape(bs, T, nx_patch, ny_patch, C, px, py).transpose(0, 1, 4, 2, 5, 3, 6)
    return np.ascontiguousarray(v).reshape(bs, T, C, nx_patch * px, ny_patch * py)


def img_to_patch(img, patch_size):
    """utils_model.py:95-109: (bs,T,C,X,Y) -> (bs,T,L,C,px,py)."""
    bs, T, C, X, Y = img.shape
    p = unfold_patches(img.reshape(bs * T, C, X, Y), patch_size)               # (bsT, C, px, py, L)
    return np.ascontiguousarray(p.transpose(0, 4, 1, 2, 3)).reshape(bs, T, -1, C, *patch_size)


def rollout_step(last_state, diff_img, mask, patch_size):
    """model.py:164,206,210: diffs = img_to_patch(pred); diffs[mask] = 0; next = last + diffs."""
    diffs = img_to_patch(diff_img, patch_size).copy()
    diffs[mask] = 0
    return (last_state + diffs).astype(last_state.dtype), diffs


IMG_EAGLE_MEAN = np.array([-0.0147, 0.2125, -0.5327, 3.7694], dtype=np.float32)      # eagle/Dataloader/IMG_Eagle.py:76,86
IMG_EAGLE_STD = np.array([1.5943, 1.8824, 6.3553, 9.0565], dtype=np.float32)         # :77,87


def img_eagle_normalize(state):
    """eagle/Dataloader/IMG_Eagle.py:72-80: (state - mean) / std over the 4 channels, float32 throughout."""
    s = np.asarray(state, dtype=np.float32)
    return ((s.reshape(-1, 4) - IMG_EAGLE_MEAN) / IMG_EAGLE_STD).astype(np.float32).reshape(s.shape)


def img_eagle_denormalize(state):
    """eagle/Dataloader/IMG_Eagle.py:82-90: state * std + mean (a rounded product, then a rounded sum)."""
    s = np.asarray(state, dtype=np.float32)
    return ((s.reshape(-1, 4) * IMG_EAGLE_STD).astype(np.float32) + IMG_EAGLE_MEAN).astype(np.float32).reshape(s.shape)


def img_eagle_item(states, pixel_type, window_length, mode):
    """eagle/Dataloader/IMG_Eagle.py:37-49 for the deterministic modes: window start 550 (or 1 for the full 990), the
    normalised window, the pixel-type mask."""
    t = 1 if window_length == 990 else 550
    assert mode in ("test", "valid")
    return img_eagle_normalize(states[t:t + window_length]), np.array(pixel_type)


def _floor_divide_f(a, b, dtype):
    """numpy's npy_floor_divide for floats, evaluated in `dtype` (NumPy C source semantics)."""
    a = np.asarray(a, dtype=dtype)
    b = dtype(b)
    mod = np.fmod(a, b).astype(dtype)
    div = ((a - mod).astype(dtype) / b).astype(dtype)
    adj = (mod != 0) & ((b < 0) != (mod < 0))
    div = np.where(adj, (div - dtype(1)).astype(dtype), div)
    fl = np.floor(div).astype(dtype)
    fl = np.where((div - fl).astype(dtype) > dtype(0.5), (fl + dtype(1)).astype(dtype), fl)
    zero = np.copysign(dtype(0), (a / b).astype(dtype))
    return np.where(div != 0, fl, zero).astype(dtype)


def grid2mesh_index(mesh_pos_t, numpy_semantics="1.26"):
    """IMG_Eagle.py:95-112 -> (index_y, index_x) int64 for one timestep's node positions."""
    Xmin, Xmax, Ymin, Ymax, LENGTH, HEIGHT = -2.5, 2.5, -1.7, 1.5, 256, 128
    x, y = np.linspace(Xmin, Xmax, LENGTH), np.linspace(Ymax, Ymin, HEIGHT)
    step_x, step_y = x[1] - x[0], y[1] - y[0]
    px = np.asarray(mesh_pos_t[:, 0], dtype=F32)
    py = np.asarray(mesh_pos_t[:, 1], dtype=F32)
    if numpy_semantics == "1.26":   # float32 array with python/np.float64 scalars stays float32
        ax = ((px - F32(Xmin)).astype(F32) + F32(step_x / 2)).astype(F32)
        ay = ((py - F32(Ymin)).astype(F32) + F32(step_y / 2)).astype(F32)
        ix = _floor_divide_f(ax, F32(step_x), F32)
        iy = _floor_divide_f(ay, F32(-step_y), F32)
    else:                           # NEP 50: np.float64 scalars are strong -> float64
        ax = (px - F32(Xmin)).astype(F32).astype(F64) + step_x / 2
        ay = (py - F32(Ymin)).astype(F32).astype(F64) + step_y / 2
        ix = _floor_divide_f(ax, F64(step_x), F64)
        iy = _floor_divide_f(ay, F64(-step_y), F64)
    return iy.astype(np.int64), ix.astype(np.int64)


def grid2mesh(velocity_grid, pressure_grid, mesh_pos, numpy_semantics="1.26"):
    """IMG_Eagle.py:93-123: nearest-cell gather after flipping the grid rows."""
    vg = np.flip(np.asarray(velocity_grid), axis=1)
    pg = np.flip(np.asarray(pressure_grid), axis=1)
    vm, pm = [], []
    for t in range(mesh_pos.shape[0]):
        iy, ix = grid2mesh_index(mesh_pos[t], numpy_semantics)
        vm.append(vg[t][iy, ix])
        pm.append(pg[t][iy, ix])
    return np.stack(vm), np.stack(pm)


def get_nrmse(true_states, pred_states, mesh_pos, faces, numpy_semantics="1.26"):
    """eagle/eagle_utils.py:60-130 -> float32 (1, seq_len): velocity + pressure N-RMSE on the grid."""
    seq_len = true_states.shape[1]
    triang, tri_index, gx, gy = get_mesh_interpolation(mesh_pos[0, 0], faces[0, 0], 238, numpy_semantics)
    t_img, p_img = [], []
    for i in range(seq_len):
        ts, ps = [], []
        for j in range(3):
            a, mask = to_grid(true_states[0, i][:, j], gx, gy, triang, tri_index)
            b, _ = to_grid(pred_states[0, i][:, j], gx, gy, triang, tri_index)
            ts.append(a)
            ps.append(b)
        t_img.append(np.stack(ts))
        p_img.append(np.stack(ps))
    t_img, p_img = np.stack(t_img)[None], np.stack(p_img)[None]
    m = np.broadcast_to(mask[None, None, None], t_img.shape)

    def aux(p, t, mm):
        err = ((p - t) * (~mm)).astype(F32)
        return np.sqrt((err * err).mean(axis=(-1, -2, -3), dtype=F32))
    return aux(p_img[:, :, :2], t_img[:, :, :2], m[:, :, :2]) + aux(p_img[:, :, 2:], t_img[:, :, 2:], m[:, :, 2:])


# --------------------------------------------------------------------------------------------
# dataset statistics  (max/compute_ds_stats.py)
# --------------------------------------------------------------------------------------------
def update_variance_batch(agg, new_values):
    """compute_ds_stats.py:20-30, evaluated in float64 (the product's definition)."""
    count, mean, m2 = agg
    v = np.asarray(new_values, dtype=F64)
    new_count = count + len(v)
    delta = v - mean
    mean = mean + np.sum(delta) / new_count
    m2 = m2 + np.sum(delta * (v - mean))
    return new_count, float(mean), float(m2)


def get_std(agg):
    """compute_ds_stats.py:33-34."""
    return float(np.sqrt(agg[2] / agg[0]))


def chan_merge(a, b):
    """Pairwise merge of (n, mean, M2) aggregates (Chan et al.); used for the rank merge."""
    na, ma, sa = a
    nb, mb, sb = b
    if na == 0:
        return b
    if nb == 0:
        return a
    n = na + nb
    d = mb - ma
    return n, ma + d * nb / n, sa + sb + d * d * na * nb / n


def ds_stats(states, diffs, bc_mask):
    """compute_ds_stats.py:52-62 for one sample: per-channel (n, mean, M2) of states and diffs
    over unmasked pixels.  states/diffs (T, L, 3, px, py), bc_mask bool same shape."""
    out = []
    for arr in (states, diffs):
        for j in range(3):
            sel = arr[:, :, j][~bc_mask[:, :, j]]
            out.append(update_variance_batch((0, 0.0, 0.0), sel))
    return out
