"""Per-frame dynamic meshes: CUDA path vs the oracle applied frame by frame."""
import ctypes

import numpy as np
import pytest
import torch

from fluid_llm_b200 import synth
from oracle import mpl_tri
from oracle import pipeline as P

from helpers import PATCH, tie_mesh

pytestmark = pytest.mark.gpu


def _patch_order_tri(tri_index, patch, crop, flip_y):
    """tri_index (T, nx, ny) of the oracle -> (T, L, px, py) in output-pixel order (padding = -1)."""
    T, nx, ny = tri_index.shape
    (bx0, bx1), (by0, by1) = P.pad_amounts(nx, ny, patch)
    img = np.pad(tri_index, ((0, 0), (bx0, bx1), (by0, by1)), constant_values=-1)[:, None]
    if flip_y:
        img = np.ascontiguousarray(img[:, :, :, ::-1])
    if crop:
        img = img[:, :, patch[0] * crop:-patch[0] * crop, patch[1] * crop:-patch[1] * crop]
    p = P.unfold_patches(img, patch)                   # (T, 1, px, py, L)
    return np.ascontiguousarray(p[:, 0].transpose(0, 3, 1, 2))


@pytest.mark.parametrize("kind,pers_name", [("eagle", "cylinder"), ("cylinder", "cylinder"), ("airfoil", "airfoil")])
def test_dynamic_window_matches_oracle(kind, pers_name):
    from fluid_llm_b200.dynamic_mesh import DynamicTrajectory
    from fluid_llm_b200.field_path import AIRFOIL, CYLINDER
    pers = AIRFOIL if pers_name == "airfoil" else CYLINDER
    tr = synth.make_dynamic_trajectory(kind, 7, mesh_seed=1, field_seed=3, flip_frac=0.1)
    if kind == "airfoil":      # the dataset's node crop, done once on the static connectivity's node set
        sel = (tr["mesh_pos"][0][:, 0] > -.5) & (tr["mesh_pos"][0][:, 0] < 2) & (tr["mesh_pos"][0][:, 1] > -.75) & (tr["mesh_pos"][0][:, 1] < .75)
        # keep the bounding box identical in all frames: clamp the kept nodes' drift to the first frame's box
        remap = np.cumsum(sel) - 1
        cells = []
        for t in range(7):
            c = tr["cells"][t]
            keep = sel[c].all(axis=1)
            cells.append(remap[c[keep]].astype(np.int32))
        n_f = min(len(c) for c in cells)
        tr = {"mesh_pos": np.ascontiguousarray(tr["mesh_pos"][:, sel]), "cells": np.stack([c[:n_f] for c in cells]),
              "velocity": np.ascontiguousarray(tr["velocity"][:, sel]), "pressure": np.ascontiguousarray(tr["pressure"][:, sel])}
        lo, hi = tr["mesh_pos"][0].min(axis=0), tr["mesh_pos"][0].max(axis=0)
        extents = (lo[0], hi[0], lo[1], hi[1])
    else:
        extents = None
    want = P.dynamic_ds_get(tr["mesh_pos"], tr["cells"], tr["velocity"], tr["pressure"], 1, 3, 2, personality=pers_name,
                            extents=extents)
    dt = DynamicTrajectory(tr["mesh_pos"], tr["cells"], tr["velocity"], tr["pressure"], extents=extents)
    states, mask, tri = dt.interp_patchify(1, 3, 2, PATCH, pers, want_tri=True)
    assert (dt.nx, dt.ny) == want["tri_index"].shape[1:]
    tri_want = _patch_order_tri(want["tri_index"], PATCH, pers.crop_patches, pers.flip_y)
    assert np.array_equal(tri.cpu().numpy(), tri_want), "triangle ids must be bit-exact in every frame"
    assert np.array_equal(mask.cpu().numpy().astype(bool), want["masks"].astype(bool))
    s = states.cpu().numpy()
    np.testing.assert_allclose(s, want["states"], rtol=1e-6, atol=0)
    assert (s != want["states"]).mean() < 1e-5


def test_dynamic_equals_static_path_on_a_static_mesh():
    """A 'dynamic' trajectory that repeats one mesh must give what the static path (fl_locate + fl_interp_patchify) gives."""
    from fluid_llm_b200.dynamic_mesh import DynamicTrajectory
    from fluid_llm_b200.field_path import CYLINDER, DeviceTrajectory, interp_patchify
    from fluid_llm_b200.mesh_utils import MeshPlan
    tr = synth.make_trajectory("eagle", 5, mesh_seed=2, field_seed=4)
    T = 5
    dt = DynamicTrajectory(np.repeat(tr["mesh_pos"][None], T, 0), np.repeat(tr["cells"][None], T, 0), tr["velocity"], tr["pressure"])
    s_dyn, m_dyn, tri_dyn = dt.interp_patchify(0, T, 1, PATCH, CYLINDER, want_tri=True)
    plan = MeshPlan(tr["mesh_pos"], tr["cells"], 238)
    s_st, m_st, tab = interp_patchify(DeviceTrajectory(tr["velocity"], tr["pressure"], plan), 0, T, 1, PATCH, CYLINDER)
    assert torch.equal(s_dyn, s_st) and torch.equal(m_dyn, m_st)
    assert torch.equal(tri_dyn[0].reshape(-1), tab.idx[:, 3])
    # the binned form (grids too large for shared memory) gives the same bits as the rasterising form
    s_bin, m_bin, tri_bin = dt.interp_patchify(0, T, 1, PATCH, CYLINDER, want_tri=True, force_binned=True)
    assert torch.equal(s_bin, s_dyn) and torch.equal(m_bin, m_dyn) and torch.equal(tri_bin, tri_dyn)


@pytest.mark.parametrize("patch", [(5, 5), (10, 7), (32, 32), (24, 16), (8, 8)])
def test_dynamic_other_patch_sizes(patch):
    """Any patch size: both forms against the static path on a repeated mesh (which is checked against the oracle for the same
    sizes in test_gpu_parity.py::test_other_patch_sizes)."""
    from fluid_llm_b200.dynamic_mesh import DynamicTrajectory
    from fluid_llm_b200.field_path import CYLINDER, DeviceTrajectory, interp_patchify
    from fluid_llm_b200.mesh_utils import MeshPlan
    tr = synth.make_trajectory("cylinder", 3, mesh_seed=1, field_seed=2)
    T = 3
    dt = DynamicTrajectory(np.repeat(tr["mesh_pos"][None], T, 0), np.repeat(tr["cells"][None], T, 0), tr["velocity"], tr["pressure"])
    plan = MeshPlan(tr["mesh_pos"], tr["cells"], 238)
    s_st, m_st, tab = interp_patchify(DeviceTrajectory(tr["velocity"], tr["pressure"], plan), 0, T, 1, patch, CYLINDER)
    for force_binned in (False, True):
        s_dyn, m_dyn, tri_dyn = dt.interp_patchify(0, T, 1, patch, CYLINDER, want_tri=True, force_binned=force_binned)
        assert torch.equal(s_dyn, s_st) and torch.equal(m_dyn, m_st)
        assert torch.equal(tri_dyn[0].reshape(-1), tab.idx[:, 3])


def test_dynamic_binned_and_raster_forms_agree_on_moving_meshes():
    from fluid_llm_b200.dynamic_mesh import DynamicTrajectory
    from fluid_llm_b200.field_path import AIRFOIL, CYLINDER
    for kind, pers in (("eagle", CYLINDER), ("cylinder", AIRFOIL)):
        tr = synth.make_dynamic_trajectory(kind, 9, mesh_seed=4, field_seed=6, flip_frac=0.1)
        dt = DynamicTrajectory(tr["mesh_pos"], tr["cells"], tr["velocity"], tr["pressure"])
        a = dt.interp_patchify(0, 9, 1, PATCH, pers, want_tri=True)
        b = dt.interp_patchify(0, 9, 1, PATCH, pers, want_tri=True, force_binned=True)
        for x, y in zip(a, b):
            assert torch.equal(x, y)


def test_dynamic_five_tuple_and_errors():
    from fluid_llm_b200.dynamic_mesh import DynamicTrajectory
    from fluid_llm_b200.field_path import CYLINDER
    tr = synth.make_dynamic_trajectory("cylinder", 5, mesh_seed=0, field_seed=1)
    dt = DynamicTrajectory(tr["mesh_pos"], tr["cells"], tr["velocity"], tr["pressure"])
    inp, nxt, diffs, masks, pos_ids = dt.ds_get(0, 4, 1, PATCH, CYLINDER)
    want = P.dynamic_ds_get(tr["mesh_pos"], tr["cells"], tr["velocity"], tr["pressure"], 0, 4, 1)
    # values: 1e-6 relative (the north-star tolerance); in practice all but ~1 in 10^5 are bit-identical
    np.testing.assert_allclose(inp.cpu().numpy(), want["states"][:-1], rtol=1e-6, atol=0)
    np.testing.assert_allclose(nxt.cpu().numpy(), want["states"][1:], rtol=1e-6, atol=0)
    assert (nxt.cpu().numpy() != want["states"][1:]).mean() < 1e-4
    assert torch.equal(diffs, nxt - inp)
    assert masks.shape == inp.shape and masks.dtype == torch.bool
    assert np.array_equal(pos_ids.cpu().numpy(), P.get_pos_id(4, want["N_x_patch"], want["N_y_patch"]))
    with pytest.raises(ValueError):
        dt.interp_patchify(3, 4, 1, PATCH, CYLINDER)                  # runs past the end
    bad = tr["cells"].copy()
    bad[2, 5, 1] = tr["mesh_pos"].shape[1] + 3
    with pytest.raises(ValueError):
        DynamicTrajectory(tr["mesh_pos"], bad, tr["velocity"], tr["pressure"]).interp_patchify(0, 4, 1, PATCH, CYLINDER)
    moved = tr["mesh_pos"].copy()
    moved[3, :, 0] *= 1.01
    with pytest.raises(ValueError):
        DynamicTrajectory(moved, tr["cells"], tr["velocity"], tr["pressure"])   # bounding boxes differ, no extents given


def test_dynamic_c_abi_argument_errors():
    from fluid_llm_b200._lib import load
    lib = load()
    assert lib.fl_dyn_workspace_bytes(0, 10, 10, 10) == 0
    assert lib.fl_dyn_workspace_bytes(4, 1000, 238, 60) > 0
    rc = lib.fl_dyn_interp_patchify(None, None, None, None, 1, 1, 1, None, None, 1, 1, 16, 16, 0, None, None, 0, None, None,
                                    None, None, None, 0, None)
    assert rc == -1 and b"null pointer" in lib.fl_last_error()


def test_dynamic_long_window_chunks():
    """More frames than one binning chunk (128): frame 0 and the last frame against the oracle."""
    from fluid_llm_b200.dynamic_mesh import DynamicTrajectory
    from fluid_llm_b200.field_path import CYLINDER
    tr = synth.make_dynamic_trajectory("cylinder", 4, mesh_seed=3, field_seed=5)
    reps = 70                                                   # 280 frames = 3 chunks, the last one partial
    big = {k: np.concatenate([v] * reps, axis=0) for k, v in tr.items()}
    dt = DynamicTrajectory(big["mesh_pos"], big["cells"], big["velocity"], big["pressure"])
    states, mask, _ = dt.interp_patchify(0, 4 * reps, 1, PATCH, CYLINDER)
    want = P.dynamic_ds_get(tr["mesh_pos"], tr["cells"], tr["velocity"], tr["pressure"], 0, 4, 1)
    s = states.cpu().numpy().reshape(reps, 4, *states.shape[1:])
    for r in (0, 31, 32, 63, 64, reps - 1):
        assert np.array_equal(s[r], want["states"]), r
    assert np.array_equal(mask.cpu().numpy().reshape(reps, 4, *mask.shape[1:])[reps - 1].astype(bool), want["masks"].astype(bool))


@pytest.mark.parametrize("force_binned", [False, True])
def test_dynamic_tie_breaks_on_the_hand_built_mesh(force_binned):
    """Grid points on vertices, on horizontal / vertical / oblique edges, on the boundary and in holes: both forms of the
    dynamic path must return the trapezoid map's triangle for every one of them (small patches so the 9 x 9 ... 33 x 33
    grids are not all padding)."""
    from fluid_llm_b200.dynamic_mesh import DynamicTrajectory
    from fluid_llm_b200.field_path import CYLINDER
    pos, tris = tie_mesh()
    rng = np.random.default_rng(3)
    T = 3
    vel = rng.standard_normal((T, len(pos), 2)).astype(np.float32)
    prs = rng.standard_normal((T, len(pos))).astype(np.float32)
    perm = [tris, tris[::-1].copy(), tris[rng.permutation(len(tris))]]          # triangle ids differ per frame
    for res in (9, 5, 17, 33):
        dt = DynamicTrajectory(np.repeat(pos[None], T, 0), np.stack(perm), vel, prs, grid_res=res)
        _, mask, tri = dt.interp_patchify(0, T, 1, (8, 4), CYLINDER, want_tri=True, force_binned=force_binned)
        for t in range(T):
            triang = mpl_tri.Triangulation(pos[:, 0], pos[:, 1], triangles=perm[t])
            gx, gy = P.grid_pos(pos[:, 0].min(), pos[:, 0].max(), pos[:, 1].min(), pos[:, 1].max(), res)
            want = triang.get_trifinder()(gx, gy)
            got = _patch_order_tri(want[None], (8, 4), 0, False)[0]
            assert np.array_equal(tri[t].cpu().numpy(), got), (res, t)
            assert np.array_equal(mask[t].cpu().numpy().astype(bool), got < 0)


@pytest.mark.parametrize("force_binned", [False, True])
def test_dynamic_writes_stay_inside_the_output_buffers(force_binned):
    """Straight through the C ABI with guard bands around every output: nothing outside [T, L, ...] may change."""
    import ctypes
    from fluid_llm_b200._lib import FL_FORCE_GATHER, check, load, ptr, stream_ptr
    from fluid_llm_b200.mesh_utils import _grid_axes
    tr = synth.make_dynamic_trajectory("cylinder", 5, mesh_seed=5, field_seed=7)
    T, N, F = 5, tr["mesh_pos"].shape[1], tr["cells"].shape[1]
    lo, hi = tr["mesh_pos"][0].min(axis=0), tr["mesh_pos"][0].max(axis=0)
    ax, ay = _grid_axes(float(lo[0]), float(hi[0]), float(lo[1]), float(hi[1]), 238, "1.26")
    nx, ny = len(ax), len(ay)
    L = ((nx + 15) // 16) * ((ny + 15) // 16)
    G = 4096                                                    # guard elements on both sides
    dev = "cuda"
    pos = torch.from_numpy(tr["mesh_pos"]).to(dev)
    cells = torch.from_numpy(tr["cells"]).to(dev)
    vel = torch.from_numpy(tr["velocity"]).to(dev)
    prs = torch.from_numpy(np.ascontiguousarray(tr["pressure"][:, :, 0])).to(dev)
    ax_d, ay_d = torch.from_numpy(ax).to(dev), torch.from_numpy(ay).to(dev)
    n_out = T * L * 256
    states = torch.full((G + 3 * n_out + G,), 777.0, dtype=torch.float32, device=dev)
    mask = torch.full((G + n_out + G,), 77, dtype=torch.uint8, device=dev)
    tri = torch.full((G + n_out + G,), -777, dtype=torch.int32, device=dev)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    lib = load()
    ws = torch.empty(int(lib.fl_dyn_workspace_bytes(T, F, nx, ny)) + 256, dtype=torch.uint8, device=dev)
    ws_tail = ws[-256:]
    ws_tail.fill_(99)
    m = (ctypes.c_float * 3)(0.823, 0.0005865, 0.04763)
    s = (ctypes.c_float * 3)(0.275, 0.275, 0.275)
    check(lib.fl_dyn_interp_patchify(ptr(pos), ptr(cells), ptr(vel), ptr(prs), T, N, F, ptr(ax_d), ptr(ay_d), nx, ny, 16, 16, 0, m, s,
                                     FL_FORCE_GATHER if force_binned else 0, ptr(states[G:]), ptr(mask[G:]), ptr(tri[G:]),
                                     ptr(status), ptr(ws), ws.numel() - 256, stream_ptr()), "fl_dyn_interp_patchify")
    torch.cuda.synchronize()
    assert int(status[0]) == 0
    for buf, fill in ((states, 777.0), (mask, 77), (tri, -777)):
        assert bool((buf[:G] == fill).all()) and bool((buf[-G:] == fill).all())
    assert bool((ws_tail == 99).all())
    inner = states[G:-G]
    assert bool(torch.isfinite(inner).all()) and not bool((inner == 777.0).any())          # every output element was written
    assert not bool((mask[G:-G] == 77).any()) and not bool((tri[G:-G] == -777).any())


@pytest.mark.parametrize("pers_name", ["cylinder", "airfoil"])
def test_dynamic_nonfinite_node_values(pers_name):
    """NaN / inf node values on per-frame meshes: per channel zeroing (mesh_utils.py:86-89), only the pressure mask is kept,
    the mask-aware personality leaves masked pixels un-normalised -- in both forms of the dynamic path."""
    from fluid_llm_b200.dynamic_mesh import DynamicTrajectory
    from fluid_llm_b200.field_path import AIRFOIL, CYLINDER
    pers = AIRFOIL if pers_name == "airfoil" else CYLINDER
    tr = synth.make_dynamic_trajectory("cylinder", 5, mesh_seed=6, field_seed=8)
    tr["pressure"][2, 50:60, 0] = np.nan
    tr["velocity"][3, 70, 0] = np.inf
    tr["velocity"][1, 80, 1] = -np.inf
    with np.errstate(all="ignore"):
        want = P.dynamic_ds_get(tr["mesh_pos"], tr["cells"], tr["velocity"], tr["pressure"], 0, 5, 1, personality=pers_name)
    dt = DynamicTrajectory(tr["mesh_pos"], tr["cells"], tr["velocity"], tr["pressure"])
    for fb in (False, True):
        states, mask, _ = dt.interp_patchify(0, 5, 1, PATCH, pers, force_binned=fb)
        assert np.array_equal(mask.cpu().numpy().astype(bool), want["masks"].astype(bool))
        s = states.cpu().numpy()
        assert np.isfinite(s).all()
        np.testing.assert_allclose(s, want["states"], rtol=1e-6, atol=0)
