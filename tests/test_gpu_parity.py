"""-m gpu: the CUDA path (through the C ABI) against the oracle on the same seeded inputs."""
import numpy as np
import pytest
import torch

from oracle import mpl_tri
from oracle import pipeline as P

from helpers import PATCH, oracle_ds_get, personality_of, tie_mesh, trajectory

pytestmark = pytest.mark.gpu


KERNELS = ["staged", "staged_idx32", "gather", "tiled", "ring"]      # staged_idx32: 16-byte table records instead of the compact ones


def _kernel_args(kernel, ppx=256):
    """interp_patchify / TrajBatch arguments that force one of the per-step kernels (csrc/fl_interp.cu, fl_tiled.cu, and the
    experimental fl_ring.cu)."""
    return {"staged": dict(tile_patches=0), "staged_idx32": dict(tile_patches=0, compact_idx=False),
            "gather": dict(tile_patches=0, force_gather=True),
            "tiled": dict(tile_patches=14 * 128 // ppx), "ring": dict(tile_patches=12 * 128 // ppx, force_ring=True)}[kernel]


def _same_bits(a, b):
    """bitwise equality of two float32 arrays (array_equal would accept -0.0 for +0.0)"""
    return np.array_equal(np.ascontiguousarray(a).view(np.int32), np.ascontiguousarray(b).view(np.int32))


def _plan(kind, sem="1.26", res=238):
    from fluid_llm_b200.airfoil_ds import crop_airfoil_mesh
    from fluid_llm_b200.mesh_utils import MeshPlan
    tr = trajectory(kind)
    pos, faces = tr["mesh_pos"], tr["cells"]
    if kind == "airfoil":
        _, pos, faces = crop_airfoil_mesh(pos, faces)
    return MeshPlan(pos, faces, res, sem), pos, faces


@pytest.mark.parametrize("kind", ["cylinder", "airfoil", "eagle"])
@pytest.mark.parametrize("sem", ["1.26", "2.x"])
def test_locate_tri_ids_bit_exact(kind, sem):
    plan, pos, faces = _plan(kind, sem)
    triang, tri_o, gx, gy = P.get_mesh_interpolation(pos, faces, 238, sem)
    assert plan.tri_index.dtype == np.int32 and plan.tri_index.shape == tri_o.shape
    assert np.array_equal(plan.grid_x, gx) and np.array_equal(plan.grid_y, gy)
    assert np.array_equal(plan.tri_index, tri_o)
    assert (tri_o == -1).sum() > 0 or kind == "airfoil"


def test_locate_tie_breaks_hand_built():
    from fluid_llm_b200.mesh_utils import MeshPlan
    pos, tris = tie_mesh()
    # resolution 9 puts grid points on every lattice vertex and edge midpoint of the 4x4 lattice
    for res in (9, 5, 17, 33):
        plan = MeshPlan(pos, tris, res, "1.26")
        triang, tri_o, gx, gy = P.get_mesh_interpolation(pos, tris, res, "1.26")
        assert np.array_equal(plan.tri_index, tri_o), res
        assert np.array_equal(plan.tri_index, mpl_tri.rule_find_many(triang, gx, gy, bucketed=False))


def test_locate_weights():
    plan, pos, faces = _plan("cylinder")
    triang, tri_o, gx, gy = P.get_mesh_interpolation(pos, faces, 238, "1.26")
    idx = plan.cell_idx_d.cpu().numpy().reshape(plan.nx, plan.ny, 4)
    w = plan.cell_w_d.cpu().numpy().reshape(plan.nx, plan.ny, 2)
    ct = triang.corrected_triangles
    inside = tri_o >= 0
    assert np.array_equal(idx[..., 3], tri_o)
    assert np.array_equal(idx[inside][:, :3], ct[tri_o[inside]])
    x, y = pos[:, 0].astype(np.float64), pos[:, 1].astype(np.float64)
    v = idx[inside][:, :3]
    w1, w2 = w[inside][:, 0], w[inside][:, 1]
    w0 = 1 - w1 - w2
    rx = w0 * x[v[:, 0]] + w1 * x[v[:, 1]] + w2 * x[v[:, 2]]
    ry = w0 * y[v[:, 0]] + w1 * y[v[:, 1]] + w2 * y[v[:, 2]]
    assert np.abs(rx - gx[inside]).max() < 1e-12 and np.abs(ry - gy[inside]).max() < 1e-12
    assert w0.min() > -1e-12 and w1.min() > -1e-12 and w2.min() > -1e-12    # partition of unity, inside


def test_locate_rejects_bad_indices():
    from fluid_llm_b200.mesh_utils import MeshPlan
    pos, tris = tie_mesh()
    bad = tris.copy()
    bad[3, 1] = 99
    with pytest.raises(ValueError):
        MeshPlan(pos, bad, 9)
    with pytest.raises(ValueError):
        MeshPlan(pos, tris[:, :2], 9)


@pytest.mark.parametrize("kind", ["cylinder", "airfoil", "eagle"])
def test_to_grid_matches_oracle(kind):
    from fluid_llm_b200.mesh_utils import get_mesh_interpolation, to_grid
    tr = trajectory(kind)
    plan, pos, faces = _plan(kind)
    nmask = slice(None)
    if kind == "airfoil":
        from fluid_llm_b200.airfoil_ds import crop_airfoil_mesh
        nmask, _, _ = crop_airfoil_mesh(tr["mesh_pos"], tr["cells"])
    triang_o, tri_o, gx, gy = P.get_mesh_interpolation(pos, faces, 238, "1.26")
    triang, tri_index, grid_x, grid_y = get_mesh_interpolation(pos, faces, 238, "1.26")
    assert np.array_equal(tri_index, tri_o)
    for ch, val in enumerate([tr["velocity"][3][nmask][:, 0], tr["velocity"][3][nmask][:, 1], tr["pressure"][3][nmask][:, 0]]):
        d, m = to_grid(val, grid_x, grid_y, triang, tri_index)
        d_o, m_o = P.to_grid(val, gx, gy, triang_o, tri_o)
        assert d.dtype == np.float32 and m.dtype == bool
        assert np.array_equal(m, m_o)
        assert np.array_equal(d, d_o), f"{kind} ch{ch}: {(d != d_o).sum()} of {d.size} values differ"
    with pytest.raises(ValueError):
        to_grid(val[:-1], grid_x, grid_y, triang, tri_index)
    with pytest.raises(ValueError):
        to_grid(val, grid_x, grid_y, triang, tri_index[:-1])


def test_to_grid_nonfinite_values_are_masked():
    from fluid_llm_b200.mesh_utils import get_mesh_interpolation, to_grid
    tr = trajectory("cylinder")
    pos, faces = tr["mesh_pos"], tr["cells"]
    triang_o, tri_o, gx, gy = P.get_mesh_interpolation(pos, faces, 238)
    triang, tri_index, grid_x, grid_y = get_mesh_interpolation(pos, faces, 238)
    val = tr["pressure"][0][:, 0].copy()
    val[100] = np.nan
    val[200] = np.inf
    d, m = to_grid(val, grid_x, grid_y, triang, tri_index)
    d_o, m_o = P.to_grid(val, gx, gy, triang_o, tri_o)
    assert np.array_equal(m, m_o) and np.array_equal(d, d_o)
    assert m.sum() > (tri_o == -1).sum()


@pytest.mark.parametrize("kind", ["cylinder", "airfoil", "eagle"])
@pytest.mark.parametrize("normalize", [True, False])
@pytest.mark.parametrize("kernel", KERNELS)
def test_interp_patchify_matches_oracle(kind, normalize, kernel):
    """states / mask of the fused kernels == the oracle's unfold+normalise pipeline, bit for bit, signed zeros included
    (north_star tolerance: 1e-6 relative; achieved: exact on these inputs)."""
    from fluid_llm_b200.field_path import AIRFOIL, CYLINDER, DeviceTrajectory, interp_patchify
    from fluid_llm_b200.airfoil_ds import crop_airfoil_mesh
    tr = trajectory(kind)
    plan, pos, faces = _plan(kind)
    vel, prs = tr["velocity"], tr["pressure"]
    if kind == "airfoil":
        nmask, _, _ = crop_airfoil_mesh(tr["mesh_pos"], tr["cells"])
        vel, prs = vel[:, nmask], prs[:, nmask]
    pers = AIRFOIL if kind == "airfoil" else CYLINDER
    states, mask, tab = interp_patchify(DeviceTrajectory(vel, prs, plan), 1, 3, 2, PATCH, pers, normalize=normalize,
                                        **_kernel_args(kernel))
    (_, _, _, _, _), extra = oracle_ds_get(kind, 1, 3, 2, normalize=normalize)
    assert (tab.n_bx, tab.n_by) == (extra["N_x_patch"], extra["N_y_patch"])
    assert np.array_equal(mask.cpu().numpy().astype(bool), extra["masks"].astype(bool))
    s, so = states.cpu().numpy(), extra["states"]
    assert s.shape == so.shape and s.dtype == so.dtype
    np.testing.assert_allclose(s, so, rtol=1e-6, atol=0)
    assert np.array_equal(s, so), f"{(s != so).sum()} of {s.size} values differ in the last bit"
    assert _same_bits(s, so), "a signed zero differs from the reference's +0.0"


@pytest.mark.parametrize("kernel", KERNELS)
def test_interp_patchify_batch_of_meshes(kernel):
    """One launch over several trajectories with different meshes, different start frames."""
    import fluid_llm_b200
    from fluid_llm_b200.field_path import CYLINDER, DeviceTrajectory, TrajBatch
    from fluid_llm_b200.mesh_utils import MeshPlan
    trajs, tabs, t0s, want = [], [], [1, 0, 3], []
    for seed, t0 in zip((0, 1, 2), t0s):
        tr = trajectory("cylinder", 8, seed, 10 + seed)
        plan = MeshPlan(tr["mesh_pos"], tr["cells"], 238)
        trajs.append(DeviceTrajectory(tr["velocity"], tr["pressure"], plan))
        tabs.append(plan.patch_table(PATCH))
        want.append(oracle_ds_get("cylinder", t0, 4, 1, mesh_seed=seed, field_seed=10 + seed)[1])
    ka = _kernel_args(kernel)
    batch = TrajBatch(trajs, tabs, t0s, 1, 4, tile_patches=ka["tile_patches"], compact_idx=ka.get("compact_idx"))
    assert (batch.tile_plans is not None) == (kernel in ("tiled", "ring"))
    states, mask = batch.run(CYLINDER, force_gather=ka.get("force_gather", False), force_ring=ka.get("force_ring", False))
    assert fluid_llm_b200.load().fl_last_interp_kernel().decode() == "k_interp_patchify_" + kernel.split("_")[0]
    assert [d.idx_slot_format for d in batch.host_desc] == [1 if kernel in ("staged", "gather") else 0] * 3   # (gather ignores the slot tables)
    torch.cuda.synchronize()
    for i, ex in enumerate(want):
        assert np.array_equal(states[i].cpu().numpy(), ex["states"])
        assert np.array_equal(mask[i].cpu().numpy().astype(bool), ex["masks"].astype(bool))
    with pytest.raises(ValueError):
        TrajBatch(trajs, tabs, [0, 0, 7], 1, 4)      # runs past the end of the trajectory


def test_host_pipeline_matches_oracle():
    """Host buffers in / host buffers out through the three-stream pipeline: five trajectories with different meshes
    over two device slots, run twice (slot reuse across calls)."""
    from fluid_llm_b200.field_path import CYLINDER, HostPipeline
    from fluid_llm_b200.mesh_utils import MeshPlan
    seeds = (0, 1, 2, 0, 2)
    plans, tabs, hv, hp, want = [], [], [], [], []
    for k, seed in enumerate(seeds):
        tr = trajectory("cylinder", 8, seed, 20 + k)
        plan = MeshPlan(tr["mesh_pos"], tr["cells"], 238)
        plans.append(plan)
        tabs.append(plan.patch_table(PATCH))
        hv.append(torch.from_numpy(np.ascontiguousarray(tr["velocity"], dtype=np.float32)).pin_memory())
        hp.append(torch.from_numpy(np.ascontiguousarray(tr["pressure"], dtype=np.float32).reshape(8, -1, 1)).pin_memory())
        want.append(oracle_ds_get("cylinder", 1, 3, 2, mesh_seed=seed, field_seed=20 + k)[1])
    pipe = HostPipeline(plans, tabs, CYLINDER, n_steps=8, t0=1, interval=2, n_frames=3, depth=2)
    L = tabs[0].n_patches
    for rep in range(2):
        hs = [torch.full((3, L, 3, 16, 16), float("nan")).pin_memory() for _ in seeds]
        hm = [torch.full((3, L, 16, 16), 7, dtype=torch.uint8).pin_memory() for _ in seeds]
        pipe.run(hv, hp, hs, hm)
        pipe.wait()
        for i, ex in enumerate(want):
            assert np.array_equal(hs[i].numpy(), ex["states"]), (rep, i)
            assert np.array_equal(hm[i].numpy().astype(bool), ex["masks"].astype(bool)), (rep, i)
    with pytest.raises(ValueError):
        pipe.run(hv[:2], hp[:2], hs[:2], hm[:2])
    # a consumer on the GPU: called per trajectory with the device slot, on the compute stream
    got = {}

    def consume(i, states, mask):
        got[i] = (states.clone(), mask.clone())

    pipe.run(hv, hp, on_device=consume)
    torch.cuda.synchronize()
    for i, ex in enumerate(want):
        assert np.array_equal(got[i][0].cpu().numpy(), ex["states"]) and np.array_equal(got[i][1].cpu().numpy().astype(bool), ex["masks"].astype(bool))


def test_custom_mean_std():
    from fluid_llm_b200.field_path import CYLINDER, DeviceTrajectory, interp_patchify
    tr = trajectory("eagle")
    plan, _, _ = _plan("eagle")
    means, stds = (0.1, -0.2, 0.3), (1.5, 1.9, 6.3)
    states, _, _ = interp_patchify(DeviceTrajectory(tr["velocity"], tr["pressure"], plan), 0, 2, 1, PATCH, CYLINDER,
                                   means=means, stds=stds)
    _, extra = oracle_ds_get("eagle", 0, 2, 1, means=means, stds=stds)
    assert np.array_equal(states.cpu().numpy(), extra["states"])


@pytest.mark.parametrize("kind", ["cylinder", "airfoil"])
def test_nonfinite_node_values_take_the_checked_path(kind):
    """NaN / inf node values: the staged kernel's scan sends the item down the checked path; mask and
    zeroing follow mesh_utils.py:86-89 per channel, only the pressure mask is kept."""
    from fluid_llm_b200.airfoil_ds import crop_airfoil_mesh
    from fluid_llm_b200.field_path import AIRFOIL, CYLINDER, DeviceTrajectory, interp_patchify
    import copy
    tr = copy.deepcopy(trajectory(kind))
    tr["pressure"][2, 50:60, 0] = np.nan
    tr["velocity"][3, 70, 0] = np.inf
    tr["velocity"][1, 80, 1] = -np.inf
    plan, pos, faces = _plan(kind)
    vel, prs = tr["velocity"], tr["pressure"]
    if kind == "airfoil":
        nmask, _, _ = crop_airfoil_mesh(tr["mesh_pos"], tr["cells"])
        vel, prs = vel[:, nmask], prs[:, nmask]
    pers = AIRFOIL if kind == "airfoil" else CYLINDER
    with np.errstate(all="ignore"):
        _, extra = P.ds_get(tr, 0, 5, 1, 238, PATCH, personality_of(kind), return_all=True)
    for kernel in KERNELS:
        states, mask, _ = interp_patchify(DeviceTrajectory(vel, prs, plan), 0, 5, 1, PATCH, pers, **_kernel_args(kernel))
        assert np.array_equal(mask.cpu().numpy().astype(bool), extra["masks"].astype(bool)), kernel
        assert _same_bits(states.cpu().numpy(), extra["states"]), kernel


@pytest.mark.parametrize("kernel", ["staged", "tiled", "ring"])
def test_long_sequence_many_items(kernel):
    """More frames than one work item holds (several frame groups per unit in the tiled kernel), odd node count (padded
    frame pitch), interval 3."""
    from fluid_llm_b200.field_path import CYLINDER, DeviceTrajectory, interp_patchify
    from fluid_llm_b200.mesh_utils import MeshPlan
    tr = trajectory("cylinder", 64, 3, 7)
    plan = MeshPlan(tr["mesh_pos"], tr["cells"], 238)
    states, mask, _ = interp_patchify(DeviceTrajectory(tr["velocity"], tr["pressure"], plan), 2, 20, 3, PATCH, CYLINDER,
                                      **_kernel_args(kernel))
    _, extra = oracle_ds_get("cylinder", 2, 20, 3, T=64, mesh_seed=3, field_seed=7)
    assert _same_bits(states.cpu().numpy(), extra["states"])
    assert np.array_equal(mask.cpu().numpy().astype(bool), extra["masks"].astype(bool))


def test_tiled_kernel_many_frame_groups_and_no_mask():
    """The tiled kernel over 64 frames (4+ frame groups per unit, runs of groups split into units), with and without the
    mask output, against the staged kernel's bits."""
    from fluid_llm_b200.field_path import CYLINDER, DeviceTrajectory, TrajBatch
    from fluid_llm_b200.mesh_utils import MeshPlan
    tr = trajectory("cylinder", 64, 3, 7)
    plan = MeshPlan(tr["mesh_pos"], tr["cells"], 238)
    dt = DeviceTrajectory(tr["velocity"], tr["pressure"], plan)
    tab = plan.patch_table(PATCH)
    ref = TrajBatch([dt], [tab], [0], 1, 64, tile_patches=0)
    rs, rm = ref.run(CYLINDER)
    for tp in (7, 6, 5, 3):
        for want_mask in (True, False):
            b = TrajBatch([dt], [tab], [0], 1, 64, want_mask=want_mask, tile_patches=tp)
            s, m = b.run(CYLINDER)
            assert torch.equal(s.view(torch.int32), rs.view(torch.int32)), (tp, want_mask)
            assert m is None or torch.equal(m, rm)
    _, extra = oracle_ds_get("cylinder", 0, 64, 1, T=64, mesh_seed=3, field_seed=7)
    assert _same_bits(rs[0].cpu().numpy(), extra["states"])


def test_big_mesh_tiled_and_gather_paths_match_oracle():
    """BASELINE config 5 in miniature: a structured ~80k-triangle mesh on a 512-point grid.  Whole frames of the node
    arrays do not fit shared memory, so the default is the tiled kernel; the gather-from-global kernel is checked too."""
    from fluid_llm_b200 import synth
    from fluid_llm_b200.field_path import CYLINDER, DeviceTrajectory, interp_patchify
    from fluid_llm_b200.mesh_utils import MeshPlan
    pos, cells = synth.make_mesh("big", 0, nx=281, ny=141)
    vel, prs = synth.make_fields("big", pos, 3, 1)
    plan = MeshPlan(pos, cells, 512)
    triang, tri_o, gx, gy = P.get_mesh_interpolation(pos, cells, 512)
    assert (plan.nx, plan.ny) == tri_o.shape == (512, 256)
    assert np.array_equal(plan.tri_index, tri_o)
    assert np.array_equal(tri_o, mpl_tri.rule_find_many(triang, gx, gy, bucketed=True))
    from fluid_llm_b200.field_path import TrajBatch
    dt = DeviceTrajectory(vel, prs, plan)
    assert TrajBatch([dt], [plan.patch_table(PATCH)], [0], 1, 3).tile_plans is not None      # the default picks tiles here
    tr = {"mesh_pos": pos, "cells": cells, "velocity": vel, "pressure": prs}
    _, extra = P.ds_get(tr, 0, 3, 1, 512, PATCH, "cylinder", return_all=True)
    for kw in (dict(), dict(tile_patches=7), dict(tile_patches=0, force_gather=True)):
        states, mask, tab = interp_patchify(dt, 0, 3, 1, PATCH, CYLINDER, **kw)
        assert (tab.n_bx, tab.n_by) == (32, 16)
        assert np.array_equal(mask.cpu().numpy().astype(bool), extra["masks"].astype(bool)), kw
        assert _same_bits(states.cpu().numpy(), extra["states"]), kw


@pytest.mark.parametrize("patch", [(8, 8), (16, 8), (8, 16), (4, 8), (5, 5), (10, 7), (32, 32), (24, 16), (20, 20), (64, 8)])
def test_other_patch_sizes(patch):
    """Patch sizes other than the reference's 16 x 16: multiples of 128 pixels run the staged kernel ((16,8), (8,16), (32,32),
    (24,16), (64,8)), everything else the gather kernel -- any size the reference's F.unfold takes."""
    from fluid_llm_b200.field_path import CYLINDER, DeviceTrajectory, interp_patchify
    tr = trajectory("cylinder")
    plan, _, _ = _plan("cylinder")
    _, extra = P.ds_get(tr, 0, 3, 1, 238, patch, "cylinder", return_all=True)
    variants = [dict()] + ([dict(tile_patches=14), dict(tile_patches=9)] if patch[0] * patch[1] == 128 else [])
    for kw in variants:        # 128-pixel patches also run the tiled kernel (one warp per patch)
        states, mask, tab = interp_patchify(DeviceTrajectory(tr["velocity"], tr["pressure"], plan), 0, 3, 1, patch, CYLINDER, **kw)
        assert (tab.n_bx, tab.n_by) == (extra["N_x_patch"], extra["N_y_patch"])
        assert np.array_equal(states.cpu().numpy(), extra["states"]), kw
        assert np.array_equal(mask.cpu().numpy().astype(bool), extra["masks"].astype(bool)), kw


def test_compact_table_records_are_what_the_header_says():
    """fl_pack_idx16 (include/fluidgrid.h): {v0 | v1 << 16, v2 | outside << 16} from the slot form of the table; refused when the
    slots do not fit 16 bits."""
    import ctypes
    from fluid_llm_b200._lib import load, ptr
    plan, _, _ = _plan("airfoil")
    tab = plan.patch_table(PATCH, 1, True)
    packed = tab.idx_slot16(plan.n_padded).cpu().numpy().view(np.uint32)
    src = tab.idx_slot.cpu().numpy()
    inside = src[:, 3] >= 0
    assert inside.any() and (~inside).any() and src[inside, :3].max() < plan.n_padded
    want = np.zeros((len(src), 2), dtype=np.uint32)
    want[inside, 0] = src[inside, 0].astype(np.uint32) | (src[inside, 1].astype(np.uint32) << 16)
    want[inside, 1] = src[inside, 2].astype(np.uint32)
    want[~inside, 1] = 1 << 16
    assert np.array_equal(packed, want)
    out = torch.empty((len(src), 2), dtype=torch.int32, device="cuda")
    assert load().fl_pack_idx16(ptr(tab.idx_slot), len(src), 65537, ptr(out), None) == -1
    assert b"16 bits" in load().fl_last_error()


def test_other_patch_sizes_kernel_choice():
    from fluid_llm_b200._lib import load
    from fluid_llm_b200.field_path import CYLINDER, DeviceTrajectory, interp_patchify
    tr = trajectory("cylinder")
    plan, _, _ = _plan("cylinder")
    for patch, kernel in (((32, 32), b"k_interp_patchify_staged"), ((20, 20), b"k_interp_patchify_gather"), ((5, 5), b"k_interp_patchify_gather")):
        interp_patchify(DeviceTrajectory(tr["velocity"], tr["pressure"], plan), 0, 2, 1, patch, CYLINDER)
        assert load().fl_last_interp_kernel() == kernel
    with pytest.raises(ValueError):
        interp_patchify(DeviceTrajectory(tr["velocity"], tr["pressure"], plan), 0, 2, 1, (0, 5), CYLINDER)


def test_c_abi_called_directly_as_integration_md_shows():
    """fl_interp_patchify with host-side descriptors, bound with plain ctypes exactly like INTEGRATION.md section 5
    (unpadded node arrays straight from the pickle layout -> the gather kernel)."""
    import ctypes
    from fluid_llm_b200._lib import LIB_PATH
    tr = trajectory("cylinder")
    plan, _, _ = _plan("cylinder")
    tab = plan.patch_table(PATCH)
    lib = ctypes.CDLL(LIB_PATH)

    class FlTraj(ctypes.Structure):
        _fields_ = [("d_velocity", ctypes.c_void_p), ("d_pressure", ctypes.c_void_p),
                    ("d_idx", ctypes.c_void_p), ("d_w", ctypes.c_void_p),
                    ("d_idx_slot", ctypes.c_void_p), ("d_node_slot", ctypes.c_void_p),
                    ("d_states", ctypes.c_void_p), ("d_mask", ctypes.c_void_p),
                    ("n_nodes", ctypes.c_int32), ("t0", ctypes.c_int32), ("interval", ctypes.c_int32), ("n_frames", ctypes.c_int32),
                    ("vel_stride", ctypes.c_int32), ("prs_stride", ctypes.c_int32),
                    ("d_idx_tile", ctypes.c_void_p), ("d_tile_nodes", ctypes.c_void_p), ("d_tile_desc", ctypes.c_void_p),
                    ("d_tile_patches", ctypes.c_void_p), ("d_tile_quads", ctypes.c_void_p), ("d_tile_qslots", ctypes.c_void_p),
                    ("n_tiles", ctypes.c_int32), ("max_tile_nodes", ctypes.c_int32),
                    ("max_tile_patches", ctypes.c_int32), ("idx_slot_format", ctypes.c_int32)]

    lib.fl_interp_patchify.restype = ctypes.c_int
    lib.fl_interp_patchify.argtypes = [ctypes.POINTER(FlTraj), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.c_uint, ctypes.c_void_p]
    lib.fl_last_error.restype = ctypes.c_char_p
    vel = torch.from_numpy(tr["velocity"]).cuda()           # [T, N, 2], unpadded: N = 1855 is odd
    prs = torch.from_numpy(tr["pressure"]).cuda()           # [T, N, 1]
    N, seq_len, step, interval = vel.shape[1], 3, 1, 2
    states = torch.empty((seq_len, tab.n_patches, 3, 16, 16), dtype=torch.float32, device="cuda")
    mask = torch.empty((seq_len, tab.n_patches, 16, 16), dtype=torch.uint8, device="cuda")
    traj = FlTraj(vel.data_ptr(), prs.data_ptr(), tab.idx.data_ptr(), tab.w.data_ptr(), None, None,
                  states.data_ptr(), mask.data_ptr(), N, step, interval, seq_len, vel.stride(0), prs.stride(0),
                  None, None, None, None, None, None, 0, 0, 0, 0)      # no tile plan: the library picks the staged / gather kernel
    mean = (ctypes.c_float * 3)(0.823, 0.0005865, 0.04763)
    std = (ctypes.c_float * 3)(0.275, 0.275, 0.275)
    rc = lib.fl_interp_patchify(ctypes.byref(traj), 1, tab.n_patches, 16, 16, mean, std, 0,
                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.fl_last_error().decode()
    torch.cuda.synchronize()
    _, extra = oracle_ds_get("cylinder", step, seq_len, interval)
    assert np.array_equal(states.cpu().numpy(), extra["states"])
    assert np.array_equal(mask.cpu().numpy().astype(bool), extra["masks"].astype(bool))
    # argument errors come back as negative codes with a message
    traj.vel_stride = N          # too small for [N, 2]
    assert lib.fl_interp_patchify(ctypes.byref(traj), 1, tab.n_patches, 16, 16, mean, std, 0, None) == -1
    assert b"stride" in lib.fl_last_error()


@pytest.mark.parametrize("kind,kernel", [("cylinder", "staged"), ("airfoil", "staged"), ("eagle", "staged"), ("cylinder", "gather"),
                                         ("cylinder", "tiled"), ("airfoil", "tiled"), ("eagle", "tiled"), ("airfoil", "ring")])
def test_interp_patchify_writes_stay_inside_the_output_buffers(kind, kernel):
    """Outputs placed inside guarded buffers (TrajBatch(out=...)): the staged and the gather kernel write every element of
    [n_traj, n_frames, L, ...] and nothing around it."""
    from fluid_llm_b200.field_path import AIRFOIL, CYLINDER, DeviceTrajectory, TrajBatch
    pers = AIRFOIL if kind == "airfoil" else CYLINDER
    plan, _, _ = _plan(kind)
    tr = trajectory(kind)
    vel, prs = tr["velocity"], tr["pressure"]
    if kind == "airfoil":
        from fluid_llm_b200.airfoil_ds import crop_airfoil_mesh
        nmask, _, _ = crop_airfoil_mesh(tr["mesh_pos"], tr["cells"])
        vel, prs = vel[:, nmask], prs[:, nmask]
    tab = plan.patch_table(PATCH, pers.crop_patches, pers.flip_y)
    trajs = [DeviceTrajectory(vel, prs, plan) for _ in range(2)]
    n_traj, T, L, G = 2, 7, tab.n_patches, 4096
    n = n_traj * T * L * 256
    sbuf = torch.full((G + 3 * n + G,), 777.0, dtype=torch.float32, device="cuda")
    mbuf = torch.full((G + n + G,), 77, dtype=torch.uint8, device="cuda")
    states = sbuf[G:G + 3 * n].view(n_traj, T, L, 3, 16, 16)
    mask = mbuf[G:G + n].view(n_traj, T, L, 16, 16)
    ka = _kernel_args(kernel)
    batch = TrajBatch(trajs, [tab, tab], [0, 1], 1, T, out=(states, mask), tile_patches=ka["tile_patches"], compact_idx=ka.get("compact_idx"))
    batch.run(pers, force_gather=ka.get("force_gather", False), force_ring=ka.get("force_ring", False))
    torch.cuda.synchronize()
    assert bool((sbuf[:G] == 777.0).all()) and bool((sbuf[-G:] == 777.0).all())
    assert bool((mbuf[:G] == 77).all()) and bool((mbuf[-G:] == 77).all())
    assert not bool((states == 777.0).any()) and not bool((mask == 77).any())
    _, extra = oracle_ds_get(kind, 1, T, 1)
    assert np.array_equal(states[1].cpu().numpy(), extra["states"])


def test_c5_full_size_1m_triangles():
    """BASELINE config 5 at FULL size: ~1M triangles / ~0.5M nodes on a 2048 x 1024 grid (8192 patches), the tiled kernel.
    Triangle ids against the stated rule evaluated on all 2M cells and against the trapezoid map (the restated matplotlib
    trifinder) on a strided subset plus the boundary rows and columns; states and mask of two frames against the oracle's
    plane interpolation + pad + unfold + normalise."""
    from fluid_llm_b200 import synth
    from fluid_llm_b200.field_path import CYLINDER, DeviceTrajectory, TrajBatch
    from fluid_llm_b200.mesh_utils import MeshPlan
    pos, cells = synth.make_mesh("big", 0)
    assert len(cells) == 1_000_000 and len(pos) > 500_000
    plan = MeshPlan(pos, cells, 2048)
    assert (plan.nx, plan.ny) == (2048, 1024)
    triang = mpl_tri.Triangulation(pos[:, 0], pos[:, 1], cells)
    gx, gy = plan.grid_x, plan.grid_y
    assert np.array_equal(gx, P.grid_pos(pos[:, 0].min(), pos[:, 0].max(), pos[:, 1].min(), pos[:, 1].max(), 2048)[0])
    tri_rule = mpl_tri.rule_find_many(triang, gx, gy, bucketed=True)
    assert np.array_equal(plan.tri_index, tri_rule)                       # all 2 097 152 cells
    finder = triang.get_trifinder()                                       # trapezoid map over 1M triangles
    sub = np.zeros(gx.shape, dtype=bool)
    sub[::8, ::8] = True
    sub[[0, -1], :] = True
    sub[:, [0, -1]] = True                                               # the boundary cells sit exactly on boundary edges
    assert np.array_equal(finder(gx[sub], gy[sub]), plan.tri_index[sub])
    vel, prs = synth.make_fields("big", pos, 2, 1)
    dt = DeviceTrajectory(vel, prs, plan)
    tab = plan.patch_table(PATCH)
    assert (tab.n_bx, tab.n_by) == (128, 64)
    batch = TrajBatch([dt], [tab], [0], 1, 2)
    assert batch.tile_plans is not None
    states, mask = batch.run(CYLINDER)
    import fluid_llm_b200
    assert fluid_llm_b200.load().fl_last_interp_kernel() == b"k_interp_patchify_tiled"
    seq = []
    for t in range(2):
        state, m = P.get_step(triang, tri_rule, gx, gy, vel, prs, t, PATCH)
        seq.append(np.concatenate([state, m[None].astype(state.dtype)], axis=0))
    patches = P.unfold_patches(np.stack(seq).astype(np.float32), PATCH)
    want = np.ascontiguousarray(patches[:, :-1].transpose(0, 4, 1, 2, 3))
    want_mask = np.ascontiguousarray(patches[:, -1].transpose(0, 3, 1, 2))
    want = P.normalize(want, want_mask, "cylinder")
    assert np.array_equal(mask[0].cpu().numpy().astype(bool), want_mask.astype(bool))
    got = states[0].cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=0)                  # north_star's tolerance, every value
    # the reference rounds the fp64 plane a*x + b*y + c, the kernel the fp64 barycentric sum: the two fp64 values differ in
    # their last bits, so a handful of the 12.6M fp32 roundings land on the other side (1 ulp); everything else is bit-equal
    diff = got.view(np.int32) != want.view(np.int32)
    assert diff.sum() <= 12, f"{diff.sum()} of {got.size} values differ in their bits"
    assert np.all(np.abs(got.view(np.int32)[diff].astype(np.int64) - want.view(np.int32)[diff].astype(np.int64)) <= 1)
    print(f"C5 full size: {int(diff.sum())} of {got.size} values differ by one ulp")
    # the gather-from-global kernel agrees on the same table
    s2, m2 = TrajBatch([dt], [tab], [0], 1, 2, tile_patches=0).run(CYLINDER, force_gather=True)
    assert torch.equal(s2.view(torch.int32), states.view(torch.int32)) and torch.equal(m2, mask)


def test_locate_agrees_with_the_exact_rational_rule():
    """CUDA triangle ids == the stated rule in exact rational arithmetic (oracle/exact_locator.py) on the tie mesh, where
    every grid point sits on a vertex, an edge or in a hole."""
    from fluid_llm_b200.mesh_utils import MeshPlan
    from oracle.exact_locator import exact_find_many
    pos, tris = tie_mesh()
    for res in (9, 17):
        plan = MeshPlan(pos, tris, res, "1.26")
        assert np.array_equal(plan.tri_index, exact_find_many(pos, tris, plan.grid_x, plan.grid_y)), res


def test_zero_area_triangles_are_refused_or_follow_the_rule():
    """A triangle of three colinear nodes makes matplotlib's trapezoid map invalid (overlapping edges) and its plane fit take
    the pseudo-inverse branch.  MeshPlan refuses such a mesh by default; with allow_degenerate=True the ids follow the
    stated rule (brute-force fp64 and exact rational evaluations agree) and a cell located in the degenerate triangle
    gets its first vertex's value -- the documented divergence from calculate_plane_coefficients."""
    from fluid_llm_b200.mesh_utils import MeshPlan, to_grid
    from oracle.exact_locator import exact_find_many
    pos, tris = tie_mesh()
    idx = np.arange(25).reshape(5, 5)
    flat = np.array([[idx[0, 0], idx[2, 0], idx[1, 0]]], dtype=np.int32)          # three nodes on the bottom boundary y = 0
    bad = np.concatenate([flat, tris]).astype(np.int32)                           # lowest index: wins every tie it takes part in
    with pytest.raises(ValueError, match="zero area"):
        MeshPlan(pos, bad, 9)
    plan = MeshPlan(pos, bad, 9, "1.26", allow_degenerate=True)
    assert plan.n_degenerate == 1
    triang = mpl_tri.Triangulation(pos[:, 0], pos[:, 1], bad)
    rule = mpl_tri.rule_find_many(triang, plan.grid_x, plan.grid_y, bucketed=False)
    assert np.array_equal(plan.tri_index, rule)
    assert np.array_equal(plan.tri_index, exact_find_many(pos, bad, plan.grid_x, plan.grid_y))
    z = np.linspace(1.0, 2.0, 25).astype(np.float32)
    d, m = to_grid(z, plan.grid_x, plan.grid_y, plan, plan.tri_index)
    d_o, m_o = P.to_grid(z, plan.grid_x, plan.grid_y, triang, rule)
    in_flat = plan.tri_index == 0
    assert np.array_equal(m, m_o) and np.array_equal(d[~in_flat], d_o[~in_flat])
    assert np.all(d[in_flat] == z[flat[0, 0]])                                     # first vertex's value


def test_to_grid_refuses_a_foreign_tri_index():
    from fluid_llm_b200.mesh_utils import get_mesh_interpolation, to_grid
    tr = trajectory("cylinder")
    triang, tri_index, gx, gy = get_mesh_interpolation(tr["mesh_pos"], tr["cells"], 238)
    val = tr["pressure"][0][:, 0]
    to_grid(val, gx, gy, triang, tri_index.copy())                                  # an equal copy is fine
    other = tri_index.copy()
    other[10, 10] = (other[10, 10] + 1) % triang.n_cells
    with pytest.raises(ValueError, match="tri_index differs"):
        to_grid(val, gx, gy, triang, other)


@pytest.mark.parametrize("kernel", KERNELS)
def test_ragged_batch_through_the_c_abi(kernel):
    """One call over trajectories with DIFFERENT frame counts, start frames and intervals (FlTraj::n_frames / t0 / interval are
    per trajectory): every kernel writes exactly the frames each trajectory asks for -- bit-identical to that trajectory
    computed alone -- and nothing behind them."""
    from fluid_llm_b200.field_path import CYLINDER, DeviceTrajectory, TrajBatch
    from fluid_llm_b200.mesh_utils import MeshPlan
    spec = [(7, 0, 1), (3, 2, 3), (1, 5, 1), (5, 1, 2)]          # (n_frames, t0, interval)
    trajs, tabs = [], []
    for seed in range(len(spec)):
        tr = trajectory("cylinder", 12, seed, 20 + seed)
        plan = MeshPlan(tr["mesh_pos"], tr["cells"], 238)
        trajs.append(DeviceTrajectory(tr["velocity"], tr["pressure"], plan))
        tabs.append(plan.patch_table(PATCH))
    ka = _kernel_args(kernel)
    run_kw = dict(force_gather=ka.get("force_gather", False), force_ring=ka.get("force_ring", False))
    n_max = max(s[0] for s in spec)
    batch = TrajBatch(trajs, tabs, [0] * len(spec), 1, n_max, tile_patches=ka["tile_patches"], compact_idx=ka.get("compact_idx"))
    for i, (nf, t0, iv) in enumerate(spec):
        batch.host_desc[i].n_frames, batch.host_desc[i].t0, batch.host_desc[i].interval = nf, t0, iv
    batch.desc = torch.from_numpy(np.frombuffer(bytes(batch.host_desc), dtype=np.uint8).copy()).cuda()
    batch.states.fill_(777.0)
    batch.mask.fill_(77)
    states, mask = batch.run(CYLINDER, **run_kw)
    torch.cuda.synchronize()
    for i, (nf, t0, iv) in enumerate(spec):
        alone = TrajBatch([trajs[i]], [tabs[i]], [t0], iv, nf, tile_patches=0)
        s1, m1 = alone.run(CYLINDER)
        assert torch.equal(states[i, :nf].view(torch.int32), s1[0].view(torch.int32)), (kernel, i)
        assert torch.equal(mask[i, :nf], m1[0]), (kernel, i)
        assert bool((states[i, nf:] == 777.0).all()) and bool((mask[i, nf:] == 77).all()), (kernel, i)
