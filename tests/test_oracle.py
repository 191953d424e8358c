"""CPU: the oracle against known answers, against the frozen reference outputs (tests/golden/) and,
when /root/reference is present (development container only), against the reference itself."""
import copy
import os

import numpy as np
import pytest

from oracle import mpl_tri, ref_import
from oracle import pipeline as P

from helpers import PATCH, tie_mesh, trajectory

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold(name):
    return np.load(os.path.join(GOLDEN, name))


# ---- matplotlib._tri restatement ------------------------------------------------------------
@pytest.mark.parametrize("kind", ["cylinder", "eagle"])
def test_trapezoid_map_equals_stated_rule(kind):
    tr = trajectory(kind)
    triang, tri_index, gx, gy = P.get_mesh_interpolation(tr["mesh_pos"], tr["cells"], 120)
    assert np.array_equal(tri_index, mpl_tri.rule_find_many(triang, gx, gy, bucketed=True))
    sub = (slice(None, None, 7), slice(None, None, 3))
    assert np.array_equal(tri_index[sub], mpl_tri.rule_find_many(triang, gx[sub], gy[sub], bucketed=False))
    # queries exactly on vertices and on edge midpoints
    x, y = triang.x, triang.y
    finder = triang.get_trifinder()
    assert np.array_equal(finder(x, y), mpl_tri.rule_find_many(triang, x, y, bucketed=True))
    ct = triang.corrected_triangles
    mx, my = (x[ct[:, 0]] + x[ct[:, 1]]) / 2, (y[ct[:, 0]] + y[ct[:, 1]]) / 2
    assert np.array_equal(finder(mx, my), mpl_tri.rule_find_many(triang, mx, my, bucketed=True))
    # vertex rule: the lowest-index triangle listing the vertex
    first = np.full(len(x), -1)
    for t in range(len(ct) - 1, -1, -1):
        first[ct[t]] = t
    assert np.array_equal(finder(x, y), first)


def test_tie_mesh_matches_frozen_ids_and_rule():
    g = _gold("tie_mesh.npz")
    pos, tris = tie_mesh()
    assert np.array_equal(pos, g["pos"]) and np.array_equal(tris, g["tris"])
    for res in (5, 9, 17, 33):
        triang, ti, gx, gy = P.get_mesh_interpolation(pos, tris, res)
        assert np.array_equal(ti, g[f"tri_index_{res}"])
        assert np.array_equal(ti, mpl_tri.rule_find_many(triang, gx, gy, bucketed=False))
    # spot checks of the stated rule on the 4x4 lattice (res 9: points on every vertex / edge midpoint)
    triang, ti, gx, gy = P.get_mesh_interpolation(pos, tris, 9)
    assert ti[0, 0] == 0                      # corner vertex -> lowest-index triangle listing it
    assert ti[3, 3] == -1 and ti[5, 5] == -1  # centres of the two removed squares (holes)
    assert (ti[[0, -1], :] >= 0).all() and (ti[:, [0, -1]] >= 0).all()   # boundary edges belong to the mesh


def test_orientation_fix_and_neighbours():
    pos, tris = tie_mesh()
    t = mpl_tri.Triangulation(pos[:, 0], pos[:, 1], tris)
    ct = t.corrected_triangles
    p = pos.astype(np.float64)[ct]
    area2 = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 1, 1] - p[:, 0, 1]) * (p[:, 2, 0] - p[:, 0, 0])
    assert (area2 > 0).all()
    assert np.array_equal(np.sort(ct, axis=1), np.sort(tris, axis=1))
    nb = t.neighbors
    for ti in range(len(ct)):
        for k in range(3):
            if nb[ti, k] >= 0:
                assert ti in nb[nb[ti, k]]


def test_plane_coefficients():
    tr = trajectory("cylinder")
    pos = tr["mesh_pos"].astype(np.float64)
    t = mpl_tri.Triangulation(pos[:, 0], pos[:, 1], tr["cells"])
    z = tr["pressure"][0][:, 0]
    pc = t.calculate_plane_coefficients(z)
    ct = t.corrected_triangles
    for k in range(3):
        v = ct[:, k]
        np.testing.assert_allclose(pc[:, 0] * pos[v, 0] + pc[:, 1] * pos[v, 1] + pc[:, 2], z[v].astype(np.float64), rtol=0, atol=1e-9)
    # colinear triangle -> Moore-Penrose branch (matplotlib _tri.cpp calculate_plane_coefficients)
    t2 = mpl_tri.Triangulation([0.0, 1.0, 2.0, 0.0], [0.0, 1.0, 2.0, 1.0], [[0, 1, 2], [0, 1, 3]])
    pc2 = t2.calculate_plane_coefficients(np.array([1.0, 2.0, 3.0, 5.0]))
    np.testing.assert_allclose(pc2[0], [0.5, 0.5, 1.0], atol=1e-15)
    with pytest.raises(ValueError):
        t2.calculate_plane_coefficients(np.zeros(3))
    with pytest.raises(ValueError):
        mpl_tri.Triangulation([0.0, 1.0, 2.0], [0.0, 1.0, 0.0], [[0, 1, 5]])


# ---- Python-side restatement: known answers -------------------------------------------------
def test_cylinder_known_answers():
    """Weak pins the reference does hold: 60 = 15 x 4 patches (src/models/layers/GNN/decoders.py:52,123),
    9 x 60 = 540 tokens (decoders.py:118), padded pixel = (0 - 0.823) / 0.275 (simple_dataloader.py:205-214)."""
    (inp, nxt, diffs, mask, pos_ids), extra = P.ds_get(trajectory("cylinder", 12), 0, 10, 1, 238, PATCH, "cylinder", return_all=True)
    assert (extra["N_x_patch"], extra["N_y_patch"]) == (15, 4)
    assert inp.shape == (9, 60, 3, 16, 16) and inp.shape[0] * inp.shape[1] == 540
    assert mask.dtype == bool and mask.shape == inp.shape and pos_ids.shape == (9, 60, 3) and pos_ids.dtype == np.int64
    assert inp[0, 0, 0, 0, 0] == np.float32(np.float32(0 - np.float32(0.823)) / np.float32(0.275))
    assert np.isclose(float(inp[0, 0, 0, 0, 0]), -2.99272728)
    assert mask[0, 0, :, 0, 0].all()            # padding is masked
    assert np.array_equal(diffs, nxt - inp)


def test_pos_id_quirk_table():
    ids = P.get_pos_id(3, 15, 4)
    assert ids[0, :6].tolist() == [[0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0], [4, 0, 0], [5, 0, 0]]
    assert ids[0, 15].tolist() == [0, 1, 0] and ids[0, 59].tolist() == [14, 3, 0] and ids[1, 0].tolist() == [0, 0, 1]
    a = np.arange(2 * 60)
    assert np.array_equal(ids.reshape(-1, 3), np.stack([a % 15, (a // 15) % 4, a // 60], axis=1))


def test_grid_pos_numpy_semantics():
    gx, gy = P.grid_pos(np.float32(0), np.float32(1.6), np.float32(0), np.float32(0.41), 238)
    assert gx.shape == (238, 60) and gx.dtype == np.float32
    assert gx[0, 0] == 0 and gx[-1, 0] == np.float32(1.6) and gy[0, -1] == np.float32(0.41)
    gx2, gy2 = P.grid_pos(np.float32(0), np.float32(1.6), np.float32(0), np.float32(0.41), 238, "2.x")
    assert gx2.shape == gx.shape and 0 < (gx2[:, 0] != gx[:, 0]).sum() < 80      # a few 1-ulp differences
    # the "2.x" mode is what the installed NumPy's np.mgrid does with float32 bounds
    mx, my = np.mgrid[np.float32(0):np.float32(1.6):238j, np.float32(0):np.float32(0.41):60j]
    if int(np.__version__.split(".")[0]) >= 2:
        assert np.array_equal(mx.astype(np.float32), gx2) and np.array_equal(my.astype(np.float32), gy2)
    # tall domain: y is the long axis
    assert P.grid_shape(0, 1, 0, 2, 100) == (50, 100)


def test_pad_unfold_fold_roundtrip():
    rng = np.random.default_rng(0)
    img = rng.standard_normal((2, 3, 3, 48, 32)).astype(np.float32)
    pt = P.img_to_patch(img, PATCH)
    assert pt.shape == (2, 3, 6, 3, 16, 16)
    assert np.array_equal(P.patch_to_img(pt, 3, 2), img)
    assert np.array_equal(pt[0, 0, 1 * 2 + 1, 2], img[0, 0, 2, 16:32, 16:32])     # l = bx * n_by + by
    assert P.pad_amounts(238, 60, PATCH) == ((1, 1), (2, 2))
    assert P.num_patches(240, 16, 16) == 15


def test_stats_oracle():
    rng = np.random.default_rng(1)
    v = rng.standard_normal(10000) * 3 + 7
    agg = (0, 0.0, 0.0)
    for part in np.array_split(v, 7):
        agg = P.update_variance_batch(agg, part)
    assert agg[0] == 10000 and np.isclose(agg[1], v.mean()) and np.isclose(P.get_std(agg), v.std())
    a = P.update_variance_batch((0, 0.0, 0.0), v[:3000])
    b = P.update_variance_batch((0, 0.0, 0.0), v[3000:])
    m = P.chan_merge(a, b)
    assert m[0] == 10000 and np.isclose(m[1], v.mean()) and np.isclose(np.sqrt(m[2] / m[0]), v.std())


# ---- frozen reference outputs ---------------------------------------------------------------
@pytest.mark.parametrize("kind", ["cylinder", "airfoil"])
def test_oracle_reproduces_frozen_reference(kind):
    """tests/golden/ref_*.npz were produced by the reference's unmodified datasets (oracle/make_golden.py)
    under NumPy 2.x; the oracle in "2.x" mode must reproduce them bit for bit."""
    g = _gold(f"ref_{kind}.npz")
    tr = {"mesh_pos": g["mesh_pos"], "cells": g["cells"], "velocity": g["velocity"], "pressure": g["pressure"]}
    out, extra = P.ds_get(tr, int(g["step_num"]), int(g["seq_len"]), int(g["seq_interval"]), 238, PATCH,
                          "airfoil" if kind == "airfoil" else "cylinder", numpy_semantics="2.x", return_all=True)
    assert (extra["N_x_patch"], extra["N_y_patch"]) == (int(g["N_x_patch"]), int(g["N_y_patch"]))
    assert np.array_equal(extra["tri_index"], g["tri_index"])
    masks = np.unpackbits(g["masks"])[:np.prod(g["masks_shape"])].reshape(g["masks_shape"]).astype(bool)
    for a, b in zip(out, (g["input_states"], g["next_state"], g["diffs"], masks, g["pos_ids"])):
        assert a.dtype == b.dtype and np.array_equal(a, b)


@pytest.mark.parametrize("kind", ["cylinder", "airfoil"])
def test_oracle_reproduces_the_reference_for_other_strides_and_pad_false(kind):
    """tests/golden/ref_strided.npz (oracle/make_golden_strided.py): the reference's unmodified data sets with stride != patch,
    pad=False and a non-square patch, over the inputs stored in ref_{kind}.npz; SHA-256 of each of the five tensors."""
    import hashlib
    g, s = _gold(f"ref_{kind}.npz"), _gold("ref_strided.npz")
    tr = {"mesh_pos": g["mesh_pos"], "cells": g["cells"], "velocity": g["velocity"], "pressure": g["pressure"]}
    from fluid_llm_b200 import synth
    probe = synth.make_trajectory(kind, 2, mesh_seed=1, field_seed=11)
    for ci, case in enumerate(s["cases"]):
        patch, stride, pad = tuple(int(v) for v in case[:2]), tuple(int(v) for v in case[2:4]), bool(case[4])
        k = f"{kind}_{ci}"
        # the reference takes N_x_patch / N_y_patch from its probe file, save_files[1] (the second seeded trajectory)
        pos1 = probe["mesh_pos"]
        if kind == "airfoil":
            _, pos1, _ = P.airfoil_crop(pos1, probe["cells"])
        nx1, ny1 = P.grid_shape(pos1[:, 0].min(), pos1[:, 0].max(), pos1[:, 1].min(), pos1[:, 1].max(), 238, "2.x")
        if pad:
            nx1, ny1 = nx1 + (-nx1) % patch[0], ny1 + (-ny1) % patch[1]
        ring = 2 if kind == "airfoil" else 0
        attrs = (P.num_patches(nx1, patch[0], stride[0]) - ring, P.num_patches(ny1, patch[1], stride[1]) - ring)
        assert attrs == tuple(s[k + "_n_patch"]), k
        out = P.ds_get(tr, 1, 3, 2, 238, patch, kind, pad=pad, numpy_semantics="2.x", stride=stride, n_patch=attrs)
        for t, shape, digest in zip(out, s[k + "_shapes"], s[k + "_sha256"]):
            assert list(t.shape) == [int(v) for v in shape[:t.ndim]], k
            assert hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest() == str(digest), k
        assert np.array_equal(out[0][0, ::7, :, ::5, ::5], s[k + "_sample"])


def test_grid2mesh_oracle_reproduces_frozen_reference():
    g = _gold("ref_grid2mesh.npz")
    vg, pg = g["velocity_grid"].astype(np.float32), g["pressure_grid"].astype(np.float32)
    vm, pm = P.grid2mesh(vg, pg, g["mesh_pos"], numpy_semantics="2.x")
    assert np.array_equal(vm.astype(np.float16), g["velocity_mesh"]) and np.array_equal(pm.astype(np.float16), g["pressure_mesh"])
    iy, ix = P.grid2mesh_index(g["mesh_pos"][0], "1.26")
    iy2, ix2 = P.grid2mesh_index(g["mesh_pos"][0], "2.x")
    assert (iy != iy2).mean() < 0.01 and (ix != ix2).mean() < 0.01      # the two NumPy semantics differ only at cell edges


# ---- live cross-check against the reference (development container only) ---------------------
needs_ref = pytest.mark.skipif(not ref_import.available(), reason="/root/reference is not present on this machine")


@needs_ref
@pytest.mark.parametrize("kind", ["cylinder", "airfoil"])
def test_oracle_vs_reference_datasets_live(kind):
    R = ref_import.modules()
    trajs = [trajectory(kind, 140, s, 10 + s) for s in (0, 1)]
    d = ref_import.write_pickles(copy.deepcopy(trajs))
    DS = R["simple_dataloader"].MGNDataset if kind == "cylinder" else R["airfoil_ds"].AirfoilDataset
    ds = DS(load_dir=d, resolution=238, patch_size=PATCH, stride=PATCH, seq_len=4, seq_interval=2, mode="valid")
    ref = ds[1]
    out = P.ds_get(trajs[1], 100, 4, 2, 238, PATCH, "airfoil" if kind == "airfoil" else "cylinder", numpy_semantics="2.x")
    for a, b in zip(ref, out):
        assert np.array_equal(a.numpy(), b)
    raw = DS(load_dir=d, resolution=238, patch_size=PATCH, stride=PATCH, seq_len=3, seq_interval=1, mode="valid", normalize=False)[0]
    out = P.ds_get(trajs[0], 100, 3, 1, 238, PATCH, "airfoil" if kind == "airfoil" else "cylinder", normalize_ds=False,
                   numpy_semantics="2.x")
    for a, b in zip(raw, out):
        assert np.array_equal(a.numpy(), b)


@needs_ref
def test_oracle_vs_reference_patch_ops_and_interpolator_live():
    import torch
    R = ref_import.modules()
    props = R["ds_props"].DSProps(15, 4, PATCH, 3)
    x = torch.randn(2, 3, 60, 3, 16, 16)
    img = R["utils_model"].patch_to_img(x, props)
    assert np.array_equal(img.numpy(), P.patch_to_img(x.numpy(), 15, 4))
    assert np.array_equal(R["utils_model"].img_to_patch(img, props).numpy(), P.img_to_patch(img.numpy(), PATCH))
    # the vendored interpolator on top of the stub Triangulation == oracle to_grid
    tr = trajectory("eagle")
    triang, tri_index, gx, gy = R["mesh_utils"].get_mesh_interpolation(tr["mesh_pos"], tr["cells"], 238)
    d_ref, m_ref = R["mesh_utils"].to_grid(tr["pressure"][2][:, 0], gx, gy, triang, tri_index)
    t2, ti2, gx2, gy2 = P.get_mesh_interpolation(tr["mesh_pos"], tr["cells"], 238, "2.x")
    d, m = P.to_grid(tr["pressure"][2][:, 0], gx2, gy2, t2, ti2)
    assert np.array_equal(ti2, tri_index) and np.array_equal(d, d_ref) and np.array_equal(m, m_ref)


def test_dynamic_oracle_on_a_repeated_static_mesh_equals_ds_get():
    """dynamic_ds_get (per-frame meshes) degenerates to ds_get when every frame carries the same mesh."""
    from fluid_llm_b200 import synth
    tr = synth.make_trajectory("cylinder", 5, mesh_seed=1, field_seed=2)
    T = 5
    dyn = P.dynamic_ds_get(np.repeat(tr["mesh_pos"][None], T, 0), np.repeat(tr["cells"][None], T, 0), tr["velocity"],
                           tr["pressure"], 1, 3, 1)
    _, extra = P.ds_get(tr, 1, 3, 1, return_all=True)
    assert np.array_equal(dyn["states"], extra["states"]) and np.array_equal(dyn["masks"], extra["masks"])
    assert np.array_equal(dyn["tri_index"][0], extra["tri_index"]) and np.array_equal(dyn["tri_index"][2], extra["tri_index"])


def test_dynamic_synth_frames_are_valid_triangulations():
    """Every frame of the synthetic dynamic trajectory: same bounding box, triangles cover the same area as frame 0's
    node set allows (edge flips keep the area), trapezoid map and rule-based locator agree."""
    from fluid_llm_b200 import synth
    from oracle import mpl_tri
    tr = synth.make_dynamic_trajectory("cylinder", 4, mesh_seed=2, field_seed=1, flip_frac=0.2)
    lo, hi = tr["mesh_pos"].min(axis=1), tr["mesh_pos"].max(axis=1)
    assert (lo == lo[0]).all() and (hi == hi[0]).all()
    assert (tr["cells"][1] != tr["cells"][0]).any()
    gx, gy = P.grid_pos(lo[0, 0], hi[0, 0], lo[0, 1], hi[0, 1], 238)
    for t in range(4):
        pos = tr["mesh_pos"][t]
        triang = mpl_tri.Triangulation(pos[:, 0], pos[:, 1], triangles=tr["cells"][t])
        assert np.array_equal(triang.get_trifinder()(gx, gy), mpl_tri.rule_find_many(triang, gx, gy))


def test_exact_rational_locator_agrees_with_both_fp64_locators():
    """Third witness for the matplotlib layer (restated, never executed): the stated tie-break rule in exact rational
    arithmetic == the trapezoid map == the brute-force fp64 rule, on the hand-built tie mesh (every grid point on a vertex,
    an edge or in a hole) and on the boundary rows / columns of the data-set shaped meshes (grid points exactly on
    boundary edges) plus a strided interior sample."""
    from helpers import tie_mesh, trajectory
    from oracle import mpl_tri
    from oracle.exact_locator import exact_find_many
    pos, tris = tie_mesh()
    for res in (9, 5, 17, 33):
        triang, tri_o, gx, gy = P.get_mesh_interpolation(pos, tris, res, "1.26")
        e = exact_find_many(pos, tris, gx, gy)
        assert np.array_equal(e, tri_o), res
        assert np.array_equal(e, mpl_tri.rule_find_many(triang, gx, gy, bucketed=False)), res
    for kind in ("cylinder", "airfoil", "eagle"):
        tr = trajectory(kind)
        pos, faces = tr["mesh_pos"], tr["cells"]
        if kind == "airfoil":
            _, pos, faces = P.airfoil_crop(pos, faces)
        for sem in ("1.26", "2.x"):
            triang, tri_o, gx, gy = P.get_mesh_interpolation(pos, faces, 238, sem)
            sel = np.zeros(gx.shape, dtype=bool)
            sel[[0, -1], :] = True
            sel[:, [0, -1]] = True
            sel[::17, ::13] = True
            e = exact_find_many(pos, faces, gx[sel], gy[sel])
            assert np.array_equal(e, tri_o[sel]), (kind, sem)
            assert np.array_equal(e, mpl_tri.rule_find_many(triang, gx[sel], gy[sel], bucketed=True)), (kind, sem)


@pytest.mark.skipif(not ref_import.available(), reason="needs /root/reference (development container only)")
def test_reference_callers_reproduce_the_frozen_fixture_live():
    """The reference's unmodified `_generate` / `gen_seq` (src/models/model.py:154-233), over the reference's own patch ops,
    reproduce tests/golden/ref_callers.npz -- the fixture the GPU tests hold the drop-ins against is not stale."""
    import hashlib
    import torch
    from oracle import make_golden, ref_callers
    g = np.load(os.path.join(GOLDEN, "ref_callers.npz"))
    R = ref_import.modules()
    r = make_golden.ROLLOUT
    props = R["ds_props"].DSProps(Nx_patch=r["Nx"], Ny_patch=r["Ny"], patch_size=(16, 16), seq_len=r["init_len"] + r["n_steps"])
    Ref = ref_callers.reference_rollout_class(R["utils_model"].img_to_patch, R["utils_model"].patch_to_img)
    m = Ref()
    m.ds_props, m.max_ctx_len, m.forward_see_init = props, r["max_ctx_len"], ref_callers.stub_forward(props)
    init, mask, pos = make_golden.rollout_inputs()
    with torch.no_grad():
        all_states, all_diffs = m._generate(init, mask, pos, r["n_steps"])
        full = torch.cat([init, torch.zeros_like(all_states[:, r["init_len"]:])], dim=1)
        img_s, img_d = m.gen_seq((full, None, None, mask, pos), r["n_steps"], start_state=r["init_len"])
    assert np.array_equal(all_states.numpy(), g["ro_all_states"]) and np.array_equal(all_diffs.numpy(), g["ro_all_diffs"])
    sha = [hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest() for t in (img_s, img_d)]
    assert sha == list(g["ro_img_sha256"])


def test_img_eagle_oracle_reproduces_frozen_reference():
    """oracle.pipeline.img_eagle_* == the reference's EagleDataset over pre-gridded states (fixture made by
    oracle/make_golden_img_eagle.py from the unmodified eagle/Dataloader/IMG_Eagle.py)."""
    from oracle.make_golden_img_eagle import WINDOW, img_eagle_inputs
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_img_eagle.npz"))
    states, pixel_type = img_eagle_inputs()
    got, mask = P.img_eagle_item(states, pixel_type, WINDOW, "test")
    assert np.array_equal(got.view(np.int32), g["states"].view(np.int32)) and np.array_equal(mask, g["mask"])
    assert np.array_equal(P.img_eagle_denormalize(got).view(np.int32), g["denormalized"].view(np.int32))
