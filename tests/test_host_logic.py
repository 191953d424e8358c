"""CPU: host-side logic of the product (grid axes, crops, patch geometry, position ids, sharding)
against the oracle's restatement of the reference."""
import numpy as np
import pytest

from oracle import pipeline as P

from helpers import trajectory


@pytest.mark.parametrize("sem", ["1.26", "2.x"])
@pytest.mark.parametrize("box", [(0, 1.6, 0, 0.41), (-0.49, 1.99, -0.74, 0.73), (-2.5, 2.5, -1.7, 1.5), (0, 1, 0, 3.3)])
def test_grid_pos_matches_oracle(sem, box):
    from fluid_llm_b200.mesh_utils import grid_pos
    b = [np.float32(v) for v in box]
    gx, gy = grid_pos(b[0], b[1], b[2], b[3], 238, sem)
    ox, oy = P.grid_pos(b[0], b[1], b[2], b[3], 238, sem)
    assert gx.dtype == np.float32 and np.array_equal(gx, ox) and np.array_equal(gy, oy)


def test_default_semantics_is_the_pinned_numpy(monkeypatch):
    from fluid_llm_b200 import mesh_utils
    monkeypatch.delenv("FLUIDGRID_NUMPY_SEMANTICS", raising=False)
    assert mesh_utils.default_numpy_semantics() == "1.26"
    monkeypatch.setenv("FLUIDGRID_NUMPY_SEMANTICS", "2.x")
    assert mesh_utils.default_numpy_semantics() == "2.x"


def test_airfoil_crop_matches_oracle():
    from fluid_llm_b200.airfoil_ds import crop_airfoil_mesh
    tr = trajectory("airfoil")
    m, pos, faces = crop_airfoil_mesh(tr["mesh_pos"], tr["cells"])
    mo, poso, faceso = P.airfoil_crop(tr["mesh_pos"], tr["cells"])
    assert np.array_equal(m, mo) and np.array_equal(pos, poso) and np.array_equal(faces, faceso)
    assert faces.max() < len(pos) and 0 < len(pos) < len(tr["mesh_pos"])


def test_position_ids_and_patch_counts():
    from fluid_llm_b200.simple_dataloader import num_patches, position_ids
    from fluid_llm_b200.ds_props import DSProps
    assert np.array_equal(position_ids(10, 15, 4).numpy(), P.get_pos_id(10, 15, 4))
    assert position_ids(4, 13, 7).dtype.is_floating_point is False and position_ids(4, 13, 7).shape == (3, 91, 3)
    assert num_patches(240, 16, 16) == 15 and num_patches(64, 16, 16) == 4
    p = DSProps(15, 4, (16, 16), 10)
    assert (p.input_tot_size, p.N_patch, p.out_patch_size, p.channel) == ((240, 64), 60, (16, 16), 3)


def test_shard_assignment_covers_everything_once():
    from fluid_llm_b200.field_path import shard_range
    for n in (1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                got += list(range(lo, hi))
            assert got == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_synthetic_meshes_are_deterministic_and_valid():
    from fluid_llm_b200 import synth
    for kind in ("cylinder", "airfoil", "eagle"):
        p1, c1 = synth.make_mesh(kind, 3)
        p2, c2 = synth.make_mesh(kind, 3)
        assert np.array_equal(p1, p2) and np.array_equal(c1, c2)
        assert p1.dtype == np.float32 and c1.dtype == np.int32 and c1.min() == 0 and c1.max() == len(p1) - 1
    with pytest.raises(ValueError):
        synth.make_mesh("torus")


def test_fgt_file_roundtrip(tmp_path):
    import pickle
    from fluid_llm_b200.traj_store import TrajectoryFile, convert_pickle, write_fgt
    tr = trajectory("airfoil", 5)
    pkl = tmp_path / "0.pkl"
    with open(pkl, "wb") as f:
        pickle.dump(tr, f)
    hdr = convert_pickle(str(pkl), str(tmp_path / "0.fgt"), airfoil_crop=True)
    tf = TrajectoryFile(str(tmp_path / "0.fgt"))
    m, pos, faces = P.airfoil_crop(tr["mesh_pos"], tr["cells"])
    assert tf.n_nodes == len(pos) and tf.n_steps == 5 and tf.prs_stride % 4 == 0 and tf.vel_stride == 2 * tf.prs_stride
    assert np.array_equal(tf.mesh_pos, pos) and np.array_equal(tf.cells, faces.astype(np.int32))
    vel, prs = tf.array("velocity"), tf.array("pressure")
    assert np.array_equal(vel[:, :2 * tf.n_nodes].reshape(5, -1, 2), tr["velocity"][:, m])
    assert np.array_equal(prs[:, :tf.n_nodes], tr["pressure"][:, m, 0])
    assert not vel[:, 2 * tf.n_nodes:].any() and not prs[:, tf.n_nodes:].any()          # pad nodes are zero
    assert all(a["offset"] % 4096 == 0 for a in hdr["arrays"].values())
    with pytest.raises(ValueError):
        write_fgt(str(tmp_path / "bad.fgt"), pos, faces, tr["velocity"], tr["pressure"])   # uncropped fields, cropped mesh
    (tmp_path / "junk.fgt").write_bytes(b"not a trajectory")
    with pytest.raises(ValueError):
        TrajectoryFile(str(tmp_path / "junk.fgt"))


def test_load_eagle_sim_reads_the_reference_layout(tmp_path):
    """max/ds_download/eagle.py:123-144 (get_data): sim.npz with pointcloud / VX / VY / PS / PG / mask + triangles.npy."""
    from fluid_llm_b200 import synth
    from fluid_llm_b200.dynamic_mesh import load_eagle_sim
    tr = synth.make_dynamic_trajectory("cylinder", 6, mesh_seed=0, field_seed=1, flip_frac=0.0)
    T, N = tr["mesh_pos"].shape[:2]
    np.savez(tmp_path / "sim.npz", pointcloud=tr["mesh_pos"], VX=tr["velocity"][..., 0], VY=tr["velocity"][..., 1],
             PS=tr["pressure"][..., 0], PG=np.zeros((T, N), np.float32), mask=np.zeros((T, N), np.int32))
    np.save(tmp_path / "triangles.npy", tr["cells"].astype(np.int64))
    pos, cells, vel, prs = load_eagle_sim(str(tmp_path), t0=1, window_length=4)
    assert pos.shape == (4, N, 2) and cells.shape == (4, tr["cells"].shape[1], 3) and cells.dtype == np.int32
    assert np.array_equal(pos, tr["mesh_pos"][1:5]) and np.array_equal(cells, tr["cells"][1:5])
    assert np.array_equal(vel, tr["velocity"][1:5]) and np.array_equal(prs, tr["pressure"][1:5, :, 0])
    pos_all, _, _, _ = load_eagle_sim(str(tmp_path))
    assert pos_all.shape[0] == T


def test_prepare_plan_validates_like_matplotlibs_triangulation():
    """`_plan_host.prepare_plan` (NumPy only; run by the ingest workers and by MeshPlan): the errors `matplotlib.tri.Triangulation`
    raises for the reference at mesh_utils.py:103, the grid of mesh_utils.py:64-79, and slots that are a permutation."""
    from fluid_llm_b200._plan_host import prepare_plan
    from oracle import pipeline as P
    pos = np.array([[0, 0], [1, 0], [0, 1], [1, 1], [2, 0.5]], dtype=np.float32)
    tri = np.array([[0, 1, 2], [1, 3, 2], [1, 4, 3]], dtype=np.int16)
    h = prepare_plan(pos, tri, 17, "1.26")
    assert h["tri"].dtype == np.int32 and h["pos32"].dtype == np.float32 and h["n_degenerate"] == 0
    gx, gy = P.grid_pos(0.0, 2.0, 0.0, 1.0, 17, "1.26")
    assert np.array_equal(h["ax"], gx[:, 0]) and np.array_equal(h["ay"], gy[0, :])
    assert h["slots"].shape == (8,) and sorted(h["slots"][:5]) == [0, 1, 2, 3, 4] and list(h["slots"][5:]) == [5, 6, 7]
    with pytest.raises(ValueError, match=r"triangles must be a \(N, 3\) int array"):
        prepare_plan(pos, tri[:, :2], 17)
    with pytest.raises(ValueError, match=r"with N >= 1"):
        prepare_plan(pos, np.zeros((0, 3), dtype=np.int32), 17)
    with pytest.raises(ValueError, match="0 <= i < 5 but found value 5"):
        prepare_plan(pos, np.array([[0, 1, 5]]), 17)
    with pytest.raises(ValueError, match="but found value -1"):
        prepare_plan(pos, np.array([[0, 1, -1]]), 17)
    with pytest.raises(ValueError, match="equal-length 1D arrays"):
        prepare_plan(pos[:, :1], tri, 17)
    with pytest.raises(ValueError, match="zero area"):
        prepare_plan(np.array([[0, 0], [1, 0], [2, 0], [0, 1]], dtype=np.float32), np.array([[0, 1, 2], [0, 1, 3]]), 17)
    assert prepare_plan(np.array([[0, 0], [1, 0], [2, 0], [0, 1]], dtype=np.float32), np.array([[0, 1, 2], [0, 1, 3]]), 17,
                        allow_degenerate=True)["n_degenerate"] == 1
