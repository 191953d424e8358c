"""-m gpu: the drop-in datasets (MGNDataset / AirfoilDataset) and the frozen reference outputs."""
import copy
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import pipeline as P

from helpers import PATCH, trajectory

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _write(tmp_path, trajs):
    for i, tr in enumerate(trajs):
        with open(tmp_path / f"{i}.pkl", "wb") as f:
            pickle.dump(tr, f)
    return str(tmp_path)


@pytest.mark.parametrize("kind", ["cylinder", "airfoil"])
def test_dataset_five_tuple_matches_oracle(kind, tmp_path):
    from fluid_llm_b200.airfoil_ds import AirfoilDataset
    from fluid_llm_b200.simple_dataloader import MGNDataset
    trajs = [trajectory(kind, 140, s, 10 + s) for s in (0, 1)]
    d = _write(tmp_path, copy.deepcopy(trajs))
    DS = MGNDataset if kind == "cylinder" else AirfoilDataset
    ds = DS(load_dir=d, resolution=238, patch_size=PATCH, stride=PATCH, seq_len=4, seq_interval=2, mode="valid")
    pers = "airfoil" if kind == "airfoil" else "cylinder"
    assert len(ds) == 2 and ds.save_files == ["0.pkl", "1.pkl"]
    want, extra = P.ds_get(trajs[1], 100, 4, 2, 238, PATCH, pers, return_all=True)     # valid mode: step 100
    assert (ds.N_x_patch, ds.N_y_patch, ds.N_patch) == (extra["N_x_patch"], extra["N_y_patch"], extra["N_x_patch"] * extra["N_y_patch"])
    got = ds[1]
    assert len(got) == 5
    for a, b in zip(got, want):
        assert a.is_cuda and tuple(a.shape) == b.shape
        assert str(a.dtype).replace("torch.", "") == str(b.dtype)
        assert np.array_equal(a.cpu().numpy(), b)
    # ds_get with explicit arguments, int index, clamped step (simple_dataloader.py:177-179)
    ds.max_step_num = 120                      # the synthetic trajectories are shorter than the datasets' 600 steps
    got = ds.ds_get(0, 10 ** 6)
    want = P.ds_get(trajs[0], 120, 4, 2, 238, PATCH, pers)
    for a, b in zip(got, want):
        assert np.array_equal(a.cpu().numpy(), b)
    # ds_min_max probes file [1], step 20, raw values of the whole padded frame -- for the airfoil too, whose ring crop
    # happens later (simple_dataloader.py:45-50, airfoil_ds.py:46-50)
    raw = P.full_seq(trajs[1], 20, 1, 1, 238, PATCH, pers)[0][0, :3]
    for c in range(3):
        assert ds.ds_min_max[c][0] == raw[c].min() and ds.ds_min_max[c][1] == raw[c].max()
    # the same sample with the pool switched off: unpickled (lazily) in this process, one pinned hop
    ds.ingest_workers = 0
    ds._cache.clear()
    for a, b in zip(ds.ds_get(0, 10 ** 6), want):
        assert np.array_equal(a.cpu().numpy(), b)
    # normalize=False and host output
    ds2 = DS(load_dir=d, resolution=238, patch_size=PATCH, stride=PATCH, seq_len=3, seq_interval=1, mode="test",
             normalize=False, output_device="cpu")
    got = ds2[0]
    want = P.ds_get(trajs[0], 100, 3, 1, 238, PATCH, pers, normalize_ds=False)
    for a, b in zip(got, want):
        assert not a.is_cuda and np.array_equal(a.numpy(), b)


def test_dataset_train_mode_and_dataloader(tmp_path):
    from fluid_llm_b200.simple_dataloader import MGNDataset
    from torch.utils.data import DataLoader
    trajs = [trajectory("cylinder", 620, s, 10 + s) for s in (0, 1)]
    d = _write(tmp_path, copy.deepcopy(trajs))
    ds = MGNDataset(load_dir=d, resolution=238, patch_size=PATCH, stride=PATCH, seq_len=10, seq_interval=1, mode="train")
    assert ds.max_step_num == 590
    import random
    random.seed(3)
    step = random.randint(0, 590)
    random.seed(3)
    got = ds[0]
    want = P.ds_get(trajs[0], step, 10, 1, 238, PATCH, "cylinder")
    assert np.array_equal(got[0].cpu().numpy(), want[0])
    dl = DataLoader(ds, batch_size=2, shuffle=False, num_workers=0)
    batch = next(iter(dl))
    assert tuple(batch[0].shape) == (2, 9, 60, 3, 16, 16) and batch[3].dtype == torch.bool and batch[4].dtype == torch.int64


@pytest.mark.parametrize("kind", ["cylinder", "airfoil"])
def test_cuda_path_reproduces_frozen_reference(kind):
    """tests/golden/ref_*.npz = the reference's own datasets run under NumPy 2.x (oracle/make_golden.py);
    the CUDA path in "2.x" grid semantics must reproduce them (tri ids bit-exact, values <= 1e-6 rel)."""
    from fluid_llm_b200.airfoil_ds import crop_airfoil_mesh
    from fluid_llm_b200.field_path import AIRFOIL, CYLINDER, DeviceTrajectory, interp_patchify
    from fluid_llm_b200.mesh_utils import MeshPlan
    g = np.load(os.path.join(GOLDEN, f"ref_{kind}.npz"))
    pos, faces, vel, prs = g["mesh_pos"], g["cells"], g["velocity"], g["pressure"]
    if kind == "airfoil":
        m, pos, faces = crop_airfoil_mesh(pos, faces)
        vel, prs = vel[:, m], prs[:, m]
    plan = MeshPlan(pos, faces, 238, "2.x")
    assert np.array_equal(plan.tri_index, g["tri_index"])
    states, mask, tab = interp_patchify(DeviceTrajectory(vel, prs, plan), int(g["step_num"]), int(g["seq_len"]),
                                        int(g["seq_interval"]), PATCH, AIRFOIL if kind == "airfoil" else CYLINDER)
    assert (tab.n_bx, tab.n_by) == (int(g["N_x_patch"]), int(g["N_y_patch"]))
    s = states.cpu().numpy()
    np.testing.assert_allclose(s[:-1], g["input_states"], rtol=1e-6, atol=0)
    assert np.array_equal(s[:-1], g["input_states"]) and np.array_equal(s[1:], g["next_state"])
    masks = np.unpackbits(g["masks"])[:np.prod(g["masks_shape"])].reshape(g["masks_shape"]).astype(bool)
    assert np.array_equal(np.repeat(mask.cpu().numpy()[1:, :, None], 3, axis=2).astype(bool), masks)


@pytest.mark.parametrize("kind", ["cylinder", "airfoil"])
def test_datasets_with_other_strides_and_pad_false(kind, tmp_path):
    """tests/golden/ref_strided.npz: the reference's unmodified data sets with stride != patch_size, pad=False and a
    non-square patch (oracle/make_golden_strided.py, NumPy 2.x).  The drop-in data sets reproduce every tensor bit for bit,
    the N_x_patch / N_y_patch attributes (taken from the probe file like the reference's; for the airfoil with
    stride != patch they are NOT the number of patches unfold yields, and the position ids follow the attributes), ds_min_max,
    and agree with the oracle in the pinned NumPy 1.26 semantics; one batched call equals the per-sample calls."""
    import hashlib
    from fluid_llm_b200.airfoil_ds import AirfoilDataset
    from fluid_llm_b200.simple_dataloader import MGNDataset
    fix = np.load(os.path.join(GOLDEN, "ref_strided.npz"))
    trajs = [trajectory(kind, 30, s, 10 + s) for s in (0, 1)]
    g = np.load(os.path.join(GOLDEN, f"ref_{kind}.npz"))
    assert np.array_equal(trajs[0]["mesh_pos"], g["mesh_pos"]) and np.array_equal(trajs[0]["velocity"][:8], g["velocity"])
    d = _write(tmp_path, copy.deepcopy(trajs))
    DS = MGNDataset if kind == "cylinder" else AirfoilDataset
    for ci, case in enumerate(fix["cases"]):
        patch, stride, pad = tuple(int(v) for v in case[:2]), tuple(int(v) for v in case[2:4]), bool(case[4])
        k = f"{kind}_{ci}"
        ds = DS(load_dir=d, resolution=238, patch_size=patch, stride=stride, seq_len=3, seq_interval=2, pad=pad, mode="valid",
                numpy_semantics="2.x")
        ds.max_step_num = 10
        assert (ds.N_x_patch, ds.N_y_patch) == tuple(fix[k + "_n_patch"]), k
        assert np.array_equal(np.array(ds.ds_min_max, dtype=np.float32), fix[k + "_min_max"]), k
        got = [t.cpu().numpy() for t in ds.ds_get(ds.save_files[0], 1)]
        for t, shape, digest in zip(got, fix[k + "_shapes"], fix[k + "_sha256"]):
            assert list(t.shape) == [int(v) for v in shape[:t.ndim]], k
            assert hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest() == str(digest), k
        many = ds.ds_get_many([(0, 1), (0, 3) if not pad else (1, 3)])   # (without padding two meshes need not share a patch grid)
        for a, b in zip(many[0], got):
            assert np.array_equal(a.cpu().numpy(), b)
        # pinned NumPy 1.26 semantics: against the oracle
        ds = DS(load_dir=d, resolution=238, patch_size=patch, stride=stride, seq_len=3, seq_interval=2, pad=pad, mode="valid")
        ds.max_step_num = 10
        want = P.ds_get(trajs[1], 3, 3, 2, 238, patch, "airfoil" if kind == "airfoil" else "cylinder", pad=pad, stride=stride,
                        n_patch=(ds.N_x_patch, ds.N_y_patch))
        for a, b in zip(ds.ds_get(1, 3), want):
            assert tuple(a.shape) == b.shape and np.array_equal(a.cpu().numpy(), b), k


@pytest.mark.parametrize("patch,stride", [((10, 7), (10, 7)), ((5, 5), (3, 4)), ((32, 32), (32, 32))])
def test_datasets_with_any_patch_size(patch, stride, tmp_path):
    """Patch sizes the vectorised kernels do not take (pixel counts that are not multiples of 4 / 32) and one larger than 256 pixels."""
    from fluid_llm_b200.airfoil_ds import AirfoilDataset
    from fluid_llm_b200.simple_dataloader import MGNDataset
    for kind, DS in (("cylinder", MGNDataset), ("airfoil", AirfoilDataset)):
        trajs = [trajectory(kind, 30, s, 10 + s) for s in (0, 1)]
        (tmp_path / kind).mkdir()
        d = _write(tmp_path / kind, copy.deepcopy(trajs))
        ds = DS(load_dir=d, resolution=238, patch_size=patch, stride=stride, seq_len=3, seq_interval=2, mode="valid")
        ds.max_step_num = 10
        want = P.ds_get(trajs[1], 3, 3, 2, 238, patch, kind, stride=stride, n_patch=(ds.N_x_patch, ds.N_y_patch))
        for a, b in zip(ds.ds_get(1, 3), want):
            assert tuple(a.shape) == b.shape and np.array_equal(a.cpu().numpy(), b), (kind, patch)


def test_ingest_slots_grow_when_a_trajectory_does_not_fit(tmp_path):
    from fluid_llm_b200.simple_dataloader import MGNDataset
    trajs = [trajectory("cylinder", 140, s, 10 + s) for s in (0, 1, 2)]
    d = _write(tmp_path, copy.deepcopy(trajs))

    class Small(MGNDataset):
        ingest_slot_bytes = 1 << 20            # 140 steps x 1 855 nodes x 12 B = 3.1 MB do not fit
        ingest_workers = 2
    ds = Small(load_dir=d, resolution=238, patch_size=PATCH, stride=PATCH, seq_len=4, seq_interval=2, mode="valid")
    ds.cache_size = 0
    for i in (0, 1, 2, 0):
        ds.prefetch([(j, 100) for j in range(3)])
        want = P.ds_get(trajs[i], 100, 4, 2, 238, PATCH, "cylinder")
        for a, b in zip(ds[i], want):
            assert np.array_equal(a.cpu().numpy(), b)
    assert ds.ingest_slot_bytes > 3 << 20 and ds._ingest.slot_bytes == ds.ingest_slot_bytes and ds._ingest.grow_to == 0


def test_a_patch_larger_than_the_frame_is_refused(tmp_path):
    from fluid_llm_b200.simple_dataloader import MGNDataset
    d = _write(tmp_path, [trajectory("cylinder", 30, s, 10 + s) for s in (0, 1)])
    with pytest.raises(ValueError, match="no patches left"):
        MGNDataset(load_dir=d, resolution=238, patch_size=(16, 64), stride=(16, 64), seq_len=3, pad=False)


def test_full_size_properties_cylinder():
    """BASELINE config 1 at full size (T = 600): properties that need no oracle pass."""
    from fluid_llm_b200.field_path import CYLINDER, DeviceTrajectory, interp_patchify
    from fluid_llm_b200.mesh_utils import MeshPlan
    tr = trajectory("cylinder", 600)
    plan = MeshPlan(tr["mesh_pos"], tr["cells"], 238)
    dt = DeviceTrajectory(tr["velocity"], tr["pressure"], plan)
    raw, mask, tab = interp_patchify(dt, 0, 600, 1, PATCH, CYLINDER, normalize=False)
    assert tuple(raw.shape) == (600, 60, 3, 16, 16)
    # mask <=> tri == -1 (incl. padding) and is the same for every frame; masked pixels are exactly 0
    m0 = mask[0]
    assert bool((mask == m0).all())
    tri = tab.idx[:, 3].reshape(60, 16, 16)
    assert torch.equal(m0.bool(), tri < 0)
    assert float(raw.abs().amax(dim=(0, 2))[m0.bool()].max()) == 0.0
    # interpolation is a convex combination: values stay inside the per-frame node range
    vmin = torch.from_numpy(np.stack([tr["velocity"][..., 0].min(1), tr["velocity"][..., 1].min(1), tr["pressure"][..., 0].min(1)], 1)).cuda()
    vmax = torch.from_numpy(np.stack([tr["velocity"][..., 0].max(1), tr["velocity"][..., 1].max(1), tr["pressure"][..., 0].max(1)], 1)).cuda()
    inside = ~m0.bool()
    sel = raw.permute(0, 2, 1, 3, 4)[:, :, inside]            # (T, 3, n_inside)
    assert bool((sel >= vmin[:, :, None] - 1e-6).all()) and bool((sel <= vmax[:, :, None] + 1e-6).all())
    # linearity: interp(a*f + g) == a*interp(f) + interp(g) up to fp32 rounding, via a constant field: exact
    const = DeviceTrajectory(np.full((2, plan.n_nodes, 2), 0.75, np.float32), np.full((2, plan.n_nodes, 1), -3.5, np.float32), plan)
    c, _, _ = interp_patchify(const, 0, 2, 1, PATCH, CYLINDER, normalize=False)
    cs = c.permute(0, 2, 1, 3, 4)[:, :, inside]
    assert bool((cs[:, :2] == 0.75).all()) and bool((cs[:, 2] == -3.5).all())
    # normalised output == (raw - mean) / std computed by torch in fp32, bit for bit
    norm, _, _ = interp_patchify(dt, 0, 600, 1, PATCH, CYLINDER)
    mean = torch.tensor(CYLINDER.means, device="cuda").view(1, 1, 3, 1, 1)
    std = torch.tensor(CYLINDER.stds, device="cuda").view(1, 1, 3, 1, 1)
    assert torch.equal(norm, (raw - mean) / std)
    # staged and gather kernels agree bit for bit at full size
    norm2, _, _ = interp_patchify(dt, 0, 600, 1, PATCH, CYLINDER, force_gather=True)
    assert torch.equal(norm, norm2)


@pytest.mark.parametrize("kind", ["cylinder", "airfoil"])
def test_img_mgn_dataset_matches_oracle(kind, tmp_path):
    """The DilResNet image loader (eagle/Dataloader/IMG_MGN.py): channel-last frames, all-pixel normalise, airfoil crop."""
    from fluid_llm_b200.img_mgn import EagleDataset
    root = tmp_path / f"{kind}_dataset" / "test"
    root.mkdir(parents=True)
    trajs = [trajectory(kind, 140, s, 10 + s) for s in (0, 1)]
    _write(root, copy.deepcopy(trajs))
    ds = EagleDataset(str(tmp_path / f"{kind}_dataset"), mode="test", window_length=6)
    assert len(ds) == 2
    item = ds[1]
    want_s, want_m = P.img_mgn_item(trajs[1], 100, 6, kind)
    assert item["states"].dtype == torch.float32 and item["mask"].dtype == bool
    assert np.array_equal(item["mask"], want_m)
    np.testing.assert_allclose(item["states"].numpy(), want_s, rtol=1e-6, atol=0)
    assert np.array_equal(item["states"].numpy(), want_s)
    back = ds.denormalize(item["states"])
    assert back.shape == item["states"].shape


@pytest.mark.parametrize("kind", ["cylinder", "airfoil"])
def test_datasets_on_flat_fgt_files(kind, tmp_path):
    """The flat on-disk format (traj_store.py) feeds the same datasets: identical 5-tuples to the pickle route."""
    from fluid_llm_b200.airfoil_ds import AirfoilDataset
    from fluid_llm_b200.simple_dataloader import MGNDataset
    from fluid_llm_b200.traj_store import convert_pickle
    trajs = [trajectory(kind, 140, s, 10 + s) for s in (0, 1)]
    pk = tmp_path / "pkl"
    fg = tmp_path / "fgt"
    pk.mkdir()
    fg.mkdir()
    _write(pk, copy.deepcopy(trajs))
    for i in range(2):
        convert_pickle(str(pk / f"{i}.pkl"), str(fg / f"{i}.fgt"), airfoil_crop=(kind == "airfoil"))
    DS = MGNDataset if kind == "cylinder" else AirfoilDataset
    kw = dict(resolution=238, patch_size=PATCH, stride=PATCH, seq_len=4, seq_interval=2, mode="valid")
    a, b = DS(load_dir=str(pk), **kw), DS(load_dir=str(fg), **kw)
    b.cache_size = 1            # fewer slots than files: file 0 is not resident and goes through the window-load route
    assert b.save_files == ["0.fgt", "1.fgt"] and (a.N_x_patch, a.N_y_patch) == (b.N_x_patch, b.N_y_patch)
    for x, y in zip(a[1], b[1]):
        assert torch.equal(x, y)
    want = P.ds_get(trajs[0], 100, 4, 2, 238, PATCH, "airfoil" if kind == "airfoil" else "cylinder")
    for x, w in zip(b[0], want):
        assert np.array_equal(x.cpu().numpy(), w)
    assert len(b._plans) == 1                                   # the window route cached file 0's mesh plan, not its node fields
    for x, w in zip(b[0], want):                                # second hit: plan from the cache
        assert np.array_equal(x.cpu().numpy(), w)


def test_ds_get_many_equals_per_sample_calls(tmp_path):
    """One launch for a DataLoader batch (`__getitems__` / `ds_get_many`) returns what the per-sample calls return."""
    from fluid_llm_b200.simple_dataloader import MGNDataset
    from torch.utils.data import DataLoader
    trajs = [trajectory("cylinder", 140, s, 30 + s) for s in (0, 1, 2)]
    _write(tmp_path, copy.deepcopy(trajs))
    ds = MGNDataset(load_dir=str(tmp_path), resolution=238, patch_size=PATCH, stride=PATCH, seq_len=5, seq_interval=2, mode="valid")
    ds.max_step_num = 120
    reqs = [(0, 3), (2, 100), (1, 57), (2, 500)]                 # the last one is clamped to max_step_num
    many = ds.ds_get_many(reqs)
    for (f, st), got in zip(reqs, many):
        for x, y in zip(got, ds.ds_get(f, st)):
            assert torch.equal(x, y)
    batch = next(iter(DataLoader(ds, batch_size=3, shuffle=False)))      # goes through __getitems__ + default collate
    assert batch[0].shape == (3, 4, ds.N_patch, 3, 16, 16) and batch[3].dtype == torch.bool and batch[4].shape == (3, 4, ds.N_patch, 3)
    for b in range(3):
        for x, y in zip(batch, ds.ds_get(b, 100)):
            assert torch.equal(x[b], y)


def test_sample_assemble_equals_the_reference_tensor_ops():
    """simple_dataloader.py:93,100 in one launch: diffs = states[1:] - states[:-1], masks = mask[1:] x 3 channels as bool."""
    from fluid_llm_b200.simple_dataloader import sample_assemble
    g = torch.Generator(device="cuda").manual_seed(3)
    for B, T, L, px, py in ((1, 10, 60, 16, 16), (3, 4, 7, 16, 8), (2, 2, 5, 16, 16), (2, 1, 5, 16, 16), (2, 3, 6, 5, 5), (1, 4, 3, 10, 7)):
        states = torch.randn(B, T, L, 3, px, py, device="cuda", generator=g)
        mask = (torch.rand(B, T, L, px, py, device="cuda", generator=g) < 0.3).to(torch.uint8)
        diffs, m3 = sample_assemble(states, mask)
        assert torch.equal(diffs, states[:, 1:] - states[:, :-1])
        assert m3.dtype == torch.bool and torch.equal(m3, mask[:, 1:].unsqueeze(3).repeat(1, 1, 1, 3, 1, 1).bool())
