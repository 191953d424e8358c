"""-m gpu: inverse path (unpatchify / patchify, rollout step, grid -> node resample) and dataset statistics."""
import os

import numpy as np
import pytest
import torch

from oracle import pipeline as P

from helpers import PATCH, oracle_ds_get

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _props(nx=15, ny=4, seq=3):
    from fluid_llm_b200.ds_props import DSProps
    return DSProps(nx, ny, PATCH, seq)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("geom", [(15, 4), (13, 7), (1, 1)])
def test_patch_ops_match_fold_unfold(dtype, geom):
    import torch.nn.functional as F
    from fluid_llm_b200.utils_model import img_to_patch, patch_to_img
    nx, ny = geom
    props = _props(nx, ny)
    x = torch.randn(2, 3, nx * ny, 3, 16, 16, device="cuda").to(dtype)
    img = patch_to_img(x, props)
    assert img.shape == (2, 3, 3, nx * 16, ny * 16) and img.dtype == dtype
    # the reference's formulation (src/utils_model.py:86-91) on the same device
    ref = F.fold(x.float().reshape(-1, nx * ny, 768).transpose(-1, -2), output_size=(nx * 16, ny * 16), kernel_size=PATCH, stride=PATCH)
    assert torch.equal(img.float().view(-1, 3, nx * 16, ny * 16), ref)
    assert np.array_equal(img.float().cpu().numpy(), P.patch_to_img(x.float().cpu().numpy(), nx, ny))
    back = img_to_patch(img, props)
    assert torch.equal(back, x)
    with pytest.raises(ValueError):
        patch_to_img(x[:, :, :-1], props) if nx * ny > 1 else patch_to_img(x[..., :8], props)


def test_patch_ops_odd_patch_and_errors():
    from fluid_llm_b200.ds_props import DSProps
    from fluid_llm_b200.utils_model import img_to_patch, patch_to_img
    props = DSProps(3, 5, (6, 5), 2)            # row bytes not a multiple of 16 -> element kernel
    x = torch.randn(1, 2, 15, 3, 6, 5, device="cuda")
    img = patch_to_img(x, props)
    assert np.array_equal(img.cpu().numpy(), P.patch_to_img(x.cpu().numpy(), 3, 5))
    assert torch.equal(img_to_patch(img, props), x)
    with pytest.raises(Exception):
        patch_to_img(x.cpu(), props)


def test_patch_ops_are_differentiable():
    """The reference uses both ops inside the training graph (src/trainer.py:95-98)."""
    from fluid_llm_b200.utils_model import img_to_patch, patch_to_img
    props = _props()
    x = torch.randn(1, 2, 60, 3, 16, 16, device="cuda", requires_grad=True)
    w = torch.randn(1, 2, 3, 240, 64, device="cuda")
    (patch_to_img(x, props) * w).sum().backward()
    assert torch.equal(x.grad, img_to_patch(w, props))
    y = torch.randn(1, 2, 3, 240, 64, device="cuda", requires_grad=True)
    v = torch.randn(1, 2, 60, 3, 16, 16, device="cuda")
    (img_to_patch(y, props) * v).sum().backward()
    assert torch.equal(y.grad, patch_to_img(v, props))


def test_rollout_step_matches_reference_glue():
    from fluid_llm_b200.utils_model import img_to_patch, rollout_step
    props = _props()
    (inp, _, _, bc, _), _ = oracle_ds_get("cylinder", 0, 3, 1)
    last = torch.from_numpy(inp[:1]).cuda().unsqueeze(0)                # (1, 1, 60, 3, 16, 16)
    mask = torch.from_numpy(bc[:1]).cuda().unsqueeze(0)
    pred = torch.randn(1, 1, 3, 240, 64, device="cuda") * 0.05
    nxt, diffs = rollout_step(last, pred, mask, props)
    d = img_to_patch(pred, props).clone()                               # model.py:164
    d[mask] = 0.                                                        # model.py:206
    assert torch.equal(diffs, d) and torch.equal(nxt, last + d)         # model.py:210
    want_next, want_d = P.rollout_step(last.cpu().numpy(), pred.cpu().numpy(), mask.cpu().numpy(), PATCH)
    assert np.array_equal(nxt.cpu().numpy(), want_next) and np.array_equal(diffs.cpu().numpy(), want_d)
    n2, d2, tok = rollout_step(last, pred, mask, props, tokens_bf16=True)       # the same call also emits the embedding's tokens
    assert torch.equal(n2, nxt) and torch.equal(d2, diffs) and tok.dtype == torch.bfloat16 and torch.equal(tok, nxt.bfloat16())
    buf = torch.zeros(60, 768, dtype=torch.bfloat16, device="cuda")              # a caller-owned token buffer (RolloutEmbedCache.token_buffer)
    _, _, tok2 = rollout_step(last, pred, mask, props, tokens_out=buf)
    assert tok2 is buf and torch.equal(buf.view(-1), tok.view(-1))
    with pytest.raises(ValueError):
        rollout_step(last, pred[..., :32], mask, props)


@pytest.mark.parametrize("patch", [(5, 5), (10, 7), (8, 6), (16, 12)])
def test_rollout_step_and_patch_ops_for_other_patch_sizes(patch):
    """utils_model.py:77-109 / model.py:204-210 for patch sizes whose width is not a multiple of 4 (scalar kernels)."""
    from fluid_llm_b200.ds_props import DSProps
    from fluid_llm_b200.utils_model import img_to_patch, patch_to_img, rollout_step
    props = DSProps(3, 4, patch, 2)
    X, Y, L = 3 * patch[0], 4 * patch[1], 12
    g = torch.Generator(device="cuda").manual_seed(5)
    last = torch.randn(2, 1, L, 3, *patch, device="cuda", generator=g)
    mask = torch.rand(2, 1, L, 3, *patch, device="cuda", generator=g) < 0.25
    pred = torch.randn(2, 1, 3, X, Y, device="cuda", generator=g)
    nxt, diffs, tok = rollout_step(last, pred, mask, props, tokens_bf16=True)
    want_next, want_d = P.rollout_step(last.cpu().numpy(), pred.cpu().numpy(), mask.cpu().numpy(), patch)
    assert np.array_equal(nxt.cpu().numpy(), want_next) and np.array_equal(diffs.cpu().numpy(), want_d)
    assert torch.equal(tok.view(-1), nxt.bfloat16().view(-1))
    img = patch_to_img(last, props)
    assert np.array_equal(img.cpu().numpy(), P.patch_to_img(last.cpu().numpy(), 3, 4)) and torch.equal(img_to_patch(img, props), last)


def test_rollout_loop_stays_exact_over_many_steps():
    """251 rollout steps (src/inference.py:87) through the fused kernel == the reference's three-op glue."""
    from fluid_llm_b200.utils_model import img_to_patch, rollout_step
    props = _props()
    g = torch.Generator(device="cuda").manual_seed(0)
    state = torch.randn(2, 1, 60, 3, 16, 16, device="cuda", generator=g)
    ref = state.clone()
    mask = torch.rand(2, 1, 60, 3, 16, 16, device="cuda", generator=g) < 0.2
    for _ in range(251):
        pred = torch.randn(2, 1, 3, 240, 64, device="cuda", generator=g) * 0.05
        state, _ = rollout_step(state, pred, mask, props)
        d = img_to_patch(pred, props).clone()
        d[mask] = 0.
        ref = ref + d
    assert torch.equal(state, ref)


def test_grid2mesh_matches_oracle_and_frozen_reference():
    from fluid_llm_b200.img_eagle import grid2mesh
    g = np.load(os.path.join(GOLDEN, "ref_grid2mesh.npz"))
    vg, pg, mp = g["velocity_grid"].astype(np.float32), g["pressure_grid"].astype(np.float32), g["mesh_pos"]
    vm, pm = grid2mesh(vg, pg, mp)
    assert not vm.is_cuda and vm.shape == (3, 500, 2)
    want_v, want_p = P.grid2mesh(vg, pg, mp, "1.26")                 # the pinned NumPy's float32 index arithmetic
    assert np.array_equal(vm.numpy(), want_v) and np.array_equal(pm.numpy(), want_p)
    # the frozen outputs of the reference itself were produced under NumPy 2.x (float64 index arithmetic):
    # nodes whose index differs between the two semantics sit within one ulp of a cell edge
    same = (P.grid2mesh_index(mp[0], "1.26")[0] == P.grid2mesh_index(mp[0], "2.x")[0]) & \
           (P.grid2mesh_index(mp[0], "1.26")[1] == P.grid2mesh_index(mp[0], "2.x")[1])
    assert same.mean() > 0.99
    assert np.array_equal(vm.numpy()[0][same].astype(np.float16), g["velocity_mesh"][0][same])
    # device tensors in -> device tensors out; per-timestep positions; negative rows wrap like NumPy indexing
    vm2, _ = grid2mesh(torch.from_numpy(vg).cuda(), torch.from_numpy(pg).cuda(), torch.from_numpy(mp).cuda())
    assert vm2.is_cuda and torch.equal(vm2.cpu(), vm)
    low = np.array([[[0.0, -1.72], [1.0, -1.699]]], dtype=np.float32)
    v_low, _ = grid2mesh(vg[:1], pg[:1], low)
    w_low, _ = P.grid2mesh(vg[:1], pg[:1], low, "1.26")
    assert np.array_equal(v_low.numpy(), w_low)
    # a node beyond the grid: the reference's fancy indexing raises IndexError, and so does the mirror (the oracle too)
    far = np.array([[[9.0, 0.0], [1.0, 0.0]]], dtype=np.float32)
    with pytest.raises(IndexError):
        P.grid2mesh(vg[:1], pg[:1], far, "1.26")
    with pytest.raises(IndexError):
        grid2mesh(vg[:1], pg[:1], far)
    v_far, _ = grid2mesh(vg[:1], pg[:1], far, check_bounds=False)           # unchecked: clamped to the last column
    assert torch.isfinite(v_far).all()


@pytest.mark.parametrize("kind", ["cylinder", "airfoil"])
def test_ds_stats_matches_oracle(kind):
    from fluid_llm_b200.compute_ds_stats import ds_stats, get_std, mean_std, merge_stats, update_variance_batch
    (inp, nxt, diffs, bc, _), extra = oracle_ds_get(kind, 0, 8, 1, normalize=False)
    states = torch.from_numpy(extra["states"]).cuda()
    mask = torch.from_numpy(extra["masks"].astype(np.uint8)).cuda()
    agg = ds_stats(states, mask).cpu().numpy()
    want = P.ds_stats(inp, diffs, bc)
    for a, w in zip(agg, want):
        assert a[0] == w[0]
        np.testing.assert_allclose(a[1], w[1], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(a[2], w[2], rtol=1e-9)
    # host helpers keep the reference's names (max/compute_ds_stats.py:20-34)
    sel = inp[:, :, 0][~bc[:, :, 0]]
    a = update_variance_batch((0, 0.0, 0.0), sel)
    np.testing.assert_allclose(get_std(a), sel.astype(np.float64).std(), rtol=1e-12)
    # splitting the frames over two "ranks" and merging reproduces the single aggregate
    half = 4
    a1 = ds_stats(states[:half + 1], mask[:half + 1])
    a2 = ds_stats(states[half:], mask[half:])
    merged = merge_stats(torch.stack([a1, a2])).cpu().numpy()
    np.testing.assert_allclose(merged[:, 0], agg[:, 0])
    np.testing.assert_allclose(merged[:, 1], agg[:, 1], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(merged[:, 2], agg[:, 2], rtol=1e-10)
    # deterministic: same bits on every call
    assert torch.equal(ds_stats(states, mask), ds_stats(states, mask))
    means, stds = mean_std(torch.from_numpy(agg))
    assert means.shape == (6,) and (stds >= 0).all()


def test_get_nrmse_matches_oracle():
    """eagle/eagle_utils.py:89-130: true and predicted node states gridded and compared (second caller of to_grid)."""
    from fluid_llm_b200.eagle_utils import get_nrmse
    from helpers import trajectory
    tr = trajectory("cylinder", 8)
    T, N = 5, len(tr["mesh_pos"])
    true = np.concatenate([tr["velocity"][:T], tr["pressure"][:T]], axis=2)[None]            # (1, T, N, 3)
    pred = (true + np.random.default_rng(0).standard_normal(true.shape).astype(np.float32) * 0.05).astype(np.float32)
    pos = np.broadcast_to(tr["mesh_pos"], (1, T, N, 2)).copy()
    faces = np.broadcast_to(tr["cells"], (1, T) + tr["cells"].shape).copy()
    got = get_nrmse(torch.from_numpy(true), torch.from_numpy(pred), torch.from_numpy(pos), torch.from_numpy(faces))
    want = P.get_nrmse(true, pred, pos, faces)
    assert got.shape == (1, T)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-5)


def test_img_eagle_dataset_reproduces_frozen_reference(tmp_path):
    """img_eagle.EagleDataset (window of states.npy -> pinned -> device -> fl_affine_channels) == the reference's EagleDataset
    on the same files, bit for bit; normalize / denormalize on CUDA and on CPU tensors."""
    import os
    from oracle.make_golden_img_eagle import WINDOW, img_eagle_inputs, write_tree
    from fluid_llm_b200.img_eagle import EagleDataset
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_img_eagle.npz"))
    data_path = write_tree(str(tmp_path))
    ds = EagleDataset(data_path, mode="test", window_length=WINDOW, splits_dir=str(tmp_path / "Splits"))
    assert len(ds) == int(g["n"][0])
    for _ in range(2):                                  # the second call reuses the pinned staging buffer
        out = ds[0]
        assert out["states"].is_cuda and out["states"].dtype == torch.float32
        assert np.array_equal(out["states"].cpu().numpy().view(np.int32), g["states"].view(np.int32))
        assert np.array_equal(out["mask"], g["mask"]) and np.array_equal(out["example"].numpy(), g["example"])
    den = ds.denormalize(out["states"])
    assert np.array_equal(den.cpu().numpy().view(np.int32), g["denormalized"].view(np.int32))
    assert np.array_equal(ds.denormalize(out["states"].cpu()).numpy().view(np.int32), g["denormalized"].view(np.int32))
    host = EagleDataset(data_path, mode="test", window_length=WINDOW, splits_dir=str(tmp_path / "Splits"), output_device="cpu")[0]
    assert not host["states"].is_cuda and np.array_equal(host["states"].numpy().view(np.int32), g["states"].view(np.int32))
    states, pixel_type = img_eagle_inputs()
    want, _ = P.img_eagle_item(states, pixel_type, WINDOW, "test")
    assert np.array_equal(out["states"].cpu().numpy(), want)
    with pytest.raises(RuntimeError):
        ds.normalize(torch.zeros(3, 5, device="cuda"))
