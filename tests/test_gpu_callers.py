"""-m gpu: the drop-ins under the reference's own CALLERS.

tests/golden/ref_callers.npz was produced in the development container by the reference's UNMODIFIED `get_data_loader`
(src/utils_model.py:9-45) and `_generate` / `gen_seq` (src/models/model.py:154-233) -- see oracle/ref_callers.py and
oracle/make_golden.py.  Here the same calls go through fluid_llm_b200 (`utils_model.get_data_loader`, `model_glue.RolloutGlue`)
on the GPU and must reproduce those outputs bit for bit."""
import hashlib
import os
import pickle

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes()).hexdigest()


def test_get_data_loader_reproduces_the_reference_batch(tmp_path, monkeypatch):
    from fluid_llm_b200.utils_model import get_data_loader
    from oracle.make_golden import CALLER_CFG
    g = np.load(os.path.join(GOLDEN, "ref_callers.npz"))
    d = tmp_path / "cylinder" / "valid"
    d.mkdir(parents=True)
    for i in range(2):
        pos, T = g[f"dl_pos{i}"], 110
        vel = np.zeros((T, len(pos), 2), np.float32)
        prs = np.zeros((T, len(pos), 1), np.float32)
        vel[100:105], prs[100:105] = g[f"dl_vel{i}"], g[f"dl_prs{i}"]         # valid mode reads steps 100, 102, 104
        with open(d / f"{i}.pkl", "wb") as f:
            pickle.dump({"mesh_pos": pos, "cells": g[f"dl_cells{i}"], "velocity": vel, "pressure": prs}, f)
    monkeypatch.chdir(tmp_path)
    dl, props = get_data_loader(dict(CALLER_CFG), mode="valid", numpy_semantics="2.x")     # the fixture ran under NumPy 2.x
    assert [props.Nx_patch, props.Ny_patch, props.seq_len, props.N_patch, props.channel] == list(g["dl_props"])
    batch = next(iter(dl))
    assert len(batch) == 5 and all(t.is_cuda for t in batch)
    for t, shape, sha in zip(batch, g["dl_shapes"], g["dl_sha256"]):
        assert list(t.shape) == [int(x) for x in shape[: t.dim()]]
        assert _sha(t) == str(sha)
    assert np.array_equal(batch[0][:, :, :2, :, :2, :2].cpu().numpy(), g["dl_sample"])


def test_rollout_glue_reproduces_the_reference_generate():
    from fluid_llm_b200.ds_props import DSProps
    from fluid_llm_b200.model_glue import RolloutGlue
    from oracle.make_golden import ROLLOUT as r, rollout_inputs
    from oracle.ref_callers import stub_forward
    g = np.load(os.path.join(GOLDEN, "ref_callers.npz"))
    props = DSProps(r["Nx"], r["Ny"], (16, 16), r["init_len"] + r["n_steps"])
    init, mask, pos = (t.cuda() for t in rollout_inputs())
    glue = RolloutGlue(stub_forward(props), props, r["max_ctx_len"])
    all_states, all_diffs = glue._generate(init, mask, pos, r["n_steps"])
    assert np.array_equal(all_states.cpu().numpy(), g["ro_all_states"])
    assert np.array_equal(all_diffs.cpu().numpy(), g["ro_all_diffs"])
    full = torch.cat([init, torch.zeros_like(all_states[:, r["init_len"]:])], dim=1)
    img_s, img_d = glue.gen_seq((full, None, None, mask, pos), r["n_steps"], start_state=r["init_len"])
    assert [_sha(img_s), _sha(img_d)] == list(g["ro_img_sha256"])
