"""ORACLE test infrastructure: freeze golden vectors under tests/golden/.

Run in the development container only (needs /root/reference):

    python -m oracle.make_golden

What is frozen, and from where:
  ref_cylinder.npz / ref_airfoil.npz   inputs (mesh, 4 frames of node fields) and the 5-tuple +
      tri_index produced by the reference's UNMODIFIED MGNDataset / AirfoilDataset imported from
      /root/reference behind the oracle/stubs modules (matplotlib -> oracle/tri_oracle.cpp).  The
      container has NumPy 2.x, so these are the reference's numerics under NEP 50 ("2.x").
  ref_grid2mesh.npz   inputs and outputs of the reference's own eagle/Dataloader/IMG_Eagle.grid2mesh
      (no stubs needed), again under NumPy 2.x.
  ref_callers.npz     the reference's own CALLERS of the path run unmodified (oracle/ref_callers.py): `get_data_loader`
      (src/utils_model.py:9-45) over two Cylinder-shaped pickles in `valid` mode -> DSProps fields and the SHA-256 of the five
      batch tensors; `_generate` / `gen_seq` (src/models/model.py:154-233) with a deterministic stand-in for the backbone ->
      all_states / all_diffs and the two images.
  tie_mesh.npz        the hand-built tie-break mesh with triangle ids from the trapezoid map at
      several resolutions (oracle output; pins the stated tie-break rule against regressions).
"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    from oracle import pipeline as P, ref_import
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import tie_mesh
    from fluid_llm_b200 import synth
    os.makedirs(OUT, exist_ok=True)
    R = ref_import.modules()
    T_KEEP, STEP, SEQ, INTERVAL = 8, 1, 3, 2
    for kind in ("cylinder", "airfoil"):
        trajs = [synth.make_trajectory(kind, 30, mesh_seed=s, field_seed=10 + s) for s in (0, 1)]
        d = ref_import.write_pickles(copy.deepcopy(trajs))
        DS = R["simple_dataloader"].MGNDataset if kind == "cylinder" else R["airfoil_ds"].AirfoilDataset
        ds = DS(load_dir=d, resolution=238, patch_size=(16, 16), stride=(16, 16), seq_len=SEQ, seq_interval=INTERVAL, mode="valid")
        ds.max_step_num = 10
        ref = [t.numpy() for t in ds.ds_get(ds.save_files[0], STEP)]
        tr = trajs[0]
        pos, faces = tr["mesh_pos"], tr["cells"]
        if kind == "airfoil":
            _, pos_c, faces_c = P.airfoil_crop(pos, faces)
        else:
            pos_c, faces_c = pos, faces
        _, tri_index, _, _ = R["mesh_utils"].get_mesh_interpolation(pos_c, faces_c, 238)
        np.savez_compressed(os.path.join(OUT, f"ref_{kind}.npz"), mesh_pos=pos, cells=faces,
                            velocity=tr["velocity"][:T_KEEP], pressure=tr["pressure"][:T_KEEP],
                            step_num=STEP, seq_len=SEQ, seq_interval=INTERVAL,
                            N_x_patch=ds.N_x_patch, N_y_patch=ds.N_y_patch, tri_index=tri_index,
                            input_states=ref[0], next_state=ref[1], diffs=ref[2], masks=np.packbits(ref[3]),
                            masks_shape=np.array(ref[3].shape), pos_ids=ref[4])
        print(kind, ds.N_x_patch, ds.N_y_patch, ref[0].shape)
    # grid2mesh: the reference function itself
    rng = np.random.default_rng(5)
    Tn, N = 3, 500
    mesh_pos = np.stack([rng.uniform(-2.45, 2.45, (Tn, N)), rng.uniform(-1.65, 1.45, (Tn, N))], axis=2).astype(np.float32)
    vg = rng.standard_normal((Tn, 128, 256, 2)).astype(np.float32)
    pg = rng.standard_normal((Tn, 128, 256, 2)).astype(np.float32)
    vm, pm = R["IMG_Eagle"].grid2mesh(vg, pg, mesh_pos)
    np.savez_compressed(os.path.join(OUT, "ref_grid2mesh.npz"), mesh_pos=mesh_pos, velocity_grid=vg.astype(np.float16),
                        pressure_grid=pg.astype(np.float16), velocity_mesh=vm.numpy().astype(np.float16),
                        pressure_mesh=pm.numpy().astype(np.float16))
    callers(R, synth)
    # tie mesh
    pos, tris = tie_mesh()
    out = {"pos": pos, "tris": tris}
    for res in (5, 9, 17, 33):
        _, ti, _, _ = P.get_mesh_interpolation(pos, tris, res)
        out[f"tri_index_{res}"] = ti
    np.savez_compressed(os.path.join(OUT, "tie_mesh.npz"), **out)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KB")


CALLER_CFG = dict(load_dir="cylinder", resolution=238, patch_size=(16, 16), stride=(16, 16), seq_len=3, seq_interval=2,
                  normalize_ds=True, batch_size=2, num_workers=1)
ROLLOUT = dict(Nx=3, Ny=2, bs=2, init_len=2, n_steps=4, max_ctx_len=3, seed=11)


def rollout_inputs():
    """Seeded inputs of the rollout fixture (regenerated identically by the tests)."""
    import torch
    r = ROLLOUT
    rng = np.random.default_rng(r["seed"])
    L, T = r["Nx"] * r["Ny"], r["init_len"] + r["n_steps"]
    init = torch.from_numpy(rng.standard_normal((r["bs"], r["init_len"], L, 3, 16, 16)).astype(np.float32))
    mask = torch.from_numpy(rng.random((r["bs"], T, L, 3, 16, 16)) < 0.2)
    pos = torch.from_numpy(rng.integers(0, 9, (r["bs"], T, L, 3)).astype(np.int64))
    pos[:, :, :, 2] = torch.arange(T).view(1, T, 1) + 5
    return init, mask, pos


def callers(R, synth):
    import hashlib
    import torch
    from oracle import ref_callers, ref_import
    out = {}
    # ---- get_data_loader over two pickles, valid mode (fixed step 100) ----
    trajs = [synth.make_trajectory("cylinder", 110, mesh_seed=s, field_seed=30 + s) for s in (0, 1)]
    root = ref_import.write_pickles(copy.deepcopy(trajs), None)
    os.makedirs(os.path.join(root, "cylinder"), exist_ok=True)
    os.rename(root, root + "_v")
    os.makedirs(os.path.join(root, "cylinder"))
    os.rename(root + "_v", os.path.join(root, "cylinder", "valid"))
    cwd = os.getcwd()
    os.chdir(root)
    try:
        dl, props = R["utils_model"].get_data_loader(dict(CALLER_CFG), mode="valid")
        batch = next(iter(dl))
    finally:
        os.chdir(cwd)
    out["dl_props"] = np.array([props.Nx_patch, props.Ny_patch, props.seq_len, props.N_patch, props.channel], dtype=np.int64)
    out["dl_shapes"] = np.array([list(t.shape) + [0] * (6 - t.dim()) for t in batch], dtype=np.int64)
    out["dl_sha256"] = np.array([hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest() for t in batch])
    out["dl_sample"] = batch[0][:, :, :2, :, :2, :2].numpy()
    for i, tr in enumerate(trajs):
        out[f"dl_pos{i}"], out[f"dl_cells{i}"] = tr["mesh_pos"], tr["cells"]
        out[f"dl_vel{i}"], out[f"dl_prs{i}"] = tr["velocity"][100:105], tr["pressure"][100:105]
    # ---- _generate / gen_seq with the reference's own patch ops ----
    r = ROLLOUT
    props = R["ds_props"].DSProps(Nx_patch=r["Nx"], Ny_patch=r["Ny"], patch_size=(16, 16), seq_len=r["init_len"] + r["n_steps"])
    Ref = ref_callers.reference_rollout_class(R["utils_model"].img_to_patch, R["utils_model"].patch_to_img)
    m = Ref()
    m.ds_props, m.max_ctx_len, m.forward_see_init = props, r["max_ctx_len"], ref_callers.stub_forward(props)
    init, mask, pos = rollout_inputs()
    with torch.no_grad():
        all_states, all_diffs = m._generate(init, mask, pos, r["n_steps"])
        full = torch.cat([init, torch.zeros_like(all_states[:, r["init_len"]:])], dim=1)
        img_s, img_d = m.gen_seq((full, None, None, mask, pos), r["n_steps"], start_state=r["init_len"])
    out["ro_all_states"], out["ro_all_diffs"] = all_states.numpy(), all_diffs.numpy()
    out["ro_img_sha256"] = np.array([hashlib.sha256(np.ascontiguousarray(t.numpy()).tobytes()).hexdigest() for t in (img_s, img_d)])
    np.savez_compressed(os.path.join(OUT, "ref_callers.npz"), **out)
    print("callers:", out["dl_props"], out["dl_shapes"][0], all_states.shape)


if __name__ == "__main__":
    main()
