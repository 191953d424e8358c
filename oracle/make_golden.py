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
    # tie mesh
    pos, tris = tie_mesh()
    out = {"pos": pos, "tris": tris}
    for res in (5, 9, 17, 33):
        _, ti, _, _ = P.get_mesh_interpolation(pos, tris, res)
        out[f"tri_index_{res}"] = ti
    np.savez_compressed(os.path.join(OUT, "tie_mesh.npz"), **out)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KB")


if __name__ == "__main__":
    main()
