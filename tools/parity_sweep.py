"""Count last-bit differences between the CUDA path and the oracle over many meshes / frames.
usage: python tools/parity_sweep.py [n_meshes_per_kind] [frames_per_mesh]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from fluid_llm_b200 import synth
from fluid_llm_b200.airfoil_ds import crop_airfoil_mesh
from fluid_llm_b200.field_path import AIRFOIL, CYLINDER, DeviceTrajectory, interp_patchify
from fluid_llm_b200.mesh_utils import MeshPlan
from oracle import pipeline as P


def main():
    n_meshes = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    total = diff_vals = diff_tri = diff_mask = 0
    t0 = time.time()
    for kind in ("cylinder", "airfoil", "eagle"):
        pers_name = "airfoil" if kind == "airfoil" else "cylinder"
        pers = AIRFOIL if kind == "airfoil" else CYLINDER
        for seed in range(n_meshes):
            tr = synth.make_trajectory(kind, T, mesh_seed=50 + seed, field_seed=70 + seed)
            pos, faces, vel, prs = tr["mesh_pos"], tr["cells"], tr["velocity"], tr["pressure"]
            if kind == "airfoil":
                m, pos, faces = crop_airfoil_mesh(pos, faces)
                vel, prs = vel[:, m], prs[:, m]
            plan = MeshPlan(pos, faces, 238)
            for normalize in (True, False):
                states, mask, _ = interp_patchify(DeviceTrajectory(vel, prs, plan), 0, T, 1, (16, 16), pers, normalize=normalize)
                _, extra = P.ds_get(tr, 0, T, 1, 238, (16, 16), pers_name, normalize_ds=normalize, return_all=True)
                s = states.cpu().numpy()
                total += s.size
                bad = np.argwhere(s != extra["states"])
                diff_vals += len(bad)
                for b in bad[:5]:
                    x, y = s[tuple(b)], extra["states"][tuple(b)]
                    print(f"  {kind} seed {seed} normalised={normalize} idx {tuple(b)}: cuda {x!r} oracle {y!r} rel {abs(float(x) - float(y)) / max(abs(float(y)), 1e-300):.3e}")
                diff_mask += int((mask.cpu().numpy().astype(bool) != extra["masks"].astype(bool)).sum())
            diff_tri += int((plan.tri_index != extra["tri_index"]).sum())
    print(f"{total} values over {3 * n_meshes} meshes x {T} frames x (normalised, raw): {diff_vals} value differences, "
          f"{diff_mask} mask differences, {diff_tri} triangle-id differences   ({time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
