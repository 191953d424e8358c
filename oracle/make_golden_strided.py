"""ORACLE test infrastructure: freeze what the reference's data sets return for `stride != patch_size` and `pad=False`.

Run in the development container only (needs /root/reference):

    python -m oracle.make_golden_strided

The reference's UNMODIFIED MGNDataset / AirfoilDataset (imported from /root/reference behind oracle/stubs, NumPy 2.x numerics)
are built over the same two seeded trajectories as tests/golden/ref_{cylinder,airfoil}.npz -- whose stored inputs are therefore
the inputs of this fixture as well -- once per (stride, pad) case; `ds_get(save_files[0], 1)` is frozen as the SHA-256 of each of
the five tensors plus their shapes, the N_x_patch / N_y_patch attributes and ds_min_max.  -> tests/golden/ref_strided.npz
"""
import copy
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "ref_strided.npz")

CASES = [((16, 16), (8, 12), True), ((16, 16), (16, 16), False), ((16, 16), (20, 24), False), ((12, 8), (12, 8), True)]
STEP, SEQ, INTERVAL = 1, 3, 2


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest()


def main():
    from oracle import ref_import
    from fluid_llm_b200 import synth
    R = ref_import.modules()
    out = {"cases": np.array([[*p, *s, int(pad)] for p, s, pad in CASES], dtype=np.int64)}
    for kind in ("cylinder", "airfoil"):
        trajs = [synth.make_trajectory(kind, 30, mesh_seed=s, field_seed=10 + s) for s in (0, 1)]
        d = ref_import.write_pickles(copy.deepcopy(trajs))
        DS = R["simple_dataloader"].MGNDataset if kind == "cylinder" else R["airfoil_ds"].AirfoilDataset
        for ci, (patch, stride, pad) in enumerate(CASES):
            ds = DS(load_dir=d, resolution=238, patch_size=patch, stride=stride, seq_len=SEQ, seq_interval=INTERVAL, pad=pad,
                    mode="valid")
            ds.max_step_num = 10
            ref = [t.numpy() for t in ds.ds_get(ds.save_files[0], STEP)]
            k = f"{kind}_{ci}"
            out[k + "_sha256"] = np.array([sha(t) for t in ref])
            out[k + "_shapes"] = np.array([list(t.shape) + [0] * (5 - t.ndim) for t in ref], dtype=np.int64)
            out[k + "_n_patch"] = np.array([ds.N_x_patch, ds.N_y_patch], dtype=np.int64)
            out[k + "_min_max"] = np.array(ds.ds_min_max, dtype=np.float32)
            out[k + "_sample"] = ref[0][0, ::7, :, ::5, ::5]
            print(k, patch, stride, pad, ds.N_x_patch, ds.N_y_patch, ref[0].shape, ref[4].shape)
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT) // 1024, "KB")


if __name__ == "__main__":
    main()
